/* Plain-C host of the model-level ABI (include/topaz_b200.h): reads a network description + weights + an image + the expected
 * scores from a binary file, scores the image through tpz_model_create / tpz_workspace_bytes / tpz_resnet_dense_forward and
 * prints the parity metric.  No Python, no torch: libtopaz_b200.so + the CUDA runtime only.
 *   usage: score_c model.bin        exit status 0 iff max-rel and rel-L2 <= 1e-3
 * File layout (little endian, written by tests/test_gpu_model_abi.py): int32 nlayers; per layer int32 {kind, cin, cout, k,
 * dil0, dil1, has_b0, has_b1, has_proj, has_bn0, has_bn1}, float32 {slope0, slope1, eps0, eps1}, then the present arrays in the
 * order w0, b0, w1, b1, proj, bn0[4][C], bn1[4][C]; then int32 c_last, float cls_w[c_last], float cls_b, int32 pad, int32 B, H, W,
 * float x[B*H*W], float y_ref[B*H*W]. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "topaz_b200.h"

#define CK(e) do { cudaError_t r_ = (e); if (r_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(r_), __FILE__, __LINE__); return 2; } } while (0)
#define TZ(e) do { int r_ = (e); if (r_) { fprintf(stderr, "topaz_b200 error %d: %s\n", r_, tpz_last_error()); return 3; } } while (0)

static FILE* f;
static int rd_i(void) { int v; if (fread(&v, 4, 1, f) != 1) { fprintf(stderr, "short file\n"); exit(4); } return v; }
static float rd_f(void) { float v; if (fread(&v, 4, 1, f) != 1) { fprintf(stderr, "short file\n"); exit(4); } return v; }
static const float* rd_dev(size_t n) {            /* n floats from the file -> device */
  float* h = (float*)malloc(n * sizeof(float));
  float* d = NULL;
  if (fread(h, sizeof(float), n, f) != n) { fprintf(stderr, "short file\n"); exit(4); }
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(2); }
  cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice);
  free(h);
  return d;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s model.bin\n", argv[0]); return 1; }
  f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  const int nl = rd_i();
  TpzLayerDesc* L = (TpzLayerDesc*)calloc((size_t)nl, sizeof(TpzLayerDesc));
  for (int i = 0; i < nl; ++i) {
    TpzLayerDesc* l = &L[i];
    l->kind = rd_i(); l->cin = rd_i(); l->cout = rd_i(); l->k = rd_i(); l->dil0 = rd_i(); l->dil1 = rd_i();
    const int hb0 = rd_i(), hb1 = rd_i(), hp = rd_i(), hn0 = rd_i(), hn1 = rd_i();
    l->slope0 = rd_f(); l->slope1 = rd_f(); l->eps0 = rd_f(); l->eps1 = rd_f();
    const int mid = l->kind == TPZ_LAYER_RESID ? l->cin : l->cout;      /* output channels of w0 */
    l->w0 = rd_dev((size_t)mid * l->cin * l->k * l->k);
    if (hb0) l->b0 = rd_dev((size_t)mid);
    if (l->kind == TPZ_LAYER_RESID) {
      l->w1 = rd_dev((size_t)l->cout * l->cin * 9);
      if (hb1) l->b1 = rd_dev((size_t)l->cout);
      if (hp) l->proj = rd_dev((size_t)l->cout * l->cin);
    }
    if (hn0) l->bn0 = rd_dev((size_t)4 * mid);
    if (hn1) l->bn1 = rd_dev((size_t)4 * l->cout);
  }
  const int c_last = rd_i();
  const float* cls_w = rd_dev((size_t)c_last);
  const float* cls_b = rd_dev(1);
  const int pad = rd_i(), B = rd_i(), H = rd_i(), W = rd_i();
  const size_t n = (size_t)B * H * W;
  const float* x = rd_dev(n);
  float* ref = (float*)malloc(n * sizeof(float));
  if (fread(ref, sizeof(float), n, f) != n) { fprintf(stderr, "short file\n"); return 4; }
  fclose(f);

  TpzModel* model = NULL;
  TZ(tpz_model_create(L, nl, cls_w, cls_b, pad, &model, NULL));
  const long long wsb = tpz_workspace_bytes(model, B, H, W);
  if (wsb < 0) { fprintf(stderr, "image too small\n"); return 5; }
  void* ws = NULL; float* y = NULL;
  CK(cudaMalloc(&ws, (size_t)wsb));
  CK(cudaMalloc((void**)&y, n * sizeof(float)));
  TZ(tpz_resnet_dense_forward(model, x, B, H, W, y, ws, wsb, NULL));
  /* a second call after tpz_model_update_weights (same parameters) must give the same bytes */
  float* y1 = (float*)malloc(n * sizeof(float));
  CK(cudaMemcpy(y1, y, n * sizeof(float), cudaMemcpyDeviceToHost));
  TZ(tpz_model_update_weights(model, L, nl, cls_w, cls_b, NULL));
  TZ(tpz_resnet_dense_forward(model, x, B, H, W, y, ws, wsb, NULL));
  float* y2 = (float*)malloc(n * sizeof(float));
  CK(cudaMemcpy(y2, y, n * sizeof(float), cudaMemcpyDeviceToHost));
  double dmax = 0, rmax = 0, d2 = 0, r2 = 0; size_t differ = 0;
  for (size_t i = 0; i < n; ++i) {
    const double d = fabs((double)y1[i] - ref[i]);
    if (d > dmax) dmax = d;
    if (fabs((double)ref[i]) > rmax) rmax = fabs((double)ref[i]);
    d2 += d * d; r2 += (double)ref[i] * ref[i];
    differ += y1[i] != y2[i];
  }
  const double mx = dmax / (rmax > 0 ? rmax : 1), l2 = sqrt(d2 / (r2 > 0 ? r2 : 1));
  printf("score_c: %d layers, image %dx%dx%d, workspace %.1f MB: max-rel %.3e rel-L2 %.3e, repack differs in %zu values\n", nl, B, H, W,
         wsb / 1e6, mx, l2, differ);
  TZ(tpz_model_destroy(model));
  return (mx <= 1e-3 && l2 <= 1e-3 && differ == 0) ? 0 : 10;
}
