/* Plain-C host of the U-Net entry points of the model-level ABI (include/topaz_b200.h): reads a denoiser description + weights +
 * a patch + the expected output from a binary file and runs what Denoise._denoise does (topaz/denoise.py:274-296): mean / unbiased
 * std on the device, normalise, tpz_unet2d_forward / tpz_unet3d_forward with the de-normalising epilogue.  No Python, no torch.
 *   usage: denoise_c model.bin        exit status 0 iff max-rel and rel-L2 <= tol (both measured after removing the mean of the
 *                                     expected output: the denoised signal is a small variation on a large offset)
 * File layout (little endian, written by tests/test_gpu_unet_abi.py): int32 dims, depth; then depth + 2*(depth-1) + 1 convolutions
 * in the order enc1..enc{depth}, dec{depth-1}.0, dec{depth-1}.2, .., dec1.0, dec1.2, dec1.4, each int32 {cout, cin, k, has_bias}
 * + float32 w[cout*cin*k^dims] (+ b[cout]); then int32 B, D, H, W, float tol, float x[B*D*H*W], float y_ref[B*D*H*W]. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "topaz_b200.h"

#define CK(e) do { cudaError_t r_ = (e); if (r_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(r_), __FILE__, __LINE__); return 2; } } while (0)
#define TZ(e) do { int r_ = (e); if (r_) { fprintf(stderr, "topaz_b200 error %d: %s\n", r_, tpz_last_error()); return 3; } } while (0)

static FILE* f;
static int rd_i(void) { int v; if (fread(&v, 4, 1, f) != 1) { fprintf(stderr, "short file\n"); exit(4); } return v; }
static float rd_f(void) { float v; if (fread(&v, 4, 1, f) != 1) { fprintf(stderr, "short file\n"); exit(4); } return v; }
static float* rd_dev(size_t n) {                   /* n floats from the file -> device */
  float* h = (float*)malloc(n * sizeof(float));
  float* d = NULL;
  if (fread(h, sizeof(float), n, f) != n) { fprintf(stderr, "short file\n"); exit(4); }
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(2); }
  cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice);
  free(h);
  return d;
}
static TpzConvDesc rd_conv(int dims) {
  TpzConvDesc c;
  c.cout = rd_i(); c.cin = rd_i(); c.k = rd_i();
  const int has_b = rd_i();
  size_t n = (size_t)c.cout * c.cin;
  for (int i = 0; i < dims; ++i) n *= (size_t)c.k;
  c.w = rd_dev(n);
  c.b = has_b ? rd_dev((size_t)c.cout) : NULL;
  return c;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s model.bin\n", argv[0]); return 1; }
  f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  TpzUnetDesc d;
  memset(&d, 0, sizeof(d));
  d.dims = rd_i(); d.depth = rd_i(); d.slope = 0.1f;
  if (d.depth < 2 || d.depth > TPZ_UNET_MAX_DEPTH) { fprintf(stderr, "bad depth %d\n", d.depth); return 1; }
  for (int i = 0; i < d.depth; ++i) d.enc[i] = rd_conv(d.dims);
  for (int l = d.depth - 1; l >= 1; --l) { d.dec_a[l] = rd_conv(d.dims); d.dec_b[l] = rd_conv(d.dims); }
  d.last = rd_conv(d.dims);
  const int B = rd_i(), D = rd_i(), H = rd_i(), W = rd_i();
  const float tol = rd_f();
  const size_t n = (size_t)B * D * H * W;
  float* x = rd_dev(n);
  float* ref = (float*)malloc(n * sizeof(float));
  if (fread(ref, sizeof(float), n, f) != n) { fprintf(stderr, "short file\n"); return 4; }
  fclose(f);

  TpzUnet* u = NULL;
  TZ(tpz_unet_create(&d, &u, NULL));
  const long long wsb = tpz_unet_workspace_bytes(u, B, D, H, W);
  if (wsb < 0) { fprintf(stderr, "patch too small\n"); return 5; }
  void* ws = NULL; float *y = NULL, *xn = NULL, *stats = NULL; double* work4 = NULL;
  CK(cudaMalloc(&ws, (size_t)wsb));
  CK(cudaMalloc((void**)&y, n * sizeof(float)));
  CK(cudaMalloc((void**)&xn, n * sizeof(float)));
  CK(cudaMalloc((void**)&stats, 2 * sizeof(float)));
  CK(cudaMalloc((void**)&work4, 4 * sizeof(double)));
  TZ(tpz_meanstd(x, (long long)n, 1, stats, work4, NULL));              /* denoise.py:283 */
  TZ(tpz_affine(x, (long long)n, stats, 0, xn, NULL));                  /* denoise.py:284 */
  if (d.dims == 2) TZ(tpz_unet2d_forward(u, xn, B, H, W, stats, y, ws, wsb, NULL));
  else TZ(tpz_unet3d_forward(u, xn, B, D, H, W, stats, y, ws, wsb, NULL));
  float* yh = (float*)malloc(n * sizeof(float));
  CK(cudaMemcpy(yh, y, n * sizeof(float), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (size_t i = 0; i < n; ++i) mean += ref[i];
  mean /= (double)n;
  double dmax = 0, rmax = 0, d2 = 0, r2 = 0;
  for (size_t i = 0; i < n; ++i) {
    const double r = (double)ref[i] - mean, dd = fabs((double)yh[i] - ref[i]);
    if (dd > dmax) dmax = dd;
    if (fabs(r) > rmax) rmax = fabs(r);
    d2 += dd * dd; r2 += r * r;
  }
  const double mx = dmax / (rmax > 0 ? rmax : 1), l2 = sqrt(d2 / (r2 > 0 ? r2 : 1));
  printf("denoise_c: %d-D U-Net depth %d, patch %dx%dx%dx%d, %d launches, workspace %.1f MB: DC-free max-rel %.3e rel-L2 %.3e (tol %.1e)\n",
         d.dims, d.depth, B, D, H, W, tpz_unet_launch_count(u, B, D, H, W), wsb / 1e6, mx, l2, tol);
  TZ(tpz_unet_destroy(u));
  return (mx <= tol && l2 <= tol) ? 0 : 10;
}
