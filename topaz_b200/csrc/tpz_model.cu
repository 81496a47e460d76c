// Model-level C ABI: a dense ("filled") classifier network -- ResNet8/16, conv31/63/127 + LinearClassifier -- as ONE handle.
//
// What the reference's FFI for this path would bind: create a model from its layer list and DEVICE weights,
// score images, update the weights after an optimizer epoch, destroy.  Plan building (k-block tables, channel padding,
// eval-mode BatchNorm folding, identity / projection skips as extra sources), weight repacking (ON the device: no host copy
// of any parameter) and the launch sequence live here; the Python engine (topaz_b200/engine.py) builds the same plans for
// the nn.Module drop-ins and is checked against this path bit for bit (tests/test_gpu_model_abi.py).
// Reference call sites: topaz/model/classifier.py:48-66, features/resnet.py:50-251, features/basic.py:12-111,
// extract.py:224-256 (score_images: model.eval(); model.fill(); per-image forward).
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"
#include <new>
#include <vector>

namespace {

inline int rup(int c, int m = 32) { return (c + m - 1) / m * m; }

// eval-mode BatchNorm as y = a*x + sh from [4][C] = gamma, beta, running_mean, running_var
__global__ void bn_affine_kernel(const float* __restrict__ bn, float eps, int C, float* __restrict__ a, float* __restrict__ sh) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // separately rounded operations (no FMA contraction): the same arithmetic as the Python packer's torch ops, so the folded
  // biases -- and with them every fp16 rounding downstream -- are bit-identical between the two plan builders
  const float s = __fdiv_rn(bn[c], __fsqrt_rn(__fadd_rn(bn[3 * C + c], eps)));
  a[c] = s;
  sh[c] = __fsub_rn(bn[C + c], __fmul_rn(s, bn[2 * C + c]));
}
// bias'[c] = b[c]*a[c] + sh[c] (each optional), zero in the channel padding
__global__ void bias_fold_kernel(const float* __restrict__ b, const float* __restrict__ a, const float* __restrict__ sh, int C,
                                 int Cstore, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cstore) return;
  float v = 0.f;
  if (c < C) v = __fadd_rn(__fmul_rn(b ? b[c] : 0.f, a ? a[c] : 1.f), sh ? sh[c] : 0.f);
  out[c] = v;
}
// k-blocks of one conv source: out[kb0 + tap*nchunks + chunk][co][j] = fp16(w[co][chunk*KC + j][tap] * a[co]), zero in the
// channel padding.  identity = 1: the source is an identity skip (w = I, one tap).
__global__ void pack_blocks_kernel(const float* __restrict__ w, const float* __restrict__ a, int Co, int Ci, int taps, int nchunks,
                                   int KC, int CoStore, int identity, __half* __restrict__ out, int* __restrict__ rowmax) {
  const long long total = (long long)taps * nchunks * CoStore * KC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % KC);
    long long q = i / KC;
    const int co = (int)(q % CoStore); q /= CoStore;
    const int chunk = (int)(q % nchunks);
    const int tap = (int)(q / nchunks);
    const int ci = chunk * KC + j;
    float v = 0.f;
    if (co < Co && ci < Ci) v = (identity ? (co == ci ? 1.f : 0.f) : w[((long long)co * Ci + ci) * taps + tap]) * (a ? a[co] : 1.f);
    out[i] = __float2half_rn(v);
    // range guard (ops._row_scales): track max|w| per output row (non-negative floats order like their bit patterns; NaN / inf
    // have the largest patterns, so they survive the max)
    if (rowmax && co < Co && ci < Ci && !identity) atomicMax(rowmax + co, (int)(__float_as_uint(v) & 0x7fffffffu));
  }
}
// flags[0] |= 1: a non-finite weight; flags[0] |= 2: a row that the fp16 image cannot hold (max|w| > 2^14 or in (0, 2^-10)): the
// Python packer rescales such rows (TpzTcConvArgs.oscale); this packer reports them and the caller uses those plans
__global__ void range_check_kernel(int* __restrict__ rowmax, int n, int* __restrict__ flags) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const float mx = __int_as_float(rowmax[c]);
  rowmax[c] = 0;
  if (!(mx <= 3.0e38f)) atomicOr(flags, 1);
  else if (mx > 16384.f || (mx > 0.f && mx < 9.765625e-4f)) atomicOr(flags, 2);
}
// first layer for tpz_conv_first_tc: out[kb][n][j] = fp16(w[n][kb*64 + j] * a[n]) (tap index = kb*64 + j)
__global__ void pack_first_kernel(const float* __restrict__ w, const float* __restrict__ a, int Co, int taps, int KB, int Cp,
                                  __half* __restrict__ out, int* __restrict__ rowmax) {
  const int total = KB * Cp * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % 64, n = (i / 64) % Cp, kb = i / (64 * Cp);
    const int t = kb * 64 + j;
    const float v = (n < Co && t < taps) ? w[(long long)n * taps + t] * (a ? a[n] : 1.f) : 0.f;
    out[i] = __float2half_rn(v);
    if (rowmax && n < Co && t < taps) atomicMax(rowmax + n, (int)(__float_as_uint(v) & 0x7fffffffu));
  }
}
// fallback first layer (tpz_conv_first, fp32 CUDA cores): w'[n][t] = w[n][t] * a[n]
__global__ void scale_rows_kernel(const float* __restrict__ w, const float* __restrict__ a, int Co, int taps, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Co * taps) out[i] = w[i] * (a ? a[i / taps] : 1.f);
}
__global__ void pad_copy_kernel(const float* __restrict__ src, int n, int nstore, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nstore) dst[i] = i < n ? src[i] : 0.f;
}

struct Part {                    // one input source of a tensor-core conv
  const float* w; int co, ci, k, dil; int c_store; int org; bool identity;
};
struct Step {
  TpzTcConvArgs args;            // launch-invariant fields filled at create time
  std::vector<Part> parts;
  const float* bn; float eps; const float* bias_src;
  int co, co_store, shrink;
  bool save_in, two_src, dot;
  __half* weights; float* bias;  // device, owned
  long long weight_elems;
};

}  // namespace

struct TpzModel {
  int pad = 0, k0 = 0, c0 = 0, c0_store = 0, dil0 = 1;
  float slope0 = 0.f;
  bool first_tc = false;
  __half* first_w16 = nullptr;   // [KB][Cp][64]
  float* first_w32 = nullptr;    // [Co][k*k] (fallback)
  float* first_b = nullptr;      // [Cp]
  std::vector<Step> steps;
  float* dot_w = nullptr;        // [C_last_store]
  float* scratch = nullptr;      // BN affine a | sh (2 x 256 floats) | per-row max|w| of the step being packed (256) | range flags (2)
  int c_last = 0;
  // optional per-launch timing of the LAST conv step (the dominant kernel of a dense forward): a ring of CUDA event pairs recorded
  // on the launch stream, read back by tpz_model_timing_read (bench.py's roofline figure)
  std::vector<cudaEvent_t> ev;   // 2 x kTimingRing when enabled
  long long timed = 0;
};
constexpr int kTimingRing = 64;

namespace {

#define ST(s) reinterpret_cast<cudaStream_t>(s)

int free_model(TpzModel* m) {
  if (!m) return 0;
  cudaFree(m->first_w16); cudaFree(m->first_w32); cudaFree(m->first_b); cudaFree(m->dot_w); cudaFree(m->scratch);
  for (auto& s : m->steps) { cudaFree(s.weights); cudaFree(s.bias); }
  for (cudaEvent_t e : m->ev) cudaEventDestroy(e);
  delete m;
  return 0;
}

// (re)pack every weight of the model from the DEVICE parameters named by `layers` (same shapes as at create time)
int pack_model(TpzModel* m, const TpzLayerDesc* L, int nlayers, const float* cls_w, const float* cls_b, cudaStream_t st) {
  float* a = m->scratch;
  float* sh = m->scratch + 256;
  int* rowmax = reinterpret_cast<int*>(m->scratch + 512);
  int* flags = rowmax + 256;
  TPZ_CUDA(cudaMemsetAsync(rowmax, 0, (256 + 4) * sizeof(int), st));
  auto affine = [&](const float* bn, float eps, int C) -> bool {
    if (!bn) return false;
    bn_affine_kernel<<<tpz_div_up(C, 128), 128, 0, st>>>(bn, eps, C, a, sh);
    return true;
  };
  {                               // first layer
    const TpzLayerDesc& f = L[0];
    const bool has = affine(f.bn0, f.eps0, f.cout);
    const int taps = f.k * f.k;
    if (m->first_tc) {
      const int KB = (taps + 63) / 64;
      pack_first_kernel<<<tpz_div_up(KB * m->c0_store * 64, 256), 256, 0, st>>>(f.w0, has ? a : nullptr, f.cout, taps, KB, m->c0_store, m->first_w16, rowmax);
      range_check_kernel<<<1, 256, 0, st>>>(rowmax, 256, flags);
    } else {
      scale_rows_kernel<<<tpz_div_up(f.cout * taps, 256), 256, 0, st>>>(f.w0, has ? a : nullptr, f.cout, taps, m->first_w32);
    }
    bias_fold_kernel<<<tpz_div_up(m->c0_store, 128), 128, 0, st>>>(f.b0, has ? a : nullptr, has ? sh : nullptr, f.cout, m->c0_store, m->first_b);
  }
  for (auto& s : m->steps) {
    const bool has = affine(s.bn, s.eps, s.co);
    long long off = 0;
    for (const Part& p : s.parts) {
      const int taps = p.k * p.k, nchunks = tpz_div_up(p.ci, s.args.KC);
      const long long n = (long long)taps * nchunks * s.co_store * s.args.KC;
      pack_blocks_kernel<<<(int)(tpz_div_up(n, 256) > 148 * 16 ? 148 * 16 : tpz_div_up(n, 256)), 256, 0, st>>>(
          p.w, has ? a : nullptr, p.co, p.ci, taps, nchunks, s.args.KC, s.co_store, p.identity ? 1 : 0, s.weights + off, rowmax);
      off += n;
    }
    range_check_kernel<<<1, 256, 0, st>>>(rowmax, 256, flags);
    bias_fold_kernel<<<tpz_div_up(s.co_store, 128), 128, 0, st>>>(s.bias_src, has ? a : nullptr, has ? sh : nullptr, s.co, s.co_store, s.bias);
  }
  pad_copy_kernel<<<tpz_div_up(rup(m->c_last), 128), 128, 0, st>>>(cls_w, m->c_last, rup(m->c_last), m->dot_w);
  float db = 0.f;                 // the fused dot takes its bias by value: one 4-byte read-back per (re)pack (+ the range flags)
  int hflags = 0;
  TPZ_CUDA(cudaMemcpyAsync(&db, cls_b, sizeof(float), cudaMemcpyDeviceToHost, st));
  TPZ_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  TPZ_CUDA(cudaStreamSynchronize(st));
  m->steps.back().args.dot_b = db;
  TPZ_CUDA(cudaGetLastError());
  if (hflags & 1) return tpz_fail(3, "topaz_b200: non-finite convolution weights (after BatchNorm folding)");
  if (hflags & 2) return tpz_fail(TPZ_E_WEIGHT_RANGE, "tpz_model: a weight row leaves the fp16 range (needs row-scaled plans)");
  return 0;
}

// k-block table + static argument block of one conv step
int plan_step(Step& s, float slope) {
  TpzTcConvArgs& a = s.args;
  memset(&a, 0, sizeof(a));
  bool all64 = true;
  for (const Part& p : s.parts) all64 = all64 && (p.c_store % 64 == 0);
  a.KC = all64 ? 64 : 32;
  a.nsrc = (int)s.parts.size();
  int nkb = 0, lattice = -1;
  for (int si = 0; si < a.nsrc; ++si) {
    const Part& p = s.parts[si];
    TPZ_CHECK(p.c_store % a.KC == 0, "tpz_model: source channels %d not a multiple of KC=%d", p.c_store, a.KC);
    a.src[si].C = p.c_store; a.src[si].org[0] = p.org; a.src[si].org[1] = p.org; a.src[si].org[2] = 0;
    a.src[si].kw = p.k; a.src[si].kh = p.k;
    if (p.k > 1) {
      if (lattice == -1) lattice = p.dil; else if (lattice != p.dil) lattice = 0;
    }
    const int nchunks = tpz_div_up(p.ci, a.KC);          // chunks that hold real channels (all-padding chunks are dropped)
    for (int r = 0; r < p.k; ++r)
      for (int t = 0; t < p.k; ++t)
        for (int ch = 0; ch < nchunks; ++ch) {
          TPZ_CHECK(nkb < TPZ_TC_MAX_KB, "tpz_model: conv needs more than %d k-blocks", TPZ_TC_MAX_KB);
          TcKBlock& kb = a.kb[nkb++];
          kb.dx = (int16_t)(t * p.dil); kb.dy = (int16_t)(r * p.dil); kb.dz = 0; kb.c0 = (int16_t)(ch * a.KC); kb.src = si;
        }
  }
  a.nkb = nkb;
  a.lattice = lattice == -1 ? 1 : lattice;
  a.Co = s.co_store; a.TW = 16; a.TH = 8; a.lattice_z = 1;
  a.neg_slope = slope;
  s.weight_elems = (long long)nkb * s.co_store * a.KC;
  return 0;
}

}  // namespace

extern "C" int tpz_model_create(const TpzLayerDesc* layers, int nlayers, const float* cls_w, const float* cls_b, int pad,
                                TpzModel** out, void* stream) {
  TPZ_CHECK(layers && nlayers >= 2 && out && cls_w && cls_b, "tpz_model_create: bad arguments");
  TPZ_CHECK(layers[0].kind == TPZ_LAYER_CONV && layers[0].cin == 1, "tpz_model_create: the first layer must be a Cin = 1 convolution");
  TPZ_CHECK(layers[nlayers - 1].kind == TPZ_LAYER_CONV, "tpz_model_create: the last feature layer must be a convolution (the 1x1 classifier is fused into it)");
  TpzModel* m = new (std::nothrow) TpzModel();
  TPZ_CHECK(m != nullptr, "tpz_model_create: out of memory");
  const TpzLayerDesc& f = layers[0];
  m->pad = pad; m->k0 = f.k; m->dil0 = f.dil0; m->c0 = f.cout; m->c0_store = rup(f.cout); m->slope0 = f.slope0;
  m->first_tc = f.dil0 == 1 && tpz_conv_first_tc_supported(f.k, m->c0_store);
  int rc = 0;
#define MCHK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { free_model(m); return tpz_fail(1000 + (int)e_, "tpz_model_create: %s", cudaGetErrorString(e_)); } } while (0)
  MCHK(cudaMalloc(&m->scratch, (512 + 256 + 4) * sizeof(float)));   // BN a | sh | row maxima | range flags
  MCHK(cudaMalloc(&m->first_b, m->c0_store * sizeof(float)));
  if (m->first_tc) MCHK(cudaMalloc(&m->first_w16, (size_t)((f.k * f.k + 63) / 64) * m->c0_store * 64 * sizeof(__half)));
  else MCHK(cudaMalloc(&m->first_w32, (size_t)f.cout * f.k * f.k * sizeof(float)));
  int c_store = m->c0_store;
  for (int li = 1; li < nlayers && rc == 0; ++li) {
    const TpzLayerDesc& l = layers[li];
    if (l.cout > 256 || l.cin > 32767) { rc = tpz_fail(2, "tpz_model_create: layer %d has unsupported channel counts %d -> %d", li, l.cin, l.cout); break; }
    if (l.kind == TPZ_LAYER_CONV) {
      Step s;
      s.parts.push_back(Part{l.w0, l.cout, l.cin, l.k, l.dil0, c_store, 0, false});
      s.bn = l.bn0; s.eps = l.eps0; s.bias_src = l.b0; s.co = l.cout; s.co_store = rup(l.cout);
      s.shrink = (l.k - 1) * l.dil0; s.save_in = false; s.two_src = false; s.dot = (li == nlayers - 1);
      s.weights = nullptr; s.bias = nullptr;
      rc = plan_step(s, l.slope0);
      m->steps.push_back(s);
      c_store = rup(l.cout);
    } else if (l.kind == TPZ_LAYER_RESID) {
      Step s0;                                                 // conv0: cin -> cin, 3x3, dilation dil0
      s0.parts.push_back(Part{l.w0, l.cin, l.cin, 3, l.dil0, c_store, 0, false});
      s0.bn = l.bn0; s0.eps = l.eps0; s0.bias_src = l.b0; s0.co = l.cin; s0.co_store = rup(l.cin);
      s0.shrink = 2 * l.dil0; s0.save_in = true; s0.two_src = false; s0.dot = false; s0.weights = nullptr; s0.bias = nullptr;
      rc = plan_step(s0, l.slope0);
      m->steps.push_back(s0);
      if (rc) break;
      Step s1;                                                 // conv1 (+ skip): the block input is a second source, cropped by `edge`
      const int edge = l.dil0 + l.dil1;
      s1.parts.push_back(Part{l.w1, l.cout, l.cin, 3, l.dil1, rup(l.cin), 0, false});
      if (l.proj) s1.parts.push_back(Part{l.proj, l.cout, l.cin, 1, 1, c_store, edge, false});
      else s1.parts.push_back(Part{nullptr, l.cout, l.cout, 1, 1, c_store, edge, true});     // identity skip: B = I
      s1.bn = l.bn1; s1.eps = l.eps1; s1.bias_src = l.b1; s1.co = l.cout; s1.co_store = rup(l.cout);
      s1.shrink = 2 * l.dil1; s1.save_in = false; s1.two_src = true; s1.dot = false; s1.weights = nullptr; s1.bias = nullptr;
      rc = plan_step(s1, l.slope1);
      m->steps.push_back(s1);
      c_store = rup(l.cout);
    } else {
      rc = tpz_fail(2, "tpz_model_create: unknown layer kind %d", l.kind);
    }
  }
  if (rc) { free_model(m); return rc; }
  m->c_last = layers[nlayers - 1].cout;
  for (auto& s : m->steps) {
    MCHK(cudaMalloc(&s.weights, (size_t)s.weight_elems * sizeof(__half)));
    MCHK(cudaMalloc(&s.bias, s.co_store * sizeof(float)));
    s.args.weights = reinterpret_cast<const tpz_half*>(s.weights);
    s.args.bias = s.bias;
  }
  MCHK(cudaMalloc(&m->dot_w, rup(m->c_last) * sizeof(float)));
#undef MCHK
  m->steps.back().args.dot_w = m->dot_w;
  rc = pack_model(m, layers, nlayers, cls_w, cls_b, ST(stream));
  if (rc) { free_model(m); return rc; }
  *out = m;
  return 0;
}

extern "C" int tpz_model_update_weights(TpzModel* m, const TpzLayerDesc* layers, int nlayers, const float* cls_w, const float* cls_b,
                                        void* stream) {
  TPZ_CHECK(m && layers && cls_w && cls_b, "tpz_model_update_weights: bad arguments");
  // the step list must match the architecture the handle was created with
  size_t steps = 0;
  for (int li = 1; li < nlayers; ++li) steps += layers[li].kind == TPZ_LAYER_RESID ? 2 : 1;
  TPZ_CHECK(steps == m->steps.size() && layers[0].cout == m->c0 && layers[0].k == m->k0,
            "tpz_model_update_weights: layer list does not match the model");
  // re-point the sources at the (possibly moved) parameter tensors; activation slopes are parameters too (nn.PReLU)
  m->slope0 = layers[0].slope0;
  size_t si = 0;
  for (int li = 1; li < nlayers; ++li) {
    const TpzLayerDesc& l = layers[li];
    if (l.kind == TPZ_LAYER_CONV) {
      Step& s = m->steps[si++];
      s.parts[0].w = l.w0; s.bn = l.bn0; s.eps = l.eps0; s.bias_src = l.b0; s.args.neg_slope = l.slope0;
    } else {
      Step& s0 = m->steps[si++];
      s0.parts[0].w = l.w0; s0.bn = l.bn0; s0.eps = l.eps0; s0.bias_src = l.b0; s0.args.neg_slope = l.slope0;
      Step& s1 = m->steps[si++];
      s1.parts[0].w = l.w1; s1.parts[1].w = l.proj; s1.bn = l.bn1; s1.eps = l.eps1; s1.bias_src = l.b1; s1.args.neg_slope = l.slope1;
      TPZ_CHECK((l.proj != nullptr) == !s1.parts[1].identity, "tpz_model_update_weights: projection presence changed in layer %d", li);
    }
  }
  return pack_model(m, layers, nlayers, cls_w, cls_b, ST(stream));
}

extern "C" int tpz_model_destroy(TpzModel* m) { return free_model(m); }

// activation buffers: three rotating slots (current, saved block input, next), each sized for the largest layer output
static long long slot_bytes(const TpzModel* m, int B, int H, int W) {
  long long h = H + 2 * m->pad - (m->k0 - 1) * m->dil0, w = W + 2 * m->pad - (m->k0 - 1) * m->dil0;
  long long mx = (long long)B * h * w * m->c0_store * 2;
  for (const auto& s : m->steps) {
    h -= s.shrink; w -= s.shrink;
    if (h <= 0 || w <= 0) return -1;
    if (!s.dot) { const long long b = (long long)B * h * w * s.co_store * 2; if (b > mx) mx = b; }
  }
  return (mx + 1023) / 1024 * 1024;
}

extern "C" long long tpz_workspace_bytes(const TpzModel* m, int B, int H, int W) {
  if (!m || B <= 0 || H <= 0 || W <= 0) return -1;
  const long long sb = slot_bytes(m, B, H, W);
  return sb < 0 ? -1 : 3 * sb + 1024;
}

extern "C" int tpz_resnet_dense_forward(TpzModel* m, const float* x, int B, int H, int W, float* y, void* workspace,
                                        long long workspace_bytes, void* stream) {
  TPZ_CHECK(m && x && y && workspace, "tpz_resnet_dense_forward: null argument");
  const long long sb = slot_bytes(m, B, H, W);
  TPZ_CHECK(sb > 0, "tpz_resnet_dense_forward: image %dx%d is smaller than the receptive field allows", H, W);
  TPZ_CHECK(workspace_bytes >= 3 * sb + 1024, "tpz_resnet_dense_forward: workspace of %lld bytes, need %lld (tpz_workspace_bytes)",
            workspace_bytes, 3 * sb + 1024);
  TPZ_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tpz_resnet_dense_forward: workspace must be 256-byte aligned");
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  float* range = reinterpret_cast<float*>(ws);                 // [2] + the reduction's scratch word
  unsigned* work = reinterpret_cast<unsigned*>(ws + 16);
  tpz_half* slot[3] = {reinterpret_cast<tpz_half*>(ws + 1024), reinterpret_cast<tpz_half*>(ws + 1024 + sb),
                       reinterpret_cast<tpz_half*>(ws + 1024 + 2 * sb)};
  TPZ_CUDA(cudaMemsetAsync(work, 0, sizeof(unsigned), ST(stream)));
  int rc = tpz_range_scale(x, (long long)B * H * W, range, work, stream);
  if (rc) return rc;
  int h = H + 2 * m->pad - (m->k0 - 1) * m->dil0, w = W + 2 * m->pad - (m->k0 - 1) * m->dil0;
  int cur = 0, saved = -1;
  if (m->first_tc)
    rc = tpz_conv_first_tc(x, B, H, W, reinterpret_cast<const tpz_half*>(m->first_w16), m->first_b, m->c0_store, m->k0, m->pad,
                           m->slope0, 0, slot[cur], range, stream);
  else
    rc = tpz_conv_first(x, B, 1, H, W, m->first_w32, m->first_b, m->c0, 1, m->k0, m->k0, m->dil0, m->pad, m->slope0, 1, slot[cur],
                        m->c0_store, range, 0, stream);
  if (rc) return rc;
  int c_cur = m->c0_store, c_saved = 0, h_saved = 0, w_saved = 0;
  for (auto& s : m->steps) {
    TpzTcConvArgs& a = s.args;
    const int ho = h - s.shrink, wo = w - s.shrink;
    if (s.save_in) { saved = cur; c_saved = c_cur; h_saved = h; w_saved = w; }
    a.src[0].ptr = slot[cur]; a.src[0].N = B; a.src[0].D = 1; a.src[0].H = h; a.src[0].W = w; a.src[0].ld = c_cur;
    if (s.two_src) {
      a.src[1].ptr = slot[saved]; a.src[1].N = B; a.src[1].D = 1; a.src[1].H = h_saved; a.src[1].W = w_saved; a.src[1].ld = c_saved;
    }
    a.N = B; a.Do = 1; a.Ho = ho; a.Wo = wo;
    a.range = range;
    int nxt = 0;
    while (nxt == cur || nxt == saved) ++nxt;
    if (s.dot) { a.out = nullptr; a.dot_out = y; }
    else { a.out = slot[nxt]; a.out_ld = s.co_store; a.out_coff = 0; a.dot_out = nullptr; }
    const bool timed = s.dot && !m->ev.empty();
    if (timed) cudaEventRecord(m->ev[2 * (m->timed % kTimingRing)], ST(stream));
    rc = tpz_tc_conv(&a, stream);
    if (rc) return rc;
    if (timed) { cudaEventRecord(m->ev[2 * (m->timed % kTimingRing) + 1], ST(stream)); ++m->timed; }
    if (!s.dot) { cur = nxt; c_cur = s.co_store; }
    if (s.two_src) saved = -1;
    h = ho; w = wo;
  }
  return 0;
}

// Test hook: copies the packed buffers of conv step `step` (0-based, in launch order) into caller-provided DEVICE buffers:
// fp16 weights [nkb][Co][KC] (weights_out, if non-NULL and weight_capacity >= the element count) and fp32 bias [Co] (bias_out, if
// non-NULL); always reports the sizes.  step = -1 names the first layer (weights = the fp16 [KB][Cp][64] block; 0 elements on the
// CUDA-core fallback).
extern "C" int tpz_model_step_buffers(const TpzModel* m, int step, void* weights_out, long long weight_capacity, float* bias_out,
                                      long long* weight_elems, int* co_store, int* kc, int* nkb, void* stream) {
  TPZ_CHECK(m && step >= -1 && step < (int)m->steps.size(), "tpz_model_step_buffers: bad step %d", step);
  const void* w; const float* b; long long n; int co, k, nb;
  if (step < 0) {
    nb = (m->k0 * m->k0 + 63) / 64;
    w = m->first_w16; n = m->first_w16 ? (long long)nb * m->c0_store * 64 : 0; b = m->first_b; co = m->c0_store; k = 64;
  } else {
    const Step& s = m->steps[step];
    w = s.weights; n = s.weight_elems; b = s.bias; co = s.co_store; k = s.args.KC; nb = s.args.nkb;
  }
  if (weight_elems) *weight_elems = n;
  if (co_store) *co_store = co;
  if (kc) *kc = k;
  if (nkb) *nkb = nb;
  if (weights_out && n > 0) {
    TPZ_CHECK(weight_capacity >= n, "tpz_model_step_buffers: weight buffer of %lld elements, need %lld", weight_capacity, n);
    TPZ_CUDA(cudaMemcpyAsync(weights_out, w, (size_t)n * sizeof(__half), cudaMemcpyDeviceToDevice, ST(stream)));
  }
  if (bias_out) TPZ_CUDA(cudaMemcpyAsync(bias_out, b, (size_t)co * sizeof(float), cudaMemcpyDeviceToDevice, ST(stream)));
  return 0;
}

// Test hook: the launch-invariant argument block of conv step `step` (pointers as last used; geometry of the last forward).
extern "C" int tpz_model_step_args(const TpzModel* m, int step, TpzTcConvArgs* out) {
  TPZ_CHECK(m && out && step >= 0 && step < (int)m->steps.size(), "tpz_model_step_args: bad step %d", step);
  *out = m->steps[step].args;
  return 0;
}

// Timing of the last conv step (+ fused classifier): enable != 0 starts recording an event pair around it at every forward (ring of
// 64), 0 stops and frees the events.  tpz_model_timing_read synchronises on the recorded pairs and returns their durations in ms
// (oldest first, at most `capacity`), then resets the counter.
extern "C" int tpz_model_timing(TpzModel* m, int enable) {
  TPZ_CHECK(m, "tpz_model_timing: null model");
  for (cudaEvent_t e : m->ev) cudaEventDestroy(e);
  m->ev.clear();
  m->timed = 0;
  if (enable) {
    m->ev.resize(2 * kTimingRing);
    for (auto& e : m->ev) TPZ_CUDA(cudaEventCreate(&e));
  }
  return 0;
}

extern "C" int tpz_model_timing_read(TpzModel* m, float* ms, int capacity, int* count) {
  TPZ_CHECK(m && ms && count, "tpz_model_timing_read: bad arguments");
  const long long n = m->timed < kTimingRing ? m->timed : kTimingRing;
  const long long first = m->timed - n;
  int out = 0;
  for (long long i = first; i < m->timed && out < capacity; ++i) {
    const int slot = (int)(i % kTimingRing);
    TPZ_CUDA(cudaEventSynchronize(m->ev[2 * slot + 1]));
    TPZ_CUDA(cudaEventElapsedTime(&ms[out], m->ev[2 * slot], m->ev[2 * slot + 1]));
    ++out;
  }
  *count = out;
  m->timed = 0;
  return 0;
}
