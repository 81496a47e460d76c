// Model-level C ABI, part 2: the U-Net denoisers -- UDenoiseNet / UDenoiseNetSmall / UDenoiseNet3D -- as ONE handle.
//
// What the reference's FFI for the denoising path would bind (tpz_unet2d_forward / tpz_unet3d_forward): create a
// model from its convolution list and fp32 OIHW / OIDHW weights, run patches through it, destroy.  Plan building (k-block
// tables, channel padding, the poly-phase plans of the fused nearest-2x up-sampling, the Cin = 1 / Cout = 1 ends), weight
// repacking and the launch sequence live here; the Python engine (topaz_b200/engine.py: _build_unet_plan / unet_forward)
// builds the same plans for the nn.Module drop-ins and tests/test_unet_abi.py holds the two bit-identical: packed bytes,
// argument blocks and -- through the launch hook, which hands every launch of this file to the CPU simulation of the kernels --
// the network output.
// Reference call sites: topaz/denoising/models.py:74-175 (UDenoiseNet), :178-244 (UDenoiseNetSmall), :452-564
// (UDenoiseNet3D); topaz/denoise.py:274-296 (Denoise._denoise: normalise, forward, de-normalise).
//
// Precision (TpzUnetDesc.precision, the engine's TPZ_PRECISION): fast = fp16 operands / fp32 accumulation everywhere; strict = every
// activation and weight carried as a (hi, lo) fp16 pair (22 significand bits), each product as hi*hi + hi*lo + lo*hi on the same
// kernels; auto = fast except the last four convolutions of a 3-D network (engine._unet_precision has the measurements behind it).
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <new>
#include <utility>
#include <vector>

namespace {

#define ST(s) reinterpret_cast<cudaStream_t>(s)

inline int rup(int c, int m = 32) { return (c + m - 1) / m * m; }
inline int tap_ld(int taps) { int ld = 32; while (ld < taps) ld *= 2; return ld; }     // engine._tap_ld
inline int floor_div(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
inline int ipow(int k, int e) { int r = 1; while (e-- > 0) r *= k; return r; }
inline uint16_t half_bits(float v) { return static_cast<__half_raw>(__float2half_rn(v)).x; }

tpz_launch_hook g_hook = nullptr;       // test hook: every launch of this file goes to it instead of the device
void* g_hook_user = nullptr;

// where the handle's packed buffers live: device memory, or (test handles, TpzUnetDesc.host_weights) plain host memory
struct Mem {
  bool host = false;
  cudaStream_t st = nullptr;     // uploads are ordered on the caller's stream (and complete before upload() returns)
  void* alloc(size_t bytes) const {
    void* p = nullptr;
    if (host) return malloc(bytes ? bytes : 1);
    return cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr;
  }
  void release(void* p) const { if (host) free(p); else cudaFree(p); }
  bool upload(void* dst, const void* src, size_t bytes) const {
    if (host) { memcpy(dst, src, bytes); return true; }
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
  }
};

// one convolution of the network on the host, in the reference's layout [co][ci][kd][kh][kw]
struct HostConv {
  std::vector<float> w, b;
  int co = 0, ci = 0, kd = 1, kh = 1, kw = 1;
  float at(int o, int c, int q, int r, int s) const { return w[((((size_t)o * ci + c) * kd + q) * kh + r) * kw + s]; }
};

// one input source of a tensor-core conv (ops.ConvPart)
struct PartSpec {
  std::vector<float> w;          // [co][ci][kd][kh][kw]
  int co = 0, ci = 0, kd = 1, kh = 1, kw = 1;
  int c_store = 0, dil = 1, org[3] = {0, 0, 0}, lat = 0, lat_z = 0;
  bool phase = true;
  bool split = false;            // the source tensor stores (hi, lo) pairs: channels [0, c_store) = hi, [c_store, 2*c_store) = lo
  float at(int o, int c, int q, int r, int s) const { return w[((((size_t)o * ci + c) * kd + q) * kh + r) * kw + s]; }
};

PartSpec part_of(const HostConv& c, int c_begin, int c_end, int c_store, const int org[3]) {
  PartSpec p;
  p.co = c.co; p.ci = c_end - c_begin; p.kd = c.kd; p.kh = c.kh; p.kw = c.kw;
  p.c_store = c_store; p.org[0] = org[0]; p.org[1] = org[1]; p.org[2] = org[2];
  p.w.resize((size_t)p.co * p.ci * p.kd * p.kh * p.kw);
  size_t i = 0;
  for (int o = 0; o < p.co; ++o)
    for (int ch = c_begin; ch < c_end; ++ch)
      for (int q = 0; q < p.kd; ++q)
        for (int r = 0; r < p.kh; ++r)
          for (int s = 0; s < p.kw; ++s) p.w[i++] = c.at(o, ch, q, r, s);
  return p;
}

struct Plan {
  TpzTcConvArgs a;               // launch-invariant fields (ops._static_tc_args)
  int co_store = 0;
  bool split_out = false;        // the output is written as (hi, lo) pairs: 2*co_store channels, lo at channel + co_store
  int out_channels() const { return split_out ? 2 * co_store : co_store; }
  long long weight_elems = 0;
  void *w_mem = nullptr, *bias_mem = nullptr, *dotw_mem = nullptr;
};

// ops.pack_tc_conv: OIHW fp32 -> [k-block][Co][KC] fp16, k-blocks ordered (source, tap, chunk), all-zero blocks dropped.  strict: the
// weights are split w = w_hi + w_lo and every (tap, chunk) becomes up to three k-blocks -- x_hi*w_hi, x_hi*w_lo (dropped when w_lo
// rounds to zero) and, for a (hi, lo) source, x_lo*w_hi reading the lo half of the source's channels
int pack(const Mem& mem, std::vector<void*>& owned, const std::vector<PartSpec>& parts, const float* bias, int co_store, float slope,
         int lattice, int phase_sel, int lattice_z, int phase_z, const float* dot_w, float dot_b, bool strict, bool split_out, Plan* out) {
  Plan& P = *out;
  memset(&P.a, 0, sizeof(P.a));
  TpzTcConvArgs& a = P.a;
  TPZ_CHECK(!parts.empty() && parts.size() <= 2, "tpz_unet: a conv takes one or two sources");
  bool all64 = true;
  for (const PartSpec& p : parts) all64 = all64 && p.c_store % 64 == 0;
  const int KC = all64 ? 64 : 32;
  const int co_real = parts[0].co;
  TPZ_CHECK(co_store % 16 == 0 && co_store >= co_real, "tpz_unet: bad stored channel count %d for %d outputs", co_store, co_real);
  // fp16 range of the weight rows (ops._row_scales): rows that need the row-scaled plans are reported, not packed
  std::vector<float> mx(co_store, 0.f);
  bool finite = true;
  for (const PartSpec& p : parts) {
    TPZ_CHECK(strict || !p.split, "tpz_unet: a (hi, lo) source needs a strict plan");
    TPZ_CHECK(p.c_store % KC == 0 && p.ci <= p.c_store && p.co == co_real, "tpz_unet: inconsistent conv source");
    const size_t per_row = (size_t)p.ci * p.kd * p.kh * p.kw;
    for (int o = 0; o < p.co; ++o)
      for (size_t i = 0; i < per_row; ++i) {
        const float v = fabsf(p.w[(size_t)o * per_row + i]);
        if (!(v <= 3.0e38f)) finite = false;
        if (v > mx[o]) mx[o] = v;
      }
  }
  if (!finite) return tpz_fail(3, "topaz_b200: non-finite convolution weights");
  for (int o = 0; o < co_store; ++o)
    if (mx[o] > 16384.f || (mx[o] > 0.f && mx[o] < 9.765625e-4f))
      return tpz_fail(TPZ_E_WEIGHT_RANGE, "tpz_unet: a weight row leaves the fp16 range (needs row-scaled plans)");
  std::vector<uint16_t> wt;
  int nkb = 0;
  for (int si = 0; si < (int)parts.size(); ++si) {
    const PartSpec& p = parts[si];
    for (int q = 0; q < p.kd; ++q)
      for (int r = 0; r < p.kh; ++r)
        for (int s = 0; s < p.kw; ++s)
          for (int c0 = 0; c0 < p.c_store; c0 += KC) {
            bool any = false;
            for (int o = 0; o < p.co && !any; ++o)
              for (int j = 0; j < KC && c0 + j < p.ci; ++j)
                if (p.at(o, c0 + j, q, r, s) != 0.f) { any = true; break; }
            if (!any) continue;
            auto add_block = [&](int c_first, int which) -> bool {      // which: 0 = fp16(w), 1 = fp16(w - fp16(w))
              if (nkb >= TPZ_TC_MAX_KB) return false;
              TcKBlock& kb = a.kb[nkb++];
              kb.dx = (int16_t)(s * p.dil); kb.dy = (int16_t)(r * p.dil); kb.dz = (int16_t)(q * p.dil); kb.c0 = (int16_t)c_first; kb.src = si;
              const size_t base = wt.size();
              wt.resize(base + (size_t)co_store * KC, 0);
              for (int o = 0; o < p.co; ++o)
                for (int j = 0; j < KC && c0 + j < p.ci; ++j) {
                  const float v = p.at(o, c0 + j, q, r, s);
                  wt[base + (size_t)o * KC + j] = which == 0 ? half_bits(v) : half_bits(v - __half2float(__float2half_rn(v)));
                }
              return true;
            };
            bool ok = add_block(c0, 0);                                   // x (or x_hi) * w_hi
            if (strict) {
              bool lo_any = false;
              for (int o = 0; o < p.co && !lo_any; ++o)
                for (int j = 0; j < KC && c0 + j < p.ci; ++j) {
                  const float v = p.at(o, c0 + j, q, r, s);
                  if (half_bits(v - __half2float(__float2half_rn(v))) & 0x7fff) { lo_any = true; break; }
                }
              if (lo_any) ok = ok && add_block(c0, 1);                    // x_hi * w_lo
              if (p.split) ok = ok && add_block(c0 + p.c_store, 0);       // x_lo * w_hi
            }
            TPZ_CHECK(ok, "tpz_unet: conv needs more than %d k-blocks", TPZ_TC_MAX_KB);
          }
  }
  if (nkb == 0) {                // degenerate all-zero conv: keep one block so the kernel has work
    memset(&a.kb[0], 0, sizeof(TcKBlock));
    nkb = 1;
    wt.assign((size_t)co_store * KC, 0);
  }
  a.nsrc = (int)parts.size();
  int lat_auto = -1;             // the one dilation shared by all multi-tap sources, 1 if none, 0 if they differ
  for (int si = 0; si < a.nsrc; ++si) {
    const PartSpec& p = parts[si];
    TpzTcSrc& s = a.src[si];
    s.C = (p.split ? 2 : 1) * p.c_store; s.org[0] = p.org[0]; s.org[1] = p.org[1]; s.org[2] = p.org[2];
    s.kw = p.kw; s.kh = p.kh; s.lat = p.lat; s.no_phase = p.phase ? 0 : 1; s.lat_z = p.lat_z;
    if (!(p.kw == 1 && p.kh == 1)) {
      if (lat_auto == -1) lat_auto = p.dil; else if (lat_auto != p.dil) lat_auto = 0;
    }
  }
  a.KC = KC; a.nkb = nkb; a.Co = co_store; a.TW = 16; a.TH = 8;
  a.lattice = lattice >= 0 ? lattice : (lat_auto == -1 ? 1 : lat_auto);
  a.phase_sel = phase_sel; a.lattice_z = lattice_z; a.phase_z = phase_z;
  a.neg_slope = slope;
  P.co_store = co_store;
  P.split_out = split_out;
  P.weight_elems = (long long)wt.size();
  std::vector<float> b(co_store, 0.f);
  if (bias) for (int o = 0; o < co_real; ++o) b[o] = bias[o];
  P.w_mem = mem.alloc(wt.size() * sizeof(uint16_t));
  P.bias_mem = mem.alloc(b.size() * sizeof(float));
  TPZ_CHECK(P.w_mem && P.bias_mem, "tpz_unet: out of memory for packed weights");
  owned.push_back(P.w_mem); owned.push_back(P.bias_mem);
  TPZ_CHECK(mem.upload(P.w_mem, wt.data(), wt.size() * sizeof(uint16_t)) && mem.upload(P.bias_mem, b.data(), b.size() * sizeof(float)),
            "tpz_unet: weight upload failed");
  a.weights = reinterpret_cast<const tpz_half*>(P.w_mem);
  a.bias = reinterpret_cast<const float*>(P.bias_mem);
  if (dot_w) {
    std::vector<float> dw(co_store, 0.f);
    for (int o = 0; o < co_real; ++o) dw[o] = dot_w[o];
    P.dotw_mem = mem.alloc(dw.size() * sizeof(float));
    TPZ_CHECK(P.dotw_mem, "tpz_unet: out of memory for packed weights");
    owned.push_back(P.dotw_mem);
    TPZ_CHECK(mem.upload(P.dotw_mem, dw.data(), dw.size() * sizeof(float)), "tpz_unet: weight upload failed");
    a.dot_b = dot_b;             // dot_w / dot_out are set per launch
  }
  return 0;
}

struct DecLevel {
  Plan a, b;
  std::vector<Plan> up2;         // one plan per output phase of the fused nearest-2x up-sampling (4 in 2-D, 8 in 3-D)
};

}  // namespace

struct TpzUnet {
  Mem mem;
  int dims = 2, depth = 0, nf = 0;
  float slope = 0.1f;
  // first layer (Cin = 1): 0 = one tcgen05 kernel with the 2x max-pool fused (2-D), 1 = im2col + tensor-core GEMM, 2 = fp32 CUDA cores
  int first_mode = 0, k1 = 0, first_ld = 0;
  void *first_w16 = nullptr, *first_b = nullptr, *first_w32 = nullptr;
  Plan first_plan;
  std::vector<Plan> enc;         // enc2 .. enc{depth}
  std::vector<DecLevel> dec;     // index l = 1 .. depth-1
  // dec1 tail
  int k_top = 0, ntap_store = 0, last_k = 0, last_c = 0, last_cstore = 0;
  bool last_simt = false, last_strict = false;
  bool first_split = false, raw_split = false;     // enc1's output / the raw-image im2col stored as (hi, lo) pairs
  float last_b = 0.f;
  void* last_w = nullptr;        // [taps][last_cstore] fp32
  Plan last_tc;
  std::vector<void*> owned;
};

namespace {

int free_unet(TpzUnet* m) {
  if (!m) return 0;
  for (void* p : m->owned) m->mem.release(p);
  delete m;
  return 0;
}

int fetch_conv(const TpzConvDesc& d, int dims, bool host, cudaStream_t st, HostConv* out, const char* what) {
  TPZ_CHECK(d.w && d.cout > 0 && d.cin > 0 && d.k > 0 && (d.k & 1), "tpz_unet_create: bad description of %s", what);
  HostConv& c = *out;
  c.co = d.cout; c.ci = d.cin; c.kd = dims == 3 ? d.k : 1; c.kh = d.k; c.kw = d.k;
  c.w.resize((size_t)c.co * c.ci * c.kd * c.kh * c.kw);
  c.b.assign(d.b ? c.co : 0, 0.f);
  if (host) {
    memcpy(c.w.data(), d.w, c.w.size() * sizeof(float));
    if (d.b) memcpy(c.b.data(), d.b, c.b.size() * sizeof(float));
  } else {
    TPZ_CUDA(cudaMemcpyAsync(c.w.data(), d.w, c.w.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (d.b) TPZ_CUDA(cudaMemcpyAsync(c.b.data(), d.b, c.b.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    TPZ_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

// engine._up2_phase_plans: conv(cat[nearest_up2(h), other]) per output phase, straight from the half-resolution tensor h: the taps
// of the k^dims kernel that alias onto the same half-res voxel are summed (5x5 -> 3x3, 3x3(x3) -> 2x2(x2) per phase)
int up2_plans(TpzUnet* m, const HostConv& ca, int up_c, const std::function<PartSpec(int, int, int)>& second, bool strict, bool split_out,
              std::vector<Plan>* out) {
  const int k = ca.kh, pad = k / 2, dims = m->dims, kz = ca.kd;
  const int co_store = rup(ca.co);
  out->clear();
  out->reserve(dims == 3 ? 8 : 4);
  for (int pz = 0; pz < (dims == 3 ? 2 : 1); ++pz) {
    std::vector<int> offz(kz, 0);
    if (dims == 3) for (int q = 0; q < kz; ++q) offz[q] = floor_div(pz + q - pad, 2);
    for (int py = 0; py < 2; ++py) {
      std::vector<int> offy(k);
      for (int r = 0; r < k; ++r) offy[r] = floor_div(py + r - pad, 2);
      for (int px = 0; px < 2; ++px) {
        std::vector<int> offx(k);
        for (int t = 0; t < k; ++t) offx[t] = floor_div(px + t - pad, 2);
        const int az0 = *std::min_element(offz.begin(), offz.end()), ay0 = *std::min_element(offy.begin(), offy.end()),
                  ax0 = *std::min_element(offx.begin(), offx.end());
        PartSpec up;
        up.co = ca.co; up.ci = up_c;
        up.kd = *std::max_element(offz.begin(), offz.end()) - az0 + 1;
        up.kh = *std::max_element(offy.begin(), offy.end()) - ay0 + 1;
        up.kw = *std::max_element(offx.begin(), offx.end()) - ax0 + 1;
        up.w.assign((size_t)up.co * up.ci * up.kd * up.kh * up.kw, 0.f);
        for (int q = 0; q < kz; ++q)               // same accumulation order as the Python packer (fp32 sums)
          for (int r = 0; r < k; ++r)
            for (int t = 0; t < k; ++t)
              for (int o = 0; o < up.co; ++o)
                for (int c = 0; c < up_c; ++c) {
                  float& dst = up.w[((((size_t)o * up.ci + c) * up.kd + (offz[q] - az0)) * up.kh + (offy[r] - ay0)) * up.kw + (offx[t] - ax0)];
                  dst = dst + ca.at(o, c, q, r, t);
                }
        up.c_store = rup(up_c); up.dil = 1; up.org[0] = ax0; up.org[1] = ay0; up.org[2] = az0; up.lat = 1; up.lat_z = 1; up.phase = false;
        up.split = strict;
        std::vector<PartSpec> parts;
        parts.push_back(std::move(up));
        parts.push_back(second(px, py, pz));
        out->emplace_back();
        int rc = pack(m->mem, m->owned, parts, ca.b.empty() ? nullptr : ca.b.data(), co_store, m->slope, 2, py * 2 + px + 1,
                      dims == 3 ? 2 : 1, pz, nullptr, 0.f, strict, split_out, &out->back());
        if (rc) return rc;
      }
    }
  }
  return 0;
}

int build_unet(TpzUnet* m, const TpzUnetDesc* d, cudaStream_t st) {
  const int dims = d->dims, depth = d->depth, ndec = depth - 1;
  const bool host = m->mem.host;
  std::vector<HostConv> enc(depth), da(depth), db(depth);
  HostConv last;
  int rc;
  for (int i = 0; i < depth; ++i) if ((rc = fetch_conv(d->enc[i], dims, host, st, &enc[i], "an encoder convolution"))) return rc;
  for (int l = 1; l <= ndec; ++l) {
    if ((rc = fetch_conv(d->dec_a[l], dims, host, st, &da[l], "a decoder convolution"))) return rc;
    if ((rc = fetch_conv(d->dec_b[l], dims, host, st, &db[l], "a decoder convolution"))) return rc;
  }
  if ((rc = fetch_conv(d->last, dims, host, st, &last, "the output convolution"))) return rc;
  TPZ_CHECK(enc[0].ci == 1, "tpz_unet_create: the first convolution must take one channel");
  TPZ_CHECK(last.co == 1 && !last.b.empty(), "tpz_unet_create: the output convolution must produce one channel and have a bias");
  const int nf = enc[0].co;
  m->nf = nf;
  for (int i = 1; i < depth; ++i) TPZ_CHECK(enc[i].ci == enc[i - 1].co, "tpz_unet_create: encoder channel counts do not chain");
  const float slope = m->slope;
  // engine._unet_precision: which layers run with split operands (S), and which tensors must therefore be stored as (hi, lo) pairs
  // (a tensor is split when one of its consumers is in S)
  const int prec = d->precision;
  auto S_enc = [&](int i) { (void)i; return prec == 2; };
  auto S_dec = [&](int l, int idx) { return prec == 2 || (prec == 1 && dims == 3 && ndec >= 2 && ((l == 2 && idx == 2) || l == 1)); };
  const bool S_last = S_dec(1, 4);
  auto split_enc = [&](int i) {                    // output of enc{i}: read by enc{i+1} and, as skip, by dec{i+1}.0
    if (i < depth) return S_enc(i + 1) || (i + 1 <= ndec && S_dec(i + 1, 0));
    return S_dec(ndec, 0);
  };
  auto split_dec = [&](int l, int idx) {           // output of dec{l}.{idx}
    if (idx == 0) return S_dec(l, 2);
    return l > 1 ? S_dec(l - 1, 0) : S_last;
  };
  m->raw_split = S_dec(1, 0);
  m->last_strict = S_last;
  auto same_org = [&](int k, int org[3]) { org[0] = -(k / 2); org[1] = -(k / 2); org[2] = dims == 3 ? -(k / 2) : 0; };
  auto bias_of = [](const HostConv& c) { return c.b.empty() ? (const float*)nullptr : c.b.data(); };
  int org[3];

  // ---- first layer ----
  const HostConv& c1 = enc[0];
  const int k1 = c1.kh;
  m->k1 = k1;
  m->first_split = split_enc(1);
  const bool first_fp32 = S_enc(1) || m->first_split;      // Cin = 1 first conv on the fp32 CUDA-core kernel, (hi, lo) output
  if (!first_fp32 && dims == 2 && tpz_conv_first_tc_supported(k1, rup(nf))) {
    m->first_mode = 0;                            // ops.pack_first_tc: [KB][Cp][64] fp16, tap t = r*k + s
    const int cp = rup(nf), KB = (k1 * k1 + 63) / 64;
    std::vector<uint16_t> wp((size_t)KB * cp * 64, 0);
    for (int o = 0; o < nf; ++o)
      for (int t = 0; t < k1 * k1; ++t) wp[((size_t)(t / 64) * cp + o) * 64 + t % 64] = half_bits(c1.w[(size_t)o * k1 * k1 + t]);
    std::vector<float> bp(cp, 0.f);
    if (!c1.b.empty()) for (int o = 0; o < nf; ++o) bp[o] = c1.b[o];
    m->first_w16 = m->mem.alloc(wp.size() * 2); m->first_b = m->mem.alloc(bp.size() * 4);
    TPZ_CHECK(m->first_w16 && m->first_b, "tpz_unet_create: out of memory");
    m->owned.push_back(m->first_w16); m->owned.push_back(m->first_b);
    TPZ_CHECK(m->mem.upload(m->first_w16, wp.data(), wp.size() * 2) && m->mem.upload(m->first_b, bp.data(), bp.size() * 4),
              "tpz_unet_create: weight upload failed");
  } else if (!first_fp32 && k1 * k1 <= 128) {
    m->first_mode = 1;                            // in-plane im2col (k*k taps -> channels) + GEMM; in 3-D the k z-taps stay taps
    const int taps = k1 * k1, ld = tap_ld(taps);
    m->first_ld = ld;
    PartSpec p;
    p.co = nf; p.ci = taps; p.kd = dims == 3 ? k1 : 1; p.kh = 1; p.kw = 1; p.c_store = ld; p.dil = 1;
    p.org[0] = 0; p.org[1] = 0; p.org[2] = dims == 3 ? -(k1 / 2) : 0;
    p.w.resize((size_t)nf * taps * p.kd);
    for (int o = 0; o < nf; ++o)
      for (int t = 0; t < taps; ++t)
        for (int q = 0; q < p.kd; ++q) p.w[((size_t)o * taps + t) * p.kd + q] = c1.w[((size_t)o * p.kd + q) * taps + t];
    if ((rc = pack(m->mem, m->owned, {p}, bias_of(c1), rup(nf), slope, -1, 0, 1, 0, nullptr, 0.f, false, false, &m->first_plan))) return rc;
  } else {
    m->first_mode = 2;                            // fp32 CUDA-core kernel on the reference's own [Co][kd][kh][kw] weights
    std::vector<float> bp(nf, 0.f);
    if (!c1.b.empty()) bp = c1.b;
    m->first_w32 = m->mem.alloc(c1.w.size() * 4); m->first_b = m->mem.alloc(bp.size() * 4);
    TPZ_CHECK(m->first_w32 && m->first_b, "tpz_unet_create: out of memory");
    m->owned.push_back(m->first_w32); m->owned.push_back(m->first_b);
    TPZ_CHECK(m->mem.upload(m->first_w32, c1.w.data(), c1.w.size() * 4) && m->mem.upload(m->first_b, bp.data(), bp.size() * 4),
              "tpz_unet_create: weight upload failed");
  }

  // ---- encoder ----
  m->enc.resize(depth - 1);
  for (int i = 1; i < depth; ++i) {
    same_org(enc[i].kh, org);
    PartSpec p = part_of(enc[i], 0, enc[i].ci, rup(enc[i].ci), org);
    p.split = S_enc(i + 1);
    if ((rc = pack(m->mem, m->owned, {p}, bias_of(enc[i]), rup(enc[i].co), slope, -1, 0, 1, 0, nullptr, 0.f, S_enc(i + 1), split_enc(i + 1),
                   &m->enc[i - 1]))) return rc;
  }

  // ---- decoder ----
  m->dec.resize(depth);
  int up_c = enc[depth - 1].co;
  for (int l = ndec; l >= 1; --l) {
    const HostConv &ca = da[l], &cb = db[l];
    DecLevel& D = m->dec[l];
    const int k = ca.kh, pad_k = k / 2;
    TPZ_CHECK(cb.ci == ca.co, "tpz_unet_create: decoder level %d channel counts do not chain", l);
    same_org(k, org);
    const bool sa = S_dec(l, 0), sb = S_dec(l, 2);
    if (l > 1) {
      const int skip_c = enc[l - 2].co;
      TPZ_CHECK(ca.ci == up_c + skip_c, "tpz_unet_create: dec%d.0 takes %d channels, expected %d + %d", l, ca.ci, up_c, skip_c);
      std::vector<PartSpec> parts{part_of(ca, 0, up_c, rup(up_c), org), part_of(ca, up_c, ca.ci, rup(skip_c), org)};
      parts[0].split = parts[1].split = sa;
      if ((rc = pack(m->mem, m->owned, parts, bias_of(ca), rup(ca.co), slope, -1, 0, 1, 0, nullptr, 0.f, sa, split_dec(l, 0), &D.a))) return rc;
      auto second = [&](int px, int py, int pz) {
        int o2[3] = {px - pad_k, py - pad_k, dims == 3 ? pz - pad_k : 0};
        PartSpec p = part_of(ca, up_c, ca.ci, rup(skip_c), o2);
        p.lat = 2; p.lat_z = dims == 3 ? 2 : 0; p.phase = false; p.split = sa;
        return p;
      };
      if ((rc = up2_plans(m, ca, up_c, second, sa, split_dec(l, 0), &D.up2))) return rc;
    } else {
      // dec1: [up-sampled (up_c channels), raw image (1 channel)]; the raw slice is a second source of k^dims im2col channels
      TPZ_CHECK(ca.ci == up_c + 1, "tpz_unet_create: dec1.0 takes %d channels, expected %d + 1", ca.ci, up_c);
      const int ntap = ipow(k, dims);
      m->k_top = k;
      m->ntap_store = tap_ld(ntap);
      auto raw_part = [&](int ox, int oy, int oz) {
        PartSpec p;
        p.co = ca.co; p.ci = ntap; p.kd = p.kh = p.kw = 1; p.c_store = tap_ld(ntap); p.dil = 1;
        p.org[0] = ox; p.org[1] = oy; p.org[2] = oz;
        p.w.resize((size_t)ca.co * ntap);
        for (int o = 0; o < ca.co; ++o)
          for (int t = 0; t < ntap; ++t) p.w[(size_t)o * ntap + t] = ca.w[((size_t)o * ca.ci + up_c) * ntap + t];
        return p;
      };
      std::vector<PartSpec> parts{part_of(ca, 0, up_c, rup(up_c), org), raw_part(0, 0, 0)};
      parts[0].split = parts[1].split = sa;
      if ((rc = pack(m->mem, m->owned, parts, bias_of(ca), rup(ca.co), slope, -1, 0, 1, 0, nullptr, 0.f, sa, split_dec(l, 0), &D.a))) return rc;
      auto second = [&](int px, int py, int pz) {
        PartSpec p = raw_part(px, py, dims == 3 ? pz : 0);
        p.lat = 2; p.lat_z = dims == 3 ? 2 : 0; p.phase = false; p.split = sa;
        return p;
      };
      if ((rc = up2_plans(m, ca, up_c, second, sa, split_dec(l, 0), &D.up2))) return rc;
    }
    int orgb[3];
    same_org(cb.kh, orgb);
    PartSpec pb = part_of(cb, 0, cb.ci, rup(ca.co), orgb);
    pb.split = sb;
    if ((rc = pack(m->mem, m->owned, {pb}, bias_of(cb), rup(cb.co), slope, -1, 0, 1, 0, nullptr, 0.f, sb, split_dec(l, 2), &D.b))) return rc;
    up_c = cb.co;
  }

  // ---- dec1.4: Cout = 1 ----
  TPZ_CHECK(last.ci == db[1].co, "tpz_unet_create: the output convolution takes %d channels, dec1.2 produces %d", last.ci, db[1].co);
  const int kl = last.kh, cin = last.ci, taps = ipow(kl, dims);
  m->last_k = kl; m->last_c = cin; m->last_cstore = rup(cin); m->last_b = last.b[0];
  {
    // [taps][C]: wl[t][c] = w[0][c][t]; a split (hi, lo) input is read as 2*C channels with the weights repeated
    const int cs = rup(cin), reps = S_last ? 2 : 1;
    std::vector<float> wl((size_t)taps * cs * reps, 0.f);
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < taps; ++t)
        for (int h = 0; h < reps; ++h) wl[((size_t)t * reps + h) * cs + c] = last.w[(size_t)c * taps + t];
    m->last_w = m->mem.alloc(wl.size() * 4);
    TPZ_CHECK(m->last_w, "tpz_unet_create: out of memory");
    m->owned.push_back(m->last_w);
    TPZ_CHECK(m->mem.upload(m->last_w, wl.data(), wl.size() * 4), "tpz_unet_create: weight upload failed");
  }
  // CUDA-core Cout = 1 tail where a tiled kernel exists (2-D 3x3 / 5x5, 3-D 3x3x3 on 32 channels: 1600 FLOP/px is too little for
  // the tensor-core path); elsewhere a 16-column tensor-core GEMM whose fused "dot" epilogue picks column 0
  const bool tiled3d = dims == 3 && rup(cin) == 32 && kl == 3;      // the z-marching kernel also takes split inputs
  m->last_simt = S_last ? tiled3d : (rup(cin) == 32 && ((dims == 2 && (kl == 3 || kl == 5)) || tiled3d));
  same_org(kl, org);
  PartSpec pl = part_of(last, 0, cin, rup(cin), org);
  pl.split = S_last;
  const float onehot0 = 1.f;
  if ((rc = pack(m->mem, m->owned, {pl}, nullptr, 16, 1.0f, -1, 0, 1, 0, &onehot0, last.b[0], S_last, false, &m->last_tc))) return rc;
  return 0;
}

// ---- workspace: first-fit allocator over byte offsets (the same schedule sizes the workspace and runs the network) ----
struct Arena {
  std::vector<std::pair<long long, long long>> holes;   // (offset, bytes), sorted by offset
  long long peak = 0;
  Arena() { holes.emplace_back(0, (long long)1 << 60); }
  long long take(long long bytes) {
    bytes = (bytes + 255) / 256 * 256;
    if (bytes == 0) bytes = 256;
    for (size_t i = 0; i < holes.size(); ++i)
      if (holes[i].second >= bytes) {
        const long long off = holes[i].first;
        holes[i].first += bytes; holes[i].second -= bytes;
        if (holes[i].second == 0) holes.erase(holes.begin() + i);
        peak = std::max(peak, off + bytes);
        return off;
      }
    return -1;
  }
  void give(long long off, long long bytes) {
    bytes = (bytes + 255) / 256 * 256;
    if (bytes == 0) bytes = 256;
    size_t i = 0;
    while (i < holes.size() && holes[i].first < off) ++i;
    holes.insert(holes.begin() + i, std::make_pair(off, bytes));
    if (i + 1 < holes.size() && holes[i].first + holes[i].second == holes[i + 1].first) {
      holes[i].second += holes[i + 1].second; holes.erase(holes.begin() + i + 1);
    }
    if (i > 0 && holes[i - 1].first + holes[i - 1].second == holes[i].first) {
      holes[i - 1].second += holes[i].second; holes.erase(holes.begin() + i);
    }
  }
};

struct Buf {                      // one fp16 NDHWC activation in the workspace
  long long off = -1, bytes = 0;
  int N = 0, D = 0, H = 0, W = 0, ld = 0;
};

constexpr long long kHeader = 1024;   // range scale (2 floats) + its reduction word, then the arena

struct Run {
  TpzUnet* m;
  unsigned char* ws;              // NULL: sizing pass (no launches)
  cudaStream_t st;
  Arena arena;
  int launches = 0;

  Buf make(int N, int D, int H, int W, int ld) {
    Buf b;
    b.N = N; b.D = D; b.H = H; b.W = W; b.ld = ld;
    b.bytes = (long long)N * D * H * W * ld * 2;
    b.off = arena.take(b.bytes);
    return b;
  }
  void drop(Buf& b) { if (b.off >= 0) arena.give(b.off, b.bytes); b.off = -1; }
  tpz_half* ptr(const Buf& b) const { return reinterpret_cast<tpz_half*>(ws + kHeader + b.off); }

  int op(int code, const void* args, int kernels = 1) {
    launches += kernels;
    if (!ws) return 0;
    if (g_hook) return g_hook(g_hook_user, code, args);
    TPZ_CHECK(!m->mem.host, "tpz_unet: a host-weights handle runs only under the launch hook");
    const TpzOpArgs* o = reinterpret_cast<const TpzOpArgs*>(args);
    const int* i = o->i;
    switch (code) {
      case TPZ_OP_RANGE_SCALE:
        TPZ_CUDA(cudaMemsetAsync(const_cast<void*>(o->p[2]), 0, sizeof(unsigned), st));
        return tpz_range_scale((const float*)o->p[0], o->n, (float*)o->p[1], (unsigned*)o->p[2], st);
      case TPZ_OP_CONV_FIRST_TC:
        return tpz_conv_first_tc((const float*)o->p[0], i[0], i[1], i[2], (const tpz_half*)o->p[1], (const float*)o->p[2], i[3], i[4], i[5],
                                 o->f[0], i[6], (tpz_half*)o->p[3], (const float*)o->p[4], st);
      case TPZ_OP_IM2COL_FIRST:
        return tpz_im2col_first((const float*)o->p[0], i[0], i[1], i[2], i[3], i[4], (tpz_half*)o->p[1], i[5], (const float*)o->p[2], i[6], st);
      case TPZ_OP_IM2COL3D_FIRST:
        return tpz_im2col3d_first((const float*)o->p[0], i[0], i[1], i[2], i[3], i[4], i[5], (tpz_half*)o->p[1], i[6], (const float*)o->p[2],
                                  i[7], st);
      case TPZ_OP_CONV_FIRST:
        return tpz_conv_first((const float*)o->p[0], i[0], i[1], i[2], i[3], (const float*)o->p[1], (const float*)o->p[2], i[4], i[5], i[6],
                              i[7], i[8], i[9], o->f[0], i[10], (tpz_half*)o->p[3], i[11], (const float*)o->p[4], i[12], st);
      case TPZ_OP_TC_CONV:
        return tpz_tc_conv(reinterpret_cast<const TpzTcConvArgs*>(args), st);
      case TPZ_OP_MAXPOOL2:
        return tpz_maxpool2((const tpz_half*)o->p[0], i[0], i[1], i[2], i[3], i[4], i[5], i[6], (tpz_half*)o->p[1], i[7], i[8], st);
      case TPZ_OP_UPSAMPLE:
        return tpz_upsample_nearest((const tpz_half*)o->p[0], i[0], i[1], i[2], i[3], i[4], i[5], i[6], i[7], i[8], (tpz_half*)o->p[1],
                                    i[9], i[10], st);
      case TPZ_OP_CONV_LAST:
        return tpz_conv_last((const tpz_half*)o->p[0], i[0], i[1], i[2], i[3], i[4], i[5], (const float*)o->p[1], o->f[0], i[6], i[7], i[8],
                             i[9], i[10], o->f[1], o->f[2], (const float*)o->p[2], (float*)o->p[3], (const float*)o->p[4], st);
    }
    return tpz_fail(2, "tpz_unet: unknown launch code %d", code);
  }

  // one tensor-core conv: `plan`'s static block + this launch's tensors (ops.fill_tc_args)
  int conv(Plan& plan, const Buf* s0, const Buf* s1, int N, int D, int H, int W, const Buf* out, float* dot_out, const float* dot_affine) {
    TpzTcConvArgs& a = plan.a;
    const Buf* srcs[2] = {s0, s1};
    for (int si = 0; si < a.nsrc; ++si) {
      const Buf& b = *srcs[si];
      TpzTcSrc& s = a.src[si];
      s.ptr = ws ? ptr(b) : nullptr; s.N = b.N; s.D = b.D; s.H = b.H; s.W = b.W; s.ld = b.ld;
    }
    a.N = N; a.Do = D; a.Ho = H; a.Wo = W;
    a.res = nullptr; a.res_scale = nullptr;
    if (out) { a.out = ws ? ptr(*out) : nullptr; a.out_ld = out->ld; a.out_coff = 0; a.out_lo = plan.split_out ? plan.co_store : 0; }
    else { a.out = nullptr; a.out_lo = 0; }
    a.range = reinterpret_cast<const float*>(ws);
    if (dot_out) { a.dot_w = reinterpret_cast<const float*>(plan.dotw_mem); a.dot_out = dot_out; a.dot_affine = dot_affine; }
    else { a.dot_w = nullptr; a.dot_out = nullptr; a.dot_affine = nullptr; }
    return op(TPZ_OP_TC_CONV, &a);
  }

  int pool(const Buf& in, bool split, Buf* out) {   // split: channels are [hi | lo] halves, the maximum is over hi + lo
    const int dims = m->dims;
    *out = make(in.N, dims == 3 ? in.D / 2 : in.D, in.H / 2, in.W / 2, in.ld);
    TpzOpArgs o;
    memset(&o, 0, sizeof(o));
    o.p[0] = ws ? ptr(in) : nullptr; o.p[1] = ws ? ptr(*out) : nullptr;
    const int C = split ? in.ld / 2 : in.ld;
    const int v[] = {in.N, in.D, in.H, in.W, C, in.ld, dims, in.ld, split ? C : 0};
    memcpy(o.i, v, sizeof(v));
    return op(TPZ_OP_MAXPOOL2, &o);
  }

  int forward(const float* x, int N, int D, int H, int W, const float* stats, float* y) {
    const int dims = m->dims, depth = m->depth, ndec = depth - 1;
    const float* range = reinterpret_cast<const float*>(ws);
    int rc;
    TpzOpArgs o;
    {                             // fp16 range guard: activations are stored multiplied by a power of two chosen from max|x|
      memset(&o, 0, sizeof(o));
      o.p[0] = x; o.p[1] = ws; o.p[2] = ws ? ws + 16 : nullptr; o.n = (long long)N * D * H * W;
      if ((rc = op(TPZ_OP_RANGE_SCALE, &o, 2))) return rc;
    }
    // ---- enc1 ----
    Buf h;
    const int cp = rup(m->nf);
    const bool pool1 = depth > 1;
    bool pooled = false;
    if (m->first_mode == 0) {
      h = pool1 ? make(N, 1, H / 2, W / 2, cp) : make(N, 1, H, W, cp);
      memset(&o, 0, sizeof(o));
      o.p[0] = x; o.p[1] = m->first_w16; o.p[2] = m->first_b; o.p[3] = ws ? ptr(h) : nullptr; o.p[4] = range;
      const int v[] = {N, H, W, cp, m->k1, m->k1 / 2, pool1 ? 1 : 0};
      memcpy(o.i, v, sizeof(v));
      o.f[0] = m->slope;
      if ((rc = op(TPZ_OP_CONV_FIRST_TC, &o))) return rc;
      pooled = pool1;
    } else if (m->first_mode == 1) {
      Buf col = make(N, D, H, W, m->first_ld);
      memset(&o, 0, sizeof(o));
      o.p[0] = x; o.p[1] = ws ? ptr(col) : nullptr; o.p[2] = range;
      const int v[] = {N * D, H, W, m->k1, m->k1 / 2, m->first_ld, 0};
      memcpy(o.i, v, sizeof(v));
      if ((rc = op(TPZ_OP_IM2COL_FIRST, &o))) return rc;
      h = make(N, D, H, W, m->first_plan.out_channels());
      if ((rc = conv(m->first_plan, &col, nullptr, N, D, H, W, &h, nullptr, nullptr))) return rc;
      drop(col);
    } else {
      const int ld = m->first_split ? 2 * cp : cp;
      h = make(N, D, H, W, ld);
      memset(&o, 0, sizeof(o));
      o.p[0] = x; o.p[1] = m->first_w32; o.p[2] = m->first_b; o.p[3] = ws ? ptr(h) : nullptr; o.p[4] = range;
      const int v[] = {N, D, H, W, m->nf, dims == 3 ? m->k1 : 1, m->k1, m->k1, 1, m->k1 / 2, 1, ld, m->first_split ? cp : 0};
      memcpy(o.i, v, sizeof(v));
      o.f[0] = m->slope;
      if ((rc = op(TPZ_OP_CONV_FIRST, &o))) return rc;
    }
    if (pool1 && !pooled) {
      Buf p;
      if ((rc = pool(h, m->first_split, &p))) return rc;
      drop(h);
      h = p;
    }
    // ---- encoder: skips = [p1 .. p_{depth-1}] (pooled outputs); enc{depth} has no pool ----
    std::vector<Buf> skips;
    skips.push_back(h);
    for (int i = 2; i <= depth; ++i) {
      Plan& pl = m->enc[i - 2];
      Buf out = make(h.N, h.D, h.H, h.W, pl.out_channels());
      if ((rc = conv(pl, &h, nullptr, h.N, h.D, h.H, h.W, &out, nullptr, nullptr))) return rc;
      if (i < depth) {
        Buf p;
        if ((rc = pool(out, pl.split_out, &p))) return rc;
        drop(out);
        h = p;
        skips.push_back(h);
      } else {
        h = out;
      }
    }
    if (depth > 1) drop(skips.back());              // p_{depth-1} feeds only enc{depth}
    // ---- decoder: level l joins p_{l-1} (level 1 joins the raw image) ----
    for (int l = ndec; l >= 1; --l) {
      DecLevel& L = m->dec[l];
      Buf other;
      int oN, oD, oH, oW;
      if (l > 1) {
        other = skips[l - 2];
        oN = other.N; oD = other.D; oH = other.H; oW = other.W;
      } else {
        oN = N; oD = D; oH = H; oW = W;
        other = make(N, D, H, W, m->raw_split ? 2 * m->ntap_store : m->ntap_store);
        memset(&o, 0, sizeof(o));
        o.p[0] = x; o.p[1] = ws ? ptr(other) : nullptr; o.p[2] = range;
        const int lo = m->raw_split ? m->ntap_store : 0;          // (hi, lo) taps: lo half at channel t + ntap_store
        if (dims == 2) {
          const int v[] = {N, H, W, m->k_top, m->k_top / 2, m->ntap_store, lo};
          memcpy(o.i, v, sizeof(v));
          if ((rc = op(TPZ_OP_IM2COL_FIRST, &o))) return rc;
        } else {
          const int v[] = {N, D, H, W, m->k_top, m->k_top / 2, m->ntap_store, lo};
          memcpy(o.i, v, sizeof(v));
          if ((rc = op(TPZ_OP_IM2COL3D_FIRST, &o))) return rc;
        }
      }
      Buf oa = make(oN, oD, oH, oW, L.a.out_channels());
      const bool exact2 = oH == 2 * h.H && oW == 2 * h.W && (dims == 2 || oD == 2 * h.D);
      if (exact2) {
        for (Plan& ph : L.up2)                      // fused nearest-2x up-sampling: one launch per output phase, reading h itself
          if ((rc = conv(ph, &h, &other, oN, oD, oH, oW, &oa, nullptr, nullptr))) return rc;
      } else {
        Buf up = make(oN, oD, oH, oW, h.ld);
        memset(&o, 0, sizeof(o));
        o.p[0] = ws ? ptr(h) : nullptr; o.p[1] = ws ? ptr(up) : nullptr;
        const int v[] = {h.N, h.D, h.H, h.W, h.ld, h.ld, oD, oH, oW, h.ld, 0};
        memcpy(o.i, v, sizeof(v));
        if ((rc = op(TPZ_OP_UPSAMPLE, &o))) return rc;
        if ((rc = conv(L.a, &up, &other, oN, oD, oH, oW, &oa, nullptr, nullptr))) return rc;
        drop(up);
      }
      drop(h);
      drop(other);
      if (l > 1) skips[l - 2].off = -1;
      Buf ob = make(oN, oD, oH, oW, L.b.out_channels());
      if ((rc = conv(L.b, &oa, nullptr, oN, oD, oH, oW, &ob, nullptr, nullptr))) return rc;
      drop(oa);
      h = ob;
    }
    // ---- dec1.4: Cout = 1, dense fp32 out, de-normalised in the same kernel when `stats` is given ----
    if (m->last_simt) {
      memset(&o, 0, sizeof(o));
      o.p[0] = ws ? ptr(h) : nullptr; o.p[1] = m->last_w; o.p[2] = stats; o.p[3] = y; o.p[4] = range;
      const int v[] = {h.N, h.D, h.H, h.W, m->last_strict ? 2 * m->last_cstore : m->last_cstore, h.ld, dims == 3 ? m->last_k : 1, m->last_k,
                       m->last_k, 1, m->last_k / 2};
      memcpy(o.i, v, sizeof(v));
      o.f[0] = m->last_b; o.f[1] = 1.f; o.f[2] = 0.f;
      if ((rc = op(TPZ_OP_CONV_LAST, &o))) return rc;
    } else {
      if ((rc = conv(m->last_tc, &h, nullptr, h.N, h.D, h.H, h.W, nullptr, y, stats))) return rc;
    }
    drop(h);
    return 0;
  }
};

int check_geometry(const TpzUnet* m, int N, int D, int H, int W) {
  TPZ_CHECK(m && N > 0 && D > 0 && H > 0 && W > 0, "tpz_unet: bad geometry %d x %d x %d x %d", N, D, H, W);
  TPZ_CHECK(m->dims == 3 || D == 1, "tpz_unet: a 2-D model takes D = 1");
  const int div = 1 << (m->depth - 1);
  TPZ_CHECK(H >= div && W >= div && (m->dims == 2 || D >= div), "tpz_unet: input %d x %d x %d is smaller than the %d pooling stages allow",
            D, H, W, m->depth - 1);
  return 0;
}

int forward(TpzUnet* m, const float* x, int N, int D, int H, int W, const float* stats, float* y, void* workspace, long long workspace_bytes,
            void* stream) {
  TPZ_CHECK(m && x && y && workspace, "tpz_unet forward: null argument");
  int rc = check_geometry(m, N, D, H, W);
  if (rc) return rc;
  TPZ_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tpz_unet forward: workspace must be 256-byte aligned");
  Run size{m, nullptr, ST(stream)};
  if ((rc = size.forward(x, N, D, H, W, stats, y))) return rc;
  const long long need = kHeader + size.arena.peak;
  TPZ_CHECK(workspace_bytes >= need, "tpz_unet forward: workspace of %lld bytes, need %lld (tpz_unet_workspace_bytes)", workspace_bytes, need);
  Run run{m, reinterpret_cast<unsigned char*>(workspace), ST(stream)};
  return run.forward(x, N, D, H, W, stats, y);
}

}  // namespace

extern "C" int tpz_unet_create(const TpzUnetDesc* desc, TpzUnet** out, void* stream) {
  TPZ_CHECK(desc && out, "tpz_unet_create: null argument");
  TPZ_CHECK((desc->dims == 2 || desc->dims == 3) && desc->depth >= 2 && desc->depth <= TPZ_UNET_MAX_DEPTH,
            "tpz_unet_create: dims %d / depth %d not supported", desc->dims, desc->depth);
  TPZ_CHECK(desc->precision >= TPZ_PRECISION_FAST && desc->precision <= TPZ_PRECISION_STRICT, "tpz_unet_create: bad precision %d",
            desc->precision);
  TpzUnet* m = new (std::nothrow) TpzUnet();
  TPZ_CHECK(m != nullptr, "tpz_unet_create: out of memory");
  m->mem.host = desc->host_weights != 0;
  m->mem.st = ST(stream);
  m->dims = desc->dims; m->depth = desc->depth; m->slope = desc->slope;
  const int rc = build_unet(m, desc, ST(stream));
  if (rc) { free_unet(m); return rc; }
  *out = m;
  return 0;
}

extern "C" int tpz_unet_destroy(TpzUnet* m) { return free_unet(m); }

extern "C" long long tpz_unet_workspace_bytes(const TpzUnet* m, int N, int D, int H, int W) {
  if (check_geometry(m, N, D, H, W)) return -1;
  Run size{const_cast<TpzUnet*>(m), nullptr, nullptr};
  if (size.forward(nullptr, N, D, H, W, nullptr, nullptr)) return -1;
  return kHeader + size.arena.peak;
}

extern "C" int tpz_unet_launch_count(const TpzUnet* m, int N, int D, int H, int W) {
  if (check_geometry(m, N, D, H, W)) return -1;
  Run size{const_cast<TpzUnet*>(m), nullptr, nullptr};
  if (size.forward(nullptr, N, D, H, W, nullptr, nullptr)) return -1;
  return size.launches;
}

extern "C" int tpz_unet2d_forward(TpzUnet* m, const float* x, int B, int H, int W, const float* denorm_stats, float* y, void* workspace,
                                  long long workspace_bytes, void* stream) {
  TPZ_CHECK(m && m->dims == 2, "tpz_unet2d_forward: not a 2-D model");
  return forward(m, x, B, 1, H, W, denorm_stats, y, workspace, workspace_bytes, stream);
}

extern "C" int tpz_unet3d_forward(TpzUnet* m, const float* x, int B, int D, int H, int W, const float* denorm_stats, float* y,
                                  void* workspace, long long workspace_bytes, void* stream) {
  TPZ_CHECK(m && m->dims == 3, "tpz_unet3d_forward: not a 3-D model");
  return forward(m, x, B, D, H, W, denorm_stats, y, workspace, workspace_bytes, stream);
}

extern "C" int tpz_unet_set_launch_hook(tpz_launch_hook fn, void* user) {
  g_hook = fn;
  g_hook_user = user;
  return 0;
}

// Test hook: the packed buffers and static argument block of one plan.  which: 0 = first-layer GEMM (im2col mode), 1 = enc{index+2},
// 2 = dec{index}.0 (materialised up-sampling), 3 = dec{index}.2, 4 = dec{index}.0 phase plan `phase`, 5 = the tensor-core Cout = 1 tail.
extern "C" int tpz_unet_plan(const TpzUnet* m, int which, int index, int phase, TpzTcConvArgs* args, long long* weight_elems) {
  TPZ_CHECK(m && args, "tpz_unet_plan: null argument");
  const Plan* p = nullptr;
  if (which == 0 && m->first_mode == 1) p = &m->first_plan;
  else if (which == 1 && index >= 0 && index < (int)m->enc.size()) p = &m->enc[index];
  else if ((which == 2 || which == 3) && index >= 1 && index < (int)m->dec.size()) p = which == 2 ? &m->dec[index].a : &m->dec[index].b;
  else if (which == 4 && index >= 1 && index < (int)m->dec.size() && phase >= 0 && phase < (int)m->dec[index].up2.size()) p = &m->dec[index].up2[phase];
  else if (which == 5) p = &m->last_tc;
  TPZ_CHECK(p != nullptr, "tpz_unet_plan: no plan (%d, %d, %d)", which, index, phase);
  *args = p->a;
  if (weight_elems) *weight_elems = p->weight_elems;
  return 0;
}
