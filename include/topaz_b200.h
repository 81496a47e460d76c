/* topaz_b200 C ABI — B200 (sm_100a) kernels for the Topaz dense-CNN hot path.
 *
 * The reference (tbepler/topaz) has no native/FFI layer: its seam is torch.nn (cuDNN/ATen library calls).
 * Each entry point below names the reference call site(s) it replaces.  Conventions:
 *   - every function returns 0 on success, non-zero on error; tpz_last_error() gives the message
 *     (thread-local); nothing falls back to the CPU;
 *   - all pointers are DEVICE pointers unless named host_*; the library never frees or retains them;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous;
 *   - activations are channels-last fp16: [N][D][H][W][C] (2-D: D = 1), `ld` = channel stride in elements;
 *   - image-like endpoints (network input / logit or denoised output) are dense fp32 [N][D][H][W].
 */
#ifndef TOPAZ_B200_H
#define TOPAZ_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t tpz_half; /* IEEE fp16 bits */

const char* tpz_last_error(void);
/* Device / build info: writes SM count, cc major, cc minor. */
int tpz_device_info(int* num_sms, int* cc_major, int* cc_minor);

/* ---- tcgen05 implicit-GEMM convolution (stride 1, arbitrary dilation, 2-D/3-D, up to two input sources) ----
 * Replaces nn.Conv2d/Conv3d + bias (+BN eval affine folded by the caller) + ReLU/LeakyReLU/PReLU
 * + cropped residual add + 1x1 proj + 1x1 classifier at:
 *   topaz/model/features/resnet.py:101-105 (BasicConv.forward), :178-204 (ResidA.forward),
 *   topaz/model/features/basic.py:101-111, topaz/model/classifier.py:64-66,
 *   topaz/denoising/models.py:130-175 (UDenoiseNet.forward), :508-564 (UDenoiseNet3D.forward).     */
#define TPZ_TC_MAX_KB 512
typedef struct {
  int16_t dx, dy, dz; /* input offset of this k-block's tap (tap index * dilation), elements */
  int16_t c0;         /* first input channel of the chunk                                     */
  int32_t src;        /* which source (0/1)                                                    */
} TcKBlock;

typedef struct {
  const tpz_half* ptr; /* [N][D][H][W][ld] fp16 */
  int N, D, H, W, C, ld;
  int org[3];          /* (x,y,z) offset added to the output coordinate: -padding or +crop */
  int kw, kh;          /* in-plane tap grid of this source (taps at multiples of its lattice spacing)  */
  int lat;             /* lattice spacing of THIS source in its own pixels (0 = same as the output `lattice`).  With
                          lat != lattice the source has a different resolution than the output: lat = 1 under an output
                          lattice of 2 reads a half-resolution tensor, i.e. a fused nearest-neighbour 2x up-sampling      */
  int no_phase;        /* 1: the output phase offset is NOT added to this source's coordinates (half-resolution source) */
  int lat_z;           /* lattice spacing of this source along z (0 = lattice_z); z taps are separate plane loads          */
} TpzTcSrc;

typedef struct {
  int nsrc;
  TpzTcSrc src[2];
  const tpz_half* weights; /* [nkb][Co][KC] fp16, k-block order == kb[] order */
  int KC;                  /* channels per k-block: 64 or 32 */
  int nkb;
  TcKBlock kb[TPZ_TC_MAX_KB];
  int N, Do, Ho, Wo, Co;   /* output geometry */
  int TW, TH;              /* pixel tile of the per-tap kernel, TW*TH == 128, TW % 8 == 0 */
  int lattice;             /* in-plane dilation shared by all taps (halo-resident kernel); 0 = per-tap kernel only */
  int phase_sel;           /* 0: all lattice*lattice output phases; k>0: only phase k-1 (py*lattice+px) is computed    */
  int lattice_z, phase_z;  /* output z lattice (0/1 = every plane) and the z phase computed by this launch: output plane
                              index = zq*lattice_z + phase_z, source plane = zq*lat_z + org_z + dz (+phase_z unless no_phase) */
  const float* bias;       /* [Co] or NULL */
  float neg_slope;         /* activation: v>0 ? v : v*neg_slope (0 = ReLU, 1 = linear, 0.1 = LeakyReLU) */
  const tpz_half* res;     /* optional residual, added before the activation */
  const float* res_scale;  /* optional per-channel scale of the residual (BN after the add) */
  int res_ld, res_D, res_H, res_W, res_org[3];
  tpz_half* out;           /* [N][Do][Ho][Wo][out_ld] (+out_coff) or NULL */
  int out_ld, out_coff;
  const float* dot_w;      /* optional fused 1x1 "classifier": dot_out = sum_c act(.)*dot_w[c] + dot_b */
  float dot_b;
  float* dot_out;          /* [N][Do][Ho][Wo] fp32 or NULL */
  const float* dot_affine; /* optional device float[2] = (shift, scale): dot_out = dot_out*scale + shift (de-normalise) */
  const float* oscale;     /* optional [Co]: per-output-channel multiplier of the accumulator, v = acc*oscale[c] + bias[c].  The
                              packer divides weight rows whose fp16 image would overflow / go subnormal by a power of two and
                              puts the factor here (exact); NULL = 1 */
  const float* range;      /* optional device float[2] = (s, 1/s) from tpz_range_scale: the activations of this network run
                              are stored multiplied by s (a power of two chosen from max|input|, so that fp16 cannot overflow
                              on un-normalised micrographs).  ReLU / LeakyReLU / PReLU / max-pool networks are positively
                              homogeneous in (input, biases): the epilogue adds bias*s, the fused dot output is multiplied by
                              1/s.  NULL = (1, 1) */
  int out_lo;              /* 0: fp16 output.  > 0 (strict mode): every output value v is stored as the pair hi = fp16(v) at
                              channel c and lo = fp16(v - hi) at channel c + out_lo (22 significand bits) */
} TpzTcConvArgs;

int tpz_tc_conv(const TpzTcConvArgs* host_args, void* stream);   /* dispatches: halo-resident kernel when
                                                                     eligible, else the per-tap kernel   */
int tpz_tc_conv_v1(const TpzTcConvArgs* host_args, void* stream);/* per-tap TMA loads (one A tile per k-block) */
int tpz_tc_conv_v2(const TpzTcConvArgs* host_args, void* stream);/* halo-resident A tile + poly-phase lattice  */

/* ---- model-level entry points: a dense ("filled") classifier network as one handle ----
 * What a non-Python host binds to score micrographs: topaz/extract.py:224-256 (score_images: model.eval(); model.fill();
 * per-image forward) over topaz/model/classifier.py:48-66 on ResNet8/16 (features/resnet.py:50-339) or conv31/63/127
 * (features/basic.py:12-111).  The layer list describes the network in its FILLED geometry (strides turned into dilations,
 * resnet.py:208-251); every parameter pointer is a DEVICE pointer in the reference's own layout (OIHW fp32), read once by
 * tpz_model_create / tpz_model_update_weights, which fold eval-mode BatchNorm, pad channels and repack to fp16 k-blocks ON the
 * device (no parameter ever crosses to the host; the 4-byte classifier bias is the one read-back).  The handle owns the packed
 * weights; activations live in the caller's workspace.  One stream at a time per handle. */
#define TPZ_LAYER_CONV 0   /* conv k x k (dilation dil0) [+ BatchNorm bn0] + activation slope0: BasicConv, conv31/63/127 layers */
#define TPZ_LAYER_RESID 1  /* ResidA: conv0 3x3 cin->cin (dil0) [+bn0] + act slope0; conv1 3x3 cin->cout (dil1) + cropped skip
                              (1x1 `proj` when cin != cout) [+bn1] + act slope1 (resnet.py:108-204) */
typedef struct {
  int kind;
  int cin, cout;
  int k;                     /* CONV: kernel size; RESID: 3 */
  int dil0, dil1;
  float slope0, slope1;      /* v > 0 ? v : v*slope (0 = ReLU; PReLU: its learned scalar) */
  const float *w0, *b0;      /* CONV: the conv; RESID: conv0.  b may be NULL (BatchNorm models have no conv bias) */
  const float *w1, *b1;      /* RESID: conv1 */
  const float *proj;         /* RESID, cin != cout: [cout][cin][1][1]; NULL = identity skip */
  const float *bn0, *bn1;    /* eval-mode BatchNorm as [4][C] = gamma, beta, running_mean, running_var; NULL = none */
  float eps0, eps1;
} TpzLayerDesc;
typedef struct TpzModel TpzModel;
/* layers[0] must be the Cin = 1 first convolution, layers[nlayers-1] a CONV whose output feeds the 1x1 classifier
 * (cls_w [C], cls_b [1], DEVICE); `pad` = the single input padding of the filled network (width // 2, resnet.py:240-243). */
int tpz_model_create(const TpzLayerDesc* layers, int nlayers, const float* cls_w, const float* cls_b, int pad, TpzModel** out,
                     void* stream);
/* Same architecture, new parameter values (and possibly new pointers): repack on the device (after optimizer steps). */
int tpz_model_update_weights(TpzModel* model, const TpzLayerDesc* layers, int nlayers, const float* cls_w, const float* cls_b,
                             void* stream);
/* tpz_model_create / tpz_model_update_weights return TPZ_E_WEIGHT_RANGE when a (BatchNorm-folded) weight row cannot be held in fp16
 * (max|w| > 2^14 or below 2^-10): such rows need the row-scaled plans (TpzTcConvArgs.oscale), which only the layer-granular path
 * builds.  Non-finite weights return 3.  The handle must not be used for forward after either. */
#define TPZ_E_WEIGHT_RANGE 4
int tpz_model_destroy(TpzModel* model);
/* Optional timing of the last conv step + fused classifier (the dominant kernel of a dense forward; bench.py's roofline figure):
 * tpz_model_timing(m, 1) records a CUDA event pair around it at every forward (ring of 64), tpz_model_timing_read returns the
 * recorded durations in ms (oldest first) and resets, tpz_model_timing(m, 0) stops. */
int tpz_model_timing(TpzModel* m, int enable);
int tpz_model_timing_read(TpzModel* m, float* ms, int capacity, int* count);
/* Bytes of DEVICE workspace (256-byte aligned) tpz_resnet_dense_forward needs for a batch of B images H x W; -1 if the
 * image is smaller than the receptive field allows. */
long long tpz_workspace_bytes(const TpzModel* model, int B, int H, int W);
/* y[B][H][W] = classifier logits of x[B][H][W] (both dense fp32 DEVICE): range scale, first layer, tensor-core convs, fused
 * 1x1 classifier.  Asynchronous on `stream`. */
int tpz_resnet_dense_forward(TpzModel* model, const float* x, int B, int H, int W, float* y, void* workspace,
                             long long workspace_bytes, void* stream);

/* Test hook: copies the packed fp16 weights [nkb][Co][KC] / fp32 bias [Co] of conv step `step` (launch order; -1 = the first
 * layer) into caller-provided DEVICE buffers (either may be NULL) and reports the sizes. */
int tpz_model_step_buffers(const TpzModel* model, int step, void* weights_out, long long weight_capacity, float* bias_out,
                           long long* weight_elems, int* co_store, int* kc, int* nkb, void* stream);

/* Test hook: copy of the argument block conv step `step` was last launched with. */
int tpz_model_step_args(const TpzModel* model, int step, TpzTcConvArgs* out);

/* ---- model-level entry points, part 2: the U-Net denoisers as one handle (tpz_unet2d_forward / tpz_unet3d_forward) ----
 * What a non-Python host binds to denoise micrograph patches / tomogram patches: topaz/denoise.py:274-296 (Denoise._denoise:
 * normalise, model forward, de-normalise) over topaz/denoising/models.py:74-175 (UDenoiseNet), :178-244 (UDenoiseNetSmall) and
 * :452-564 (UDenoiseNet3D).  The description lists the network's convolutions in the reference's own parameter layout (fp32
 * OIHW / OIDHW, DEVICE pointers; state_dict keys enc{i}.0.*, dec{l}.{0,2}.*, dec1.4.*): enc{i} = conv + LeakyReLU (+ MaxPool(2)
 * for i < depth), dec{l} = two convs over cat[nearest-upsampled, skip p_{l-1}] (level 1: the raw image), dec1.4 the Cout = 1 tail.
 * tpz_unet_create reads the weights once (one synchronising device-to-host copy), builds the k-block plans -- including the per-phase
 * plans of the fused nearest-2x up-sampling -- and keeps the packed fp16 weights in the handle; activations live in the caller's
 * workspace.  One stream at a time per handle.
 * Returns TPZ_E_WEIGHT_RANGE / 3 like tpz_model_create. */
#define TPZ_UNET_MAX_DEPTH 8
#define TPZ_PRECISION_FAST 0
#define TPZ_PRECISION_AUTO 1
#define TPZ_PRECISION_STRICT 2
typedef struct {
  const float *w, *b;        /* [cout][cin][k]^dims fp32, bias [cout] (b may be NULL) */
  int cout, cin, k;
} TpzConvDesc;
typedef struct {
  int dims;                  /* 2 or 3 */
  int depth;                 /* encoder stages: 6 (UDenoiseNet, UDenoiseNet3D), 4 (UDenoiseNetSmall) */
  TpzConvDesc enc[TPZ_UNET_MAX_DEPTH];     /* enc[i-1] = enc{i}.0 */
  TpzConvDesc dec_a[TPZ_UNET_MAX_DEPTH];   /* dec_a[l] = dec{l}.0, l = 1 .. depth-1 (entry 0 unused) */
  TpzConvDesc dec_b[TPZ_UNET_MAX_DEPTH];   /* dec_b[l] = dec{l}.2 */
  TpzConvDesc last;          /* dec1.4 */
  float slope;               /* LeakyReLU slope (0.1) */
  int precision;             /* TPZ_PRECISION_FAST: fp16 operands, fp32 accumulation (the 11-bit significand of the reference GPU
                                path's TF32); _STRICT: every activation / weight a (hi, lo) fp16 pair, products as hi*hi + hi*lo +
                                lo*hi (22 bits; 3x the MMAs); _AUTO: fast, except the last four convolutions of a 3-D network, where
                                the pretrained unet-3d models cancel ~100x (the Python engine's TPZ_PRECISION default) */
  int host_weights;          /* test handles only: the pointers above are HOST pointers, the packed buffers stay on the host and the
                                handle runs only under tpz_unet_set_launch_hook */
} TpzUnetDesc;
typedef struct TpzUnet TpzUnet;
int tpz_unet_create(const TpzUnetDesc* desc, TpzUnet** out, void* stream);
int tpz_unet_destroy(TpzUnet* model);
/* Bytes of DEVICE workspace (256-byte aligned) a forward of N patches of D x H x W needs (2-D: D = 1); -1 if the patch is smaller
 * than the pooling stages allow.  tpz_unet_launch_count: kernels launched by that forward. */
long long tpz_unet_workspace_bytes(const TpzUnet* model, int N, int D, int H, int W);
int tpz_unet_launch_count(const TpzUnet* model, int N, int D, int H, int W);
/* y = model(x), x and y dense fp32 DEVICE [B][H][W] / [B][D][H][W].  denorm_stats (device float[2] = mean, std; may be NULL): the
 * output is de-normalised, y*std + mean, inside the last kernel (denoise.py:295).  Asynchronous on `stream`. */
int tpz_unet2d_forward(TpzUnet* model, const float* x, int B, int H, int W, const float* denorm_stats, float* y, void* workspace,
                       long long workspace_bytes, void* stream);
int tpz_unet3d_forward(TpzUnet* model, const float* x, int B, int D, int H, int W, const float* denorm_stats, float* y,
                       void* workspace, long long workspace_bytes, void* stream);
/* Test hooks.  tpz_unet_set_launch_hook(fn, user): every kernel launch of the U-Net entry points is handed to fn(user, op, args)
 * instead of the device (fn = NULL restores the device path) -- tests/test_unet_abi.py runs the whole C++ launch sequence on the CPU
 * simulation of the kernels this way.  args is a TpzTcConvArgs for TPZ_OP_TC_CONV and a TpzOpArgs otherwise, with the callee's
 * arguments in declaration order: pointers in p[], ints in i[], floats in f[] (the one long long in n).
 * tpz_unet_plan: the static argument block (weights / bias pointing at the handle's packed buffers) of one plan; which: 0 = first-layer
 * GEMM, 1 = enc{index+2}, 2 = dec{index}.0, 3 = dec{index}.2, 4 = phase plan `phase` of dec{index}.0, 5 = tensor-core Cout = 1 tail. */
enum { TPZ_OP_RANGE_SCALE = 0, TPZ_OP_CONV_FIRST_TC = 1, TPZ_OP_IM2COL_FIRST = 2, TPZ_OP_IM2COL3D_FIRST = 3, TPZ_OP_CONV_FIRST = 4,
       TPZ_OP_TC_CONV = 5, TPZ_OP_MAXPOOL2 = 6, TPZ_OP_UPSAMPLE = 7, TPZ_OP_CONV_LAST = 8 };
typedef struct { const void* p[6]; long long n; int i[16]; float f[4]; } TpzOpArgs;
typedef int (*tpz_launch_hook)(void* user, int op, const void* args);
int tpz_unet_set_launch_hook(tpz_launch_hook fn, void* user);
int tpz_unet_plan(const TpzUnet* model, int which, int index, int phase, TpzTcConvArgs* args, long long* weight_elems);

/* ---- direct (SIMT) convolutions for the thin ends and for validation ----
 * tpz_conv_first: Cin = 1 conv from a dense fp32 image, fp32 math, fused bias + activation, fp16 NDHWC out.
 *   Replaces the first BasicConv 7x7 (resnet.py:66,102), conv31/63/127 layer 0 (basic.py:47-52) and the
 *   U-Net enc1 11x11 / 7x7x7 conv (denoising/models.py:79,457).  `pad` = zero padding on every side,
 *   weights fp32 [Co][kd][kh][kw], optional fused 2x max-pool (`pool` = 1/2).                        */
int tpz_conv_first(const float* x, int N, int D, int H, int W, const float* w, const float* bias, int Co,
                   int kd, int kh, int kw, int dil, int pad, float neg_slope, int pool, tpz_half* out,
                   int out_ld, const float* range, int out_lo, void* stream);
/* tpz_range_scale: range[0] = s, range[1] = 1/s with s a power of two chosen from max|x| (device reduction, no host
 *   synchronisation): s = 1 when max|x| <= 64 (normalised micrographs: results bit-identical to the unscaled path),
 *   otherwise s < 1 with max|x*s| in [4, 8).  Never > 1: the biases are scaled with the activations, so a tiny input is left
 *   alone (its activations are bias-dominated and in range).  Non-finite input gives s = 1.  `work` is a device uint32 scratch word.
 *   The dense kernels take `range` and keep every fp16 activation scaled by s (see TpzTcConvArgs.range), which is how an
 *   un-normalised micrograph (|x| >> 65504) is scored without overflow; the reference's fp32/TF32 path has that range
 *   natively (classifier.py:48-66 applies no normalisation of its own). */
int tpz_range_scale(const float* x, long long n, float* range, unsigned* work, void* stream);
/* tpz_conv_first_tc: the same Cin = 1 convolution (2-D, dilation 1) as ONE tcgen05 kernel: each CTA builds the im2col
 *   tile of 128 output pixels in shared memory (SWIZZLE_128B K-major, generic stores + fence.proxy.async) and multiplies
 *   it with the smem-resident weights; HBM sees only the fp32 image in and the fp16 [N][Ho][Wo][Cp] map out.
 *   w_packed: fp16 [ceil(k*k/64)][Cp][64] (tap t = r*k+s, zero padded); bias fp32 [Cp]; Cp in {32,64}; k in {3,5,7,11}.
 *   pool = 1 fuses the MaxPool2d(2) that follows the U-Net's enc1 conv (denoising/models.py:80): out is then
 *   [N][Ho/2][Wo/2][Cp] and the full-resolution map is never written. */
int tpz_conv_first_tc(const float* x, int B, int H, int W, const tpz_half* w_packed, const float* bias, int Cp, int k,
                      int pad, float neg_slope, int pool, tpz_half* out, const float* range, void* stream);
int tpz_conv_first_tc_supported(int k, int Cp);
/* tpz_im2col_first: im2col of a single-channel 2-D image (k x k taps -> channels, zero padded to ld) so that
 *   Cin = 1 convs (first BasicConv 7x7, U-Net enc1 11x11, the raw-image slice of U-Net dec1.0) run as a
 *   1-tap tensor-core GEMM through tpz_tc_conv.  out: fp16 [N][1][Ho][Wo][ld].  `range`: see tpz_range_scale (taps are
 *   stored multiplied by s).  out_lo > 0 (strict mode): the fp16 rounding residuals are stored at channel t + out_lo.     */
int tpz_im2col_first(const float* x, int N, int H, int W, int k, int pad, tpz_half* out, int ld, const float* range,
                     int out_lo, void* stream);
/* tpz_im2col3d_first: the 3-D analogue ('same' padding, pad = k/2): out fp16 [N][D][H][W][ld], channel t = (dz*k+dy)*k+dx;
 *   the raw-volume slice of UDenoiseNet3D dec1.0 (denoising/models.py:555) as a second tensor-core source. */
int tpz_im2col3d_first(const float* x, int N, int D, int H, int W, int k, int pad, tpz_half* out, int ld,
                       const float* range, int out_lo, void* stream);
/* tpz_conv_last: Cout = 1 conv from fp16 NDHWC to dense fp32 (classifier 1x1, classifier.py:65; U-Net
 *   dec1.4, denoising/models.py:127,505).  out = (sum + bias) * out_scale + out_shift, then, if
 *   affine_stats (device float[2] = mean,std) is given, out = out*std + mean (denoise.py:295 de-normalise).               */
int tpz_conv_last(const tpz_half* x, int N, int D, int H, int W, int C, int ld, const float* w, float bias,
                  int kd, int kh, int kw, int dil, int pad, float out_scale, float out_shift,
                  const float* affine_stats, float* out, const float* range, void* stream);
/* tpz_conv_generic: reference-quality fp32-accumulate conv on fp16 NDHWC tensors with stride support;
 *   used for validation of tpz_tc_conv and for shapes the tensor-core kernel does not cover.
 *   weights fp32 [Co][Ci][kd][kh][kw] (reference OIHW layout). Two sources are concatenated on channels. */
int tpz_conv_generic(const tpz_half* x0, int C0, int ld0, const tpz_half* x1, int C1, int ld1, int N, int D,
                     int H, int W, const float* w, const float* bias, int Co, int kd, int kh, int kw,
                     int stride, int dil, int pad, float neg_slope, const tpz_half* res, int res_ld,
                     int res_org, tpz_half* out, int out_ld, int Do, int Ho, int Wo, void* stream);

/* ---- pooling / resampling (F.max_pool2d/3d(2), F.interpolate(mode='nearest') + torch.cat) ----
 *   denoising/models.py:82-97 (MaxPool), :140-171 (interpolate + cat).                                */
int tpz_maxpool2(const tpz_half* x, int N, int D, int H, int W, int C, int ld, int dims, tpz_half* out,
                 int out_ld, int lo_off, void* stream);   /* lo_off > 0 (strict mode): channels [0,C) are hi parts, [lo_off,
                                                             lo_off+C) lo parts; the maximum is taken over hi+lo */
int tpz_upsample_nearest(const tpz_half* x, int N, int D, int H, int W, int C, int ld, int Do, int Ho, int Wo,
                         tpz_half* out, int out_ld, int out_coff, void* stream);

/* ---- normalisation prologue (denoise.py:283-284 torch mean / unbiased std; stats.py:36-46) ----
 * stats[0]=mean, stats[1]=std (unbiased if `unbiased`), computed on device in fp64 accumulators;
 * y = (x-mean)/std written as fp32 (y may alias x).  `stats` is a device float[2], `work4` a device double[4] scratch.                  */
int tpz_meanstd(const float* x, long long n, int unbiased, float* stats, double* work4, void* stream);
int tpz_affine(const float* x, long long n, const float* stats, int inverse, float* y, void* stream);

/* ---- training step of the strided (unfilled) classifier: reference topaz/methods.py:98-165 ----
 * fp32 NHWC activations, OIHW fp32 weights (the reference's parameter layout; gradients are written in the same
 * layout so they alias torch .grad storage).  Input coordinate of a tap = o*stride + tap*dil + org.
 *  tpz_conv_fwd_f32  : y = relu?(conv(x,w) + bias + res[(o*res_stride+res_org)])      (replaces cuDNN fprop)
 *  tpz_conv_dgrad_f32: dx (=|+=) conv^T(dy,w), optionally masked by relu_mask > 0       (replaces cuDNN bwd-data)
 *  tpz_conv_wgrad_f32: dw += x (*) dy, db += sum dy   (atomic accumulation; zero first) (replaces cuDNN bwd-filter)
 *  tpz_ge_binomial_loss_grad: fused GE-binomial loss, metrics and closed-form d/dscore (methods.py:103-151);
 *      scores/labels = the whole (global) minibatch, dscores written for the local shard [lo,hi)
 *  tpz_adam_step: fused flat-buffer Adam (torch.optim.Adam defaults) + L2 term + gradient zeroing (methods.py:153-160) */
int tpz_conv_fwd_f32(const float* x, int N, int H, int W, int Ci, const float* w, const float* bias, int Co, int kh,
                     int kw, int stride, int dil, int org, const float* res, int res_H, int res_W, int res_org,
                     int res_stride, int relu, float* y, int Ho, int Wo, void* stream);
int tpz_conv_dgrad_f32(const float* dy, int N, int Ho, int Wo, int Co, const float* w, int Ci, int kh, int kw, int stride,
                       int dil, int org, const float* relu_mask, int accumulate, float* dx, int H, int W, void* stream);
int tpz_conv_wgrad_f32(const float* x, int N, int H, int W, int Ci, const float* dy, int Ho, int Wo, int Co, int kh,
                       int kw, int stride, int dil, int org, float* dw, float* db, void* stream);
/* Tensor-core (mma.sync, error-compensated 3xTF32) variants for channel counts that are multiples of 16/32.
 * Weights come from tpz_train_repack: OIHW -> [tap][ci][co] (forward) and [tap][co][ci] (dgrad); `descs` is a device
 * array of {int64 src, dst_fwd, dst_dg; int32 Co, Ci, taps, pad} (element offsets into flat_params / packed). */
/* Precision of the mma training convs: 0 (default) = error-compensated 3xTF32, fp32-level accuracy (the reference's CPU
 * results); 1 = single-pass TF32, the precision of the reference's own GPU path (torch.backends.cudnn.allow_tf32 defaults
 * to True).  Also selectable with the environment variable TPZ_TRAIN_TF32=1.  Returns the previous mode. */
int tpz_train_set_tf32(int single_pass);
int tpz_train_repack(const float* flat_params, const void* descs, int ndesc, long long max_elems, float* packed, void* stream);
int tpz_conv_fwd_mma(const float* x, int N, int H, int W, int Ci, const float* w_fwd_packed, const float* bias, int Co,
                     int kh, int kw, int stride, int dil, int org, const float* res, int res_H, int res_W, int res_org,
                     int res_stride, int relu, float* y, int Ho, int Wo, void* stream);
int tpz_conv_dgrad_mma(const float* dy, int N, int Ho, int Wo, int Co, const float* w_dg_packed, int Ci, int kh, int kw,
                       int stride, int dil, int org, const float* relu_mask, int accumulate, float* dx, int H, int W,
                       void* stream);
int tpz_conv_wgrad_mma(const float* x, int N, int H, int W, int Ci, const float* dy, int Ho, int Wo, int Co, int kh, int kw,
                       int stride, int dil, int org, float* dw, void* stream);
/* tcgen05 versions (tpz_train_tc.cu): the same three products as error-compensated 3xTF32 on `tcgen05.mma.kind::tf32` with
 * TMEM accumulators, for Ci % 32 == 0 and Co % 32 == 0.  Operands are split hi / lo once by the staging threads (activations,
 * gradients) or by tpz_train_repack_tc (weights: OIHW -> forward [tap][ci/32][co][32] and dgrad [tap][co/32][ci][32], each a hi
 * plane followed by a lo plane; `descs` as for tpz_train_repack, offsets into a packed buffer of 4 x numel per weight).
 * Same argument meaning as the _mma entry points.  TPZ_TRAIN_FLUSH = chunks of 32 K-elements accumulated in TMEM before the
 * round-to-nearest drain into registers (default 2; 0 = never). */
int tpz_train_repack_tc(const float* flat_params, const void* descs, int ndesc, long long max_elems, float* packed, void* stream);
int tpz_conv_fwd_tc(const float* x, int N, int H, int W, int Ci, const float* w_fwd_packed, const float* bias, int Co,
                    int kh, int kw, int stride, int dil, int org, const float* res, int res_H, int res_W, int res_org,
                    int res_stride, int relu, float* y, int Ho, int Wo, void* stream);
int tpz_conv_dgrad_tc(const float* dy, int N, int Ho, int Wo, int Co, const float* w_dg_packed, int Ci, int kh, int kw,
                      int stride, int dil, int org, const float* relu_mask, int accumulate, float* dx, int H, int W,
                      void* stream);
int tpz_conv_wgrad_tc(const float* x, int N, int H, int W, int Ci, const float* dy, int Ho, int Wo, int Co, int kh, int kw,
                      int stride, int dil, int org, float* dw, void* stream);
/* Fused variants (one kernel when the halo-resident path takes the layer, the separate kernels otherwise):
 *  tpz_conv_dgrad_tc_res : dx = mask(dgrad [+ dx] + res placed at (res_org, res_org)) -- ResidA.conv0's data gradient plus the gradient
 *                          of the cropped identity skip (resnet.py:198-203) and the ReLU mask of the block input
 *  tpz_conv_wgrad_tc_bias: also db[c] += sum over pixels of dy[.][c] (db may be NULL)                                                */
int tpz_conv_dgrad_tc_res(const float* dy, int N, int Ho, int Wo, int Co, const float* w_dg_packed, int Ci, int kh, int kw, int stride,
                          int dil, int org, const float* relu_mask, int accumulate, const float* res, int res_H, int res_W,
                          int res_org, float* dx, int H, int W, void* stream);
int tpz_conv_wgrad_tc_bias(const float* x, int N, int H, int W, int Ci, const float* dy, int Ho, int Wo, int Co, int kh, int kw,
                           int stride, int dil, int org, float* dw, float* db, void* stream);
/* Cin = 1 first layer of the training net (resnet.py:294, 7x7 stride 2), valid conv, org = 0 */
int tpz_first_fwd_f32(const float* x, int N, int H, int W, const float* w, const float* bias, int Co, int k, int stride,
                      int relu, float* y, int Ho, int Wo, void* stream);
int tpz_first_wgrad_f32(const float* x, int N, int H, int W, const float* dy, int Ho, int Wo, int Co, int k, int stride,
                        float* dw, void* stream);
/* The same first layer on the tensor core (3xTF32, im2col tile built in shared memory; 7 x 7, Co = 32 only).  Return -1 without
 * launching when the shape is not covered: the caller falls back to the fp32 kernels above. */
int tpz_first_fwd_tc(const float* x, int N, int H, int W, const float* w, const float* bias, int Co, int k, int stride, int relu,
                     float* y, int Ho, int Wo, void* stream);
int tpz_first_wgrad_tc(const float* x, int N, int H, int W, const float* dy, int Ho, int Wo, int Co, int k, int stride, float* dw,
                       float* db /* bias gradient += column sums of dy; may be NULL */, void* stream);
int tpz_bias_grad_f32(const float* dy, long long P, int C, float* db, void* stream);   /* db[c] += sum_p dy[p][c] */
/* Classifier head in training (classifier.py:29,65: 1x1 conv C -> 1 on M pixels, x [M][C] fp32):
 *  tpz_cls_fwd_f32: y[m] = sum_c x[m][c]*w[c] + bias[0]
 *  tpz_cls_bwd_f32: dx[m][c] = g[m]*w[c] (zero where masked and x[m][c] <= 0: the ReLU of the layer that produced x; dx may be NULL),
 *                   dw[c] += sum_m g[m]*x[m][c], db[0] += sum_m g[m]  (dw / db accumulate, as p.grad does) */
int tpz_cls_fwd_f32(const float* x, long long M, int C, const float* w, const float* bias, float* y, void* stream);
int tpz_cls_bwd_f32(const float* x, long long M, int C, const float* w, const float* g, int masked, float* dx, float* dw, float* db,
                    void* stream);
int tpz_relu_bwd_f32(float* dy, const float* y, long long n, void* stream);
int tpz_crop_add_f32(float* dx, int N, int H, int W, int C, const float* g, int Ho, int Wo, int org, int stride,
                     void* stream);
/* Training-mode BatchNorm of the strided classifier (nn.BatchNorm2d inside BasicConv / ResidA when bn=True, the default of
 * `topaz train`: topaz/commands/train.py:91, topaz/model/features/resnet.py:68-70,101-104,134-141,185-204).  NHWC fp32 [P][C].
 *  tpz_bn_stats_f32     : sums[c] += sum_p x[p][c], sums[C+c] += sum_p x[p][c]^2   (fp64; caller zeroes; all-reduced by the
 *                         host across ranks so that multi-GPU training normalises with the statistics of the GLOBAL minibatch)
 *  tpz_bn_fwd_f32       : sums != NULL (training): mean = sums[c]/count, var = sums[C+c]/count - mean^2 (biased),
 *                         y = relu?((x-mean)*rsqrt(var+eps)*gamma+beta), save = {mean[C], invstd[C]}, and when running_mean
 *                         != NULL: running = (1-momentum)*running + momentum*{mean, var*count/(count-1)};
 *                         sums == NULL (eval): mean / invstd are READ from save
 *  tpz_bn_bwd_reduce_f32: sums[c] += sum_p g[p][c], sums[C+c] += sum_p g[p][c]*xhat[p][c]   (xhat = (x-mean)*invstd)
 *  tpz_bn_bwd_f32       : dx = gamma*invstd*(g - sums[c]/count - xhat*sums[C+c]/count) (dx may alias g);
 *                         dbeta[c] += local_sums[c], dgamma[c] += local_sums[C+c] (this rank's share; gradients are summed
 *                         over ranks later by the flat-gradient all-reduce)                                                */
int tpz_bn_stats_f32(const float* x, long long P, int C, double* sums, void* stream);
int tpz_bn_fwd_f32(const float* x, long long P, int C, const double* sums, long long count, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var, int relu,
                   float* y, float* save, void* stream);
int tpz_bn_bwd_reduce_f32(const float* g, const float* x, long long P, int C, const float* save, double* sums, void* stream);
int tpz_bn_bwd_f32(const float* g, const float* x, long long P, int C, const float* save, const double* sums, long long count,
                   const float* gamma, const double* local_sums, float* dgamma, float* dbeta, float* dx, void* stream);
/* PReLU (one learnable slope) / LeakyReLU of the conv31/63/127 extractors in training (topaz/model/features/basic.py:16,51,66):
 *  tpz_act_fwd_f32: y = v > 0 ? v : a*v, a = *slope_dev (nn.PReLU().weight, device) when non-NULL else slope_const
 *  tpz_act_bwd_f32: g <- g*(v > 0 ? 1 : a) in place; *dslope += sum over v <= 0 of g*v (NULL: no slope gradient)          */
int tpz_act_fwd_f32(const float* v, long long n, const float* slope_dev, float slope_const, float* y, void* stream);
int tpz_act_bwd_f32(float* g, const float* v, long long n, const float* slope_dev, float slope_const, float* dslope,
                    void* stream);
/* nn.Dropout in training (topaz/model/features/resnet.py:296-303, basic.py:58-59,71-72; `topaz train --dropout`):
 *  tpz_dropout_fwd_f32: mask[i] = Philox(seed, subsequence i/4, offset)[i%4] > p;  y = mask ? x/(1-p) : 0
 *  tpz_dropout_bwd_f32: g <- mask ? g/(1-p) : 0                                                                            */
int tpz_dropout_fwd_f32(const float* x, long long n, float p, unsigned long long seed, unsigned long long offset, float* y,
                        unsigned char* mask, void* stream);
int tpz_dropout_bwd_f32(float* g, const unsigned char* mask, long long n, float p, void* stream);
int tpz_ge_binomial_loss_grad(const float* scores, const double* labels, int B, double pi, double slack, int lo, int hi,
                              float* dscores, float* out5, void* stream);
/* PN / GE_KL / PU objectives (topaz/methods.py:25-74, 168-255, 258-322): mode 0/1/2; out6 = {loss, ge_penalty, precision,
 * tpr, fpr, aux}; aux_in = running expectation (GE_KL) or beta (PU); pi <= 0 selects the unweighted PN loss.            */
int tpz_pu_objective_loss_grad(const float* scores, const double* labels, int B, int mode, double pi, double slack,
                               double momentum, double aux_in, int lo, int hi, float* dscores, float* out6, void* stream);
int tpz_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                  float beta2, float eps, int step, float l2, float grad_scale, void* stream);
/* Same update with the step count kept on the device: *step_dev is incremented first and the bias corrections are derived
 * from it in the kernel, so the launch can be replayed from a CUDA graph of the whole training step (methods.py). */
int tpz_adam_step_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                      float beta2, float eps, int* step_dev, float l2, float grad_scale, void* stream);

/* ---- training-crop sampler + augmentation on the GPU ----
 * Replaces MultipleImageSetDataset.__getitem__ / MemoryMappedImage.get_crop (topaz/utils/data/memory_mapped_data.py:45-100,
 * 195-233).  `pixels`: all micrographs concatenated (fp32); imgs[i] = {element offset, H, W}; set_begin[nsets+1] partitions
 * the image list into sets, set_cdf = cumulative set probabilities; positives = int32 [P][3] (image, y, x) of every labelled
 * pixel; pos_mask = uint8 per pixel (same offsets) for the 'pn' rejection of labelled pixels.  Writes X [B][crop][crop] fp32
 * and Y [B] fp64; params_scratch >= 32*B bytes.  Reproducible from (seed, batch_index).                                   */
typedef struct { long long offset; int H, W; } TpzSamplerImage;
int tpz_sample_crops(int B, unsigned long long seed, unsigned long long batch_index, const TpzSamplerImage* imgs,
                     const float* pixels, const int* set_begin, const float* set_cdf, int nsets, const int* positives,
                     int num_positives, const unsigned char* pos_mask, float positive_balance, int split_pn, int rotate,
                     int flip, int crop, int big_crop, void* params_scratch, float* X, double* Y, void* stream);
int tpz_make_crops(int B, int crop, int big_crop, const TpzSamplerImage* imgs, const float* pixels, const void* params,
                   float* X, void* stream);

/* ---- micrograph preprocessing ----
 * tpz_gemm_f32: C[M][N] = A[M][K] * B[K][N], row-major fp32, 3xTF32 tensor-core product (K%16==0, N%32==0).  The
 *   Fourier-crop downsample (topaz/utils/image.py:38-61: rfft2 -> crop -> irfft2) is linear and separable; the host
 *   builds its real row / column operators once per shape and applies them with this product.
 * tpz_gmm_sums: one pass over the pixels of the 2-component GMM normalisation (topaz/stats.py:122-214) for up to 12
 *   parameter sets at once (the reference's 12 initialisations, advanced in lockstep).  sets8 (host doubles,
 *   [nsets][8]) = {mode, split, mu0-shift, mu1-shift, var0, var1, log(1-pi), log(pi)}; mode 0 = initial hard split at
 *   `split` (stats.py:136-139), mode 1 = E step.  sums (device double[nsets][7]) = {sum Z, sum p0, sum p1, sum p0*xc,
 *   sum p1*xc, sum p0*xc^2, sum p1*xc^2} with xc = x - shift.
 * tpz_select_hist: radix-select histograms over the order-preserving uint32 key of each float, for the exact order
 *   statistics behind np.quantile (stats.py:91): level 0 -> hist[4096] of key>>20; level 1 -> hist[s][4096] of
 *   (key>>8)&0xFFF for keys whose top 12 bits equal prefixes[s]; level 2 -> hist[s][256] of key&0xFF for keys whose
 *   top 24 bits equal prefixes[s] (prefixes: device uint32[nprefix <= 64]). */
int tpz_gemm_f32(const float* A, long long M, int K, const float* B, int N, float* C, void* stream);
int tpz_gmm_sums(const float* x, long long n, double shift, const double* sets8, int nsets, double* sums, void* stream);
int tpz_select_hist(const float* x, long long n, int level, const unsigned* prefixes, int nprefix, unsigned* hist,
                    void* stream);

/* ---- greedy non-maximum suppression, picks bit-identical to topaz/algorithms.py:25-63 ----
 * scores: device fp32 [H][W]; state (uint8[H*W]), list (int32[max_picks]), counters (int32[2]) are device scratch.
 * On return list[0..*host_num_picks) holds the flat indices of the picks (unordered; the caller orders them by score).  */
int tpz_nms2d(const float* scores, int H, int W, int r, float threshold, unsigned char* state, int* list, int* counters,
              int max_picks, int* host_num_picks, void* stream);
/* 3-D variant (topaz/algorithms.py:66-103): the reference suppresses FLAT indices i + delta (delta = dz*H*W + dy*W + dx
 * over the ball of radius scale*r) with no bounds handling, so neighbours wrap across rows/slices; `deltas` (device
 * int32[num_deltas], symmetric set) is built by the caller exactly as the reference builds coord_deltas. */
int tpz_nms_flat(const float* scores, long long n, const int* deltas, int num_deltas, float threshold, unsigned char* state,
                 int* list, int* counters, int max_picks, int* host_num_picks, void* stream);

/* ---- fixed filters: dense 1->1 same-padded fp32 convolution (GaussianDenoise.apply, topaz/filters.py:62-79) ---- */
int tpz_filter_f32(const float* x, int N, int D, int H, int W, const float* f, int kd, int kh, int kw, float bias, float* y,
                   void* stream);

/* ---- layout helpers ---- */
int tpz_f32_to_f16(const float* x, long long n, tpz_half* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif
