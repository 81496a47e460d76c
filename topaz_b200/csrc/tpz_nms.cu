// Greedy non-maximum suppression on the GPU with picks bit-identical to the reference's sequential algorithm
// (topaz/algorithms.py:25-63), including its quirk: suppressed coordinates are clipped to [0,H] x [0,W] (not -1), so
// a pick within r of the right border also suppresses pixel (0, y+1) of the following rows.
//
// Sequential rule: visit pixels by descending score; pick a pixel if no earlier pick suppresses it.  Parallel
// formulation (rounds): an undecided pixel q is picked when no UNDECIDED pixel p that ranks higher suppresses it
// (higher-ranked PICKED suppressors have already removed q); each new pick then removes the undecided pixels it
// suppresses THAT RANK LOWER (the relation is not symmetric at the right border, so a pick must never remove a
// higher-ranked pixel).  Rank = score descending, ties by larger flat index first.  3-4 rounds in practice.
#include "tpz_common.cuh"
#include "../../include/topaz_b200.h"

namespace {
enum : unsigned char { UND = 0, PICK = 1, SUP = 2, REJ = 3 };

__device__ __forceinline__ bool ranks_higher(const float* __restrict__ A, int p, int q) {
  const float a = A[p], b = A[q];
  return a > b || (a == b && p > q);
}
__device__ __forceinline__ int isqrt_i(int v) {
  int s = (int)sqrtf((float)v);
  while (s * s > v) --s;
  while ((s + 1) * (s + 1) <= v) ++s;
  return s;
}
// does a pick at p = (xp, yp) suppress the flat index of q = (xq, yq)?
__device__ __forceinline__ bool suppresses(int xp, int yp, int xq, int yq, int r, int W) {
  int dy = yq - yp;
  if (dy >= -r && dy <= r) {
    const int dxm = isqrt_i(r * r - dy * dy);
    if (max(xp - dxm, 0) <= xq && xq <= min(xp + dxm, W)) return true;
  }
  if (xq == 0 && yq >= 1) {            // clip-to-W quirk: (W, y') aliases onto (0, y'+1)
    dy = yq - 1 - yp;
    if (dy >= -r && dy <= r && xp + isqrt_i(r * r - dy * dy) >= W) return true;
  }
  return false;
}

__global__ void nms_init_kernel(const float* __restrict__ A, int n, float thr, unsigned char* __restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) state[i] = (A[i] > thr) ? UND : REJ;
}

// counters: [0] = number of picks so far (list length), [1] = undecided pixels left blocked this round
__global__ void nms_pick_kernel(const float* __restrict__ A, const unsigned char* __restrict__ state, int H, int W, int r,
                                int* __restrict__ list, int* __restrict__ counters, int max_picks) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= H * W || state[q] != UND) return;
  const int xq = q % W, yq = q / W;
  bool blocked = false;
  for (int pass = 0; pass < 2 && !blocked; ++pass) {
    int y0, y1, x0, x1;
    if (pass == 0) { y0 = max(0, yq - r); y1 = min(H - 1, yq + r); x0 = max(0, xq - r); x1 = min(W - 1, xq + r); }
    else {
      if (!(xq == 0 && yq >= 1)) break;
      y0 = max(0, yq - 1 - r); y1 = min(H - 1, yq - 1 + r); x0 = max(0, W - r); x1 = W - 1;
    }
    for (int yp = y0; yp <= y1 && !blocked; ++yp)
      for (int xp = x0; xp <= x1; ++xp) {
        const int p = yp * W + xp;
        if (p != q && state[p] == UND && ranks_higher(A, p, q) && suppresses(xp, yp, xq, yq, r, W)) { blocked = true; break; }
      }
  }
  if (blocked) { atomicAdd(&counters[1], 1); return; }
  const int slot = atomicAdd(&counters[0], 1);
  if (slot < max_picks) list[slot] = q;
}

__global__ void nms_commit_kernel(const int* __restrict__ list, int begin, int end, unsigned char* __restrict__ state) {
  const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < end) state[list[i]] = PICK;
}

// one block per new pick: remove the lower-ranked undecided pixels it suppresses
__global__ void nms_suppress_kernel(const float* __restrict__ A, const int* __restrict__ list, int begin, int H, int W, int r,
                                    unsigned char* __restrict__ state) {
  const int p = list[begin + blockIdx.x];
  const int xp = p % W, yp = p / W;
  const int side = 2 * r + 1;
  for (int t = threadIdx.x; t < side * side; t += blockDim.x) {
    const int dy = t / side - r, dx = t % side - r;
    if (dx * dx + dy * dy > r * r) continue;
    const int yy = min(max(yp + dy, 0), H), xx = min(max(xp + dx, 0), W);
    const long long idx = (long long)yy * W + xx;
    if (idx < (long long)H * W && state[idx] == UND && ranks_higher(A, p, (int)idx)) state[idx] = SUP;
  }
}

// ---- flat-delta variant: the reference's 3-D NMS (algorithms.py:66-103) suppresses FLAT indices i + delta with no bounds
// handling (neighbours wrap across rows / slices).  The delta set is symmetric, so "p suppresses q" == (q - p) in deltas.
__global__ void nms_flat_pick_kernel(const float* __restrict__ A, const unsigned char* __restrict__ state, long long n,
                                     const int* __restrict__ deltas, int nd, int* __restrict__ list,
                                     int* __restrict__ counters, int max_picks) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n || state[q] != UND) return;
  const float aq = A[q];
  bool blocked = false;
  for (int t = 0; t < nd; ++t) {
    const long long p = q + deltas[t];
    if (p < 0 || p >= n || p == q) continue;
    if (state[p] != UND) continue;
    const float ap = A[p];
    if (ap > aq || (ap == aq && p > q)) { blocked = true; break; }
  }
  if (blocked) { atomicAdd(&counters[1], 1); return; }
  const int slot = atomicAdd(&counters[0], 1);
  if (slot < max_picks) list[slot] = (int)q;
}
__global__ void nms_flat_suppress_kernel(const float* __restrict__ A, const int* __restrict__ list, int begin, long long n,
                                         const int* __restrict__ deltas, int nd, unsigned char* __restrict__ state) {
  const int p = list[begin + blockIdx.x];
  for (int t = threadIdx.x; t < nd; t += blockDim.x) {
    const long long idx = (long long)p + deltas[t];
    if (idx >= 0 && idx < n && state[idx] == UND && ranks_higher(A, p, (int)idx)) state[idx] = SUP;
  }
}
}  // namespace

extern "C" int tpz_nms_flat(const float* scores, long long n, const int* deltas, int num_deltas, float threshold,
                            unsigned char* state, int* list, int* counters, int max_picks, int* host_num_picks,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TPZ_CHECK(n > 0 && n < (1ll << 31) && num_deltas > 0, "tpz_nms_flat: bad sizes n=%lld deltas=%d", n, num_deltas);
  TPZ_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(int), stream));
  const int blocks = (int)((n + 255) / 256);
  nms_init_kernel<<<blocks, 256, 0, stream>>>(scores, (int)n, threshold, state);
  int picked = 0;
  for (int round = 0; round < 100000; ++round) {
    TPZ_CUDA(cudaMemsetAsync(counters + 1, 0, sizeof(int), stream));
    nms_flat_pick_kernel<<<blocks, 256, 0, stream>>>(scores, state, n, deltas, num_deltas, list, counters, max_picks);
    int h[2];
    TPZ_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    TPZ_CUDA(cudaStreamSynchronize(stream));
    TPZ_CHECK(h[0] <= max_picks, "tpz_nms_flat: more than max_picks=%d picks", max_picks);
    const int fresh = h[0] - picked;
    if (fresh > 0) {
      nms_commit_kernel<<<tpz_div_up(fresh, 256), 256, 0, stream>>>(list, picked, h[0], state);
      nms_flat_suppress_kernel<<<fresh, 128, 0, stream>>>(scores, list, picked, n, deltas, num_deltas, state);
    }
    picked = h[0];
    if (h[1] == 0) break;
    TPZ_CHECK(fresh > 0, "tpz_nms_flat: no progress (NaN scores?)");
  }
  TPZ_CUDA(cudaGetLastError());
  *host_num_picks = picked;
  return 0;
}

// scores: device fp32 [H][W]; state: device uint8 [H*W] scratch; list: device int32 [max_picks] (flat indices of the
// picks, unordered); counters: device int32[2] scratch.  *host_num_picks receives the pick count.  Synchronises the
// stream once per round (3-4 rounds).
extern "C" int tpz_nms2d(const float* scores, int H, int W, int r, float threshold, unsigned char* state, int* list,
                         int* counters, int max_picks, int* host_num_picks, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TPZ_CHECK(H > 0 && W > 0 && r >= 0 && (long long)H * W < (1ll << 31), "tpz_nms2d: bad geometry %dx%d r=%d", H, W, r);
  const int n = H * W;
  TPZ_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(int), stream));
  nms_init_kernel<<<tpz_div_up(n, 256), 256, 0, stream>>>(scores, n, threshold, state);
  int picked = 0;
  for (int round = 0; round < 100000; ++round) {
    TPZ_CUDA(cudaMemsetAsync(counters + 1, 0, sizeof(int), stream));
    nms_pick_kernel<<<tpz_div_up(n, 256), 256, 0, stream>>>(scores, state, H, W, r, list, counters, max_picks);
    int h[2];
    TPZ_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    TPZ_CUDA(cudaStreamSynchronize(stream));
    TPZ_CHECK(h[0] <= max_picks, "tpz_nms2d: more than max_picks=%d picks", max_picks);
    const int fresh = h[0] - picked;
    if (fresh > 0) {
      nms_commit_kernel<<<tpz_div_up(fresh, 256), 256, 0, stream>>>(list, picked, h[0], state);
      nms_suppress_kernel<<<fresh, 128, 0, stream>>>(scores, list, picked, H, W, r, state);
    }
    picked = h[0];
    if (h[1] == 0) break;             // nobody left undecided
    TPZ_CHECK(fresh > 0, "tpz_nms2d: no progress (NaN scores?)");
  }
  TPZ_CUDA(cudaGetLastError());
  *host_num_picks = picked;
  return 0;
}
