// Second-stage top-k merge: reduces L sorted candidate lists per query to one sorted top-k.
//
// Used twice: (1) per shard, over the per-CTA partial lists the scan kernel emits (L = #CTAs);
// (2) across ranks, over the all-gathered per-rank results (L = world size) — this replaces the
// 2*W gathers + concat + second torch.topk of the reference (src/index.py:135-157).
//
// One CTA per query, 8 warps.  Each warp folds its share of the lists into a register-resident
// sorted list of KP = 32*E entries with the bitonic merge step
//     C[i] = better(A[i], B[KP-1-i])   (C is bitonic and holds the top KP of A u B)
// followed by log2(KP) compare-exchange stages; the 8 warp results are folded by warp 0 through
// shared memory.  Order: score descending, id ascending on equal scores (total, deterministic).
#include "internal.h"

#include <math.h>

namespace mips {

constexpr int kMergeE = kMaxK / 32;   // 4 entries per lane -> KP = 128
constexpr int kMergeKP = 32 * kMergeE;
constexpr int kMergeWarps = 8;
constexpr int64_t kPadId = INT64_MAX;

struct Cand {
  uint32_t ord;  // order-preserving image of the fp32 score; 0 = padding
  int64_t id;
};

__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {
  return a.ord > b.ord || (a.ord == b.ord && a.id < b.id);
}

__device__ __forceinline__ Cand shfl_xor_cand(const Cand& c, int lmask) {
  Cand o;
  o.ord = __shfl_xor_sync(0xffffffffu, c.ord, lmask);
  o.id = __shfl_xor_sync(0xffffffffu, c.id, lmask);
  return o;
}

// v holds a bitonic sequence of KP entries (position i = lane*E + e); sorts it descending.
__device__ __forceinline__ void bitonic_merge_desc(Cand (&v)[kMergeE], int lane) {
#pragma unroll
  for (int j = kMergeKP >> 1; j >= 1; j >>= 1) {
    if (j < kMergeE) {
#pragma unroll
      for (int e = 0; e < kMergeE; ++e) {
        if ((e & j) == 0) {
          const Cand a = v[e], b = v[e ^ j];
          const bool ab = better(a, b);
          v[e] = ab ? a : b;
          v[e ^ j] = ab ? b : a;
        }
      }
    } else {
      const int lmask = j / kMergeE;
      const bool lower = (lane & lmask) == 0;
#pragma unroll
      for (int e = 0; e < kMergeE; ++e) {
        const Cand a = v[e];
        const Cand b = shfl_xor_cand(a, lmask);
        const bool ab = better(a, b);
        v[e] = (ab == lower) ? a : b;  // lower lane keeps the better one
      }
    }
  }
}

__device__ __forceinline__ Cand load_cand(const float* s, const int64_t* ids, int pos, int k_in) {
  Cand c;
  c.ord = 0;
  c.id = kPadId;
  if (pos < k_in) {
    const int64_t id = ids[pos];
    if (id >= 0) {
      c.ord = f32_to_ord(s[pos]);
      c.id = id;
    }
  }
  return c;
}

__global__ void __launch_bounds__(kMergeWarps * 32)
merge_topk_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int num_lists,
                  int64_t list_stride, int k_in, int k_out, float* __restrict__ out_scores,
                  int64_t* __restrict__ out_ids) {
  __shared__ uint32_t sh_ord[kMergeWarps][kMergeKP];
  __shared__ int64_t sh_id[kMergeWarps][kMergeKP];
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  Cand acc[kMergeE];
#pragma unroll
  for (int e = 0; e < kMergeE; ++e) { acc[e].ord = 0; acc[e].id = kPadId; }

  bool first = true;
  for (int l = warp; l < num_lists; l += kMergeWarps) {
    const float* s = scores + l * list_stride + static_cast<int64_t>(q) * k_in;
    const int64_t* id = ids + l * list_stride + static_cast<int64_t>(q) * k_in;
    if (first) {
#pragma unroll
      for (int e = 0; e < kMergeE; ++e) acc[e] = load_cand(s, id, lane * kMergeE + e, k_in);
      first = false;
    } else {
#pragma unroll
      for (int e = 0; e < kMergeE; ++e) {
        const Cand b = load_cand(s, id, kMergeKP - 1 - (lane * kMergeE + e), k_in);
        if (better(b, acc[e])) acc[e] = b;
      }
      bitonic_merge_desc(acc, lane);
    }
  }
#pragma unroll
  for (int e = 0; e < kMergeE; ++e) {
    sh_ord[warp][lane * kMergeE + e] = acc[e].ord;
    sh_id[warp][lane * kMergeE + e] = acc[e].id;
  }
  __syncthreads();
  if (warp != 0) return;
  for (int w = 1; w < kMergeWarps; ++w) {
#pragma unroll
    for (int e = 0; e < kMergeE; ++e) {
      const int pos = kMergeKP - 1 - (lane * kMergeE + e);
      Cand b;
      b.ord = sh_ord[w][pos];
      b.id = sh_id[w][pos];
      if (better(b, acc[e])) acc[e] = b;
    }
    bitonic_merge_desc(acc, lane);
  }
#pragma unroll
  for (int e = 0; e < kMergeE; ++e) {
    const int pos = lane * kMergeE + e;
    if (pos < k_out) {
      const bool ok = acc[e].id != kPadId;
      out_scores[static_cast<int64_t>(q) * k_out + pos] = ok ? ord_to_f32(acc[e].ord) : -INFINITY;
      out_ids[static_cast<int64_t>(q) * k_out + pos] = ok ? acc[e].id : -1;
    }
  }
}

cudaError_t launch_merge(const float* scores, const int64_t* ids, int num_lists, int64_t list_stride, int batch,
                         int k_in, int k_out, float* out_scores, int64_t* out_ids, cudaStream_t st) {
  if (batch == 0) return cudaSuccess;
  merge_topk_kernel<<<batch, kMergeWarps * 32, 0, st>>>(scores, ids, num_lists, list_stride, k_in, k_out, out_scores,
                                                       out_ids);
  return cudaGetLastError();
}

}  // namespace mips
