// Selection and merge kernels that follow the fused scan.
//
//  * select_topk_kernel / select_topk_big_kernel — per shard: exact top-k of the (unsorted) candidate
//    lists the scan's CTAs leave in the workspace (k <= 128: score words in registers; k <= 1024:
//    streamed from L2).  MSB-first bisection with block-wide counts, then rank-by-counting.
//  * merge_topk_kernel — across ranks: L sorted (fp32 score, int64 id) lists per query, as produced by
//    all-gathering every rank's result, are reduced to one sorted top-k.  This replaces the 2*W
//    gathers + concat + second torch.topk of the reference (src/index.py:135-157).  One CTA per
//    query, 8 warps: each warp folds its share of the lists into a register-resident sorted list of
//    KP = 32*E entries with the bitonic merge step  C[i] = better(A[i], B[KP-1-i])  (C is bitonic
//    and holds the top KP of A u B) followed by log2(KP) compare-exchange stages; the 8 warp
//    results are folded by warp 0 through shared memory.
// Order everywhere: score descending, id ascending on equal scores (total, deterministic).
#include "internal.h"

#include <math.h>

namespace mips {

__device__ __forceinline__ void ptx_griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ptx_griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

// v holds a bitonic sequence of KP = 32*E entries (position i = lane*E + e); sorts it descending.
template <int kMergeE>
__device__ __forceinline__ void bitonic_merge_desc(Cand (&v)[kMergeE], int lane) {
  constexpr int kMergeKP = 32 * kMergeE;
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

template <int kMergeE>
__device__ __forceinline__ void merge_lists(const float* scores, const int64_t* ids, int num_lists, int64_t list_stride,
                                            int64_t id_list_stride, int k_in, int k_out, float* __restrict__ out_scores,
                                            int64_t* __restrict__ out_ids) {
  constexpr int kMergeKP = 32 * kMergeE;
  extern __shared__ __align__(16) uint8_t merge_smem[];
  int64_t (*sh_id)[kMergeKP] = reinterpret_cast<int64_t (*)[kMergeKP]>(merge_smem);
  uint32_t (*sh_ord)[kMergeKP] = reinterpret_cast<uint32_t (*)[kMergeKP]>(merge_smem + sizeof(int64_t) * kMergeWarps * kMergeKP);
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  Cand acc[kMergeE];
#pragma unroll
  for (int e = 0; e < kMergeE; ++e) { acc[e].ord = 0; acc[e].id = kPadId; }

  bool first = true;
  for (int l = warp; l < num_lists; l += kMergeWarps) {
    const float* s = scores + l * list_stride + static_cast<int64_t>(q) * k_in;
    const int64_t* id = ids + l * id_list_stride + static_cast<int64_t>(q) * k_in;
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
      bitonic_merge_desc<kMergeE>(acc, lane);
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
    bitonic_merge_desc<kMergeE>(acc, lane);
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

template <int kMergeE>
__global__ void __launch_bounds__(kMergeWarps * 32)
merge_topk_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int num_lists,
                  int64_t list_stride, int64_t id_list_stride, int k_in, int k_out, float* __restrict__ out_scores,
                  int64_t* __restrict__ out_ids) {
  merge_lists<kMergeE>(scores, ids, num_lists, list_stride, id_list_stride, k_in, k_out, out_scores, out_ids);
}

// Fused exchange + merge, receiving half: the W candidate blocks of this step were (or are being) stored into
// this GPU's exchange slot by the peers' xchg_push_kernel over NVLink.  Every CTA waits until all W arrival flags
// carry this step's epoch, then merges straight out of the slot.  The wait is bounded in wall-clock time (host-set,
// minutes by default): a peer that never pushes makes this kernel write padding and raise the exchange's error word
// (the next call on the exchange fails with MIPS_ETIMEOUT); nothing traps, the context stays alive.
template <int kMergeE>
__global__ void __launch_bounds__(kMergeWarps * 32)
xchg_merge_kernel(uint8_t* local_base, int world, size_t cap, size_t s_bytes, int k_in, int k_out,
                  float* __restrict__ out_scores, int64_t* __restrict__ out_ids, unsigned long long timeout_ns,
                  int* err_word) {
  XchgCtrl* ctrl = reinterpret_cast<XchgCtrl*>(local_base);
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch);   // set by our own push
  const int slot = static_cast<int>((epoch - 1) & 1);
  if (!xchg_wait_flags(ctrl, slot, epoch, world, timeout_ns, err_word)) {
    for (int pos = threadIdx.x; pos < k_out; pos += blockDim.x) {
      out_scores[static_cast<int64_t>(blockIdx.x) * k_out + pos] = -INFINITY;
      out_ids[static_cast<int64_t>(blockIdx.x) * k_out + pos] = -1;
    }
    return;
  }
  const uint8_t* slot_base = local_base + kXchgCtrlBytes + static_cast<size_t>(slot) * world * cap;
  merge_lists<kMergeE>(reinterpret_cast<const float*>(slot_base), reinterpret_cast<const int64_t*>(slot_base + s_bytes), world,
                       static_cast<int64_t>(cap / 4), static_cast<int64_t>(cap / 8), k_in, k_out, out_scores, out_ids);
}

// ------------------------------------------------------------------------------------------------
// Per-shard select: exact top-k of the (unsorted) candidate lists the scan kernel leaves behind.
//
// One CTA of 1024 threads per query.  Every CTA of the scan contributes <= kEmit packed keys
// (orderable score << 32 | ~row, unique).  The score words are held in registers (kSelKPT per
// thread); the k-th largest is found exactly by MSB-first bisection with block-wide counts; if
// several candidates tie on that boundary score a second bisection over their row words decides
// (smaller row wins).  The k winners are sorted by one warp and written as (fp32 score, int64 id).
// ------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelListsPerWarp = 5;                  // 32 warps * 5 lists = 160 >= 148 CTAs
constexpr int kSelChunks = kEmit / 32;               // 8 keys of one list per lane
constexpr int kSelKPT = kSelListsPerWarp * kSelChunks;  // 40 score words per thread
constexpr int kSelMaxLists = kSelWarps * kSelListsPerWarp;

template <int E>
__device__ __forceinline__ void warp_sort_desc_u64(uint64_t (&key)[E], int lane) {
  constexpr int N = 32 * E;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j < E) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & j) == 0) {
            const int i = lane * E + e;
            const bool desc = (i & k) == 0;
            const uint64_t a = key[e], b = key[e ^ j];
            const uint64_t hi = a > b ? a : b, lo = a > b ? b : a;
            key[e] = desc ? hi : lo;
            key[e ^ j] = desc ? lo : hi;
          }
        }
      } else {
        const int lmask = j / E;
        const bool lower = (lane & lmask) == 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = lane * E + e;
          const bool desc = (i & k) == 0;
          const uint64_t a = key[e];
          const uint64_t b = __shfl_xor_sync(0xffffffffu, a, lmask);
          const bool keep_max = (desc == lower);
          key[e] = keep_max ? (a > b ? a : b) : (a > b ? b : a);
        }
      }
    }
  }
}

// Block-wide MSB-first bisection over the non-zero 32-bit values v[NV] (NV per thread).
// Invariant: at least kk values are >= prefix.  `prefix` enters with the bits already known,
// `bit` is the first undecided bit (left at the next undecided bit, -1 when none remain), `n_ge`
// the number of values >= prefix.  Stops as soon as no more than `stop_at` values remain >= prefix,
// or when all bits are decided (prefix is then exactly the kk-th largest value).
template <int NV>
__device__ __forceinline__ uint32_t block_bisect(const uint32_t (&v)[NV], int kk, uint32_t prefix, int& bit, int& n_ge,
                                                 int stop_at, int* s_cnt /*[32]*/) {
  if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll 1
  while (bit >= 0 && n_ge > stop_at) {
    const uint32_t cand = prefix | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j) c += v[j] >= cand ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt[bit], c);
    __syncthreads();
    const int tot = s_cnt[bit];
    if (tot >= kk) { prefix = cand; n_ge = tot; }
    --bit;
  }
  __syncthreads();
  return prefix;
}

constexpr int kSelStage2 = kSelThreads;  // survivors of phase 1 are re-bisected one per thread

__global__ void __launch_bounds__(kSelThreads, 1)
select_topk_kernel(const uint64_t* __restrict__ cand, const int* __restrict__ part_cnt, int num_lists, int nblk, int k,
                   int64_t id_base, int64_t id_stride, float* __restrict__ out_scores,
                   int64_t* __restrict__ out_ids) {
  __shared__ int s_cnt[32];
  __shared__ int s_misc[6];       // [0] #greater, [2] winner slots, [3] total candidates, [4] stage-2 slots
  __shared__ uint32_t s_bits[2];  // AND / OR of all candidate score words
  __shared__ uint64_t s_win[kSmallK];
  __shared__ uint64_t s_key[kSelStage2];
  ptx_griddep_wait();                // the candidate lists come from the scan kernel just before
  ptx_griddep_launch_dependents();
  const int q = blockIdx.x;
  // query q of the launch lives in query block q / kNQ; its lists are those of CTAs l * nblk + q / kNQ
  const int qblk = q / kNQ, ql = q % kNQ;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x < kSmallK) s_win[threadIdx.x] = 0;
  if (threadIdx.x < 6) s_misc[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_bits[0] = 0xFFFFFFFFu; s_bits[1] = 0u; }
  s_key[threadIdx.x] = 0;
  __syncthreads();

  // warp w owns lists w, w+32, ...; lane owns positions lane, lane+32, ... of each (coalesced)
  const uint64_t* lptr[kSelListsPerWarp];
  uint32_t v[kSelKPT];
  uint32_t and_v = 0xFFFFFFFFu, or_v = 0u;
  int mine = 0;
#pragma unroll
  for (int i = 0; i < kSelListsPerWarp; ++i) {
    const int l = warp + i * kSelWarps;
    const size_t cta = static_cast<size_t>(l < num_lists ? l : 0) * nblk + qblk;
    int c = l < num_lists ? part_cnt[cta * kNQ + ql] : 0;
    c = c > kEmit ? kEmit : c;
    lptr[i] = cand + (cta * kNQ + ql) * kCap;
#pragma unroll
    for (int ch = 0; ch < kSelChunks; ++ch) {
      const int pos = lane + 32 * ch;
      uint32_t hi = 0u;
      if (pos < c) {
        hi = static_cast<uint32_t>(lptr[i][pos] >> 32);
        and_v &= hi;
        or_v |= hi;
        ++mine;
      }
      v[i * kSelChunks + ch] = hi;
    }
  }
  and_v = __reduce_and_sync(0xffffffffu, and_v);
  or_v = __reduce_or_sync(0xffffffffu, or_v);
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0) {
    atomicAnd(&s_bits[0], and_v);
    atomicOr(&s_bits[1], or_v);
    if (mine) atomicAdd(&s_misc[3], mine);
  }
  __syncthreads();
  const int total = s_misc[3];

  // ---- phase 1 (all candidates, 40 score words per thread): raise the threshold until at most
  //      kSelStage2 candidates remain above it ----
  uint32_t t_hi = 0;
  int n_ge = total;
  int bit = -1;
  if (total > kSmallK) {  // block-uniform
    const uint32_t diff = s_bits[0] ^ s_bits[1];
    if (diff == 0u) {
      t_hi = s_bits[1];  // every candidate has the same score word
    } else {
      bit = 31 - __clz(diff);
      const uint32_t common = bit == 31 ? 0u : (s_bits[1] & ~((2u << bit) - 1u));
      t_hi = block_bisect<kSelKPT>(v, k, common, bit, n_ge, kSelStage2, s_cnt);
    }
  }
  if (n_ge <= kSelStage2) {
    // ---- phase 2 (<= 1024 survivors, one per thread): full keys to shared memory, keep bisecting
    //      until they fit the final sort ----
#pragma unroll
    for (int i = 0; i < kSelListsPerWarp; ++i) {
#pragma unroll
      for (int ch = 0; ch < kSelChunks; ++ch) {
        const uint32_t hv = v[i * kSelChunks + ch];
        if (hv != 0u && hv >= t_hi) s_key[atomicAdd(&s_misc[4], 1)] = lptr[i][lane + 32 * ch];
      }
    }
    __syncthreads();
    const uint64_t my_key = s_key[threadIdx.x];   // 0 beyond the survivors
    uint32_t w1[1] = {static_cast<uint32_t>(my_key >> 32)};
    if (n_ge > kSmallK) t_hi = block_bisect<1>(w1, k, t_hi, bit, n_ge, kSmallK, s_cnt);
    if (n_ge <= kSmallK) {
      // the <= kSmallK survivors go to the final sort, which orders full 64-bit keys (score, then row)
      // and therefore resolves ties exactly
      if (my_key != 0ull && w1[0] >= t_hi) s_win[atomicAdd(&s_misc[2], 1)] = my_key;
    } else {
      // > kSmallK candidates share the exact k-th largest score word: everything above wins, and a
      // bisection over the row words of the tied candidates keeps the `need` smallest rows
      const bool gtr = w1[0] > t_hi;
      const int gt = __syncthreads_count(gtr);
      if (gtr) s_win[atomicAdd(&s_misc[2], 1)] = my_key;
      uint32_t lo1[1] = {(my_key != 0ull && w1[0] == t_hi) ? static_cast<uint32_t>(my_key) : 0u};
      int n_tie = n_ge - gt, b2 = 31;
      const uint32_t t_lo = block_bisect<1>(lo1, k - gt, 0u, b2, n_tie, 0, s_cnt);
      if (lo1[0] != 0u && lo1[0] >= t_lo) s_win[atomicAdd(&s_misc[2], 1)] = my_key;
    }
  } else {
    // ---- more than kSelStage2 candidates share the exact k-th largest score word t_hi (massive
    //      duplicates): same tie rule, in the 40-words-per-thread domain ----
    int gt = 0;
#pragma unroll
    for (int j = 0; j < kSelKPT; ++j) gt += v[j] > t_hi ? 1 : 0;
    gt = __reduce_add_sync(0xffffffffu, gt);
    if (lane == 0 && gt) atomicAdd(&s_misc[0], gt);
#pragma unroll
    for (int i = 0; i < kSelListsPerWarp; ++i) {
#pragma unroll
      for (int ch = 0; ch < kSelChunks; ++ch) {
        const int j = i * kSelChunks + ch;
        if (v[j] > t_hi) {
          const int slot = atomicAdd(&s_misc[2], 1);
          if (slot < kSmallK) s_win[slot] = lptr[i][lane + 32 * ch];
        }
        const bool tie = v[j] == t_hi && v[j] != 0u;
        v[j] = tie ? static_cast<uint32_t>(lptr[i][lane + 32 * ch]) : 0u;
      }
    }
    __syncthreads();
    const int need = k - s_misc[0];  // >= 1
    int n_tie = n_ge - s_misc[0], b2 = 31;
    const uint32_t t_lo = block_bisect<kSelKPT>(v, need, 0u, b2, n_tie, 0, s_cnt);  // exact need-th largest row word
#pragma unroll
    for (int j = 0; j < kSelKPT; ++j) {
      if (v[j] != 0u && v[j] >= t_lo) {
        const int slot = atomicAdd(&s_misc[2], 1);
        if (slot < kSmallK) s_win[slot] = (static_cast<uint64_t>(t_hi) << 32) | v[j];
      }
    }
  }
  __syncthreads();
  // ---- final order: rank by counting (128 threads x 128 broadcast reads beat a one-warp sorting
  //      network by ~10x here); keys are unique, empty slots are 0 ----
  if (threadIdx.x >= kSmallK) return;
  const uint64_t mykey = s_win[threadIdx.x];
  int rank = 0, nvalid = 0;
#pragma unroll 8
  for (int j = 0; j < kSmallK; ++j) {
    const uint64_t o = s_win[j];
    rank += o > mykey ? 1 : 0;
    nvalid += o != 0ull ? 1 : 0;
  }
  if (mykey != 0ull && rank < k) {
    const uint32_t r = 0xFFFFFFFFu - static_cast<uint32_t>(mykey);
    out_scores[static_cast<int64_t>(q) * k + rank] = ord_to_f32(static_cast<uint32_t>(mykey >> 32));
    out_ids[static_cast<int64_t>(q) * k + rank] = id_base + static_cast<int64_t>(r) * id_stride;
  }
  const int p = threadIdx.x;  // padding for queries with fewer than k candidates
  if (p >= nvalid && p < k) {
    out_scores[static_cast<int64_t>(q) * k + p] = -INFINITY;
    out_ids[static_cast<int64_t>(q) * k + p] = -1;
  }
}

// ------------------------------------------------------------------------------------------------
// Big-k select (128 < k <= 1024, lists of up to kCapBig entries): same bisection, but the score
// words are re-read from the L2-resident lists on every step instead of living in registers; once
// at most kBigStage (2048) candidates remain they are ranked by counting in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kBigStage = 2048;

__global__ void __launch_bounds__(kSelThreads, 1)
select_topk_big_kernel(const uint64_t* __restrict__ cand, const int* __restrict__ part_cnt, int num_lists, int nblk,
                       int cap, int k, int64_t id_base, int64_t id_stride, float* __restrict__ out_scores,
                       int64_t* __restrict__ out_ids) {
  __shared__ int s_misc[4];       // [0] scratch count, [2] survivor slots
  __shared__ uint32_t s_bits[2];
  __shared__ int s_len[kSelMaxLists];
  __shared__ uint64_t s_key[kBigStage];
  ptx_griddep_wait();                // the candidate lists come from the scan kernel just before
  ptx_griddep_launch_dependents();
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kBigStage; i += kSelThreads) s_key[i] = 0;
  if (threadIdx.x < 4) s_misc[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_bits[0] = 0xFFFFFFFFu; s_bits[1] = 0u; }
  if (threadIdx.x < kSelMaxLists) {
    int c = threadIdx.x < num_lists ? part_cnt[(static_cast<size_t>(threadIdx.x) * nblk + q / kNQ) * kNQ + q % kNQ] : 0;
    s_len[threadIdx.x] = c > cap ? cap : c;
  }
  __syncthreads();

  // one pass over every candidate of this query: warp w takes lists w, w+32, ...; f(hi, lo) per entry
  auto for_each = [&](auto&& f) {
    for (int l = warp; l < num_lists; l += kSelWarps) {
      const uint64_t* lst = cand + ((static_cast<size_t>(l) * nblk + q / kNQ) * kNQ + q % kNQ) * cap;
      const int c = s_len[l];
      for (int i = lane; i < c; i += 32) f(lst[i]);
    }
  };
  auto block_sum = [&](int v) -> int {   // all threads must call
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if (threadIdx.x == 0) s_misc[0] = 0;
    __syncthreads();
    if (lane == 0 && v) atomicAdd(&s_misc[0], v);
    __syncthreads();
    return s_misc[0];
  };

  uint32_t and_v = 0xFFFFFFFFu, or_v = 0u;
  int mine = 0;
  for_each([&](uint64_t kk) { const uint32_t h = static_cast<uint32_t>(kk >> 32); and_v &= h; or_v |= h; ++mine; });
  and_v = __reduce_and_sync(0xffffffffu, and_v);
  or_v = __reduce_or_sync(0xffffffffu, or_v);
  if (lane == 0) { atomicAnd(&s_bits[0], and_v); atomicOr(&s_bits[1], or_v); }
  const int total = block_sum(mine);

  uint32_t t_hi = 0;
  int n_ge = total;
  if (total > kBigStage) {
    const uint32_t diff = s_bits[0] ^ s_bits[1];
    int bit = diff == 0u ? -1 : 31 - __clz(diff);
    t_hi = (diff == 0u || bit == 31) ? (diff == 0u ? s_bits[1] : 0u) : (s_bits[1] & ~((2u << bit) - 1u));
    while (bit >= 0 && n_ge > kBigStage) {
      const uint32_t c0 = t_hi | (1u << bit);
      int n = 0;
      for_each([&](uint64_t kk) { n += static_cast<uint32_t>(kk >> 32) >= c0 ? 1 : 0; });
      const int tot = block_sum(n);
      if (tot >= k) { t_hi = c0; n_ge = tot; }
      --bit;
    }
  }
  uint64_t thrkey = static_cast<uint64_t>(t_hi) << 32;
  if (n_ge > kBigStage) {
    // more than kBigStage candidates share the exact k-th largest score word: bisect the row words of the ties
    int n = 0;
    for_each([&](uint64_t kk) { n += static_cast<uint32_t>(kk >> 32) > t_hi ? 1 : 0; });
    const int need = k - block_sum(n);
    uint32_t t_lo = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t c0 = t_lo | (1u << bit);
      int m2 = 0;
      for_each([&](uint64_t kk) {
        m2 += (static_cast<uint32_t>(kk >> 32) == t_hi && static_cast<uint32_t>(kk) >= c0) ? 1 : 0;
      });
      if (block_sum(m2) >= need) t_lo = c0;
    }
    thrkey |= t_lo;   // now exactly k candidates have key >= thrkey
  }
  for_each([&](uint64_t kk) {
    if (kk >= thrkey) {
      const int slot = atomicAdd(&s_misc[2], 1);
      if (slot < kBigStage) s_key[slot] = kk;
    }
  });
  __syncthreads();
  // rank by counting: 2 keys per thread against all survivors (broadcast reads); keys are unique
  const int nsurv = s_misc[2] < kBigStage ? s_misc[2] : kBigStage;
  const uint64_t k0 = s_key[threadIdx.x], k1 = s_key[threadIdx.x + kSelThreads];
  int r0 = 0, r1 = 0;
  for (int j = 0; j < nsurv; ++j) {
    const uint64_t o = s_key[j];
    r0 += o > k0 ? 1 : 0;
    r1 += o > k1 ? 1 : 0;
  }
  auto emit = [&](uint64_t kk, int rank) {
    if (kk != 0ull && rank < k) {
      const uint32_t r = 0xFFFFFFFFu - static_cast<uint32_t>(kk);
      out_scores[static_cast<int64_t>(q) * k + rank] = ord_to_f32(static_cast<uint32_t>(kk >> 32));
      out_ids[static_cast<int64_t>(q) * k + rank] = id_base + static_cast<int64_t>(r) * id_stride;
    }
  };
  emit(k0, r0);
  emit(k1, r1);
  for (int p = nsurv + threadIdx.x; p < k; p += kSelThreads) {   // fewer than k candidates: pad
    out_scores[static_cast<int64_t>(q) * k + p] = -INFINITY;
    out_ids[static_cast<int64_t>(q) * k + p] = -1;
  }
}

cudaError_t launch_select(const uint64_t* cand, const int* part_cnt, int num_lists, int nblk, int cap, int batch, int k,
                          int64_t id_base, int64_t id_stride, float* out_scores, int64_t* out_ids, cudaStream_t st) {
  if (batch == 0) return cudaSuccess;
  if (num_lists > kSelMaxLists) return cudaErrorInvalidValue;
  if (cap == kCap && k <= kSmallK)
    return launch_pdl(select_topk_kernel, dim3(batch), dim3(kSelThreads), 0, st, g_use_pdl, cand, part_cnt, num_lists, nblk,
                      k, id_base, id_stride, out_scores, out_ids);
  return launch_pdl(select_topk_big_kernel, dim3(batch), dim3(kSelThreads), 0, st, g_use_pdl, cand, part_cnt, num_lists,
                    nblk, cap, k, id_base, id_stride, out_scores, out_ids);
}

template <int E>
static cudaError_t launch_merge_e(const float* scores, const int64_t* ids, int num_lists, int64_t list_stride,
                                  int64_t id_list_stride, int batch, int k_in, int k_out, float* out_scores,
                                  int64_t* out_ids, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(kMergeWarps) * 32 * E * (sizeof(int64_t) + sizeof(uint32_t));
  merge_topk_kernel<E><<<batch, kMergeWarps * 32, smem, st>>>(scores, ids, num_lists, list_stride, id_list_stride, k_in,
                                                             k_out, out_scores, out_ids);
  return cudaGetLastError();
}

cudaError_t configure_merge() {
  const int smem = kMergeWarps * kMaxK * static_cast<int>(sizeof(int64_t) + sizeof(uint32_t));
  cudaError_t e = cudaFuncSetAttribute(merge_topk_kernel<kMaxK / 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(xchg_merge_kernel<kMaxK / 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

cudaError_t launch_xchg_merge(uint8_t* local_base, int world, size_t cap, size_t s_bytes, int batch, int k_in, int k_out,
                              float* out_scores, int64_t* out_ids, unsigned long long timeout_ns, int* err_word,
                              cudaStream_t st) {
  if (batch == 0) return cudaSuccess;
  const int kk = k_in > k_out ? k_in : k_out;
  if (kk <= kSmallK) {
    constexpr int E = kSmallK / 32;
    const size_t smem = static_cast<size_t>(kMergeWarps) * 32 * E * (sizeof(int64_t) + sizeof(uint32_t));
    xchg_merge_kernel<E><<<batch, kMergeWarps * 32, smem, st>>>(local_base, world, cap, s_bytes, k_in, k_out, out_scores, out_ids,
                                                                timeout_ns, err_word);
    return cudaGetLastError();
  }
  static cudaError_t cfg = configure_merge();
  if (cfg != cudaSuccess) return cfg;
  constexpr int E = kMaxK / 32;
  const size_t smem = static_cast<size_t>(kMergeWarps) * 32 * E * (sizeof(int64_t) + sizeof(uint32_t));
  xchg_merge_kernel<E><<<batch, kMergeWarps * 32, smem, st>>>(local_base, world, cap, s_bytes, k_in, k_out, out_scores, out_ids,
                                                                timeout_ns, err_word);
  return cudaGetLastError();
}

cudaError_t launch_merge(const float* scores, const int64_t* ids, int num_lists, int64_t list_stride,
                         int64_t id_list_stride, int batch, int k_in, int k_out, float* out_scores, int64_t* out_ids,
                         cudaStream_t st) {
  if (batch == 0) return cudaSuccess;
  const int kk = k_in > k_out ? k_in : k_out;
  if (kk <= kSmallK)
    return launch_merge_e<kSmallK / 32>(scores, ids, num_lists, list_stride, id_list_stride, batch, k_in, k_out,
                                        out_scores, out_ids, st);
  static cudaError_t cfg = configure_merge();
  if (cfg != cudaSuccess) return cfg;
  return launch_merge_e<kMaxK / 32>(scores, ids, num_lists, list_stride, id_list_stride, batch, k_in, k_out, out_scores,
                                    out_ids, st);
}

}  // namespace mips
