// Fused exact-MIPS scan for sm_100a: score GEMM on tcgen05 tensor cores + in-kernel top-k select.
//
// Replaces  scores = torch.matmul(allqueries.half(), self.embeddings); torch.topk(scores, k)
// (reference src/index.py:118-119) without ever writing the [batch, n_local] score matrix.
//
// One persistent CTA per SM.  Passage rows (K-major, [n_local, dim]) are streamed from HBM exactly
// once per query pass by TMA into a multi-stage shared-memory ring; the <=64 queries of the pass
// stay resident in shared memory as the B operand.  Per 128-passage tile the MMA warp issues
// dim/16 tcgen05.mma (M=128 passages x N=64 queries x K=16) into one of two TMEM accumulator
// buffers; four epilogue warps read the accumulators back (tcgen05.ld, one passage per thread),
// compare against per-query running thresholds held in shared memory and append the rare
// survivors to small L2-resident candidate lists.  When a list nears capacity one warp bitonic-
// sorts it in registers, keeps the best k and raises the query's threshold.  At the end each CTA
// emits its sorted top-k per query; merge.cu reduces the per-CTA lists.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue/select (TMEM lane quarter = warp_id % 4).
#include "internal.h"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace mips {

// ------------------------------------------------------------------------------------------------
// Warp-level exact selection over 32*E candidates held E per lane (no sorting).
// ------------------------------------------------------------------------------------------------
// k-th largest of the 32*E unsigned values v (0 = "not a candidate", never selected): MSB-first
// bisection with one warp-wide population count per bit.  Requires >= kk non-zero values.
template <int E>
__device__ __forceinline__ uint32_t warp_kth_largest(const uint32_t (&v)[E], int kk) {
  uint32_t prefix = 0;
#pragma unroll 1
  for (int b = 31; b >= 0; --b) {
    const uint32_t cand = prefix | (1u << b);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (v[e] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= kk) prefix = cand;
  }
  return prefix;
}

__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(f32_to_ord(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - row);
}

// Keeps exactly the k best (largest-key) of the c > k candidates of one list, in place and
// unsorted, and publishes the k-th best as the query's new threshold.  Keys are unique (they
// embed the row), so "key >= k-th largest key" selects exactly k entries.  Whole warp.
__device__ __forceinline__ void compact_list(uint64_t* __restrict__ list, int c, int k, int lane,
                                             uint64_t* thrkey_s, float* thr_s, int* cnt_s, int q) {
  uint64_t key[kSortE];
  uint32_t hi[kSortE];
  const ulonglong2* src = reinterpret_cast<const ulonglong2*>(list + lane * kSortE);
#pragma unroll
  for (int e = 0; e < kSortE; e += 2) {
    const ulonglong2 v = src[e >> 1];
    const int i = lane * kSortE + e;
    key[e] = (i < c) ? v.x : 0ull;
    key[e + 1] = (i + 1 < c) ? v.y : 0ull;
  }
#pragma unroll
  for (int e = 0; e < kSortE; ++e) hi[e] = static_cast<uint32_t>(key[e] >> 32);
  // 1) k-th largest score image
  const uint32_t t_hi = warp_kth_largest<kSortE>(hi, k);
  // 2) among entries tied on the score, the (k - #greater)-th largest low word (= smallest rows)
  int gt = 0, eq = 0;
#pragma unroll
  for (int e = 0; e < kSortE; ++e) {
    gt += hi[e] > t_hi ? 1 : 0;
    eq += hi[e] == t_hi ? 1 : 0;
  }
  gt = __reduce_add_sync(0xffffffffu, gt);
  eq = __reduce_add_sync(0xffffffffu, eq);
  const int need = k - gt;  // 1 <= need <= eq
  uint32_t t_lo;
  if (need == eq) {  // common case: every tied entry is kept -> threshold is the smallest tied low word
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int e = 0; e < kSortE; ++e)
      if (hi[e] == t_hi) m = min(m, static_cast<uint32_t>(key[e]));
    t_lo = __reduce_min_sync(0xffffffffu, m);
  } else {
    uint32_t lo[kSortE];
#pragma unroll
    for (int e = 0; e < kSortE; ++e) lo[e] = hi[e] == t_hi ? static_cast<uint32_t>(key[e]) : 0u;
    t_lo = warp_kth_largest<kSortE>(lo, need);
  }
  const uint64_t thrkey = (static_cast<uint64_t>(t_hi) << 32) | t_lo;
  // 3) stream compaction of the keepers to the head of the list
  int mine = 0;
#pragma unroll
  for (int e = 0; e < kSortE; ++e) mine += key[e] >= thrkey ? 1 : 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  int pos = incl - mine;
#pragma unroll
  for (int e = 0; e < kSortE; ++e)
    if (key[e] >= thrkey) list[pos++] = key[e];
  if (lane == 0) {
    thrkey_s[q] = thrkey;
    thr_s[q] = ord_to_f32(t_hi);
    cnt_s[q] = k;
  }
}

// ------------------------------------------------------------------------------------------------
// The scan kernel
// ------------------------------------------------------------------------------------------------
#define SCORE(q) __uint_as_float((q) < 32 ? r0[(q)&31] : r1[(q)&31])

__global__ void __launch_bounds__(kScanThreads, 1)
mips_scan_kernel(const __grid_constant__ CUtensorMap tmap_e, const __grid_constant__ CUtensorMap tmap_q,
                 const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int nk = p.num_kchunks;
  const int S = p.num_stages;
  const uint32_t q_smem = base;                                // nk chunks of [64 q x 64 el]
  const uint32_t st_smem = base + nk * kQChunkBytes;           // S stages of [128 p x 64 el]
  uint8_t* ctrl = smem + nk * kQChunkBytes + S * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);          // full[8] empty[8] tfull[2] tempty[2] qfull
  uint64_t* thrkey_s = bars + 24;                              // [64]
  float* thr_s = reinterpret_cast<float*>(thrkey_s + kNQ);     // [64]
  int* cnt_s = reinterpret_cast<int*>(thr_s + kNQ);            // [64]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(cnt_s + kNQ);

  const uint32_t bar_full = ptx::smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t bar_qfull = bar_tempty + 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---------------- one-time setup ----------------
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_e);
    ptx::prefetch_tensormap(&tmap_q);
    for (int s = 0; s < kMaxStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_tfull + 8 * b, 1);
      ptx::mbar_init(bar_tempty + 8 * b, 4);  // one arrival per epilogue warp
    }
    ptx::mbar_init(bar_qfull, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kNQ) {
    const int q = threadIdx.x - 64;
    const bool live = q < p.batch;  // padded query columns never pass the filter
    // initial threshold: the k-th best score of the sampled pre-pass when there is one (valid lower
    // bound of the final k-th score), else -inf
    const float seed = p.seed ? p.seed[static_cast<size_t>(q) * p.k + (p.k - 1)] : -INFINITY;
    thr_s[q] = live ? seed : INFINITY;
    thrkey_s[q] = live ? (static_cast<uint64_t>(f32_to_ord(seed)) << 32) : ~0ull;
    cnt_s[q] = 0;
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_s), kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int first_tile = blockIdx.x;
  const int tile_step = gridDim.x;
  // optional per-CTA cycle counters (diagnostics; zeroed by the host)
  const bool want_stats = p.stats != nullptr;
  unsigned long long* my_stats = want_stats ? p.stats + static_cast<size_t>(blockIdx.x) * kNumStats : nullptr;
  const long long t_start = clock64();
  long long st_a = 0, st_b = 0, st_c = 0, st_d = 0, st_e = 0;
  int st_m = 0, st_n = 0;

  // Roles run warp-uniformly (all 32 lanes take the same path and wait on the same barriers);
  // only the asynchronous issue instructions are predicated on one elected lane.  This keeps
  // addresses and descriptors in uniform registers and the issue loops short.
  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar_qfull, nk * kQChunkBytes);
      for (int kc = 0; kc < nk; ++kc)
        ptx::tma_load_2d(&tmap_q, bar_qfull, q_smem + kc * kQChunkBytes, kc * kKChunk, p.q_row0, ptx::kEvictLast);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int t = first_tile; t < p.num_tiles; t += tile_step) {
      for (int kc = 0; kc < nk; ++kc) {
        const long long w0 = want_stats ? clock64() : 0;
        ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        if (want_stats) st_a += clock64() - w0;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, kStageBytes);
          ptx::tma_load_2d(&tmap_e, bar_full + 8 * stage, st_smem + stage * kStageBytes, kc * kKChunk, t * kTileM,
                           ptx::kEvictFirst);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
    if (want_stats && lane == 0) my_stats[kStProdWait] = st_a;
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    ptx::mbar_wait(bar_qfull, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    const bool no_mma = (p.flags & kDbgNoMma) != 0;
    for (int t = first_tile; t < p.num_tiles; t += tile_step, ++it) {
      const int buf = it & 1;
      long long w0 = want_stats ? clock64() : 0;
      ptx::mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1u);
      if (want_stats) st_b += clock64() - w0;
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * kNQ;
      for (int kc = 0; kc < nk; ++kc) {
        w0 = want_stats ? clock64() : 0;
        ptx::mbar_wait(bar_full + 8 * stage, phase);
        if (want_stats) st_a += clock64() - w0;
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          if (no_mma) {
            ptx::mbar_arrive(bar_empty + 8 * stage);
            if (kc == nk - 1) ptx::mbar_arrive(bar_tfull + 8 * buf);
          } else {
            // K-major SWIZZLE_128B descriptors; advancing K by 16 elements = +32 bytes (= +2 in the
            // 16-byte-granular start-address field) inside the 128-byte swizzle row
            const uint64_t da = ptx::make_kmajor_sw128_desc(st_smem + stage * kStageBytes);
            const uint64_t db = ptx::make_kmajor_sw128_desc(q_smem + kc * kQChunkBytes);
#pragma unroll
            for (int k4 = 0; k4 < kKChunk / kUmmaK; ++k4)
              ptx::umma_f16(d_tmem, da + 2 * k4, db + 2 * k4, p.idesc, (kc | k4) != 0 ? 1u : 0u);
            ptx::umma_commit(bar_empty + 8 * stage);  // stage reusable once these MMAs retire
            if (kc == nk - 1) ptx::umma_commit(bar_tfull + 8 * buf);
          }
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
    if (want_stats && lane == 0) { my_stats[kStMmaWaitFull] = st_a; my_stats[kStMmaWaitTmem] = st_b; }
  } else {
    // =========================== epilogue / select ===========================
    const int ew = warp - 2;        // 0..3: which 16 queries this warp compacts
    const int quarter = warp & 3;   // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint64_t* my_cand = p.cand + static_cast<size_t>(blockIdx.x) * kNQ * kCap;
    const bool no_select = (p.flags & kDbgNoSelect) != 0;
    const uint32_t cnt_addr = ptx::smem_u32(cnt_s);

    int it = 0;
    for (int t = first_tile; t < p.num_tiles; t += tile_step, ++it) {
      const int buf = it & 1;
      long long w0 = want_stats ? clock64() : 0;
      ptx::mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      if (want_stats) { const long long w1 = clock64(); st_a += w1 - w0; w0 = w1; }
      ptx::tc_fence_after();
      uint32_t r0[32], r1[32];
      ptx::tmem_ld_32x32b_x32(t_lane + buf * kNQ, r0);
      ptx::tmem_ld_32x32b_x32(t_lane + buf * kNQ + 32, r1);
      ptx::tmem_ld_wait();
      if (want_stats) { const long long w1 = clock64(); st_d += w1 - w0; w0 = w1; }
      if (no_select) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * buf);
        continue;
      }

      const int64_t row = static_cast<int64_t>(t) * kTileM + quarter * 32 + lane;
      const bool valid = row < p.n_local;
      const uint32_t row32 = static_cast<uint32_t>(row);
      bool need_compact = false;

      // ---- select ----
      // A (static, ~2 instructions per score): per-thread bitmask of the queries whose running
      //   threshold this passage reaches, OR-reduced over the warp into the set of non-empty queries.
      // B (dynamic, compact code, 4 queries per round): the score column of each non-empty query is
      //   re-read from TMEM (tcgen05.ld x1 takes a runtime column, registers cannot be indexed
      //   dynamically), the exact 64-bit key test decides score ties by row, lanes 0..3 reserve
      //   slots with ONE shared-memory atomic instruction, and the survivors are stored.
      // Keeping B out of the unrolled code keeps the per-tile loop inside the instruction cache.
      const float4* thr4 = reinterpret_cast<const float4*>(thr_s);
      const uint32_t lt_mask = (1u << lane) - 1u;
      uint32_t pm0 = 0, pm1 = 0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 th = thr4[g];
        if (__uint_as_float(r0[4 * g + 0]) >= th.x) pm0 |= 1u << (4 * g + 0);
        if (__uint_as_float(r0[4 * g + 1]) >= th.y) pm0 |= 1u << (4 * g + 1);
        if (__uint_as_float(r0[4 * g + 2]) >= th.z) pm0 |= 1u << (4 * g + 2);
        if (__uint_as_float(r0[4 * g + 3]) >= th.w) pm0 |= 1u << (4 * g + 3);
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 th = thr4[8 + g];
        if (__uint_as_float(r1[4 * g + 0]) >= th.x) pm1 |= 1u << (4 * g + 0);
        if (__uint_as_float(r1[4 * g + 1]) >= th.y) pm1 |= 1u << (4 * g + 1);
        if (__uint_as_float(r1[4 * g + 2]) >= th.z) pm1 |= 1u << (4 * g + 2);
        if (__uint_as_float(r1[4 * g + 3]) >= th.w) pm1 |= 1u << (4 * g + 3);
      }
      if (!valid) pm0 = pm1 = 0u;
      uint32_t ne0 = __reduce_or_sync(0xffffffffu, pm0);
      uint32_t ne1 = __reduce_or_sync(0xffffffffu, pm1);
#pragma unroll 1
      while ((ne0 | ne1) != 0u) {  // warp-uniform
        int qs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (ne0 != 0u) { qs[u] = __ffs(ne0) - 1; ne0 &= ne0 - 1u; }
          else if (ne1 != 0u) { qs[u] = 32 + __ffs(ne1) - 1; ne1 &= ne1 - 1u; }
          else qs[u] = -1;
        }
        uint32_t sv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) sv[u] = ptx::tmem_ld_32x32b_x1(t_lane + buf * kNQ + (qs[u] < 0 ? 0 : qs[u]));
        ptx::tmem_ld_wait();
        uint64_t kk[4];
        uint32_t m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = qs[u] < 0 ? 0 : qs[u];
          const uint32_t bits = q < 32 ? pm0 : pm1;
          const bool mine = qs[u] >= 0 && ((bits >> (q & 31)) & 1u);
          kk[u] = make_key(__uint_as_float(sv[u]), row32);
          m[u] = __ballot_sync(0xffffffffu, mine && kk[u] > thrkey_s[q]);
        }
        const uint32_t mym = lane == 0 ? m[0] : lane == 1 ? m[1] : lane == 2 ? m[2] : lane == 3 ? m[3] : 0u;
        const int myq = lane == 0 ? qs[0] : lane == 1 ? qs[1] : lane == 2 ? qs[2] : qs[3];
        int old = 0;
        if (mym != 0u) old = atomicAdd(&cnt_s[myq], __popc(mym));  // lanes 0..3, distinct addresses
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int basep = __shfl_sync(0xffffffffu, old, u);
          if ((m[u] >> lane) & 1u) {
            const int pos = basep + __popc(m[u] & lt_mask);
            my_cand[qs[u] * kCap + pos] = kk[u];
            need_compact |= pos >= kCap - kTileM;
            ++st_n;
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * buf);  // TMEM buffer may be overwritten now
      // all appends of this tile are done; lists that could overflow during the next tile are
      // compacted (one barrier with an OR-reduction tells every warp whether any list needs it)
      if (want_stats) { const long long w1 = clock64(); st_b += w1 - w0; w0 = w1; }
      const bool do_compact = ptx::named_bar_red_or(1, 128, need_compact);
      if (want_stats) { const long long w1 = clock64(); st_e += w1 - w0; w0 = w1; }
      if (do_compact) {
#pragma unroll 1
        for (int i = 0; i < kNQ / 4; ++i) {
          const int q = ew * (kNQ / 4) + i;
          const int c = cnt_s[q];
          if (c > kCap - kTileM) {
            compact_list(my_cand + q * kCap, c, p.k, lane, thrkey_s, thr_s, cnt_s, q);
            ++st_m;
          }
        }
        ptx::named_bar_sync(1, 128);
        if (want_stats) st_c += clock64() - w0;
      }
    }

    // ---------------- final: publish this CTA's candidate counts ----------------
    // The candidate lists stay where they are (L2-resident workspace); the select kernel reads
    // them directly.  Only lists longer than kEmit are first cut down to their best k.
    if (!no_select) {
#pragma unroll 1
      for (int i = 0; i < kNQ / 4; ++i) {
        const int q = ew * (kNQ / 4) + i;
        int c = cnt_s[q];
        if (c > kEmit) {
          compact_list(my_cand + q * kCap, c, p.k, lane, thrkey_s, thr_s, cnt_s, q);
          c = p.k;
          ++st_m;
        }
        if (lane == 0) p.part_cnt[static_cast<size_t>(blockIdx.x) * kNQ + q] = c;
      }
    }
    if (want_stats && lane == 0) {
      atomicAdd(&my_stats[kStEpiWaitTmem], static_cast<unsigned long long>(st_a));
      atomicAdd(&my_stats[kStEpiSelect], static_cast<unsigned long long>(st_b));
      atomicAdd(&my_stats[kStEpiCompact], static_cast<unsigned long long>(st_c));
      atomicAdd(&my_stats[kStNumCompact], static_cast<unsigned long long>(st_m));
      atomicAdd(&my_stats[kStEpiLd], static_cast<unsigned long long>(st_d));
      atomicAdd(&my_stats[kStEpiBar], static_cast<unsigned long long>(st_e));
    }
    if (want_stats) atomicAdd(&my_stats[kStNumAppend], static_cast<unsigned long long>(st_n));
  }

  // ---------------- teardown ----------------
  if (want_stats && threadIdx.x == 0) my_stats[kStTotal] = clock64() - t_start;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}
#undef SCORE

cudaError_t configure_scan(size_t smem_bytes) {
  return cudaFuncSetAttribute(mips_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              static_cast<int>(smem_bytes));
}

cudaError_t launch_scan(const CUtensorMap& tmap_e, const CUtensorMap& tmap_q, const ScanParams& p, int grid,
                        size_t smem_bytes, cudaStream_t st) {
  mips_scan_kernel<<<grid, kScanThreads, smem_bytes, st>>>(tmap_e, tmap_q, p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Query preparation: cast to the index dtype (== allqueries.half(), src/index.py:118), optional
// L2 normalisation in fp32 (faiss.normalize_L2, build_server/server_start.py:142), zero padding
// of the rows [batch, batch_pad).  One CTA per output row.
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__device__ __forceinline__ float load_as_float(const TIn* p, int64_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, int64_t i) { return __half2float(p[i]); }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}

template <typename TIn>
__global__ void prep_queries_kernel(const TIn* __restrict__ q, int64_t q_ld, int batch, int dim, int out_dtype,
                                    int normalize, void* __restrict__ out) {
  const int row = blockIdx.x;
  __shared__ float red[32];
  float scale = 1.0f;
  const bool live = row < batch;
  if (live && normalize) {
    float ss = 0.f;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
      const float v = load_as_float<TIn>(q, row * q_ld + c);
      ss += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const float nrm = sqrtf(red[0]);
    scale = nrm > 0.f ? 1.0f / nrm : 1.0f;  // zero rows stay zero, like faiss
  }
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float v = 0.f;
    if (live) {
      v = load_as_float<TIn>(q, row * q_ld + c);
      if (normalize) v *= scale;
    }
    if (out_dtype == 0)
      reinterpret_cast<__half*>(out)[static_cast<int64_t>(row) * dim + c] = __float2half_rn(v);
    else
      reinterpret_cast<__nv_bfloat16*>(out)[static_cast<int64_t>(row) * dim + c] = __float2bfloat16_rn(v);
  }
}

cudaError_t launch_prep_queries(const void* q, int q_dtype, int64_t q_ld, int batch, int batch_pad, int dim,
                                int out_dtype, int normalize, void* out, cudaStream_t st) {
  const int threads = 256;
  switch (q_dtype) {
    case 0:
      prep_queries_kernel<__half><<<batch_pad, threads, 0, st>>>(static_cast<const __half*>(q), q_ld, batch, dim,
                                                                 out_dtype, normalize, out);
      break;
    case 1:
      prep_queries_kernel<__nv_bfloat16><<<batch_pad, threads, 0, st>>>(static_cast<const __nv_bfloat16*>(q), q_ld,
                                                                        batch, dim, out_dtype, normalize, out);
      break;
    default:
      prep_queries_kernel<float><<<batch_pad, threads, 0, st>>>(static_cast<const float*>(q), q_ld, batch, dim,
                                                                out_dtype, normalize, out);
      break;
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Row gather for the 3-tuple search_knn variant (build_server/index.py:228-229).
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const uint16_t* __restrict__ emb, int64_t ld, int dim, int64_t n_local,
                                   const int64_t* __restrict__ rows, int64_t n, uint16_t* __restrict__ out) {
  const int64_t i = blockIdx.x;
  if (i >= n) return;
  const int64_t r = rows[i];
  const bool ok = r >= 0 && r < n_local;
  const int vec = dim / 8;  // dim % 64 == 0 -> 16-byte vectors
  const uint4* src = reinterpret_cast<const uint4*>(emb + r * ld);
  uint4* dst = reinterpret_cast<uint4*>(out + i * dim);
  for (int c = threadIdx.x; c < vec; c += blockDim.x) dst[c] = ok ? src[c] : make_uint4(0, 0, 0, 0);
}

cudaError_t launch_gather_rows(const void* emb, int64_t ld, int dim, int64_t n_local, const int64_t* rows, int64_t n,
                               void* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  gather_rows_kernel<<<static_cast<unsigned>(n), 128, 0, st>>>(static_cast<const uint16_t*>(emb), ld, dim, n_local,
                                                               rows, n, static_cast<uint16_t*>(out));
  return cudaGetLastError();
}

}  // namespace mips
