// Fused exact-MIPS scan for sm_100a: score GEMM on tcgen05 tensor cores + in-kernel top-k select.
//
// Replaces  scores = torch.matmul(allqueries.half(), self.embeddings); torch.topk(scores, k)
// (reference src/index.py:118-119) without ever writing the [batch, n_local] score matrix.
//
// Query-stationary design, one persistent CTA per SM:
//  * the <=128 queries of a pass are the A operand and live in TENSOR MEMORY for the whole kernel
//    (128 lanes x dim/2 columns; up to 768 dims — a longer K tail stays in shared memory);
//  * passage rows (K-major, [n_local, dim]) are streamed from HBM exactly once per pass by TMA into
//    a shared-memory ring and are the B operand; each byte of the index is read from shared memory
//    once (the smem->tensor-core operand path, ~64 B/clk/SM, is the second-tightest resource);
//  * per 64-passage tile the MMA warp issues dim/16 tcgen05.mma (M=128 queries x N=64 passages x
//    K=16) into one of two TMEM accumulator buffers;
//  * four epilogue warps own 32 queries each — one thread per query: tcgen05.ld brings the 64 scores
//    of the tile, each is compared with the query's running threshold (a register), and the rare
//    survivors are appended to the query's candidate list (L2-resident, per CTA).  No atomics, no
//    cross-warp synchronisation.  When a list nears capacity its warp cuts it back to the best k
//    (exact bisection select, no sort) and raises the threshold.
// The per-CTA lists are reduced by select_topk_kernel (merge.cu).
//
// Three instances of one body: mips_scan_kernel ([n, dim] storage, <= 128 queries per CTA), mips_scan_dn_kernel
// ([dim, n] storage, MN-major B operand) and mips_scan_pair_kernel (batches > 128: the two CTAs of a cluster form a
// tcgen05 CTA pair, UMMA M = 256, each CTA loads half of every passage tile; see scan_body).
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
// unsorted, and returns the k-th best key (the list owner's new threshold).  Keys are unique (they
// embed the row), so "key >= k-th largest key" selects exactly k entries.  Whole warp, converged.
__device__ __noinline__ uint64_t compact_list(uint64_t* __restrict__ list, int c, int k, int lane) {
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
  __syncwarp();
  return thrkey;
}

// Same contract for lists of any length (big-k mode): nothing is kept in registers, every bisection
// step re-reads the (L2-resident) list.  Slow (~32 passes) but only reached when a list overflows,
// which seeded thresholds make a rare event.  Whole warp, converged; requires c > k.
__device__ __noinline__ uint64_t compact_list_stream(uint64_t* __restrict__ list, int c, int k, int lane) {
  const uint32_t* hi_words = reinterpret_cast<const uint32_t*>(list) + 1;   // little endian: high word second
  const uint32_t* lo_words = reinterpret_cast<const uint32_t*>(list);
  uint32_t t_hi = 0;
#pragma unroll 1
  for (int b = 31; b >= 0; --b) {
    const uint32_t cand = t_hi | (1u << b);
    int n = 0;
    for (int i = lane; i < c; i += 32) n += hi_words[2 * i] >= cand ? 1 : 0;
    n = __reduce_add_sync(0xffffffffu, n);
    if (n >= k) t_hi = cand;
  }
  int gt = 0, eq = 0;
  for (int i = lane; i < c; i += 32) {
    const uint32_t h = hi_words[2 * i];
    gt += h > t_hi ? 1 : 0;
    eq += h == t_hi ? 1 : 0;
  }
  gt = __reduce_add_sync(0xffffffffu, gt);
  eq = __reduce_add_sync(0xffffffffu, eq);
  const int need = k - gt;
  uint32_t t_lo = 0;   // need == eq: every tied entry is kept
  if (need < eq) {
#pragma unroll 1
    for (int b = 31; b >= 0; --b) {
      const uint32_t cand = t_lo | (1u << b);
      int n = 0;
      for (int i = lane; i < c; i += 32) n += (hi_words[2 * i] == t_hi && lo_words[2 * i] >= cand) ? 1 : 0;
      n = __reduce_add_sync(0xffffffffu, n);
      if (n >= need) t_lo = cand;
    }
  } else {
    uint32_t m = 0xFFFFFFFFu;
    for (int i = lane; i < c; i += 32)
      if (hi_words[2 * i] == t_hi) m = min(m, lo_words[2 * i]);
    t_lo = __reduce_min_sync(0xffffffffu, m);
  }
  const uint64_t thrkey = (static_cast<uint64_t>(t_hi) << 32) | t_lo;
  // in-place stream compaction, 32 entries at a time (writes never pass the reads)
  int out = 0;
  for (int base = 0; base < c; base += 32) {
    const int i = base + lane;
    const uint64_t kk = i < c ? list[i] : 0ull;
    const bool keep = kk >= thrkey && i < c;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) list[out + __popc(m & ((1u << lane) - 1u))] = kk;
    out += __popc(m);
  }
  __syncwarp();
  return thrkey;
}

// r[c] for a runtime c in [0, 64): registers cannot be indexed dynamically, so pick through a
// 6-level select tree (63 selects; only executed for the rare survivors).
__device__ __forceinline__ uint32_t pick64(const uint32_t (&r0)[32], const uint32_t (&r1)[32], int c) {
  uint32_t t[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) t[i] = (c & 32) ? r1[i] : r0[i];
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
#pragma unroll
    for (int i = 0; i < w; ++i) t[i] = (c & w) ? t[i + w] : t[i];
  }
  return t[0];
}

// Cross-CTA hand-offs of the in-kernel sampled seeding use flag words tagged with a per-launch token
// (no counting barrier): a stale flag of an earlier launch never matches, nothing has to be reset, and
// every wait is bounded — a CTA that is not there in time (e.g. because another kernel holds its SM) is
// simply left out, which only makes the seed lower.  Seeds are lower bounds, so any subset is valid.
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr int kSeedOwnerSpins = 1024;    // ~0.15 ms: how long a query's owner waits for the CTAs' samples
constexpr int kSeedReaderSpins = 4096;   // ~0.6 ms: how long a thread waits for its query's seed

// ------------------------------------------------------------------------------------------------
// The scan kernel
// ------------------------------------------------------------------------------------------------
// kPair = false: one CTA per SM works alone (UMMA M = 64 / 128, cta_group::1).
// kPair = true : the two CTAs of a cluster (one TPC) form a tcgen05 CTA pair for 256 queries (UMMA M = 256,
//   cta_group::2).  Each CTA keeps its own 128 queries in its own tensor memory and gets its own 128 x 64
//   accumulator, but TMA-loads only HALF of every passage tile (32 of the 64 rows); the pair's MMAs, issued by the
//   leader CTA, read both halves.  Per SM that is half the shared-memory fill per unit of tensor work — the limiter
//   of the single-CTA kernel for batches >= 256 (one SM's TMA ingest, ~46 B/clk, against 96 KB per 1536 tensor
//   cycles).  Barriers: full[] lives in the leader (one arrival per CTA + both CTAs' TMA bytes), empty[] / tfull[]
//   are signalled in both CTAs by multicast tcgen05.commit, tempty[] / qready live in the leader and collect the
//   epilogue warps of both CTAs.
template <bool kPair, bool kBMn>
__device__ __forceinline__ void scan_body(const CUtensorMap& tmap_e, const CUtensorMap& tmap_q, const ScanParams& p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int nk = p.num_kchunks;                               // K chunks of 64 elements
  const int nk_ts = nk < kMaxTsChunks ? nk : kMaxTsChunks;    // ... of which the queries sit in TMEM
  const int nk_ss = nk - nk_ts;                               // ... and in shared memory (dim > 768)
  const int S = p.num_stages;
  const int cps = p.chunks_per_stage;                            // K chunks one pipeline stage carries
  constexpr int kJ = kPair ? kTopJPair : kTopJ;                  // sampled-seeding values per (CTA, query)
  constexpr int kSlots = kPair ? kSeedSlotsPair : kSeedSlots;    // ... and lists a lane of the selecting warp covers
  constexpr int kBoxRows = kPair ? kTileN / 2 : kTileN;          // passage rows this CTA loads per tile
  constexpr int kBoxBytes = kBoxRows * kKChunk * 2;              // one K chunk of them (SWIZZLE_128B)
  const int stage_bytes = cps * kBoxBytes;
  const uint32_t cta_rank = kPair ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  const int stages_per_tile = (nk + cps - 1) / cps;
  const uint32_t q_smem = base;                               // nk_ss chunks of [128 q x 64 el]
  const uint32_t st_smem = base + nk_ss * kQChunkBytes;       // S stages of 2 x [64 p x 64 el]
  uint8_t* ctrl = smem + nk_ss * kQChunkBytes + S * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);         // full[12] empty[12] tfull[2] tempty[2] qfull qready
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 32);

  const uint32_t bar_full = ptx::smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t bar_qfull = bar_tempty + 16;
  const uint32_t bar_qready = bar_qfull + 8;                  // pair mode: both CTAs' queries are in tensor memory

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---------------- one-time setup ----------------
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_e);
    if (nk_ss > 0) ptx::prefetch_tensormap(&tmap_q);
    for (int s = 0; s < kMaxStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, kPair ? 2 : 1);   // pair: one arrival per CTA (the leader's carries the bytes)
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_tfull + 8 * b, 1);
      ptx::mbar_init(bar_tempty + 8 * b, kPair ? 8 : 4);  // one arrival per epilogue warp (of both CTAs)
    }
    ptx::mbar_init(bar_qfull, kPair ? 2 : 1);
    ptx::mbar_init(bar_qready, 8);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) { ptx::tmem_alloc_pair(ptx::smem_u32(tmem_ptr_s), kTmemCols); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_s), kTmemCols); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync_all(); else __syncthreads();   // pair: the peer's barriers are initialised too
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  // Programmatic dependent launch: let the next kernel of the stream begin its own prologue now.  Up to
  // here nothing produced by an earlier kernel was touched; the passage index and the tensor maps are
  // static, so the TMA producer may start streaming right away.  Whoever reads upstream results
  // (prepared queries, seed thresholds) or writes the shared workspace calls griddep_wait() first.
  ptx::griddep_launch_dependents();

  // Large batches: nblk query blocks (128 queries each) are scanned in the same launch.  CTA c works on
  // block c % nblk and walks the tile sequence c / nblk, c / nblk + grid / nblk, ... — the nblk CTAs that
  // share a tile read it at about the same time, so HBM delivers it once and the others hit the L2.
  const int qblk = blockIdx.x % p.nblk;
  const int first_tile = blockIdx.x / p.nblk;
  const int tile_step = gridDim.x / p.nblk;
  const int q_row0 = p.q_row0 + qblk * kNQ;
  const int blk_batch = p.batch - qblk * kNQ < kNQ ? p.batch - qblk * kNQ : kNQ;   // may be <= 0: idle block
  const uint64_t stream_hint = p.nblk > 1 ? ptx::kEvictNormal : ptx::kEvictFirst;
  // Tile schedule of this CTA: n_samp sample tiles (its own first tiles, scanned once more afterwards) followed
  // by its n_my tiles first_tile, first_tile + tile_step, ...  All three roles walk the same sequence.
  const int n_my = first_tile < p.num_tiles ? (p.num_tiles - first_tile + tile_step - 1) / tile_step : 0;
  const int n_samp = p.sample_tiles < n_my ? p.sample_tiles : n_my;
  const int n_iter = n_samp + n_my;
  // optional per-CTA cycle counters (diagnostics; zeroed by the host)
  const bool want_stats = p.stats != nullptr;
  unsigned long long* my_stats = want_stats ? p.stats + static_cast<size_t>(blockIdx.x) * kNumStats : nullptr;
  const long long t_start = clock64();
  long long st_a = 0, st_b = 0, st_c = 0, st_d = 0;
  int st_m = 0, st_n = 0;

  // Roles run warp-uniformly (all 32 lanes take the same path and wait on the same barriers);
  // only the asynchronous issue instructions are predicated on one elected lane.  This keeps
  // addresses and descriptors in uniform registers and the issue loops short.
  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (nk_ss > 0) ptx::griddep_wait();   // the K tail is read from the prepared-query buffer
    if (nk_ss > 0 && ptx::elect_one()) {  // K tail of the queries: resident in shared memory
      if (kPair) {
        const uint32_t qf = ptx::mapa(bar_qfull, 0);
        if (leader) ptx::mbar_arrive_expect_tx(bar_qfull, 2 * nk_ss * kQChunkBytes); else ptx::mbar_arrive_cluster(qf);
        for (int kc = 0; kc < nk_ss; ++kc)
          ptx::tma_load_2d_pair(&tmap_q, qf, q_smem + kc * kQChunkBytes, (nk_ts + kc) * kKChunk, q_row0, ptx::kEvictLast);
      } else {
        ptx::mbar_arrive_expect_tx(bar_qfull, nk_ss * kQChunkBytes);
        for (int kc = 0; kc < nk_ss; ++kc)
          ptx::tma_load_2d(&tmap_q, bar_qfull, q_smem + kc * kQChunkBytes, (nk_ts + kc) * kKChunk, q_row0,
                           ptx::kEvictLast);
      }
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    // lock-step of the pairs that share this tile sequence (see ScanParams::progress)
    bool lockstep = kPair && leader && p.progress != nullptr && p.nblk > 2;
    const int npb = p.nblk >> 1, my_pair = static_cast<int>(blockIdx.x >> 1);
    const int group0 = npb > 0 ? (my_pair / npb) * npb : 0;
    unsigned long long tag = 0;
    if (lockstep) {
      ptx::griddep_wait();   // the search token is bumped by the query-preparation kernel of this search
      tag = static_cast<unsigned long long>(*reinterpret_cast<const volatile uint32_t*>(p.token) * 4096u +
                                            static_cast<uint32_t>(p.scan_seq)) << 32;
    }
    for (int i = 0; i < n_iter; ++i) {
      const int t = first_tile + (i < n_samp ? i : i - n_samp) * tile_step;
      if (lockstep && i >= p.lock_window && (i % kLockEvery) == 0) {
        if (lane == 0) {
          const unsigned long long need = tag | static_cast<unsigned long long>(i - p.lock_window);
          for (int sblg = 0; sblg < npb && lockstep; ++sblg) {
            if (group0 + sblg == my_pair) continue;
            const unsigned long long* slot = p.progress + group0 + sblg;
            int polls = 0;
            // a sibling that carries another tag has not started this launch yet; one that never shows up (its SMs
            // are busy with another kernel) must not stall the scan: give up on the lock-step after ~20k polls
            for (;;) {
              const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(slot);
              if ((v >> 32) == (tag >> 32) && v >= need) break;
              if (++polls > 20000) { lockstep = false; break; }
              __nanosleep(200);
            }
          }
        }
        lockstep = __shfl_sync(0xffffffffu, lockstep ? 1 : 0, 0) != 0;
      }
      for (int si = 0; si < stages_per_tile; ++si) {
        const long long w0 = want_stats ? clock64() : 0;
        ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        if (want_stats) st_a += clock64() - w0;
        if ((p.flags & kDbgNoTma) != 0) {
          if (ptx::elect_one()) {
            if (!kPair || leader) ptx::mbar_arrive(bar_full + 8 * stage);
            else ptx::mbar_arrive_cluster(ptx::mapa(bar_full + 8 * stage, 0));
          }
        } else if (ptx::elect_one()) {
          const int kc0 = si * cps;
          const int nch = nk - kc0 < cps ? nk - kc0 : cps;
          if (kPair) {
            // this CTA's half of the tile; the bytes of both halves are counted on the leader's barrier
            const uint32_t full = ptx::mapa(bar_full + 8 * stage, 0);
            if (leader) ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * nch * kBoxBytes);
            else ptx::mbar_arrive_cluster(full);
            const int c_row = t * kTileN + static_cast<int>(cta_rank) * kBoxRows;
            for (int c = 0; c < nch; ++c)
              ptx::tma_load_2d_pair(&tmap_e, full, st_smem + stage * stage_bytes + c * kBoxBytes, (kc0 + c) * kKChunk, c_row,
                                    stream_hint);
          } else {
            ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, nch * kBoxBytes);
            for (int c = 0; c < nch; ++c) {
              // tensor-map coordinates are (inner, outer): (dim, passage) for [n, dim], (passage, dim) for [dim, n]
              const int c_dim = (kc0 + c) * kKChunk, c_row = t * kTileN;
              ptx::tma_load_2d(&tmap_e, bar_full + 8 * stage, st_smem + stage * stage_bytes + c * kBoxBytes,
                               kBMn ? c_row : c_dim, kBMn ? c_dim : c_row, stream_hint);
            }
          }
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      if (kPair && leader && p.progress != nullptr && p.nblk > 2 && lane == 0)
        *reinterpret_cast<volatile unsigned long long*>(p.progress + my_pair) = tag | static_cast<unsigned long long>(i);
    }
    if (want_stats && lane == 0) my_stats[kStProdWait] = st_a;
  } else if (warp == 1) {
   if (!kPair || leader) {
    // =========================== MMA issuer (pair: the leader CTA's only) ===========================
    if (nk_ss > 0) { if (kPair) ptx::mbar_wait_cluster(bar_qfull, 0); else ptx::mbar_wait(bar_qfull, 0); }
    // the epilogue warps (of both CTAs) have written the queries to TMEM
    if (kPair) ptx::mbar_wait_cluster(bar_qready, 0); else ptx::named_bar_sync(2, 160);
    ptx::tc_fence_after();
    const bool no_mma = !kPair && (p.flags & kDbgNoMma) != 0;
    const uint32_t idesc = p.idesc;
    // B: [rows x 64 el] boxes, SWIZZLE_128B.  K-major ([n, dim] index): rows are passages, a K=16 step advances 32 B
    // inside the swizzle row (+2 in the 16-byte-granular address field).  MN-major ([dim, n] index): rows are dims,
    // a K=16 step advances 16 rows = 2048 B (+128).
    constexpr uint32_t kBStep = kBMn ? 128u : 2u;
    // The MMAs of chunks [c0, c1) of one pipeline stage.  tcgen05.mma of this shape retires in 32 tensor cycles, so
    // the issue loop itself has to stay well below that per MMA: one asm block per K chunk (4 MMAs) whose operand
    // addresses are 32-bit adds inside the block, nothing branched on in between.
    auto issue = [&](int c0, int c1, int kc0, uint32_t b_lo0, uint32_t d_tmem) {
      int kc = kc0 + c0;
      uint32_t b_lo = b_lo0 + static_cast<uint32_t>(c0) * (kBoxBytes >> 4);
      uint32_t a_tmem = tmem_base + kc * (kKChunk / 2);             // A from TMEM: 8 columns per K=16 step
      const int kc_end = kc0 + c1, ts_end = kc_end < nk_ts ? kc_end : nk_ts;
#pragma unroll 1
      for (; kc < ts_end; ++kc, b_lo += kBoxBytes >> 4, a_tmem += kKChunk / 2)
        ptx::umma_ts_x4<kPair, kBStep>(d_tmem, a_tmem, b_lo, idesc, kc != 0 ? 1u : 0u);
      uint32_t a_lo = ptx::sw128_desc_lo(q_smem + (kc - nk_ts) * kQChunkBytes);   // K tail: A from shared memory
#pragma unroll 1
      for (; kc < kc_end; ++kc, b_lo += kBoxBytes >> 4, a_lo += kQChunkBytes >> 4)
        ptx::umma_ss_x4<kPair, kBStep>(d_tmem, a_lo, b_lo, idesc, kc != 0 ? 1u : 0u);
    };
    // A step = one pipeline stage.  What a step needs: its stage filled and, on a tile's first stage, the accumulator
    // buffer drained.  The next step's barriers are probed (non-blocking) before the current step's last chunk is
    // issued, so that in the steady state the hand-off costs the issuer nothing while MMAs are still queued.
    const int n_steps = n_iter * stages_per_tile;
    int stage = 0, si = 0, it = 0;
    uint32_t phase = 0;
    bool ready = false;
    for (int s = 0; s < n_steps; ++s) {
      const int buf = it & 1;
      if (!ready) {
        if (si == 0) {
          const long long w0 = want_stats ? clock64() : 0;
          ptx::mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1u);
          if (want_stats) st_b += clock64() - w0;
        }
        const long long w0 = want_stats ? clock64() : 0;
        ptx::mbar_wait(bar_full + 8 * stage, phase);
        if (want_stats) st_a += clock64() - w0;
      }
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + kAccCol0 + buf * kTileN;
      const bool last = si == stages_per_tile - 1;
      const int kc0 = si * cps;
      const int nch = nk - kc0 < cps ? nk - kc0 : cps;
      const uint32_t b_lo0 = ptx::sw128_desc_lo(st_smem + stage * stage_bytes);
      int nstage = stage + 1, nsi = si + 1, nit = it;
      uint32_t nphase = phase;
      if (nstage == S) { nstage = 0; nphase ^= 1u; }
      if (nsi == stages_per_tile) { nsi = 0; ++nit; }
      if (!no_mma && ptx::elect_one()) issue(0, nch - 1, kc0, b_lo0, d_tmem);
      __syncwarp();
      ready = false;
      if (s + 1 < n_steps) {
        ready = ptx::mbar_test_wait(bar_full + 8 * nstage, nphase);
        if (nsi == 0) ready = ready && ptx::mbar_test_wait(bar_tempty + 8 * (nit & 1), ((nit >> 1) & 1) ^ 1u);
        ready = __all_sync(0xffffffffu, ready);
      }
      if (ptx::elect_one()) {
        if (no_mma) {
          ptx::mbar_arrive(bar_empty + 8 * stage);
          if (last) ptx::mbar_arrive(bar_tfull + 8 * buf);
        } else {
          issue(nch - 1, nch, kc0, b_lo0, d_tmem);
          // stage reusable (in both CTAs of a pair) once these MMAs retire
          if (kPair) ptx::umma_commit_pair(bar_empty + 8 * stage); else ptx::umma_commit(bar_empty + 8 * stage);
          if (last) { if (kPair) ptx::umma_commit_pair(bar_tfull + 8 * buf); else ptx::umma_commit(bar_tfull + 8 * buf); }
        }
      }
      __syncwarp();
      stage = nstage; phase = nphase; si = nsi; it = nit;
    }
    if (want_stats && lane == 0) { my_stats[kStMmaWaitFull] = st_a; my_stats[kStMmaWaitTmem] = st_b; }
   }
  } else {
    // =========================== epilogue / select ===========================
    // UMMA M=128: query m sits in TMEM lane m.  UMMA M=64 (passes of <= 64 queries; half the tensor
    // work and power): query m sits in lane 32*(m/16) + m%16, i.e. lanes 0..15 of every lane quarter.
    const int quarter = warp & 3;           // TMEM lanes [32*quarter, +32) are accessible to this warp
    const int per_warp = p.m64 ? 16 : 32;   // queries owned by this warp
    const bool lane_ok = lane < per_warp;
    const int ql = quarter * per_warp + (lane_ok ? lane : 0);   // my query within the pass
    const bool live = lane_ok && ql < blk_batch;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int cap = p.cap;
    uint64_t* warp_lists = p.cand + (static_cast<size_t>(blockIdx.x) * kNQ + quarter * per_warp) * cap;
    uint64_t* my_list = warp_lists + static_cast<size_t>(lane_ok ? lane : 0) * cap;
    const bool no_select = (p.flags & kDbgNoSelect) != 0;

    // ---- queries -> TMEM (A operand, K-major: column c of a chunk holds elements 2c, 2c+1) ----
    ptx::griddep_wait();   // prepared queries, seed thresholds and the candidate workspace belong to earlier kernels
    {
      const uint32_t* qrow = reinterpret_cast<const uint32_t*>(
          static_cast<const uint16_t*>(p.qbuf) + static_cast<size_t>(q_row0 + ql) * p.dim);
      // software-pipelined: the global loads of chunk kc+1 are in flight while chunk kc is stored
      uint4 nxt[8];
      {
        const uint4* src = reinterpret_cast<const uint4*>(qrow);
#pragma unroll
        for (int i = 0; i < 8; ++i) nxt[i] = nk_ts > 0 ? src[i] : make_uint4(0, 0, 0, 0);
      }
      for (int kc = 0; kc < nk_ts; ++kc) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          w[4 * i + 0] = nxt[i].x; w[4 * i + 1] = nxt[i].y; w[4 * i + 2] = nxt[i].z; w[4 * i + 3] = nxt[i].w;
        }
        if (kc + 1 < nk_ts) {
          const uint4* src = reinterpret_cast<const uint4*>(qrow + (kc + 1) * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) nxt[i] = src[i];
        }
        ptx::tmem_st_32x32b_x32(t_lane + kc * 32, w);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      if (kPair) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(bar_qready, 0));
      } else {
        ptx::named_bar_sync(2, 160);
      }
    }
    // where this warp reports a drained accumulator buffer: the MMA issuer's CTA
    const uint32_t tempty_at = kPair ? ptx::mapa(bar_tempty, 0) : bar_tempty;

    // ---- per-query state lives in registers ----
    // initial threshold: the k-th best score of the sampled pre-pass when there is one (a valid
    // lower bound of the final k-th score), else -inf; dead (padding) queries never pass
    float seed = (p.seed && live) ? p.seed[static_cast<size_t>(qblk * kNQ + ql) * p.k + (p.k - 1)] : -INFINITY;
    int it = 0;

    if (p.sample_tiles > 0) {
      // ---- phase A: the kJ best scores of my query over this CTA's sample tiles (registers only) ----
      // kJ = 4 when 148 single CTAs sample for a query block, 16 in the pair kernel, where only 37 (or 74) CTAs
      // do: the k-th best of the union of per-CTA top-kJ lists is only a tight bound while a CTA rarely holds
      // more than kJ of the sample's top k.
      float tj[kJ];
#pragma unroll
      for (int j = 0; j < kJ; ++j) tj[j] = -INFINITY;
      for (; it < n_samp; ++it) {
        const int buf = it & 1;
        ptx::mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
        ptx::tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t acc = t_lane + kAccCol0 + buf * kTileN;
        ptx::tmem_ld_32x32b_x32(acc, r0);
        ptx::tmem_ld_32x32b_x32(acc + 32, r1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (kPair) ptx::mbar_arrive_cluster(tempty_at + 8 * buf); else ptx::mbar_arrive(bar_tempty + 8 * buf); }
        uint32_t pm0 = 0, pm1 = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (__uint_as_float(r0[c]) > tj[kJ - 1]) pm0 |= 1u << c;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (__uint_as_float(r1[c]) > tj[kJ - 1]) pm1 |= 1u << c;
        const int64_t nvalid = p.n_local - static_cast<int64_t>(first_tile + it * tile_step) * kTileN;
        if (nvalid < kTileN) {
          const uint64_t vm = (1ull << nvalid) - 1ull;
          pm0 &= static_cast<uint32_t>(vm);
          pm1 &= static_cast<uint32_t>(vm >> 32);
        }
        while ((pm0 | pm1) != 0u) {
          int c;
          if (pm0 != 0u) { c = __ffs(pm0) - 1; pm0 &= pm0 - 1u; }
          else { c = 32 + __ffs(pm1) - 1; pm1 &= pm1 - 1u; }
          const float s = __uint_as_float(pick64(r0, r1, c));
          if (s > tj[kJ - 1]) {       // insert into the sorted (descending) list: one bubble pass from the tail
            tj[kJ - 1] = s;
#pragma unroll
            for (int j = kJ - 1; j > 0; --j)
              if (tj[j] > tj[j - 1]) { const float x = tj[j - 1]; tj[j - 1] = tj[j]; tj[j] = x; }
          }
        }
        __syncwarp();
      }
      const uint32_t token = *reinterpret_cast<const volatile uint32_t*>(p.token) * 64u + static_cast<uint32_t>(p.launch_idx);
      if (lane_ok) {
        uint4* dst = reinterpret_cast<uint4*>(p.top) + (static_cast<size_t>(qblk * kNQ + ql) * tile_step + first_tile) * (kJ / 4);
#pragma unroll
        for (int j = 0; j < kJ / 4; ++j)
          dst[j] = make_uint4(f32_to_ord(tj[4 * j]), f32_to_ord(tj[4 * j + 1]), f32_to_ord(tj[4 * j + 2]), f32_to_ord(tj[4 * j + 3]));
      }
      __threadfence();
      ptx::named_bar_sync(3, 128);
      if (threadIdx.x == 64) st_release_u32(p.top_flag + blockIdx.x, token);   // this CTA's samples are published
      // ---- phase B: epilogue warp e of CTA c owns launch query c + e * grid: k-th best of the CTAs' samples ----
      {
        const int qq = static_cast<int>(blockIdx.x) + (warp - 2) * static_cast<int>(gridDim.x);
        if (qq < p.batch) {
          const int qb = qq / kNQ;                           // lists of query block qb come from CTAs s * nblk + qb
          bool ok[kSlots];
#pragma unroll
          for (int j = 0; j < kSlots; ++j) ok[j] = lane + 32 * j >= tile_step;   // slots that do not exist count as done
          for (int spins = 0; spins < kSeedOwnerSpins; ++spins) {
            bool all = true;
#pragma unroll
            for (int j = 0; j < kSlots; ++j) {
              if (!ok[j]) ok[j] = ld_acquire_u32(p.top_flag + (lane + 32 * j) * p.nblk + qb) == token;
              all = all && ok[j];
            }
            if (__all_sync(0xffffffffu, all)) break;
            __nanosleep(100);
          }
          const uint4* src = reinterpret_cast<const uint4*>(p.top) + static_cast<size_t>(qq) * tile_step * (kJ / 4);
          const uint32_t absent = f32_to_ord(-INFINITY);    // a CTA that did not make it contributes -inf
          uint32_t v[kSlots * kJ];
#pragma unroll
          for (int j = 0; j < kSlots; ++j) {
            const int slot = lane + 32 * j;
#pragma unroll
            for (int u = 0; u < kJ / 4; ++u) {
              uint4 x = make_uint4(0u, 0u, 0u, 0u);               // 0 = not a candidate
              if (slot < tile_step) x = ok[j] ? __ldcg(src + static_cast<size_t>(slot) * (kJ / 4) + u) : make_uint4(absent, absent, absent, absent);
              v[kJ * j + 4 * u + 0] = x.x; v[kJ * j + 4 * u + 1] = x.y; v[kJ * j + 4 * u + 2] = x.z; v[kJ * j + 4 * u + 3] = x.w;
            }
          }
          const uint32_t kth = warp_kth_largest<kSlots * kJ>(v, p.k);
          if (lane == 0) st_release_u64(p.seed_tag + qq, (static_cast<uint64_t>(token) << 32) | kth);
        }
      }
      // ---- every thread picks up the seed of its own query (or goes on unseeded if it never arrives) ----
      if (live) {
        const uint64_t* mine = p.seed_tag + qblk * kNQ + ql;
        for (int spins = 0; spins < kSeedReaderSpins; ++spins) {
          const uint64_t tag = ld_acquire_u64(mine);
          if (static_cast<uint32_t>(tag >> 32) == token) { seed = ord_to_f32(static_cast<uint32_t>(tag)); break; }
          __nanosleep(100);
        }
      }
      __syncwarp();
    }

    float thr = live ? seed : INFINITY;
    uint64_t thrkey = live ? (static_cast<uint64_t>(f32_to_ord(seed)) << 32) : ~0ull;
    int cnt = 0;

    for (; it < n_iter; ++it) {
      const int t = first_tile + (it - n_samp) * tile_step;
      const int buf = it & 1;
      long long w0 = want_stats ? clock64() : 0;
      ptx::mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      if (want_stats) { const long long w1 = clock64(); st_a += w1 - w0; w0 = w1; }
      ptx::tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t acc = t_lane + kAccCol0 + buf * kTileN;
      ptx::tmem_ld_32x32b_x32(acc, r0);
      ptx::tmem_ld_32x32b_x32(acc + 32, r1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // the 64 scores are in registers now
        if (kPair) ptx::mbar_arrive_cluster(tempty_at + 8 * buf); else ptx::mbar_arrive(bar_tempty + 8 * buf);
      }
      if (want_stats) { const long long w1 = clock64(); st_d += w1 - w0; w0 = w1; }
      if (no_select) continue;

      // ---- filter: bitmask of the passages of this tile that reach my query's threshold ----
      uint32_t pm0 = 0, pm1 = 0;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (__uint_as_float(r0[c]) >= thr) pm0 |= 1u << c;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (__uint_as_float(r1[c]) >= thr) pm1 |= 1u << c;
      const int64_t row0 = static_cast<int64_t>(t) * kTileN;
      const int64_t nvalid = p.n_local - row0;  // rows past the end are zero-filled by TMA: mask them
      if (nvalid < kTileN) {
        const uint64_t vm = (1ull << nvalid) - 1ull;
        pm0 &= static_cast<uint32_t>(vm);
        pm1 &= static_cast<uint32_t>(vm >> 32);
      }
      // ---- append the survivors (exact key test decides score ties by row) ----
      if ((pm0 & pm1) == 0xFFFFFFFFu) {
        // every passage of the tile passes (unseeded first tiles): statically indexed, no select tree
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const uint64_t kk = make_key(__uint_as_float(r0[c]), static_cast<uint32_t>(row0) + c);
          if (kk > thrkey) my_list[cnt++] = kk;
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const uint64_t kk = make_key(__uint_as_float(r1[c]), static_cast<uint32_t>(row0) + 32 + c);
          if (kk > thrkey) my_list[cnt++] = kk;
        }
        st_n += 64;
        pm0 = pm1 = 0u;
      }
      // rare path: a few survivors per tile
      while ((pm0 | pm1) != 0u) {
        int c;
        if (pm0 != 0u) { c = __ffs(pm0) - 1; pm0 &= pm0 - 1u; }
        else { c = 32 + __ffs(pm1) - 1; pm1 &= pm1 - 1u; }
        const float s = __uint_as_float(pick64(r0, r1, c));
        const uint64_t kk = make_key(s, static_cast<uint32_t>(row0) + c);
        if (kk > thrkey) {
          my_list[cnt++] = kk;
          ++st_n;
        }
      }
      __syncwarp();
      if (want_stats) { const long long w1 = clock64(); st_b += w1 - w0; w0 = w1; }
      // ---- lists that could overflow during the next tile are cut back to their best k ----
      uint32_t need = __ballot_sync(0xffffffffu, cnt > cap - kTileN);
      while (need != 0u) {  // warp-uniform
        const int l = __ffs(need) - 1;
        need &= need - 1u;
        const int c = __shfl_sync(0xffffffffu, cnt, l);
        uint64_t* lst = warp_lists + static_cast<size_t>(l) * cap;
        const uint64_t nk_key = cap == kCap ? compact_list(lst, c, p.k, lane) : compact_list_stream(lst, c, p.k, lane);
        if (lane == l) {
          thrkey = nk_key;
          thr = ord_to_f32(static_cast<uint32_t>(nk_key >> 32));
          cnt = p.k;
        }
        ++st_m;
      }
      if (want_stats) st_c += clock64() - w0;
    }

    // ---------------- final: publish this CTA's candidate counts ----------------
    // The candidate lists stay where they are (L2-resident workspace); the select kernel reads
    // them directly.  Only lists longer than p.emit are first cut down to their best k.
    if (!no_select) {
      uint32_t need = __ballot_sync(0xffffffffu, cnt > p.emit);
      while (need != 0u) {
        const int l = __ffs(need) - 1;
        need &= need - 1u;
        const int c = __shfl_sync(0xffffffffu, cnt, l);
        uint64_t* lst = warp_lists + static_cast<size_t>(l) * cap;
        if (cap == kCap) compact_list(lst, c, p.k, lane); else compact_list_stream(lst, c, p.k, lane);
        if (lane == l) cnt = p.k;
        ++st_m;
      }
      if (lane_ok) p.part_cnt[static_cast<size_t>(blockIdx.x) * kNQ + ql] = cnt;
    }
    if (want_stats && lane == 0) {
      atomicAdd(&my_stats[kStEpiWaitTmem], static_cast<unsigned long long>(st_a));
      atomicAdd(&my_stats[kStEpiSelect], static_cast<unsigned long long>(st_b));
      atomicAdd(&my_stats[kStEpiCompact], static_cast<unsigned long long>(st_c));
      atomicAdd(&my_stats[kStNumCompact], static_cast<unsigned long long>(st_m));
      atomicAdd(&my_stats[kStEpiLd], static_cast<unsigned long long>(st_d));
    }
    if (want_stats) atomicAdd(&my_stats[kStNumAppend], static_cast<unsigned long long>(st_n));
  }

  // ---------------- teardown ----------------
  if (want_stats && threadIdx.x == 0) my_stats[kStTotal] = clock64() - t_start;
  ptx::tc_fence_before();
  // pair: neither CTA may exit (or free tensor memory) while the other can still signal its barriers
  if (kPair) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (kPair) ptx::tmem_dealloc_pair(tmem_base, kTmemCols); else ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

__global__ void __launch_bounds__(kScanThreads, 1)
mips_scan_kernel(const __grid_constant__ CUtensorMap tmap_e, const __grid_constant__ CUtensorMap tmap_q,
                 const ScanParams p) {
  scan_body<false, false>(tmap_e, tmap_q, p);
}

// Index stored [dim, n_local] (the reference's own layout): the B operand is MN-major.
__global__ void __launch_bounds__(kScanThreads, 1)
mips_scan_dn_kernel(const __grid_constant__ CUtensorMap tmap_e, const __grid_constant__ CUtensorMap tmap_q,
                    const ScanParams p) {
  scan_body<false, true>(tmap_e, tmap_q, p);
}

// CTA-pair variant: launched with cluster dimension 2 (see launch_scan_pair).
__global__ void __launch_bounds__(kScanThreads, 1)
mips_scan_pair_kernel(const __grid_constant__ CUtensorMap tmap_e, const __grid_constant__ CUtensorMap tmap_q,
                      const ScanParams p) {
  scan_body<true, false>(tmap_e, tmap_q, p);
}

// The attribute is per function and device, not per handle: always the maximum, so that handles of different
// dims on one GPU cannot lower each other's limit.
cudaError_t configure_scan() {
  cudaError_t e = cudaFuncSetAttribute(mips_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(mips_scan_dn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(mips_scan_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
}

cudaError_t launch_scan(const CUtensorMap& tmap_e, const CUtensorMap& tmap_q, const ScanParams& p, int grid,
                        size_t smem_bytes, cudaStream_t st) {
  if (p.b_mn)
    return launch_pdl(mips_scan_dn_kernel, dim3(grid), dim3(kScanThreads), smem_bytes, st, g_use_pdl, tmap_e, tmap_q, p);
  return launch_pdl(mips_scan_kernel, dim3(grid), dim3(kScanThreads), smem_bytes, st, g_use_pdl, tmap_e, tmap_q, p);
}

static void pair_launch_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, size_t smem_bytes,
                               cudaStream_t st) {
  cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kScanThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
}

// grid must be even: CTAs 2c and 2c+1 form pair c.
cudaError_t launch_scan_pair(const CUtensorMap& tmap_e, const CUtensorMap& tmap_q, const ScanParams& p, int grid,
                             size_t smem_bytes, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  pair_launch_config(cfg, attr, grid, smem_bytes, st);
  return cudaLaunchKernelEx(&cfg, mips_scan_pair_kernel, tmap_e, tmap_q, p);
}

// How many CTA pairs the device keeps resident at once with this much shared memory (74 on a full B200).
cudaError_t max_resident_pairs(size_t smem_bytes, int* out) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  pair_launch_config(cfg, attr, 2, smem_bytes, nullptr);
  cfg.numAttrs = 1;
  return cudaOccupancyMaxActiveClusters(out, mips_scan_pair_kernel, &cfg);
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
                                    int normalize, void* __restrict__ out, uint32_t* token) {
  ptx::griddep_wait();               // the query buffer may still be read by the previous search
  ptx::griddep_launch_dependents();
  const int row = blockIdx.x;
  // one new token per search: tags the flags of the scan's in-kernel seeding (also on every graph replay)
  if (token != nullptr && row == 0 && threadIdx.x == 0) *token = *token + 1u;
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
                                int out_dtype, int normalize, void* out, uint32_t* token, cudaStream_t st) {
  const int threads = 256;
  switch (q_dtype) {
    case 0:
      return launch_pdl(prep_queries_kernel<__half>, dim3(batch_pad), dim3(threads), 0, st, g_use_pdl,
                        static_cast<const __half*>(q), q_ld, batch, dim, out_dtype, normalize, out, token);
    case 1:
      return launch_pdl(prep_queries_kernel<__nv_bfloat16>, dim3(batch_pad), dim3(threads), 0, st, g_use_pdl,
                        static_cast<const __nv_bfloat16*>(q), q_ld, batch, dim, out_dtype, normalize, out, token);
    default:
      return launch_pdl(prep_queries_kernel<float>, dim3(batch_pad), dim3(threads), 0, st, g_use_pdl,
                        static_cast<const float*>(q), q_ld, batch, dim, out_dtype, normalize, out, token);
  }
}

// ------------------------------------------------------------------------------------------------
// Row gather for the 3-tuple search_knn variant (build_server/index.py:228-229).
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const uint16_t* __restrict__ emb, int64_t ld, int dim, int64_t n_local, int layout,
                                   const int64_t* __restrict__ rows, int64_t n, uint16_t* __restrict__ out) {
  const int64_t i = blockIdx.x;
  if (i >= n) return;
  const int64_t r = rows[i];
  const bool ok = r >= 0 && r < n_local;
  if (layout == 1) {
    const int vec = dim / 8;  // dim % 64 == 0 -> 16-byte vectors
    const uint4* src = reinterpret_cast<const uint4*>(emb + r * ld);
    uint4* dst = reinterpret_cast<uint4*>(out + i * dim);
    for (int c = threadIdx.x; c < vec; c += blockDim.x) dst[c] = ok ? src[c] : make_uint4(0, 0, 0, 0);
  } else {  // [dim, n]: column gather
    for (int c = threadIdx.x; c < dim; c += blockDim.x) out[i * dim + c] = ok ? emb[static_cast<int64_t>(c) * ld + r] : 0;
  }
}

cudaError_t launch_gather_rows(const void* emb, int64_t ld, int dim, int64_t n_local, int layout, const int64_t* rows,
                               int64_t n, void* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  gather_rows_kernel<<<static_cast<unsigned>(n), 128, 0, st>>>(static_cast<const uint16_t*>(emb), ld, dim, n_local, layout,
                                                               rows, n, static_cast<uint16_t*>(out));
  return cudaGetLastError();
}

}  // namespace mips
