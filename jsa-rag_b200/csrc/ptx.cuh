// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Only what the MIPS scan kernel needs; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe of a barrier phase (one poll, never suspends).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug traps (-> cudaErrorLaunchFailure) instead of hanging the GPU.  Every wait of the scan
// is on work of the same CTA (pair), so the bound is a wall-clock one far above any legitimate stall (SM preemption,
// a debugger): 20 s measured with %globaltimer, looked at every 64k polls.
constexpr uint64_t kMbarTimeoutNs = 20ull * 1000000000ull;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFFu) == 0u) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kMbarTimeoutNs) __trap();
    }
  }
}
// Barriers that also collect arrivals of the other CTA of a pair are waited on the same way: what crosses the
// pair is tensor-memory / async-proxy state ordered by the tcgen05 fences around the wait, no generic-proxy data,
// so the default (CTA-scope) acquire is sufficient — a cluster-scope one would cost an L1 invalidation per poll.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }

// ------------------------------------------------------------------ CTA pairs (clusters of 2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ named barriers
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One lane of the (converged) warp is elected; returns true on that lane only.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ programmatic dependent launch
// A kernel launched with programmatic stream serialization may start before its predecessor in the
// stream has finished: everything before griddep_wait() overlaps the predecessor's tail, everything
// after it sees the predecessor's memory.  griddep_launch_dependents() lets the successor start early.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ TMA
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void prefetch_tensormap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled load global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint32_t bar, uint32_t dst_smem, int32_t c0,
                                            int32_t c1, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "l"(cache_hint)
      : "memory");
}

// Same for a CTA pair (cta_group::2): the data lands in the executing CTA's shared memory, the completion
// bytes are counted on an mbarrier of the pair's leader CTA (`bar_cluster` is a shared::cluster address).
__device__ __forceinline__ void tma_load_2d_pair(const void* tmap, uint32_t bar_cluster, uint32_t dst_smem, int32_t c0,
                                                 int32_t c1, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(cache_hint)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// Pair forms: the same warp of BOTH CTAs of the pair executes them.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A * B, kind::f16 (fp16 and bf16 operands, fp32 accumulation); A from tensor memory (TS: a_tmem
// addresses 128 lanes x 8 columns of packed 16-bit values per K = 16 step) or shared memory (SS), B from shared memory.
// CTA-pair MMAs (cta_group::2, issued by the leader CTA only): M = 256 — rows [0,128) are the leader's A operand and
// accumulator lanes, rows [128,256) the peer's; the B operand's N rows are split, the first N/2 in the leader's
// shared memory and the second N/2 in the peer's, at the same offsets (one descriptor serves both).
// Arrive on the mbarrier at the same shared-memory offset in BOTH CTAs of the pair once all previously issued
// tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i reads TMEM lane base+i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// registers -> TMEM: thread i writes 32 consecutive 32-bit columns of lane base+i
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (tile rows are 128 B = 64
// 16-bit elements wide, 8-row groups are 1024 B apart).  Bit layout: start>>4 [0,14),
// LBO>>4 [16,30) (ignored for swizzled K-major), SBO>>4 [32,46), version=1 [46,48),
// layout_type=2 (SWIZZLE_128B) [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// The same descriptor as two 32-bit words: only the low word (start address) changes from MMA to MMA, so the issue
// loop does 32-bit adds on it (the address field cannot carry out: shared memory is < 256 KiB).
constexpr uint32_t kSw128DescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t sw128_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }

// One K chunk (64 elements = four K=16 MMAs) in ONE asm block: the operand addresses of the four instructions are
// derived inside the block (tensor-memory A: +8 columns per step; descriptor low words: +kBStep / +2), so the compiler
// keeps them next to their MMA instead of hoisting dozens of them into (spilling) uniform registers.
template <bool kPair, uint32_t kBStep>
__device__ __forceinline__ void umma_ts_x4(uint32_t tmem_d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc_first) {
  if (kPair) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 a1, a2, a3, b1, b2, b3;\n\t.reg .b64 d0, d1, d2, d3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 d0, {%2, %5};\n\t"
        "add.u32 b1, %2, %6;\n\tadd.u32 b2, b1, %6;\n\tadd.u32 b3, b2, %6;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "mov.b64 d1, {b1, %5};\n\tmov.b64 d2, {b2, %5};\n\tmov.b64 d3, {b3, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], d0, %3, {%7, %7, %7, %7, %7, %7, %7, %7}, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a1], d1, %3, {%7, %7, %7, %7, %7, %7, %7, %7}, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a2], d2, %3, {%7, %7, %7, %7, %7, %7, %7, %7}, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [a3], d3, %3, {%7, %7, %7, %7, %7, %7, %7, %7}, 1;\n\t}"
        :
        : "r"(tmem_d), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc_first), "r"(kSw128DescHi), "n"(kBStep), "r"(0u)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 a1, a2, a3, b1, b2, b3;\n\t.reg .b64 d0, d1, d2, d3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 d0, {%2, %5};\n\t"
        "add.u32 b1, %2, %6;\n\tadd.u32 b2, b1, %6;\n\tadd.u32 b3, b2, %6;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "mov.b64 d1, {b1, %5};\n\tmov.b64 d2, {b2, %5};\n\tmov.b64 d3, {b3, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d0, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], d1, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], d2, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], d3, %3, 1;\n\t}"
        :
        : "r"(tmem_d), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc_first), "r"(kSw128DescHi), "n"(kBStep)
        : "memory");
  }
}
template <bool kPair, uint32_t kBStep>
__device__ __forceinline__ void umma_ss_x4(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc_first) {
  if (kPair) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 a1, a2, a3, b1, b2, b3;\n\t.reg .b64 d0, d1, d2, d3, e0, e1, e2, e3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 d0, {%2, %5};\n\tmov.b64 e0, {%1, %5};\n\t"
        "add.u32 b1, %2, %6;\n\tadd.u32 b2, b1, %6;\n\tadd.u32 b3, b2, %6;\n\t"
        "add.u32 a1, %1, 2;\n\tadd.u32 a2, %1, 4;\n\tadd.u32 a3, %1, 6;\n\t"
        "mov.b64 d1, {b1, %5};\n\tmov.b64 d2, {b2, %5};\n\tmov.b64 d3, {b3, %5};\n\t"
        "mov.b64 e1, {a1, %5};\n\tmov.b64 e2, {a2, %5};\n\tmov.b64 e3, {a3, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], e0, d0, %3, {%7, %7, %7, %7, %7, %7, %7, %7}, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], e1, d1, %3, {%7, %7, %7, %7, %7, %7, %7, %7}, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], e2, d2, %3, {%7, %7, %7, %7, %7, %7, %7, %7}, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], e3, d3, %3, {%7, %7, %7, %7, %7, %7, %7, %7}, 1;\n\t}"
        :
        : "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc_first), "r"(kSw128DescHi), "n"(kBStep), "r"(0u)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 a1, a2, a3, b1, b2, b3;\n\t.reg .b64 d0, d1, d2, d3, e0, e1, e2, e3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 d0, {%2, %5};\n\tmov.b64 e0, {%1, %5};\n\t"
        "add.u32 b1, %2, %6;\n\tadd.u32 b2, b1, %6;\n\tadd.u32 b3, b2, %6;\n\t"
        "add.u32 a1, %1, 2;\n\tadd.u32 a2, %1, 4;\n\tadd.u32 a3, %1, 6;\n\t"
        "mov.b64 d1, {b1, %5};\n\tmov.b64 d2, {b2, %5};\n\tmov.b64 d3, {b3, %5};\n\t"
        "mov.b64 e1, {a1, %5};\n\tmov.b64 e2, {a2, %5};\n\tmov.b64 e3, {a3, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], e0, d0, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], e1, d1, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], e2, d2, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], e3, d3, %3, 1;\n\t}"
        :
        : "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc_first), "r"(kSw128DescHi), "n"(kBStep)
        : "memory");
  }
}

// Instruction descriptor for kind::f16: fp32 accumulate, A/B both K-major.
// c_format [4,6)=1 (F32); a_format [7,10), b_format [10,13): 0 = F16, 1 = BF16;
// n_dim [17,23) = N>>3; m_dim [24,29) = M>>4.
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n, int ab_format) {
  return (1u << 4) | (static_cast<uint32_t>(ab_format) << 7) | (static_cast<uint32_t>(ab_format) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace ptx
