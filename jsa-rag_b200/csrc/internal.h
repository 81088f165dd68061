// Internal declarations shared by the translation units of libjsa_mips.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mips {

// ---- compile-time geometry of the fused scan kernel ----
// D[query, passage] = Q[query, :] . P[passage, :]   with tcgen05.mma M=128 (queries), N=64 (passages)
constexpr int kNQ = 128;         // queries per pass (UMMA M): one TMEM lane / one epilogue thread per query
constexpr int kTileN = 64;       // passages per tile (UMMA N): one fp32 TMEM column per passage
constexpr int kKChunk = 64;      // elements per K chunk: 64 x 2 B = one 128-byte swizzle row
constexpr int kUmmaK = 16;       // K per tcgen05.mma for 16-bit operands
constexpr int kChunkBytes = kTileN * kKChunk * 2;                 // 8 KiB: [64 passages x 64 el], SWIZZLE_128B
constexpr int kQChunkBytes = kNQ * kKChunk * 2;                   // 16 KiB: [128 queries x 64 el] (smem-resident K tail)
constexpr int kMaxQBlocks = 8;   // query blocks (of kNQ) one launch can carry: 4 single CTAs or 4 CTA pairs
constexpr int kMaxStages = 12;
constexpr int kTmemCols = 512;   // whole TMEM: queries (A operand) + two accumulator buffers
constexpr int kAccCol0 = kTmemCols - 2 * kTileN;                  // accumulators live in the last 128 columns
constexpr int kMaxTsChunks = kAccCol0 / (kKChunk / 2);            // 12 K chunks (768 dims) of the queries fit in TMEM
// Two list geometries: k <= kSmallK keeps 512-slot lists (register-resident compaction / select),
// k <= kMaxK uses 2048-slot lists and the streaming variants.
constexpr int kCap = 512;        // candidate slots per (CTA, query), small-k mode
constexpr int kSortE = kCap / 32;  // keys per lane when a warp compacts one small list
constexpr int kEmit = 256;       // small-k mode: a CTA hands at most this many candidates per query to the select kernel
constexpr int kSmallK = 128;     // largest k of the small (fast) mode
constexpr int kCapBig = 2048;    // candidate slots per (CTA, query), big-k mode
constexpr int kMaxK = 1024;      // largest supported top-k
constexpr int kMaxDim = 1024;
constexpr int kScanThreads = 192;  // warp0 TMA, warp1 MMA + TMEM alloc, warps2-5 epilogue/select
constexpr int kCtrlBytes = 1024;   // barriers + TMEM base pointer
constexpr int kMaxSmem = 232448;   // 227 KiB opt-in dynamic shared memory per CTA on sm_100

struct ScanParams {
  int64_t n_local;       // rows in this shard
  int num_tiles;         // tiles (of kTileN passages) this launch scans: tiles [0, num_tiles)
  int num_kchunks;       // dim / 64
  int num_stages;        // pipeline depth that fits next to the resident queries
  int chunks_per_stage;  // K chunks (8 KiB each) one pipeline stage carries
  int k;                 // top-k (<= kMaxK)
  int batch;             // valid queries in this launch (<= kNQ * nblk)
  int nblk;              // query blocks (of kNQ) scanned concurrently by this launch (1, 2 or 4; grid % nblk == 0);
                         // CTA c serves block c % nblk — in the pair kernel CTAs 2j, 2j+1 are one tcgen05 CTA pair
  int m64;               // 1: UMMA M=64 (batch <= 64), 0: UMMA M=128
  int b_mn;              // 1: index stored [dim, n_local] (MN-major B operand), 0: [n_local, dim] (K-major)
  int q_row0;            // first row of this pass in the prepared query buffer
  int dim;               // embedding dimension
  const void* qbuf;      // prepared queries [batch_pad, dim] in the index dtype (row-major, zero padded)
  uint32_t idesc;        // tcgen05 instruction descriptor (dtype dependent)
  int cap;               // slots per candidate list (kCap or kCapBig)
  int emit;              // lists longer than this are cut back to their best k before they are handed over
  uint64_t* cand;        // [grid][kNQ][cap] packed (orderable score << 32 | ~row) keys
  int* part_cnt;         // [grid][kNQ] number of candidates each CTA leaves at the head of its lists
  int64_t id_base, id_stride;
  const float* seed;     // optional [kNQ][k] sorted scores of the sampled pre-pass (NULL = none)
  int flags;             // diagnostics: kDbgNoSelect / kDbgNoMma (results are then meaningless)
  unsigned long long* stats;  // optional [grid][kNumStats] per-CTA cycle counters (NULL = off)
  // In-kernel sampled seeding (k <= kSmallK): every CTA first scans its own first `sample_tiles` tiles keeping only
  // the kTopJ best scores per query, publishes them with a token-tagged flag, one warp per query takes the k-th best
  // of all CTAs' values (a valid lower bound of the final k-th score) and publishes it with the same tag, and the
  // full scan starts with that threshold.  0 = off (the thresholds come from `seed` or start at -inf).
  int sample_tiles;
  int launch_idx;           // scan launch number within the search (token = search token * 64 + launch_idx)
  uint32_t* top;            // [nblk * kNQ][grid / nblk][kTopJ or kTopJPair] orderable score images (nblk <= kMaxQBlocks)
  uint32_t* top_flag;       // [grid] token of the launch whose samples CTA c has published
  uint64_t* seed_tag;       // [nblk * kNQ] token << 32 | seed image
  const uint32_t* token;    // search token, bumped by prep_queries_kernel
  // Pair kernel with several pair blocks per launch: the pairs that walk the same tile sequence publish how far
  // they are and none runs more than a few tiles ahead of the slowest, so a tile is still in the L2 when its other
  // readers ask for it (without this the siblings drift apart and DRAM delivers every tile 1.5 - 2.5 times).
  unsigned long long* progress;   // [grid / 2] (tag << 32 | tiles issued) per CTA pair; NULL = no lock-step
  int scan_seq;                   // number of this scan launch within the search (tag = token * 4096 + scan_seq)
  int lock_window;                // tiles a pair may be ahead of its slowest sibling
};
constexpr int kLockWindow = 8;    // default window (JSA_MIPS_LOCK_WINDOW overrides)
constexpr int kLockEvery = 4;     // the siblings' progress is looked at every this many tiles
constexpr int kTopJ = 4;
constexpr int kSeedSlots = 5;   // a lane of the selecting warp looks after CTAs lane, lane + 32, ... (5 * 32 = 160 >= 148)
constexpr int kTopJPair = 16;   // pair kernel: 37 or 74 lists per query block, hence deeper per-CTA samples
constexpr int kSeedSlotsPair = 3;   // 3 * 32 = 96 >= 74 pairs

constexpr int kDbgNoSelect = 1;  // epilogue only drains TMEM (isolates GEMM + streaming)
constexpr int kDbgNoMma = 2;     // no tcgen05.mma, stages are released immediately (isolates TMA streaming)
constexpr int kDbgNoSeed = 4;    // disable the sampled pre-passes (thresholds start at -inf)
constexpr int kDbgForceM128 = 16; // always use UMMA M=128 (A/B test of the M=64 small-batch mode)
constexpr int kDbgOneBlock = 32;  // one query block per launch even for large batches (A/B test of the L2-shared multi-block scan)
constexpr int kDbgHostPrepass = 64; // seed thresholds with separate sampled scan + select launches (the pre-fusion path; always used for k > 128)
constexpr int kDbgTimeScan = 8;  // record CUDA events around every full-shard scan launch (mips_scan_times_ms)
constexpr int kDbgNoTma = 256;   // the producer hands over stages without loading them (power/latency split; results meaningless)
constexpr int kDbgNoLockstep = 512;  // pairs sharing a tile sequence run free (A/B test of the L2 lock-step)
constexpr int kDbgFourPairBlocks = 2048;  // 4 pair blocks per launch whenever > 512 queries are left (normally only for long tile sequences)
constexpr int kDbgNoPair = 128;  // batches > 128 without tcgen05 CTA pairs (the round-1 multi-block path; A/B test)
enum Stat { kStProdWait = 0, kStMmaWaitFull, kStMmaWaitTmem, kStEpiWaitTmem, kStEpiSelect, kStEpiCompact,
            kStNumCompact, kStNumAppend, kStTotal, kStEpiLd, kNumStats };

// Launch helper: programmatic dependent launch lets the prologue of kernel N+1 (barrier init, TMEM
// allocation, the first TMA loads of the static index) overlap the tail of kernel N.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
extern bool g_use_pdl;   // JSA_MIPS_PDL=0 disables it

// launchers (defined in scan.cu / merge.cu); return cudaError_t of the launch
cudaError_t launch_prep_queries(const void* q, int q_dtype, int64_t q_ld, int batch, int batch_pad, int dim,
                                int out_dtype, int normalize, void* out, uint32_t* token, cudaStream_t st);
cudaError_t launch_scan(const CUtensorMap& tmap_e, const CUtensorMap& tmap_q, const ScanParams& p, int grid,
                        size_t smem_bytes, cudaStream_t st);
cudaError_t launch_scan_pair(const CUtensorMap& tmap_e, const CUtensorMap& tmap_q, const ScanParams& p, int grid,
                             size_t smem_bytes, cudaStream_t st);
cudaError_t max_resident_pairs(size_t smem_bytes, int* out);
cudaError_t configure_scan();
cudaError_t launch_merge(const float* scores, const int64_t* ids, int num_lists, int64_t list_stride,
                         int64_t id_list_stride, int batch, int k_in, int k_out, float* out_scores, int64_t* out_ids,
                         cudaStream_t st);
cudaError_t launch_select(const uint64_t* cand, const int* part_cnt, int num_lists, int nblk, int cap, int batch, int k,
                          int64_t id_base, int64_t id_stride, float* out_scores, int64_t* out_ids, cudaStream_t st);
cudaError_t configure_merge();
cudaError_t launch_gather_rows(const void* emb, int64_t ld, int dim, int64_t n_local, int layout, const int64_t* rows,
                               int64_t n, void* out, cudaStream_t st);

// ---- peer exchange (exchange.cu / merge.cu): candidate blocks pushed into every peer's slot over NVLink ----
constexpr int kXchgMaxWorld = 16;
constexpr size_t kXchgCtrlBytes = 512;
struct XchgCtrl {                       // at the start of every rank's exchange buffer
  unsigned long long epoch;             // number of pushes this rank has completed
  unsigned int ticket;                  // CTA arrival counter of the running push
  unsigned int pad;
  unsigned long long flags[2][kXchgMaxWorld];   // flags[slot][src] = epoch of the last block src stored into `slot`
};
static_assert(sizeof(XchgCtrl) <= kXchgCtrlBytes, "control block too large");
struct XchgPeers { uint8_t* base[kXchgMaxWorld]; };
cudaError_t launch_xchg_push(const XchgPeers& peers, int rank, int world, const void* local_block, size_t block_bytes,
                             size_t cap, cudaStream_t st);
cudaError_t launch_xchg_gather(uint8_t* local_base, int world, size_t cap, size_t block_bytes, void* out,
                               unsigned long long timeout_ns, int* err_word, cudaStream_t st);
cudaError_t launch_xchg_merge(uint8_t* local_base, int world, size_t cap, size_t s_bytes, int batch, int k_in, int k_out,
                              float* out_scores, int64_t* out_ids, unsigned long long timeout_ns, int* err_word,
                              cudaStream_t st);

#ifdef __CUDACC__
// Block-wide wait of the receiving kernels: thread p < world polls flag[slot][p] until it carries `epoch`.  The
// bound is wall-clock time (%globaltimer); on expiry the host-mapped error word is set and false is returned to the
// whole block — the caller writes padding / nothing and returns, the CUDA context stays usable.
__device__ __forceinline__ bool xchg_wait_flags(XchgCtrl* ctrl, int slot, unsigned long long epoch, int world,
                                                unsigned long long timeout_ns, int* err_word) {
  int late = 0;
  if (threadIdx.x < world) {
    const unsigned long long* flag = &ctrl->flags[slot][threadIdx.x];
    unsigned long long seen = 0, t0 = 0;
    for (unsigned spins = 0;; ++spins) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
      if (seen >= epoch) break;
      if ((spins & 1023u) == 1023u) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) { late = 1; break; }
      }
      __nanosleep(spins < 64 ? 20 : (spins < 4096 ? 200 : 2000));
    }
  }
  late = __syncthreads_or(late);
  if (late && threadIdx.x == 0 && err_word != nullptr) {
    *reinterpret_cast<volatile int*>(err_word) = 1;
    __threadfence_system();
  }
  return late == 0;
}
#endif

cudaError_t launch_rerank(const void* q, int64_t q_ld, const void* cand, int dtype, int batch, int num_cand, int dim,
                          int k, float* out_scores, int64_t* out_pos, int64_t* out_rank, void* out_emb, cudaStream_t st);
int rerank_max_candidates();

// ---- order-preserving float <-> uint32 (larger float -> larger uint) ----
__host__ __device__ inline uint32_t f32_to_ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f + 0.0f);  // +0.0f canonicalises -0 to +0
#else
  union { float f; uint32_t u; } c; c.f = f + 0.0f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

}  // namespace mips
