// C ABI of libjsa_mips.so (declared in include/jsa_mips.h).  Host-side orchestration only:
// descriptor encoding, workspace carving, kernel launches.  No allocation on the hot path once the
// workspace exists; no CPU fallback — on a non-sm_100 device every call fails with MIPS_EUNSUPPORTED.
#include "../../include/jsa_mips.h"
#include "internal.h"
#include "ptx.cuh"

#include <nvtx3/nvToolsExt.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace mips;

struct mips_handle {
  int device = 0;
  int dim = 0;
  int dtype = MIPS_DTYPE_F16;
  int num_sms = 0;
  // bound index
  const void* emb = nullptr;
  int64_t n_local = 0, ld = 0, id_base = 0, id_stride = 1;
  int layout = 1;   // 1: [n_local, dim] K-major rows; 0: [dim, n_local] (the reference's layout, MN-major operand)
  CUtensorMap tmap_e;
  CUtensorMap tmap_e_half;   // boxes of kTileN / 2 passages: what each CTA of a tcgen05 pair loads per tile
  bool bound = false;
  // kernel geometry
  int num_kchunks = 0, num_stages = 0, chunks_per_stage = 2;
  size_t smem_bytes = 0;
  // CTA-pair kernel (batches > 128): stages of half tiles, pairs the device keeps resident
  int pair_stages = 0, max_pairs = 0;
  size_t pair_smem_bytes = 0;
  // internal buffers
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int ws_pins = 0;                 // captured graphs holding pointers into ws (mips_workspace_pin)
  std::vector<void*> ws_retired;   // outgrown while pinned: freed when the last pin goes
  void* sync = nullptr;  // zero-initialised: search token | per-CTA sample flags | tagged seeds (in-kernel sampled seeding)
  void* io = nullptr;  // device staging for mips_search_host: queries | scores | ids
  size_t io_bytes = 0;
  int last_launches = 0;
  int dbg_flags = 0;
  unsigned long long* dbg_stats = nullptr;
  // optional CUDA-event timing of the dominant kernel (the full-shard scan), for roofline reporting
  static constexpr int kMaxTimed = 256;
  cudaEvent_t ev0[kMaxTimed], ev1[kMaxTimed];
  int n_timed = 0;
  bool timing_ready = false;
  std::string err;
};

namespace mips {
bool g_use_pdl = []() { const char* e = getenv("JSA_MIPS_PDL"); return !(e && e[0] == '0'); }();
}

namespace {
// NVTX ranges around the launches of a search (JSA_MIPS_NVTX=1; off by default: a range is two extra driver calls)
const bool g_nvtx = []() { const char* e = getenv("JSA_MIPS_NVTX"); return e && e[0] == '1'; }();
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) : on(g_nvtx) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};
}  // namespace

namespace {

// layout of mips_handle::sync: [0] search token | kSyncFlagOff: uint32 flag per CTA | + kSyncSeedOff: uint64 seed tags
constexpr size_t kSyncFlagOff = 256, kSyncSeedOff = 4096, kSyncProgBytes = 1024;   // progress words: one per CTA pair

std::string g_create_err;
std::mutex g_mu;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int fail(mips_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else { std::lock_guard<std::mutex> lk(g_mu); g_create_err = buf; }
  return code;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

#define CUDA_TRY(h, expr)                                                                          \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) return fail((h), MIPS_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Generic row-major [rows, cols] 16-bit matrix -> 2-D tensor map with box {64 cols (128 B), box_rows}.
int encode_2d_map(mips_handle* h, CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(h, MIPS_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kKChunk), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = h->dtype == MIPS_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, MIPS_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MIPS_OK;
}

struct WsLayout {
  size_t q_off, cand_off, pk_off, seed_s_off, seed_i_off, top_off, total;
};

WsLayout ws_layout(const mips_handle* h, int max_batch, int max_k) {
  WsLayout w;
  const size_t bpad = align_up(static_cast<size_t>(max_batch > 0 ? max_batch : 1), 2 * kNQ);   // whole CTA pairs
  const size_t grid = static_cast<size_t>(h->num_sms);
  const size_t cap = max_k <= kSmallK ? kCap : kCapBig;
  size_t off = 0;
  w.q_off = off;    off += align_up(bpad * h->dim * 2, 1024);
  w.cand_off = off; off += align_up(grid * kNQ * cap * sizeof(uint64_t), 1024);
  w.pk_off = off;   off += align_up(grid * kNQ * sizeof(int), 1024);
  w.seed_s_off = off; off += align_up(static_cast<size_t>(kMaxQBlocks * kNQ) * max_k * sizeof(float), 1024);
  w.seed_i_off = off; off += align_up(static_cast<size_t>(kMaxQBlocks * kNQ) * max_k * sizeof(int64_t), 1024);
  w.top_off = off; off += align_up(static_cast<size_t>(kMaxQBlocks * kNQ) * grid * kTopJPair * sizeof(uint32_t), 1024);
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

int mips_abi_version(void) { return JSA_MIPS_ABI_VERSION; }
int mips_max_k(void) { return kMaxK; }
int mips_max_dim(void) { return kMaxDim; }

const char* mips_last_error(const mips_handle* h) {
  if (h) return h->err.c_str();
  return g_create_err.c_str();
}

int mips_create(mips_handle** out, int device, int dim, int index_dtype) {
  if (!out) return fail(nullptr, MIPS_EINVAL, "out is NULL");
  *out = nullptr;
  if (dim <= 0 || dim % kKChunk != 0 || dim > kMaxDim)
    return fail(nullptr, MIPS_EINVAL, "dim=%d unsupported: must be a multiple of %d and <= %d", dim, kKChunk, kMaxDim);
  if (index_dtype != MIPS_DTYPE_F16 && index_dtype != MIPS_DTYPE_BF16)
    return fail(nullptr, MIPS_EINVAL, "index dtype %d unsupported (fp16=0, bf16=1)", index_dtype);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, MIPS_EUNSUPPORTED, "no CUDA device: this engine has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, MIPS_EINVAL, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, MIPS_EUNSUPPORTED, "device %d is sm_%d%d; this library contains sm_100a code only", device,
                prop.major, prop.minor);
  DeviceGuard g(device);
  if (!g.ok) return fail(nullptr, MIPS_ECUDA, "cudaSetDevice(%d) failed", device);
  mips_handle* h = new mips_handle();
  h->device = device;
  h->dim = dim;
  h->dtype = index_dtype;
  h->num_sms = prop.multiProcessorCount;
  h->num_kchunks = dim / kKChunk;
  // queries live in tensor memory (up to kMaxTsChunks K chunks); only a longer K tail needs shared memory
  const int q_bytes = (h->num_kchunks > kMaxTsChunks ? h->num_kchunks - kMaxTsChunks : 0) * kQChunkBytes;
  // Two pipeline stages, each as deep in K as fits: every stage hand-off (mbarrier round trip +
  // tcgen05.commit) costs the MMA issuer ~200 cycles during which the tensor pipe drains, so the
  // fewer, larger stages win (measured: 2 x 96 KiB stages beat 12 x 16 KiB by 20 % at dim 768).
  int cps = ((kMaxSmem - 1024 - kCtrlBytes - q_bytes) / 2) / kChunkBytes;
  if (const char* e = getenv("JSA_MIPS_CPS")) cps = atoi(e);   // tuning knob: K chunks per pipeline stage
  if (cps < 1) cps = 1;
  if (cps > h->num_kchunks) cps = h->num_kchunks;
  const int stage_bytes = cps * kChunkBytes;
  int stages = (kMaxSmem - 1024 - kCtrlBytes - q_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) { delete h; return fail(nullptr, MIPS_EINVAL, "dim=%d leaves no room for the TMA pipeline", dim); }
  h->num_stages = stages;
  h->chunks_per_stage = cps;
  h->smem_bytes = 1024 + q_bytes + static_cast<size_t>(stages) * stage_bytes + kCtrlBytes;
  cudaError_t e = configure_scan();
  if (e != cudaSuccess) {
    delete h;
    return fail(nullptr, MIPS_ECUDA, "cudaFuncSetAttribute(smem=%d) failed: %s", kMaxSmem, cudaGetErrorString(e));
  }
  // CTA-pair kernel: every CTA loads half tiles (32 passages x the whole K extent per stage), so twice as many
  // stages fit; the pair count comes from the occupancy query (74 on a full B200: every TPC holds one pair)
  {
    const int half_stage = h->num_kchunks * (kChunkBytes / 2);
    int ps = (kMaxSmem - 1024 - kCtrlBytes - q_bytes) / half_stage;
    if (ps > kMaxStages) ps = kMaxStages;
    if (const char* e2 = getenv("JSA_MIPS_PAIR_STAGES")) { const int v = atoi(e2); if (v >= 2 && v < ps) ps = v; }
    h->pair_stages = ps;
    h->pair_smem_bytes = 1024 + q_bytes + static_cast<size_t>(ps) * half_stage + kCtrlBytes;
    int np = 0;
    if (ps >= 2 && max_resident_pairs(h->pair_smem_bytes, &np) == cudaSuccess) h->max_pairs = np;
    cudaGetLastError();
    if (const char* e2 = getenv("JSA_MIPS_PAIRS")) { if (e2[0] == '0') h->max_pairs = 0; }
  }
  constexpr size_t kSyncBytes = kSyncFlagOff + kSyncSeedOff + kMaxQBlocks * kNQ * sizeof(uint64_t) + kSyncProgBytes;
  e = cudaMalloc(&h->sync, kSyncBytes);
  if (e == cudaSuccess) e = cudaMemset(h->sync, 0, kSyncBytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (h->sync) cudaFree(h->sync);
    delete h;
    return fail(nullptr, MIPS_ECUDA, "allocating the seeding flag words failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return MIPS_OK;
}

void mips_destroy(mips_handle* h) {
  if (!h) return;
  DeviceGuard g(h->device);
  if (h->sync) cudaFree(h->sync);
  if (h->ws) cudaFree(h->ws);
  for (void* p : h->ws_retired) cudaFree(p);
  if (h->io) cudaFree(h->io);
  if (h->timing_ready)
    for (int i = 0; i < mips_handle::kMaxTimed; ++i) { cudaEventDestroy(h->ev0[i]); cudaEventDestroy(h->ev1[i]); }
  delete h;
}

int mips_bind_index_layout(mips_handle* h, const void* emb, int64_t n_local, int64_t ld, int layout, int64_t id_base,
                           int64_t id_stride) {
  if (!h) return MIPS_EINVAL;
  if (layout != 0 && layout != 1) return fail(h, MIPS_EINVAL, "layout=%d invalid (0 = [dim, n], 1 = [n, dim])", layout);
  if (n_local < 0 || n_local > 0x7FFFFF00ll) return fail(h, MIPS_EINVAL, "n_local=%lld out of range", (long long)n_local);
  if (n_local > 0 && !emb) return fail(h, MIPS_EINVAL, "emb is NULL");
  const int64_t min_ld = layout == 1 ? h->dim : n_local;
  if (ld < min_ld || (ld * 2) % 16 != 0)
    return fail(h, MIPS_EINVAL, "row stride ld=%lld must be >= %lld and 16-byte aligned", (long long)ld, (long long)min_ld);
  if (reinterpret_cast<uintptr_t>(emb) % 16 != 0) return fail(h, MIPS_EINVAL, "emb must be 16-byte aligned");
  h->emb = emb;
  h->n_local = n_local;
  h->ld = ld;
  h->layout = layout;
  h->id_base = id_base;
  h->id_stride = id_stride;
  h->bound = false;
  if (n_local > 0) {
    // [n, dim]: box = 64 passages x 64 dims, rows are passages (K-major B operand)
    // [dim, n]: box = 64 dims x 64 passages, rows are dims   (MN-major B operand, no transpose needed)
    int rc = layout == 1 ? encode_2d_map(h, &h->tmap_e, emb, n_local, h->dim, ld, kTileN)
                         : encode_2d_map(h, &h->tmap_e, emb, h->dim, n_local, ld, kKChunk);
    if (rc != MIPS_OK) return rc;
    if (layout == 1) {
      rc = encode_2d_map(h, &h->tmap_e_half, emb, n_local, h->dim, ld, kTileN / 2);
      if (rc != MIPS_OK) return rc;
    }
  }
  h->bound = true;
  return MIPS_OK;
}

int mips_bind_index(mips_handle* h, const void* emb, int64_t n_local, int64_t ld, int64_t id_base, int64_t id_stride) {
  return mips_bind_index_layout(h, emb, n_local, ld, 1, id_base, id_stride);
}

int mips_workspace_bytes(const mips_handle* h, int max_batch, int max_k, size_t* out) {
  if (!h || !out || max_batch < 0 || max_k <= 0 || max_k > kMaxK) return MIPS_EINVAL;
  *out = ws_layout(h, max_batch, max_k).total;
  return MIPS_OK;
}

int mips_workspace_pin(mips_handle* h, int delta) {
  if (!h) return MIPS_EINVAL;
  if (h->ws_pins + delta < 0) return fail(h, MIPS_EINVAL, "workspace pin count would become negative");
  h->ws_pins += delta;
  if (h->ws_pins == 0 && !h->ws_retired.empty()) {
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    for (void* p : h->ws_retired) cudaFree(p);
    h->ws_retired.clear();
  }
  return MIPS_OK;
}

int mips_search_local(mips_handle* h, const void* queries, int q_dtype, int64_t q_ld, int batch, int k,
                      int normalize, float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (!h) return MIPS_EINVAL;
  h->last_launches = 0;
  if (!h->bound) return fail(h, MIPS_ENOTBOUND, "mips_bind_index has not been called");
  if (batch < 0 || k <= 0) return fail(h, MIPS_EINVAL, "batch=%d k=%d invalid", batch, k);
  if (k > h->n_local) return fail(h, MIPS_EKRANGE, "selected index k out of range (k=%d > n_local=%lld)", k, (long long)h->n_local);
  if (k > kMaxK) return fail(h, MIPS_EINVAL, "k=%d exceeds the fused top-k limit %d", k, kMaxK);
  if (batch == 0) return MIPS_OK;
  if (!queries || !out_scores || !out_ids) return fail(h, MIPS_EINVAL, "NULL queries/outputs");
  if (q_dtype < 0 || q_dtype > 2) return fail(h, MIPS_EINVAL, "q_dtype=%d invalid", q_dtype);
  if (q_ld < h->dim) return fail(h, MIPS_EINVAL, "q_ld=%lld < dim", (long long)q_ld);
  DeviceGuard g(h->device);
  if (!g.ok) return fail(h, MIPS_ECUDA, "cudaSetDevice(%d) failed", h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const WsLayout w = ws_layout(h, batch, k);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  if (ws) {
    if (workspace_bytes < w.total) return fail(h, MIPS_EWORKSPACE, "workspace %zu < required %zu bytes", workspace_bytes, w.total);
    if (reinterpret_cast<uintptr_t>(ws) % 1024 != 0) return fail(h, MIPS_EINVAL, "workspace must be 1024-byte aligned");
  } else {
    if (h->ws_bytes < w.total) {
      // grow-only internal workspace (sized for the largest request seen); not on the steady-state path
      if (h->ws && h->ws_pins > 0) {
        h->ws_retired.push_back(h->ws);   // a captured graph still points into it
      } else {
        CUDA_TRY(h, cudaStreamSynchronize(st));
        if (h->ws) cudaFree(h->ws);
      }
      h->ws = nullptr; h->ws_bytes = 0;
      CUDA_TRY(h, cudaMalloc(&h->ws, w.total));
      h->ws_bytes = w.total;
    }
    ws = static_cast<uint8_t*>(h->ws);
  }
  void* qbuf = ws + w.q_off;
  const int bpad = static_cast<int>(align_up(batch, 2 * kNQ));

  NvtxRange nv_search("mips.search_local");
  {
    NvtxRange nv("mips.prep");
    CUDA_TRY(h, launch_prep_queries(queries, q_dtype, q_ld, batch, bpad, h->dim, h->dtype, normalize, qbuf,
                                    static_cast<uint32_t*>(h->sync), st));
  }
  h->last_launches++;

  CUtensorMap tmap_q;
  int rc = encode_2d_map(h, &tmap_q, qbuf, bpad, h->dim, h->dim, kNQ);
  if (rc != MIPS_OK) return rc;

  const int num_tiles = static_cast<int>((h->n_local + kTileN - 1) / kTileN);
  const int grid = num_tiles < h->num_sms ? num_tiles : h->num_sms;
  ScanParams p = {};
  p.n_local = h->n_local;
  p.num_tiles = num_tiles;
  p.num_kchunks = h->num_kchunks;
  p.num_stages = h->num_stages;
  p.chunks_per_stage = h->chunks_per_stage;
  p.k = k;
  p.b_mn = h->layout == 0 ? 1 : 0;
  p.dim = h->dim;
  p.qbuf = qbuf;
  p.cap = k <= kSmallK ? kCap : kCapBig;
  p.emit = k <= kSmallK ? kEmit : kCapBig;
  p.cand = reinterpret_cast<uint64_t*>(ws + w.cand_off);
  p.part_cnt = reinterpret_cast<int*>(ws + w.pk_off);
  p.id_base = h->id_base;
  p.id_stride = h->id_stride;
  p.flags = h->dbg_flags;
  p.stats = h->dbg_stats;

  float* seed_scores = reinterpret_cast<float*>(ws + w.seed_s_off);
  int64_t* seed_ids = reinterpret_cast<int64_t*>(ws + w.seed_i_off);

  // Queries are processed in launches of up to kMaxBlocks blocks of kNQ (128) queries.  With nblk > 1
  // blocks per launch the CTAs split into nblk groups that scan the same tile sequence side by side
  // (one HBM read feeds nblk blocks through the L2), which moves large batches from HBM-bound passes
  // towards the tensor-core bound.
  //
  // More than 128 queries left: tcgen05 CTA pairs (cta_group::2, UMMA M = 256).  A pair serves 256 queries and each
  // of its CTAs loads only half of every passage tile, which halves the shared-memory fill per unit of tensor work;
  // one launch carries 1 or 2 pair blocks (256 / 512 queries; with 74 pairs, 2 blocks x 37 tile sequences).
  constexpr int kMaxBlocks = 4;
  const bool pairs_ok = h->layout == 1 && h->max_pairs >= 1 && !(h->dbg_flags & (kDbgNoPair | kDbgOneBlock));
  int n_launch = 0, scan_seq = 0;
  for (int q0 = 0; q0 < batch;) {
    const int rem = batch - q0;
    const bool pair = pairs_ok && rem > kNQ;
    int nblk, launch_grid;
    if (pair) {
      // pair blocks of this launch: 1, 2 or 4.  4 blocks (1024 queries per pass over the index) halve the HBM traffic
      // once more but leave 2 of 74 pairs idle and keep four pairs in lock-step: measured +2..3 % on one GPU at 33M
      // rows, -7 % on the shards of 2-8 GPUs.  JSA_MIPS_PAIR_BLOCKS=1|2|4 forces the maximum.
      // 0 (default) = automatic: 4 blocks only when a tile sequence is long (>= 20k tiles, i.e. >= ~23M rows on a
      // full device), where the saved HBM power outweighs the stalls
      static const int env_npb = []() { const char* e = getenv("JSA_MIPS_PAIR_BLOCKS"); const int v = e ? atoi(e) : 0; return (v == 4 || v == 2 || v == 1) ? v : 0; }();
      int pairs = grid / 2 < h->max_pairs ? grid / 2 : h->max_pairs;
      const int max_npb = (h->dbg_flags & kDbgFourPairBlocks) ? 4
                          : (env_npb ? env_npb : ((pairs >= 4 && num_tiles / (pairs / 4) >= 20000) ? 4 : 2));
      int npb = rem > 4 * kNQ && max_npb >= 4 ? 4 : (rem > 2 * kNQ && max_npb >= 2 ? 2 : 1);
      nblk = 2 * npb;
      launch_grid = pairs >= npb ? (pairs / npb) * nblk : nblk;    // tiny indices: surplus pairs just idle
    } else {
      nblk = (rem + kNQ - 1) / kNQ;
      if (nblk > kMaxBlocks) nblk = kMaxBlocks;
      if (nblk == 3) nblk = 2;                       // grid (148) must divide evenly
      if (h->dbg_flags & kDbgOneBlock) nblk = 1;
      launch_grid = grid >= nblk ? (grid / nblk) * nblk : nblk;   // tiny indices: surplus CTAs just idle
    }
    const int nslots = launch_grid / nblk;         // CTAs (= candidate lists) per query block
    p.nblk = nblk;
    p.batch = rem < kNQ * nblk ? rem : kNQ * nblk;
    p.q_row0 = q0;
    p.seed = nullptr;
    p.m64 = (!pair && nblk == 1 && p.batch <= 64 && !(h->dbg_flags & kDbgForceM128)) ? 1 : 0;
    p.idesc = ptx::make_idesc_f16(pair ? 2 * kNQ : (p.m64 ? 64 : kNQ), kTileN, h->dtype == MIPS_DTYPE_BF16 ? 1 : 0) |
              (p.b_mn ? (1u << 16) : 0u);   // bit 16: B operand is MN-major
    p.token = static_cast<const uint32_t*>(h->sync);
    p.progress = nullptr;
    static const int lock_window = []() { const char* e = getenv("JSA_MIPS_LOCK_WINDOW"); const int v = e ? atoi(e) : kLockWindow; return v < 1 ? 1 : v; }();
    p.lock_window = lock_window;
    if (pair && nblk > 2 && !(h->dbg_flags & kDbgNoLockstep) && static_cast<size_t>(launch_grid / 2) * 8 <= kSyncProgBytes)
      p.progress = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(h->sync) + kSyncFlagOff + kSyncSeedOff +
                                                         kMaxQBlocks * kNQ * sizeof(uint64_t));
    p.num_stages = pair ? h->pair_stages : h->num_stages;
    p.chunks_per_stage = pair ? h->num_kchunks : h->chunks_per_stage;
    const CUtensorMap& tmap_e = pair ? h->tmap_e_half : h->tmap_e;
    const size_t smem_bytes = pair ? h->pair_smem_bytes : h->smem_bytes;
    auto scan = pair ? launch_scan_pair : launch_scan;

    // Sampled pre-passes.  The k-th best score of any sample of the shard is a valid lower bound of
    // the final k-th score, so it can seed the thresholds of a larger pass, which then appends only
    // ~k * tiles / (lists * sample_tiles) candidates per (CTA, query): few enough that no list is ever
    // compacted in-stream and the select kernel takes the raw lists.  levels[] holds the tiles per CTA
    // of each pre-pass (the first is unseeded and short enough to leave <= emit candidates per list).
    const int tiles_per_cta = (num_tiles + nslots - 1) / nslots;
    int levels[6];
    int n_levels = 0;
    if (tiles_per_cta > p.emit / kTileN && !(h->dbg_flags & kDbgNoSeed)) {   // longer unseeded scans would overflow `emit`
      // target appended candidates per (CTA, query): ~150 for the 512-slot lists, ~1200 for the 2048-slot ones
      const int64_t denom = static_cast<int64_t>(nslots) * (k <= kSmallK ? 150 : 1200);
      int64_t need = tiles_per_cta;
      int tmp[6];
      int nt = 0;
      while (nt < 6) {
        int64_t nxt = (need * k + denom - 1) / denom;
        if (nxt < 1) nxt = 1;
        if (nxt >= need) break;
        tmp[nt++] = static_cast<int>(nxt);
        need = nxt;
        if (nxt <= p.emit / kTileN) break;   // an unseeded pass this short leaves <= emit candidates per list
      }
      for (int i = nt - 1; i >= 0; --i) levels[n_levels++] = tmp[i];
    }
    // k <= 128: the last (largest) sample is scanned by the full-shard launch itself (token-tagged hand-offs inside
    // the kernel), so a search is prep + scan + select.  Needs one epilogue warp per launch query for the selection
    // and 4x more per-CTA values than k: with fewer (large batches split the CTAs over 2-4 query blocks) the top-kTopJ
    // truncation loosens the seed and the separate sampled launches win (measured at batch 1024: 48.3 vs 49.5 ms).
    p.sample_tiles = 0;
    if (n_levels > 0 && k <= kSmallK && !(h->dbg_flags & kDbgHostPrepass) &&
        nslots * (pair ? kTopJPair : kTopJ) >= 4 * k && nslots <= 32 * (pair ? kSeedSlotsPair : kSeedSlots) && launch_grid * 4 >= p.batch && launch_grid * sizeof(uint32_t) <= kSyncSeedOff &&
        n_launch < 64) {
      p.sample_tiles = levels[n_levels - 1];
      p.launch_idx = n_launch;
      p.top = reinterpret_cast<uint32_t*>(ws + w.top_off);
      p.top_flag = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(h->sync) + kSyncFlagOff);
      p.seed_tag = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(h->sync) + kSyncFlagOff + kSyncSeedOff);
      n_levels = 0;
    }
    ++n_launch;
    for (int lv = 0; lv < n_levels; ++lv) {
      NvtxRange nv("mips.sampled_prepass");
      ScanParams pp = p;
      pp.num_tiles = levels[lv] * nslots < num_tiles ? levels[lv] * nslots : num_tiles;
      pp.stats = nullptr;
      pp.scan_seq = (scan_seq++) & 4095;
      CUDA_TRY(h, scan(tmap_e, tmap_q, pp, launch_grid, smem_bytes, st));
      CUDA_TRY(h, launch_select(p.cand, p.part_cnt, nslots, nblk, p.cap, p.batch, k, 0, 1, seed_scores, seed_ids, st));
      h->last_launches += 2;
      p.seed = seed_scores;
    }
    p.scan_seq = (scan_seq++) & 4095;
    const bool timed = (h->dbg_flags & kDbgTimeScan) && h->timing_ready && h->n_timed < mips_handle::kMaxTimed;
    if (timed) CUDA_TRY(h, cudaEventRecord(h->ev0[h->n_timed], st));
    {
      NvtxRange nv(pair ? "mips.scan_pair" : "mips.scan");
      CUDA_TRY(h, scan(tmap_e, tmap_q, p, launch_grid, smem_bytes, st));
    }
    if (timed) { CUDA_TRY(h, cudaEventRecord(h->ev1[h->n_timed], st)); h->n_timed++; }
    NvtxRange nv_sel("mips.select");
    CUDA_TRY(h, launch_select(p.cand, p.part_cnt, nslots, nblk, p.cap, p.batch, k, h->id_base, h->id_stride,
                              out_scores + static_cast<size_t>(q0) * k, out_ids + static_cast<size_t>(q0) * k, st));
    h->last_launches += 2;
    q0 += p.batch;
  }
  return MIPS_OK;
}

int mips_merge_topk_strided(int device, const float* scores, const int64_t* ids, int num_lists,
                            int64_t score_list_stride, int64_t id_list_stride, int batch, int k_in, int k_out,
                            float* out_scores, int64_t* out_ids, void* stream) {
  if (num_lists <= 0 || batch < 0 || k_in <= 0 || k_out <= 0 || k_in > kMaxK || k_out > kMaxK)
    return fail(nullptr, MIPS_EINVAL, "mips_merge_topk: bad sizes lists=%d batch=%d k_in=%d k_out=%d", num_lists, batch, k_in, k_out);
  if (batch == 0) return MIPS_OK;
  if (!scores || !ids || !out_scores || !out_ids) return fail(nullptr, MIPS_EINVAL, "mips_merge_topk: NULL pointer");
  DeviceGuard g(device);
  if (!g.ok) return fail(nullptr, MIPS_ECUDA, "cudaSetDevice(%d) failed", device);
  cudaError_t e = launch_merge(scores, ids, num_lists, score_list_stride, id_list_stride, batch, k_in, k_out, out_scores,
                               out_ids, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(nullptr, MIPS_ECUDA, "merge launch failed: %s", cudaGetErrorString(e));
  return MIPS_OK;
}

int mips_merge_topk(int device, const float* scores, const int64_t* ids, int num_lists, int batch, int k_in, int k_out,
                    float* out_scores, int64_t* out_ids, void* stream) {
  const int64_t stride = static_cast<int64_t>(batch) * k_in;
  return mips_merge_topk_strided(device, scores, ids, num_lists, stride, stride, batch, k_in, k_out, out_scores, out_ids,
                                 stream);
}

int mips_max_rerank_candidates(void) { return rerank_max_candidates(); }

int mips_rerank(int device, const void* queries, int64_t q_ld, const void* cand, int dtype, int batch, int num_cand,
                int dim, int k, float* out_scores, int64_t* out_pos, int64_t* out_rank, void* out_emb, void* stream) {
  if (batch < 0 || num_cand <= 0 || dim <= 0 || k <= 0 || q_ld < dim || num_cand > rerank_max_candidates() ||
      dim > 16384 || (dtype != MIPS_DTYPE_F16 && dtype != MIPS_DTYPE_BF16 && dtype != MIPS_DTYPE_F32))
    return fail(nullptr, MIPS_EINVAL, "mips_rerank: bad arguments batch=%d num_cand=%d dim=%d k=%d dtype=%d", batch,
                num_cand, dim, k, dtype);
  if (k > num_cand) return fail(nullptr, MIPS_EKRANGE, "selected index k out of range (k=%d > %d candidates)", k, num_cand);
  if (batch == 0) return MIPS_OK;
  if (!queries || !cand || !out_scores || !out_pos) return fail(nullptr, MIPS_EINVAL, "mips_rerank: NULL pointer");
  DeviceGuard g(device);
  if (!g.ok) return fail(nullptr, MIPS_ECUDA, "cudaSetDevice(%d) failed", device);
  cudaError_t e = launch_rerank(queries, q_ld, cand, dtype, batch, num_cand, dim, k, out_scores, out_pos, out_rank, out_emb,
                                static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(nullptr, MIPS_ECUDA, "rerank launch failed: %s", cudaGetErrorString(e));
  return MIPS_OK;
}

int mips_gather_rows(mips_handle* h, const int64_t* local_rows, int64_t n, void* out, void* stream) {
  if (!h) return MIPS_EINVAL;
  if (!h->bound) return fail(h, MIPS_ENOTBOUND, "mips_bind_index has not been called");
  if (n < 0 || (n > 0 && (!local_rows || !out))) return fail(h, MIPS_EINVAL, "bad gather arguments");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(h, MIPS_ECUDA, "cudaSetDevice(%d) failed", h->device);
  CUDA_TRY(h, launch_gather_rows(h->emb, h->ld, h->dim, h->n_local, h->layout, local_rows, n, out, static_cast<cudaStream_t>(stream)));
  return MIPS_OK;
}

static int search_host_impl(mips_handle* h, const float* host_queries, int batch, int k, int normalize, float* host_scores,
                            int64_t* host_ids, void* stream, bool wait) {
  if (!h) return MIPS_EINVAL;
  h->last_launches = 0;
  if (batch < 0 || k <= 0) return fail(h, MIPS_EINVAL, "batch=%d k=%d invalid", batch, k);
  if (batch == 0) return MIPS_OK;
  if (!host_queries || !host_scores || !host_ids) return fail(h, MIPS_EINVAL, "NULL host buffer");
  DeviceGuard g(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t qb = align_up(static_cast<size_t>(batch) * h->dim * sizeof(float), 256);
  const size_t sb = align_up(static_cast<size_t>(batch) * k * sizeof(float), 256);
  const size_t ib = align_up(static_cast<size_t>(batch) * k * sizeof(int64_t), 256);
  if (h->io_bytes < qb + sb + ib) {
    CUDA_TRY(h, cudaStreamSynchronize(st));
    if (h->io) cudaFree(h->io);
    h->io = nullptr; h->io_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->io, qb + sb + ib));
    h->io_bytes = qb + sb + ib;
  }
  uint8_t* io = static_cast<uint8_t*>(h->io);
  float* dq = reinterpret_cast<float*>(io);
  float* ds = reinterpret_cast<float*>(io + qb);
  int64_t* di = reinterpret_cast<int64_t*>(io + qb + sb);
  CUDA_TRY(h, cudaMemcpyAsync(dq, host_queries, static_cast<size_t>(batch) * h->dim * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = mips_search_local(h, dq, MIPS_DTYPE_F32, h->dim, batch, k, normalize, ds, di, nullptr, 0, stream);
  if (rc != MIPS_OK) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(host_scores, ds, static_cast<size_t>(batch) * k * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaMemcpyAsync(host_ids, di, static_cast<size_t>(batch) * k * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (wait) CUDA_TRY(h, cudaStreamSynchronize(st));
  return MIPS_OK;
}

int mips_search_host(mips_handle* h, const float* host_queries, int batch, int k, int normalize, float* host_scores,
                     int64_t* host_ids, void* stream) {
  return search_host_impl(h, host_queries, batch, k, normalize, host_scores, host_ids, stream, true);
}

int mips_search_host_async(mips_handle* h, const float* host_queries, int batch, int k, int normalize,
                           float* host_scores, int64_t* host_ids, void* stream) {
  return search_host_impl(h, host_queries, batch, k, normalize, host_scores, host_ids, stream, false);
}

int mips_last_launch_count(const mips_handle* h) { return h ? h->last_launches : 0; }

// JSON float list -> fp32 (host only).  The reference's wire format sends the query matrix as a flat list of
// decimal floats (build_server/server_start.py:18-21,186): 65 536 numbers of ~18 digits for a 64 x 1024 batch,
// which Python's json needs ~40 ms to parse.  This is a plain decimal scanner: up to 19 significant digits are
// accumulated exactly in 64 bits and scaled by a power of ten in double precision (exact powers up to 1e22); longer
// mantissas or exponents outside that range go through strtod.  The result can differ from a correctly rounded
// strtod by one double ulp before the rounding to fp32 — far below the fp16 rounding the queries undergo next.
int64_t mips_parse_float_list(const char* text, size_t len, float* out, int64_t max_out) {
  static const double kPow10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  if (!text || !out) return -1;
  const char* p = text;
  const char* end = text + len;
  int64_t n = 0;
  while (p < end) {
    while (p < end && (*p == ' ' || *p == ',' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '[')) ++p;
    if (p >= end || *p == ']') break;
    const char* start = p;
    bool neg = false;
    if (*p == '-') { neg = true; ++p; } else if (*p == '+') { ++p; }
    uint64_t mant = 0;
    int digits = 0, dropped = 0, frac = 0;
    bool any = false, seen_dot = false;
    for (; p < end; ++p) {
      const char c = *p;
      if (c >= '0' && c <= '9') {
        any = true;
        if (digits < 19) { mant = mant * 10 + static_cast<uint64_t>(c - '0'); if (mant != 0 || seen_dot) ++digits; else digits = 0; if (seen_dot) ++frac; }
        else { ++dropped; if (seen_dot) { /* beyond 19 digits: ignored fraction digits */ } else { --frac; } }
      } else if (c == '.' && !seen_dot) {
        seen_dot = true;
      } else {
        break;
      }
    }
    int exp10 = 0;
    if (p < end && (*p == 'e' || *p == 'E')) {
      const char* q = p + 1;
      bool eneg = false;
      if (q < end && (*q == '-' || *q == '+')) { eneg = *q == '-'; ++q; }
      int e = 0;
      bool edig = false;
      for (; q < end && *q >= '0' && *q <= '9'; ++q) { edig = true; if (e < 10000) e = e * 10 + (*q - '0'); }
      if (edig) { exp10 = eneg ? -e : e; p = q; }
    }
    if (!any) return -2;                 // not a number (NaN / Infinity / garbage): the caller falls back to json
    if (n >= max_out) return -3;
    const int scale = exp10 - frac;
    double v;
    if (dropped == 0 && scale >= -22 && scale <= 22) {
      v = static_cast<double>(mant);
      v = scale < 0 ? v / kPow10[-scale] : v * kPow10[scale];
    } else {
      char buf[64];
      const size_t l = static_cast<size_t>(p - start);
      if (l >= sizeof(buf)) return -2;
      memcpy(buf, start, l);
      buf[l] = 0;
      v = strtod(buf, nullptr);
      neg = false;                       // strtod already applied the sign
    }
    out[n++] = static_cast<float>(neg ? -v : v);
  }
  return n;
}

int mips_debug_config(mips_handle* h, int flags, void* stats_dev) {
  if (!h) return MIPS_EINVAL;
  h->dbg_flags = flags;
  h->dbg_stats = static_cast<unsigned long long*>(stats_dev);
  h->n_timed = 0;
  if ((flags & kDbgTimeScan) && !h->timing_ready) {
    DeviceGuard g(h->device);
    for (int i = 0; i < mips_handle::kMaxTimed; ++i) {
      CUDA_TRY(h, cudaEventCreate(&h->ev0[i]));
      CUDA_TRY(h, cudaEventCreate(&h->ev1[i]));
    }
    h->timing_ready = true;
  }
  return MIPS_OK;
}

int mips_scan_times_ms(mips_handle* h, float* out, int max_out, int* n_out) {
  if (!h || !out || !n_out) return MIPS_EINVAL;
  DeviceGuard g(h->device);
  int n = h->n_timed < max_out ? h->n_timed : max_out;
  for (int i = 0; i < n; ++i) {
    CUDA_TRY(h, cudaEventSynchronize(h->ev1[i]));
    CUDA_TRY(h, cudaEventElapsedTime(&out[i], h->ev0[i], h->ev1[i]));
  }
  *n_out = n;
  h->n_timed = 0;
  return MIPS_OK;
}
int mips_debug_num_stats(void) { return kNumStats; }
int mips_num_sms(const mips_handle* h) { return h ? h->num_sms : 0; }

}  // extern "C"
