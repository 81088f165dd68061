// Peer exchange: the one exchange step of the distributed search (every rank's [B, k] candidates to every rank,
// reference src/index.py:135-142 = 2*W gathers of pickled lists) done with NVLink peer stores instead of a
// collective call.  Every rank owns one buffer  [control | slot 0: W blocks | slot 1: W blocks]  that all peers map
// through CUDA IPC.  A step is two launches on each rank:
//   xchg_push_kernel   stores this rank's candidate block into block[rank] of the current slot on EVERY peer (and
//                      itself), fences, and the last CTA to finish raises flag[slot][rank] = epoch on every peer;
//   xchg_merge_kernel  (merge.cu) waits for the W flags of this epoch and merges straight out of the slot.
// The wait is a wall-clock one (%globaltimer) with a bound the host sets (default 30 minutes, a collective
// watchdog's order of magnitude: a peer that is saving a checkpoint or paused in the allocator is merely late).
// On expiry the kernel writes a host-visible error word, fills its outputs with padding and returns — the context
// stays alive and the next call on the exchange fails with MIPS_ETIMEOUT.
// Two slots alternate by epoch parity.  That is enough: a rank can only start push t+2 after its merge t+1 returned,
// which needed every peer's push t+1, which each peer issued (stream order) after its own merge t had finished
// reading slot t&1.  The epoch lives in device memory so that a captured CUDA graph advances it on every replay.
#include "internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/jsa_mips.h"

namespace mips {

namespace {
constexpr int kPushThreads = 256;

__global__ void __launch_bounds__(kPushThreads)
xchg_push_kernel(XchgPeers peers, int rank, int world, const uint2* __restrict__ local_block, size_t n8, size_t cap) {
  XchgCtrl* ctrl = reinterpret_cast<XchgCtrl*>(peers.base[rank]);
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch);
  const int slot = static_cast<int>(epoch & 1);
  const size_t block_off = kXchgCtrlBytes + (static_cast<size_t>(slot) * world + rank) * cap;
  const size_t stride = static_cast<size_t>(gridDim.x) * kPushThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kPushThreads + threadIdx.x; i < n8; i += stride) {
    const uint2 v = local_block[i];
    for (int p = 0; p < world; ++p) reinterpret_cast<uint2*>(peers.base[p] + block_off)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(&ctrl->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x < world) {
    unsigned long long* flag = &reinterpret_cast<XchgCtrl*>(peers.base[threadIdx.x])->flags[slot][rank];
    const unsigned long long v = epoch + 1;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(v) : "memory");
  }
  if (threadIdx.x == 0) {
    ctrl->ticket = 0;
    *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch) = epoch + 1;
  }
}

// Receiving half of a plain all-gather (used for the queries): waits for the W arrival flags of this epoch, then
// copies the W blocks out of the slot into one contiguous [W, block] array, so the slot can be reused two steps later.
__global__ void __launch_bounds__(kPushThreads)
xchg_gather_kernel(uint8_t* local_base, int world, size_t cap, size_t n8, uint2* __restrict__ out,
                   unsigned long long timeout_ns, int* err_word) {
  XchgCtrl* ctrl = reinterpret_cast<XchgCtrl*>(local_base);
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch);
  const int slot = static_cast<int>((epoch - 1) & 1);
  if (!xchg_wait_flags(ctrl, slot, epoch, world, timeout_ns, err_word)) return;   // late peer: `out` is left untouched
  const uint8_t* slot_base = local_base + kXchgCtrlBytes + static_cast<size_t>(slot) * world * cap;
  const size_t total = n8 * world;
  const size_t stride = static_cast<size_t>(gridDim.x) * kPushThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kPushThreads + threadIdx.x; i < total; i += stride) {
    const size_t src = i / n8, off = i - src * n8;
    out[i] = reinterpret_cast<const uint2*>(slot_base + src * cap)[off];
  }
}
}  // namespace

cudaError_t launch_xchg_gather(uint8_t* local_base, int world, size_t cap, size_t block_bytes, void* out,
                               unsigned long long timeout_ns, int* err_word, cudaStream_t st) {
  const size_t n8 = block_bytes / 8;
  size_t grid = (n8 * world + kPushThreads * 4 - 1) / (kPushThreads * 4);
  if (grid < 1) grid = 1;
  if (grid > 128) grid = 128;
  xchg_gather_kernel<<<static_cast<unsigned>(grid), kPushThreads, 0, st>>>(local_base, world, cap, n8, static_cast<uint2*>(out),
                                                                          timeout_ns, err_word);
  return cudaGetLastError();
}

cudaError_t launch_xchg_push(const XchgPeers& peers, int rank, int world, const void* local_block, size_t block_bytes,
                             size_t cap, cudaStream_t st) {
  const size_t n8 = block_bytes / 8;
  size_t grid = (n8 + kPushThreads * 4 - 1) / (kPushThreads * 4);
  if (grid < 1) grid = 1;
  if (grid > 64) grid = 64;
  xchg_push_kernel<<<static_cast<unsigned>(grid), kPushThreads, 0, st>>>(peers, rank, world,
                                                                        static_cast<const uint2*>(local_block), n8, cap);
  return cudaGetLastError();
}

}  // namespace mips

using namespace mips;

struct mips_xchg {
  int device = 0, rank = 0, world = 1;
  size_t cap = 0, bytes = 0;
  uint8_t* local = nullptr;
  XchgPeers peers = {};
  bool connected = false;
  bool ipc = false;                        // peers[] were opened through CUDA IPC (and must be closed)
  unsigned long long timeout_ns = 1800ull * 1000000000ull;
  int* err_host = nullptr;                 // page-locked, mapped: written by a wait kernel whose peers never arrived
  int* err_dev = nullptr;
  char err[256] = {0};
};

namespace {
int xfail(mips_xchg* x, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (x) snprintf(x->err, sizeof(x->err), "%s%s%s", what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
  cudaGetLastError();
  return code;
}
struct XDeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit XDeviceGuard(int d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != d) ok = cudaSetDevice(d) == cudaSuccess;
  }
  ~XDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" {

int mips_xchg_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

int mips_xchg_create(mips_xchg** out, int device, int rank, int world, size_t block_capacity_bytes) {
  if (!out || world < 1 || world > kXchgMaxWorld || rank < 0 || rank >= world || block_capacity_bytes == 0) return MIPS_EINVAL;
  *out = nullptr;
  mips_xchg* x = new mips_xchg();
  x->device = device; x->rank = rank; x->world = world;
  x->cap = (block_capacity_bytes + 255) / 256 * 256;
  x->bytes = kXchgCtrlBytes + 2 * static_cast<size_t>(world) * x->cap;
  if (const char* t = getenv("JSA_MIPS_XCHG_TIMEOUT_S")) {
    const double v = atof(t);
    if (v > 0) x->timeout_ns = static_cast<unsigned long long>(v * 1e9);
  }
  XDeviceGuard g(device);
  cudaError_t e = g.ok ? cudaMalloc(&x->local, x->bytes) : cudaErrorInvalidDevice;
  if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->bytes);
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&x->err_host), sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) { *x->err_host = 0; e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&x->err_dev), x->err_host, 0); }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (x->local) cudaFree(x->local);
    if (x->err_host) cudaFreeHost(x->err_host);
    delete x;
    cudaGetLastError();
    return MIPS_ECUDA;
  }
  x->peers.base[rank] = x->local;
  *out = x;
  return MIPS_OK;
}

int mips_xchg_set_timeout_ms(mips_xchg* x, int64_t ms) {
  if (!x || ms <= 0) return MIPS_EINVAL;
  x->timeout_ns = static_cast<unsigned long long>(ms) * 1000000ull;
  return MIPS_OK;
}

int mips_xchg_status(mips_xchg* x) {
  if (!x) return MIPS_EINVAL;
  if (x->err_host && *reinterpret_cast<volatile int*>(x->err_host) != 0)
    return xfail(x, MIPS_ETIMEOUT, "a peer's block did not arrive within the exchange timeout; results of that step are padding");
  return MIPS_OK;
}

int mips_xchg_connect_local(mips_xchg* x, mips_xchg* const* all, int n) {
  if (!x || !all || n != x->world) return MIPS_EINVAL;
  if (x->connected) return xfail(x, MIPS_EINVAL, "exchange already connected");
  for (int p = 0; p < n; ++p) {
    if (!all[p] || all[p]->device != x->device || all[p]->cap != x->cap || all[p]->world != x->world || all[p]->rank != p)
      return xfail(x, MIPS_EINVAL, "local wiring needs W exchanges of one device, equal capacity, ranks 0..W-1 in order");
    x->peers.base[p] = all[p]->local;
  }
  x->connected = true;
  x->ipc = false;
  return MIPS_OK;
}

const char* mips_xchg_last_error(mips_xchg* x) { return x ? x->err : "null exchange"; }

int mips_xchg_export(mips_xchg* x, void* out_handle) {
  if (!x || !out_handle) return MIPS_EINVAL;
  XDeviceGuard g(x->device);
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, x->local);
  if (e != cudaSuccess) return xfail(x, MIPS_ECUDA, "cudaIpcGetMemHandle", e);
  memcpy(out_handle, &h, sizeof(h));
  return MIPS_OK;
}

int mips_xchg_connect(mips_xchg* x, const void* all_handles) {
  if (!x || !all_handles) return MIPS_EINVAL;
  if (x->connected) return xfail(x, MIPS_EINVAL, "exchange already connected");
  XDeviceGuard g(x->device);
  const uint8_t* hs = static_cast<const uint8_t*>(all_handles);
  for (int p = 0; p < x->world; ++p) {
    if (p == x->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + static_cast<size_t>(p) * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int c = 0; c < p; ++c)
        if (c != x->rank && x->peers.base[c]) { cudaIpcCloseMemHandle(x->peers.base[c]); x->peers.base[c] = nullptr; }
      return xfail(x, MIPS_EUNSUPPORTED, "cudaIpcOpenMemHandle (no peer mapping between these GPUs/processes)", e);
    }
    x->peers.base[p] = static_cast<uint8_t*>(ptr);
  }
  x->connected = true;
  x->ipc = true;
  return MIPS_OK;
}

size_t mips_xchg_capacity(mips_xchg* x) { return x ? x->cap : 0; }

namespace {
int check_live(mips_xchg* x) {
  if (!x) return MIPS_EINVAL;
  if (!x->connected) return xfail(x, MIPS_ENOTBOUND, "mips_xchg_connect has not been called");
  return mips_xchg_status(x);
}
}  // namespace

int mips_xchg_push(mips_xchg* x, const void* local_block, size_t block_bytes, void* stream) {
  int rc = check_live(x);
  if (rc != MIPS_OK) return rc;
  if (block_bytes % 8 || block_bytes > x->cap) return xfail(x, MIPS_EINVAL, "block must be a multiple of 8 bytes and <= capacity");
  if (block_bytes == 0) return MIPS_OK;
  if (!local_block) return xfail(x, MIPS_EINVAL, "NULL pointer");
  XDeviceGuard g(x->device);
  cudaError_t e = launch_xchg_push(x->peers, x->rank, x->world, local_block, block_bytes, x->cap, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return xfail(x, MIPS_ECUDA, "push launch failed", e);
  return MIPS_OK;
}

int mips_xchg_merge_wait(mips_xchg* x, size_t block_bytes, size_t score_bytes, int batch, int k_in, int k_out,
                         float* out_scores, int64_t* out_ids, void* stream) {
  int rc = check_live(x);
  if (rc != MIPS_OK) return rc;
  if (batch < 0 || k_in <= 0 || k_out <= 0 || k_in > kMaxK || k_out > kMaxK) return xfail(x, MIPS_EINVAL, "bad sizes");
  const size_t need = score_bytes + static_cast<size_t>(batch) * k_in * sizeof(int64_t);
  if (score_bytes % 8 || score_bytes < static_cast<size_t>(batch) * k_in * sizeof(float) || block_bytes != need ||
      block_bytes > x->cap)
    return xfail(x, MIPS_EINVAL, "block layout does not fit the exchange ([scores | ids], 8-byte aligned, <= capacity)");
  if (batch == 0) return MIPS_OK;   // the global batch is the same on every rank, so all of them skip together
  if (!out_scores || !out_ids) return xfail(x, MIPS_EINVAL, "NULL pointer");
  XDeviceGuard g(x->device);
  cudaError_t e = launch_xchg_merge(x->local, x->world, x->cap, score_bytes, batch, k_in, k_out, out_scores, out_ids,
                                    x->timeout_ns, x->err_dev, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return xfail(x, MIPS_ECUDA, "merge launch failed", e);
  return MIPS_OK;
}

int mips_xchg_gather_wait(mips_xchg* x, size_t block_bytes, void* out, void* stream) {
  int rc = check_live(x);
  if (rc != MIPS_OK) return rc;
  if (block_bytes % 8 || block_bytes > x->cap) return xfail(x, MIPS_EINVAL, "block must be a multiple of 8 bytes and <= capacity");
  if (block_bytes == 0) return MIPS_OK;
  if (!out) return xfail(x, MIPS_EINVAL, "NULL pointer");
  XDeviceGuard g(x->device);
  cudaError_t e = launch_xchg_gather(x->local, x->world, x->cap, block_bytes, out, x->timeout_ns, x->err_dev,
                                     static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return xfail(x, MIPS_ECUDA, "gather launch failed", e);
  return MIPS_OK;
}

int mips_xchg_merge(mips_xchg* x, const void* local_block, size_t block_bytes, size_t score_bytes, int batch, int k_in,
                    int k_out, float* out_scores, int64_t* out_ids, void* stream) {
  if (batch == 0) return check_live(x);
  int rc = mips_xchg_push(x, local_block, block_bytes, stream);
  if (rc != MIPS_OK) return rc;
  return mips_xchg_merge_wait(x, block_bytes, score_bytes, batch, k_in, k_out, out_scores, out_ids, stream);
}

int mips_xchg_gather(mips_xchg* x, const void* local_block, size_t block_bytes, void* out, void* stream) {
  int rc = mips_xchg_push(x, local_block, block_bytes, stream);
  if (rc != MIPS_OK) return rc;
  return mips_xchg_gather_wait(x, block_bytes, out, stream);
}

int mips_xchg_destroy(mips_xchg* x) {
  if (!x) return MIPS_OK;
  XDeviceGuard g(x->device);
  cudaDeviceSynchronize();
  if (x->ipc)
    for (int p = 0; p < x->world; ++p)
      if (p != x->rank && x->peers.base[p]) cudaIpcCloseMemHandle(x->peers.base[p]);
  if (x->local) cudaFree(x->local);
  if (x->err_host) cudaFreeHost(x->err_host);
  cudaGetLastError();
  delete x;
  return MIPS_OK;
}

}  // extern "C"
