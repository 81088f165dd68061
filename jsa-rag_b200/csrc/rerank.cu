// Re-rank of a short candidate list per query: scores[b, j] = <q_b, cand[b, j, :]>, sorted descending,
// top-k kept, the winners' embeddings gathered.  Replaces the tail of RAG.retrieve_with_rerank
// (reference src/rag.py:228-233: einsum("id,ijd->ij"), torch.sort, slice, torch.gather) and the
// re-selection of the 3-tuple search_knn (build_server/index.py:253-255) with one launch.
//
// One CTA per query.  HBM-bound and tiny (B*L*dim elements read once: 19.7 MB at B=64, L=100,
// dim=768 fp32), so the design rule is only: coalesced 16-byte loads, one pass, no intermediate
// in global memory.  Warps stride over the candidates; a lane strides over the dims of one
// candidate; fp32 accumulation; keys (orderable score << 32 | ~position) are sorted by a bitonic
// network in shared memory, which makes the order total: score descending, position ascending.
#include "internal.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace mips {

namespace {

constexpr int kRerankThreads = 256;
constexpr int kRerankMaxCand = 1024;

template <typename T> struct Vec;   // 16-byte vector view of T
template <> struct Vec<float> {
  static constexpr int kN = 4;
  static __device__ inline void load(const float* p, float (&v)[4]) {
    const float4 x = *reinterpret_cast<const float4*>(p);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  }
};
template <> struct Vec<__half> {
  static constexpr int kN = 8;
  static __device__ inline void load(const __half* p, float (&v)[8]) {
    const uint4 x = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&x);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int kN = 8;
  static __device__ inline void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 x = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};

__device__ inline float to_f32(float x) { return x; }
__device__ inline float to_f32(__half x) { return __half2float(x); }
__device__ inline float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T>
__global__ void __launch_bounds__(kRerankThreads)
rerank_kernel(const T* __restrict__ q, int64_t q_ld, const T* __restrict__ cand, int num_cand, int dim, int k, int vec_ok,
              float* __restrict__ out_scores, int64_t* __restrict__ out_pos, int64_t* __restrict__ out_rank,
              T* __restrict__ out_emb) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);                     // [pad]
  int pad = 2;   // >= 2 keeps the query copy behind the keys 16-byte aligned
  while (pad < num_cand) pad <<= 1;
  T* qs = reinterpret_cast<T*>(smem_raw + static_cast<size_t>(pad) * sizeof(uint64_t));   // [dim]

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* qrow = q + static_cast<int64_t>(b) * q_ld;
  const T* crow = cand + static_cast<int64_t>(b) * num_cand * dim;

  for (int d = tid; d < dim; d += kRerankThreads) qs[d] = qrow[d];
  for (int i = num_cand + tid; i < pad; i += kRerankThreads) keys[i] = 0;   // padding sorts last
  __syncthreads();

  constexpr int V = Vec<T>::kN;
  for (int j = warp; j < num_cand; j += kRerankThreads / 32) {
    const T* c = crow + static_cast<int64_t>(j) * dim;
    float acc = 0.f;
    if (vec_ok) {
      for (int d = lane * V; d < dim; d += 32 * V) {
        float a[V], x[V];
        Vec<T>::load(c + d, a);
        Vec<T>::load(qs + d, x);
#pragma unroll
        for (int i = 0; i < V; ++i) acc = fmaf(a[i], x[i], acc);
      }
    } else {
      for (int d = lane; d < dim; d += 32) acc = fmaf(to_f32(c[d]), to_f32(qs[d]), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0)
      keys[j] = (static_cast<uint64_t>(f32_to_ord(acc)) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(j));
  }
  __syncthreads();

  // bitonic sort, descending
  for (int size = 2; size <= pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (pad >> 1); i += kRerankThreads) {
        const int lo = ((i / stride) * (stride << 1)) + (i % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = keys[lo], c2 = keys[hi];
        if ((a < c2) == desc) { keys[lo] = c2; keys[hi] = a; }
      }
      __syncthreads();
    }
  }

  for (int i = tid; i < num_cand; i += kRerankThreads) {
    const uint64_t key = keys[i];
    const int pos = static_cast<int>(~static_cast<uint32_t>(key));
    if (i < k) {
      out_scores[static_cast<int64_t>(b) * k + i] = ord_to_f32(static_cast<uint32_t>(key >> 32));
      out_pos[static_cast<int64_t>(b) * k + i] = pos;
    }
    if (out_rank) out_rank[static_cast<int64_t>(b) * num_cand + pos] = i;
  }
  if (out_emb) {
    for (int j = warp; j < k; j += kRerankThreads / 32) {
      const int pos = static_cast<int>(~static_cast<uint32_t>(keys[j]));
      const T* src = crow + static_cast<int64_t>(pos) * dim;
      T* dst = out_emb + (static_cast<int64_t>(b) * k + j) * dim;
      if (vec_ok) {
        for (int d = lane * V; d < dim; d += 32 * V)
          *reinterpret_cast<uint4*>(dst + d) = *reinterpret_cast<const uint4*>(src + d);
      } else {
        for (int d = lane; d < dim; d += 32) dst[d] = src[d];
      }
    }
  }
}

template <typename T>
cudaError_t launch_rerank_t(const void* q, int64_t q_ld, const void* cand, int batch, int num_cand, int dim, int k,
                            float* out_scores, int64_t* out_pos, int64_t* out_rank, void* out_emb, cudaStream_t st) {
  int pad = 2;
  while (pad < num_cand) pad <<= 1;
  const size_t smem = static_cast<size_t>(pad) * sizeof(uint64_t) + static_cast<size_t>(dim) * sizeof(T);
  constexpr int V = Vec<T>::kN;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const int vec_ok = (dim % V == 0) && aligned(cand) && (!out_emb || aligned(out_emb));
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rerank_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  rerank_kernel<T><<<batch, kRerankThreads, smem, st>>>(static_cast<const T*>(q), q_ld, static_cast<const T*>(cand),
                                                        num_cand, dim, k, vec_ok, out_scores, out_pos, out_rank,
                                                        static_cast<T*>(out_emb));
  return cudaGetLastError();
}

}  // namespace

int rerank_max_candidates() { return kRerankMaxCand; }

cudaError_t launch_rerank(const void* q, int64_t q_ld, const void* cand, int dtype, int batch, int num_cand, int dim,
                          int k, float* out_scores, int64_t* out_pos, int64_t* out_rank, void* out_emb, cudaStream_t st) {
  switch (dtype) {
    case 0: return launch_rerank_t<__half>(q, q_ld, cand, batch, num_cand, dim, k, out_scores, out_pos, out_rank, out_emb, st);
    case 1: return launch_rerank_t<__nv_bfloat16>(q, q_ld, cand, batch, num_cand, dim, k, out_scores, out_pos, out_rank, out_emb, st);
    default: return launch_rerank_t<float>(q, q_ld, cand, batch, num_cand, dim, k, out_scores, out_pos, out_rank, out_emb, st);
  }
}

}  // namespace mips
