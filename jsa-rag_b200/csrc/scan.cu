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
// Warp-wide bitonic sort, descending, of 32*E 64-bit keys; key i lives in lane i/E, slot i%E.
// ------------------------------------------------------------------------------------------------
template <int E>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&key)[E], int lane) {
  constexpr int N = 32 * E;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j < E) {  // partner in the same lane
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
      } else {  // partner in lane ^ (j / E)
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

__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(f32_to_ord(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - row);
}

// Sort candidate list q (c valid entries) descending, keep the best `k` at its head, update the
// query's threshold.  Returns the sorted keys in `key` (position i = lane*E + e).  Whole warp.
__device__ __forceinline__ void compact_list(uint64_t* __restrict__ list, int c, int k, int lane,
                                             uint64_t (&key)[kSortE], uint64_t* thrkey_s, float* thr_s, int* cnt_s,
                                             int q) {
  const ulonglong2* src = reinterpret_cast<const ulonglong2*>(list + lane * kSortE);
#pragma unroll
  for (int e = 0; e < kSortE; e += 2) {
    ulonglong2 v = src[e >> 1];
    const int i = lane * kSortE + e;
    key[e] = (i < c) ? v.x : 0ull;
    key[e + 1] = (i + 1 < c) ? v.y : 0ull;
  }
  warp_sort_desc<kSortE>(key, lane);
  // write back the head (positions < kMaxK); only the first min(c, k) are meaningful afterwards
  if (lane * kSortE < kMaxK) {
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(list + lane * kSortE);
#pragma unroll
    for (int e = 0; e < kSortE; e += 2) dst[e >> 1] = make_ulonglong2(key[e], key[e + 1]);
  }
  if (c >= k) {
    uint64_t kth = 0;
#pragma unroll
    for (int e = 0; e < kSortE; ++e)
      if (e == ((k - 1) % kSortE)) kth = key[e];
    kth = __shfl_sync(0xffffffffu, kth, (k - 1) / kSortE);
    if (lane == 0) {
      thrkey_s[q] = kth;
      thr_s[q] = ord_to_f32(static_cast<uint32_t>(kth >> 32));
      cnt_s[q] = k;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The scan kernel
// ------------------------------------------------------------------------------------------------
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
    thr_s[q] = live ? -INFINITY : INFINITY;
    thrkey_s[q] = live ? 0ull : ~0ull;
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

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar_qfull, nk * kQChunkBytes);
      for (int kc = 0; kc < nk; ++kc)
        ptx::tma_load_2d(&tmap_q, bar_qfull, q_smem + kc * kQChunkBytes, kc * kKChunk, p.q_row0, ptx::kEvictLast);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = first_tile; t < p.num_tiles; t += tile_step) {
        for (int kc = 0; kc < nk; ++kc) {
          ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, kStageBytes);
          ptx::tma_load_2d(&tmap_e, bar_full + 8 * stage, st_smem + stage * kStageBytes, kc * kKChunk, t * kTileM,
                           ptx::kEvictFirst);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      ptx::mbar_wait(bar_qfull, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = first_tile; t < p.num_tiles; t += tile_step, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kNQ;
        for (int kc = 0; kc < nk; ++kc) {
          ptx::mbar_wait(bar_full + 8 * stage, phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = st_smem + stage * kStageBytes;
          const uint32_t b_addr = q_smem + kc * kQChunkBytes;
#pragma unroll
          for (int k4 = 0; k4 < kKChunk / kUmmaK; ++k4) {
            // advancing K by 16 elements = 32 bytes inside the 128-byte swizzle row
            const uint64_t da = ptx::make_kmajor_sw128_desc(a_addr + k4 * 32);
            const uint64_t db = ptx::make_kmajor_sw128_desc(b_addr + k4 * 32);
            ptx::umma_f16(d_tmem, da, db, p.idesc, (kc | k4) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(bar_empty + 8 * stage);  // stage reusable once these MMAs retire
          if (kc == nk - 1) ptx::umma_commit(bar_tfull + 8 * buf);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue / select ===========================
    const int ew = warp - 2;        // 0..3: which 16 queries this warp compacts
    const int quarter = warp & 3;   // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint64_t* my_cand = p.cand + static_cast<size_t>(blockIdx.x) * kNQ * kCap;
    uint64_t key[kSortE];

    int it = 0;
    for (int t = first_tile; t < p.num_tiles; t += tile_step, ++it) {
      const int buf = it & 1;
      ptx::mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t r0[32], r1[32];
      ptx::tmem_ld_32x32b_x32(t_lane + buf * kNQ, r0);
      ptx::tmem_ld_32x32b_x32(t_lane + buf * kNQ + 32, r1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * buf);  // accumulators are in registers now

      const int64_t row = static_cast<int64_t>(t) * kTileM + quarter * 32 + lane;
      const bool valid = row < p.n_local;

      // fast filter: does any of my 64 scores reach its query's threshold?
      bool any = false;
      const float4* thr4 = reinterpret_cast<const float4*>(thr_s);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 th = thr4[g];
        any |= __uint_as_float(r0[4 * g + 0]) >= th.x;
        any |= __uint_as_float(r0[4 * g + 1]) >= th.y;
        any |= __uint_as_float(r0[4 * g + 2]) >= th.z;
        any |= __uint_as_float(r0[4 * g + 3]) >= th.w;
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 th = thr4[8 + g];
        any |= __uint_as_float(r1[4 * g + 0]) >= th.x;
        any |= __uint_as_float(r1[4 * g + 1]) >= th.y;
        any |= __uint_as_float(r1[4 * g + 2]) >= th.z;
        any |= __uint_as_float(r1[4 * g + 3]) >= th.w;
      }
      if (any && valid) {
        const uint32_t row32 = static_cast<uint32_t>(row);
#pragma unroll
        for (int q = 0; q < kNQ; ++q) {
          const float s = __uint_as_float(q < 32 ? r0[q & 31] : r1[q & 31]);
          if (s >= thr_s[q]) {
            const uint64_t kk = make_key(s, row32);
            if (kk > thrkey_s[q]) {  // exact (score desc, row asc) order against the current k-th best
              const int pos = atomicAdd(&cnt_s[q], 1);
              my_cand[q * kCap + pos] = kk;
            }
          }
        }
      }
      // all appends of this tile are done -> lists that could overflow on the next tile are compacted
      ptx::named_bar_sync(1, 128);
#pragma unroll 1
      for (int i = 0; i < kNQ / 4; ++i) {
        const int q = ew * (kNQ / 4) + i;
        const int c = cnt_s[q];
        if (c > kCap - kTileM) compact_list(my_cand + q * kCap, c, p.k, lane, key, thrkey_s, thr_s, cnt_s, q);
      }
      ptx::named_bar_sync(1, 128);
    }

    // ---------------- final: sorted per-CTA top-k for every query ----------------
#pragma unroll 1
    for (int i = 0; i < kNQ / 4; ++i) {
      const int q = ew * (kNQ / 4) + i;
      const int c = cnt_s[q];
      compact_list(my_cand + q * kCap, c, p.k, lane, key, thrkey_s, thr_s, cnt_s, q);
      const int nvalid = c < p.k ? c : p.k;
      float* out_s = p.part_scores + (static_cast<size_t>(blockIdx.x) * kNQ + q) * p.k;
      int64_t* out_i = p.part_ids + (static_cast<size_t>(blockIdx.x) * kNQ + q) * p.k;
#pragma unroll
      for (int e = 0; e < kSortE; ++e) {
        const int pos = lane * kSortE + e;
        if (pos < p.k) {
          const bool ok = pos < nvalid;
          const uint32_t r = 0xFFFFFFFFu - static_cast<uint32_t>(key[e]);
          out_s[pos] = ok ? ord_to_f32(static_cast<uint32_t>(key[e] >> 32)) : -INFINITY;
          out_i[pos] = ok ? p.id_base + static_cast<int64_t>(r) * p.id_stride : -1;
        }
      }
    }
  }

  // ---------------- teardown ----------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

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
