/*
 * jsa_mips.h — C ABI of the B200-native exact maximum-inner-product-search engine.
 *
 * This is the drop-in boundary for the passage-retrieval hot path of Caohy23/JSA-RAG.  The
 * reference reaches that path through a duck-typed Python object (DistributedIndex,
 * reference src/index.py:44-161), not through an FFI; every entry point below replaces the
 * library call(s) named next to it, and the Python host layer in jsa-rag_b200/ binds them
 * with ctypes (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every function returns 0 on success and a negative MIPS_E* code on failure, never throws
 *     or aborts; mips_last_error() returns a human-readable message for the last failure;
 *   - device pointers are *borrowed* for the duration of a call; the handle owns only its own
 *     descriptors and (optionally) an internal workspace;
 *   - all device work is stream-ordered on the cudaStream_t passed as `stream` (void*).
 */
#ifndef JSA_MIPS_H_
#define JSA_MIPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JSA_MIPS_ABI_VERSION 1

/* element types */
#define MIPS_DTYPE_F16 0
#define MIPS_DTYPE_BF16 1
#define MIPS_DTYPE_F32 2 /* queries only */

/* error codes */
#define MIPS_OK 0
#define MIPS_EINVAL (-1)      /* bad argument (NULL, negative size, unsupported dim/dtype) */
#define MIPS_EKRANGE (-2)     /* k > number of indexed rows: reference raises RuntimeError("selected index k out of range") */
#define MIPS_ECUDA (-3)       /* a CUDA runtime/driver call failed */
#define MIPS_ENOTBOUND (-4)   /* search before mips_bind_index */
#define MIPS_EWORKSPACE (-5)  /* caller-provided workspace too small */
#define MIPS_EUNSUPPORTED (-6)/* not running on an sm_100 device / feature not built */
#define MIPS_ETIMEOUT    (-7) /* peer exchange: a rank's block did not arrive within the exchange timeout */

typedef struct mips_handle mips_handle;

/* Limits of this build (queryable so host code never hard-codes them). */
int mips_abi_version(void);
int mips_max_k(void);          /* largest supported top-k (fused select) */
int mips_max_dim(void);        /* largest supported embedding dimension */

/*
 * Creates an engine bound to CUDA device `device` for `dim`-dimensional embeddings stored as
 * `index_dtype` (MIPS_DTYPE_F16 — the reference's only storage type, src/index.py:52 — or
 * MIPS_DTYPE_BF16 for the JSA bf16 configuration).  dim must be a multiple of 64.
 * Replaces: DistributedIndex.__init__ (src/index.py:45-48).
 */
int mips_create(mips_handle** out, int device, int dim, int index_dtype);
void mips_destroy(mips_handle* h);
const char* mips_last_error(const mips_handle* h); /* h may be NULL: message of the last failed mips_create */

/*
 * Binds the passage-embedding matrix of this rank's shard.  `emb` is a device pointer to
 * n_local rows of `dim` elements, row stride `ld` elements (K-major: one passage per row; this is
 * the transposed *view* of the reference's [dim, n_local] `.embeddings`, src/index.py:52,
 * so `index.embeddings[:, a:b] = x.T` (src/rag.py:120) writes straight into it).
 * Global passage id of local row r is id_base + r * id_stride
 *   round-robin jsonl sharding (src/index_io.py:41):  id_base = rank, id_stride = world_size
 *   contiguous shard files      (src/index.py:97-100): id_base = first row, id_stride = 1
 * Rebinding is cheap (re-encodes one TMA descriptor) and must be repeated whenever the tensor
 * is re-allocated.  Replaces: the `self.embeddings` operand of torch.matmul (src/index.py:118).
 */
int mips_bind_index(mips_handle* h, const void* emb, int64_t n_local, int64_t ld,
                    int64_t id_base, int64_t id_stride);

/*
 * Same, with an explicit storage layout: layout 1 = [n_local, dim] rows (ld >= dim); layout 0 =
 * [dim, n_local] — the reference's own `.embeddings` layout (src/index.py:52), ld >= n_local — which
 * is consumed as an MN-major tcgen05 operand straight from the reference-format tensor (zero copy).
 */
int mips_bind_index_layout(mips_handle* h, const void* emb, int64_t n_local, int64_t ld, int layout,
                           int64_t id_base, int64_t id_stride);

/* Bytes of device workspace mips_search_local needs for up to max_batch queries and top max_k. */
int mips_workspace_bytes(const mips_handle* h, int max_batch, int max_k, size_t* out);

/*
 * Lifetime of the INTERNAL workspace (used when `workspace` is NULL): it grows to the largest (batch, k) seen.  A
 * captured CUDA graph of a search holds raw pointers into it, so the owner of such a graph calls
 * mips_workspace_pin(h, +1) before capture and mips_workspace_pin(h, -1) after destroying the graph: while the pin
 * count is positive an outgrown workspace is retired (kept allocated) instead of freed, and freed when the count
 * returns to zero.
 */
int mips_workspace_pin(mips_handle* h, int delta);

/*
 * Exact top-k inner-product search of `batch` queries over the bound shard.
 *   queries    device pointer, [batch, dim] row-major with row stride q_ld elements, q_dtype
 *              F32 / F16 / BF16; cast to the index dtype exactly like `allqueries.half()`
 *              (src/index.py:118).  normalize != 0 L2-normalises each query first in fp32
 *              (faiss.normalize_L2, build_server/server_start.py:142).
 *   out_scores device [batch, k] fp32, descending (fp32 accumulator values, not fp16-rounded)
 *   out_ids    device [batch, k] int64 global passage ids; ties broken by ascending id
 *   workspace  device scratch of at least mips_workspace_bytes(); NULL => the handle uses (and
 *              grows) an internal workspace.
 * Replaces: torch.matmul + torch.topk in _compute_scores_and_indices (src/index.py:114-121),
 * faiss GpuIndexFlatIP.search (build_server/server_start.py:143).  The [batch, n_local] score
 * matrix is never written to memory.
 */
int mips_search_local(mips_handle* h, const void* queries, int q_dtype, int64_t q_ld, int batch, int k,
                      int normalize, float* out_scores, int64_t* out_ids,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * Merges num_lists sorted candidate lists per query into the global top k_out:
 *   scores [num_lists, batch, k_in] fp32 descending, ids [num_lists, batch, k_in] int64
 *   -> out_scores [batch, k_out], out_ids [batch, k_out]   (score desc, id asc on ties).
 * Entries with id < 0 are padding and are ignored.  k_in, k_out <= mips_max_k().
 * Replaces: the 2*W gathers + concat + second torch.topk of search_knn (src/index.py:135-157)
 * after ONE all-gather of (score, id) candidates.
 */
int mips_merge_topk(int device, const float* scores, const int64_t* ids, int num_lists, int batch,
                    int k_in, int k_out, float* out_scores, int64_t* out_ids, void* stream);

/* Same with independent list strides (in elements) for the score and id arrays, so that both can
 * live in ONE all-gathered buffer per rank ([scores | ids] blocks) and the exchange is a single collective. */
int mips_merge_topk_strided(int device, const float* scores, const int64_t* ids, int num_lists,
                            int64_t score_list_stride, int64_t id_list_stride, int batch, int k_in, int k_out,
                            float* out_scores, int64_t* out_ids, void* stream);

/*
 * out[i, :] = embeddings[local_rows[i], :]   (index dtype, [n, dim] row-major).
 * Replaces: self.embeddings[:, indices.view(-1)] of the 3-tuple search_knn
 * (build_server/index.py:228-229).
 */
int mips_gather_rows(mips_handle* h, const int64_t* local_rows, int64_t n, void* out, void* stream);

/*
 * Peer exchange: the exchange step of the row-sharded search fused with the merge, over NVLink peer stores
 * (one node, one process per GPU).  Replaces the 2*W gathers + concat + topk of src/index.py:135-157 and the
 * all-gather in front of mips_merge_topk.  Collective: every rank creates an exchange with the same world size
 * and capacity, exports its 64-byte IPC handle, gives the handles of all ranks (rank order) to mips_xchg_connect,
 * and then calls mips_xchg_merge once per search with the same (batch, k_in, k_out):
 *   local_block = this rank's [fp32 scores [batch,k_in] padded to score_bytes | int64 ids [batch,k_in]]
 *   -> out_scores / out_ids [batch, k_out] = merge of all W ranks' blocks (score desc, id asc), on every rank.
 * Two launches per call (push, wait+merge), stream-ordered, capturable in a CUDA graph.
 * mips_xchg_connect returns MIPS_EUNSUPPORTED when the GPUs cannot map each other's memory (callers then use the
 * all-gather + mips_merge_topk path).
 */
typedef struct mips_xchg mips_xchg;
int mips_xchg_handle_bytes(void);
int mips_xchg_create(mips_xchg** out, int device, int rank, int world, size_t block_capacity_bytes);
int mips_xchg_export(mips_xchg* x, void* out_handle);
int mips_xchg_connect(mips_xchg* x, const void* all_handles);
size_t mips_xchg_capacity(mips_xchg* x);
int mips_xchg_merge(mips_xchg* x, const void* local_block, size_t block_bytes, size_t score_bytes, int batch, int k_in,
                    int k_out, float* out_scores, int64_t* out_ids, void* stream);
/* Plain all-gather over the same mechanism (the query all-gather of src/index.py:128): every rank's `block_bytes`
 * (equal on all ranks, multiple of 8) -> out [W, block_bytes] in rank order on every rank.  Two launches. */
int mips_xchg_gather(mips_xchg* x, const void* local_block, size_t block_bytes, void* out, void* stream);
/* The two halves of mips_xchg_merge / mips_xchg_gather as separate calls (one launch each): a rank may push as soon
 * as its block is ready and wait later.  Every push must be followed by exactly one *_wait before the next push. */
int mips_xchg_push(mips_xchg* x, const void* local_block, size_t block_bytes, void* stream);
int mips_xchg_merge_wait(mips_xchg* x, size_t block_bytes, size_t score_bytes, int batch, int k_in, int k_out,
                         float* out_scores, int64_t* out_ids, void* stream);
int mips_xchg_gather_wait(mips_xchg* x, size_t block_bytes, void* out, void* stream);
/* Stragglers: the receiving kernels wait for the peers' blocks in wall-clock time — 30 minutes by default (the order
 * of a collective watchdog; the reference tolerates 100000 s, src/slurm.py:181), JSA_MIPS_XCHG_TIMEOUT_S or
 * mips_xchg_set_timeout_ms change it.  On expiry nothing traps: the kernel writes padding (score -inf, id -1; a
 * gather leaves `out` untouched), raises a host-visible error word, and this and every later call on the exchange
 * returns MIPS_ETIMEOUT (mips_xchg_status reads the word without launching anything).  The CUDA context stays usable;
 * the caller tears the exchange down and falls back to the all-gather path. */
int mips_xchg_set_timeout_ms(mips_xchg* x, int64_t ms);
int mips_xchg_status(mips_xchg* x);
/* Test wiring: connects W exchanges that live in ONE process on ONE device by their device pointers (no IPC), so
 * that the push / merge / gather kernels can be driven rank by rank on a single GPU: launch all W pushes of a step
 * first, then the W waits (a wait launched before its peers' pushes would only sit out its timeout). */
int mips_xchg_connect_local(mips_xchg* x, mips_xchg* const* all, int n);
const char* mips_xchg_last_error(mips_xchg* x);
int mips_xchg_destroy(mips_xchg* x);

/*
 * Re-rank of a short candidate list per query (one launch):
 *   s[b, j] = <queries[b, :], cand[b, j, :]>   (fp32 accumulation), j < num_cand <= mips_max_rerank_candidates()
 *   out_scores [batch, k] fp32 / out_pos [batch, k] int64 = the k best (score desc, position asc on ties),
 *   out_rank   [batch, num_cand] int64 (optional, may be NULL) = rank of every candidate in the sorted order,
 *   out_emb    [batch, k, dim] (optional, may be NULL; same dtype as cand) = cand[b, out_pos[b, j], :].
 * queries [batch, dim] with row stride q_ld elements and cand [batch, num_cand, dim] contiguous share `dtype`
 * (MIPS_DTYPE_F16 / BF16 / F32).  k > num_cand -> MIPS_EKRANGE.
 * Replaces: einsum("id,ijd->ij") + torch.sort + slice + torch.gather of RAG.retrieve_with_rerank
 * (src/rag.py:228-233) and the re-selection of the 3-tuple search_knn (build_server/index.py:253-255).
 */
int mips_rerank(int device, const void* queries, int64_t q_ld, const void* cand, int dtype, int batch, int num_cand,
                int dim, int k, float* out_scores, int64_t* out_pos, int64_t* out_rank, void* out_emb, void* stream);
int mips_max_rerank_candidates(void);

/*
 * End-to-end convenience for callers that hold HOST buffers (server / ctypes clients):
 * H2D copy of fp32 queries [batch, dim], mips_search_local, D2H copy of results, stream sync.
 * host_* pointers should be page-locked for full speed but need not be.
 */
int mips_search_host(mips_handle* h, const float* host_queries, int batch, int k, int normalize,
                     float* host_scores, int64_t* host_ids, void* stream);
/* Same without the final stream synchronisation: copies and kernels are only enqueued on `stream`; host_scores /
 * host_ids are valid once the stream (or an event recorded after the call) has completed, and host_queries must stay
 * untouched until then.  Lets a server keep the next request in flight while it serialises the previous answer
 * (everything stays ordered on the one stream, so no extra device buffers are involved). */
int mips_search_host_async(mips_handle* h, const float* host_queries, int batch, int k, int normalize,
                           float* host_scores, int64_t* host_ids, void* stream);

/*
 * Host-only helper of the server's JSON route: parses a JSON list of decimal numbers ("[0.12, -3e-4, ...]", the
 * brackets optional) into fp32.  Returns the count, or < 0 (-2: a token that is not a plain number, -3: more than
 * max_out numbers) — callers then fall back to a general JSON parser.
 * Replaces: torch.tensor(request.query_embs) over a pydantic-validated list (build_server/server_start.py:186).
 */
int64_t mips_parse_float_list(const char* text, size_t len, float* out, int64_t max_out);

/* Number of kernels the last mips_search_local / mips_search_host on this handle launched. */
int mips_last_launch_count(const mips_handle* h);

/*
 * Diagnostics (not used on the product path).  flags: 1 = skip the select epilogue, 2 = skip the
 * MMAs (pure TMA streaming) — results are meaningless with either set; 4 = no sampled pre-pass;
 * 8 = time the scan launches; 16 = always UMMA M=128; 32 = one query block per launch; 64 = seed the thresholds
 * with separate sampled scan + select launches instead of inside the scan kernel; 128 = batches > 128 without
 * tcgen05 CTA pairs (the round-1 multi-block path); 256 = the producer hands stages over without loading them
 * (results meaningless: power / latency split); 512 = CTA pairs that share a tile sequence run without the L2
 * lock-step; 2048 = 4 pair blocks (1024 queries) per launch at any index size (automatic only for >= ~23M rows).  Environment switches read once per process: JSA_MIPS_PDL=0 (no programmatic dependent launch),
 * JSA_MIPS_PAIRS=0, JSA_MIPS_PAIR_BLOCKS=1|2|4 (0 = automatic), JSA_MIPS_LOCK_WINDOW=<tiles>, JSA_MIPS_NVTX=1.  stats_dev: device array of
 * [mips_num_sms()][mips_debug_num_stats()] uint64 per-CTA cycle counters the scan kernel fills
 * (caller zeroes it), or NULL.  Counter order: producer wait, MMA wait(full), MMA wait(TMEM),
 * epilogue wait(TMEM), epilogue select, epilogue compaction, #compactions, #appends, total cycles, epilogue tcgen05.ld.
 */
int mips_debug_config(mips_handle* h, int flags, void* stats_dev);
/* With flag 8 set, every full-shard scan launch is bracketed by CUDA events on the caller's stream;
 * this returns (and clears) the recorded launch durations in milliseconds (at most 256 kept). */
int mips_scan_times_ms(mips_handle* h, float* out, int max_out, int* n_out);
int mips_debug_num_stats(void);
int mips_num_sms(const mips_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* JSA_MIPS_H_ */
