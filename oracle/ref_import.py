"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference index module.

Only usable in the build container (``/root/reference`` does not exist on the
GPU box).  It is used by ``oracle/make_golden.py`` to freeze golden vectors and
by the ``-m "not gpu"`` tests (when the reference tree is present) to validate
the restatement in ``oracle/flat_index_oracle.py``.

The reference does not import as shipped: ``src/index.py:11-12`` needs faiss
and ``src/index.py:16`` pulls ``src/retrievers.py`` -> ``src/modeling_bert.py:44``
which needs a transformers-4.18 symbol.  Two stub modules are injected
(SURVEY.md Appendix A); nothing of the reference is copied.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("JSA_REFERENCE_ROOT", "/root/reference")

_FAISS_NAMES = [
    "GpuIndexIVFFlat", "GpuIndexIVFPQ", "GpuIndexIVFScalarQuantizer", "GpuIndexFlatIP", "IndexPQ",
    "GpuIndexIVFPQConfig", "GpuIndexIVFFlatConfig", "GpuIndexIVFScalarQuantizerConfig",
    "GpuIndexFlatConfig", "GpuMultipleClonerOptions",
]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "index.py"))


def import_reference_index():
    """Returns the reference's ``src.index`` module (with faiss / retrievers stubbed)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if "faiss" not in sys.modules:
        faiss = types.ModuleType("faiss")
        for n in _FAISS_NAMES:  # names dereferenced at import time by src/index.py:18-28,280
            setattr(faiss, n, type(n, (), {}))
        contrib = types.ModuleType("faiss.contrib")
        tu = types.ModuleType("faiss.contrib.torch_utils")
        faiss.contrib, contrib.torch_utils = contrib, tu
        sys.modules.update({"faiss": faiss, "faiss.contrib": contrib, "faiss.contrib.torch_utils": tu})
    if "src.retrievers" not in sys.modules:
        retr = types.ModuleType("src.retrievers")
        retr.EMBEDDINGS_DIM = 768  # src/retrievers.py:14
        sys.modules["src.retrievers"] = retr
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src.index as ref_index  # noqa: E402

    return ref_index


def import_reference_index_io():
    import_reference_index()
    import src.index_io as ref_index_io  # noqa: E402

    return ref_index_io


def make_reference_cpu_index(passages, embeddings_nd, dim=None):
    """Builds the reference DistributedIndex on CPU exactly as its callers do.

    ``embeddings_nd``: torch tensor [N, D]; written with the same slice assignment as
    src/rag.py:120 (``index.embeddings[:, a:b] = emb.T``), which casts to fp16.
    """
    ref = import_reference_index()
    idx = ref.DistributedIndex()
    idx.is_in_gpu = False  # src/index.py:48,53 — keep the fp16 matrix on the host
    idx.init_embeddings(passages, dim=dim if dim is not None else embeddings_nd.shape[1])
    idx.embeddings[:, :] = embeddings_nd.T
    return idx


def import_reference_rag():
    """Returns the reference's ``src.rag`` module (for RAG.retrieve_with_rerank, src/rag.py:176-246).

    On top of the two stubs above it needs: ``turtle`` (src/rag.py:15 imports two unused names; no tkinter
    here), ``UntiedDualEncoderRetriever`` on the retrievers stub (src/rag.py:23), and the grpc / http client
    modules it only calls in server mode (src/rag.py:27-33)."""
    import_reference_index()

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m

    stub("turtle", pos=None, register_shape=None)
    if not hasattr(sys.modules["src.retrievers"], "UntiedDualEncoderRetriever"):
        sys.modules["src.retrievers"].UntiedDualEncoderRetriever = type("UntiedDualEncoderRetriever", (), {})
    stub("rebuildgrpc")
    stub("rebuildgrpc.async_init_build_client", run_retrieve=None, run_build=None)
    stub("post", call_retrieve_api=None)
    import src.rag as ref_rag  # noqa: E402

    return ref_rag


class _HostTensor:
    """A tokenizer output whose ``.cuda()`` stays on the host (src/rag.py:2364-2365 moves every value)."""

    def __init__(self, t):
        self.t = t

    def cuda(self):
        return self


def run_reference_rerank(query_emb, passage_emb, topk, batch_size=7):
    """Runs the UNMODIFIED ``RAG.retrieve_with_rerank`` (src/rag.py:176-246) on the host.

    The encoder is out of scope, so ``self`` is a stand-in whose retriever returns rows of ``passage_emb``
    ([B, L, D]): passage (i, j) has text "i*L+j", ``retriever_format`` is "{text}", the tokenizer maps the
    strings back to row numbers and the retriever looks them up.  Everything after the encoder — einsum, sort,
    slice, gather, the MRR statistics, the output lists — is the reference's own code.
    Returns (output_passages, output_scores, query_emb, topk_passage_embd, iter_stats)."""
    import torch

    rag = import_reference_rag()
    bsz, n_cand, dim = passage_emb.shape
    table = passage_emb.reshape(bsz * n_cand, dim)
    passages = [[{"id": i * n_cand + j, "title": "t", "text": str(i * n_cand + j)} for j in range(n_cand)]
                for i in range(bsz)]

    class Opt:
        n_to_rerank_with_retrieve_with_rerank = n_cand
        retriever_format = "{text}"
        per_gpu_embedder_batch_size = batch_size
        text_maxlength = 512

    class Retriever:
        def eval(self):
            return self

        def __call__(self, idx=None, is_passages=False):
            assert is_passages
            return table[idx.t]

    class Self:
        opt = Opt()

        def _retrieve(self, index, topk_, query, ids, mask, batch_metadata, filtering_fun, iter_stats, posterior):
            assert topk_ == n_cand
            return passages, None, query_emb, None

        def _get_fp16_retriever_copy(self, posterior=False):
            return Retriever()

        def retriever_tokenizer(self, batch, **kw):
            return {"idx": _HostTensor(torch.tensor([int(s) for s in batch], dtype=torch.int64))}

    stats = {}
    fn = rag.RAG.retrieve_with_rerank
    fn = getattr(fn, "__wrapped__", fn)          # peel torch.no_grad
    with torch.no_grad():
        out_p, out_s, q, emb = fn(Self(), None, topk, ["q"] * bsz, None, None, iter_stats=stats)
    return out_p, out_s, q, emb, stats
