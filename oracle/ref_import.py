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
