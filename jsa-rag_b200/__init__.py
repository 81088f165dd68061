"""jsa-rag_b200 — B200-native exact maximum-inner-product search for the JSA-RAG retrieval path.

The directory name carries a hyphen (repo convention), so import it with
``importlib.import_module("jsa-rag_b200")`` or through the ``jsa_rag_b200`` alias module at the repo
root.  Everything a caller needs is re-exported here.
"""
import sys as _sys

from . import _native, dist_utils  # noqa: F401
from .index import B200Index, B200IndexWithEmbeddings, EMBEDDINGS_DIM  # noqa: F401
from .index_io import load_or_initialize_index, load_passages, save_embeddings_and_index  # noqa: F401
from .filtering import filter_results_by_id  # noqa: F401


def __getattr__(name):  # lazy: engine imports need the CUDA extension
    if name in ("MipsEngine", "merge_topk", "rerank_topk"):
        from . import engine
        return getattr(engine, name)
    if name in ("B200ServerIndex", "IndexHolder", "create_app", "RetrieveRequest", "rebuildRequest",
                "append_embedding_batch", "iter_embedding_stream", "get_pkl_files_in_directory"):
        from . import server
        return getattr(server, name)
    if name == "rerank_passages":
        from . import rerank
        return rerank.rerank_passages
    if name == "call_retrieve_api":
        from . import client
        return client.call_retrieve_api
    raise AttributeError(name)


_sys.modules.setdefault("jsa_rag_b200", _sys.modules[__name__])
