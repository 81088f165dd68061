"""Index I/O + passage sharding — mirrors reference src/index_io.py (same names and behaviour)."""
from __future__ import annotations

import argparse
import json
import logging

import torch

from . import dist_utils
from .index import B200Index

logger = logging.getLogger(__name__)


def load_passages(filenames, maxload=-1):
    """jsonl -> this rank's passages.  Line ``c`` (counted across files) goes to rank ``c % W``
    (src/index_io.py:36-44); ``title`` gets ``": section"`` appended (src/index_io.py:30-31).
    Like the reference, a blank line owned by this rank contributes a ``None`` entry."""
    counter = 0
    passages = []
    global_rank = dist_utils.get_rank()
    world_size = dist_utils.get_world_size()
    for fname in filenames:
        with open(fname) as fin:
            for line in fin:
                if maxload > -1 and counter >= maxload:
                    break
                if (counter % world_size) == global_rank:
                    ex = None
                    if line.strip() != "":
                        ex = json.loads(line)
                        assert "id" in ex
                        if "title" in ex and "section" in ex and len(ex["section"]) > 0:
                            ex["title"] = f"{ex['title']}: {ex['section']}"
                    else:
                        print("empty line")
                    passages.append(ex)
                counter += 1
    return passages


def save_embeddings_and_index(index, opt: argparse.Namespace) -> None:
    """src/index_io.py:65-69."""
    index.save_index(opt.save_index_path, opt.save_index_n_shards)


def load_or_initialize_index(opt):
    """src/index_io.py:72-95.  ``index_mode`` "flat" (the reference's exact index) and "b200" both
    return the B200-native exact index; "faiss" is accepted only for ``faiss_index_type == "flat"``
    (exact IP, src/index.py:323-325) — the approximate IVF/PQ/SQ types are out of scope."""
    mode = getattr(opt, "index_mode", "flat")
    dtype = torch.bfloat16 if getattr(opt, "index_dtype", "fp16") in ("bf16", "bfloat16") else torch.float16
    if mode in ("flat", "b200"):
        index = B200Index(dtype=dtype)
    elif mode == "faiss":
        if getattr(opt, "faiss_index_type", "flat") != "flat":
            raise ValueError(f"unsupported faiss index type {opt.faiss_index_type}: only exact 'flat' search is provided")
        index = B200Index(dtype=dtype)
    else:
        raise ValueError(f"unsupported index mode {mode}")

    if getattr(opt, "load_index_path", None) is not None:
        logger.info(f"Loading index from: {opt.load_index_path} with index mode: {mode}")
        index.load_index(opt.load_index_path, opt.save_index_n_shards)
        passages = [index.doc_map[i] for i in range(len(index.doc_map))]
    else:
        passages = []
        if not getattr(opt, "use_file_passages", False):
            logger.info(f"Loading passages from: {opt.passages}")
            passages = load_passages(opt.passages, opt.max_passages)
            dim = 1024 if "bge" in str(getattr(opt, "retriever_model_path", "")) else 768   # src/index_io.py:92
            index.init_embeddings(passages, dim=dim)
    return index, passages
