"""Tail of the retrieve-then-rerank path — mirrors reference src/rag.py:228-246.

The reference re-encodes the ``n_to_rerank`` retrieved passages with the current retriever (out of
scope here: that is the encoder), then scores them against the query embedding with
``einsum("id,ijd->ij")``, sorts, keeps ``topk`` and gathers the winners' embeddings.  Everything
after the encoder is ONE launch of ``mips_rerank`` here.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from . import engine


@torch.no_grad()
def rerank_passages(query_emb: torch.Tensor, passage_emb: torch.Tensor, passages: Sequence[Sequence[dict]], topk: int,
                    iter_stats: Dict = None) -> Tuple[List[List[dict]], List[List[float]], torch.Tensor]:
    """(output_passages, output_scores, topk_passage_embd) of RAG.retrieve_with_rerank (src/rag.py:228-246).

    query_emb [B, D]; passage_emb [B, L, D] (or [B*L, D] as the reference builds it, src/rag.py:211-228);
    passages: B lists of L passage dicts in the candidate order.  ``iter_stats`` receives the
    reference's two statistics, "MRR" and "MRR_rev" (src/rag.py:236-240)."""
    bsz = query_emb.shape[0]
    if passage_emb.dim() == 2:
        passage_emb = passage_emb.view(bsz, -1, passage_emb.shape[-1])
    if bsz == 0:
        return [], [], passage_emb.new_zeros((0, topk, passage_emb.shape[-1]))
    scores, pos, rank, emb = engine.rerank_topk(query_emb, passage_emb, topk, want_rank=iter_stats is not None)
    scores = scores.to(query_emb.dtype)             # einsum returns the operands' dtype
    pos_h = pos.tolist()
    out_passages = [[passages[i][j] for j in row] for i, row in enumerate(pos_h)]
    out_scores = scores.tolist()
    if iter_stats is not None:
        # the reference's formulas verbatim in meaning: 1 / mean(position + 1) of the kept candidates,
        # and the same over the new ranks of the first topk candidates
        iter_stats["MRR"] = (1 / (pos.float() + 1).mean(-1).mean().item(), bsz)
        iter_stats["MRR_rev"] = (1 / (rank[:, :topk].float() + 1).mean(-1).mean().item(), bsz)
    return out_passages, out_scores, emb
