"""Post-filter of over-retrieved results — same contract as the reference's task filter
(src/tasks/base.py:96-148, handed to retrieval as ``filtering_fun`` at train.py:223 and applied at
src/rag.py:165-166 after ``search_knn(query_emb, topk * filtering_overretrieve_ratio)``).

The reference walks every (query, passage) pair in Python.  Here the decision is made on an id ARRAY: one
elementwise comparison + one stable argsort give, per query, the positions of the non-violating passages followed by
the violating ones; dicts and scores are only touched to gather the topk outputs."""
import logging

import numpy as np

logger = logging.getLogger(__name__)


def filter_positions(source_ids, candidate_ids, topk):
    """``source_ids`` [b], ``candidate_ids`` [b, k'] (any comparable dtype, e.g. object arrays of strings or int64
    global ids) -> (positions [b, min(topk, k')], kept [b]): per row the indices of the candidates whose id differs
    from the row's source id, in their original order, followed by the violating ones; ``kept`` counts the former."""
    cand = np.asarray(candidate_ids)
    src = np.asarray(source_ids).reshape(-1, 1)
    viol = cand == src                                         # [b, k']
    order = np.argsort(viol, axis=1, kind="stable")            # False (keep) first, original order preserved
    return order[:, :topk], (~viol).sum(axis=1)


def filter_results_by_id(batch_metadata, passages, scores, topk, training=None):
    """Removes, per query, the passages whose ``id`` equals the source example's ``metadata["id"]``
    (so a model cannot retrieve the passage it is denoising); if fewer than ``topk`` remain, the
    violating passages are appended back, with a warning.  Returns (passages, scores) cut to topk."""
    if batch_metadata is None:
        logger.warning("Trying to filter a batch with no metadata - probably a padding instance - just return the topk")
        return [ps[:topk] for ps in passages], [ss[:topk] for ss in scores]
    b = min(len(batch_metadata), len(passages), len(scores))   # zip() semantics of the reference
    if b == 0:
        return [], []
    widths = {len(ps) for ps in passages[:b]}
    if len(widths) != 1 or 0 in widths:
        return _filter_rows(batch_metadata, passages, scores, topk)        # ragged / empty rows: row by row
    ids = np.empty((b, widths.pop()), dtype=object)
    for r in range(b):
        ids[r] = [p["id"] for p in passages[r]]
    src = np.empty(b, dtype=object)
    src[:] = [m["id"] for m in batch_metadata[:b]]
    pos, kept = filter_positions(src, ids, topk)
    for n_kept in kept[kept < topk].tolist():
        logger.warning(f"{n_kept} passages after filtering for topk = {topk}")
    out_p, out_s = [], []
    for r, row in enumerate(pos.tolist()):
        ps, ss = passages[r], scores[r]
        out_p.append(tuple(ps[j] for j in row))
        out_s.append(tuple(ss[j] for j in row))
    return out_p, out_s


def _filter_rows(batch_metadata, passages, scores, topk):
    out_p, out_s = [], []
    for metadata, passage_li, scores_li in zip(batch_metadata, passages, scores):
        src_id = metadata["id"]
        keep = [(p, s) for p, s in zip(passage_li, scores_li) if p["id"] != src_id]
        viol = [(p, s) for p, s in zip(passage_li, scores_li) if p["id"] == src_id]
        if topk > len(keep):
            logger.warning(f"{len(keep)} passages after filtering for topk = {topk}")
        merged = keep + viol
        ps, ss = zip(*merged) if merged else ((), ())
        out_p.append(ps[:topk])
        out_s.append(ss[:topk])
    return out_p, out_s
