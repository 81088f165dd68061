"""Post-filter of over-retrieved results — same contract as the reference's task filter
(src/tasks/base.py:96-148, handed to retrieval as ``filtering_fun`` at train.py:223 and applied at
src/rag.py:165-166 after ``search_knn(query_emb, topk * filtering_overretrieve_ratio)``)."""
import logging

logger = logging.getLogger(__name__)


def filter_results_by_id(batch_metadata, passages, scores, topk, training=None):
    """Removes, per query, the passages whose ``id`` equals the source example's ``metadata["id"]``
    (so a model cannot retrieve the passage it is denoising); if fewer than ``topk`` remain, the
    violating passages are appended back, with a warning.  Returns (passages, scores) cut to topk."""
    if batch_metadata is None:
        logger.warning("Trying to filter a batch with no metadata - probably a padding instance - just return the topk")
        return [ps[:topk] for ps in passages], [ss[:topk] for ss in scores]
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
