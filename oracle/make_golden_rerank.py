"""TEST INFRASTRUCTURE ONLY — freezes golden vectors of the re-rank tail from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  ``python oracle/make_golden_rerank.py``

``RAG.retrieve_with_rerank`` (src/rag.py:176-246) is executed on the host by
``oracle/ref_import.run_reference_rerank`` (stand-in encoder that returns the rows of a seeded embedding table;
everything after the encoder is the reference's code).  Written to ``tests/golden/rerank_<case>.npz``: inputs,
returned positions (passage ids), scores, gathered embeddings and the two MRR statistics.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name, B, L, D, topk, dtype, seed
CASES = [
    ("rerank_fp32_b6_l40_d128_k8", 6, 40, 128, 8, torch.float32, 31),
    ("rerank_fp32_b3_l100_d768_k100", 3, 100, 768, 100, torch.float32, 32),     # full sort, real dim
    ("rerank_bf16_b5_l64_d64_k16", 5, 64, 64, 16, torch.bfloat16, 33),          # bf16-rounded scores -> near-ties
    ("rerank_fp32_b2_l1_d24_k1", 2, 1, 24, 1, torch.float32, 34),               # single candidate, odd dim
]


def main():
    for name, b, n_cand, d, k, dtype, seed in CASES:
        g = torch.Generator().manual_seed(seed)
        q = torch.nn.functional.normalize(torch.randn(b, d, generator=g), dim=-1).to(dtype)
        p = torch.nn.functional.normalize(torch.randn(b, n_cand, d, generator=g), dim=-1).to(dtype)
        out_p, out_s, _, emb, stats = ref_import.run_reference_rerank(q, p, k)
        pos = np.array([[doc["id"] - i * n_cand for doc in row] for i, row in enumerate(out_p)], dtype=np.int64)
        # the gathered embeddings are stored for the small cases; a checksum of them for the big one
        emb_np = emb.float().numpy()
        extra = dict(emb=emb_np) if emb_np.size <= 50_000 else dict(emb_sum=emb_np.astype(np.float64).sum(-1))
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + ".npz"),
            query_emb=q.float().numpy(), passage_emb=p.float().numpy(), dtype=str(dtype).replace("torch.", ""),
            topk=k, positions=pos, scores=np.array(out_s, dtype=np.float32), **extra,
            mrr=stats["MRR"][0], mrr_rev=stats["MRR_rev"][0], torch_version=torch.__version__)
        print(name, pos.shape, stats)


if __name__ == "__main__":
    main()
