"""Client of the index server — mirrors reference src/post.py (``call_retrieve_api``)."""
import requests
import torch

DEFAULT_URL = "http://127.0.0.1:29501/retrieve"   # the reference hard-codes its cluster host (src/post.py:21)


def call_retrieve_api(query_embs=None, topk=10, url: str = DEFAULT_URL, session=None):
    """POSTs the flattened fp32 query embeddings; returns (docs, scores) on HTTP 200, prints and
    returns None otherwise — exactly like src/post.py:6-31."""
    bsz = query_embs.size(0)
    query_embs = query_embs.to(torch.float32)
    query_emb_list = query_embs.cpu().numpy().flatten().tolist()
    data = {"query_embs": query_emb_list, "bsz": bsz, "topk": topk}
    response = (session or requests).post(url, json=data)
    if response.status_code == 200:
        results = response.json()
        return results[0], results[1]
    print(f"请求失败，状态码: {response.status_code}")
    return None
