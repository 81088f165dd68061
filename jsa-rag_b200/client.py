"""HTTP client of the index server.

``call_retrieve_api(query_embs, topk)`` keeps the calling convention of the reference's client
(src/post.py:6-31): it returns ``(docs, scores)`` when the server answers 200 and, like the
reference, reports the status code on stdout and returns ``None`` otherwise.  The server address is
configurable here (the reference hard-codes its cluster host, src/post.py:21).
"""
from typing import Optional, Tuple

import requests
import torch

DEFAULT_URL = "http://127.0.0.1:29501/retrieve"


def encode_request(query_embs: torch.Tensor, topk: int) -> dict:
    """JSON body of POST /retrieve: the fp32 embeddings flattened row-major, their count and k
    (RetrieveRequest, build_server/server_start.py:18-21)."""
    flat = query_embs.detach().to(device="cpu", dtype=torch.float32).reshape(-1)
    return {"query_embs": flat.tolist(), "bsz": int(query_embs.shape[0]), "topk": int(topk)}


class RetrieveClient:
    """Keeps one HTTP session (connection reuse) to an index server."""

    def __init__(self, url: str = DEFAULT_URL, session=None):
        self.url = url
        self.session = session if session is not None else requests.Session()

    def retrieve(self, query_embs: torch.Tensor, topk: int = 10) -> Optional[Tuple[list, list]]:
        reply = self.session.post(self.url, json=encode_request(query_embs, topk))
        if reply.status_code != 200:
            print(f"请求失败，状态码: {reply.status_code}")   # same message as the reference client
            return None
        docs, scores = reply.json()[:2]
        return docs, scores

    def retrieve_binary(self, query_embs: torch.Tensor, topk: int = 10, half: bool = False):
        """POST /retrieve_bin: raw fp32 (or fp16) query bytes instead of a JSON float list."""
        q = query_embs.detach().to(device="cpu", dtype=torch.float16 if half else torch.float32).contiguous()
        url = self.url.rsplit("/", 1)[0] + "/retrieve_bin"
        reply = self.session.post(url, params={"bsz": int(q.shape[0]), "topk": int(topk), "dtype": "fp16" if half else "fp32"},
                                  data=q.numpy().tobytes(), headers={"Content-Type": "application/octet-stream"})
        if reply.status_code != 200:
            print(f"请求失败，状态码: {reply.status_code}")
            return None
        docs, scores = reply.json()[:2]
        return docs, scores

    def search_binary(self, query_embs: torch.Tensor, topk: int = 10, half: bool = False):
        """POST /search_bin: binary both ways — returns (scores fp32 [b, k], passage ids int64 [b, k]) as numpy arrays for
        callers that hold the passage store themselves (76.8 KB instead of ~0.6 MB of JSON for 64 x 100 results)."""
        import numpy as np
        q = query_embs.detach().to(device="cpu", dtype=torch.float16 if half else torch.float32).contiguous()
        b = int(q.shape[0])
        url = self.url.rsplit("/", 1)[0] + "/search_bin"
        reply = self.session.post(url, params={"bsz": b, "topk": int(topk), "dtype": "fp16" if half else "fp32"},
                                  data=q.numpy().tobytes(), headers={"Content-Type": "application/octet-stream"})
        if reply.status_code != 200:
            print(f"请求失败，状态码: {reply.status_code}")
            return None
        raw = reply.content
        scores = np.frombuffer(raw[:b * topk * 4], dtype="<f4").reshape(b, topk)
        ids = np.frombuffer(raw[b * topk * 4:], dtype="<i8").reshape(b, topk)
        return scores, ids


def call_retrieve_api(query_embs=None, topk=10, url: str = DEFAULT_URL, session=None):
    """Drop-in for src/post.py:call_retrieve_api."""
    return RetrieveClient(url, session=session if session is not None else requests).retrieve(query_embs, topk)
