"""Worker for the world_size-2 gloo tests.  The fused CUDA search and the device merge are replaced
by oracle-backed TEST DOUBLES (this file lives in tests/; the product has no CPU path) so that the
host-side plumbing — query all-gather, candidate all-gather packing, row-ownership / global ids,
passage resolution, shard file I/O — runs for real across two processes."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def install_test_doubles(index):
    from oracle import flat_index_oracle as O

    def _local_search(allqueries, topk, normalize=False):
        n_local = index._store.shape[0]
        if topk > n_local:
            raise RuntimeError("selected index k out of range")
        s = allqueries.to(index._store.dtype).float() @ index._store.float().t()   # fp32 scores like the engine
        # engine order: score desc, global id asc
        order = torch.argsort(-s.double(), dim=1, stable=True)[:, :topk]
        return torch.gather(s, 1, order), index._id_base + order * index._id_stride

    def _merge_lists(scores, ids, topk):
        L, b, k = scores.shape
        s = scores.permute(1, 0, 2).reshape(b, L * k)
        i = ids.permute(1, 0, 2).reshape(b, L * k)
        key = np.lexsort((i.numpy(), -s.double().numpy()), axis=1)[:, :topk]
        key = torch.from_numpy(key)
        return torch.gather(s, 1, key), torch.gather(i, 1, key)

    index._local_search = _local_search
    index._merge_lists = _merge_lists
    index._gather_embeddings = None
    return index


def worker(rank, world, port, tmpdir, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = importlib.import_module("jsa-rag_b200")
        from conftest import load_golden
        from oracle import flat_index_oracle as O
        g = load_golden("flat_n1003_d768_b8_k20")
        n, k = int(g["n"]), 20
        e = torch.from_numpy(g["embeddings"])
        q = torch.from_numpy(g["queries"])
        exact = O.exact_scores(g["queries"], g["embeddings"])
        passages_all = [{"id": str(i), "title": f"t{i}", "text": f"passage {i}"} for i in range(n)]
        my_q = q[:3] if rank == 0 else q[3:]
        my_rows = slice(0, 3) if rank == 0 else slice(3, 8)

        if mode == "round_robin":
            # load_passages sharding: line c -> rank c % W (src/index_io.py:41)
            jsonl = os.path.join(tmpdir, "p.jsonl")
            if rank == 0:
                import json
                with open(jsonl, "w") as f:
                    for p in passages_all:
                        f.write(json.dumps(p) + "\n")
            dist.barrier()
            mine = eng.load_passages([jsonl])
            assert [int(p["id"]) for p in mine] == list(range(rank, n, world))
            index = eng.B200Index(device="cpu")
            index.init_embeddings(mine, dim=768)
            index.embeddings[:, :] = e[rank::world].T            # write site of src/rag.py:120
        else:
            # save with 1 "virtual" layout then load contiguous shards [r*spw, (r+1)*spw)
            if rank == 0:
                full = eng.B200Index(device="cpu")
                full._id_base, full._id_stride = 0, 1
                full.doc_map = {i: p for i, p in enumerate(passages_all)}
                full._store = e.clone()
                import jsa_rag_b200  # noqa: F401
                # emulate a single-worker save of 4 shards (rank/world are 0/1 inside save via monkeypatch)
                du = eng.dist_utils
                gr, gw = du.get_rank, du.get_world_size
                du.get_rank, du.get_world_size = (lambda: 0), (lambda: 1)
                try:
                    full.save_index(tmpdir, 4)
                finally:
                    du.get_rank, du.get_world_size = gr, gw
            dist.barrier()
            index = eng.B200Index(device="cpu")
            index.load_index(tmpdir, 4)
            assert index._sharding == "contiguous"
            per = [251, 251, 251, 250]
            assert index._store.shape[0] == sum(per[rank * 2:(rank + 1) * 2])
            assert index._id_base == (0 if rank == 0 else 502)

        install_test_doubles(index)
        stats = {}
        index.iter_stats = stats                                  # the trainer's dict (src/rag.py:143,170)
        docs, scores = index.search_knn(my_q, k)
        assert isinstance(stats["runtime/search"], tuple) and stats["runtime/search"][0] > 0 and stats["runtime/search"][1] == 1
        ids = np.array([[int(d["id"]) for d in row] for row in docs], dtype=np.int64)
        rep = O.compare_topk(ids, np.array(scores, dtype=np.float64), g["ids"][my_rows], g["scores"][my_rows].astype(np.float32),
                             exact[my_rows])
        assert rep["ok"], rep["errors"][:3]
        assert all(d["text"] == f"passage {d['id']}" for row in docs for d in row)
        # winners owned by the other rank came out of the node-shared passage store, without a collective ...
        assert "node-shared passage store" in index.last_passage_path
        own = [d for row in docs for d in row if (int(d["id"]) % world == rank if mode == "round_robin" else
                                                  index._id_base <= int(d["id"]) < index._id_base + index._store.shape[0])]
        assert own and all(any(d is v for v in (index.doc_map[(int(d["id"]) - index._id_base) // index._id_stride],)) for d in own), \
            "own-shard winners must be the doc_map objects themselves"
        # ... and the all-to-all of pickled winners (ranks on different hosts) returns the same passages
        os.environ["JSA_MIPS_PASSAGES"] = "a2a"
        docs_a, scores_a = index.search_knn(my_q, k)
        os.environ.pop("JSA_MIPS_PASSAGES")
        assert "all-to-all" in index.last_passage_path and docs_a == docs and scores_a == scores

        # empty batch on one rank: still participates in the collectives, gets ([], [])
        docs2, scores2 = index.search_knn(my_q[:0] if rank == 1 else my_q, k)
        if rank == 1:
            assert docs2 == [] and scores2 == []
        else:
            assert len(docs2) == 3

        # tensor-level API + candidate packing round trip
        s_t, i_t = index.search(my_q, k)
        assert s_t.dtype == torch.float32 and i_t.dtype == torch.int64 and tuple(i_t.shape) == (my_q.shape[0], k)
        gs, gi = eng.dist_utils.all_gather_candidates(s_t[:2], i_t[:2] + (1 << 40))
        assert gs.shape == (world, 2, k) and torch.equal(gi[rank], i_t[:2] + (1 << 40)) and torch.equal(gs[rank], s_t[:2])

        # k larger than the local shard -> RuntimeError like torch.topk in the reference
        try:
            index.search_knn(my_q, 600)
            raise AssertionError("expected RuntimeError")
        except RuntimeError as ex:
            assert "selected index k out of range" in str(ex)
        dist.barrier()
    finally:
        dist.destroy_process_group()
