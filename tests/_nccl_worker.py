"""Run under torchrun on >= 2 GPUs: the real distributed search (NCCL all-gathers + device merge) against a
single-GPU search over the un-sharded index."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = importlib.import_module("jsa-rag_b200")
    n, d, k = 200_003, 768, 100
    g = torch.Generator(device=dev).manual_seed(7)
    e = torch.nn.functional.normalize(torch.randn(n, d, generator=g, device=dev), dim=1).half()
    q_all = torch.nn.functional.normalize(torch.randn(37, d, generator=g, device=dev), dim=1)
    sizes = [37 // world + (1 if r < 37 % world else 0) for r in range(world)]
    offs = [sum(sizes[:r]) for r in range(world + 1)]
    my_q = q_all[offs[rank]:offs[rank + 1]]

    # reference answer: one engine over the whole index on this GPU
    full = eng.MipsEngine(d, torch.float16, dev)
    full.bind(e)
    fs, fi = full.search(q_all, k)

    passages = [{"id": str(i), "text": f"passage {i}"} for i in range(rank, n, world)]      # src/index_io.py:41
    index = eng.B200Index()
    index.init_embeddings(passages, dim=d)
    index.embeddings[:, :] = e[rank::world].T
    index._xchg_min_bytes = 1024            # start with a tight exchange so that the k = 1000 search below has to grow it
    s, i = index.search(my_q, k)
    assert torch.equal(i, fi[offs[rank]:offs[rank + 1]]), "distributed ids differ from the single-GPU answer"
    assert torch.equal(s, fs[offs[rank]:offs[rank + 1]]), "distributed scores differ from the single-GPU answer"

    # the exchange step ran as NVLink peer stores + wait-and-merge (mips_xchg_merge); the NCCL all-gather +
    # mips_merge_topk path must give the same bits
    p2p = bool(getattr(index, "_xchg", None))
    if os.environ.get("JSA_MIPS_EXCHANGE", "p2p") == "p2p" and os.environ.get("JSA_REQUIRE_P2P", "1") == "1":
        assert p2p, "peer exchange was not set up on this box"
    saved, index._xchg = index._xchg, False
    s_n, i_n = index.search(my_q, k)
    index._xchg = saved
    assert torch.equal(i_n, i) and torch.equal(s_n, s)
    # a straggler: the last rank is 15 s late for one search (a checkpoint write, a GC pause).  Its peers' wait
    # kernels keep polling (the bound is 30 minutes of wall-clock time, nothing traps) and the answer is unchanged.
    if p2p and os.environ.get("JSA_TEST_STRAGGLER", "1") == "1":
        import time
        torch.cuda.synchronize()
        dist.barrier()
        if rank == world - 1:
            time.sleep(15.0)
        s_l, i_l = index.search(my_q, k)
        torch.cuda.synchronize()
        assert torch.equal(i_l, i) and torch.equal(s_l, s), "search after a 15 s straggler differs"
        index._xchg.status()
    # back-to-back searches without host syncs: slots alternate, a fast rank may run ahead by one step
    outs = []
    for t in range(40):
        outs.append(index.search(my_q.roll(t, 0), k)[1])
    for t, o in enumerate(outs):
        assert torch.equal(o, fi[offs[rank]:offs[rank + 1]].roll(t, 0)), f"step {t} differs"
    # growing the exchange (bigger blocks: k = 1000) is collective and transparent
    cap0 = index._xchg.capacity if p2p else 0
    fs2, fi2 = full.search(q_all, 1000)
    s2, i2 = index.search(my_q, 1000)
    assert torch.equal(i2, fi2[offs[rank]:offs[rank + 1]]) and torch.equal(s2, fs2[offs[rank]:offs[rank + 1]])
    assert not p2p or index._xchg.capacity > cap0
    s2, i2 = index.search(my_q, k)
    assert torch.equal(i2, i)

    docs, scores = index.search_knn(my_q, k)
    ids = torch.tensor([[int(x["id"]) for x in row] for row in docs])
    assert torch.equal(ids, fi[offs[rank]:offs[rank + 1]].cpu())
    assert all(x["text"] == f"passage {x['id']}" for row in docs for x in row)
    assert "node-shared passage store" in index.last_passage_path          # remote winners: no text over NVLink
    os.environ["JSA_MIPS_PASSAGES"] = "a2a"                                 # ranks on different hosts: all-to-all of winners
    docs_a, scores_a = index.search_knn(my_q, k)
    os.environ.pop("JSA_MIPS_PASSAGES")
    assert docs_a == docs and scores_a == scores and "all-to-all" in index.last_passage_path

    d3, s3, emb = index.search_knn(my_q, 7, return_embeddings=True)                         # build_server/index.py:217-261
    ids3 = torch.tensor([[int(x["id"]) for x in row] for row in d3], device=dev)
    assert torch.equal(emb, e[ids3])

    # CUDA-graph replay of the whole distributed search (collectives captured) equals the single-GPU answer
    nb = min(sizes)
    ref_rows = torch.cat([fi[offs[r]:offs[r] + nb] for r in range(world)])[rank * nb:(rank + 1) * nb]
    # equal per-rank batches: the query all-gather also runs over the peer exchange (mips_xchg_gather)
    for t in range(6):
        es, ei = index.search(my_q[:nb].roll(t, 0), k)
        assert torch.equal(ei, ref_rows.roll(t, 0)), "equal-batch search differs from the single-GPU answer"
    assert not p2p or bool(getattr(index, "_xchg_q", None)), "queries did not go through the peer exchange"
    run = index.make_graphed_search(nb, k)
    for _ in range(2):
        gs, gi = run(my_q[:nb])
        torch.cuda.synchronize()
        assert torch.equal(gi, ref_rows), "graphed search differs from the single-GPU answer"
    run.release()
    del run, gs, gi

    # replicated queries (a front end broadcasting a request): no query exchange, every rank gets every row
    rs_, ri_ = index.search(q_all, k, replicated=True)
    assert torch.equal(ri_, fi) and torch.equal(rs_, fs), "replicated-query search differs from the single-GPU answer"
    run = index.make_graphed_search(q_all.shape[0], k, replicated=True)
    gs, gi = run(q_all)
    torch.cuda.synchronize()
    assert torch.equal(gi, fi)
    run.release()
    del run, gs, gi

    # uneven / empty local batches still take part in the collectives
    d4, s4 = index.search_knn(my_q[:0] if rank == world - 1 else my_q, k)
    assert (d4 == [] and s4 == []) if rank == world - 1 else len(d4) == my_q.shape[0]
    index.close_exchange()
    dist.barrier()
    if rank == 0:
        print(f"nccl worker ok: world={world} exchange={'p2p' if p2p else 'nccl'}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
