"""The peer-exchange kernels (xchg_push / xchg_merge / xchg_gather — the cross-rank step of reference
src/index.py:135-157) driven on ONE GPU: W exchange objects of one process are wired by device pointer
(mips_xchg_connect_local) and stepped rank by rank — all W pushes of a step, then the W waits — so no kernel ever
waits for a later launch.  Checked against the oracle merge (concatenate in rank order + torch.topk)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import flat_index_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


class LocalRanks:
    """W `mips_xchg` objects on one device, connected without IPC."""

    def __init__(self, eng, world, capacity, device):
        self.lib = eng._native.load()
        self.world, self.dev = world, device
        self.h = []
        for r in range(world):
            h = ctypes.c_void_p()
            assert self.lib.mips_xchg_create(ctypes.byref(h), device.index or 0, r, world, capacity) == 0
            self.h.append(h)
        arr = (ctypes.c_void_p * world)(*[h.value for h in self.h])
        for h in self.h:
            assert self.lib.mips_xchg_connect_local(h, arr, world) == 0, self.lib.mips_xchg_last_error(h)
        self.stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)

    def close(self):
        torch.cuda.synchronize(self.dev)
        for h in self.h:
            self.lib.mips_xchg_destroy(h)
        self.h = []

    def push(self, r, block):
        return self.lib.mips_xchg_push(self.h[r], ctypes.c_void_p(block.data_ptr()), block.numel(), self.stream)

    def merge_wait(self, r, block_bytes, s_bytes, batch, k_in, k_out):
        out_s = torch.empty((batch, k_out), dtype=torch.float32, device=self.dev)
        out_i = torch.empty((batch, k_out), dtype=torch.int64, device=self.dev)
        rc = self.lib.mips_xchg_merge_wait(self.h[r], block_bytes, s_bytes, batch, k_in, k_out,
                                           ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()), self.stream)
        return rc, out_s, out_i

    def gather_wait(self, r, block_bytes):
        out = torch.empty((self.world, block_bytes), dtype=torch.uint8, device=self.dev)
        rc = self.lib.mips_xchg_gather_wait(self.h[r], block_bytes, ctypes.c_void_p(out.data_ptr()), self.stream)
        return rc, out


def _rank_lists(world, batch, k, gen, dev, ties=False):
    """Per rank: sorted-descending fp32 scores [B, k] and distinct global ids (round-robin sharding: id % W == rank)."""
    scores, ids = [], []
    for r in range(world):
        s = torch.randn(batch, k, generator=gen, device=dev)
        if ties:
            s = (s * 4).round() / 4          # many equal scores across ranks: order must be score desc, id asc
        loc = torch.stack([torch.randperm(100_000, generator=gen, device=dev)[:k] for _ in range(batch)])
        gid, _ = torch.sort((loc * world + r).to(torch.int64), dim=1)
        # a rank's list arrives in the engine's total order: score descending, id ascending among equal scores
        s, _ = torch.sort(s, dim=1, descending=True, stable=True)
        scores.append(s.contiguous())
        ids.append(gid.contiguous())
    return scores, ids


@pytest.mark.parametrize("world,batch,k", [(2, 64, 100), (4, 64, 100), (8, 64, 100), (8, 5, 1000), (4, 3, 10), (2, 130, 128)])
def test_push_merge_matches_oracle_over_40_steps(eng, dev, world, batch, k):
    from importlib import import_module
    packed = import_module("jsa-rag_b200.engine").packed_result_buffer
    s_bytes = (batch * k * 4 + 7) // 8 * 8
    block = s_bytes + batch * k * 8
    ranks = LocalRanks(eng, world, block, dev)
    gen = torch.Generator(device=dev).manual_seed(1000 * world + k)
    try:
        for step in range(40):                                   # slot parity alternates; epochs keep counting
            scores, ids = _rank_lists(world, batch, k, gen, dev, ties=(step % 5 == 4))
            bufs = []
            for r in range(world):
                buf, vs, vi = packed(batch, k, dev)
                vs[0].copy_(scores[r]); vi[0].copy_(ids[r])
                bufs.append(buf)
            for r in range(world):
                assert ranks.push(r, bufs[r][0]) == 0
            outs = [ranks.merge_wait(r, block, s_bytes, batch, k, k) for r in range(world)]
            torch.cuda.synchronize()
            ref_s, ref_i = O.merge_rank_results([s.cpu() for s in scores], [i.cpu() for i in ids], k)
            pool_s = torch.cat([s.cpu() for s in scores], dim=1).numpy()
            pool_i = torch.cat([i.cpu() for i in ids], dim=1).numpy()
            # rows whose candidate pool holds no two equal scores have ONE right answer; elsewhere torch.topk's order
            # among ties is unspecified and ours is (score desc, id asc)
            tie_free = np.array([len(np.unique(row)) == row.size for row in pool_s])
            for r, (rc, out_s, out_i) in enumerate(outs):
                assert rc == 0
                got_s, got_i = out_s.cpu(), out_i.cpu()
                assert torch.equal(got_s, ref_s), f"step {step} rank {r}: merged scores differ from the oracle"
                assert torch.equal(got_i[tie_free], ref_i[tie_free]), f"step {step} rank {r}: merged ids differ from the oracle"
                for b in np.nonzero(~tie_free)[0]:
                    row_s, row_i = got_s[b].numpy(), got_i[b].numpy()
                    assert (np.lexsort((row_i, -row_s)) == np.arange(k)).all(), "ties must be ordered by ascending id"
                    pool = set(zip(pool_s[b].tolist(), pool_i[b].tolist()))
                    assert set(zip(row_s.tolist(), row_i.tolist())) <= pool
                    kth = ref_s[b, -1].item()
                    mine = {(x, int(y)) for x, y in zip(row_s.tolist(), row_i.tolist()) if x > kth}
                    theirs = {(x, int(y)) for x, y in zip(ref_s[b].tolist(), ref_i[b].tolist()) if x > kth}
                    assert mine == theirs
                    # among the candidates tied at the k-th score the smallest ids win
                    tied_pool = sorted(int(y) for x, y in pool if x == kth)
                    tied_mine = sorted(int(y) for x, y in zip(row_s.tolist(), row_i.tolist()) if x == kth)
                    assert tied_mine == tied_pool[:len(tied_mine)]
        for h in ranks.h:
            assert ranks.lib.mips_xchg_status(h) == 0
    finally:
        ranks.close()


@pytest.mark.parametrize("world", [2, 8])
def test_push_gather_is_an_all_gather(eng, dev, world):
    nbytes = 64 * 768 * 4 // world // 8 * 8
    ranks = LocalRanks(eng, world, nbytes, dev)
    gen = torch.Generator(device=dev).manual_seed(5)
    try:
        for step in range(12):
            blocks = [torch.randint(0, 256, (nbytes,), dtype=torch.uint8, generator=gen, device=dev) for _ in range(world)]
            for r in range(world):
                assert ranks.push(r, blocks[r]) == 0
            outs = [ranks.gather_wait(r, nbytes) for r in range(world)]
            torch.cuda.synchronize()
            want = torch.stack(blocks)
            for rc, out in outs:
                assert rc == 0 and torch.equal(out, want)
    finally:
        ranks.close()


def test_late_peer_times_out_without_killing_the_context(eng, dev):
    """A peer that never pushes: the wait kernel gives up after the (here 150 ms) timeout, writes padding, raises
    the exchange's error word — no trap, the CUDA context keeps working, later calls report MIPS_ETIMEOUT."""
    N = eng._native
    batch, k = 4, 10
    s_bytes = (batch * k * 4 + 7) // 8 * 8
    block = s_bytes + batch * k * 8
    ranks = LocalRanks(eng, 2, block, dev)
    try:
        assert ranks.lib.mips_xchg_set_timeout_ms(ranks.h[0], 150) == 0
        buf = torch.zeros(block, dtype=torch.uint8, device=dev)
        assert ranks.push(0, buf) == 0                        # rank 1 stays silent
        rc, out_s, out_i = ranks.merge_wait(0, block, s_bytes, batch, k, k)
        assert rc == 0                                        # the launch itself is fine
        torch.cuda.synchronize()                              # returns after ~150 ms, no launch failure
        assert ranks.lib.mips_xchg_status(ranks.h[0]) == N.MIPS_ETIMEOUT
        assert bool((out_i == -1).all()) and bool(torch.isinf(out_s).all())
        assert int(torch.arange(10, device=dev).sum().item()) == 45          # context alive
        assert ranks.push(0, buf) == N.MIPS_ETIMEOUT                          # sticky: the exchange must be rebuilt
        assert b"timeout" in ranks.lib.mips_xchg_last_error(ranks.h[0])
        assert ranks.lib.mips_xchg_status(ranks.h[1]) == 0
    finally:
        ranks.close()
