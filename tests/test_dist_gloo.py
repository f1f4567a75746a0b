"""World-size-2 tests of the sharding layer on CPU (gloo backend): partitioning, gathers, restart-sharded fit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from bobe_b200 import dist as bd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_match_array_split():
    for total in (0, 1, 7, 64, 1000003):
        for ws in (1, 2, 3, 8):
            chunks = np.array_split(np.arange(total), ws)
            for r in range(ws):
                lo, hi = bd.shard_bounds(total, r, ws)
                assert hi - lo == len(chunks[r]) and (len(chunks[r]) == 0 or chunks[r][0] == lo)


class _FakeGP:
    """Stands in for GP: deterministic 'prediction' so that the gather can be checked on CPU."""

    def _predict(self, x, want_mean, want_var, standardised):
        x = np.asarray(x)
        return x.sum(axis=1), (x ** 2).sum(axis=1) if want_var else None

    def neg_mll_and_grad_batched(self, lp):
        lp = np.atleast_2d(lp)
        return (lp ** 2).sum(axis=1), 2 * lp

    def fantasy_acquisition(self, mc, cand=None, std=False):
        """mean over the MC points of a per-(candidate, MC point) quantity: what WIPV / WIPStd are structurally."""
        mc = np.asarray(mc)
        cand = mc if cand is None else np.atleast_2d(np.asarray(cand))
        s = 1.0 + np.cos(cand.sum(1))[:, None] * np.sin(mc.sum(1))[None, :] ** 2
        return (np.sqrt(s) if std else s).mean(axis=1)


def _worker(rank, ws, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        rng = np.random.default_rng(0)
        xq = rng.uniform(0, 1, (101, 3))
        mean, var = bd.predict_sharded(_FakeGP(), xq)
        assert np.allclose(mean, xq.sum(1)) and np.allclose(var, (xq ** 2).sum(1))
        # restart-sharded fit: each rank's "fit" returns the best of its chunk; all ranks agree on the global best
        x0 = rng.uniform(-2, 2, (5, 4))
        score = lambda chunk: {"mll": float(-np.min((chunk ** 2).sum(1))), "params": chunk[np.argmin((chunk ** 2).sum(1))]}
        res = bd.fit_sharded(score, x0)
        best = np.argmin((x0 ** 2).sum(1))
        assert np.allclose(res["params"], x0[best]) and np.isclose(res["mll"], -(x0[best] ** 2).sum())
        # a rank whose fit fails (non-finite) must not win
        res2 = bd.fit_sharded(lambda c: {"mll": float("nan"), "params": c[0]} if rank == 0 else score(c), x0)
        lo, hi = bd.shard_bounds(5, 1, ws)
        assert np.isclose(res2["mll"], -np.min((x0[lo:hi] ** 2).sum(1)))
        v, g = bd.mll_grad_sharded(_FakeGP(), x0)
        assert np.allclose(v, (x0 ** 2).sum(1)) and np.allclose(g, 2 * x0)
        vals = bd.acquisition_sharded(lambda c: c.sum(1), xq[:7])
        assert np.allclose(vals, xq[:7].sum(1))
        # WIPV with the MC columns sharded: size-weighted partial means, gathered and added in rank order
        mc = rng.uniform(0, 1, (37, 3))  # odd count: shards of 19 and 18
        cand = rng.uniform(0, 1, (8, 3))
        for std in (False, True):
            w = bd.wipv_sharded(_FakeGP(), mc, cand, std=std)
            assert w.shape == (8,) and np.allclose(w, _FakeGP().fantasy_acquisition(mc, cand, std), rtol=1e-13, atol=0)
            ws_self = bd.wipv_sharded(_FakeGP(), mc, None, std=std)
            assert ws_self.shape == (37,) and np.allclose(ws_self, _FakeGP().fantasy_acquisition(mc, None, std), rtol=1e-13)
        one = bd.wipv_sharded(_FakeGP(), mc[:1], cand)  # fewer MC points than ranks: an empty shard contributes zero
        assert np.allclose(one, _FakeGP().fantasy_acquisition(mc[:1], cand), rtol=1e-13)
        t = bd.allgather_rows(torch.arange(rank * 3, rank * 3 + (3 if rank == 0 else 2), dtype=torch.float64), 5)
        assert torch.equal(t, torch.arange(5, dtype=torch.float64))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        tdist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_single_process_paths():
    xq = np.random.default_rng(1).uniform(0, 1, (10, 2))
    m, v = bd.predict_sharded(_FakeGP(), xq)
    assert np.allclose(m, xq.sum(1))
    assert bd.world() == (0, 1)
    mc = np.random.default_rng(2).uniform(0, 1, (9, 2))
    assert np.array_equal(bd.wipv_sharded(_FakeGP(), mc, mc[:3]), _FakeGP().fantasy_acquisition(mc, mc[:3]))
