"""CPU, world_size 2 over gloo: the host logic of the row-sharded round (partition, int64 key encoding, ONE MIN
all-reduce per round, collective convergence test) with the oracle's sweep standing in for the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import load_golden
from scip_b200 import sharded, synth


class OracleEngine:
    """local row block swept by the CPU oracle (tests only); mirrors CudaEngine's interface"""

    def __init__(self, prob, rows, **numerics):
        self.prob, self.rows, self.numerics = prob, rows, numerics
        self.vartype = np.asarray(prob["vartype"])

    def set_bounds(self, lb, ub):
        self.lb = np.array(lb, dtype=np.float64) + 0.0
        self.ub = np.array(ub, dtype=np.float64) + 0.0

    def get_bounds(self):
        return self.lb.copy(), self.ub.copy()

    def round_begin(self):
        pass

    def round_sweep(self):
        cutoff, nlb, nub = oracle.sweep(self.prob, self.lb, self.ub, self.rows[0], self.rows[1], **self.numerics)
        keys = sharded.encode_keys(nlb, nub)
        keys[-2] = -1 if cutoff else 0
        self.keys = torch.from_numpy(keys)
        return self.keys

    # the packed exchange of the product's multi-GPU round (sharded.ReplicatedPropagator)
    def get_keys(self):
        return sharded.encode_keys(self.lb, self.ub)

    def set_keys(self, keys):
        self.keys = torch.from_numpy(np.ascontiguousarray(keys))

    def round_apply(self):
        nlb, nub, cutoff = sharded.decode_keys(self.keys.numpy())
        cross = nlb > nub
        feastol = self.numerics.get("feastol", 1e-6)
        rel = (nlb - nub) / np.maximum(1.0, np.maximum(np.abs(nlb), np.abs(nub)))
        if np.any(cross & (rel > feastol)):
            cutoff = True
        nlb = np.where(cross, nub, nlb)
        nchg = int((nlb != self.lb).sum() + (nub != self.ub).sum())
        self.lb, self.ub = nlb, nub
        return nchg, int(cutoff)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, probs, out, packed=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for name, (prob, numerics) in probs.items():
            cuts = sharded.partition_rows(prob["rowptr"], world)
            eng = OracleEngine(prob, (int(cuts[rank]), int(cuts[rank + 1])), **numerics)
            eng.set_bounds(prob["lb"], prob["ub"])
            driver = sharded.ReplicatedPropagator if packed else sharded.ShardedPropagator
            res = driver(eng).propagate(maxrounds=500)
            lb, ub = eng.get_bounds()
            out[(name, rank)] = (res, lb, ub)
    finally:
        dist.destroy_process_group()


def test_key_encoding_is_order_preserving_and_an_involution():
    x = np.array([-1e20, -3.5, -0.0, 0.0, 1e-300, 2.0, 1e20])
    keys = sharded.encode_keys(x, x)
    assert np.all(np.diff(keys[1:-2:2]) >= 0)          # ub keys ascend with the value
    assert np.all(np.diff(keys[0:-2:2]) <= 0)          # lb keys descend: MIN tightens both sides
    lb, ub, cutoff = sharded.decode_keys(keys)
    assert np.array_equal(lb, x + 0.0) and np.array_equal(ub, x + 0.0) and not cutoff
    assert keys[2 * 2] == keys[2 * 3]                  # -0.0 and 0.0 share a key


def test_partition_rows_balances_nonzeros():
    prob = synth.mixed_knapsack(500, 5000, 60_000, seed=3, dense_range=(400, 900))
    for parts in (1, 2, 3, 8):
        cuts = sharded.partition_rows(prob["rowptr"], parts)
        assert cuts[0] == 0 and cuts[-1] == len(prob["lhs"]) and np.all(np.diff(cuts) >= 0)
        nnz = np.diff(prob["rowptr"][cuts])
        assert nnz.max() - nnz.min() <= 2 * np.diff(prob["rowptr"]).max()


def test_packed_exchange_equals_the_dense_min():
    # pack_changes / merge_changes (the host mirror of peer_push_kernel / peer_merge_kernel): merging the packets of two
    # ranks into either rank's keys == the elementwise MIN of the two key vectors
    rng = np.random.default_rng(1)
    lb, ub = rng.normal(size=50), rng.normal(size=50) + 5.0
    base = sharded.encode_keys(lb, ub)
    a, b = base.copy(), base.copy()
    ia, ib = rng.integers(0, 100, size=30), rng.integers(0, 100, size=30)
    a[ia] -= rng.integers(1, 1000, size=30)
    b[ib] -= rng.integers(1, 1000, size=30)
    want = np.minimum(a, b)
    for mine, other in ((a, b), (b, a)):
        got = mine.copy()
        cols, pairs = sharded.pack_changes(base, other)
        assert len(cols) <= 30
        sharded.merge_changes(got, cols, pairs)
        assert np.array_equal(got, want)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("packed", [False, True])
def test_two_ranks_reach_the_single_process_fixpoint(packed):
    probs = {
        "bell5": (load_golden("bell5", "1e-9")[0], dict(boundstreps=1e-9)),
        "egout": (load_golden("egout", "1e-9")[0], dict(boundstreps=1e-9)),
        "setcover": (synth.setcover(3000, 3000, 30_000, seed=4), {}),
        "setcover_infeasible": (synth.setcover(3000, 3000, 30_000, seed=5, infeasible=True), {}),
        "mixedknap": (synth.mixed_knapsack(300, 3000, 30_000, seed=6, dense_range=(300, 900)), dict(boundstreps=1e-9)),
    }
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, probs, out, packed), nprocs=world, join=True)
        out = dict(out)
    for name, (prob, numerics) in probs.items():
        want = oracle.propagate(prob, **numerics)
        r0, lb0, ub0 = out[(name, 0)]
        r1, lb1, ub1 = out[(name, 1)]
        assert r0 == r1                                         # collective convergence: same rounds, same verdict
        assert r0["status"] == want["status"], name
        if want["status"] != oracle.STATUS_CUTOFF:
            assert r0["nrounds"] == want["nrounds"] and r0["nchanges"] == want["nchanges"], name
            for lb, ub in ((lb0, ub0), (lb1, ub1)):
                assert np.array_equal(lb, want["lb"]) and np.array_equal(ub, want["ub"]), name
