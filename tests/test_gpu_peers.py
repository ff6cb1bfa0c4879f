"""GPU, 2 devices: rows sharded over two ranks, candidates committed into both ranks' key vectors through peer memory
(CUDA IPC + system-scope atomics), device-side barriers, the fixpoint loop stays in the CUDA graph on each GPU."""
import os
import socket

import numpy as np
import pytest

import oracle
from conftest import assert_bounds_match, load_golden
from scip_b200 import synth

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        from scip_b200 import propagator
        return propagator.load_library().gpulin_device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, probs, out):
    import torch.distributed as dist
    from scip_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for name, (prob, numerics) in probs.items():
            pp = sharded.PeerPropagator(prob, rank, world, device=rank, **numerics)
            for rep in range(2):                       # twice: the barrier epochs must survive a second call
                pp.set_bounds(prob["lb"], prob["ub"])
                res = pp.propagate(0)
            lb, ub = pp.get_bounds()
            out[(name, rank)] = (res, lb, ub)
            dist.barrier()
            pp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.timeout(600)
def test_two_gpus_peer_exchange_reaches_the_fixpoint():
    import torch.multiprocessing as mp
    probs = {
        "egout": (load_golden("egout", "1e-9")[0], dict(boundstreps=1e-9)),
        "setcover": (synth.setcover(100_000, 100_000, 1_000_000, seed=2), {}),
        "setcover_infeasible": (synth.setcover(50_000, 50_000, 500_000, seed=5, infeasible=True), {}),
        "mixedknap": (synth.mixed_knapsack(4000, 40_000, 1_000_000, seed=21, dense_range=(1500, 6000), eq_frac=0.2),
                      dict(boundstreps=1e-9)),
    }
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, probs, out), nprocs=world, join=True)
        out = dict(out)
    for name, (prob, numerics) in probs.items():
        want = oracle.propagate(prob, **numerics)
        for rank in range(world):
            res, lb, ub = out[(name, rank)]
            assert res["status"] == want["status"], (name, rank)
            if want["status"] != oracle.STATUS_CUTOFF:
                assert res["nrounds"] == want["nrounds"] and res["nchanges"] == want["nchanges"], (name, rank)
                assert_bounds_match(lb, ub, want["lb"], want["ub"], prob["vartype"], what=f"{name} rank {rank}")
