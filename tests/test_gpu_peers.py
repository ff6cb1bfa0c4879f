"""GPU, 2+ devices of one node: every rank holds the whole problem, dense rounds are shared (one packed exchange per
dense round through peer memory inside the round's kernels), small rounds run redundantly; the fixpoint loop stays in the
CUDA graph on each GPU.  Covered: one process per GPU (CUDA IPC), one process driving all GPUs (gpulin_group_connect, what
the SCIP plugin does), and the host-driven NCCL variant (one int64 MIN all-reduce per round) that is kept for comparison."""
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


def _problems():
    return {
        "egout": (load_golden("egout", "1e-9")[0], dict(boundstreps=1e-9)),
        "setcover": (synth.setcover(100_000, 100_000, 1_000_000, seed=2), {}),
        "setcover_big": (synth.setcover(300_000, 300_000, 3_000_000, seed=7), {}),      # bit-table sweep, several dense rounds
        "setcover_infeasible": (synth.setcover(50_000, 50_000, 500_000, seed=5, infeasible=True), {}),
        "mixedknap": (synth.mixed_knapsack(4000, 40_000, 1_000_000, seed=21, dense_range=(1500, 6000), eq_frac=0.2),
                      dict(boundstreps=1e-9)),
        "unitnet": (synth.unit_network(200_000, 150_000, 1_000_000, seed=4), {}),
    }


def _check(out, probs, world, what):
    for name, (prob, numerics) in probs.items():
        want = oracle.propagate(prob, **numerics)
        for rank in range(world):
            res, lb, ub = out[(name, rank)]
            assert res["status"] == want["status"], (what, name, rank)
            if want["status"] != oracle.STATUS_CUTOFF:
                assert res["nrounds"] == want["nrounds"] and res["nchanges"] == want["nchanges"], (what, name, rank)
                assert_bounds_match(lb, ub, want["lb"], want["ub"], prob["vartype"], what=f"{what} {name} rank {rank}")


def _worker(rank, world, port, probs, out, mode):
    import torch
    import torch.distributed as dist
    from scip_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if mode == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for name, (prob, numerics) in probs.items():
            if mode == "nccl":
                cuts = sharded.partition_rows(prob["rowptr"], world)
                eng = sharded.CudaEngine(prob, (int(cuts[rank]), int(cuts[rank + 1])), rank, **numerics)
                pp = sharded.ShardedPropagator(eng)
                pp.set_bounds, pp.get_bounds, pp.close = eng.set_bounds, eng.get_bounds, eng.close
            else:
                pp = sharded.PeerPropagator(prob, rank, world, device=rank, **numerics)
            for rep in range(2):                       # twice: the exchange numbers must survive a second call
                pp.set_bounds(prob["lb"], prob["ub"])
                res = pp.propagate(0)
            lb, ub = pp.get_bounds()
            out[(name, rank)] = (res, lb, ub)
            dist.barrier()
            pp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.timeout(900)
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_one_process_per_gpu_reaches_the_fixpoint(mode):
    import torch.multiprocessing as mp
    probs = _problems()
    world = min(_ngpus(), 4)
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, probs, out, mode), nprocs=world, join=True)
        out = dict(out)
    _check(out, probs, world, mode)


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.timeout(900)
def test_one_process_drives_all_gpus():
    """gpulin_group_connect: the in-process connection of the SCIP plugin; also an incremental call (few updated bounds
    run in one block on every device, redundantly) and a round limit with resume"""
    from scip_b200 import sharded
    world = min(_ngpus(), 4)
    probs = _problems()
    out = {}
    for name, (prob, numerics) in probs.items():
        gp = sharded.GroupPropagator(prob, list(range(world)), **numerics)
        try:
            for rep in range(2):
                gp.set_bounds(prob["lb"], prob["ub"])
                res = gp.propagate(0)
            for rank in range(world):
                lb, ub = gp.get_bounds(rank)
                out[(name, rank)] = (res, lb, ub)
            if name == "setcover":
                lb, ub = gp.get_bounds(0)
                free = np.flatnonzero(lb < ub)[:3]
                gp.update_bounds(free, np.ones(3), np.ones(3))
                r2 = gp.propagate(0)
                lb2, ub2 = lb.copy(), ub.copy()
                lb2[free] = 1.0
                want = oracle.propagate(prob, lb=lb2, ub=ub2)
                assert r2["status"] == want["status"]
                if want["status"] != oracle.STATUS_CUTOFF:
                    assert r2["nrounds"] == want["nrounds"] and r2["nchanges"] == want["nchanges"]
                    for rank in range(world):
                        glb, gub = gp.get_bounds(rank)
                        assert_bounds_match(glb, gub, want["lb"], want["ub"], prob["vartype"], what=f"group update rank {rank}")
                gp.set_bounds(prob["lb"], prob["ub"])
                r3 = gp.propagate(2)
                assert r3["status"] == 2 and r3["nrounds"] == 2
                r4 = gp.propagate(0)
                full = oracle.propagate(prob)
                assert r3["nrounds"] + r4["nrounds"] == full["nrounds"]
                glb, gub = gp.get_bounds(world - 1)
                assert_bounds_match(glb, gub, full["lb"], full["ub"], prob["vartype"], what="group resume")
        finally:
            gp.close()
    _check(out, probs, world, "group")
