#!/usr/bin/env python
"""The NCCL exchange against the peer-memory exchange on the same GPUs (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/probe_nccl.py [c3|c4] [steps]

NCCL variant (sharded.ShardedPropagator): rows partitioned over the ranks, per round local sweeps + exact rules, ONE
all_reduce(MIN) over the int64 key vector (2 ncols + 2 keys), dense apply -- a host loop with one read-back per round.
Peer variant (sharded.PeerPropagator, the product's default): whole matrix on every rank, dense rounds shared, one packed
exchange per dense round inside the round's kernels, the loop on the devices.  Prints one JSON line from rank 0 with the
wall time per fixpoint of both (max over ranks), after checking that they end with the same bounds."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_b200 import sharded, synth  # noqa: E402


def timed(pp, prob, steps, device):
    out = []
    for _ in range(steps + 2):
        pp.set_bounds(prob["lb"], prob["ub"])
        dist.barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        res = pp.propagate(0)
        torch.cuda.synchronize(device)
        out.append(time.perf_counter() - t0)
    t = torch.tensor([float(np.median(out[2:]))], device=f"cuda:{device}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return res, float(t.item())


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{device}"))
    prob = synth.setcover() if which == "c3" else synth.mixed_knapsack()

    cuts = sharded.partition_rows(prob["rowptr"], world)
    eng = sharded.CudaEngine(prob, (int(cuts[rank]), int(cuts[rank + 1])), device)
    nccl = sharded.ShardedPropagator(eng)
    nccl.set_bounds = eng.set_bounds
    res_n, t_n = timed(nccl, prob, steps, device)
    lb_n, ub_n = eng.get_bounds()
    eng.close()

    peer = sharded.PeerPropagator(prob, rank, world, device=device)
    res_p, t_p = timed(peer, prob, steps, device)
    lb_p, ub_p = peer.get_bounds()
    peer.close()

    same = bool(np.array_equal(lb_n, lb_p) and np.array_equal(ub_n, ub_p) and res_n["nrounds"] == res_p["nrounds"]
                and res_n["nchanges"] == res_p["nchanges"])
    if rank == 0:
        print(json.dumps(dict(workload=which, n_gpus=world, steps=steps, rounds=res_p["nrounds"], changes=res_p["nchanges"],
                              nccl_ms_per_fixpoint=t_n * 1e3, peer_ms_per_fixpoint=t_p * 1e3, same_result=same,
                              note="wall time per gpulin fixpoint incl. the host's launch and read-back, median over the "
                                   "steps, max over the ranks; nccl = one all_reduce(MIN) over 2 ncols + 2 int64 keys per "
                                   "round from a host loop, peer = the product's packed exchange inside the dense rounds")))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
