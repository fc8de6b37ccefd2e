"""BASELINE config 5 on N GPUs (torchrun): assign_to_nearest_center with 10k centres against a
10M-frame x 500-atom trajectory sharded over the ranks (centres replicated, no exchange inside
the pass), followed by k-medoids (PAM) proposals on the sharded state.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        scripts/bench_c5_multi.py [--frames-per-gpu 1250000] [--centres 10000] [--proposals 24]

Rank 0 prints one JSON object.  Times are CUDA-synchronised wall clock, max over ranks.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--frames-per-gpu", type=int, default=1_250_000)
    p.add_argument("--atoms", type=int, default=500)
    p.add_argument("--centres", type=int, default=10_000)
    p.add_argument("--proposals", type=int, default=24)
    args = p.parse_args()
    from enspara_b200 import mpi, synth
    from enspara_b200.cluster import _ops, util
    from enspara_b200.cluster._pam import PamEngine
    from enspara_b200.cluster.kcenters import _SingleComm
    from enspara_b200.device import DeviceTrajectory
    mpi.init_from_env("nccl")
    rank, size = mpi.rank(), mpi.size()
    comm = mpi.comm if size > 1 else _SingleComm()
    n, A, k = args.frames_per_gpu, args.atoms, args.centres
    X = synth.device_trajectory(n, A, seed=0, first_frame=rank * n)

    # centres: every rank contributes k/size evenly spaced frames of its shard, all-gathered
    per = k // size
    k = per * size
    # an odd stride visits all 64 base conformers of the synthetic set (an even one would put
    # every centre on a handful of conformers and leave most frames without a near centre)
    stride = max(1, (n // per - 1) | 1)
    loc = torch.arange(per, device="cuda", dtype=torch.int64) * stride
    mine = X.gather(loc)
    cen = DeviceTrajectory.empty(k, A)
    if size > 1:
        dist.all_gather_into_tensor(cen.xyz.view(-1), mine.xyz.view(-1))
        dist.all_gather_into_tensor(cen.traces, mine.traces)
    else:
        cen.xyz.copy_(mine.xyz)
        cen.traces.copy_(mine.traces)
    medoid_global = [int(r * n + i) for r in range(size) for i in loc.cpu().tolist()]

    def barrier():
        if size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.cpu()[0])

    ws = {}
    _ops.assign_device_tc(util.RMSD, synth.device_trajectory(8192, A, seed=3), cen, workspace={})
    barrier()
    t = time.perf_counter()
    stats = {}
    d, a = _ops.assign_device_tc(util.RMSD, X, cen, workspace=ws, stats=stats)
    torch.cuda.synchronize()
    dt_assign = tmax(time.perf_counter() - t)
    ok = bool((a[loc].cpu() == (torch.arange(per, dtype=torch.int32) + rank * per)).all())
    del ws

    pam = PamEngine(X, util.RMSD, comm, d, a, medoid_global)
    pam.sweep(random_state=0, max_proposals=4)
    barrier()
    t = time.perf_counter()
    acc = pam.sweep(random_state=1, max_proposals=args.proposals)
    torch.cuda.synchronize()
    dt_pam = tmax(time.perf_counter() - t)
    if rank == 0:
        n_total = n * size
        print(json.dumps({
            "config": "C5: assign_to_nearest_center %d centres x %d frames x %d atoms over %d "
                      "GPU(s), then PAM proposals on the sharded state" % (k, n_total, A, size),
            "n_gpus": size, "assign_seconds": dt_assign,
            "assign_evals_per_s": n_total * k / dt_assign,
            "assign_algorithmic_TFLOPs": n_total * k / dt_assign * 18 * A / 1e12,
            "centres_assigned_to_themselves": ok,
            "survivors_per_frame_rank0": stats.get("survivors_mean"),
            "overflow_frames_rank0": stats.get("overflow_frames"),
            "pam_proposals": args.proposals, "pam_accepted": int(acc),
            "pam_ms_per_proposal": 1e3 * dt_pam / args.proposals,
            "pam_full_pass_evals_per_s": n_total * args.proposals / dt_pam}), flush=True)
    if size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
