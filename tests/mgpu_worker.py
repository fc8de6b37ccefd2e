"""Worker for the multi-rank GPU test (launched by torchrun, one process per GPU):
sharded k-centers / k-hybrid must equal the single-GPU run bit for bit (contiguous shards keep
the serial lowest-index tie rule)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from enspara_b200 import mpi, synth
    from enspara_b200.cluster import hybrid, kcenters
    mpi.init_from_env("nccl")
    rank, size = mpi.rank(), mpi.size()
    ok = True
    for (n, A, kw) in ((3001, 50, dict(n_clusters=25)), (2000, 264, dict(dist_cutoff=1.9)),
                       (5, 10, dict(n_clusters=4)), (1, 10, dict(n_clusters=1)),
                       (3, 10, dict(n_clusters=3))):      # the last two leave shards EMPTY
        X = synth.trajectory(n, A, seed=7)
        bounds = np.linspace(0, n, size + 1).astype(int)
        mine = X[bounds[rank]:bounds[rank + 1]]
        serial = kcenters.kcenters(X, "rmsd", mpi_mode=False, **kw)
        shard = kcenters.kcenters(mine, "rmsd", mpi_mode=True, **kw)
        glob = [int(bounds[r] + l) for r, l in shard.center_indices]
        ok &= glob == [int(c) for c in serial.center_indices]
        ok &= np.array_equal(shard.assignments, serial.assignments[bounds[rank]:bounds[rank + 1]])
        ok &= np.array_equal(shard.distances, serial.distances[bounds[rank]:bounds[rank + 1]])
        ok &= len(shard.centers) == len(serial.centers)
        tri = kcenters.kcenters(mine, "rmsd", mpi_mode=True, use_triangle_inequality=True, **kw)
        ok &= [tuple(c) for c in tri.center_indices] == [tuple(c) for c in shard.center_indices]
        ok &= np.array_equal(tri.assignments, shard.assignments)
        ok &= np.array_equal(tri.distances, shard.distances)
        for a, b in zip(shard.centers, serial.centers):
            ok &= np.array_equal(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1))
    # features, euclidean, exact
    F = synth.features(10007, 16, seed=3)
    bounds = np.linspace(0, len(F), size + 1).astype(int)
    serial = kcenters.kcenters(F, "euclidean", n_clusters=30)
    shard = kcenters.kcenters(F[bounds[rank]:bounds[rank + 1]], "euclidean", n_clusters=30,
                              mpi_mode=True)
    ok &= [int(bounds[r] + l) for r, l in shard.center_indices] == \
        [int(c) for c in serial.center_indices]
    ok &= np.array_equal(shard.distances, serial.distances[bounds[rank]:bounds[rank + 1]])
    # k-hybrid: sharded PAM == serial PAM (same seeded RandomState on every rank)
    X = synth.trajectory(1200, 30, seed=11)
    bounds = np.linspace(0, len(X), size + 1).astype(int)
    serial = hybrid.hybrid(X, "rmsd", n_clusters=8, n_iters=2, random_state=0)
    shard = hybrid.hybrid(X[bounds[rank]:bounds[rank + 1]], "rmsd", n_clusters=8, n_iters=2,
                          random_state=0, mpi_mode=True)
    ok &= [int(bounds[r] + l) for r, l in shard.center_indices] == \
        [int(c) for c in serial.center_indices]
    ok &= np.array_equal(shard.assignments, serial.assignments[bounds[rank]:bounds[rank + 1]])
    ok &= np.array_equal(shard.distances, serial.distances[bounds[rank]:bounds[rank + 1]])
    # ---- config-4-shaped shards (BASELINE configs[3]): >= 76k frames x 500 atoms per rank, so
    # the kernel is k_kcenters_step_rmsd_tma and the exchange the fused peer-memory one -- the
    # combination bench.py times at N > 1.  Sharded == single GPU on centres, assignments AND
    # distances (the reference's MPI bar, test_cluster.py:270-275, 309-314), for a run bounded
    # by n_clusters and for one terminated by dist_cutoff (kcenters.py:217).
    from enspara_b200 import _lib
    per = int(os.environ.get("EB_MGPU_C4_FRAMES", "100000"))
    A = 500
    full = synth.device_trajectory(per * size, A, seed=0, first_frame=0)
    mine = synth.device_trajectory(per, A, seed=0, first_frame=rank * per)
    lo, hi = rank * per, (rank + 1) * per
    p2p_wanted = os.environ.get("ENSPARA_B200_P2P", "1") != "0"
    report = {}

    def same(shard, serial, tag):
        g = [int(r * per + l) for r, l in shard.center_indices]
        c = g == [int(v) for v in serial.center_indices]
        a = np.array_equal(shard.assignments, serial.assignments[lo:hi])
        d = np.array_equal(shard.distances, serial.distances[lo:hi])
        report[tag] = dict(k=len(g), centres=bool(c), assignments=bool(a), distances=bool(d))
        return c and a and d

    serial = kcenters.kcenters(full, "rmsd", n_clusters=24, mpi_mode=False)
    shard, eng = kcenters.kcenters(mine, "rmsd", n_clusters=24, mpi_mode=True,
                                   _return_engine=True)
    ok &= same(shard, serial, "c4_n_clusters")
    ok &= bool(_lib.load().eb_kcenters_step_rmsd_uses_tma(per, A))
    ok &= (eng.p2p == p2p_wanted)
    report["c4_kernel_is_tma"] = bool(_lib.load().eb_kcenters_step_rmsd_uses_tma(per, A))
    report["c4_exchange_is_fused_p2p"] = bool(eng.p2p)
    report["c4_wait_us_per_step"] = 1e-3 * getattr(eng, "wait_ns", 0) / 24
    # cutoff-terminated: a radius between the 17th and the 24th centre's max-min-distance, so
    # the stop rule (not n_clusters) ends the run, after several polls of the device state
    cut = float(serial.distances.max()) * 1.03
    serial_c = kcenters.kcenters(full, "rmsd", dist_cutoff=cut, mpi_mode=False)
    shard_c = kcenters.kcenters(mine, "rmsd", dist_cutoff=cut, mpi_mode=True)
    ok &= same(shard_c, serial_c, "c4_dist_cutoff")
    ok &= 1 < len(serial_c.center_indices) <= 24
    ok &= float(serial_c.distances.max()) <= cut
    # both bounds at once, and the exchange buffer re-used by a third run in a row
    serial_b = kcenters.kcenters(full, "rmsd", n_clusters=7, dist_cutoff=cut, mpi_mode=False)
    shard_b = kcenters.kcenters(mine, "rmsd", n_clusters=7, dist_cutoff=cut, mpi_mode=True)
    ok &= same(shard_b, serial_b, "c4_both")
    del full, mine, serial, shard, eng, serial_c, shard_c, serial_b, shard_b
    torch.cuda.empty_cache()

    # ---- k-hybrid WITHOUT a seed (random_state=None): every rank's global RandomState is
    # seeded differently; the sweep must still consume ONE stream (rank 0's), or ranks pick
    # different owners / broadcast roots and hang or diverge (the reference: rank 0 draws,
    # mpi/ops.py:247-253)
    np.random.seed(1000 + rank)
    X = synth.trajectory(1500, 30, seed=13)
    bounds = np.linspace(0, len(X), size + 1).astype(int)
    un = hybrid.hybrid(X[bounds[rank]:bounds[rank + 1]], "rmsd", n_clusters=9, n_iters=2,
                       random_state=None, mpi_mode=True)
    mine_ids = torch.tensor([int(bounds[r] + l) for r, l in un.center_indices],
                            dtype=torch.int64, device="cuda")
    every = [torch.empty_like(mine_ids) for _ in range(size)]
    dist.all_gather(every, mine_ids)
    ok &= all(bool(torch.equal(e, every[0])) for e in every)
    # ... and the state is self-consistent: every frame sits with its nearest medoid
    from enspara_b200.cluster import util as cutil
    med = [X[int(g)] for g in every[0].cpu().numpy()]
    a2, d2 = cutil.assign_to_nearest_center(X[bounds[rank]:bounds[rank + 1]], med, "rmsd")
    ok &= np.array_equal(a2, un.assignments)
    ok &= np.allclose(d2, un.distances, rtol=1e-6, atol=1e-7)
    report["khybrid_unseeded_ranks_agree"] = bool(
        all(bool(torch.equal(e, every[0])) for e in every))

    # the `cluster` app under torchrun: file i on rank i % size, results re-assembled on every
    # rank, rank 0 writes -- must equal the serial estimator (test_apps_cluster_mpi.py:100-139)
    import tempfile
    from enspara_b200 import ra
    from enspara_b200.apps import cluster as app
    from enspara_b200.cluster import KCenters
    tmp = os.environ.get("EB_MGPU_TMP") or tempfile.gettempdir()
    G = synth.features(4 * 300, 6, seed=9)
    files = [os.path.join(tmp, "eb_mgpu_f%d.npy" % i) for i in range(2 * size)]
    cuts = np.linspace(0, len(G), 2 * size + 1).astype(int)
    if rank == 0:
        for i, f in enumerate(files):
            np.save(f, G[cuts[i]:cuts[i + 1]])
    dist.barrier()
    outs = {k: os.path.join(tmp, "eb_mgpu_%s" % k) for k in
            ("dist.h5", "assig.h5", "ctrs.npy", "inds.npy")}
    rc = app.main(["cluster", "--features"] + files + [
        "--algorithm", "kcenters", "--cluster-distance", "euclidean", "--cluster-radius", "0.7",
        "--distances", outs["dist.h5"], "--assignments", outs["assig.h5"],
        "--center-features", outs["ctrs.npy"], "--center-indices", outs["inds.npy"]])
    ok &= rc == 0
    dist.barrier()
    if rank == 0:
        serial = KCenters("euclidean", cluster_radius=0.7, mpi_mode=False).fit(G)
        d = ra.load(outs["dist.h5"])
        a = ra.load(outs["assig.h5"])
        dflat = d.flatten() if hasattr(d, "flatten") and not isinstance(d, np.ndarray) \
            else np.asarray(d).reshape(-1)
        aflat = a.flatten() if hasattr(a, "flatten") and not isinstance(a, np.ndarray) \
            else np.asarray(a).reshape(-1)
        ok &= np.array_equal(dflat, serial.distances_)
        ok &= len(np.unique(aflat)) == len(serial.center_indices_)
        ctrs = np.load(outs["ctrs.npy"])
        ok &= ctrs.shape == (len(serial.center_indices_), 6)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        import json
        report["ranks"] = size
        report["ok"] = bool(int(flag.item()) == 1)
        print("MGPU_REPORT " + json.dumps(report), flush=True)
        print("MGPU_OK" if int(flag.item()) == 1 else "MGPU_FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
