"""Measure every BASELINE.json config that fits one B200 (the bench.py line covers config 4's
per-GPU shard; this script covers configs 1, 2, 3 and the per-GPU shard of config 5) through
the PUBLIC estimator / function API, and write one JSON object per config.

    python scripts/bench_configs.py [--out gpurun_out/configs.json] [--only c1,c2,c3,c5,reassign]

Timing: wall clock around the public call with a device synchronise on both sides (these are
whole-algorithm numbers with host control flow inside, not kernel times); inputs are generated
in HBM (`data: synthetic, resident`) unless the entry says `host`.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

PEAK = 6550.4
TENSOR_PEAK = 1575.1       # TFLOP/s, dense bf16/fp16 burst (MEASURED_PEAKS.json)
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                           "MEASURED_PEAKS.json")) as fh:
        _pk = json.load(fh)
        PEAK = float(_pk["hbm_gbs"])
        TENSOR_PEAK = float(_pk.get("bf16_tflops", TENSOR_PEAK))
except Exception:
    pass


def tensor_peaks():
    """cuBLAS dense GEMM peaks on THIS box (8192^3, best of 10, CUDA events): fp16 with FP32
    accumulation (what the K3t screen executes) and TF32 (what SURVEY 8d asked to be measured
    for the 3xTF32 formulation of round 1).  cuBLAS is used for the PEAK only."""
    out = {}
    n = 8192
    for name, dt, tf32 in (("fp16", torch.float16, False), ("tf32", torch.float32, True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a = torch.randn(n, n, device="cuda", dtype=dt)
        b = torch.randn(n, n, device="cuda", dtype=dt)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name + "_tflops"] = 2 * n ** 3 / (best * 1e-3) / 1e12
        del a, b
    torch.backends.cuda.matmul.allow_tf32 = False
    out["how"] = "torch.matmul 8192^3 (cuBLAS), best of 10, CUDA events"
    return out


def sync():
    torch.cuda.synchronize()


def timed(fn):
    sync()
    t = time.perf_counter()
    out = fn()
    sync()
    return out, time.perf_counter() - t


def c1():
    """KCenters rmsd n_clusters=100, 20k frames x 264 atoms (config 1; L2-resident: 63 MB)."""
    from enspara_b200 import synth
    from enspara_b200.cluster import KCenters
    X = synth.trajectory(20_000, 264, seed=0)
    KCenters("rmsd", n_clusters=5).fit(X[:2000])
    est, dt = timed(lambda: KCenters("rmsd", n_clusters=100).fit(X))
    dev = synth.device_trajectory(20_000, 264, seed=0)
    est2, dt2 = timed(lambda: KCenters("rmsd", n_clusters=100).fit(dev))
    same = [int(c) for c in est.result_.center_indices] == \
        [int(c) for c in est2.result_.center_indices]
    # the device loop alone (CUDA events around the queued iterations)
    from enspara_b200.cluster._engine import KCentersEngine
    from enspara_b200.cluster.kcenters import _SingleComm
    loop_us = None
    for _ in range(2):
        eng = KCentersEngine(dev, "rmsd", _SingleComm())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record()
        eng.run(100, 0.0)
        e1.record()
        sync()
        loop_us = 1e3 * e0.elapsed_time(e1) / 100
    return {"config": "C1 KCenters rmsd k=100, 20k x 264 (host ndarray in, results out)",
            "device_loop_us_per_iteration": loop_us,
            "seconds_host_in": dt, "evals_per_s_host_in": 20_000 * 100 / dt,
            "seconds_resident": dt2, "evals_per_s_resident": 20_000 * 100 / dt2,
            "us_per_iteration_resident": 1e6 * dt2 / 100,
            "note": "63 MB of coordinates: L2-resident, launch-latency regime; not an HBM "
                    "fraction", "host_equals_resident": same}


def c2():
    """KCenters euclidean on 1M x 64 f32 (config 2), k=1000 (SURVEY 8d assumption)."""
    from enspara_b200 import synth
    from enspara_b200.cluster import KCenters
    X = synth.device_features(1_000_000, 64, seed=0)
    KCenters("euclidean", n_clusters=5).fit(X)
    k = 1000
    est, dt = timed(lambda: KCenters("euclidean", n_clusters=k).fit(X))
    evs = 1_000_000 * k / dt
    # the device loop alone (CUDA events around the queued iterations; no result read-back)
    from enspara_b200.cluster._engine import KCentersEngine
    from enspara_b200.cluster.kcenters import _SingleComm
    eng = KCentersEngine(X, "euclidean", _SingleComm())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    eng.run(k, 0.0)
    e1.record()
    sync()
    loop_us = 1e3 * e0.elapsed_time(e1) / k
    st = eng.read_state()
    return {"config": "C2 KCenters euclidean k=1000, 1M x 64 f32 (resident)",
            "seconds": dt, "us_per_iteration": 1e6 * dt / k, "evals_per_s": evs,
            "algorithmic_GBps": evs * 264 / 1e9, "frac_of_hbm_peak": evs * 264 / 1e9 / PEAK,
            "device_loop_us_per_iteration": loop_us,
            "device_loop_frac_of_hbm_peak": 1e6 / loop_us * 1_000_000 * 264 / 1e9 / PEAK,
            "arg_max_barrier_us_per_iteration_block0": 1e-3 * st.wait_ns / max(1, k - 1),
            "note": "`seconds` is KCenters.fit wall time incl. the D2H of 1M assignments / "
                    "distances and building 1000 centre rows; device_loop_* is the persistent "
                    "multi-iteration kernel alone",
            "n_centers": len(est.result_.center_indices)}


def c3(n=1_000_000, A=500, k=1000, sweeps=5):
    """KHybrid rmsd k=1000, 5 sweeps, 1M x 500 (config 3)."""
    from enspara_b200 import synth
    from enspara_b200.cluster import KCenters, KHybrid
    X = synth.device_trajectory(n, A, seed=0)
    KHybrid("rmsd", n_clusters=4, kmedoids_updates=1, random_state=0).fit(
        synth.device_trajectory(4096, A, seed=1))
    kc, dt_kc = timed(lambda: KCenters("rmsd", n_clusters=k).fit(X))
    est, dt = timed(lambda: KHybrid("rmsd", n_clusters=k, kmedoids_updates=sweeps,
                                    random_state=0).fit(X))
    dt_pam = dt - dt_kc
    d0 = kc.result_.distances
    d1 = est.result_.distances
    return {"config": "C3 KHybrid rmsd k=%d, %d sweeps, %d x %d (resident)" % (k, sweeps, n, A),
            "seconds_total": dt, "seconds_kcenters_phase": dt_kc,
            "kcenters_evals_per_s": n * k / dt_kc,
            "kcenters_frac_of_hbm_peak": n * k / dt_kc * (12 * A + 8) / 1e9 / PEAK,
            "seconds_pam_phase": dt_pam, "ms_per_proposal": 1e3 * dt_pam / (k * sweeps),
            "proposal_full_pass_evals_per_s": n * k * sweeps / dt_pam,
            "msq_cost_before": float(np.mean(d0 * d0)), "msq_cost_after": float(np.mean(d1 * d1)),
            "n_centers": len(est.result_.center_indices)}


def c5(n=1_250_000, A=500, k=10_000):
    """assign_to_nearest_center, 10k centres vs one GPU's shard of 10M x 500 (config 5)."""
    from enspara_b200 import synth
    from enspara_b200.cluster import _ops, util
    X = synth.device_trajectory(n, A, seed=0)
    idx = torch.arange(0, n, n // k, device="cuda")[:k]
    cen = X.gather(idx)
    _ops.assign_device_auto(util.RMSD, synth.device_trajectory(8192, A, seed=2), cen)
    stats = {}
    (d, a), dt = timed(lambda: _ops.assign_device_tc(util.RMSD, X, cen, stats=stats))
    evs = n * k / dt
    # every centre is a frame of X: it must be assigned to itself at distance ~0
    self_ok = bool((a[idx].cpu() == torch.arange(k, dtype=torch.int32)).all())
    out = {"config": "C5 assign_to_nearest_center, %d centres x %d frames x %d atoms (one "
                     "GPU's shard of 10M; centres replicated, no exchange)" % (k, n, A),
           "seconds": dt, "evals_per_s": evs, "algorithmic_TFLOPs": evs * 18 * A / 1e12,
           "executed_FP16_TFLOPs": evs * 54 * 512 / 1e12,
           "frac_of_measured_bf16_peak_executed": evs * 54 * 512 / 1e12 / TENSOR_PEAK,
           "frac_of_measured_bf16_peak_algorithmic": evs * 18 * A / 1e12 / TENSOR_PEAK,
           "centres_assigned_to_themselves": self_ok,
           "survivors_per_frame": stats.get("survivors_mean"),
           "overflow_frames": stats.get("overflow_frames")}
    # the same pass with centres CHOSEN BY K-CENTERS (spread out: what assign/reassign sees in
    # practice; evenly spaced frames of the 64-conformer synthetic set are ~156 near-duplicates
    # per conformer, the worst case for any bound-based screen)
    from enspara_b200.cluster import kcenters as kcm
    kk = 2000
    (res, eng), dt_kc = timed(lambda: kcm.kcenters(X, "rmsd", n_clusters=kk, _return_engine=True))
    cen2 = X.gather(torch.as_tensor([int(c) for c in res.center_indices], device="cuda"))
    stats2 = {}
    (d2, a2), dt2 = timed(lambda: _ops.assign_device_tc(util.RMSD, X, cen2, stats=stats2))
    out["kcenters_chosen_centres"] = {
        "k": kk, "seconds": dt2, "evals_per_s": n * kk / dt2,
        "survivors_per_frame": stats2.get("survivors_mean"),
        "overflow_frames": stats2.get("overflow_frames"),
        "identical_to_kcenters_state": bool(torch.equal(a2, eng.assign))
        and bool(torch.equal(d2, eng.dist)),
        "kcenters_seconds": dt_kc, "kcenters_evals_per_s": n * kk / dt_kc}
    del eng, res, cen2, d2, a2
    # PAM refinement on the same shard: a handful of proposals with k = 10k medoids
    from enspara_b200.cluster._pam import PamEngine
    from enspara_b200.cluster.kcenters import _SingleComm
    pam = PamEngine(X, util.RMSD, _SingleComm(), d, a, [int(i) for i in idx.cpu()])
    nprop = 16
    _, dtp = timed(lambda: pam.sweep(random_state=0, max_proposals=nprop))
    out["pam_ms_per_proposal_k10000"] = 1e3 * dtp / nprop
    return out


def tri(n=1_000_000, A=500, k=1000):
    """use_triangle_inequality: same result, most frames never read."""
    from enspara_b200 import synth
    from enspara_b200.cluster import kcenters as kcm
    X = synth.device_trajectory(n, A, seed=0)
    kcm.kcenters(X, "rmsd", n_clusters=8, use_triangle_inequality=True)
    a, dt_plain = timed(lambda: kcm.kcenters(X, "rmsd", n_clusters=k))
    b, dt_tri = timed(lambda: kcm.kcenters(X, "rmsd", n_clusters=k,
                                           use_triangle_inequality=True))
    same = [int(c) for c in a.center_indices] == [int(c) for c in b.center_indices] and \
        bool(np.array_equal(a.assignments, b.assignments)) and \
        bool(np.array_equal(a.distances, b.distances))
    return {"config": "KCenters rmsd k=%d, %d x %d, use_triangle_inequality" % (k, n, A),
            "seconds_plain": dt_plain, "seconds_triangle": dt_tri,
            "speedup": dt_plain / dt_tri, "identical_results": same,
            "effective_evals_per_s_triangle": n * k / dt_tri}


def reassign(tmp="/tmp/eb_reassign"):
    """Streaming reassign of on-disk .npy trajectories."""
    from enspara_b200 import synth
    from enspara_b200.cluster import reassign as rz
    os.makedirs(tmp, exist_ok=True)
    A, per, nfiles, k = 500, 100_000, 8, 1000
    files = []
    for i in range(nfiles):
        p = os.path.join(tmp, "t%d.npy" % i)
        if not os.path.exists(p):
            np.save(p, synth.device_trajectory_aos(per, A, 0, i * per).cpu().numpy())
        files.append(p)
    centers = np.load(files[0], mmap_mode="r")[::per // k][:k].copy()
    lengths = [per] * nfiles
    rz.batch_reassign([(f, None, None) for f in files[:1]], centers, lengths[:1], 0.5)
    orig = rz.determine_batch_size
    rz.determine_batch_size = lambda *a, **kw: (2 * per + 1, 0.0)   # 4 batches of 2 files
    stats = {}
    try:
        (a, d), dt = timed(lambda: rz.batch_reassign([(f, None, None) for f in files], centers,
                                                     lengths, 0.5, stats=stats))
    finally:
        rz.determine_batch_size = orig
    n = per * nfiles
    # steady state: a second run reuses nothing but shows the cost without the one-off pinned
    # allocation (2 x %d-frame staging buffers)
    rz.determine_batch_size = lambda *a, **kw: (2 * per + 1, 0.0)
    stats2 = {}
    try:
        _, dt_b = timed(lambda: rz.batch_reassign([(f, None, None) for f in files], centers,
                                                  lengths, 0.5, stats=stats2))
    finally:
        rz.determine_batch_size = orig
    stats["seconds_second_run"] = dt_b
    return {"config": "reassign: %d files x %d frames x %d atoms from disk (.npy, page cache) "
                      "vs %d centres" % (nfiles, per, A, k),
            "seconds": dt, "frames_per_s": n / dt, "evals_per_s": n * k / dt,
            "disk_read_GBps": n * A * 12 / dt / 1e9, **stats}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--out", default="gpurun_out/configs.json")
    p.add_argument("--only", default="c1,c2,c3,c5,reassign")
    p.add_argument("--no-cpu", action="store_true",
                   help="skip the CPU baselines beside c1/c2/c3/c5 (scripts/cpu_baselines.py)")
    args = p.parse_args()
    torch.cuda.set_device(0)
    fns = {"c1": c1, "c2": c2, "c3": c3, "c5": c5, "reassign": reassign, "tri": tri,
           "peaks": tensor_peaks}
    results = {}
    for name in args.only.split(","):
        try:
            results[name] = fns[name]()
        except Exception as exc:  # keep going: one failing config must not hide the others
            import traceback
            results[name] = {"error": repr(exc), "trace": traceback.format_exc()[-1500:]}
        if not args.no_cpu and name in ("c1", "c2", "c3", "c5") and "error" not in results[name]:
            # the reference's CPU path for the same config, same run, same host
            from scripts import cpu_baselines
            results[name]["cpu_baseline"] = cpu_baselines.run((name,)).get(name)
        print(name, json.dumps(results[name]), flush=True)
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
