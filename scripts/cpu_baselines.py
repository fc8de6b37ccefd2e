#!/usr/bin/env python
"""CPU baselines beside every BASELINE config (SURVEY.md 8d "CPU baseline beside it"): the
reference's path on THIS host's cores, bounded samples (~5-15 s each), evals/s.

  C1  full size: reference k-centers loop (restated, kcenters.py:243-311) around the restated
      mdtraj RMSD (float32 SSE lanes + OpenMP, incl. mdtraj's per-call copy + centre)
  C2  the reference's OWN compiled Cython libdist.euclidean (oracle/_ref, libdist.pyx:122-164)
      inside the restated loop: `kind: "reference"` for the distance code, which is > 95 % of
      the time; the Python package itself cannot travel to the GPU box
  C3  k-centers phase at reduced n; one PAM sweep (kmedoids.py:520-699, restated) at reduced
      n and k
  C5  assign_to_nearest_center (cluster/util.py:159-205, restated) at reduced n and k

No GPU and no product code: only oracle/ (test infrastructure) is used.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def _rmsd_note():
    return ("restated mdtraj RMSD (float32 SSE 4-atom lanes, OpenMP over frames, per-call "
            "copy+centre like md.rmsd); mdtraj itself is not installable offline")


def c1(od, oc, cores):
    X = od.synth_trajectory(20_000, 264)
    T = od.Trajectory(X)
    t = time.perf_counter()
    r = oc.kcenters(T, od.rmsd_f32_sse, n_clusters=100)
    dt = time.perf_counter() - t
    return {"value": 20_000 * 100 / dt, "unit": "evals/s", "cores": cores, "kind": "port",
            "seconds": dt, "sample": "the full config: 100 iterations over 20000 x 264; "
            + _rmsd_note(), "n_centers": len(r.center_indices)}


def c2(od, oc, cores, seconds=8.0):
    from oracle import refharness
    ld = refharness.load_compiled_libdist()
    X = od.synth_features(1_000_000, 64)
    t = time.perf_counter()
    oc.kcenters(X, ld.euclidean, n_clusters=3)
    one = (time.perf_counter() - t) / 3
    k = int(max(5, min(1000, seconds / one)))
    t = time.perf_counter()
    r = oc.kcenters(X, ld.euclidean, n_clusters=k)
    dt = time.perf_counter() - t
    return {"value": 1_000_000 * k / dt, "unit": "evals/s", "cores": cores, "kind": "reference",
            "seconds": dt, "ms_per_iteration": 1e3 * dt / k,
            "algorithmic_GBps": 1_000_000 * k / dt * 264 / 1e9,
            "sample": "%d k-centers iterations over the full 1M x 64 float32 matrix: the "
                      "reference's own compiled Cython libdist.euclidean (oracle/_ref, built "
                      "from /root/reference/enspara/geometry/libdist.pyx) inside the restated "
                      "loop of kcenters.py:243-311" % k, "n_centers": len(r.center_indices)}


def c3(od, oc, cores):
    A = 500
    n = 150_000
    X = od.synth_trajectory(n, A)
    T = od.Trajectory(X)
    kk = 24
    t = time.perf_counter()
    r = oc.kcenters(T, od.rmsd_f32_sse, n_clusters=kk)
    dt_kc = time.perf_counter() - t
    # one PAM sweep on a smaller problem (k metric calls per proposal on the ambiguous subset)
    n2, k2 = 60_000, 32
    T2 = od.Trajectory(X[:n2])
    r2 = oc.kcenters(T2, od.rmsd_f32_sse, n_clusters=k2)
    t = time.perf_counter()
    oc.pam_update(T2, od.rmsd_f32_sse, list(r2.center_indices), r2.assignments.copy(),
                  r2.distances.copy(), random_state=0)
    dt_pam = time.perf_counter() - t
    return {"kcenters_phase": {"value": n * kk / dt_kc, "unit": "evals/s", "cores": cores,
                               "kind": "port", "seconds": dt_kc,
                               "sample": "%d iterations over %d x %d; " % (kk, n, A)
                               + _rmsd_note()},
            "pam_phase": {"value": n2 * k2 / dt_pam, "unit": "proposal full-pass evals/s",
                          "ms_per_proposal": 1e3 * dt_pam / k2, "cores": cores, "kind": "port",
                          "seconds": dt_pam,
                          "sample": "one sweep of %d proposals over %d x %d (kmedoids.py:520-699 "
                                    "restated: full pass + k metric calls on the ambiguous "
                                    "subset per proposal); " % (k2, n2, A) + _rmsd_note()}}


def c5(od, oc, cores):
    A, n, k = 500, 40_000, 250
    X = od.synth_trajectory(n, A)
    T = od.Trajectory(X)
    centers = [T[int(i)] for i in np.linspace(0, n - 1, k).astype(int)]
    t = time.perf_counter()
    oc.assign_to_nearest_center(T, centers, od.rmsd_f32_sse)
    dt = time.perf_counter() - t
    return {"value": n * k / dt, "unit": "evals/s", "cores": cores, "kind": "port",
            "seconds": dt, "sample": "assign_to_nearest_center of %d x %d frames to %d centres "
            "(cluster/util.py:159-205 restated: one md.rmsd call per centre); " % (n, A, k)
            + _rmsd_note()}


def run(only=("c1", "c2", "c3", "c5")):
    from oracle import cluster as oc
    from oracle import distances as od
    cores = od.use_all_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)      # the reference's libdist reads it
    out = {}
    for name, fn in (("c1", c1), ("c2", c2), ("c3", c3), ("c5", c5)):
        if name not in only:
            continue
        try:
            out[name] = fn(od, oc, cores)
        except Exception as exc:
            out[name] = {"error": repr(exc)}
        print(name, json.dumps(out[name]), flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/r2_cpu_baselines.json")
    ap.add_argument("--only", default="c1,c2,c3,c5")
    a = ap.parse_args()
    res = run(tuple(a.only.split(",")))
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
