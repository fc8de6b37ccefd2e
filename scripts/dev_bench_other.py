"""Developer probe: K2 (euclidean k-centers), K3 (many-centres assign), PAM sweep timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from enspara_b200 import synth
from enspara_b200.cluster import util, _ops, kcenters as kc, kmedoids as km
from enspara_b200.cluster._engine import KCentersEngine
from enspara_b200.cluster._pam import PamEngine
from enspara_b200.cluster.kcenters import _SingleComm


def ev():
    return torch.cuda.Event(enable_timing=True)


def bench_k2(n=1_000_000, F=64, k=200):
    X = synth.device_features(n, F, seed=0)
    for metric in ("euclidean", "manhattan"):
        eng = KCentersEngine(X, metric, _SingleComm())
        eng.run(5, 0.0)
        eng = KCentersEngine(X, metric, _SingleComm())
        e0, e1 = ev(), ev()
        torch.cuda.synchronize(); e0.record()
        c, md = eng.run(k, 0.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        evs = n * k / (ms * 1e-3)
        st = eng.read_state()
        print("K2 %s n=%d F=%d k=%d: %.1f us/iter %.2f G evals/s %.0f GB/s (%.1f%% of 6543); "
              "block 0 between two bodies: %.2f us/iter"
              % (metric, n, F, k, 1e3 * ms / k, evs / 1e9, evs * (4 * F + 8) / 1e9,
                 100 * evs * (4 * F + 8) / 1e9 / 6543.1, 1e-3 * st.wait_ns / max(1, k - 1)),
              flush=True)


def bench_k3(n=200_000, A=500, k=1000):
    data = synth.device_trajectory(n, A, seed=0)
    cen = data.gather(torch.arange(0, n, n // k, device="cuda")[:k])
    for _ in range(2):
        e0, e1 = ev(), ev()
        torch.cuda.synchronize(); e0.record()
        d, a = _ops.assign_device(util.RMSD, data, cen)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    evs = n * k / (ms * 1e-3)
    print("K3 exact n=%d A=%d k=%d: %.1f ms  %.3f G evals/s  (%.1f TFLOP/s fp64-equivalent 18A)"
          % (n, A, k, ms, evs / 1e9, evs * 18 * A / 1e12), flush=True)
    return data


def bench_pam(data, k=200):
    n = len(data)
    t = time.perf_counter()
    res, eng = kc.kcenters(data, "rmsd", n_clusters=k, _return_engine=True)
    torch.cuda.synchronize()
    t_kc = time.perf_counter() - t
    pam = PamEngine(data, util.RMSD, _SingleComm(), eng.dist, eng.assign,
                    [int(c) for c in res.center_indices])
    t = time.perf_counter()
    acc = pam.sweep(random_state=0)
    torch.cuda.synchronize()
    t_sw = time.perf_counter() - t
    print("PAM n=%d A=%d k=%d: kcenters %.3f s, one sweep %.3f s (%.2f ms/proposal, %d accepted)"
          % (n, data.n_atoms, k, t_kc, t_sw, 1e3 * t_sw / k, acc), flush=True)


if __name__ == "__main__":
    bench_k2()
    data = bench_k3()
    bench_pam(data)
    data = synth.device_trajectory(1_000_000, 500, seed=0)
    bench_pam(data, k=100)
