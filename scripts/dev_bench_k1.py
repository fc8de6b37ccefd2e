"""Developer probe: raw K1 step throughput on device-generated frames (not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from enspara_b200.device import DeviceTrajectory
from enspara_b200.cluster._engine import KCentersEngine
from enspara_b200.cluster.kcenters import _SingleComm


def make(n, A):
    d = DeviceTrajectory.empty(n, A)
    per = 100000
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for lo in range(0, n, per):
        m = min(per, n - lo)
        aos = torch.randn((m, A, 3), device="cuda", generator=g) * 2.0
        d.ingest_aos(aos, lo)
    torch.cuda.synchronize()
    return d


def run(n, A, k, exact):
    d = make(n, A)
    eng = KCentersEngine(d, "rmsd", _SingleComm(), exact=exact)
    eng.run(5, 0.0)  # warm-up
    eng = KCentersEngine(d, "rmsd", _SingleComm(), exact=exact)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    c, md = eng.run(k, 0.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    evals = n * k / (ms * 1e-3)
    gbs = evals * (12 * A + 8) / 1e9
    print("n=%d A=%d k=%d exact=%d: %.3f ms/iter  %.3f G evals/s  %.0f GB/s algorithmic (%.1f%% of 6543)  centers[:4]=%s maxdist=%.4f"
          % (n, A, k, exact, ms / k, evals / 1e9, gbs, 100 * gbs / 6543.1, c[:4], md), flush=True)


if __name__ == "__main__":
    print("EB_K1_VARIANT=%s" % os.environ.get("EB_K1_VARIANT", "0"))
    sizes = ((1250000, 500), (1250000, 264), (1000000, 100)) if len(sys.argv) < 2 else \
        [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for (n, A) in sizes:
        for exact in (True, False):
            run(n, A, 50, exact)
