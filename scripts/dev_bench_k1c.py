"""Developer probe: sustained (power-capped) K1 throughput, 1.25M x 500, k = 600 iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from enspara_b200 import synth
from enspara_b200.cluster._engine import KCentersEngine
from enspara_b200.cluster.kcenters import _SingleComm
n, k = 1_250_000, 600
d = synth.device_trajectory(n, 500, seed=0)
eng = KCentersEngine(d, "rmsd", _SingleComm())
eng.run(200, 0.0)          # heat up
for rep in range(2):
    eng = KCentersEngine(d, "rmsd", _SingleComm())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    eng.run(k, 0.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    print("EB_K1_FR=%s EB_K1_TMA=%s sustained: %.4f ms/iter (%.3f of 6550)" % (
        os.environ.get("EB_K1_FR", "-"), os.environ.get("EB_K1_TMA", "-"), ms,
        n * 6008 / ms / 1e6 / 6550.4), flush=True)
