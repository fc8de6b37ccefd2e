"""Developer probe: K1 device-loop time per iteration vs number of iterations and shard size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from enspara_b200 import synth
from enspara_b200.cluster._engine import KCentersEngine
from enspara_b200.cluster.kcenters import _SingleComm
for n in (1_000_000, 1_250_000):
    d = synth.device_trajectory(n, 500, seed=0)
    for k in (50, 200, 1000):
        eng = KCentersEngine(d, "rmsd", _SingleComm())
        eng.run(5, 0.0)
        eng = KCentersEngine(d, "rmsd", _SingleComm())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        eng.run(k, 0.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k
        print("n=%d k=%d: %.4f ms/iter = %.1f GB/s algorithmic (%.3f of 6550)" % (
            n, k, ms, n * 6008 / ms / 1e6, n * 6008 / ms / 1e6 / 6550.4), flush=True)
    del d
