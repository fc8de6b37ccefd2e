import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from enspara_b200 import synth
from enspara_b200.cluster import _ops, util

def run(n, A, k):
    data = synth.device_trajectory(n, A, seed=0)
    cen = data.gather(torch.arange(0, n, max(1, n // k), device="cuda")[:k])
    for name, fn in (("tc   ", _ops.assign_device_tc), ("exact", _ops.assign_device)):
        if name == "exact" and n * k > 4e8:
            continue
        fn(util.RMSD, data, cen)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        stats = {}
        if name.startswith("tc"):
            d, a = fn(util.RMSD, data, cen, stats=stats)
        else:
            d, a = fn(util.RMSD, data, cen)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ev = n * k / (ms * 1e-3)
        print("%s n=%d A=%d k=%d: %.1f ms  %.2f G evals/s  %.1f TFLOP/s (18A flop/eval; x3 executed for 3xTF32) %s"
              % (name, n, A, k, ms, ev / 1e9, ev * 18 * A / 1e12, stats), flush=True)

def run_kcenters_centers(n, A, k):
    """Centres chosen by k-centers (spread out, like the real use of assign/reassign)."""
    from enspara_b200.cluster import kcenters as kc
    data = synth.device_trajectory(n, A, seed=0)
    res, eng = kc.kcenters(data, "rmsd", n_clusters=k, _return_engine=True)
    cen = data.gather(torch.as_tensor([int(c) for c in res.center_indices], device="cuda"))
    _ops.assign_device_tc(util.RMSD, data, cen)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    stats = {}
    d, a = _ops.assign_device_tc(util.RMSD, data, cen, stats=stats)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    same = bool(torch.equal(a, eng.assign)) and bool(torch.equal(d, eng.dist))
    print("tc (k-centers centres) n=%d A=%d k=%d: %.1f ms  %.2f G evals/s  identical to k-centers state: %s %s"
          % (n, A, k, ms, n * k / ms / 1e6, same, stats), flush=True)


if __name__ == "__main__":
    run(200_000, 500, 1000)
    run(1_000_000, 500, 1000)
    run(1_000_000, 500, 10000)
    run_kcenters_centers(1_000_000, 500, 1000)
