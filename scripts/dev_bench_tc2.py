"""Developer probe: screen pass timing (events) for one shape; EB_TC_SEG selects the segment count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from enspara_b200 import synth
from enspara_b200.cluster import _ops, util
n, A, k = 262144, 500, int(sys.argv[1]) if len(sys.argv) > 1 else 1008
data = synth.device_trajectory(n, A, seed=0)
cen = data.gather(torch.arange(0, n, n // k, device="cuda")[:k])
ws = {}
_ops.assign_device_tc(util.RMSD, data, cen, workspace=ws)
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.synchronize()
e[0].record()
for _ in range(3):
    stats = {}
    _ops.assign_device_tc(util.RMSD, data, cen, workspace=ws, stats=stats)
e[1].record()
torch.cuda.synchronize()
ms = e[0].elapsed_time(e[1]) / 3
print("EB_TC_SEG=%s n=%d k=%d: %.2f ms per pass (pack + screen + rescore)  %.2f G evals/s %s"
      % (os.environ.get("EB_TC_SEG", "default"), n, k, ms, n * k / ms / 1e6, stats), flush=True)
