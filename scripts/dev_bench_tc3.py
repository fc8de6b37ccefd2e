"""Developer probe: per-kernel durations of the screen pass in a live (un-replayed) run, via
torch.profiler (CUPTI), plus SM clocks before / after."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from enspara_b200 import synth
from enspara_b200.cluster import _ops, util


def clocks():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active",
                               "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    except Exception as exc:
        return repr(exc)


n, A, k = 262144, 500, int(sys.argv[1]) if len(sys.argv) > 1 else 1008
data = synth.device_trajectory(n, A, seed=0)
cen = data.gather(torch.arange(0, n, n // k, device="cuda")[:k])
ws = {}
_ops.assign_device_tc(util.RMSD, data, cen, workspace=ws)
torch.cuda.synchronize()
print("clocks before:", clocks())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(3):
        _ops.assign_device_tc(util.RMSD, data, cen, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
print("clocks after:", clocks())
print("events: %.2f ms per pass" % (e0.elapsed_time(e1) / 3))
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
for r in rows[:10]:
    print("%-70s n=%3d total=%9.1f us mean=%9.1f us" % (r.key[:70], r.count, r.device_time_total,
                                                        r.device_time_total / max(r.count, 1)))
