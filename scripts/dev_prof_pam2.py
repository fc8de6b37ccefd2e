"""Developer probe: GPU time per kernel of PAM proposals (torch profiler / CUPTI), and the wall
time per proposal, at config 3 (1M x 500, k = 1000)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from enspara_b200 import synth
from enspara_b200.cluster import util, kcenters as kc
from enspara_b200.cluster._pam import PamEngine
from enspara_b200.cluster.kcenters import _SingleComm

n, A, k, nprop = 1_000_000, 500, int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 200
data = synth.device_trajectory(n, A, seed=0)
res, eng = kc.kcenters(data, "rmsd", n_clusters=k, _return_engine=True)
pam = PamEngine(data, util.RMSD, _SingleComm(), eng.dist, eng.assign,
                [int(c) for c in res.center_indices])
pam.sweep(random_state=0, max_proposals=20)
torch.cuda.synchronize()
t = time.perf_counter()
pam.sweep(random_state=1, max_proposals=nprop)
torch.cuda.synchronize()
wall = (time.perf_counter() - t) / nprop
print("host issue per proposal: %.1f us; sync wait per proposal: %.1f us; count refresh per "
      "proposal: %.1f us" % (1e6 * pam.host_issue_s / (nprop + 20),
                             1e6 * pam.sync_wait_s / (nprop + 20),
                             1e6 * pam.refresh_s / (nprop + 20)))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pam.sweep(random_state=2, max_proposals=nprop)
    torch.cuda.synchronize()
rows = []
tot = 0.0
for e in prof.key_averages():
    if e.device_time_total > 0:
        rows.append((e.device_time_total / nprop, e.count / nprop, e.key[:70]))
        tot += e.device_time_total / nprop
rows.sort(reverse=True)
print("wall per proposal: %.1f us; GPU busy per proposal: %.1f us" % (1e6 * wall, tot))
for us, cnt, name in rows[:22]:
    print("%8.1f us  x%.2f  %s" % (us, cnt, name))
