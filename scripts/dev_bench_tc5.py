"""Developer probe: per-kernel times of one dense tensor-core assignment pass (pack / screen /
re-score), survivors per frame for several kappa, and the end-to-end rate at the C5 shard shape."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from enspara_b200 import _lib, synth
from enspara_b200.cluster import _ops, util
from enspara_b200.device import ptr, stream_ptr

n, A, k = 262144, 500, 1008
data = synth.device_trajectory(n, A, seed=0)
cen = data.gather(torch.arange(0, n, n // k, device="cuda")[:k])
L = _lib.load()
scratch = torch.empty(int(L.eb_tc_scratch_bytes(n, A, k)), dtype=torch.uint8, device="cuda")
od = torch.empty(n, dtype=torch.float32, device="cuda")
oa = torch.empty(n, dtype=torch.int32, device="cuda")
cand = torch.empty(n, dtype=torch.int32, device="cuda")
dbg = torch.zeros(16, dtype=torch.float32, device="cuda")
out = {}


def one(kappa, mode=1):
    _lib.call("eb_rmsd_assign_tc", ptr(data.xyz), ptr(data.traces), n, A, ptr(cen.xyz),
              ptr(cen.traces), k, float(kappa), None, 0, ptr(od), ptr(oa), ptr(cand),
              ptr(scratch), ptr(dbg), mode, stream_ptr())


for name, kappa in (("rigorous", _ops.tc_kappa(data.a_pad)), ("round1_8x", 8 * 2.0 ** -24 * data.a_pad),
                    ("2x_measured", 2 * 2.0 ** -24 * data.a_pad)):
    one(kappa)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            one(kappa)
        torch.cuda.synchronize()
    times = {}
    for e in prof.key_averages():
        times[e.key.split("(")[0][-40:]] = round(e.device_time_total / e.count / 1e3, 3)
    surv = float(cand.clamp(min=0).float().mean())
    if name == "rigorous":
        c = cand.clamp(min=0)
        passes = (c + 1) // 2
        per_warp = passes.view(-1, 4).max(dim=1).values.float().mean()
        q = torch.quantile(c.float()[::64], torch.tensor([0.5, 0.9, 0.99, 1.0], device="cuda"))
        print("re-score load balance: mean passes per frame %.2f, mean of max over a warp's 4 "
              "frames %.2f (x%.2f); survivors p50/p90/p99/max = %s"
              % (float(passes.float().mean()), float(per_warp),
                 float(per_warp) / float(passes.float().mean()), q.tolist()), flush=True)
    ovf = int((cand < 0).sum())
    out[name] = {"kappa": kappa, "kernel_ms": times, "survivors_per_frame": surv, "overflow": ovf}
    print(name, json.dumps(out[name]), flush=True)
one(_ops.tc_kappa(data.a_pad), mode=2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        one(_ops.tc_kappa(data.a_pad), mode=2)
    torch.cuda.synchronize()
print("mode2 (no QCP epilogue):", {e.key.split("(")[0][-40:]: round(e.device_time_total / e.count / 1e3, 3)
                                   for e in prof.key_averages()}, flush=True)

# end to end through the host wrapper (audit on), C5 shard shape: 1.25M frames x 10000 centres
del scratch
n2, k2 = 1_250_000, 10_000
data2 = synth.device_trajectory(n2, A, seed=0)
cen2 = data2.gather(torch.arange(0, n2, n2 // k2, device="cuda")[:k2])
ws = {}
for rep in range(2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); ev[0].record()
    d, a = _ops.assign_device_tc(util.RMSD, data2, cen2, workspace=ws)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    print("C5 shard: %d x %d in %.1f ms = %.2f G evals/s (audit level %d, audited passes %d)"
          % (n2, k2, ms, n2 * k2 / ms / 1e6, _ops._audit_level(), _ops.audit_stats["passes_audited"]), flush=True)
out["c5_shard_ms"] = ms
out["c5_shard_gevals"] = n2 * k2 / ms / 1e6
json.dump(out, open("gpurun_out/r2_tc_probe.json", "w"), indent=1)
