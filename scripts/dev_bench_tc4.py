"""Developer probe: how much of the screen kernel is its QCP epilogue (mode 2 = no epilogue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from enspara_b200 import _lib, synth
from enspara_b200.cluster import _ops
from enspara_b200.device import ptr, stream_ptr
n, A, k = 262144, 500, 1008
data = synth.device_trajectory(n, A, seed=0)
cen = data.gather(torch.arange(0, n, n // k, device="cuda")[:k])
L = _lib.load()
scratch = torch.empty(int(L.eb_tc_scratch_bytes(n, A, k)), dtype=torch.uint8, device="cuda")
od = torch.empty(n, dtype=torch.float32, device="cuda"); oa = torch.empty(n, dtype=torch.int32, device="cuda")
cand = torch.empty(n, dtype=torch.int32, device="cuda")
dbg = torch.zeros(16, dtype=torch.float32, device="cuda")
kappa = _ops.tc_kappa(data.a_pad)
for mode in (1, 2, 1, 2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(3):
        _lib.call("eb_rmsd_assign_tc", ptr(data.xyz), ptr(data.traces), n, A, ptr(cen.xyz),
                  ptr(cen.traces), k, float(kappa), None, 0, ptr(od), ptr(oa), ptr(cand),
                  ptr(scratch), ptr(dbg), mode, stream_ptr())
    ev[1].record(); torch.cuda.synchronize()
    print("mode %d: %.2f ms per call (mode 1 = pack + screen + rescore, mode 2 = pack + screen "
          "without QCP epilogue)" % (mode, ev[0].elapsed_time(ev[1]) / 3), flush=True)
