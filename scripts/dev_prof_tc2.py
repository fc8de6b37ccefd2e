"""One dense tensor-core assignment pass for ncu (screen + re-score), 262144 x 1008 x 500."""
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
od = torch.empty(n, dtype=torch.float32, device="cuda")
oa = torch.empty(n, dtype=torch.int32, device="cuda")
cand = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(2):
    _lib.call("eb_rmsd_assign_tc", ptr(data.xyz), ptr(data.traces), n, A, ptr(cen.xyz),
              ptr(cen.traces), k, float(_ops.tc_kappa(data.a_pad)), None, 0, ptr(od), ptr(oa),
              ptr(cand), ptr(scratch), None, 1, stream_ptr())
torch.cuda.synchronize()
