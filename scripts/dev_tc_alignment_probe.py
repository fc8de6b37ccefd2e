"""How many extra bits does the tcgen05 FP32 accumulation keep when it aligns the 16 products of
one kind::f16 MMA k-step (block FMA model of Fasi, Higham et al. 2021)?  One big product 2^30 and
fifteen small ones 2^(6-t) in the SAME k-step: a small product survives the alignment iff the
adder keeps at least t+1 bits below the FP32 ulp of the big one (ulp = 2^7).  Reads the raw
accumulators with mode 0 of eb_rmsd_assign_tc."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from enspara_b200 import _lib
from enspara_b200.device import DeviceTrajectory, ptr, stream_ptr


def probe(t, n_small=15, second_step=False):
    L = _lib.load()
    A, n, k = 32, 128, 48
    fx = torch.zeros((n, 3, A), dtype=torch.float32, device="cuda")
    cx = torch.zeros((k, 3, A), dtype=torch.float32, device="cuda")
    # operands are scaled by 2^8 inside the pack kernel: x = 2^7 -> h1 = 2^15
    fx[:, 0, 0] = 2.0 ** 7
    cx[:, 0, 0] = 2.0 ** 7                      # product 2^30
    lo = 16 if second_step else 1               # small products in the other k-step or the same
    fx[:, 0, lo:lo + n_small] = 2.0 ** -5       # h1 = 2^3
    cx[:, 0, lo:lo + n_small] = 2.0 ** (-5 - t) # h1 = 2^(3-t): product 2^(6-t)
    data = DeviceTrajectory(fx, torch.ones(n, dtype=torch.float64, device="cuda"), A)
    cen = DeviceTrajectory(cx, torch.ones(k, dtype=torch.float64, device="cuda"), A)
    scratch = torch.zeros(int(L.eb_tc_scratch_bytes(n, A, k)), dtype=torch.uint8, device="cuda")
    dbg = torch.zeros((n, k, 9), dtype=torch.float32, device="cuda")
    _lib.call("eb_rmsd_assign_tc", ptr(data.xyz), ptr(data.traces), n, A, ptr(cen.xyz),
              ptr(cen.traces), k, 0.0, None, 0, None, None, None, ptr(scratch), ptr(dbg), 0,
              stream_ptr())
    torch.cuda.synchronize()
    got = float(dbg[0, 0, 0].double()) * 65536.0      # mode 0 un-scales by 2^-16
    exact = 2.0 ** 30 + n_small * 2.0 ** (6 - t)
    return got - 2.0 ** 30, exact - 2.0 ** 30


out = {}
for t in range(0, 8):
    g, e = probe(t)
    g2, e2 = probe(t, second_step=True)
    out[t] = {"same_step": [g, e], "other_step": [g2, e2]}
    print("t=%d small product = 2^%d (ulp of the sum = 2^7): same k-step: got +%g of exact +%g ; "
          "other k-step: got +%g of +%g" % (t, 6 - t, g, e, g2, e2), flush=True)
json.dump(out, open("gpurun_out/r2_tc_alignment_probe.json", "w"), indent=1)
