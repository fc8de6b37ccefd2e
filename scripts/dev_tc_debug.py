"""Developer probe for the tcgen05 screen: mode 0 dumps the approximate inner-product matrices."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from enspara_b200 import _lib, synth
from enspara_b200.device import DeviceTrajectory, ptr, stream_ptr

def run(n, k, A):
    L = _lib.load()
    X = synth.trajectory(n, A, seed=3)
    data = DeviceTrajectory.from_host(X)
    cidx = np.linspace(0, n - 1, k).astype(np.int64)
    cen = data.gather(cidx)
    scratch = torch.zeros(int(L.eb_tc_scratch_bytes(n, A, k)), dtype=torch.uint8, device="cuda")
    dbg = torch.zeros((n, k, 9), dtype=torch.float32, device="cuda")
    _lib.call("eb_rmsd_assign_tc", ptr(data.xyz), ptr(data.traces), n, A, ptr(cen.xyz),
              ptr(cen.traces), k, 0.0, None, 0, None, None, None, ptr(scratch), ptr(dbg), 0,
              stream_ptr())
    torch.cuda.synchronize()
    xs = data.xyz.cpu().numpy().astype(np.float64)      # (n,3,Ap)
    cs = cen.xyz.cpu().numpy().astype(np.float64)       # (k,3,Ap)
    M = np.einsum("fia,cja->fcij", xs, cs).reshape(n, k, 9)
    got = dbg.cpu().numpy().astype(np.float64)
    G = np.sqrt(data.traces.cpu().numpy()[:, None] * cen.traces.cpu().numpy()[None, :])
    err = np.abs(got - M).max(axis=2) / G
    print("n=%d k=%d A=%d: max |dM|/sqrt(GaGb) = %.3e  mean = %.3e   (M scale %.1f)"
          % (n, k, A, err.max(), err.mean(), np.abs(M).max()))
    bad = np.argwhere(err > 1e-3)
    if len(bad):
        print("  first bad pairs:", bad[:5].tolist())
        f, c = bad[0]
        print("  got ", got[f, c]); print("  want", M[f, c])

if __name__ == "__main__":
    run(300, 70, 500)
    run(1000, 33, 500)
    run(129, 1, 496)


def lam_max(M):
    S = M.reshape(-1, 3, 3)
    K = np.empty((len(S), 4, 4))
    K[:, 0, 0] = S[:, 0, 0] + S[:, 1, 1] + S[:, 2, 2]
    K[:, 0, 1] = K[:, 1, 0] = S[:, 1, 2] - S[:, 2, 1]
    K[:, 0, 2] = K[:, 2, 0] = S[:, 2, 0] - S[:, 0, 2]
    K[:, 0, 3] = K[:, 3, 0] = S[:, 0, 1] - S[:, 1, 0]
    K[:, 1, 1] = S[:, 0, 0] - S[:, 1, 1] - S[:, 2, 2]
    K[:, 1, 2] = K[:, 2, 1] = S[:, 0, 1] + S[:, 1, 0]
    K[:, 1, 3] = K[:, 3, 1] = S[:, 2, 0] + S[:, 0, 2]
    K[:, 2, 2] = -S[:, 0, 0] + S[:, 1, 1] - S[:, 2, 2]
    K[:, 2, 3] = K[:, 3, 2] = S[:, 1, 2] + S[:, 2, 1]
    K[:, 3, 3] = -S[:, 0, 0] - S[:, 1, 1] + S[:, 2, 2]
    return np.linalg.eigvalsh(K)[:, -1]


def calibrate(n, k, A, seed):
    L = _lib.load()
    X = synth.trajectory(n, A, seed=seed)
    data = DeviceTrajectory.from_host(X)
    cidx = np.linspace(0, n - 1, k).astype(np.int64)
    cen = data.gather(cidx)
    scratch = torch.zeros(int(L.eb_tc_scratch_bytes(n, A, k)), dtype=torch.uint8, device="cuda")
    dbg = torch.zeros((n, k, 9), dtype=torch.float32, device="cuda")
    _lib.call("eb_rmsd_assign_tc", ptr(data.xyz), ptr(data.traces), n, A, ptr(cen.xyz),
              ptr(cen.traces), k, 0.0, None, 0, None, None, None, ptr(scratch), ptr(dbg), 0,
              stream_ptr())
    torch.cuda.synchronize()
    xs = data.xyz.cpu().numpy().astype(np.float64)
    cs = cen.xyz.cpu().numpy().astype(np.float64)
    M = np.einsum("fia,cja->fcij", xs, cs).reshape(n * k, 9)
    got = dbg.cpu().numpy().astype(np.float64).reshape(n * k, 9)
    G = np.sqrt(data.traces.cpu().numpy()[:, None] * cen.traces.cpu().numpy()[None, :]).reshape(-1)
    dl = 2 * np.abs(lam_max(got) - lam_max(M)) / G
    print("calib n=%d k=%d A=%d: max 2|dlambda|/sqrt(GaGb) = %.3e  p99.9 = %.3e  mean = %.3e"
          % (n, k, A, dl.max(), np.quantile(dl, 0.999), dl.mean()))


if __name__ == "__main__":
    calibrate(2000, 64, 500, 1)
    calibrate(2000, 64, 240, 2)
    calibrate(3000, 32, 1000, 3)
    calibrate(3000, 32, 48, 4)
