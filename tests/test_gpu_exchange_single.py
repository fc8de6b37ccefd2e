"""GPU (ONE device is enough): the fused peer-memory candidate exchange of the RMSD step
(csrc/eb_common.cuh `struct Exch`, entry points eb_kcenters_{seed,step}_rmsd_p2p), driven
directly through the C ABI with R "ranks" emulated inside one process -- one exchange buffer,
state block and stream per rank, `peers` = the R buffer addresses on the same GPU.  Exercises
the protocol itself (sequence parity, flags, record slots, stop rule, empty shards, timeout)
where only one GPU is available; tests/test_gpu_multi.py does the same across real GPUs.

What must hold (the reference's MPI bar, enspara/test/test_cluster.py:270-275, 309-314):
centres, assignments and distances of the R-shard run == the one-shard run, bit for bit.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


class _Rank:
    def __init__(self, torch, lib, data, lo, hi, rank, size, nbytes, rec_bytes, cap):
        from enspara_b200.device import DeviceTrajectory
        dev = data.xyz.device
        self.lo, self.n = lo, hi - lo
        self.data = DeviceTrajectory(data.xyz[lo:hi], data.traces[lo:hi], data.n_atoms)
        self.rank, self.size = rank, size
        self.exch = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.dist = torch.full((self.n,), float("inf"), dtype=torch.float32, device=dev)
        self.assign = torch.full((self.n,), -1, dtype=torch.int32, device=dev)
        self.state = torch.zeros(64, dtype=torch.uint8, device=dev)
        self.partials = torch.empty(int(lib.eb_kc_partials_bytes()), dtype=torch.uint8,
                                    device=dev)
        self.cand_out = torch.zeros(rec_bytes, dtype=torch.uint8, device=dev)
        self.centers = torch.full((cap,), -1, dtype=torch.int64, device=dev)
        self.stream = torch.cuda.Stream(device=dev)


def _run_emulated(torch, data, bounds, n_clusters, cutoff, lockstep):
    """k-centers over len(bounds)-1 emulated ranks.  ``lockstep``: a rank's step s+1 is queued
    behind every rank's step s with events (needed when one launch fills the GPU, e.g. the TMA
    kernel at one CTA per SM: a spinning consumer would otherwise starve its producer);
    otherwise the ranks' launches really overlap and wait for each other's flags."""
    from enspara_b200 import _lib
    from enspara_b200.device import ptr
    lib = _lib.load()
    size = len(bounds) - 1
    A = data.n_atoms
    nbytes = int(lib.eb_exch_bytes(A, size))
    rec = int(lib.eb_rmsd_record_bytes(A))
    cap = n_clusters + 2
    ranks = [_Rank(torch, lib, data, int(bounds[r]), int(bounds[r + 1]), r, size, nbytes, rec,
                   cap) for r in range(size)]
    peers = torch.tensor([r.exch.data_ptr() for r in ranks], dtype=torch.int64,
                         device=data.xyz.device)
    torch.cuda.synchronize()

    def sp(r):
        return ctypes.c_void_p(r.stream.cuda_stream)

    def wait_all(evs):
        for r in ranks:
            for e in evs:
                r.stream.wait_event(e)

    evs = []
    for r in ranks:
        d = r.data
        _lib.call("eb_kcenters_seed_rmsd_p2p", ptr(d.xyz), ptr(d.traces), r.n, A, r.lo,
                  ptr(peers), size, r.rank, ptr(r.dist), 0, ptr(r.state), ptr(r.partials),
                  ptr(r.cand_out), sp(r))
        e = torch.cuda.Event()
        e.record(r.stream)
        evs.append(e)
    for _ in range(n_clusters + 1):        # one launch more than needed: must be a no-op
        if lockstep:
            wait_all(evs)
        evs = []
        for r in ranks:
            d = r.data
            _lib.call("eb_kcenters_step_rmsd_p2p", ptr(d.xyz), ptr(d.traces), r.n, A, r.lo,
                      ptr(peers), size, r.rank, ptr(r.dist), ptr(r.assign), n_clusters,
                      float(cutoff), ptr(r.state), ptr(r.centers), ptr(r.partials),
                      ptr(r.cand_out), 1, 1, sp(r))
            e = torch.cuda.Event()
            e.record(r.stream)
            evs.append(e)
    torch.cuda.synchronize()
    states = [_lib.KcState.from_buffer_copy(r.state.cpu().numpy().tobytes()) for r in ranks]
    return ranks, states


def _check(torch, n, A, bounds, n_clusters, cutoff, lockstep, seed=3):
    from enspara_b200 import synth
    from enspara_b200.cluster import kcenters
    data = synth.device_trajectory(n, A, seed=seed)
    kw = dict(n_clusters=n_clusters)
    if cutoff > 0:
        kw["dist_cutoff"] = cutoff
    serial = kcenters.kcenters(data, "rmsd", **kw)
    ranks, states = _run_emulated(torch, data, bounds, n_clusters, cutoff, lockstep)
    k = len(serial.center_indices)
    for r, st in zip(ranks, states):
        assert st.error == 0
        assert st.n_centers == k
        got = r.centers[:k].cpu().numpy().tolist()
        assert got == [int(c) for c in serial.center_indices], "rank %d" % r.rank
        np.testing.assert_array_equal(r.assign.cpu().numpy().astype(np.int64),
                                      serial.assignments[r.lo:r.lo + r.n])
        np.testing.assert_array_equal(r.dist.cpu().numpy().astype(np.float64),
                                      serial.distances[r.lo:r.lo + r.n])
    return serial, states


def test_overlapping_ranks_small_shards(cuda):
    """2, 3 and 8 emulated ranks whose launches genuinely overlap on the GPU (small grids):
    every consumer spins on its flags until the producers publish."""
    for size, n, A in ((2, 3001, 50), (3, 2500, 264), (8, 4000, 22)):
        bounds = np.linspace(0, n, size + 1).astype(int)
        _check(cuda, n, A, bounds, n_clusters=20, cutoff=0.0, lockstep=False)


def test_empty_shards_and_stop_rule(cuda):
    """Shards of zero frames publish 'empty' records (index -1); a cutoff ends the run on every
    rank at the same step and the queued extra launch is a no-op that waits for nothing."""
    n, A = 1200, 40
    bounds = np.array([0, 0, 700, 700, 1200])
    serial, states = _check(cuda, n, A, bounds, n_clusters=30, cutoff=0.0, lockstep=False)
    cut = float(serial.distances.max()) * 1.1
    serial_c, states = _check(cuda, n, A, bounds, n_clusters=30, cutoff=cut, lockstep=False)
    assert 1 < len(serial_c.center_indices) < 30
    assert all(st.done == 1 for st in states)


def test_tma_step_kernel_with_exchange(cuda):
    """Config-4-shaped shards (500 atoms, >= 75 776 frames per rank): the step kernel is
    k_kcenters_step_rmsd_tma, one CTA per SM -- the combination bench.py times at N > 1."""
    from enspara_b200 import _lib
    per = 80_000
    assert _lib.load().eb_kcenters_step_rmsd_uses_tma(per, 500)
    bounds = np.array([0, per, 2 * per])
    _check(cuda, 2 * per, 500, bounds, n_clusters=10, cutoff=0.0, lockstep=True, seed=0)


def test_missing_peer_times_out_instead_of_hanging(cuda, monkeypatch):
    """A rank whose peer never publishes gives up after EB_EXCH_TIMEOUT_S and flags the state
    (the host raises) -- it does not spin forever inside a kernel."""
    import os
    import subprocess
    import sys
    code = r'''
import ctypes, sys, numpy as np, torch
sys.path.insert(0, %r)
from enspara_b200 import _lib, synth
from enspara_b200.device import ptr
lib = _lib.load()
data = synth.device_trajectory(500, 30, seed=1)
nb = int(lib.eb_exch_bytes(30, 2)); rec = int(lib.eb_rmsd_record_bytes(30))
dev = data.xyz.device
bufs = [torch.zeros(nb, dtype=torch.uint8, device=dev) for _ in range(2)]
peers = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
dist = torch.full((500,), float("inf"), dtype=torch.float32, device=dev)
assign = torch.full((500,), -1, dtype=torch.int32, device=dev)
state = torch.zeros(64, dtype=torch.uint8, device=dev)
part = torch.empty(int(lib.eb_kc_partials_bytes()), dtype=torch.uint8, device=dev)
cand = torch.zeros(rec, dtype=torch.uint8, device=dev)
ctr = torch.full((8,), -1, dtype=torch.int64, device=dev)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.call("eb_kcenters_seed_rmsd_p2p", ptr(data.xyz), ptr(data.traces), 500, 30, 0, ptr(peers),
          2, 0, ptr(dist), 0, ptr(state), ptr(part), ptr(cand), s)
_lib.call("eb_kcenters_step_rmsd_p2p", ptr(data.xyz), ptr(data.traces), 500, 30, 0, ptr(peers),
          2, 0, ptr(dist), ptr(assign), 5, 0.0, ptr(state), ptr(ctr), ptr(part), ptr(cand), 1, 2,
          s)
torch.cuda.synchronize()
st = _lib.KcState.from_buffer_copy(state.cpu().numpy().tobytes())
print("ERROR=%%d DONE=%%d K=%%d" %% (st.error, st.done, st.n_centers))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, EB_EXCH_TIMEOUT_S="0.5")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                         timeout=300, env=env)
    assert "ERROR=1 DONE=1 K=0" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
