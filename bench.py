#!/usr/bin/env python
"""bench.py -- RMSD frame-centre evaluations per second for KCenters on B200 (BASELINE.json).

A "step" is ONE k-centers iteration (one pass of the fused step kernel over every resident
frame, plus the candidate all-gather when N > 1).  Workload = BASELINE config 4's per-GPU
shard: 1,250,000 frames x 500 atoms per GPU (7.5 GB of coordinates per GPU, >> 126 MB of L2, so
no flush is needed between steps); N GPUs hold N such shards (weak scaling; N = 8 is the
10M-frame north-star run).  Frames are synthetic (enspara_b200/synth.py, generated in HBM).

Arms
  default            the B200 path.  `value`: inputs resident in HBM, K steps timed with CUDA
                     events, max over ranks.  `e2e`: the same metric through the public API
                     (KCenters(...).fit on a HOST array: H2D + centring + K steps + D2H of the
                     results inside the timed region).  `roofline`: per-launch event times of
                     the step kernel against MEASURED_PEAKS.json.  `cpu_baseline`: the oracle's
                     restated mdtraj path on the host cores, bounded sample (rank 0, N=1 only).
  --impl reference   the reference's own CPU path for this metric (restated: mdtraj is not
                     installable here) driven by the restated k-centers loop, all host threads.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FRAMES_PER_GPU = 1_250_000
N_ATOMS = 500
METRIC = "rmsd_frame_center_evals_per_s"
UNIT = "evals/s"
WORKLOAD = "KCenters rmsd, 1.25M frames x 500 atoms per GPU (BASELINE config 4 shard: " \
           "10M x 500 over 8 GPUs), one k-centers iteration per step"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--frames-per-gpu", type=int, default=N_FRAMES_PER_GPU)
    p.add_argument("--atoms", type=int, default=N_ATOMS)
    p.add_argument("--cpu-seconds", type=float, default=12.0,
                   help="target CPU seconds for the bounded cpu_baseline sample")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-parity-check", action="store_true",
                   help="skip the untimed N-rank == 1-rank check (N > 1)")
    p.add_argument("--fast", action="store_true", help="float32-block accumulation mode")
    return p.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to all samples
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's path for this metric, restated (oracle/), bounded sample
# ---------------------------------------------------------------------------------------------
def _host_sample(n, n_atoms):
    """The first n frames of the synthetic trajectory as a host array, made by the oracle's C
    restatement of the generator (bit-identical to enspara_b200/synth.py and to the CUDA
    generator; OpenMP, about a second).  The CPU arms therefore neither load the product's
    CUDA library nor touch the GPU; generation is outside every timed region."""
    from oracle import distances as od
    return od.synth_trajectory(n, n_atoms, seed=0, first_frame=0)


def cpu_kcenters_rate(n_atoms, steps, warmup, target_seconds, sample_frames=None):
    """Times the restated reference k-centers iteration (kcenters.py:282-309 around the
    restated md.rmsd with its per-call copy + centre) on all host threads.  Returns
    (evals_per_s, info dict)."""
    from oracle import distances as od
    # every core this process may use, whatever OMP_NUM_THREADS says (torchrun exports 1)
    threads = od.use_all_cores()
    n = sample_frames or 250_000
    X = _host_sample(n, n_atoms)
    T = od.Trajectory(X)
    distances = np.full(n, np.inf)
    assignments = np.full(n, -1, dtype=np.int64)
    ctr = []

    def iteration():
        new = int(np.argmax(distances))
        d = od.rmsd_f32_sse(T, T[new])
        upd = d < distances
        distances[upd] = d[upd]
        assignments[upd] = len(ctr)
        ctr.append(new)
        return distances.max()

    for _ in range(max(1, warmup)):
        iteration()
    if steps is None:
        t = time.perf_counter()
        iteration()
        one = time.perf_counter() - t
        steps = int(min(200, max(3, target_seconds / max(one, 1e-6))))
    t = time.perf_counter()
    for _ in range(steps):
        iteration()
    dt = time.perf_counter() - t
    rate = n * steps / dt
    info = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d k-centers iterations over %d frames x %d atoms (restated mdtraj "
                      "float32 RMSD: SSE 4-atom lanes + OpenMP over frames, incl. mdtraj's "
                      "per-call copy+centre; numpy bookkeeping as kcenters.py:282-309; mdtraj "
                      "itself is not installable offline)" % (steps, n, n_atoms),
            "seconds": dt}
    return rate, info, steps, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, info, steps, dt = cpu_kcenters_rate(args.atoms, args.steps, args.warmup,
                                              args.cpu_seconds)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_atoms": args.atoms,
                   "note": "CPU arm runs a bounded sample of the same workload: "
                           + info["sample"]},
        "cpu_baseline": info,
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# N ranks == 1 rank (the reference's own MPI bar: test_cluster.py:270-275, 309-314)
# ---------------------------------------------------------------------------------------------
def _checksum(center_list, assign, dist, first_index):
    """Order-sensitive 63-bit checksum of (global centre list, assignments, float32 distance
    bits) that is ADDITIVE over shards, so the sharded run's value is an all-reduce of the
    ranks' values and must equal the single-GPU value of the same global data."""
    g = np.arange(first_index, first_index + len(assign), dtype=np.int64)
    bits = np.asarray(dist, dtype=np.float32).view(np.int32).astype(np.int64)
    s = int(((np.asarray(assign, np.int64) + 2) * (g % 65521 + 1)).sum())
    s += int(((bits % 65521) * (g % 65519 + 1)).sum())
    c = int(sum((int(v) % 1000003) * (i + 1) for i, v in enumerate(center_list)))
    return s, c


def sharded_parity_check(world, rank, A, per_rank=80_000, n_centers=12):
    """Untimed.  A down-scaled copy of the bench's global trajectory (per_rank frames per rank,
    still large enough for the TMA step kernel + the exchange that the timed steps use) is
    clustered twice through the public function: sharded over the N ranks, and in one piece on
    every rank's own GPU (the generator is counter-based, so any rank can make all of it).
    Centres, assignments and distances must be IDENTICAL."""
    import torch
    import torch.distributed as dist
    from enspara_b200 import _lib, synth
    from enspara_b200.cluster import kcenters as kc
    full = synth.device_trajectory(per_rank * world, A, seed=0, first_frame=0)
    mine = synth.device_trajectory(per_rank, A, seed=0, first_frame=rank * per_rank)
    one = kc.kcenters(full, "rmsd", n_clusters=n_centers, mpi_mode=False)
    shd, eng = kc.kcenters(mine, "rmsd", n_clusters=n_centers, mpi_mode=True,
                           _return_engine=True)
    lo, hi = rank * per_rank, (rank + 1) * per_rank
    glob = [int(r * per_rank + l) for r, l in shd.center_indices]
    ok_c = glob == [int(c) for c in one.center_indices]
    ok_a = bool(np.array_equal(shd.assignments, one.assignments[lo:hi]))
    ok_d = bool(np.array_equal(shd.distances, one.distances[lo:hi]))
    s_sh, c_sh = _checksum(glob, shd.assignments, shd.distances, lo)
    s_one, c_one = _checksum(one.center_indices, one.assignments, one.distances, 0)
    t = torch.tensor([s_sh, int(ok_c), int(ok_a), int(ok_d)], dtype=torch.int64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    s_sh_all, n_c, n_a, n_d = (int(v) for v in t.cpu())
    return {"what": "KCenters rmsd n_clusters=%d on a %d-frame x %d-atom down-scaled copy of "
                    "the bench trajectory: %d-rank sharded run vs single-GPU run"
                    % (n_centers, per_rank * world, A, world),
            "frames_total": per_rank * world, "frames_per_rank": per_rank,
            "step_kernel_is_tma": bool(_lib.load().eb_kcenters_step_rmsd_uses_tma(per_rank, A)),
            "exchange_is_fused_p2p": bool(eng.p2p),
            "centers_equal": n_c == world, "assignments_equal": n_a == world,
            "distances_equal": n_d == world,
            "checksum_sharded": [s_sh_all, c_sh], "checksum_single_gpu": [s_one, c_one],
            "equal": bool(n_c == world and n_a == world and n_d == world
                          and s_sh_all == s_one and c_sh == c_one)}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from enspara_b200 import mpi, synth
    from enspara_b200.cluster import kcenters as kc_mod
    from enspara_b200.cluster._engine import KCentersEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one JSON line: anything libraries print while the run is in
    # progress (NCCL's version banner is written to fd 1 at communicator creation) is sent to
    # stderr; the saved descriptor is restored just before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    numa_cpus = 0
    if world > 1:
        # one process per GPU, all uploading at once in the e2e leg: keep each process (and
        # the pinned buffers it first-touches) on its GPU's NUMA node
        numa_cpus = mpi.bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    comm = mpi.comm if world > 1 else kc_mod._SingleComm()

    n_local, A = args.frames_per_gpu, args.atoms
    n_total = n_local * world
    K, W = args.steps, max(args.warmup, 3)
    exact = not args.fast

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident data: this rank's shard of the global synthetic trajectory ---------------
    data = synth.device_trajectory(n_local, A, seed=0, first_frame=rank * n_local)
    torch.cuda.synchronize()

    # ---- `value`: K steps, inputs resident in HBM ------------------------------------------
    eng = KCentersEngine(data, "rmsd", comm, exact=exact)
    eng._ensure_center_list(K + W + 8)
    eng.seed(0)
    limit = K + W + 4
    for _ in range(W):
        eng.step(limit, 0.0)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    wait_ns_before = int(eng.read_state().wait_ns)   # warm-up incl. the ranks' start-up skew
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(K):
        eng.step(limit, 0.0)
        ev[i + 1].record()
    barrier()
    t1 = time.perf_counter()
    total_ms = ev[0].elapsed_time(ev[K])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    st = eng.read_state()
    assert st.n_centers == K + W, "steps were skipped: %d != %d" % (st.n_centers, K + W)
    exchange = ("none (single GPU)" if world == 1 else
                "peer-memory stores + flags fused into the step kernel (no collective launch)"
                if eng.p2p else "one NCCL all-gather of candidate records per step")
    clk = clocks.stop(t0, t1) if rank == 0 else None

    tmax = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.cpu()[0])
    value = n_total * K / (total_ms * 1e-3)

    exch_wait_us = None
    if world > 1 and eng.p2p:
        # time block 0 spent in the prologue waiting for the peers' records, per TIMED step
        # (the spread between the GPUs: every iteration waits for the slowest shard)
        exch_wait_us = 1e-3 * float(int(st.wait_ns) - wait_ns_before) / max(1, K)

    parity = None
    if world > 1 and not args.no_parity_check:
        parity = sharded_parity_check(world, rank, A)

    # roofline of the dominant kernel (the fused step): algorithmic bytes per launch
    # = (12*A + 8) bytes per frame x frames per launch, over the mean per-launch
    # event time on the launching stream
    peak, peak_src = measured_peaks()
    bytes_per_launch = (12 * A + 8) * n_local
    mean_launch_s = float(np.mean(per_launch_ms)) * 1e-3
    achieved = bytes_per_launch / mean_launch_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel": ("k_kcenters_step_rmsd_tma" if exact and
                           eng.lib.eb_kcenters_step_rmsd_uses_tma(n_local, A) else
                           "k_kcenters_step_rmsd<exact=%d>" % int(exact)),
                "peak_note": "peak is the measured device COPY bandwidth (reads + writes); this "
                             "kernel only reads, and a read-only stream can exceed it (ncu: "
                             "84.9 % of the 8.18 TB/s DRAM peak)",
                "bytes_per_launch_algorithmic": bytes_per_launch,
                "mean_launch_ms": mean_launch_s * 1e3}
    ncu_traffic = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
    if os.path.exists(ncu_traffic):
        try:
            with open(ncu_traffic) as fh:
                tr = json.load(fh)
            if tr.get("n_frames") == n_local and tr.get("n_atoms") == A:
                roofline["traffic"] = tr.get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- `e2e`: public API on a HOST array ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        from enspara_b200.cluster import KCenters
        del eng
        host = torch.empty((n_local, A, 3), dtype=torch.float32).pin_memory()
        base = synth.base_conformers(A, 64, 0)
        chunk = 125_000
        for lo in range(0, n_local, chunk):
            m = min(chunk, n_local - lo)
            host[lo:lo + m].copy_(synth.device_trajectory_aos(
                m, A, 0, rank * n_local + lo, base=base))
        torch.cuda.synchronize()
        del data
        torch.cuda.empty_cache()
        host_np = host.numpy()
        K2 = K
        est = KCenters("rmsd", n_clusters=K2, mpi_mode=(world > 1))
        # warm-up of the API path on a small slice (allocator, lazy init), untimed
        KCenters("rmsd", n_clusters=3, mpi_mode=(world > 1)).fit(host_np[:20000])
        barrier()
        t = time.perf_counter()
        est.fit(host_np)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        tm = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dt = float(tm.cpu()[0])
        assert len(est.result_.center_indices) == K2
        # where the time goes (second, UNTIMED fit with a device sync after every phase)
        os.environ["ENSPARA_B200_PHASE_TIMES"] = "1"
        KCenters("rmsd", n_clusters=K2, mpi_mode=(world > 1)).fit(host_np)
        os.environ["ENSPARA_B200_PHASE_TIMES"] = "0"
        phases = {k: round(v, 4) for k, v in kc_mod.last_phase_times.items()}
        e2e = {"value": n_total * K2 / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(n_local * A * 12 / K2),
               "d2h_bytes_per_step": int(n_local * (4 + 4) / K2),
               "seconds": dt, "steps": K2, "phases_s_rank0_untimed_rerun": phases,
               "numa_bound_cpus": numa_cpus,
               "what": "KCenters('rmsd', n_clusters=%d).fit(pinned host (n,%d,3) float32): "
                       "H2D + centring + %d iterations + D2H of assignments/distances, per "
                       "rank on its shard" % (K2, A, K2)}

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            _, cpu, _, _ = cpu_kcenters_rate(A, None, 1, args.cpu_seconds)
        except Exception as exc:  # the oracle is optional for the GPU arm
            cpu = {"error": repr(exc)}

    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 accumulate / f32 in-out" if exact else "f32 blocks + f64 sums",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": n_local, "n_atoms": A,
                       "frames_total": n_total, "l2": "inputs (%.1f GB per GPU) larger than L2; "
                       "no flush needed" % (n_local * A * 12 / 1e9),
                       "parallelism": "frames sharded contiguously; candidate exchange: "
                                      + exchange if world > 1 else "single GPU"},
            "roofline": roofline, "clocks": clk, "gpu_launches": K,
            "e2e": e2e, "cpu_baseline": cpu,
        }
        if parity is not None:
            line["parity_check"] = parity
        if exch_wait_us is not None:
            line["exchange_wait_us_per_step"] = exch_wait_us
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["equal"]:
        sys.exit("bench.py: the %d-rank run differs from the single-GPU run" % world)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
