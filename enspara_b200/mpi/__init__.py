"""Distributed substrate: the reference's ``enspara.mpi`` surface on top of torch.distributed.

The reference talks to mpi4py's COMM_WORLD and falls back to a DummyComm for one process
(/root/reference/enspara/mpi/__init__.py:11-40, mpi/util.py:6-26).  Here one process drives one
GPU; ranks are torch.distributed ranks (NCCL over NVLink on the GPU box, gloo in CPU tests),
launched with torchrun.  ``rank()`` / ``size()`` keep their meaning; when no process group is
initialised they return 0 / 1 exactly like the reference's fallback.
"""
import os

import torch
import torch.distributed as dist

from . import ops  # noqa: F401  (re-exported like the reference)


def is_distributed():
    return dist.is_available() and dist.is_initialized()


def rank():
    return dist.get_rank() if is_distributed() else 0


def size():
    return dist.get_world_size() if is_distributed() else 1


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs nearest to its GPU (NVML's ideal affinity), so that pinned
    staging buffers are first-touched on the GPU's NUMA node: with one process per GPU all
    uploading at once, remote-socket host memory is what limits host -> HBM throughput.
    Best effort: returns the number of CPUs bound to, or 0 when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                phys = ids[device_index]
                h = (pynvml.nvmlDeviceGetHandleByIndex(int(phys)) if phys.isdigit()
                     else pynvml.nvmlDeviceGetHandleByUUID(phys))
            else:
                h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
            cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64)
                    if (int(mask) >> b) & 1}
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
            return len(cpus)
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return 0


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (idempotent)."""
    if is_distributed() or "RANK" not in os.environ:
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        if os.environ.get("ENSPARA_B200_NUMA_BIND", "1") != "0":
            bind_to_gpu_numa_node(local)
        dist.init_process_group(backend=backend, device_id=torch.device("cuda", local))
        return
    dist.init_process_group(backend=backend)


class Comm:
    """The handful of collectives the clustering path needs, on device or host tensors."""

    def __init__(self, group=None):
        self.group = group

    @property
    def size(self):
        return dist.get_world_size(self.group) if is_distributed() else 1

    @property
    def rank(self):
        return dist.get_rank(self.group) if is_distributed() else 0

    def barrier(self):
        if self.size > 1:
            dist.barrier(self.group)

    def all_gather_into(self, out, inp):
        """out: (size * numel) flat tensor, inp: flat tensor; same device/dtype."""
        if self.size == 1:
            if out.data_ptr() != inp.data_ptr():
                out.view(-1)[:inp.numel()].copy_(inp.view(-1))
            return
        dist.all_gather_into_tensor(out, inp, group=self.group)

    def all_reduce_sum(self, t):
        if self.size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max(self, t):
        if self.size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def broadcast(self, t, root):
        if self.size > 1:
            dist.broadcast(t, src=root, group=self.group)
        return t

    def broadcast_object(self, obj, root=0):
        """Pickled broadcast of a small host object (used once per PAM sweep for the
        RandomState, never per proposal)."""
        if self.size == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def all_gather_object(self, obj):
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out


comm = Comm()
