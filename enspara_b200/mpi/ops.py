"""Shard-level helpers with the reference's ``enspara.mpi.ops`` names and meaning
(/root/reference/enspara/mpi/ops.py), on torch.distributed instead of mpi4py.

The k-centers / PAM device loops do not use these (their exchange is the candidate-record
all-gather, see cluster/_engine.py); they exist for the callers either side of the hot path
(CLI result assembly, tests, user scripts written against the reference).
"""
import numpy as np
import torch
from sklearn.utils import check_random_state

from ..exception import DataInvalid, ImproperlyConfigured


def _mpi():
    from .. import mpi
    return mpi


def _lengths_to_offsets(lengths):
    lengths = np.asarray(lengths, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(lengths)])


def convert_local_indices(local_ctr_inds, global_lengths):
    """(rank, local_frame) -> global frame, assuming trajectory i lives on rank i % size
    (ops.py:14-39; the round-robin file striping of mpi/io.py:188-189)."""
    mpi = _mpi()
    size = mpi.size()
    global_lengths = np.asarray(global_lengths, dtype=np.int64)
    starts = _lengths_to_offsets(global_lengths)
    out = []
    for r, local_fid in local_ctr_inds:
        owned = np.arange(len(global_lengths))[int(r)::size]
        flat = np.concatenate([np.arange(starts[t], starts[t + 1]) for t in owned]) \
            if len(owned) else np.zeros(0, np.int64)
        out.append(flat[int(local_fid)])
    return out


def assemble_striped_array(local_arr):
    """Element i of the global array lives on rank i % size (ops.py:42-79)."""
    mpi = _mpi()
    local_arr = np.asarray(local_arr)
    if mpi.size() == 1:
        return local_arr
    if not np.all(local_arr > 0):
        raise ImproperlyConfigured(
            "On rank %s, a length <= 0 was found. Lengths must be strictly greater than zero."
            % mpi.rank())
    parts = mpi.comm.all_gather_object(local_arr)
    total = sum(len(p) for p in parts)
    out = np.zeros((total,) + local_arr.shape[1:], dtype=local_arr.dtype) - 1
    for r, p in enumerate(parts):
        out[r::mpi.size()] = p
    return out


def assemble_striped_ragged_array(local_array, global_lengths):
    """Concatenated per-frame array whose rows (trajectories) are striped over ranks
    (ops.py:82-125).  Returns the flat global array in trajectory order."""
    mpi = _mpi()
    global_lengths = np.asarray(global_lengths)
    if not np.issubdtype(global_lengths.dtype, np.integer):
        raise DataInvalid("global_lengths must be integers")
    local_array = np.asarray(local_array)
    size = mpi.size()
    parts = mpi.comm.all_gather_object(local_array)
    starts = _lengths_to_offsets(global_lengths)
    out = np.zeros(int(starts[-1]), dtype=local_array.dtype)
    for r, p in enumerate(parts):
        pos = 0
        for t in range(r, len(global_lengths), size):
            L = int(global_lengths[t])
            out[starts[t]:starts[t] + L] = p[pos:pos + L]
            pos += L
    return out


def striped_array_max(local_array):
    """Global max of an array spread over ranks (ops.py:128-140)."""
    mpi = _mpi()
    local_array = np.asarray(local_array)
    local_max = local_array.max() if local_array.size else -np.inf
    t = torch.tensor([float(local_max)], dtype=torch.float64)
    if mpi.size() > 1 and torch.distributed.get_backend() == "nccl":
        t = t.cuda()
    return float(mpi.comm.all_reduce_max(t).cpu()[0])


def striped_array_mean(local_array):
    """Global mean: sum of local sums / sum of local lengths (ops.py:143-166)."""
    mpi = _mpi()
    local_array = np.asarray(local_array)
    local_sum, local_len = np.sum(local_array), len(local_array)
    if mpi.size() == 1:
        return local_sum / local_len
    t = torch.tensor([float(local_sum), float(local_len)], dtype=torch.float64)
    if torch.distributed.get_backend() == "nccl":
        t = t.cuda()
    t = mpi.comm.all_reduce_sum(t).cpu()
    return float(t[0]) / float(t[1])


def distribute_frame(data, world_index, owner_rank):
    """Broadcast one element of ``data`` from its owner to every rank (ops.py:169-212)."""
    mpi = _mpi()
    if owner_rank >= mpi.size():
        raise ImproperlyConfigured(
            "In MPI swarm of size %s, recieved owner rank == %s." % (mpi.size(), owner_rank))
    is_traj = hasattr(data, "xyz")
    if mpi.rank() == owner_rank:
        frame = np.array(data[world_index].xyz if is_traj else data[world_index])
    else:
        frame = np.empty_like(data[0].xyz if is_traj else data[0])
    if mpi.size() > 1:
        t = torch.from_numpy(np.ascontiguousarray(frame))
        nccl = torch.distributed.get_backend() == "nccl"
        if nccl:
            t = t.cuda()
        mpi.comm.broadcast(t, owner_rank)
        frame = t.cpu().numpy()
    if is_traj:
        return type(data)(xyz=frame, topology=data.top)
    return frame


def randind(local_array, random_state=None):
    """Uniform choice over an array spread across ranks -> (owner_rank, local_index)
    (ops.py:215-272).  Rank 0 draws ``randint(total)``; the global position g is mapped through
    the striped concatenation [arange(total)[r::size] for r], cut by the per-rank lengths --
    kept verbatim in meaning because it defines the reference's random stream."""
    mpi = _mpi()
    random_state = check_random_state(random_state)
    n_states = np.array(mpi.comm.all_gather_object(len(local_array)))
    total = int(n_states.sum())
    if total < 1:
        raise DataInvalid(
            "Random choice requires a non-empty array. Got shapes: %s" % n_states)
    g = random_state.randint(total) if mpi.rank() == 0 else None
    if mpi.size() > 1:
        g = mpi.comm.all_gather_object(g)[0]
    concat = np.concatenate([np.arange(total)[r::mpi.size()] for r in range(mpi.size())])
    pos = int(np.where(concat == g)[0][0])
    bounds = _lengths_to_offsets(n_states)
    owner = int(np.searchsorted(bounds, pos, side="right") - 1)
    return owner, pos - int(bounds[owner])
