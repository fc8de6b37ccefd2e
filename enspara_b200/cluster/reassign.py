"""Streaming nearest-centre re-assignment of on-disk trajectories.

Reference: /root/reference/enspara/cluster/util.py:551-734 (`compute_batches`,
`determine_batch_size`, `batch_reassign`, `reassign`) and apps/reassign.py.  Same batching rule
and return types; what changes is where the work runs:

* centres are centred and uploaded ONCE (the reference pre-centres them once too, :684-686) and
  stay in HBM, already packed for the tensor-core screen when there are enough of them;
* each batch is read from disk by a background thread into pinned host memory while the GPU
  works on the previous batch (the reference loads, then computes, then loads ...);
* a batch goes host -> HBM in double-buffered chunks overlapped with the centring kernel
  (`DeviceTrajectory.from_host`), then through the many-centres RMSD kernels (K3t screen +
  exact re-score, or K3x), i.e. `assign_to_nearest_center` semantics with the reference's
  ``precentered=True`` arithmetic (:625-629) -- centring once is what `precentered` means.

File formats: ``.npy`` arrays of shape (n, A, 3) and mdtraj ``.h5`` trajectories
(`/coordinates`) are read without mdtraj; every other format goes through ``mdtraj.load`` when
mdtraj is importable (it is not in this image) and raises ImproperlyConfigured otherwise.
"""
import logging
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .. import ra
from ..exception import DataInvalid, ImproperlyConfigured
from ..ra import partition_list
from . import util

logger = logging.getLogger(__name__)


def compute_batches(lengths, batch_size):
    """Batches (lists of indices into ``lengths``) of combined length below ``batch_size``
    (util.py:551-567, same greedy rule including its strict '<')."""
    batch_sizes = [[]]
    batch_indices = [[]]
    for i, l in enumerate(lengths):
        if sum(batch_sizes[-1]) + l < batch_size:
            batch_sizes[-1].append(l)
            batch_indices[-1].append(i)
        else:
            batch_sizes.append([l])
            batch_indices.append([i])
    return batch_indices


def nonempty_batches(lengths, batch_size):
    """``compute_batches`` without the empty batch it opens with when the first file alone
    fills one (``lengths[0] >= batch_size``, e.g. ``[100, 10, 20]`` at 100 -> ``[[], [0],
    [1, 2]]``): the streaming pass chains a prefetch from every batch to the next."""
    return [b for b in compute_batches(lengths, batch_size) if b]


#: beyond this many bytes per batch streaming gains nothing (the loader, the H2D copy and the
#: kernels already overlap); smaller batches keep the two pinned staging buffers modest
BATCH_BYTES_CAP = 8 << 30


def determine_batch_size(n_atoms, dtype_bytes, frac_mem, device_frac=0.25, largest_file=0):
    """Frames per batch.  The reference's rule is ``frac_mem`` of host RAM (util.py:570-581);
    here a batch exists twice on the host (two pinned staging buffers, one being filled while
    the other is consumed) and once on the device (SoA frames + screen scratch), so it is
    bounded by ``frac_mem`` x RAM / 2, by ``device_frac`` of free HBM and by BATCH_BYTES_CAP --
    but never below the largest single file when that still fits the RAM / HBM bounds."""
    import psutil
    bytes_per_frame = n_atoms * 3 * dtype_bytes
    hard = int(psutil.virtual_memory().total * frac_mem / 2 / bytes_per_frame)
    try:
        import torch
        if torch.cuda.is_available():
            free, _ = torch.cuda.mem_get_info()
            # device bytes per frame: padded SoA copy + AoS staging share + packed screen
            # operands for the chunk in flight, ~3x the host bytes in the worst case
            hard = min(hard, int(free * device_frac / (3 * bytes_per_frame)))
    except Exception:  # pragma: no cover
        pass
    soft = max(int(BATCH_BYTES_CAP // bytes_per_frame), min(int(largest_file), hard))
    batch_size = min(hard, soft)
    return batch_size, batch_size * bytes_per_frame / 1024 ** 3


# ---------------------------------------------------------------------------------------------
# file access
# ---------------------------------------------------------------------------------------------
def _ext(path):
    return os.path.splitext(path)[1].lower()


def _mdtraj():
    try:
        import mdtraj as md
    except ImportError:
        return None
    # a real installation has a version and a file; ignore placeholder modules
    return md if getattr(md, "__file__", None) and hasattr(md, "load") else None


def sound_trajectory(path, stride=1):
    """Number of frames in a trajectory file without loading it (util/load.py:14-49)."""
    e = _ext(path)
    if e == ".npy":
        n = np.load(path, mmap_mode="r").shape[0]
    elif e in (".h5", ".hdf5"):
        from ..util import h5min
        n = h5min.File(path)["coordinates"].shape[0]
    elif e == ".xtc":
        from ..util import traj
        n = traj.xtc_shape(path)[0]        # native: frame headers only
    else:
        md = _mdtraj()
        if md is None:
            raise ImproperlyConfigured(
                "Reading '%s' needs mdtraj, which is not installed; .xtc, .npy and mdtraj .h5 "
                "trajectories are read natively." % path)
        with md.open(path) as f:
            n = len(f)
    return len(range(0, n, stride))


def load_frames(path, top=None, atom_indices=None, out=None):
    """(n, A, 3) float32 coordinates of one file, optionally into ``out``."""
    e = _ext(path)
    if e == ".xtc":
        # native decoder: the atom selection is applied while decoding, straight into ``out``
        from ..util import traj
        if out is not None and out.flags.c_contiguous and out.dtype == np.float32:
            return traj.read_xtc(path, atom_indices=atom_indices, out=out)
        xyz = traj.read_xtc(path, atom_indices=atom_indices)
        atom_indices = None
    elif e == ".npy":
        xyz = np.load(path, mmap_mode="r")
    elif e in (".h5", ".hdf5"):
        from ..util import h5min
        xyz = h5min.read(path, "coordinates")
    else:
        md = _mdtraj()
        if md is None:
            raise ImproperlyConfigured(
                "Reading '%s' needs mdtraj, which is not installed." % path)
        xyz = md.load(path, top=top, atom_indices=atom_indices).xyz
        atom_indices = None
    if xyz.ndim != 3 or xyz.shape[2] != 3:
        raise DataInvalid("'%s' holds an array of shape %s, expected (n_frames, n_atoms, 3)"
                          % (path, xyz.shape))
    if atom_indices is not None:
        xyz = xyz[:, np.asarray(atom_indices, dtype=np.int64)]
    if out is not None:
        out[...] = xyz
        return out
    return np.ascontiguousarray(xyz, dtype=np.float32)


def _select_atoms(topfile, selection):
    """(topology object or None, atom indices or None) for one topology / selection pair."""
    if selection is None:
        return None, None
    if not isinstance(selection, str):
        return None, np.asarray(selection, dtype=np.int64)
    md = _mdtraj()
    if md is not None:
        top = md.load(topfile).top
        return top, top.select(selection)
    if selection.strip() == "all" and (topfile is None or not os.path.exists(str(topfile))
                                       or _ext(str(topfile)) not in (".pdb", ".ent")):
        return None, None
    # no mdtraj: .pdb topologies and the selection mini-language of util/traj.py
    from ..util import traj
    top = traj.load_topology(topfile)
    try:
        return top, top.select(selection)
    except ValueError as exc:
        raise ImproperlyConfigured(
            "The provided selection '%s' didn't match the topology file, %s (%s)"
            % (selection, topfile, exc))


def _centers_xyz(centers, n_atoms=None):
    """Centres (md.Trajectory, list of 1-frame trajectories / arrays, ndarray) -> (k, A, 3)."""
    if hasattr(centers, "xyz"):
        xyz = np.asarray(centers.xyz, dtype=np.float32)
    elif isinstance(centers, np.ndarray):
        xyz = centers.astype(np.float32, copy=False)
    else:
        rows = []
        for c in centers:
            a = np.asarray(c.xyz if hasattr(c, "xyz") else c, dtype=np.float32)
            rows.append(a[0] if a.ndim == 3 else a)
        xyz = np.stack(rows)
    if xyz.ndim != 3 or xyz.shape[2] != 3:
        raise DataInvalid("centres have shape %s, expected (k, n_atoms, 3)" % (xyz.shape,))
    return xyz


# ---------------------------------------------------------------------------------------------
# the streaming pass
# ---------------------------------------------------------------------------------------------
def batch_reassign(targets, centers, lengths, frac_mem, n_procs=None, stats=None):
    """Assign every frame of every target file to its nearest centre, batch by batch
    (util.py:584-649).  ``targets``: list of (path, topology, atom_indices).  Returns
    (list of int64 arrays, list of float64 arrays), one pair per file."""
    import torch
    from ..device import DeviceTrajectory
    from . import _ops

    cxyz = _centers_xyz(centers)
    n_atoms = cxyz.shape[1]
    DTYPE_BYTES = 4
    batch_size, batch_gb = determine_batch_size(
        n_atoms, DTYPE_BYTES, frac_mem, largest_file=max(list(lengths) + [0]))
    logger.info("Batch max size set to %s frames (~%.2f GB, %.1f%% of total RAM).",
                batch_size, batch_gb, frac_mem * 100)
    if len(lengths) and batch_size < max(lengths):
        raise ImproperlyConfigured(
            "Batch size of %s was smaller than largest file (size %s)."
            % (batch_size, max(lengths)))
    batches = nonempty_batches(lengths, batch_size)
    metric = util.RMSD
    cdev = DeviceTrajectory.from_host(cxyz)          # centred once, resident for all batches

    biggest = max([sum(lengths[i] for i in b) for b in batches] + [1])
    pinned = [torch.empty((biggest, n_atoms, 3), dtype=torch.float32).pin_memory()
              for _ in range(2)]

    if n_procs is None:
        n_procs = max(1, min(8, os.cpu_count() or 1))
    file_pool = ThreadPoolExecutor(max_workers=n_procs)

    def load_batch(bi):
        """Disk -> pinned host buffer bi % 2: one task per file on the file pool (the copies
        release the GIL), coordinated from the prefetch thread."""
        buf = pinned[bi % 2].numpy()
        lo, tasks = 0, []
        for i in batches[bi]:
            path, top, aids = targets[i]
            n = lengths[i]
            tasks.append(file_pool.submit(load_frames, path, top, aids, buf[lo:lo + n]))
            lo += n
        for t in tasks:
            t.result()
        return lo

    assignments, distances = [], []
    t_load = t_gpu = 0.0
    pool = ThreadPoolExecutor(max_workers=1)
    try:
        pending = pool.submit(load_batch, 0) if batches else None
        for bi, batch in enumerate(batches):
            tick = time.perf_counter()
            n_b = pending.result()
            t_load += time.perf_counter() - tick
            # the other pinned buffer is free: the previous batch's GPU work was synchronised
            # by its D2H copy below
            if bi + 1 < len(batches):
                pending = pool.submit(load_batch, bi + 1)
            tick = time.perf_counter()
            data = DeviceTrajectory.from_host(pinned[bi % 2].numpy()[:n_b])
            d, a = _ops.assign_device_auto(metric, data, cdev)
            a_host = a.cpu().numpy().astype(np.int64)
            d_host = d.cpu().numpy().astype(np.float64)
            del data, d, a
            t_gpu += time.perf_counter() - tick
            blens = [lengths[i] for i in batch]
            assignments.extend(partition_list(a_host, blens))
            distances.extend(partition_list(d_host, blens))
            logger.info("Finished batch %s of %s (%s frames).", bi + 1, len(batches), n_b)
    finally:
        pool.shutdown(wait=True)
        file_pool.shutdown(wait=True)
    if stats is not None:
        stats.update(wait_for_loader_s=t_load, gpu_s=t_gpu, batches=len(batches),
                     batch_size=batch_size)
    return assignments, distances


def reassign(topologies, trajectories, atoms, centers, frac_mem=0.5):
    """Re-assign sets of trajectory files to ``centers`` (util.py:652-734).  Returns
    (assignments, distances): 2-D ndarrays when all files have the same length, RaggedArrays
    otherwise."""
    if len(topologies) != len(trajectories):
        raise ImproperlyConfigured(
            "Number of topologies (%s) didn't match number of sets of trajectories (%s)."
            % (len(topologies), len(trajectories)))
    if len(topologies) != len(atoms):
        raise ImproperlyConfigured(
            "Number of topologies (%s) didn't match number of atom selection strings (%s)."
            % (len(topologies), len(atoms)))

    tick = time.perf_counter()
    targets = []
    for topfile, trjfiles, sel in zip(topologies, trajectories, atoms):
        top, atom_ids = _select_atoms(topfile, sel)
        for trjfile in trjfiles:
            if not os.path.exists(trjfile):
                raise ImproperlyConfigured("Trajectory file '%s' does not exist." % trjfile)
            targets.append((trjfile, top, atom_ids))
    logger.info("Sounding dataset of %s trajectories and %s topologies.",
                sum(len(t) for t in trajectories), len(topologies))
    lengths = [sound_trajectory(f) for f, _, _ in targets]
    logger.info("Sounded %s trajectories with %s frames in %.1f seconds.", len(lengths),
                sum(lengths), time.perf_counter() - tick)

    assignments, distances = batch_reassign(targets, centers, lengths, frac_mem=frac_mem)
    logger.info("Reassignment took %.1f seconds.", time.perf_counter() - tick)

    if all(len(assignments[0]) == len(a) for a in assignments):
        return np.array(assignments), np.array(distances)
    return ra.RaggedArray(assignments), ra.RaggedArray(distances)
