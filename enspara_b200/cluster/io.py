"""Loading inputs and writing outputs for the `cluster` app -- the callers either side of the
hot path.  Behaviour follows the reference's
/root/reference/enspara/cluster/util.py:315-549 and mpi/io.py:16-194; mdtraj (trajectory
files) is imported lazily and only needed for that format, as in the reference.  `.npy` and
`.h5` files need neither mdtraj nor PyTables (enspara_b200.util.h5min).
"""
import logging
import os
import pickle
from glob import glob

import numpy as np

from .. import mpi, ra
from ..exception import ImproperlyConfigured

logger = logging.getLogger(__name__)


def expand_files(pgroups):
    """Glob-expand groups of path patterns (util.py:315-321)."""
    out = []
    for pgroup in pgroups:
        out.append([])
        for p in pgroup:
            out[-1].extend(sorted(glob(p)))
    return out


def load_npy_as_striped(filenames, stride=1):
    """File i is loaded by rank i % size (mpi/io.py:68-139).  Returns (global lengths, local
    concatenated array)."""
    specs = [(h.shape, h.dtype) for h in (np.load(f, mmap_mode="r") for f in filenames)]
    shape0, dtype = specs[0]
    for i, (s, d) in enumerate(specs):
        if s[1:] != shape0[1:]:
            raise ImproperlyConfigured(
                "Subsequent dimensions of file '{}' didn't match shape of first file, '{}' "
                "({} != {})".format(filenames[0], filenames[i], shape0, s))
        if d != dtype:
            raise ImproperlyConfigured(
                "Type of file '{}' didn't match type first file, '{}' ({} != {})".format(
                    filenames[0], filenames[i], dtype, d))
    if len(filenames) < mpi.size():
        raise ImproperlyConfigured(
            "To stripe files across workers, at least 1 file per rank must be given. "
            "World size is %s, number of files is %s." % (mpi.size(), len(filenames)))
    global_lengths = [len(range(0, s[0], stride)) for s, _ in specs]
    local = [np.load(f, mmap_mode="r")[::stride] for f in filenames[mpi.rank()::mpi.size()]]
    data = np.concatenate(local) if local else np.empty((0,) + shape0[1:], dtype=dtype)
    return global_lengths, np.ascontiguousarray(data)


def load_h5_as_striped(filename, stride=1):
    """Node i of an .h5 file is loaded by rank i % size (mpi/io.py:16-65).  Every rank reads
    the (tiny) node table itself instead of a broadcast from rank 0; PyTables is not needed
    (enspara_b200.util.h5min)."""
    from ..util import h5min
    f = h5min.File(filename)
    all_keys = sorted(f.keys())
    all_shapes = [f[k].shape for k in all_keys]
    global_lengths = [len(range(0, s[0], stride)) for s in all_shapes]
    if len(all_keys) == 2 and "array" in all_keys and "lengths" in all_keys:
        raise NotImplementedError(
            "Parallel loading of RaggedArrays that have been stored as "
            "arrays and lengths cannot be loaded in parallel.")
    if len(all_keys) < mpi.size():
        raise ImproperlyConfigured(
            "To stripe nodes across workers, at least 1 node per rank must be given. "
            "World size is %s, number of nodes is %s." % (mpi.size(), len(all_keys)))
    local = ra.load(filename, keys=all_keys[mpi.rank()::mpi.size()], stride=stride)
    if hasattr(local, "_data"):
        local = local._data
    elif stride != 1:
        local = local[::stride]      # ra.load returns a single node whole (ra.py:155-158)
    return global_lengths, np.ascontiguousarray(local)


def load_features(features, stride):
    if len(features) == 1 and os.path.splitext(features[0])[1].lower() in (".h5", ".hdf5"):
        lengths, data = load_h5_as_striped(features[0], stride)
    else:
        lengths, data = load_npy_as_striped(features, stride)
    logger.info("Loaded %s trajectories with %s frames with stride %s.", len(lengths),
                len(data), stride)
    return lengths, data


def _traj_backend():
    """mdtraj when it is installed, else the native readers of util/traj.py (.xtc + .pdb)."""
    try:
        import mdtraj as md
        if getattr(md, "__file__", None) and hasattr(md, "load"):
            return md, True
    except ImportError:
        pass
    from ..util import traj
    return traj, False


def load_trajectories(topologies, trajectories, selections, stride):
    """Trajectory files -> (global lengths, Trajectory of this rank's frames); file i of the
    flattened list goes to rank i % size (util.py:350-404, mpi/io.py:142-194).  With mdtraj
    every format it reads; without it .xtc trajectories with .pdb topologies (native decoder,
    stride and atom selection applied while decoding)."""
    md, have_md = _traj_backend()
    flat, tops, inds, n_inds, top = [], [], [], None, None
    for topfile, trjset, selection in zip(topologies, trajectories, selections):
        top = md.load(topfile).top if have_md else md.load_topology(topfile)
        try:
            indices = top.select(selection)
        except Exception:
            raise ImproperlyConfigured(
                "The provided selection '{s}' didn't match the topology file, {t}".format(
                    s=selection, t=topfile))
        if n_inds is not None and n_inds != len(indices):
            raise ImproperlyConfigured(
                "Selection on topology %s selected %s atoms, but other selections selected "
                "%s atoms." % (topfile, len(indices), n_inds))
        n_inds = len(indices)
        for trj in trjset:
            flat.append(trj)
            tops.append(top)
            inds.append(indices)
    if n_inds == 0:
        raise ImproperlyConfigured("No atoms selected for clustering")   # util.py:386 (assert)
    if len(flat) < mpi.size():
        raise ImproperlyConfigured(
            "To stripe files across workers, at least 1 file per rank must be given.")
    mine = range(mpi.rank(), len(flat), mpi.size())
    loaded = [md.load(flat[i], top=tops[i], stride=stride, atom_indices=inds[i]) for i in mine]
    local_lengths = np.array([len(t) for t in loaded], dtype=int)
    lengths = mpi.ops.assemble_striped_array(local_lengths)
    xyz = np.concatenate([t.xyz for t in loaded])
    sel_top = top.subset(top.select(selections[-1]))
    return lengths, md.Trajectory(xyz=xyz, topology=sel_top)


def load_trjs_or_features(args):
    if args.features:
        return load_features(args.features, stride=args.subsample)
    return load_trajectories(args.topologies, args.trajectories, args.atoms, args.subsample)


def _intermediate_path(path, tag):
    d = os.path.join(os.path.dirname(path), "intermediate-%s" % tag)
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, os.path.basename(path))


def write_centers_indices(path, indices, intermediate_n=None):
    """(trajectory, frame) pairs as .npy (util.py:464-478)."""
    if not path:
        logger.info("--center-indices not provided, not writing center indices to file.")
        return
    if intermediate_n is not None:
        path = _intermediate_path(path, intermediate_n)
    with open(path, "wb") as f:
        np.save(f, indices)


def load_asymm_frames(center_indices, trajectories, topologies, subsample):
    """Full-topology structures of the cluster centres, re-loaded from the trajectory files at
    ``(trajectory, frame * subsample)`` (util.py:407-431): what the reference pickles, ALL
    atoms, not only the ``--atoms`` selection that was clustered."""
    md, have_md = _traj_backend()
    flat, tops = [], []
    for topfile, trjset in zip(topologies, trajectories):
        top = md.load(topfile).top if have_md else md.load_topology(topfile)
        for trj in trjset:
            flat.append(trj)
            tops.append(top)
    frames = []
    for t, f in center_indices:
        frames.append(md.load_frame(flat[int(t)], int(f) * int(subsample), top=tops[int(t)]))
    return frames


def write_centers(result, args, intermediate_n=None):
    """Feature centres -> .npy; trajectory centres -> pickle of full-topology frames re-loaded
    from the trajectory files (util.py:481-508)."""
    path = args.center_features
    if intermediate_n is not None:
        path = _intermediate_path(path, intermediate_n)
    if args.features:
        np.save(path, np.asarray(result.centers))
        return
    try:
        centers = load_asymm_frames(result.center_indices, args.trajectories, args.topologies,
                                    args.subsample)
    except ImproperlyConfigured as exc:
        # a format neither mdtraj (absent) nor the native readers handle: keep the clustered
        # (atom-sliced) frames rather than nothing, and say so
        logger.warning("Could not re-load full centre structures (%s); writing the clustered "
                       "atom selection instead.", exc)
        centers = result.centers
    with open(path, "wb") as f:
        pickle.dump(centers, f)


def write_assignments_and_distances(result, args, intermediate_n=None):
    """util.py:511-549: subsample == 1 writes the clustering's own arrays; otherwise every
    frame of the input trajectories is re-assigned to the centres found on the subsample
    (streamed through cluster/reassign.py) unless --no-reassign."""
    dpath, apath = args.distances, args.assignments
    if intermediate_n is not None:
        dpath = _intermediate_path(dpath, intermediate_n)
        apath = _intermediate_path(apath, intermediate_n)
    if args.subsample == 1:
        ra.save(dpath, result.distances)
        ra.save(apath, result.assignments)
    elif not args.no_reassign:
        from .reassign import reassign
        logger.debug("Reassigning data from subsampling of %s", args.subsample)
        assig, dist = reassign(args.topologies, args.trajectories, args.atoms,
                               centers=result.centers)
        ra.save(dpath, dist)
        ra.save(apath, assig)
    else:
        logger.debug("Got --no-reassign, not doing reassigment")


def write_intermediate(result, args, lengths, tag):
    """--save_intermediates dumps (hybrid.py:129-151, kmedoids.py:459-473)."""
    if mpi.size() > 1:
        return  # intermediates are written from assembled results only in serial runs
    part = result.partition(lengths)
    write_centers_indices(args.center_indices,
                          [(t, f * args.subsample) for t, f in part.center_indices],
                          intermediate_n=tag)
    write_centers(part, args, intermediate_n=tag)
    write_assignments_and_distances(part, args, intermediate_n=tag)
