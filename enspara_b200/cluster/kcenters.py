"""K-centers clustering on B200: same estimator / function API as the reference
(/root/reference/enspara/cluster/kcenters.py), the loop body replaced by one fused kernel
launch per iteration (see cluster/_engine.py and csrc/eb_rmsd_kcenters.cu).
"""
import logging
import time

import numpy as np
from sklearn.base import BaseEstimator, ClusterMixin
from sklearn.utils import check_random_state

from .. import mpi
from ..exception import ImproperlyConfigured
from . import util
from ._engine import KCentersEngine

logger = logging.getLogger(__name__)

#: phase -> seconds of the last kcenters() call when ENSPARA_B200_PHASE_TIMES=1 (developer
#: timing: each mark synchronises the device, so it is off by default)
last_phase_times = {}


class _Phases:
    def __init__(self):
        import os
        self.on = os.environ.get("ENSPARA_B200_PHASE_TIMES", "0") == "1"
        self.t = time.perf_counter()
        self.d = {}

    def mark(self, name):
        if not self.on:
            return
        import torch
        torch.cuda.synchronize()
        now = time.perf_counter()
        self.d[name] = self.d.get(name, 0.0) + now - self.t
        self.t = now
        last_phase_times.clear()
        last_phase_times.update(self.d)


class KCenters(BaseEstimator, ClusterMixin, util.MolecularClusterMixin):
    """Sklearn-style k-centers (Gonzalez 1985) -- reference: kcenters.py:18-100.

    Parameters
    ----------
    metric : 'rmsd' | 'euclidean' | 'manhattan' | 'cityblock' | 'sqeuclidean' | recognised
        function object (``mdtraj.rmsd``, ``libdist.euclidean`` ...).
    n_clusters : int, optional
        Stop after this many centres.
    cluster_radius : float, optional
        Stop once every frame is within this distance of its centre.
    random_first_center : bool
        Not implemented in the reference either (raises NotImplementedError).
    random_state : int or RandomState
    mpi_mode : bool, optional
        Frames are sharded over ranks (one process per GPU).  ``None`` -> automatic from the
        process-group size, like the reference's ``mpi.size() != 1``.
    """

    def __init__(self, metric, n_clusters=None, cluster_radius=None,
                 random_first_center=False, random_state=None, mpi_mode=None):
        if n_clusters is None and cluster_radius is None:
            raise ImproperlyConfigured("Either n_clusters or cluster_radius "
                                       "is required for KHybrid clustering")
        self.metric = util._get_distance_method(metric)
        self.n_clusters = n_clusters
        self.cluster_radius = cluster_radius
        self.random_first_center = random_first_center
        self.random_state = check_random_state(random_state)
        self.mpi_mode = mpi.size() != 1 if mpi_mode is None else mpi_mode

    def fit(self, X, init_centers=None):
        """Cluster ``X`` (md.Trajectory-like, ndarray, or a device container)."""
        t0 = time.perf_counter()
        self.result_ = kcenters(
            X, distance_method=self.metric, n_clusters=self.n_clusters,
            dist_cutoff=self.cluster_radius, init_centers=init_centers,
            random_first_center=self.random_first_center, mpi_mode=self.mpi_mode)
        self.runtime_ = time.perf_counter() - t0
        return self


def kcenters_mpi(*args, **kwargs):
    kwargs.pop("mpi_mode", None)
    return kcenters(*args, mpi_mode=True, **kwargs)


def _take_centers(traj, data, indices):
    """``[traj[i] for i in indices]`` like kcenters.py:283 -- slices of the caller's object
    when there is one, centred host frames when the data only exists on the device."""
    from ..device import DeviceFeatures, DeviceTrajectory
    if isinstance(traj, DeviceTrajectory):
        if not indices:
            return []
        frames = traj.gather(indices).to_host_aos()
        return [frames[i] for i in range(len(indices))]
    if isinstance(traj, DeviceFeatures):
        rows = traj.X[np.asarray(indices, dtype=np.int64)].cpu().numpy() if indices else []
        return [r for r in rows]
    return [traj[int(i)] for i in indices]


def kcenters(traj, distance_method, n_clusters=np.inf, dist_cutoff=0, init_centers=None,
             random_first_center=False, use_triangle_inequality=False, mpi_mode=False,
             exact=True, _return_engine=False):
    """Function form of k-centers; reference: kcenters.py:108-240.

    Returns ``ClusterResult(center_indices, distances, assignments, centers)`` with
    ``assignments`` int64 and ``distances`` float64 host arrays.  In ``mpi_mode`` the arrays
    cover this rank's frames and centre indices are ``(owner_rank, local_index)`` pairs
    (kcenters.py:375-376).

    ``use_triangle_inequality`` (kcenters.py:287-296): frames whose distance to their centre is
    at most half the distance between that centre and the new one are not re-evaluated -- on
    the GPU they are not even read, which removes most of the HBM traffic once there are many
    centres.  Results are the same as without it (test_cluster.py:710-770).  RMSD only; feature
    metrics always evaluate every row.  ``exact=False`` selects the float32-block accumulation
    mode of the kernel (~1e-5 relative).
    """
    if (n_clusters is np.inf) and (dist_cutoff == 0):
        raise ImproperlyConfigured("Either n_clusters or cluster_radius "
                                   "is required for KHybrid clustering")
    metric = util._get_distance_method(distance_method)

    if n_clusters is None and dist_cutoff is None:
        raise ImproperlyConfigured(
            "KCenters must specify 'n_clusters' or 'distance_cutoff'")
    elif n_clusters is None and dist_cutoff is not None:
        n_clusters = np.inf
    elif n_clusters is not None and dist_cutoff is None:
        dist_cutoff = 0

    if random_first_center:
        raise NotImplementedError(
            "We haven't implemented kcenters 'random_first_center' yet.")

    ph = _Phases()
    data = metric.to_device(traj)
    ph.mark("upload+centre")
    comm = mpi.comm if mpi_mode else _SingleComm()
    engine = KCentersEngine(data, metric.kind, comm, exact=exact,
                            triangle=use_triangle_inequality)
    ph.mark("engine_setup")

    centers = []
    ctr_inds = []
    n_existing = 0
    if init_centers is not None:
        from . import _ops
        centers = [c for c in init_centers]
        logger.info("Updating assignments to previous cluster centers")
        cdev = _ops.centers_to_device(metric, centers, data)
        _ops.assign_device(metric, data, cdev, out_dist=engine.dist,
                           out_assign=engine.assign, accumulate=False)
        a_host, d_host = engine.results_host()
        ctr_inds = list(util.find_cluster_centers(a_host, d_host))
        n_existing = len(ctr_inds)
        if mpi_mode and comm.size > 1:
            # a shard without members of some initial centre would count fewer existing
            # centres than its peers and then use other centre ids, launch counts and
            # collectives: agree on the labels present on ANY rank (one small all-reduce)
            import torch
            present = torch.zeros(max(len(centers), 1), dtype=torch.int32, device=engine.dev)
            if len(a_host):
                present[torch.as_tensor(np.unique(a_host), device=engine.dev)] = 1
            comm.all_reduce_max(present)
            n_existing = int(present.sum().item())
        if n_existing == len(centers):
            engine.preload_centers(cdev)
        else:
            # an initial centre without members: new centre ids would collide with the ids of
            # the supplied centres (a reference quirk); prune nothing rather than prune wrongly
            engine.triangle = False

    new_global, maxdist = engine.run(n_clusters, dist_cutoff, n_existing=n_existing)
    ph.mark("iterations")

    if mpi_mode:
        new_inds = [engine.shard.to_rank_local(g) for g in new_global]
        new_centers = _distribute_centers(engine, data, new_global)
    else:
        new_inds = [np.int64(g) for g in new_global]
        new_centers = _take_centers(traj, data, new_global)
    ctr_inds.extend(new_inds)
    centers.extend(new_centers)

    logger.info("Terminated k-centers with n=%s and d=%0.6f.", len(ctr_inds), maxdist)

    ph.mark("centres")
    assignments, distances = engine.results_host()
    ph.mark("results_d2h")
    result = util.ClusterResult(center_indices=ctr_inds, assignments=assignments,
                                distances=distances, centers=centers)
    if _return_engine:
        return result, engine
    return result


class _SingleComm:
    """Communicator of a run that is not sharded, whatever the process group looks like."""
    size = 1
    rank = 0

    def all_gather_object(self, obj):
        return [obj]

    def broadcast_object(self, obj, root=0):
        return obj

    def all_gather_into(self, out, inp):
        if out.data_ptr() != inp.data_ptr():
            out.view(-1)[:inp.numel()].copy_(inp.view(-1))

    def all_reduce_max(self, t):
        return t

    def all_reduce_sum(self, t):
        return t

    def broadcast(self, t, root):
        return t

    def barrier(self):
        pass


def _distribute_centers(engine, data, global_indices):
    """Every rank gets the coordinates of every centre (the reference broadcasts each one as it
    is chosen, kcenters.py:345-348 -> mpi/ops.py:169-212; here one collective at the end).
    Frames come from the owner's host object, i.e. the caller's original coordinates."""
    import torch
    from ..device import DeviceTrajectory
    shard = engine.shard
    k = len(global_indices)
    is_traj = isinstance(data, DeviceTrajectory)
    shape = (data.n_atoms, 3) if is_traj else (data.n_features,)
    dtype = np.float32 if is_traj else data.np_dtype
    host = np.zeros((k,) + shape, dtype=dtype)
    owned = [(j, int(g - shard.offset)) for j, g in enumerate(global_indices)
             if shard.offset <= g < shard.offset + engine.n]
    src = data.host
    if owned:
        if src is not None:
            for j, l in owned:
                fr = src[l]
                host[j] = np.asarray(fr.xyz if hasattr(fr, "xyz") else fr).reshape(shape)
        elif is_traj:
            frames = data.gather([l for _, l in owned]).to_host_aos()
            for (j, _), fr in zip(owned, frames):
                host[j] = fr
        else:
            rows = data.X[[l for _, l in owned]].cpu().numpy()
            for (j, _), r in zip(owned, rows):
                host[j] = r
    if shard.size > 1 and k > 0:
        t = torch.from_numpy(host).to(engine.dev)
        engine.comm.all_reduce_sum(t)
        host = t.cpu().numpy()
    if is_traj and src is not None and hasattr(src, "xyz") and hasattr(src, "top"):
        try:
            return [type(src)(xyz=host[j][None], topology=src.top) for j in range(k)]
        except Exception:  # pragma: no cover - exotic trajectory types
            pass
    return [host[j] for j in range(k)]
