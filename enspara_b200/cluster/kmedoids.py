"""K-medoids (PAM) on B200 with the reference's API
(/root/reference/enspara/cluster/kmedoids.py).  The sweep itself lives in cluster/_pam.py.
"""
import logging
import time

import numpy as np
from sklearn.base import BaseEstimator, ClusterMixin
from sklearn.utils import check_random_state

from .. import exception, mpi
from ..exception import ImproperlyConfigured
from . import util
from ._pam import PamEngine
from .kcenters import _SingleComm

logger = logging.getLogger(__name__)


class KMedoids(BaseEstimator, ClusterMixin, util.MolecularClusterMixin):
    """Sklearn-style k-medoids -- reference: kmedoids.py:28-105."""

    def __init__(self, metric, n_clusters=None, n_iters=5, args=None, lengths=None):
        self.metric = util._get_distance_method(metric)
        self.n_clusters = n_clusters
        self.n_iters = n_iters
        self.args = args
        self.lengths = lengths

    def fit(self, X, assignments=None, distances=None, cluster_center_inds=None,
            X_lengths=None, args=None):
        t0 = time.perf_counter()
        self.result_ = kmedoids(
            X, distance_method=self.metric, n_clusters=self.n_clusters, n_iters=self.n_iters,
            assignments=assignments, distances=distances,
            cluster_center_inds=cluster_center_inds, X_lengths=X_lengths, args=args)
        self.runtime_ = time.perf_counter() - t0
        return self


def _msq(x):
    """mean(x^2) over all ranks (kmedoids.py:478-479) for host arrays; the device sweeps use
    the deterministic on-device reduction instead."""
    return mpi.ops.striped_array_mean(np.square(x))


def kmedoids(X, distance_method, n_clusters=None, n_iters=5, assignments=None, distances=None,
             cluster_center_inds=None, proposals=None, X_lengths=None, args=None, lengths=None,
             random_state=None):
    """Function form -- reference: kmedoids.py:108-202 (argument checks and warm/cold start
    normalisation kept; MPI cold start is broken in the reference, SURVEY.md App. A.7(ii),
    and raises ImproperlyConfigured here)."""
    if cluster_center_inds is not None:
        if hasattr(cluster_center_inds[0], "__len__") and X_lengths is None:
            raise ImproperlyConfigured(
                "If cluster_center_inds is given as [[global_traj_id, frame_id],...]"
                "then X_lengths also needs to be supplied")

    if cluster_center_inds is None and n_clusters is None:
        if mpi.size() > 1:
            raise ImproperlyConfigured(
                "Must provide n_clusters or cluster_center_inds, assignments,"
                "and distances for KMedoids in MPI mode.")
        elif assignments is None and distances is None:
            raise ImproperlyConfigured(
                "Must provide n_clusters or cluster_center_inds or "
                " (assignments and distances) for KMedoids")

    metric = util._get_distance_method(distance_method)
    data = metric.to_device(X)

    if mpi.size() > 1:
        if not (cluster_center_inds is not None and distances is not None
                and assignments is not None):
            raise ImproperlyConfigured(
                "For KMedoids, MPI mode requires that assignments, distances and "
                "cluster_center_inds are all supplied.")
        cluster_center_inds = ctr_ids_mpi(cluster_center_inds, X_lengths)
        local = [p[1] for p in cluster_center_inds if p[0] == mpi.rank()]
        assert np.all(np.asarray(distances)[local] < 0.001)
    else:
        assignments, distances, cluster_center_inds = _kmedoids_inputs_tree(
            X, data, metric, n_clusters, assignments, distances, cluster_center_inds,
            X_lengths, random_state=random_state)
        # should be all 0s up to rounding (kmedoids.py:195-197)
        assert np.all(np.asarray(distances)[np.asarray(cluster_center_inds, dtype=int)] < 0.001)

    return _kmedoids_iterations(
        X, metric, n_iters, cluster_center_inds, assignments, distances, proposals=proposals,
        args=args, lengths=lengths, random_state=random_state, _data=data)


def _kmedoids_inputs_tree(X, data, metric, n_clusters, assignments, distances,
                          cluster_center_inds, X_lengths, random_state=None):
    """Warm / cold start normalisation -- reference: kmedoids.py:283-363."""
    rng = np.random.default_rng(seed=random_state)

    if ((assignments is not None and distances is None) or
            (assignments is None and distances is not None)):
        raise ImproperlyConfigured(
            "Assignments and distances need to both be supplied, or neither supplied.")

    if cluster_center_inds is None:
        if assignments is not None and distances is not None:
            cluster_center_inds = util.find_cluster_centers(assignments, distances)
        else:
            # redraw until every cluster got a distinct frame (kmedoids.py:343-350)
            cluster_center_inds = np.array([])
            while len(np.unique(cluster_center_inds)) < n_clusters:
                cluster_center_inds = rng.integers(0, len(data), n_clusters)
    elif hasattr(cluster_center_inds[0], "__len__"):
        cluster_center_inds = [sum(X_lengths[:cluster_center_inds[i][0]])
                               + cluster_center_inds[i][1]
                               for i in np.arange(len(cluster_center_inds))]

    if assignments is None and distances is None:
        from . import _ops
        idx = [int(i) for i in cluster_center_inds]
        if isinstance(X, type(data)):
            centers = data.gather(idx) if metric.is_rmsd else type(data)(data.X[idx])
        else:
            centers = _ops.centers_to_device(metric, X[idx], data)
        d, a = _ops.assign_device(metric, data, centers)
        assignments = a.cpu().numpy().astype(np.int64)
        distances = d.cpu().numpy().astype(np.float64)

    return assignments, distances, cluster_center_inds


def ctr_ids_mpi(cluster_center_inds, lengths):
    """[(traj, frame)] or [global index] -> [(rank, local index)] for trajectories striped over
    ranks file by file (reference: kmedoids.py:365-408, trajectory i on rank i % size)."""
    num_procs = mpi.size()
    lengths = np.asarray(lengths, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lengths)])
    pairs = []
    for c in cluster_center_inds:
        if hasattr(c, "__len__"):
            pairs.append((int(c[0]), int(c[1])))
        else:
            t = int(np.searchsorted(starts, c, side="right") - 1)
            pairs.append((t, int(c - starts[t])))
    out = []
    for traj_id, frame_id in pairs:
        r = traj_id % num_procs
        owned = np.arange(len(lengths))[r::num_procs]
        before = int(lengths[owned[:traj_id // num_procs]].sum())
        out.append((r, before + frame_id))
    return out


def _to_global(cluster_center_inds, shard):
    """Serial ints or (rank, local) pairs -> global frame indices."""
    out = []
    for c in cluster_center_inds:
        if hasattr(c, "__len__"):
            out.append(int(shard.offsets[int(c[0])]) + int(c[1]))
        else:
            out.append(int(c))
    return out


def _from_global(global_inds, shard, pairs):
    if pairs:
        return [shard.to_rank_local(g) for g in global_inds]
    return [np.int64(g) for g in global_inds]


def _medoid_coords(X, data, engine, pairs):
    """Centre coordinates to return (kmedoids.py:598-607, 699)."""
    from .kcenters import _distribute_centers, _take_centers
    if pairs:
        class _E:  # minimal engine view for _distribute_centers
            pass
        e = _E()
        e.shard, e.n, e.dev, e.comm = engine.shard, engine.n, engine.dev, engine.comm
        return _distribute_centers(e, data, engine.medoid_global)
    return _take_centers(X, data, engine.medoid_global)


def _kmedoids_iterations(X, distance_method, n_iters, cluster_center_inds, assignments,
                         distances, proposals=None, args=None, lengths=None, random_state=None,
                         _data=None):
    """Sweep driver -- reference: kmedoids.py:410-476.  ``random_state`` is handed to every
    sweep unchanged: an int re-seeds each sweep, a RandomState object carries over
    (SURVEY.md App. A.5)."""
    metric = util._get_distance_method(distance_method)
    data = _data if _data is not None else metric.to_device(X)
    pairs = len(cluster_center_inds) > 0 and hasattr(cluster_center_inds[0], "__len__")
    comm = mpi.comm if pairs else _SingleComm()
    engine = None
    result = None
    for i in range(n_iters):
        if engine is None:
            from ._engine import ShardInfo
            shard = ShardInfo(len(data), comm)
            engine = PamEngine(data, metric, comm, distances, assignments,
                               _to_global(cluster_center_inds, shard))
        _sweep(engine, proposals, random_state, pairs)
        if args is not None and getattr(args, "save_intermediates", False) \
                and i != n_iters - 1:
            from . import io as cio
            a, d = engine.results_host()
            inter = util.ClusterResult(
                center_indices=_from_global(engine.medoid_global, engine.shard, pairs),
                assignments=a, distances=d,
                centers=_medoid_coords(X, data, engine, pairs))
            cio.write_intermediate(inter, args, lengths, "kmedoids-%d" % i)
        logger.info("KMedoids update %s", i)
    if engine is None:
        return util.ClusterResult(center_indices=cluster_center_inds, assignments=assignments,
                                  distances=distances, centers=None)
    a, d = engine.results_host()
    new_inds = _from_global(engine.medoid_global, engine.shard, pairs)
    # the reference mutates the caller's list in place (kmedoids.py:689)
    try:
        for j, v in enumerate(new_inds):
            cluster_center_inds[j] = v
        center_indices = cluster_center_inds
    except TypeError:
        center_indices = new_inds
    result = util.ClusterResult(center_indices=center_indices, assignments=a, distances=d,
                                centers=_medoid_coords(X, data, engine, pairs))
    return result


def _sweep(engine, proposals, random_state, pairs):
    if proposals is not None:
        if len(proposals) != engine.k:
            raise exception.DataInvalid(
                "Length of 'proposals' didn't match length of 'medoid_inds' "
                "({} != {}).".format(len(proposals), engine.k))
        if hasattr(proposals[0], "__len__") != pairs:
            raise exception.DataInvalid(
                "Depth of 'proposals' didn't match 'medoid_inds' "
                "(proposals[0] == {})".format(proposals[0]))
        proposals = _to_global(proposals, engine.shard)
    acc = engine.sweep(proposals=proposals, random_state=random_state)
    logger.info("Kmedoid sweep reduced cost to %.7f (%.2f%% acceptance)",
                engine.last_cost, acc / max(engine.k, 1) * 100)


def _kmedoids_pam_update(X, metric, medoid_inds, assignments, distances, proposals=None,
                         cost=_msq, random_state=None):
    """One PAM sweep on host arrays -- reference: kmedoids.py:520-699.  Returns
    ``(medoid_inds, distances, assignments, medoid_coords)`` (note the order) and mutates
    ``medoid_inds`` in place like the reference (:689).  Only the default mean-square cost is
    fused on the device."""
    if cost is not _msq:
        raise ImproperlyConfigured(
            "Only the default mean-square cost is implemented on the GPU path.")
    assert np.issubdtype(np.asarray(assignments).dtype, np.integer)
    metric = util._get_distance_method(metric)
    data = metric.to_device(X)
    assert len(assignments) == len(data)
    assert len(distances) == len(data)
    pairs = len(medoid_inds) > 0 and hasattr(medoid_inds[0], "__len__")
    comm = mpi.comm if pairs else _SingleComm()
    from ._engine import ShardInfo
    shard = ShardInfo(len(data), comm)
    engine = PamEngine(data, metric, comm, distances, assignments,
                       _to_global(medoid_inds, shard))
    _sweep(engine, proposals, random_state, pairs)
    new_inds = _from_global(engine.medoid_global, engine.shard, pairs)
    try:
        for j, v in enumerate(new_inds):
            medoid_inds[j] = v
    except TypeError:
        medoid_inds = new_inds
    a, d = engine.results_host()
    return medoid_inds, d, a, _medoid_coords(X, data, engine, pairs)
