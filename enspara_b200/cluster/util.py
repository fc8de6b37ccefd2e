"""Algorithmic helpers of the clustering path, with the reference's names and behaviour
(/root/reference/enspara/cluster/util.py:46-242, 289-313).  The file/IO half of the reference
module lives in ``enspara_b200.cluster.io`` (a "next" row of SURVEY.md 8f).
"""
import logging
from collections import namedtuple

import numpy as np
import torch

from .. import _lib, ra
from ..device import DeviceFeatures, DeviceTrajectory, ptr, stream_ptr
from ..exception import DataInvalid, ImproperlyConfigured
from ..ra import partition_indices, partition_list

logger = logging.getLogger(__name__)


# ---------------------------------------------------------------------------------------------
# metrics: the b1 plug point (cluster/util.py:289-313)
# ---------------------------------------------------------------------------------------------
class Metric:
    """A distance the fused kernels implement.  Callable with the reference's protocol
    ``d = metric(X, y) -> ndarray[n]`` (kcenters.py:298, kmedoids.py:637, util.py:195,200):
    one datum against all data, computed on the GPU."""

    def __init__(self, kind):
        self.kind = kind
        self.__name__ = kind

    @property
    def is_rmsd(self):
        return self.kind == "rmsd"

    def __repr__(self):
        return "<enspara_b200 metric %r>" % self.kind

    def __eq__(self, other):
        return isinstance(other, Metric) and other.kind == self.kind

    def __hash__(self):
        return hash(("enspara_b200.Metric", self.kind))

    def to_device(self, X):
        """Upload data for this metric (no-op for device containers)."""
        if isinstance(X, (DeviceTrajectory, DeviceFeatures)):
            return X
        if self.is_rmsd:
            return DeviceTrajectory.from_host(X)
        return DeviceFeatures.from_host(X)

    def __call__(self, X, y, out=None):
        from . import _ops
        return _ops.one_to_all(self, X, y, out=out)


RMSD = Metric("rmsd")
EUCLIDEAN = Metric("euclidean")
MANHATTAN = Metric("manhattan")
SQEUCLIDEAN = Metric("sqeuclidean")

_BY_NAME = {"rmsd": RMSD, "euclidean": EUCLIDEAN, "cityblock": MANHATTAN,
            "manhattan": MANHATTAN, "sqeuclidean": SQEUCLIDEAN}

#: names the reference forwards to msmbuilder's libdistance (util.py:41-43); no kernels here
msmbuilder_libdistance_metrics = ["euclidean", "sqeuclidean", "cityblock", "chebyshev",
                                  "canberra", "braycurtis", "hamming", "jaccard"]


def _recognise_callable(fn):
    """Map well-known function objects onto fused kernels: the CLI and user code pass
    ``md.rmsd`` / ``libdist.euclidean`` objects, not strings (apps/cluster.py:179,210)."""
    mod = getattr(fn, "__module__", "") or ""
    name = getattr(fn, "__name__", "") or ""
    if name == "rmsd" and (mod.startswith("mdtraj") or mod.startswith("enspara_b200")):
        return RMSD
    if name in ("euclidean", "manhattan") and "libdist" in mod:
        return _BY_NAME[name]
    return None


def _get_distance_method(metric):
    """String / callable -> Metric, raising ImproperlyConfigured like util.py:289-313.

    Arbitrary Python callables cannot be fused into the device loop and there is no CPU
    fallback, so they are rejected with the same exception type the reference uses for unknown
    metrics."""
    if isinstance(metric, Metric):
        return metric
    if isinstance(metric, str):
        if metric in _BY_NAME:
            return _BY_NAME[metric]
        if metric in msmbuilder_libdistance_metrics:
            raise ImproperlyConfigured(
                "'{}' is an MSMBuilder libdistance metric; enspara_b200 implements 'rmsd', "
                "'euclidean', 'manhattan'/'cityblock' and 'sqeuclidean' only.".format(metric))
        raise ImproperlyConfigured("'{}' is not a recognized metric".format(metric))
    if callable(metric):
        m = _recognise_callable(metric)
        if m is not None:
            return m
        raise ImproperlyConfigured(
            "Callable metric {!r} cannot run on the GPU path: pass 'rmsd', 'euclidean', "
            "'manhattan' or 'sqeuclidean' (or mdtraj.rmsd / libdist.euclidean / "
            "libdist.manhattan objects). There is no CPU fallback.".format(metric))
    raise ImproperlyConfigured("'{}' is not a recognized metric".format(metric))


# ---------------------------------------------------------------------------------------------
# results
# ---------------------------------------------------------------------------------------------
class ClusterResult(namedtuple("ClusterResult",
                               ["center_indices", "distances", "assignments", "centers"])):
    """Same field order as the reference (util.py:105-109)."""

    def partition(self, lengths):
        """Split concatenated results into per-trajectory pieces (util.py:111-156): numpy
        arrays when all lengths agree, RaggedArray otherwise."""
        square = all(lengths[0] == l for l in lengths)
        if square:
            return ClusterResult(
                assignments=np.array(partition_list(self.assignments, lengths)),
                distances=np.array(partition_list(self.distances, lengths)),
                center_indices=partition_indices(self.center_indices, lengths),
                centers=self.centers)
        return ClusterResult(
            assignments=ra.RaggedArray(self.assignments, lengths=lengths),
            distances=ra.RaggedArray(self.distances, lengths=lengths),
            center_indices=partition_indices(self.center_indices, lengths),
            centers=self.centers)


class MolecularClusterMixin:
    """predict() + result accessors (util.py:46-102)."""

    def predict(self, X):
        if not hasattr(self, "result_"):
            raise ImproperlyConfigured(
                "To predict the clustering result for new data, the clusterer first must "
                "have fit some data.")
        pred_assigs, pred_dists = assign_to_nearest_center(
            trajectory=X, cluster_centers=self.centers_, distance_method=self.metric)
        pred_centers = find_cluster_centers(pred_assigs, pred_dists)
        return ClusterResult(assignments=pred_assigs, distances=pred_dists,
                             center_indices=pred_centers, centers=self.centers_)

    @property
    def labels_(self):
        return self.result_.assignments

    @property
    def distances_(self):
        return self.result_.distances

    @property
    def center_indices_(self):
        return self.result_.center_indices

    @property
    def centers_(self):
        return self.result_.centers


def assign_to_nearest_center(trajectory, cluster_centers, distance_method):
    """Nearest of k centres for every frame (util.py:159-205): strict '<' in centre order,
    so the lowest centre index wins ties; assignments start at 0, distances at +inf.
    Returns (assignments int64[n], distances float64[n]) on the host."""
    from . import _ops
    metric = _get_distance_method(distance_method)
    return _ops.assign_host(metric, trajectory, cluster_centers)


def find_cluster_centers(assignments, distances):
    """For each label the first frame of minimum distance (util.py:208-242)."""
    assignments = np.asarray(assignments)
    distances = np.asarray(distances)
    if len(distances) != len(assignments):
        raise DataInvalid(
            "Length of distances (%s) must match length of assignments (%s)."
            % (len(distances), len(assignments)))
    unique_centers = np.unique(assignments)
    center_inds = np.zeros_like(unique_centers)
    # one stable sort replaces the reference's O(k n) np.where loop: within a label, the
    # first position of the minimum distance
    order = np.lexsort((np.arange(len(assignments)), distances, assignments))
    sorted_labels = assignments[order]
    first = np.searchsorted(sorted_labels, unique_centers, side="left")
    center_inds[:] = order[first]
    return center_inds
