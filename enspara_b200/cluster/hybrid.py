"""K-hybrid = k-centers followed by PAM sweeps, with the reference's API
(/root/reference/enspara/cluster/hybrid.py).  Frames are uploaded once and stay in HBM for both
phases; distances / assignments are handed from the k-centers loop to the PAM engine on the
device."""
import logging
import time

import numpy as np
from sklearn.base import BaseEstimator, ClusterMixin
from sklearn.utils import check_random_state

from .. import mpi
from ..exception import ImproperlyConfigured
from . import kcenters, kmedoids, util

logger = logging.getLogger(__name__)


class KHybrid(BaseEstimator, ClusterMixin, util.MolecularClusterMixin):
    """Sklearn-style k-hybrid -- reference: hybrid.py:28-109."""

    def __init__(self, metric, n_clusters=None, cluster_radius=None, kmedoids_updates=5,
                 random_first_center=False, random_state=None, mpi_mode=None, args=None,
                 lengths=None):
        if n_clusters is None and cluster_radius is None:
            raise ImproperlyConfigured("Either n_clusters or cluster_radius "
                                       "is required for KHybrid clustering")
        self.kmedoids_updates = kmedoids_updates
        self.n_clusters = n_clusters
        self.cluster_radius = cluster_radius
        self.random_first_center = random_first_center
        self.metric = util._get_distance_method(metric)
        # one legacy RandomState spans all sweeps of the estimator (hybrid.py:78)
        self.random_state = check_random_state(random_state)
        self.mpi_mode = mpi_mode if mpi_mode is not None else mpi.size() != 1
        self.args = args
        self.lengths = lengths

    def fit(self, X, init_centers=None, args=None):
        t0 = time.perf_counter()
        self.result_ = hybrid(
            X, self.metric, n_iters=self.kmedoids_updates, n_clusters=self.n_clusters,
            dist_cutoff=self.cluster_radius, random_first_center=self.random_first_center,
            init_centers=init_centers, random_state=self.random_state,
            mpi_mode=self.mpi_mode, args=self.args, lengths=self.lengths)
        self.runtime_ = time.perf_counter() - t0
        return self


def hybrid(X, distance_method, n_iters=5, n_clusters=np.inf, dist_cutoff=0,
           random_first_center=False, init_centers=None, random_state=None, mpi_mode=False,
           args=None, lengths=None):
    """Function form -- reference: hybrid.py:112-162."""
    metric = util._get_distance_method(distance_method)
    data = metric.to_device(X)
    if not hasattr(data, "host") or data.host is None:
        data.host = X if X is not data else None

    result, kc_engine = kcenters.kcenters(
        data, metric, n_clusters=n_clusters, dist_cutoff=dist_cutoff,
        init_centers=init_centers, random_first_center=random_first_center, mpi_mode=mpi_mode,
        _return_engine=True)
    if X is not data and not mpi_mode:
        # centres are slices of the caller's object, like kcenters.py:283
        result = result._replace(centers=[X[int(i)] for i in result.center_indices])

    if args is not None and getattr(args, "save_intermediates", False):
        from . import io as cio
        cio.write_intermediate(result, args, lengths, "kcenters")

    if n_iters > 0:
        # the k-centers state goes to the PAM engine as DEVICE tensors (same values as the host
        # copies in `result`; no device -> host -> device round trip between the phases)
        return kmedoids._kmedoids_iterations(
            X, metric, n_iters, result.center_indices, kc_engine.assign, kc_engine.dist,
            args=args, lengths=lengths, random_state=random_state, _data=data)
    return result
