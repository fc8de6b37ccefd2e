"""numpy restatement of the reference's clustering loops.  TEST INFRASTRUCTURE.

Every function states the reference lines it follows (paths under /root/reference/enspara/).
Tie rules, update rules and RNG consumption are the point of this file;
the distance arithmetic comes from ``oracle.distances``.  Validated against the real reference
package (imported with stubs) by tests/test_oracle_ref.py whenever /root/reference is present.
"""
from collections import namedtuple

import numpy as np
from sklearn.utils import check_random_state

# field order as cluster/util.py:105-109
Result = namedtuple("Result", ["center_indices", "distances", "assignments", "centers"])


def _take(X, idx):
    return X[idx]


def kcenters(X, metric, n_clusters=np.inf, dist_cutoff=0.0, init_centers=None, trace=None):
    """cluster/kcenters.py:108-240 + :243-311 (serial path).

    argmax = first occurrence (:282); strict '<' update (:304-306); loop while
    len(ctr) < n_clusters and maxdist > dist_cutoff with maxdist taken after the update
    (:217, :225-226).  ``trace`` (list) receives (index, maxdist_after) per iteration.
    """
    if n_clusters is None:
        n_clusters = np.inf
    if dist_cutoff is None:
        dist_cutoff = 0.0
    n = len(X)
    if init_centers is None:
        ctr_inds, centers = [], []
        assignments = np.full(n, -1, dtype=np.int64)
        distances = np.full(n, np.inf, dtype=np.float64)
    else:
        centers = [c for c in init_centers]
        assignments, distances = assign_to_nearest_center(X, centers, metric)
        ctr_inds = list(find_cluster_centers(assignments, distances))
    maxdist = distances.max()
    while len(ctr_inds) < n_clusters and maxdist > dist_cutoff:
        new = int(np.argmax(distances))
        center = _take(X, new)
        d = np.asarray(metric(X, center))
        upd = d < distances
        distances[upd] = d[upd]
        assignments[upd] = len(ctr_inds)
        ctr_inds.append(new)
        centers.append(center)
        maxdist = distances.max()
        if trace is not None:
            trace.append((new, float(maxdist)))
    return Result(ctr_inds, distances, assignments, centers)


def assign_to_nearest_center(X, centers, metric):
    """cluster/util.py:159-205, main branch (:198-203): centre order, strict '<', init 0/inf."""
    n = len(X)
    assignments = np.zeros(n, dtype=np.int64)
    distances = np.full(n, np.inf, dtype=np.float64)
    if len(centers) > n and hasattr(centers, "xyz"):
        # alternate branch :193-197 -- per frame argmin over all centres (first minimum)
        for i in range(n):
            d = np.asarray(metric(centers, _take(X, i)))
            assignments[i] = int(np.argmin(d))
            distances[i] = d.min()
        return assignments, distances
    for i, c in enumerate(centers):
        d = np.asarray(metric(X, c))
        upd = d < distances
        distances[upd] = d[upd]
        assignments[upd] = i
    return assignments, distances


def find_cluster_centers(assignments, distances):
    """cluster/util.py:208-242: per unique label, the first frame of minimum distance."""
    assignments = np.asarray(assignments)
    distances = np.asarray(distances)
    labels = np.unique(assignments)
    out = np.zeros_like(labels)
    for i, c in enumerate(labels):
        members = np.where(assignments == c)[0]
        out[i] = members[np.argmin(distances[members])]
    return out


def msq(x):
    """kmedoids.py:478-479 via mpi/ops.py:143-166 (single process): mean of squares."""
    x = np.square(x)
    return np.sum(x) / len(x)


def pam_update(X, metric, medoid_inds, assignments, distances, proposals=None,
               random_state=None, log=None):
    """One PAM sweep: cluster/kmedoids.py:520-699 (serial path).

    Per cluster cid: state_inds = where(assignments == cid) ascending (:611); proposal =
    random_state.choice(state_inds) (:514) unless ``proposals`` given (:622-628); full pass
    (:637); three masks (:644-658); the 'up_this' subset is re-assigned against the medoid
    list with the proposal substituted at slot cid (:660-667); accept iff
    mean(new^2) < mean(old^2) strictly (:680-694).  medoid_inds is mutated in place (:689).
    """
    rs = check_random_state(random_state)
    medoid_coords = [_take(X, i) for i in medoid_inds]
    for cid in range(len(medoid_inds)):
        state_inds = np.where(assignments == cid)[0]
        if proposals is None:
            prop_ind = rs.choice(state_inds)
        else:
            prop_ind = proposals[cid]
        prop = _take(X, prop_ind)
        new_ctr_dist = np.asarray(metric(X, prop))

        new_dist = np.zeros_like(distances) - 1
        new_assig = np.zeros_like(assignments) - 1

        dn = distances > new_ctr_dist
        new_assig[dn] = cid
        new_dist[dn] = new_ctr_dist[dn]

        up_other = (distances <= new_ctr_dist) & (assignments != cid)
        new_assig[up_other] = assignments[up_other]
        new_dist[up_other] = distances[up_other]

        up_this = (distances <= new_ctr_dist) & (assignments == cid)
        new_medoids = list(medoid_coords)
        new_medoids[cid] = prop
        amb_a, amb_d = assign_to_nearest_center(_take(X, up_this), new_medoids, metric)
        new_assig[up_this] = amb_a
        new_dist[up_this] = amb_d

        old_cost, new_cost = msq(distances), msq(new_dist)
        accepted = bool(new_cost < old_cost)
        if log is not None:
            log.append((cid, int(prop_ind), float(old_cost), float(new_cost), accepted))
        if accepted:
            distances, assignments = new_dist, new_assig
            medoid_coords = new_medoids
            medoid_inds[cid] = prop_ind
    return medoid_inds, distances, assignments, medoid_coords


def kmedoids_iterations(X, metric, n_iters, center_inds, assignments, distances,
                        proposals=None, random_state=None):
    """kmedoids.py:410-476: ``random_state`` is handed to every sweep unchanged, so an int
    re-seeds each sweep while a RandomState object carries over."""
    centers = None
    for _ in range(n_iters):
        center_inds, distances, assignments, centers = pam_update(
            X, metric, center_inds, assignments, distances, proposals=proposals,
            random_state=random_state)
    return Result(center_inds, distances, assignments, centers)


def hybrid(X, metric, n_iters=5, n_clusters=np.inf, dist_cutoff=0.0, random_state=None):
    """cluster/hybrid.py:112-162."""
    r = kcenters(X, metric, n_clusters=n_clusters, dist_cutoff=dist_cutoff)
    if n_iters > 0:
        return kmedoids_iterations(X, metric, n_iters, r.center_indices, r.assignments,
                                   r.distances, random_state=random_state)
    return r
