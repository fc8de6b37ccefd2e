"""``enspara.geometry.libdist`` on the GPU: one point against many rows.

Same call signatures, checks and exceptions as the Cython module
(/root/reference/enspara/geometry/libdist.pyx:148-203); the arithmetic is reproduced bit for bit
by csrc/eb_feat.cu (typed difference, typed square, float64 accumulation in feature order,
double sqrt).  ``hamming`` is not reachable from the clustering path
(cluster/util.py:289-313 has no branch for it) and is not provided.
"""
import numpy as np

from ..exception import DataInvalid


def _check(X, y, out):
    X = np.asarray(X) if not hasattr(X, "shape") else X
    y = np.asarray(y) if not hasattr(y, "shape") else y
    if len(X.shape) != 2:
        raise DataInvalid("Data array dimension must be two, got shape %s." % str(X.shape))
    if len(y.shape) != 1:
        raise DataInvalid("Target point dimension must be one, got shape %s." % str(y.shape))
    if X.shape[1] != y.shape[0]:
        raise DataInvalid(("Target data point dimension (%s) must match data "
                           "array dimension (%s)") % (y.shape[0], X.shape[1]))
    if out is not None:
        if out.dtype != np.float64:
            raise DataInvalid("In-place output array must be np.float64, got '%s'." % out.dtype)
        if out.shape[0] != X.shape[0]:
            raise DataInvalid(("In-place output array dimension (%s) must match number of "
                               "samples in data array (%s)") % (out.shape[0], X.shape[0]))
        if len(out.shape) != 1:
            raise DataInvalid("In-place output array must be one-dimensional, got shape %s"
                              % (out.shape,))
    return X, y


def euclidean(X, y, out=None):
    """Euclidean distance between the point ``y`` and every row of ``X`` (float64[n])."""
    from ..cluster import util
    X, y = _check(X, y, out)
    return util.EUCLIDEAN(X, y, out=out)


def manhattan(X, y, out=None):
    """Manhattan distance between the point ``y`` and every row of ``X`` (float64[n])."""
    from ..cluster import util
    X, y = _check(X, y, out)
    return util.MANHATTAN(X, y, out=out)
