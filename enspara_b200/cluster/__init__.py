"""Clustering estimators and functions, laid out like the reference package
(/root/reference/enspara/cluster/__init__.py): sub-modules ``kcenters``, ``kmedoids``,
``hybrid``, ``util`` plus the three estimator classes."""
from . import util  # noqa: F401
from . import kcenters  # noqa: F401
from . import kmedoids  # noqa: F401
from . import hybrid  # noqa: F401
from .hybrid import KHybrid  # noqa: F401
from .kcenters import KCenters  # noqa: F401
from .kmedoids import KMedoids  # noqa: F401

__all__ = ["KCenters", "KHybrid", "KMedoids", "kcenters", "kmedoids", "hybrid", "util"]
