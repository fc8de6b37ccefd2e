"""Clustering estimators and functions with the reference's names
(/root/reference/enspara/cluster/__init__.py)."""
from . import util
from .kcenters import KCenters, kcenters, kcenters_mpi

__all__ = ["KCenters", "kcenters", "kcenters_mpi", "util"]
