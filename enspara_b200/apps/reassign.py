"""`reassign` app: given cluster centres, re-assign trajectories in batches
(/root/reference/enspara/apps/reassign.py).  Same flags and outputs; the work runs through
``enspara_b200.cluster.reassign`` (loader thread + pinned staging + many-centres RMSD kernels).

Centres: the pickle written by the `cluster` app (list of 1-frame trajectories / arrays) or a
``.npy`` array of shape (k, n_atoms, 3).  Trajectories: ``.npy`` / mdtraj ``.h5`` are read
natively; other formats and atom-selection strings need mdtraj.
"""
import argparse
import logging
import os
import pickle
import sys
import time

import numpy as np

from .. import exception, ra
from ..cluster import reassign as rz

logger = logging.getLogger(__name__)


def process_command_line(argv):
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--centers", required=True,
                        help="Center structures (pickle or .npy) to use for reassignment.")
    parser.add_argument("--trajectories", required=True, nargs="+", action="append",
                        help="The aligned trajectory files to reassign.")
    parser.add_argument("--topology", required=True, action="append", dest="topologies",
                        help="The topology file for the trajectories.")
    parser.add_argument("--atoms", default="(name CA or name C or name N or name CB)",
                        help="The atoms from the trajectories (MDTraj atom-selection syntax, "
                             "or 'all') to reassign based upon.")
    parser.add_argument("--output-path", default=None)
    parser.add_argument("-m", "--mem-fraction", default=0.5, type=float,
                        help="The fraction of total RAM to use in deciding the batch size.")
    parser.add_argument("--distances", required=True,
                        help="Path to h5 file where distance to nearest cluster center will "
                             "be output.")
    parser.add_argument("--assignments", required=True,
                        help="Path to h5 file where assignments to nearest center will be "
                             "output.")
    args = parser.parse_args(argv[1:])

    if args.mem_fraction >= 1 or args.mem_fraction <= 0:
        raise exception.ImproperlyConfigured(
            "Flag --mem-fraction must be in range (0, 1). Got %s" % args.mem_fraction)
    if len(args.topologies) != len(args.trajectories):
        raise exception.ImproperlyConfigured(
            "The number of --topology and --trajectory flags must agree.")
    if args.output_path is None:
        args.output_path = os.path.dirname(args.centers)
    for trjset in args.trajectories:
        for trj in trjset:
            with open(trj, "r"):
                pass
    return args


def load_centers(path, atoms):
    """Centres file -> (k, n_atoms, 3) float32 (apps/reassign.py:106-110; the atom selection is
    applied to trajectory-like centres that carry a topology)."""
    if os.path.splitext(path)[1].lower() == ".npy":
        return rz._centers_xyz(np.load(path))
    with open(path, "rb") as f:
        centers = pickle.load(f)
    if hasattr(centers, "xyz"):
        centers = [centers[i] for i in range(len(centers))]
    out = []
    for c in centers:
        if hasattr(c, "top") and c.top is not None and hasattr(c.top, "select") \
                and isinstance(atoms, str) and atoms.strip() != "all":
            c = c.atom_slice(c.top.select(atoms))
        out.append(c)
    return rz._centers_xyz(out)


def main(argv=None):
    argv = sys.argv if argv is None else argv
    logging.basicConfig(level=logging.INFO,
                        format="%(asctime)s %(name)-8s %(levelname)-7s %(message)s",
                        datefmt="%m-%d-%Y %H:%M:%S")
    args = process_command_line(argv)
    tick = time.perf_counter()
    centers = load_centers(args.centers, args.atoms)
    logger.info('Loaded %s centers with %s atoms using selection "%s" in %.1f seconds.',
                len(centers), centers.shape[1], args.atoms, time.perf_counter() - tick)
    assig, dist = rz.reassign(args.topologies, args.trajectories,
                              [args.atoms] * len(args.topologies), centers=centers,
                              frac_mem=args.mem_fraction)
    logger.info("Finished reassignments in %.1f seconds.", time.perf_counter() - tick)
    ra.save(args.distances, dist)
    ra.save(args.assignments, assig)
    logger.info("Wrote distances at %s.", args.distances)
    logger.info("Wrote assignments at %s.", args.assignments)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
