"""`cluster` command-line app with the reference's flags
(/root/reference/enspara/apps/cluster.py:69-377).  Run as
``python -m enspara_b200.apps.cluster ...`` (one process), or under
``torchrun --nproc-per-node N -m enspara_b200.apps.cluster ...`` (one process per GPU; input
files are striped over ranks exactly like the reference does under mpirun).
"""
import argparse
import logging
import os
import sys

import numpy as np

from .. import exception, mpi
from ..cluster import KCenters, KHybrid, KMedoids, util
from ..cluster import io as cio
from ..util.log import timed

logger = logging.getLogger(__name__)


class readable_dir(argparse.Action):
    """The directory of the given path must exist and be readable (apps/util.py:5-20)."""

    def __call__(self, parser, namespace, values, option_string=None):
        d = os.path.dirname(os.path.abspath(values))
        if not os.path.isdir(d):
            raise argparse.ArgumentTypeError("readable_dir:{0} is not a valid path".format(d))
        if not os.access(d, os.R_OK):
            raise argparse.ArgumentTypeError("readable_dir:{0} is not a readable dir".format(d))
        setattr(namespace, self.dest, values)


def process_command_line(argv):
    FEATURE_DISTANCES = ["euclidean", "manhattan"]
    TRAJECTORY_DISTANCES = ["rmsd"]
    ALGORITHMS = {"kcenters": KCenters, "khybrid": KHybrid, "kmedoids": KMedoids}

    parser = argparse.ArgumentParser(
        prog="cluster", formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        description="Cluster a set (or several sets) of trajectories into a single state "
                    "space based upon RMSD.")
    input_args = parser.add_argument_group("Input Settings")
    grp = parser.add_mutually_exclusive_group(required=True)
    grp.add_argument("--features", nargs="+",
                     help="The h5 file (or npy files) containing observations and features.")
    grp.add_argument("--trajectories", nargs="+", action="append",
                     help="List of paths to aligned trajectory files to cluster.")
    input_args.add_argument("--topology", action="append", dest="topologies",
                            help="The topology file for the trajectories (once per "
                                 "--trajectories flag).")
    c = parser.add_argument_group("Clustering Settings")
    c.add_argument("--algorithm", required=True, choices=["khybrid", "kcenters", "kmedoids"])
    c.add_argument("--atoms", action="append",
                   help="MDTraj atom selection to cluster on (once, or once per topology).")
    c.add_argument("--cluster-radius", default=None, type=float)
    c.add_argument("--cluster-number", default=None, type=int)
    c.add_argument("--cluster-distance", default=None,
                   choices=FEATURE_DISTANCES + TRAJECTORY_DISTANCES)
    c.add_argument("--cluster-iterations", default=None, type=int)
    c.add_argument("--save_intermediates", default=False, type=bool)
    c.add_argument("--init-center-inds", default=None, type=str)
    c.add_argument("--init-assignments", default=None, type=str)
    c.add_argument("--init-distances", default=None, type=str)
    c.add_argument("--subsample", default=1, type=int)
    o = parser.add_argument_group("Output Settings")
    o.add_argument("--no-reassign", default=False, action="store_true")
    o.add_argument("--distances", required=True, action=readable_dir)
    o.add_argument("--center-features", required=True, action=readable_dir)
    o.add_argument("--assignments", required=True, action=readable_dir)
    o.add_argument("--center-indices", required=False, action=readable_dir)

    args = parser.parse_args(argv[1:])

    if args.features:
        args.features = cio.expand_files([args.features])[0]
        if args.cluster_distance in FEATURE_DISTANCES:
            args.cluster_distance = util._get_distance_method(args.cluster_distance)
        else:
            raise exception.ImproperlyConfigured(
                "The given distance (%s) is not compatible with features."
                % args.cluster_distance)
        if args.subsample != 1 and len(args.features) == 1:
            raise exception.ImproperlyConfigured("Subsampling is not supported for h5 inputs.")
        if args.topologies:
            raise exception.ImproperlyConfigured(
                "When --features is specified, --topology is unneccessary.")
        if args.atoms:
            raise exception.ImproperlyConfigured(
                "Option --atoms is only meaningful when clustering trajectories.")
    elif args.trajectories and args.topologies:
        args.trajectories = cio.expand_files(args.trajectories)
        if not args.cluster_distance or args.cluster_distance == "rmsd":
            args.cluster_distance = util.RMSD
        else:
            raise exception.ImproperlyConfigured(
                "Option --cluster-distance must be rmsd when clustering trajectories.")
        if not args.atoms:
            raise exception.ImproperlyConfigured(
                "Option --atoms is required when clustering trajectories.")
        elif len(args.atoms) == 1:
            args.atoms = args.atoms * len(args.trajectories)
        elif len(args.atoms) != len(args.trajectories):
            raise exception.ImproperlyConfigured(
                "Flag --atoms must be provided either once (selection is applied to all "
                "trajectories) or the same number of times --trajectories is supplied.")
        if len(args.topologies) != len(args.trajectories):
            raise exception.ImproperlyConfigured(
                "The number of --topology and --trajectory flags must agree.")
    else:
        raise exception.ImproperlyConfigured(
            "Either --features or both of --trajectories and --topologies are required.")

    if args.cluster_radius is None and args.cluster_number is None:
        raise exception.ImproperlyConfigured(
            "At least one of --cluster-radius and --cluster-number is required to cluster.")

    args.Clusterer = ALGORITHMS[args.algorithm]
    if args.Clusterer is KCenters and args.cluster_iterations is not None:
        raise exception.ImproperlyConfigured(
            "--cluster-iterations only has an effect when using an interative clustering "
            "scheme (e.g. khybrid).")
    if args.Clusterer is KMedoids:
        if args.cluster_radius is not None:
            raise exception.ImproperlyConfigured(
                "--cluster-radius only has an effect when using kcenters or khybrid.")
    else:
        for name in (args.init_center_inds, args.init_distances, args.init_assignments):
            if name:
                raise exception.ImproperlyConfigured(
                    "--init-center-inds, --init-distances, and --init-assignments are only "
                    "implemented for kmedoids")
    if args.no_reassign and args.subsample == 1:
        logger.warning("When subsampling is 1 (or unspecified), --no-reassign has no effect.")
    if args.subsample != 1 and not args.no_reassign and args.features:
        # the reference can only re-assign trajectory inputs (cluster/util.py:531-534 needs
        # topologies); with features it would crash after clustering -- say so up front
        logger.warning("Re-assignment of skipped frames needs trajectory inputs; proceeding "
                       "as with --no-reassign.")
        args.no_reassign = True
    if not args.no_reassign and args.subsample > 1 and mpi.size() > 1:
        logger.warning("Reassignment is suppressed in MPI mode.")   # apps/cluster.py:266-268
        args.no_reassign = True
    return args


def main(argv=None):
    argv = sys.argv if argv is None else argv
    mpi.init_from_env()
    mpi_mode = mpi.size() > 1
    fmt = "%(asctime)s " + ("[Rank %s] " % mpi.rank() if mpi_mode else "") + \
        "%(name)-8s %(levelname)-7s %(message)s"
    logging.basicConfig(level=logging.INFO, format=fmt, datefmt="%m-%d-%Y %H:%M:%S")

    args = process_command_line(argv)

    # in sharded mode lengths are global, data is this rank's (apps/cluster.py:291-293)
    lengths, data = cio.load_trjs_or_features(args)

    kwargs = {}
    if args.cluster_iterations is not None:
        if args.Clusterer is KHybrid:
            kwargs["kmedoids_updates"] = int(args.cluster_iterations)
        elif args.Clusterer is KMedoids:
            kwargs["n_iters"] = int(args.cluster_iterations)
        if args.Clusterer is not KCenters:
            kwargs["args"] = args
            kwargs["lengths"] = lengths
    if args.cluster_radius is not None:
        kwargs["cluster_radius"] = args.cluster_radius
    if args.Clusterer is not KMedoids:
        kwargs["mpi_mode"] = mpi_mode

    clustering = args.Clusterer(metric=args.cluster_distance, n_clusters=args.cluster_number,
                                **kwargs)
    if args.Clusterer is KMedoids:
        restart = {}
        if args.init_distances:
            _, restart["distances"] = cio.load_features([args.init_distances], 1)
        if args.init_assignments:
            restart["X_lengths"], restart["assignments"] = cio.load_features(
                [args.init_assignments], 1)
        if args.init_center_inds:
            restart["cluster_center_inds"] = np.load(args.init_center_inds)
        clustering.fit(data, **restart)
    else:
        clustering.fit(data)
    del data

    logger.info("Clustered %s frames into %s clusters in %s seconds.", sum(lengths),
                len(clustering.centers_), clustering.runtime_)

    result = clustering.result_
    if mpi_mode:
        with timed("Reassembled dist and assign arrays in %.2f sec", logger.info):
            all_dists = mpi.ops.assemble_striped_ragged_array(result.distances, lengths)
            all_assigs = mpi.ops.assemble_striped_ragged_array(result.assignments, lengths)
            ctr_inds = mpi.ops.convert_local_indices(result.center_indices, lengths)
        result = util.ClusterResult(center_indices=ctr_inds, distances=all_dists,
                                    assignments=all_assigs, centers=result.centers)
    result = result.partition(lengths)

    if mpi.rank() == 0:
        with timed("Wrote center indices in %.2f sec.", logger.info):
            cio.write_centers_indices(
                args.center_indices,
                [(t, f * args.subsample) for t, f in result.center_indices])
        with timed("Wrote center structures in %.2f sec.", logger.info):
            cio.write_centers(result, args)
        cio.write_assignments_and_distances(result, args)
    mpi.comm.barrier()
    logger.info("Success! Data can be found in %s.", os.path.dirname(args.distances))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
