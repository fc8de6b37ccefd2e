"""`enspara <app> ...` style entry point for the two apps on the clustering path
(/root/reference/enspara/apps/main.py:6-62; its third app, implied timescales, is downstream
of clustering and out of scope).  ``python -m enspara_b200.apps.main cluster --features ...``."""
import argparse
import sys


def identify_app(argv):
    parser = argparse.ArgumentParser(
        prog="enspara_b200", formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        description="Main entry point for the enspara_b200 apps.")
    parser.add_argument("appname", choices=["cluster", "reassign"],
                        help="Name of the application.")
    parser.add_argument("appargs", nargs=argparse.REMAINDER,
                        help="Subsequent arguments to the app (add subcommand for more).")
    helpstack = []
    for h in ("--help", "-h"):
        while h in argv and argv.index(h) != 1:
            argv.remove(h)
            helpstack.append(h)
    args = parser.parse_args(argv[1:])
    if args.appname == "cluster":
        from .cluster import main
    else:
        from .reassign import main
    args.main = main
    args.appargs.extend(helpstack)
    return args


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    args = identify_app(argv)
    # the apps expect argv[0] to be a program name, like the reference (main.py:49)
    args.main([args.appname] + args.appargs)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
