"""One entry point for the apps on the clustering path: ``python -m enspara_b200.apps.main
cluster ...`` / ``... reassign ...`` (the reference exposes the same two, plus a timescales app
that is downstream of clustering and out of scope, through its `enspara` console script,
/root/reference/enspara/apps/main.py)."""
import importlib
import sys

APPS = {"cluster": "enspara_b200.apps.cluster", "reassign": "enspara_b200.apps.reassign"}


def _usage(stream):
    stream.write("usage: enspara_b200 {%s} [app arguments]\n"
                 "       enspara_b200 <app> --help   for the app's own options\n"
                 % ",".join(sorted(APPS)))


def main(argv=None):
    """Dispatch ``argv[1]`` to the app of that name; the app receives ``[name] + rest`` so that
    its argument parser sees a program name first, like the reference's apps do."""
    argv = list(sys.argv if argv is None else argv)
    if len(argv) < 2 or argv[1] in ("-h", "--help"):
        _usage(sys.stdout)
        return 0 if len(argv) >= 2 else 2
    name, rest = argv[1], argv[2:]
    if name not in APPS:
        _usage(sys.stderr)
        sys.stderr.write("unknown app %r\n" % name)
        raise SystemExit(2)
    app_main = importlib.import_module(APPS[name]).main
    rc = app_main([name] + rest)
    return 0 if rc is None else rc


if __name__ == "__main__":
    sys.exit(main())
