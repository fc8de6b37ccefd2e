"""Build libenspara_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The shared library is plain C ABI (include/enspara_b200.h); PyTorch is not involved in the
build.  ``python -m enspara_b200.build`` or ``__graft_entry__.build()`` runs this.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libenspara_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "-Xptxas", "-v",
    "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(os.path.dirname(HERE), "include", "enspara_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ (one nvcc per file, in parallel) and link them into
    enspara_b200/libenspara_b200.so."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + compile_flags + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        return obj, " ".join(cmd) + "\n" + res.stdout + res.stderr, res.returncode

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    log = "".join(r[1] for r in results)
    ok = all(r[2] == 0 for r in results)
    if ok:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler",
               "-fPIC", "-o", LIB] + [r[0] for r in results]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        ok = res.returncode == 0
    with open(os.path.join(HERE, "csrc", "build.log"), "w") as fh:
        fh.write(log)
    if not ok:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libenspara_b200.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
