"""CPU oracle for the clustering hot path -- TEST INFRASTRUCTURE, never imported by the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package (and there only as the checker or the timed
CPU baseline).  ``enspara_b200`` must never import it.

Contents
--------
enspara_oracle.c   C restatement of mdtraj.rmsd (absent third-party dep) and of enspara's
                   Cython libdist, see the header of that file for citations.
distances.py       ctypes bindings + numpy-facing metric callables (``rmsd``, ``euclidean`` ...).
cluster.py         numpy restatement of the reference's k-centers / PAM / assign loops.
xtc.py             minimal XTC reader used once to turn the reference's frame0.xtc test fixture
                   into tests/golden/frame0_xyz.npy.
refharness.py      imports the REAL reference package from /root/reference (only in the build
                   container) with mdtraj/tables stubbed, to validate the restatement and to
                   generate golden fixtures.
"""
