"""Trajectory input without mdtraj: a PDB topology with an atom-selection mini-language, a
light Trajectory container and loaders for .xtc (native decoder in libenspara_b200.so,
csrc/eb_xtc.cu), .npy and mdtraj-style .h5.

The reference loads trajectories with ``md.load(file, top=, stride=, atom_indices=)`` and picks
atoms with ``top.select(<mdtraj selection string>)`` (/root/reference/enspara/cluster/util.py:
350-404, apps/cluster.py:206-229).  mdtraj is not installable in this image, so the loaders
either side of the hot path get their own readers for the formats the reference's tests use
(``frame0.xtc`` + ``native.pdb``, ``beta-peptide.xtc``); when mdtraj IS importable it is
preferred for every other format.  Only what the clustering path touches is provided: ``.xyz``,
``.top`` / ``.topology``, ``len()``, integer / slice / mask / list indexing, ``n_atoms``,
``n_frames``, and ``type(traj)(xyz=, topology=)`` (mpi/ops.py:208-210).

Selection language (the subset of mdtraj's that the reference's CLI tests and docs use):
``all`` | ``none`` | ``backbone`` | ``sidechain`` | ``protein`` | ``water`` |
``name CA`` | ``resname ALA`` | ``residue 5`` (= ``resSeq``) | ``resid 3`` (0-based residue
index) | ``index 10`` | ``element C`` (= ``symbol``) | ``chainid 0`` | numeric keywords also
take ``== != < <= > >=`` and ``a to b``; combined with ``and`` / ``or`` / ``not`` and
parentheses.  Anything else raises ``ValueError`` (the callers turn it into the reference's
``ImproperlyConfigured``).
"""
import ctypes
import os
import re

import numpy as np

from .. import _lib
from ..exception import DataInvalid, ImproperlyConfigured

_BACKBONE = {"N", "CA", "C", "O"}
_PROTEIN = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS",
            "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL", "HID", "HIE", "HIP", "HSD",
            "HSE", "HSP", "CYX", "ASH", "GLH", "LYN", "ACE", "NME", "NH2", "MSE"}
_WATER = {"HOH", "WAT", "SOL", "H2O", "TIP", "TIP3", "TIP4", "SPC"}


class Topology:
    """Per-atom arrays of a PDB file (first model)."""

    def __init__(self, names, resnames, resseqs, chains, elements, serials=None):
        self.names = np.asarray(names, dtype=object)
        self.resnames = np.asarray(resnames, dtype=object)
        self.resseqs = np.asarray(resseqs, dtype=np.int64)
        self.chains = np.asarray(chains, dtype=object)
        self.elements = np.asarray(elements, dtype=object)
        self.serials = np.arange(len(self.names)) if serials is None else np.asarray(serials)
        n = len(self.names)
        # residue index: a new residue starts when (chain, resSeq, resname) changes
        rid = np.zeros(n, dtype=np.int64)
        cid = np.zeros(n, dtype=np.int64)
        chain_ids = {}
        for i in range(n):
            if i:
                same = (self.chains[i] == self.chains[i - 1]
                        and self.resseqs[i] == self.resseqs[i - 1]
                        and self.resnames[i] == self.resnames[i - 1])
                rid[i] = rid[i - 1] + (0 if same else 1)
            cid[i] = chain_ids.setdefault(self.chains[i], len(chain_ids))
        self.resids, self.chainids = rid, cid

    @property
    def n_atoms(self):
        return len(self.names)

    def __len__(self):
        return len(self.names)

    def __eq__(self, other):
        return (isinstance(other, Topology) and self.n_atoms == other.n_atoms
                and bool(np.all(self.names == other.names))
                and bool(np.all(self.resnames == other.resnames))
                and bool(np.all(self.resseqs == other.resseqs)))

    def __hash__(self):
        return hash((self.n_atoms, tuple(self.names[:8])))

    def subset(self, atom_indices):
        idx = np.asarray(atom_indices, dtype=np.int64)
        return Topology(self.names[idx], self.resnames[idx], self.resseqs[idx], self.chains[idx],
                        self.elements[idx], self.serials[idx])

    def select(self, expression):
        """Indices (ascending, int64) of the atoms matching ``expression``."""
        mask = _Selection(self, expression).evaluate()
        return np.nonzero(mask)[0].astype(np.int64)


# ---------------------------------------------------------------------------------------------
# selection mini-language
# ---------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"\s*(\(|\)|==|!=|<=|>=|<|>|[^\s()<>=!]+)")
_STRING_KEYS = {"name": "names", "resname": "resnames", "resn": "resnames",
                "element": "elements", "symbol": "elements", "type": "elements"}
_NUMERIC_KEYS = {"residue": "resseqs", "resseq": "resseqs", "resi": "resseqs",
                 "resid": "resids", "index": None, "chainid": "chainids"}
_FLAGS = {"all", "everything", "none", "nothing", "backbone", "is_backbone", "sidechain",
          "is_sidechain", "protein", "is_protein", "water", "is_water", "waters"}


class _Selection:
    def __init__(self, top, text):
        if not isinstance(text, str) or not text.strip():
            raise ValueError("empty atom selection")
        self.top = top
        self.toks = _TOKEN.findall(text)
        if "".join(self.toks).replace(" ", "") != re.sub(r"\s+", "", text):
            raise ValueError("cannot tokenise atom selection %r" % text)
        self.pos = 0

    def _peek(self):
        return self.toks[self.pos] if self.pos < len(self.toks) else None

    def _next(self):
        t = self._peek()
        if t is None:
            raise ValueError("unexpected end of atom selection")
        self.pos += 1
        return t

    def evaluate(self):
        m = self._or()
        if self._peek() is not None:
            raise ValueError("unexpected token %r in atom selection" % self._peek())
        return m

    def _or(self):
        m = self._and()
        while self._peek() is not None and self._peek().lower() in ("or", "||"):
            self._next()
            m = m | self._and()
        return m

    def _and(self):
        m = self._not()
        while self._peek() is not None and self._peek().lower() in ("and", "&&"):
            self._next()
            m = m & self._not()
        return m

    def _not(self):
        if self._peek() is not None and self._peek().lower() in ("not", "!"):
            self._next()
            return ~self._not()
        return self._atom()

    def _atom(self):
        t = self._next()
        if t == "(":
            m = self._or()
            if self._next() != ")":
                raise ValueError("missing ')' in atom selection")
            return m
        key = t.lower()
        n = self.top.n_atoms
        if key in _FLAGS:
            if key in ("all", "everything"):
                return np.ones(n, dtype=bool)
            if key in ("none", "nothing"):
                return np.zeros(n, dtype=bool)
            prot = np.array([r in _PROTEIN for r in self.top.resnames], dtype=bool)
            bb = np.array([a in _BACKBONE for a in self.top.names], dtype=bool)
            if key in ("backbone", "is_backbone"):
                return prot & bb
            if key in ("sidechain", "is_sidechain"):
                return prot & ~bb
            if key in ("protein", "is_protein"):
                return prot
            return np.array([r in _WATER for r in self.top.resnames], dtype=bool)
        if key in _STRING_KEYS:
            arr = getattr(self.top, _STRING_KEYS[key])
            op = "=="
            if self._peek() in ("==", "!="):
                op = self._next()
            val = self._next().strip("'\"")
            if val in ("(", ")") or val.lower() in ("and", "or", "not"):
                raise ValueError("missing value after %r in atom selection" % t)
            m = np.array([a == val for a in arr], dtype=bool)
            return m if op == "==" else ~m
        if key in _NUMERIC_KEYS:
            attr = _NUMERIC_KEYS[key]
            arr = np.arange(n) if attr is None else getattr(self.top, attr)
            op = "=="
            if self._peek() in ("==", "!=", "<", "<=", ">", ">="):
                op = self._next()
            lo = self._number()
            if op == "==" and self._peek() is not None and self._peek().lower() == "to":
                self._next()
                hi = self._number()
                return (arr >= lo) & (arr <= hi)
            return {"==": arr == lo, "!=": arr != lo, "<": arr < lo, "<=": arr <= lo,
                    ">": arr > lo, ">=": arr >= lo}[op]
        raise ValueError("unknown keyword %r in atom selection" % t)

    def _number(self):
        t = self._next()
        if not re.fullmatch(r"\d+", t):      # like mdtraj's grammar: no signed literals
            raise ValueError("expected a non-negative integer in atom selection, got %r" % t)
        return int(t)


def load_pdb_topology(path):
    """Atoms of the first model of a PDB file (ATOM / HETATM records, fixed columns)."""
    names, resnames, resseqs, chains, elements, serials = [], [], [], [], [], []
    with open(path) as fh:
        for line in fh:
            rec = line[:6]
            if rec in ("ATOM  ", "HETATM"):
                name = line[12:16].strip()
                names.append(name)
                resnames.append(line[17:21].strip())
                chains.append(line[21:22])
                try:
                    resseqs.append(int(line[22:26]))
                except ValueError:
                    resseqs.append(int(line[22:27].strip() or 0))
                el = line[76:78].strip() if len(line) >= 78 else ""
                if not el:
                    stripped = name.lstrip("0123456789")
                    el = stripped[:1] if stripped else ""
                elements.append(el.capitalize())
                try:
                    serials.append(int(line[6:11]))
                except ValueError:
                    serials.append(len(serials) + 1)
            elif rec.startswith("ENDMDL"):
                break
    if not names:
        raise DataInvalid("no ATOM / HETATM records in '%s'" % path)
    return Topology(names, resnames, resseqs, chains, elements, serials)


def load_pdb_xyz(path):
    """(1, n_atoms, 3) float32 coordinates of the first model, in nm (PDB is in Angstrom)."""
    xyz = []
    with open(path) as fh:
        for line in fh:
            if line[:6] in ("ATOM  ", "HETATM"):
                xyz.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
            elif line.startswith("ENDMDL"):
                break
    return (np.asarray(xyz, dtype=np.float32) / np.float32(10.0))[None]


# ---------------------------------------------------------------------------------------------
# container
# ---------------------------------------------------------------------------------------------
class Trajectory:
    """What the clustering path touches of ``mdtraj.Trajectory``."""

    def __init__(self, xyz, topology=None):
        xyz = np.asarray(xyz, dtype=np.float32)
        if xyz.ndim == 2:
            xyz = xyz[None]
        if xyz.ndim != 3 or xyz.shape[2] != 3:
            raise DataInvalid("coordinates must have shape (n_frames, n_atoms, 3), got %s"
                              % (xyz.shape,))
        self.xyz = xyz
        self.top = self.topology = topology

    @property
    def n_frames(self):
        return self.xyz.shape[0]

    @property
    def n_atoms(self):
        return self.xyz.shape[1]

    def __len__(self):
        return self.xyz.shape[0]

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return Trajectory(self.xyz[int(key)][None].copy(), self.top)
        return Trajectory(self.xyz[key], self.top)

    def __repr__(self):
        return "<enspara_b200 Trajectory with %d frames, %d atoms>" % (self.n_frames,
                                                                       self.n_atoms)

    def atom_slice(self, atom_indices):
        idx = np.asarray(atom_indices, dtype=np.int64)
        return Trajectory(np.ascontiguousarray(self.xyz[:, idx]),
                          self.top.subset(idx) if self.top is not None else None)


def join(trajs):
    """Concatenate trajectories frame-wise (``md.join``)."""
    trajs = list(trajs)
    return Trajectory(np.concatenate([t.xyz for t in trajs]), trajs[0].top)


# ---------------------------------------------------------------------------------------------
# files
# ---------------------------------------------------------------------------------------------
def xtc_shape(path):
    """(n_frames, n_atoms) of an .xtc file; headers only."""
    nf, na = ctypes.c_int64(), ctypes.c_int32()
    _lib.check(_lib.load().eb_xtc_scan(os.fsencode(path), ctypes.byref(nf), ctypes.byref(na)))
    return int(nf.value), int(na.value)


def read_xtc(path, stride=1, atom_indices=None, out=None, first=0, max_frames=None):
    """Frames first, first+stride, ... (at most ``max_frames``) of an .xtc file as
    (n, n_sel, 3) float32 nm, decoded natively (csrc/eb_xtc.cu).  ``out``: preallocated
    destination (e.g. a pinned buffer)."""
    nf, na = xtc_shape(path)
    n_out = max(0, (nf - first + stride - 1) // stride)
    if max_frames is not None:
        n_out = min(n_out, int(max_frames))
    sel = None
    n_sel = na
    if atom_indices is not None:
        sel = np.ascontiguousarray(np.asarray(atom_indices), dtype=np.int32)
        n_sel = len(sel)
        if n_sel == 0:
            raise DataInvalid("empty atom selection for '%s'" % path)
    if out is None:
        out = np.empty((n_out, n_sel, 3), dtype=np.float32)
    if out.shape != (n_out, n_sel, 3) or out.dtype != np.float32 or not out.flags.c_contiguous:
        raise DataInvalid("destination has shape %s, the file gives %s" % (out.shape,
                                                                       (n_out, n_sel, 3)))
    nr = ctypes.c_int64()
    _lib.check(_lib.load().eb_xtc_read(
        os.fsencode(path), int(first), int(stride), int(n_out),
        sel.ctypes.data_as(ctypes.c_void_p) if sel is not None else None, int(n_sel),
        out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nr)))
    if nr.value != n_out:
        raise DataInvalid("'%s' yielded %d frames, expected %d" % (path, nr.value, n_out))
    return out


def n_frames(path):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".xtc":
        return xtc_shape(path)[0]
    raise ImproperlyConfigured("no native reader for '%s'" % path)


def load_topology(path):
    ext = os.path.splitext(path)[1].lower()
    if ext in (".pdb", ".ent"):
        return load_pdb_topology(path)
    raise ImproperlyConfigured(
        "Topology file '%s': only .pdb is read without mdtraj, which is not installed." % path)


def load_frame(path, index, top=None, atom_indices=None):
    """``md.load_frame(path, index, top=)``: one frame as a 1-frame Trajectory."""
    ext = os.path.splitext(path)[1].lower()
    if isinstance(top, str):
        top = load_topology(top)
    if ext == ".xtc":
        xyz = read_xtc(path, first=int(index), max_frames=1, atom_indices=atom_indices)
        if len(xyz) != 1:
            raise DataInvalid("'%s' has no frame %d" % (path, index))
    elif ext == ".npy":
        xyz = np.asarray(np.load(path, mmap_mode="r")[int(index)], dtype=np.float32)[None]
    elif ext in (".h5", ".hdf5"):
        from . import h5min
        xyz = np.asarray(h5min.read(path, "coordinates")[int(index)], dtype=np.float32)[None]
    else:
        raise ImproperlyConfigured("no native reader for '%s'" % path)
    if atom_indices is not None and ext != ".xtc":
        xyz = xyz[:, np.asarray(atom_indices, dtype=np.int64)]
    if top is not None and atom_indices is not None:
        top = top.subset(atom_indices)
    return Trajectory(xyz, top)


def load(path, top=None, stride=1, atom_indices=None):
    """``md.load(path, top=, stride=, atom_indices=)`` for .xtc / .pdb without mdtraj."""
    ext = os.path.splitext(path)[1].lower()
    if isinstance(top, str):
        top = load_topology(top)
    if ext == ".xtc":
        xyz = read_xtc(path, stride=stride, atom_indices=atom_indices)
        if top is not None:
            if atom_indices is not None:
                top = top.subset(atom_indices)
            elif top.n_atoms != xyz.shape[1]:
                raise DataInvalid("topology has %d atoms, '%s' has %d" % (
                    top.n_atoms, path, xyz.shape[1]))
        return Trajectory(xyz, top)
    if ext in (".pdb", ".ent"):
        t = Trajectory(load_pdb_xyz(path), load_pdb_topology(path))
        return t.atom_slice(atom_indices) if atom_indices is not None else t
    raise ImproperlyConfigured(
        "Trajectory file '%s': only .xtc, .pdb, .npy and mdtraj .h5 are read without mdtraj, "
        "which is not installed." % path)
