"""Minimal ragged-array support for ``ClusterResult.partition``.

The reference's ``enspara.ra`` (854 lines, HDF5 I/O, fancy indexing) is host bookkeeping and
out of scope; the clustering path needs only: a flat buffer cut into rows
by ``lengths`` with row access, plus ``partition_list`` / ``partition_indices``
(/root/reference/enspara/ra/ra.py:223-242, 361-376).
"""
import numpy as np

from .exception import DataInvalid


class RaggedArray:
    """Rows of different length over one flat buffer (row access, iteration, equality)."""

    def __init__(self, array, lengths=None, copy=True, error_checking=True):
        if lengths is None:
            rows = [np.asarray(r) for r in array]
            lengths = [len(r) for r in rows]
            array = np.concatenate(rows) if rows else np.zeros(0)
        self._data = np.array(array, copy=copy)
        self.lengths = np.asarray(lengths, dtype=np.int64)
        if error_checking and int(self.lengths.sum()) != len(self._data):
            raise DataInvalid(
                "Sum of lengths (%d) does not match the number of elements (%d)."
                % (int(self.lengths.sum()), len(self._data)))
        self.starts = np.concatenate([[0], np.cumsum(self.lengths)[:-1]]).astype(np.int64)

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def shape(self):
        return (len(self.lengths), None)

    def __len__(self):
        return len(self.lengths)

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            if i < 0:
                i += len(self.lengths)
            s = int(self.starts[i])
            return self._data[s:s + int(self.lengths[i])]
        if isinstance(i, slice):
            rows = range(*i.indices(len(self.lengths)))
            return RaggedArray([self[r] for r in rows])
        raise TypeError("RaggedArray supports int and slice row indexing only")

    def __iter__(self):
        for i in range(len(self.lengths)):
            yield self[i]

    def flatten(self):
        return self._data

    def __eq__(self, other):
        if isinstance(other, RaggedArray):
            return (np.array_equal(self.lengths, other.lengths)
                    and np.array_equal(self._data, other._data))
        return NotImplemented


def partition_list(list_to_partition, partition_lengths):
    """Cut a concatenated sequence into consecutive pieces (ra.py:361-376)."""
    if np.sum(partition_lengths) != len(list_to_partition):
        raise DataInvalid(
            "Number of elements in list (%d) does not equal the sum of the lengths to "
            "partition (%d)" % (len(list_to_partition), np.sum(partition_lengths)))
    out, start = [], 0
    for n in partition_lengths:
        out.append(list_to_partition[start:start + n])
        start += n
    return out


def partition_indices(indices, traj_lengths):
    """Concatenated index -> (trajectory, frame) pairs (ra.py:223-242)."""
    bounds = np.concatenate([[0], np.cumsum(np.asarray(traj_lengths, dtype=np.int64))])
    out = []
    for index in indices:
        t = int(np.searchsorted(bounds, index, side="right") - 1)
        if 0 <= t < len(traj_lengths):
            out.append((t, index - int(bounds[t])))
    return out


# ---------------------------------------------------------------------------------------------
# save / load.  The reference writes one zlib CArray ``arr_<i>`` per trajectory with PyTables
# (ra.py:45-89) and needs PyTables to read them back.  PyTables is optional here: `.h5` paths
# use it when importable, `.npy` / `.npz` paths work everywhere.
# ---------------------------------------------------------------------------------------------
def _rows(obj):
    if isinstance(obj, RaggedArray):
        return [obj[i] for i in range(len(obj))], False
    arr = np.asarray(obj)
    if arr.ndim >= 2:
        return [arr[i] for i in range(arr.shape[0])], True
    return [arr], True


def save(path, obj):
    """Write a (ragged) array.  `.h5`: PyTables layout of the reference; `.npy`: dense array
    (rows of equal length only); `.npz`: flat data + lengths."""
    import os
    ext = os.path.splitext(path)[1].lower()
    rows, square = _rows(obj)
    if ext in (".h5", ".hdf5"):
        try:
            import tables
        except ImportError:
            from .exception import ImproperlyConfigured
            raise ImproperlyConfigured(
                "Writing '%s' needs PyTables (as in the reference); it is not installed. "
                "Use a .npy / .npz path instead." % path)
        compression = tables.Filters(complevel=9, complib="zlib", shuffle=True)
        with tables.open_file(path, mode="w") as handle:
            if square and not isinstance(obj, RaggedArray):
                arr = np.asarray(obj)
                atom = tables.Atom.from_dtype(arr.dtype)
                node = handle.create_carray(where="/", name="array", atom=atom,
                                            shape=arr.shape, filters=compression)
                node[:] = arr
            else:
                n_zeros = len(str(len(rows))) + 1
                for i, row in enumerate(rows):
                    row = np.asarray(row)
                    atom = tables.Atom.from_dtype(row.dtype)
                    node = handle.create_carray(
                        where="/", name="array_" + str(i).zfill(n_zeros), atom=atom,
                        shape=row.shape, filters=compression)
                    node[:] = row
        return
    if ext == ".npz" or not square or isinstance(obj, RaggedArray):
        flat = np.concatenate([np.asarray(r).reshape(-1) for r in rows]) if rows else np.zeros(0)
        np.savez(path if ext == ".npz" else path + ".npz", data=flat,
                 lengths=np.array([len(r) for r in rows], dtype=np.int64))
        return
    np.save(path, np.asarray(obj))


def load(path):
    """Inverse of ``save`` for the .npy / .npz forms (and .h5 when PyTables is present)."""
    import os
    ext = os.path.splitext(path)[1].lower()
    if ext in (".h5", ".hdf5"):
        import tables
        with tables.open_file(path) as handle:
            names = sorted(n.name for n in handle.list_nodes("/"))
            if names == ["array"]:
                return handle.get_node("/array")[:]
            rows = [handle.get_node("/" + n)[:] for n in names]
        return RaggedArray(rows)
    if ext == ".npz":
        z = np.load(path)
        return RaggedArray(z["data"], lengths=z["lengths"])
    return np.load(path)
