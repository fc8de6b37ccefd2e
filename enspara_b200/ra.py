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
# save / load.  The reference writes one zlib CArray per trajectory with PyTables, named
# ``arr_<i>`` (ra.py:45-89: tag 'arr', index zero-filled to len(str(n_rows)) + 1 digits; a
# plain ndarray becomes the single node ``arr_0``) and reads them back with ``ra.load``
# (ra.py:114-220).  PyTables is not needed here: `.h5` files go through the dependency-free
# reader/writer in ``enspara_b200.util.h5min`` (reads PyTables' chunked zlib+shuffle nodes,
# writes contiguous datasets under the same names); `.npy` / `.npz` paths are also accepted.
# ---------------------------------------------------------------------------------------------
def _rows(obj):
    if isinstance(obj, RaggedArray):
        return [obj[i] for i in range(len(obj))], False
    return [np.asarray(obj)], True


def save(path, obj, compression_level=1, tag="arr"):
    """Write a RaggedArray or ndarray.  `.h5`: the reference's node naming (ra.py:45-89);
    `.npy`: dense array; `.npz`: flat data + lengths."""
    import os
    from .util import h5min
    ext = os.path.splitext(path)[1].lower()
    rows, is_array = _rows(obj)
    if ext in (".h5", ".hdf5"):
        n_zeros = 1 if is_array else len(str(len(rows))) + 1
        h5min.write(path, {tag + "_" + str(i).zfill(n_zeros): np.asarray(r)
                           for i, r in enumerate(rows)})
        return path
    if ext == ".npz" or not is_array:
        flat = np.concatenate([np.asarray(r) for r in rows]) if rows else np.zeros(0)
        np.savez(path if ext == ".npz" else path + ".npz", data=flat,
                 lengths=np.array([len(r) for r in rows], dtype=np.int64))
        return path
    np.save(path, np.asarray(obj))
    return path


def load(path, keys=..., stride=1):
    """Inverse of ``save``; for `.h5` the semantics of ra.py:114-220: a file with one node
    gives an ndarray, several nodes give a RaggedArray with one row per node (sorted by
    name), ``stride`` slices every row, ``keys=None`` reads the old '/array' + '/lengths'
    layout."""
    import os
    from .util import h5min
    ext = os.path.splitext(path)[1].lower()
    if ext in (".h5", ".hdf5"):
        f = h5min.File(path)
        names = f.keys()
        if keys is None:
            if "lengths" in names:
                return RaggedArray(f["array"].read(), lengths=f["lengths"].read())[::stride]
            return _native(f["arr_0"].read())[::stride]
        if keys is Ellipsis:
            keys = sorted(names)
        if len(keys) == 1:
            return _native(f[keys[0]].read())[::1]
        shapes = [f[k].shape for k in keys]
        if not all(len(shapes[0]) == len(s) for s in shapes):
            raise DataInvalid(
                "Loading a RaggedArray using HDF5 file keys requires that all input arrays "
                "have the same dimension. Got shapes: %s" % shapes)
        if not all(shapes[0][1:] == s[1:] for s in shapes):
            raise DataInvalid(
                "Loading a RaggedArray using HDF5 file keys requires that all input arrays "
                "share nonragged dimensions. Got shapes: %s" % shapes)
        dtypes = [f[k].dtype for k in keys]
        if not all(dtypes[0] == d for d in dtypes):
            raise DataInvalid("Can't load keys in %s because the keys didn't have all the "
                              "same dtype. Keys were: %s" % (dtypes[0], keys))
        rows = [_native(f[k].read())[::stride] for k in keys]
        return RaggedArray(np.concatenate(rows), lengths=[len(r) for r in rows], copy=False)
    if ext == ".npz":
        z = np.load(path)
        r = RaggedArray(z["data"], lengths=z["lengths"])
        return r if stride == 1 else RaggedArray([row[::stride] for row in r])
    return np.load(path)[::stride]


def _native(a):
    return a.astype(a.dtype.newbyteorder("="), copy=False)
