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
