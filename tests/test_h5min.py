"""CPU: the dependency-free HDF5 reader/writer (enspara_b200/util/h5min.py) that stands in for
PyTables on the I/O either side of the clustering path (ra.save / ra.load, --features x.h5)."""
import os
import struct
import zlib

import numpy as np
import pytest
from numpy.testing import assert_array_equal

from enspara_b200 import ra
from enspara_b200.exception import DataInvalid
from enspara_b200.util import h5min

REF_H5 = "/root/reference/enspara/test/data/frame0.h5"


def test_write_read_round_trip(tmp_path):
    p = str(tmp_path / "t.h5")
    arrays = {"arr_0": np.arange(10, dtype=np.int64),
              "arr_1": np.random.RandomState(0).rand(5, 3),
              "f32": np.linspace(0, 1, 7, dtype=np.float32),
              "i8": np.array([-3, 4], dtype=np.int8),
              "empty": np.zeros((0, 4), dtype=np.float64)}
    h5min.write(p, arrays)
    ls = h5min.list_datasets(p)
    assert sorted(ls) == sorted("/" + k for k in arrays)
    for k, v in arrays.items():
        got = h5min.read(p, k)
        assert got.dtype == v.dtype and got.shape == v.shape
        assert_array_equal(got, v)


def test_many_nodes(tmp_path):
    p = str(tmp_path / "many.h5")
    arrays = {"arr_%04d" % i: np.full(i % 5 + 1, i, dtype=np.int64) for i in range(100)}
    h5min.write(p, arrays)
    f = h5min.File(p)
    assert sorted(f.keys()) == sorted(arrays)
    assert_array_equal(f["arr_0042"].read(), arrays["arr_0042"])


def test_ra_save_load_names_follow_the_reference(tmp_path):
    """ra.py:45-89: ndarray -> '/arr_0'; RaggedArray of n rows -> 'arr_' + zero-filled index
    with len(str(n)) + 1 digits.  ra.py:114-220: one node -> ndarray, several -> RaggedArray."""
    p = str(tmp_path / "a.h5")
    a = np.arange(12, dtype=np.int64).reshape(3, 4)
    ra.save(p, a)
    assert h5min.File(p).keys() == ["arr_0"]
    assert_array_equal(ra.load(p), a)
    # a single node is returned whole: the reference ignores ``stride`` there (ra.py:155-158)
    assert_array_equal(ra.load(p, stride=2), a)
    assert_array_equal(ra.load(p, keys=None, stride=2), a[::2])
    r = ra.RaggedArray([np.arange(3.0), np.arange(5.0), np.arange(2.0)])
    ra.save(p, r)
    assert sorted(h5min.File(p).keys()) == ["arr_00", "arr_01", "arr_02"]
    back = ra.load(p)
    assert isinstance(back, ra.RaggedArray)
    assert_array_equal(back.lengths, [3, 5, 2])
    assert_array_equal(back[1], np.arange(5.0))
    strided = ra.load(p, stride=2)
    assert_array_equal(strided.lengths, [2, 3, 1])
    one = ra.load(p, keys=["arr_01"])
    assert_array_equal(one, np.arange(5.0))


def test_ra_load_rejects_mismatched_nodes(tmp_path):
    p = str(tmp_path / "bad.h5")
    h5min.write(p, {"arr_0": np.zeros((3, 2)), "arr_1": np.zeros((3, 4))})
    with pytest.raises(DataInvalid):
        ra.load(p)
    h5min.write(p, {"arr_0": np.zeros(3), "arr_1": np.zeros(3, dtype=np.int64)})
    with pytest.raises(DataInvalid):
        ra.load(p)


def _chunked_file(path, arr, chunk_rows, leaf_size=0):
    """Hand-assemble a PyTables-style file: one chunked dataset with shuffle + deflate, to
    exercise the reader's B-tree / filter path without PyTables."""
    h5min.write(path, {"x": arr})          # start from a valid file, then graft a chunked one
    base = bytearray(open(path, "rb").read())
    f = h5min.File(path)
    ohdr = f._links["x"]
    itemsize = arr.dtype.itemsize
    rank = arr.ndim
    n = arr.shape[0]
    chunk_shape = (chunk_rows,) + arr.shape[1:]
    chunks = []
    for lo in range(0, n, chunk_rows):
        block = np.zeros(chunk_shape, dtype=arr.dtype)
        m = min(chunk_rows, n - lo)
        block[:m] = arr[lo:lo + m]
        raw = np.frombuffer(block.tobytes(), dtype=np.uint8)
        shuf = raw.reshape(-1, itemsize).T.tobytes()
        chunks.append((lo, zlib.compress(shuf, 1)))
    # append chunk data + one leaf B-tree node
    addrs = []
    for lo, comp in chunks:
        base.extend(b"\0" * (-len(base) % 8))
        addrs.append(len(base))
        base.extend(comp)
    def leaf(entries):
        base.extend(b"\0" * (-len(base) % 8))
        at = len(base)
        node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries),
                                               0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF))
        for (lo, comp), a in entries:
            node += struct.pack("<II", len(comp), 0)
            node += struct.pack("<%dQ" % (rank + 1), lo, *([0] * rank))
            node += struct.pack("<Q", a)
        node += struct.pack("<II", 0, 0) + struct.pack("<%dQ" % (rank + 1), n, *([0] * rank))
        base.extend(node)
        return at

    entries = list(zip(chunks, addrs))
    if leaf_size and len(entries) > leaf_size:
        # two-level tree: leaves of `leaf_size` chunks under one internal node (level 1)
        kids = [(entries[i][0][0], leaf(entries[i:i + leaf_size]))
                for i in range(0, len(entries), leaf_size)]
        base.extend(b"\0" * (-len(base) % 8))
        btree = len(base)
        node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 1, len(kids),
                                               0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF))
        for lo, at in kids:
            node += struct.pack("<II", 0, 0)
            node += struct.pack("<%dQ" % (rank + 1), lo, *([0] * rank))
            node += struct.pack("<Q", at)
        node += struct.pack("<II", 0, 0) + struct.pack("<%dQ" % (rank + 1), n, *([0] * rank))
        base.extend(node)
    else:
        btree = leaf(entries)
    # new object header for the dataset: dataspace, datatype, filters, chunked layout
    space = struct.pack("<BBBBI", 1, rank, 0, 0, 0) + struct.pack("<%dQ" % rank, *arr.shape)
    filt = struct.pack("<BB6x", 1, 2)
    filt += struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", itemsize, 0)
    filt += struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 1, 0)
    layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", btree) + \
        struct.pack("<%dI" % (rank + 1), *(chunk_shape + (itemsize,)))
    msgs = [h5min._msg(0x01, space), h5min._msg(0x03, h5min._dtype_msg(arr.dtype), 1),
            h5min._msg(0x0B, filt), h5min._msg(0x08, layout)]
    body = b"".join(msgs)
    base.extend(b"\0" * (-len(base) % 8))
    new_ohdr = len(base)
    base.extend(struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body)
    # repoint the symbol table entry
    idx = base.find(struct.pack("<Q", ohdr), base.find(b"SNOD"))
    base[idx:idx + 8] = struct.pack("<Q", new_ohdr)
    struct.pack_into("<Q", base, 40, len(base))      # end-of-file address
    open(path, "wb").write(bytes(base))


@pytest.mark.parametrize("dtype,shape,chunk", [(np.float32, (101, 7, 3), 16),
                                               (np.int64, (33,), 8),
                                               (np.float64, (10, 4), 10)])
def test_reads_chunked_shuffled_deflated_nodes(tmp_path, dtype, shape, chunk):
    rs = np.random.RandomState(1)
    arr = (rs.rand(*shape) * 100).astype(dtype)
    p = str(tmp_path / "c.h5")
    _chunked_file(p, arr, chunk)
    got = h5min.read(p, "x")
    assert got.dtype == arr.dtype
    assert_array_equal(got, arr)


def test_reads_two_level_chunk_btrees(tmp_path):
    """Large PyTables arrays index their chunks with multi-level B-trees."""
    arr = np.arange(97 * 5, dtype=np.float64).reshape(97, 5)
    p = str(tmp_path / "deep.h5")
    _chunked_file(p, arr, 4, leaf_size=3)       # 25 chunks, 9 leaves under one internal node
    assert_array_equal(h5min.read(p, "x"), arr)


@pytest.mark.reference
@pytest.mark.skipif(not os.path.exists(REF_H5), reason="needs /root/reference")
def test_reference_fixture_decodes_to_the_committed_golden(frame0_h5_xyz):
    """The committed golden IS the reference's frame0.h5 (PyTables CArray, zlib + shuffle)."""
    ls = h5min.list_datasets(REF_H5)
    assert ls["/coordinates"][0] == (501, 22, 3)
    assert_array_equal(h5min.read(REF_H5, "coordinates"), frame0_h5_xyz)
    assert h5min.File(REF_H5).attrs["program"] == "MDTraj"


def test_load_h5_as_striped(tmp_path):
    """mpi/io.py:16-65 on a single rank: lengths per node, rows concatenated."""
    from enspara_b200.cluster import io as cio
    p = str(tmp_path / "feat.h5")
    r = ra.RaggedArray([np.random.RandomState(i).rand(n, 4) for i, n in enumerate((5, 9, 3))])
    ra.save(p, r)
    lengths, data = cio.load_h5_as_striped(p, stride=1)
    assert lengths == [5, 9, 3]
    assert data.shape == (17, 4)
    assert_array_equal(data[5:14], r[1])
    lengths, data = cio.load_h5_as_striped(p, stride=2)
    assert lengths == [3, 5, 2] and data.shape == (10, 4)
