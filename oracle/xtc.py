"""Minimal GROMACS XTC reader (pure Python).  TEST INFRASTRUCTURE.

Used once, by scripts/make_golden.py, to turn the reference's own RMSD test fixture
(/root/reference/enspara/test/data/frame0.xtc, 501 frames x 22 atoms, the input of the golden
statistics at enspara/test/test_cluster.py:200-238) into tests/golden/frame0_xyz.npy, because
neither mdtraj nor any XTC library exists in this image.

The format is the published xdrfile "xdr3dfcoord" compressed-coordinate scheme: a big-endian
bit stream holding, per atom, either three "large" integers packed with mixed radix
(sizes = maxint-minint+1) or runs of "small" deltas packed with radix magicints[smallidx].
"""
import struct

import numpy as np

_MAGICINTS = [
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203,
    256, 322, 406, 512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192,
    10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072,
    165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510,
    2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216]
_FIRSTIDX = 9
_LASTIDX = len(_MAGICINTS)


class _Bits:
    """MSB-first bit reader over a bytes object."""

    def __init__(self, data):
        self.v = int.from_bytes(data, "big")
        self.nbits = 8 * len(data)
        self.pos = 0

    def take(self, n):
        if n == 0:
            return 0
        shift = self.nbits - self.pos - n
        if shift < 0:
            raise ValueError("XTC bit stream exhausted")
        self.pos += n
        return (self.v >> shift) & ((1 << n) - 1)

    def ints(self, nbits, sizes):
        """xdrfile receiveints(): bytes arrive low byte first, then mixed-radix decode."""
        val = 0
        k = 0
        while nbits > 8:
            val |= self.take(8) << (8 * k)
            k += 1
            nbits -= 8
        if nbits > 0:
            val |= self.take(nbits) << (8 * k)
        out = [0, 0, 0]
        for i in (2, 1):
            out[i] = val % sizes[i]
            val //= sizes[i]
        out[0] = val
        return out


def _sizeofint(size):
    n = 0
    num = 1
    while size >= num and n < 32:
        n += 1
        num <<= 1
    return n


def _decode_coords(buf, off, natoms):
    (precision,) = struct.unpack_from(">f", buf, off)
    off += 4
    minint = struct.unpack_from(">3i", buf, off)
    off += 12
    maxint = struct.unpack_from(">3i", buf, off)
    off += 12
    (smallidx,) = struct.unpack_from(">i", buf, off)
    off += 4
    (nbytes,) = struct.unpack_from(">i", buf, off)
    off += 4
    payload = buf[off:off + nbytes]
    off += (nbytes + 3) // 4 * 4

    sizeint = [maxint[i] - minint[i] + 1 for i in range(3)]
    if any(s > 0xFFFFFF for s in sizeint):
        bitsizeint = [_sizeofint(s) for s in sizeint]
        bitsize = 0
    else:
        bitsizeint = None
        bitsize = (sizeint[0] * sizeint[1] * sizeint[2]).bit_length()

    smaller = _MAGICINTS[max(_FIRSTIDX, smallidx - 1)] // 2
    smallnum = _MAGICINTS[smallidx] // 2
    sizesmall = [_MAGICINTS[smallidx]] * 3

    bits = _Bits(payload)
    out = np.empty((natoms, 3), np.int64)
    w = 0          # atoms written
    i = 0          # atoms decoded
    run = 0
    while i < natoms:
        if bitsize == 0:
            this = [bits.take(bitsizeint[0]), bits.take(bitsizeint[1]), bits.take(bitsizeint[2])]
        else:
            this = bits.ints(bitsize, sizeint)
        i += 1
        this = [this[d] + minint[d] for d in range(3)]
        prev = list(this)
        flag = bits.take(1)
        is_smaller = 0
        if flag == 1:
            run = bits.take(5)
            is_smaller = run % 3
            run -= is_smaller
            is_smaller -= 1
        if run > 0:
            for k in range(0, run, 3):
                cur = bits.ints(smallidx, sizesmall)
                i += 1
                cur = [cur[d] + prev[d] - smallnum for d in range(3)]
                if k == 0:
                    # first small atom is swapped with the large one (water O/H trick)
                    cur, prev = prev, cur
                    out[w] = prev
                    w += 1
                else:
                    prev = list(cur)
                out[w] = cur
                w += 1
        else:
            out[w] = this
            w += 1
        smallidx += is_smaller
        if is_smaller < 0:
            smallnum = smaller
            smaller = _MAGICINTS[smallidx - 1] // 2 if smallidx > _FIRSTIDX else 0
        elif is_smaller > 0:
            smaller = smallnum
            smallnum = _MAGICINTS[smallidx] // 2
        sizesmall = [_MAGICINTS[smallidx]] * 3
    if w != natoms:
        raise ValueError("XTC decode wrote %d of %d atoms" % (w, natoms))
    # xdrfile: float = int * (1/precision), both in float32
    inv = np.float32(1.0) / np.float32(precision)
    return (out.astype(np.float32) * inv).astype(np.float32), off, precision


def read_xtc(path):
    """Return (xyz float32 (n_frames, n_atoms, 3) in nm, times, steps, boxes)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    off = 0
    frames, times, steps, boxes = [], [], [], []
    while off < len(buf):
        magic, natoms, step = struct.unpack_from(">3i", buf, off)
        if magic != 1995:
            raise ValueError("bad XTC magic %d at offset %d" % (magic, off))
        off += 12
        (time,) = struct.unpack_from(">f", buf, off)
        off += 4
        box = struct.unpack_from(">9f", buf, off)
        off += 36
        (lsize,) = struct.unpack_from(">i", buf, off)
        off += 4
        if lsize != natoms:
            raise ValueError("atom count mismatch")
        if natoms <= 9:
            xyz = np.array(struct.unpack_from(">%df" % (3 * natoms), buf, off),
                           np.float32).reshape(natoms, 3)
            off += 12 * natoms
        else:
            xyz, off, _ = _decode_coords(buf, off, natoms)
        frames.append(xyz)
        times.append(time)
        steps.append(step)
        boxes.append(box)
    return (np.stack(frames), np.array(times, np.float32), np.array(steps),
            np.array(boxes, np.float32).reshape(-1, 3, 3))
