// Native GROMACS XTC reader (host code only; lives in libenspara_b200.so so that the loaders
// either side of the hot path need neither mdtraj nor xdrfile).
//
// Replaces, for .xtc inputs, the md.load(...) calls of the reference's trajectory loading
// (/root/reference/enspara/cluster/util.py:350-404 load_trajectories, mpi/io.py:142-194
// load_trajectory_as_striped, cluster/util.py:596-606 batch loading in reassign): frames are
// decoded straight into a caller-supplied (frames, selected atoms, 3) float32 array with the
// stride and atom selection applied while decoding, so a strided / atom-sliced load never
// materialises the full trajectory.
//
// Format (the published xdrfile "xdr3dfcoord" scheme): per frame a big-endian header
//   int magic = 1995, int natoms, int step, float time, float box[9], int natoms
// followed, for natoms <= 9, by 3*natoms raw floats, else by
//   float precision, int minint[3], int maxint[3], int smallidx, int nbytes, byte payload[nbytes]
// (padded to 4 bytes).  The payload is a bit stream, most significant bit first.  Every atom is
// either a "large" triple packed mixed-radix with radices size[d] = maxint[d] - minint[d] + 1
// (or three plain bit fields when a radix exceeds 24 bits), optionally followed by a run of
// "small" triples that are deltas to the previous atom, packed mixed-radix with the common
// radix magic[smallidx] into `smallidx` bits; the first small atom of a run is swapped with
// its large predecessor (the water oxygen / hydrogen trick).  smallidx adapts by +-1 per run.
// Coordinates are integer * (1 / precision) in float32, nanometres.
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <vector>

#include "eb_common.cuh"

namespace eb {
namespace xtc {

const int kMagic[] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203,
    256, 322, 406, 512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192,
    10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072,
    165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510,
    2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
constexpr int kFirstIdx = 9;
constexpr int kLastIdx = (int)(sizeof(kMagic) / sizeof(kMagic[0]));

struct Mapped {
    const unsigned char *p = nullptr;
    size_t len = 0;
    int fd = -1;
    int open_file(const char *path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return fail(EB_ERR_INVALID, "cannot open '%s' (errno %ld)", path, (long)errno);
        struct stat st;
        if (fstat(fd, &st) != 0) return fail(EB_ERR_INVALID, "cannot stat '%s'", path);
        len = (size_t)st.st_size;
        if (len == 0) return EB_OK;
        void *m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return fail(EB_ERR_INVALID, "cannot map '%s'", path);
        p = (const unsigned char *)m;
        return EB_OK;
    }
    ~Mapped()
    {
        if (p) munmap((void *)p, len);
        if (fd >= 0) ::close(fd);
    }
};

inline uint32_t be32(const unsigned char *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
inline int32_t be_i32(const unsigned char *p) { return (int32_t)be32(p); }
inline float be_f32(const unsigned char *p)
{
    const uint32_t u = be32(p);
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// MSB-first bit reader with a 64-bit window
struct BitReader {
    const unsigned char *p, *end;
    uint64_t window = 0;
    int have = 0;
    bool overrun = false;
    BitReader(const unsigned char *b, size_t n) : p(b), end(b + n) {}
    inline uint32_t take(int n)   // n <= 32
    {
        if (n == 0) return 0;
        while (have < n) {
            uint64_t byte = 0;
            if (p < end)
                byte = *p++;
            else
                overrun = true;
            window = (window << 8) | byte;
            have += 8;
        }
        have -= n;
        return (uint32_t)((window >> have) & ((n == 32) ? 0xffffffffull : ((1ull << n) - 1)));
    }
    // `nbits` bits holding a mixed-radix number whose bytes arrive least significant first
    inline void triple(int nbits, const uint32_t radix[3], int32_t out[3])
    {
        unsigned __int128 v = 0;
        int shift = 0;
        while (nbits > 8) {
            v |= (unsigned __int128)take(8) << shift;
            shift += 8;
            nbits -= 8;
        }
        if (nbits > 0) v |= (unsigned __int128)take(nbits) << shift;
        out[2] = (int32_t)(v % radix[2]);
        v /= radix[2];
        out[1] = (int32_t)(v % radix[1]);
        out[0] = (int32_t)(v / radix[1]);
    }
};

inline int bit_length_u128(unsigned __int128 v)
{
    int n = 0;
    while (v) {
        ++n;
        v >>= 1;
    }
    return n;
}
inline int bit_length_u32(uint32_t v)
{
    int n = 0;
    while (v) {
        ++n;
        v >>= 1;
    }
    return n;
}

struct FrameHeader {
    int natoms;
    int step;
    float time;
    size_t coord_off;    // offset of the coordinate block
    size_t next_off;     // offset of the next frame
};

// Parses the header at `off`; returns 0 ok, 1 clean end of file, <0 corrupt
static int parse_header(const Mapped &m, size_t off, FrameHeader *h)
{
    if (off == m.len) return 1;
    if (off + 56 > m.len) return -1;
    const unsigned char *p = m.p + off;
    if (be_i32(p) != 1995) return -2;
    h->natoms = be_i32(p + 4);
    h->step = be_i32(p + 8);
    h->time = be_f32(p + 12);
    if (h->natoms < 0 || be_i32(p + 52) != h->natoms) return -3;
    h->coord_off = off + 56;
    if (h->natoms <= 9) {
        h->next_off = h->coord_off + 12 * (size_t)h->natoms;
    } else {
        if (h->coord_off + 36 > m.len) return -1;
        const int nbytes = be_i32(m.p + h->coord_off + 32);
        if (nbytes < 0) return -4;
        h->next_off = h->coord_off + 36 + (((size_t)nbytes + 3) & ~(size_t)3);
    }
    if (h->next_off > m.len) return -1;
    return 0;
}

// Decodes one frame's coordinates; sel (n_sel indices, ascending or not) or all atoms.
static int decode_frame(const Mapped &m, const FrameHeader &h, const int32_t *sel, int n_sel,
                        std::vector<int32_t> &scratch, float *out)
{
    const int natoms = h.natoms;
    const unsigned char *p = m.p + h.coord_off;
    if (natoms <= 9) {
        for (int s = 0; s < (sel ? n_sel : natoms); ++s) {
            const int a = sel ? sel[s] : s;
            for (int d = 0; d < 3; ++d) out[3 * s + d] = be_f32(p + 12 * a + 4 * d);
        }
        return EB_OK;
    }
    const float precision = be_f32(p);
    int32_t minint[3], maxint[3];
    for (int d = 0; d < 3; ++d) {
        minint[d] = be_i32(p + 4 + 4 * d);
        maxint[d] = be_i32(p + 16 + 4 * d);
    }
    int smallidx = be_i32(p + 28);
    const int nbytes = be_i32(p + 32);
    if (smallidx < kFirstIdx || smallidx >= kLastIdx || !(precision > 0.f))
        return fail(EB_ERR_INVALID, "%s: corrupt XTC frame header", "xtc_read");
    uint32_t size[3];
    bool wide = false;
    for (int d = 0; d < 3; ++d) {
        const int64_t s = (int64_t)maxint[d] - (int64_t)minint[d] + 1;
        if (s <= 0) return fail(EB_ERR_INVALID, "%s: corrupt XTC coordinate range", "xtc_read");
        size[d] = (uint32_t)s;
        wide |= s > 0xffffff;
    }
    int bits_large = 0, bits_d[3] = {0, 0, 0};
    if (wide) {
        for (int d = 0; d < 3; ++d) bits_d[d] = bit_length_u32(size[d]);
    } else {
        bits_large = bit_length_u128((unsigned __int128)size[0] * size[1] * size[2]);
    }
    int smaller = kMagic[smallidx - 1 > kFirstIdx ? smallidx - 1 : kFirstIdx] / 2;
    int smallnum = kMagic[smallidx] / 2;
    uint32_t radix_small[3] = {(uint32_t)kMagic[smallidx], (uint32_t)kMagic[smallidx],
                               (uint32_t)kMagic[smallidx]};

    scratch.resize(3 * (size_t)natoms);
    int32_t *ints = scratch.data();
    BitReader br(p + 36, (size_t)nbytes);
    int written = 0, decoded = 0, run = 0;
    while (decoded < natoms) {
        int32_t cur[3];
        if (wide) {
            for (int d = 0; d < 3; ++d) cur[d] = (int32_t)br.take(bits_d[d]);
        } else {
            br.triple(bits_large, size, cur);
        }
        ++decoded;
        for (int d = 0; d < 3; ++d) cur[d] += minint[d];
        int32_t prev[3] = {cur[0], cur[1], cur[2]};
        int delta_idx = 0;
        if (br.take(1)) {
            run = (int)br.take(5);
            delta_idx = run % 3;
            run -= delta_idx;
            delta_idx -= 1;
        }
        if (run > 0) {
            if (written + run / 3 + 1 > natoms)
                return fail(EB_ERR_INVALID, "%s: corrupt XTC run length", "xtc_read");
            for (int k = 0; k < run; k += 3) {
                int32_t sm[3];
                br.triple(smallidx, radix_small, sm);
                ++decoded;
                for (int d = 0; d < 3; ++d) sm[d] += prev[d] - smallnum;
                if (k == 0) {
                    // the first small atom goes BEFORE its large predecessor
                    for (int d = 0; d < 3; ++d) {
                        ints[3 * written + d] = sm[d];
                        const int32_t t = prev[d];
                        prev[d] = sm[d];
                        sm[d] = t;
                    }
                    ++written;
                } else {
                    for (int d = 0; d < 3; ++d) prev[d] = sm[d];
                }
                for (int d = 0; d < 3; ++d) ints[3 * written + d] = sm[d];
                ++written;
            }
        } else {
            for (int d = 0; d < 3; ++d) ints[3 * written + d] = cur[d];
            ++written;
        }
        smallidx += delta_idx;
        if (smallidx < kFirstIdx || smallidx >= kLastIdx)
            return fail(EB_ERR_INVALID, "%s: corrupt XTC small-index", "xtc_read");
        if (delta_idx < 0) {
            smallnum = smaller;
            smaller = smallidx > kFirstIdx ? kMagic[smallidx - 1] / 2 : 0;
        } else if (delta_idx > 0) {
            smaller = smallnum;
            smallnum = kMagic[smallidx] / 2;
        }
        radix_small[0] = radix_small[1] = radix_small[2] = (uint32_t)kMagic[smallidx];
    }
    if (written != natoms || br.overrun)
        return fail(EB_ERR_INVALID, "%s: XTC frame did not decode to its atom count", "xtc_read");
    const float inv = 1.0f / precision;
    for (int s = 0; s < (sel ? n_sel : natoms); ++s) {
        const int a = sel ? sel[s] : s;
        for (int d = 0; d < 3; ++d) out[3 * s + d] = (float)ints[3 * a + d] * inv;
    }
    return EB_OK;
}

}  // namespace xtc
}  // namespace eb

using namespace eb;

extern "C" {

// Number of frames and atoms of an .xtc file (walks the frame headers, decodes nothing).
int eb_xtc_scan(const char *path, int64_t *n_frames, int32_t *n_atoms)
{
    EB_CHECK_ARG(path && n_frames && n_atoms, "xtc_scan: null pointer");
    xtc::Mapped m;
    const int rc = m.open_file(path);
    if (rc != EB_OK) return rc;
    size_t off = 0;
    int64_t n = 0;
    int atoms = -1;
    for (;;) {
        xtc::FrameHeader h;
        const int st = xtc::parse_header(m, off, &h);
        if (st == 1) break;
        if (st < 0)
            return fail(EB_ERR_INVALID, "%s: not an XTC file or truncated (frame %ld)", path,
                        (long)n);
        if (atoms < 0) atoms = h.natoms;
        if (h.natoms != atoms)
            return fail(EB_ERR_INVALID, "%s: atom count changes at frame %ld", path, (long)n);
        ++n;
        off = h.next_off;
    }
    *n_frames = n;
    *n_atoms = atoms < 0 ? 0 : atoms;
    return EB_OK;
}

// Frames first, first+stride, ... (at most max_frames of them; max_frames < 0: all) of an .xtc
// file into xyz_out[(frame, atom, 3)] float32, nanometres; atom_idx (n_sel indices into the
// file's atoms, optional) selects and orders the atoms.  *n_read receives the frames written.
int eb_xtc_read(const char *path, int64_t first, int64_t stride, int64_t max_frames,
                const int32_t *atom_idx, int32_t n_sel, float *xyz_out, int64_t *n_read)
{
    EB_CHECK_ARG(path && xyz_out && n_read, "xtc_read: null pointer");
    EB_CHECK_ARG(first >= 0 && stride >= 1, "xtc_read: need first >= 0 and stride >= 1");
    EB_CHECK_ARG(!atom_idx || n_sel > 0, "xtc_read: empty atom selection");
    xtc::Mapped m;
    const int rc = m.open_file(path);
    if (rc != EB_OK) return rc;
    std::vector<int32_t> scratch;
    size_t off = 0;
    int64_t frame = 0, written = 0;
    for (;;) {
        if (max_frames >= 0 && written >= max_frames) break;
        xtc::FrameHeader h;
        const int st = xtc::parse_header(m, off, &h);
        if (st == 1) break;
        if (st < 0)
            return fail(EB_ERR_INVALID, "%s: not an XTC file or truncated (frame %ld)", path,
                        (long)frame);
        if (frame >= first && (frame - first) % stride == 0) {
            if (atom_idx)
                for (int s = 0; s < n_sel; ++s)
                    if (atom_idx[s] < 0 || atom_idx[s] >= h.natoms)
                        return fail(EB_ERR_INVALID, "%s: atom index out of range (%ld of %ld)",
                                    path, (long)atom_idx[s], (long)h.natoms);
            const size_t per = 3 * (size_t)(atom_idx ? n_sel : h.natoms);
            const int drc = xtc::decode_frame(m, h, atom_idx, n_sel, scratch,
                                              xyz_out + (size_t)written * per);
            if (drc != EB_OK) return drc;
            ++written;
        }
        ++frame;
        off = h.next_off;
    }
    *n_read = written;
    return EB_OK;
}

}  // extern "C"
