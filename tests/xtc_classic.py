"""TEST INFRASTRUCTURE: an independent reader of the XTC format, written from the published description of the
"xdr3dfcoord" compression (the xdrfile library, LGPL, distributed with GROMACS since 3.0) in its classic shape: a
most-significant-bit-first bit stream, triples packed as one mixed-radix number that is read eight bits at a time,
least-significant byte first.  It shares no code with gmxstub/xdrtraj.hpp (which reads through a 64-bit window and
divides by multiply-high) and uses Python integers of arbitrary size instead of the byte-array long division, so an
error in the fast decoder, in the stand-in's encoder or in the frame layout shows up as a disagreement.

Pure Python: small systems only (about 10^5 atoms per second)."""
import struct

import numpy as np

XTC_MAGIC = 1995
FIRSTIDX = 9
# thresholds of the adaptive small-difference coding: 2^(i/3) rounded, with the three historical irregular entries
# (5060, 524287, 8388607) every XTC file ever written depends on
MAGICINTS = [0] * 9 + [
    8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812, 1024, 1290, 1625, 2048,
    2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570, 104031,
    131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245,
    3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216]


class _Bits:
    """Most-significant-bit-first reader over the opaque byte block of one frame."""

    def __init__(self, data):
        self.value = int.from_bytes(data, "big")
        self.left = 8 * len(data)

    def take(self, n):
        if n == 0:
            return 0
        if n > self.left:
            raise ValueError("XTC bit stream exhausted")
        self.left -= n
        return (self.value >> self.left) & ((1 << n) - 1)

    def triple(self, nbits, sizes):
        """Three integers packed as v = (a * sizes[1] + b) * sizes[2] + c; the stream holds v in 8-bit pieces, the least
        significant piece first, the last (most significant) piece with the remaining nbits % 8 bits."""
        v, shift = 0, 0
        while nbits > 8:
            v |= self.take(8) << shift
            shift += 8
            nbits -= 8
        if nbits > 0:
            v |= self.take(nbits) << shift
        v, c = divmod(v, sizes[2])
        a, b = divmod(v, sizes[1])
        return [a, b, c]


def _bits_for(size):
    """Bits needed for one integer in [0, size)  (the format counts size itself: 2^bits > size)."""
    bits, num = 0, 1
    while size >= num and bits < 32:
        bits += 1
        num <<= 1
    return bits


def _bits_for_product(sizes):
    return (sizes[0] * sizes[1] * sizes[2]).bit_length()


def read_frames(path):
    """-> list of dicts {natoms, step, time, box (9,), x (natoms, 3) float32, precision}"""
    raw = open(path, "rb").read()
    off, frames = 0, []
    while off < len(raw):
        magic, natoms, step = struct.unpack_from(">iii", raw, off)
        if magic != XTC_MAGIC:
            raise ValueError("bad XTC magic %d at byte %d" % (magic, off))
        time, = struct.unpack_from(">f", raw, off + 12)
        box = np.frombuffer(raw, ">f4", 9, off + 16).astype(np.float32)
        off += 52
        size, = struct.unpack_from(">i", raw, off)
        off += 4
        if size != natoms:
            raise ValueError("atom counts of the frame header and of the coordinate block differ")
        if size <= 9:
            x = np.frombuffer(raw, ">f4", 3 * size, off).astype(np.float32).reshape(size, 3)
            off += 12 * size
            frames.append(dict(natoms=natoms, step=step, time=time, box=box, x=x, precision=0.0))
            continue
        precision, = struct.unpack_from(">f", raw, off)
        minint = struct.unpack_from(">iii", raw, off + 4)
        maxint = struct.unpack_from(">iii", raw, off + 16)
        smallidx, nbytes = struct.unpack_from(">ii", raw, off + 28)
        off += 36
        bits = _Bits(raw[off:off + nbytes])
        off += (nbytes + 3) & ~3
        sizeint = [maxint[k] - minint[k] + 1 for k in range(3)]
        if max(sizeint) > 0xFFFFFF:
            widths, bitsize = [_bits_for(s) for s in sizeint], 0
        else:
            widths, bitsize = None, _bits_for_product(sizeint)
        smaller = MAGICINTS[max(FIRSTIDX, smallidx - 1)] // 2
        smallnum = MAGICINTS[smallidx] // 2
        sizesmall = [MAGICINTS[smallidx]] * 3
        out = np.empty((size, 3), np.int64)
        i, run = 0, 0
        while i < size:
            if bitsize == 0:
                this = [bits.take(widths[k]) for k in range(3)]
            else:
                this = bits.triple(bitsize, sizeint)
            this = [this[k] + minint[k] for k in range(3)]
            first_at = i
            i += 1
            is_smaller = 0
            if bits.take(1) == 1:
                run = bits.take(5)
                is_smaller = run % 3
                run -= is_smaller
                is_smaller -= 1
            if run > 0:
                prev = this
                for k in range(0, run, 3):
                    d = bits.triple(smallidx, sizesmall)
                    cur = [d[q] + prev[q] - smallnum for q in range(3)]
                    if k == 0:
                        # the writer exchanged the first two atoms of the run (water: O after its first H)
                        out[first_at] = cur
                        out[first_at + 1] = this
                    else:
                        out[i] = cur
                    i += 1
                    prev = cur
            else:
                out[first_at] = this
            smallidx += is_smaller
            if is_smaller < 0:
                smallnum = smaller
                smaller = MAGICINTS[smallidx - 1] // 2 if smallidx > FIRSTIDX else 0
            elif is_smaller > 0:
                smaller = smallnum
                smallnum = MAGICINTS[smallidx] // 2
            sizesmall = [MAGICINTS[smallidx]] * 3
        inv = np.float32(1.0) / np.float32(precision)
        x = out.astype(np.int32).astype(np.float32) * inv
        frames.append(dict(natoms=natoms, step=step, time=time, box=box, x=x, precision=precision))
    return frames


TRR_MAGIC = 1993


def read_trr_frames(path):
    """Independent TRR reader (layout as published with the xdrfile library: magic, strlen("GMX_trn_file") + 1, the XDR
    string, ten block sizes, natoms, step, nre, time, lambda in the file's real size, then box / x / v / f).
    -> list of dicts {natoms, step, time, box, x, v, f} (missing blocks are None)"""
    raw = open(path, "rb").read()
    off, frames = 0, []
    while off < len(raw):
        magic, slen, n = struct.unpack_from(">iii", raw, off)
        if magic != TRR_MAGIC or slen != 13 or n != 12 or raw[off + 12:off + 24] != b"GMX_trn_file":
            raise ValueError("bad TRR frame header at byte %d" % off)
        off += 24
        ir, e, box_size, vir, pres, top, sym, xs, vs, fs, natoms, step, nre = struct.unpack_from(">13i", raw, off)
        off += 52
        real = box_size // 9 if box_size else (xs or vs or fs) // (3 * natoms)
        fmt = {4: ">f4", 8: ">f8"}[real]
        time, lam = np.frombuffer(raw, fmt, 2, off)
        off += 2 * real + ir + e
        fr = dict(natoms=natoms, step=step, time=float(time), box=None, x=None, v=None, f=None)
        if box_size:
            fr["box"] = np.frombuffer(raw, fmt, 9, off).astype(np.float32)
        off += box_size + vir + pres + top + sym
        for key, size in (("x", xs), ("v", vs), ("f", fs)):
            if size:
                fr[key] = np.frombuffer(raw, fmt, 3 * natoms, off).astype(np.float32).reshape(natoms, 3)
            off += size
        frames.append(fr)
    return frames
