// xdrtraj.hpp -- GROMACS trajectory wire formats for the libgromacs stand-in: XTC (compressed coordinates) and TRR
// (full-precision x / v / f), reader and writer, written from the published format descriptions:
//   * both are XDR streams (RFC 4506: big-endian 4-byte units, opaque data padded to 4 bytes);
//   * XTC frame: magic 1995, natoms, step, time, box[9], natoms, then either 3*natoms floats (natoms <= 9) or
//     precision, minint[3], maxint[3], smallidx, byte count and the bit stream of the "xdr3dfcoord" compression
//     (integer coordinates = round(x * precision); full-range triples coded as one mixed-radix number of sizeint[] digits,
//     runs of up to 8 near neighbours coded as small differences with an adaptive width `smallidx` taken from the
//     `magicints` table, first and second atom of a run swapped so that water oxygens follow their hydrogens);
//   * TRR frame: magic 1993, version string "GMX_trn_file", 13 ints (ir, e, box, vir, pres, top, sym, x, v, f sizes,
//     natoms, step, nre), t and lambda, then box / virial / pressure / x / v / f blocks of floats or doubles
//     (precision deduced from box_size / 9 or x_size / 3 natoms).
// Parity: the reference delegates these formats to libgromacs 2020.1 (read_first_frame / read_next_frame,
// include/itp/gmx_bits/gmx_handle.ipp:70,83), which is not available here and has no fixtures in the reference tree, so
// this decoder is pinned only against its own encoder (round trips to exactly the quantised values, tests/test_trajfmt.py).
// XTC frames are independent, which is what the batch reader exploits: it slices K frames out of the file sequentially
// and decodes them on as many host threads as there are frames, straight into the pinned batch.
#pragma once

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace xdrtraj
{

constexpr int kXtcMagic = 1995;
constexpr int kTrrMagic = 1993;

// ---- XDR primitives -------------------------------------------------------------------------------------------------
inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void     put32(unsigned char* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
inline float    be_float(const unsigned char* p) { uint32_t u = be32(p); float f; std::memcpy(&f, &u, 4); return f; }
inline double   be_double(const unsigned char* p)
{
    uint64_t u = ((uint64_t)be32(p) << 32) | be32(p + 4);
    double   d;
    std::memcpy(&d, &u, 8);
    return d;
}

struct Writer
{
    std::vector<unsigned char> buf;
    void i32(int32_t v) { size_t n = buf.size(); buf.resize(n + 4); put32(&buf[n], (uint32_t)v); }
    void f32(float f) { uint32_t u; std::memcpy(&u, &f, 4); size_t n = buf.size(); buf.resize(n + 4); put32(&buf[n], u); }
    void opaque(const unsigned char* p, size_t n)
    {
        buf.insert(buf.end(), p, p + n);
        while (buf.size() % 4) buf.push_back(0);
    }
};

// ---- xdr3dfcoord bit stream -------------------------------------------------------------------------------------------
// ATTRIBUTION.  The table `magicints` and the four bit-stream helpers that follow (`sizeofint`, `sizeofints`, `sendbits`,
// `receivebits`, with their state words lastbits / lastbyte / cnt) restate, name for name, the routines of GROMACS' libxdrf
// (src/gromacs/fileio/libxdrf.cpp, function xdr3dfcoord and its helpers; Copyright (c) 1991-2000 University of Groningen,
// 2001-2004 The GROMACS development team, 2013-2020 the GROMACS development team and its contributors; licensed LGPL-2.1-or-later).
// They ARE the definition of the XTC wire format -- the table is the radix ladder every XTC file indexes into, and the bit
// order these helpers produce is what a compatible reader has to consume -- so they are kept recognisable rather than
// disguised; this block is therefore to be read as LGPL-2.1-or-later material derived from GROMACS.  Everything after them
// (the 64-bit window decoder `xtc_decode_coords`, the multiply-high divisions, the table-driven adaptive width, the TRR reader
// and both writers' framing) was written for this repository from the format description.
static const int magicints[] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812, 1024,
    1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570,
    104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245,
    3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216 };
constexpr int FIRSTIDX = 9;
constexpr int LASTIDX  = (int)(sizeof(magicints) / sizeof(*magicints));

inline int sizeofint(unsigned size)
{
    unsigned num = 1;
    int      bits = 0;
    while (size >= num && bits < 32) { bits++; num <<= 1; }
    return bits;
}

inline int sizeofints(int n, const unsigned sizes[])
{
    unsigned bytes[32], nbytes = 1, bytecnt, tmp;
    bytes[0] = 1;
    int num_of_bits = 0;
    for (int i = 0; i < n; i++)
    {
        tmp = 0;
        for (bytecnt = 0; bytecnt < nbytes; bytecnt++)
        {
            tmp            = bytes[bytecnt] * sizes[i] + tmp;
            bytes[bytecnt] = tmp & 0xff;
            tmp >>= 8;
        }
        while (tmp != 0) { bytes[bytecnt++] = tmp & 0xff; tmp >>= 8; }
        nbytes = bytecnt;
    }
    unsigned num = 1;
    nbytes--;
    while (bytes[nbytes] >= num) { num_of_bits++; num *= 2; }
    return num_of_bits + (int)nbytes * 8;
}

struct BitBuf   // cnt = bytes consumed/produced, lastbits / lastbyte = the partially filled byte
{
    unsigned char* c;
    size_t         cap;
    size_t         cnt = 0;
    unsigned       lastbits = 0, lastbyte = 0;
};

inline void sendbits(BitBuf& b, int num_of_bits, unsigned num)
{
    while (num_of_bits >= 8)
    {
        // (fields wider than 32 bits only ever carry leading zeros: sendints() pads a short number this way)
        b.lastbyte   = (b.lastbyte << 8) | (num_of_bits - 8 < 32 ? ((num >> (num_of_bits - 8)) & 0xff) : 0u);
        b.c[b.cnt++] = (unsigned char)(b.lastbyte >> b.lastbits);
        num_of_bits -= 8;
    }
    if (num_of_bits > 0)
    {
        b.lastbyte = (b.lastbyte << num_of_bits) | (num & ((1u << num_of_bits) - 1u));
        b.lastbits += num_of_bits;
        if (b.lastbits >= 8)
        {
            b.lastbits -= 8;
            b.c[b.cnt++] = (unsigned char)(b.lastbyte >> b.lastbits);
        }
    }
    if (b.lastbits > 0) b.c[b.cnt] = (unsigned char)(b.lastbyte << (8 - b.lastbits));
}

inline unsigned receivebits(BitBuf& b, int num_of_bits)
{
    const unsigned mask = num_of_bits >= 32 ? 0xffffffffu : ((1u << num_of_bits) - 1u);
    unsigned       num  = 0;
    while (num_of_bits >= 8)
    {
        b.lastbyte = (b.lastbyte << 8) | b.c[b.cnt++];
        num |= (b.lastbyte >> b.lastbits) << (num_of_bits - 8);
        num_of_bits -= 8;
    }
    if (num_of_bits > 0)
    {
        if ((int)b.lastbits < num_of_bits)
        {
            b.lastbits += 8;
            b.lastbyte = (b.lastbyte << 8) | b.c[b.cnt++];
        }
        b.lastbits -= num_of_bits;
        num |= (b.lastbyte >> b.lastbits) & ((1u << num_of_bits) - 1u);
    }
    return num & mask;
}

// one mixed-radix number  nums[0] + sizes[0]... : the triple is multiplied out byte-wise and sent in num_of_bits bits
inline void sendints(BitBuf& b, int num_of_ints, int num_of_bits, const unsigned sizes[], const unsigned nums[])
{
    unsigned bytes[32], nbytes = 0, bytecnt, tmp;
    tmp = nums[0];
    do { bytes[nbytes++] = tmp & 0xff; tmp >>= 8; } while (tmp != 0);
    for (int i = 1; i < num_of_ints; i++)
    {
        tmp = nums[i];
        for (bytecnt = 0; bytecnt < nbytes; bytecnt++)
        {
            tmp            = bytes[bytecnt] * sizes[i] + tmp;
            bytes[bytecnt] = tmp & 0xff;
            tmp >>= 8;
        }
        while (tmp != 0) { bytes[bytecnt++] = tmp & 0xff; tmp >>= 8; }
        nbytes = bytecnt;
    }
    unsigned i;
    if ((unsigned)num_of_bits >= nbytes * 8)
    {
        for (i = 0; i < nbytes; i++) sendbits(b, 8, bytes[i]);
        sendbits(b, num_of_bits - (int)nbytes * 8, 0);
    }
    else
    {
        for (i = 0; i < nbytes - 1; i++) sendbits(b, 8, bytes[i]);
        sendbits(b, num_of_bits - (int)(nbytes - 1) * 8, bytes[i]);
    }
}

inline void receiveints(BitBuf& b, int num_of_ints, int num_of_bits, const unsigned sizes[], int nums[])
{
    int bytes[32];
    bytes[1] = bytes[2] = bytes[3] = 0;
    int num_of_bytes = 0;
    while (num_of_bits > 8) { bytes[num_of_bytes++] = (int)receivebits(b, 8); num_of_bits -= 8; }
    if (num_of_bits > 0) bytes[num_of_bytes++] = (int)receivebits(b, num_of_bits);
    for (int i = num_of_ints - 1; i > 0; i--)
    {
        unsigned num = 0;
        for (int j = num_of_bytes - 1; j >= 0; j--)
        {
            num            = (num << 8) | (unsigned)bytes[j];
            const unsigned p = num / sizes[i];
            bytes[j]       = (int)p;
            num            = num - p * sizes[i];
        }
        nums[i] = (int)num;
    }
    nums[0] = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24);
}

// the decoder's lambdas must be inlined: out of line, the bit position they capture by reference lives in memory and every
// field pays a store-to-load round trip
#if defined(__GNUC__)
#define XDRTRAJ_INLINE __attribute__((always_inline))
#else
#define XDRTRAJ_INLINE
#endif

// ---- decoder helpers ----------------------------------------------------------------------------------------------------
// q = v / d by one multiply-high with m = floor((2^64 - 1) / d) + 1: m d = 2^64 + e with 0 < e <= d, so
// floor(v m / 2^64) = floor(v / d + v e / (d 2^64)) is the quotient whenever v e < 2^64, i.e. for every v <= lim =
// (2^64 - 1) / d.  The large triples of ordinary systems (three 13-bit ranges: v < 2^39, d < 2^13) stay on this path.
struct XtcDiv
{
    uint32_t d;
    uint64_t m, lim;
    static XtcDiv make(uint32_t d) { return XtcDiv{ d, d > 1 ? (~0ull / d) + 1 : 0, d > 1 ? ~0ull / d : 0 }; }
};

// constants that turn the next n <= 56 stream bits (as little-endian bytes b) into the value sendints() packed:
// (b & mask) are the whole bytes, byte number n / 8 holds the remaining n % 8 bits at its top
struct XtcField
{
    int      n, shift, drop;
    uint64_t mask;
    static XtcField make(int n)
    {
        const int full = n >> 3, rem = n & 7;
        return XtcField{ n, 8 * (full < 8 ? full : 0), 8 - rem, full >= 8 ? ~0ull : ((1ull << (8 * full)) - 1ull) };
    }
};

struct XtcSmallTable     // per adaptive width index: divisor, field constants, half range
{
    XtcDiv   div[128];
    XtcField field[128];
    int      half[128];
    static const XtcSmallTable& get()
    {
        static const XtcSmallTable t = [] {
            XtcSmallTable s{};
            for (int k = 0; k < LASTIDX && k < 128; ++k)
            {
                s.div[k]   = XtcDiv::make((uint32_t)magicints[k]);
                s.field[k] = XtcField::make(k);
                s.half[k]  = magicints[k] / 2;
            }
            return s;
        }();
        return t;
    }
};

// ---- XTC coordinate block: decode --------------------------------------------------------------------------------------
// p points at the second natoms field of a frame (after magic, natoms, step, time, box); returns bytes consumed, 0 on error
inline size_t xtc_decode_coords(const unsigned char* p, size_t avail, int natoms, float* x, float* precision_out = nullptr)
{
    if (avail < 4) return 0;
    const int lsize = (int)be32(p);
    if (lsize != natoms) return 0;
    size_t off = 4;
    if (natoms <= 9)
    {
        if (avail < off + (size_t)natoms * 12) return 0;
        for (int i = 0; i < natoms * 3; ++i) x[i] = be_float(p + off + 4 * (size_t)i);
        if (precision_out) *precision_out = -1.f;
        return off + (size_t)natoms * 12;
    }
    if (avail < off + 4 + 24 + 8) return 0;
    const float precision = be_float(p + off); off += 4;
    if (precision_out) *precision_out = precision;
    int minint[3], maxint[3];
    for (int k = 0; k < 3; ++k) { minint[k] = (int)be32(p + off); off += 4; }
    for (int k = 0; k < 3; ++k) { maxint[k] = (int)be32(p + off); off += 4; }
    int smallidx = (int)be32(p + off); off += 4;
    const size_t nbytes = be32(p + off); off += 4;
    const size_t padded = (nbytes + 3) & ~(size_t)3;
    if (avail < off + padded || smallidx < FIRSTIDX || smallidx >= LASTIDX) return 0;

    unsigned sizeint[3], bitsizeint[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; ++k) sizeint[k] = (unsigned)maxint[k] - (unsigned)minint[k] + 1u;   // unsigned: a corrupt header must not be UB
    int bitsize;
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffff)
    {
        for (int k = 0; k < 3; ++k) bitsizeint[k] = (unsigned)sizeofint(sizeint[k]);
        bitsize = 0;   // flag the use of large sizes
    }
    else bitsize = sizeofints(3, sizeint);

    // Decoder state.  The classic implementation reads the stream one byte at a time, undoes the mixed-radix packing by long
    // division over a byte array and carries the adaptive width along in three variables; here
    //  * a field of up to 56 bits comes out of one unaligned 64-bit load at the current bit position (no refill branch),
    //    and the "whole bytes least significant first, then the remaining high bits" order of sendints() is one byte swap,
    //    one mask and one shift with constants that depend only on the field width;
    //  * a packed triple of at most 64 bits is unpacked with two multiply-high divisions by an invariant divisor;
    //  * everything that depends on the adaptive width `smallidx` (divisor, field constants, offset) is a table look-up,
    //    and the run-length flag is consumed without a branch.
    // Same integers as the classic code.
    // One loop iteration consumes < 120 bytes and the window load reads 8 bytes from the current position (the position is
    // checked at the top of every iteration): with kSlack readable bytes behind the block (a frame inside a mapped file or
    // a larger buffer) the stream is decoded in place, otherwise from a zero-padded copy.
    constexpr size_t kSlack = 192;
    std::vector<unsigned char> data;
    const unsigned char*       q = p + off;
    if (avail < off + padded + kSlack)
    {
        data.assign(padded + kSlack, 0);
        std::memcpy(data.data(), p + off, nbytes);
        q = data.data();
    }
    size_t bitpos = 0;
    auto window = [&]() XDRTRAJ_INLINE -> uint64_t {                          // the next >= 57 bits of the stream, left-aligned
        uint64_t w;
        std::memcpy(&w, q + (bitpos >> 3), 8);
        return __builtin_bswap64(w) << (bitpos & 7);
    };
    auto bits = [&](int n) XDRTRAJ_INLINE -> uint32_t {                       // 0 <= n <= 32
        if (n == 0) return 0;
        const uint32_t v = (uint32_t)(window() >> (64 - n));
        bitpos += (size_t)n;
        return v;
    };
    const XtcField fbig = XtcField::make(bitsize);
    auto packed56 = [&](const XtcField& f) XDRTRAJ_INLINE -> uint64_t {       // f.n <= 56
        const uint64_t b = __builtin_bswap64(window());        // stream bytes in little-endian order
        bitpos += (size_t)f.n;
        return (b & f.mask) | ((((b >> f.shift) & 0xffu) >> f.drop) << f.shift);
    };
    // the value of an n-bit field of any width, 32 bits at a time
    auto packed = [&](int n) -> uint64_t {                     // 0 < n <= 64
        uint64_t v = 0;
        int      sh = 0;
        for (; n >= 32; n -= 32, sh += 32) v |= (uint64_t)__builtin_bswap32(bits(32)) << sh;
        if (n > 0) v |= packed56(XtcField::make(n)) << sh;
        return v;
    };
    auto divmod = [](uint64_t v, const XtcDiv& dv, uint32_t& r) XDRTRAJ_INLINE -> uint64_t {
        if (v > dv.lim) { const uint64_t qq = dv.d ? v / dv.d : 0; r = (uint32_t)(v - qq * dv.d); return qq; }
        const uint64_t qq = (uint64_t)(((unsigned __int128)v * dv.m) >> 64);
        r = (uint32_t)(v - qq * dv.d);
        return qq;
    };
    const XtcDiv dl1 = XtcDiv::make(sizeint[1]), dl2 = XtcDiv::make(sizeint[2]);
    auto triple = [&](const XtcField& f, const XtcDiv& d1, const XtcDiv& d2, int out[3]) XDRTRAJ_INLINE {
        if (f.n <= 64)
        {
            uint64_t v = f.n <= 56 ? packed56(f) : packed(f.n);
            uint32_t r;
            v = divmod(v, d2, r); out[2] = (int)r;
            v = divmod(v, d1, r); out[1] = (int)r;
            out[0] = (int)v;
            return;
        }
        // wider than 64 bits (coordinate ranges close to 2^24 per component): 128-bit arithmetic
        unsigned __int128 v = 0;
        int sh = 0, n = f.n;
        for (; n >= 32; n -= 32, sh += 32) v |= (unsigned __int128)__builtin_bswap32(bits(32)) << sh;
        if (n > 0) v |= (unsigned __int128)packed(n) << sh;
        out[2] = (int)(v % d2.d); v /= d2.d;
        out[1] = (int)(v % d1.d); v /= d1.d;
        out[0] = (int)v;
    };
    const XtcSmallTable& small = XtcSmallTable::get();

    const float inv = 1.0f / precision;
    float*      lfp = x;
    int         run = 0, i = 0;
    int         thiscoord[3], prevcoord[3];
    while (i < lsize)
    {
        if ((bitpos >> 3) > nbytes) return 0;   // corrupt stream
        if (bitsize == 0)
        {
            thiscoord[0] = (int)bits((int)bitsizeint[0]);
            thiscoord[1] = (int)bits((int)bitsizeint[1]);
            thiscoord[2] = (int)bits((int)bitsizeint[2]);
        }
        else triple(fbig, dl1, dl2, thiscoord);
        i++;
        // (sums in unsigned arithmetic: identical for valid files, and a corrupt stream wraps instead of overflowing)
        for (int k = 0; k < 3; ++k) { thiscoord[k] = (int)((unsigned)thiscoord[k] + (unsigned)minint[k]); prevcoord[k] = thiscoord[k]; }
        // one flag bit, and five more bits (new run length and width change) only when it is set
        const uint32_t six  = (uint32_t)(window() >> 58);
        const int      flag = (int)(six >> 5), r5 = (int)(six & 31u), m3 = r5 % 3;
        bitpos += (size_t)(1 + 5 * flag);
        const int is_smaller = (m3 - 1) & -flag;
        run ^= (run ^ (r5 - m3)) & -flag;
        if (run > 0)
        {
            if (i + run / 3 > lsize) return 0;
            const XtcDiv&   ds = small.div[smallidx];
            const XtcField& fs = small.field[smallidx];
            const int       smallnum = small.half[smallidx];
            for (int k = 0; k < run; k += 3)
            {
                triple(fs, ds, ds, thiscoord);
                i++;
                for (int d = 0; d < 3; ++d) thiscoord[d] = (int)((unsigned)thiscoord[d] + (unsigned)prevcoord[d] - (unsigned)smallnum);
                if (k == 0)
                {
                    // the first two atoms of a run were interchanged by the writer (water: O after its first H)
                    for (int d = 0; d < 3; ++d) std::swap(thiscoord[d], prevcoord[d]);
                    *lfp++ = prevcoord[0] * inv; *lfp++ = prevcoord[1] * inv; *lfp++ = prevcoord[2] * inv;
                }
                else
                    for (int d = 0; d < 3; ++d) prevcoord[d] = thiscoord[d];
                *lfp++ = thiscoord[0] * inv; *lfp++ = thiscoord[1] * inv; *lfp++ = thiscoord[2] * inv;
            }
        }
        else
        {
            *lfp++ = thiscoord[0] * inv; *lfp++ = thiscoord[1] * inv; *lfp++ = thiscoord[2] * inv;
        }
        smallidx += is_smaller;                 // the offset and divisor of the small triples follow through the tables
        if (smallidx < FIRSTIDX || smallidx >= LASTIDX) return 0;
    }
    return off + padded;
}

// ---- XTC coordinate block: encode ----------------------------------------------------------------------------------------
inline bool xtc_encode_coords(Writer& w, int natoms, const float* x, float precision)
{
    w.i32(natoms);
    if (natoms <= 9)
    {
        for (int i = 0; i < natoms * 3; ++i) w.f32(x[i]);
        return true;
    }
    if (precision <= 0) precision = 1000.f;
    w.f32(precision);
    std::vector<int> ip((size_t)natoms * 3);
    int minint[3] = { INT_MAX, INT_MAX, INT_MAX }, maxint[3] = { INT_MIN, INT_MIN, INT_MIN }, mindiff = INT_MAX;
    int old[3] = { 0, 0, 0 };
    for (int a = 0; a < natoms; ++a)
    {
        int l[3];
        for (int k = 0; k < 3; ++k)
        {
            float lf = x[3 * (size_t)a + k] * precision;
            lf += lf >= 0 ? 0.5f : -0.5f;
            if (std::fabs(lf) > (float)(INT_MAX - 2)) return false;   // scaling would overflow an int
            l[k] = (int)lf;
            minint[k] = std::min(minint[k], l[k]);
            maxint[k] = std::max(maxint[k], l[k]);
            ip[3 * (size_t)a + k] = l[k];
        }
        const int diff = std::abs(old[0] - l[0]) + std::abs(old[1] - l[1]) + std::abs(old[2] - l[2]);
        if (diff < mindiff && a > 0) mindiff = diff;
        old[0] = l[0]; old[1] = l[1]; old[2] = l[2];
    }
    for (int k = 0; k < 3; ++k) w.i32(minint[k]);
    for (int k = 0; k < 3; ++k) w.i32(maxint[k]);
    for (int k = 0; k < 3; ++k)
        if ((float)maxint[k] - (float)minint[k] >= (float)(INT_MAX - 2)) return false;
    unsigned sizeint[3], bitsizeint[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; ++k) sizeint[k] = (unsigned)(maxint[k] - minint[k] + 1);
    int bitsize;
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffff)
    {
        for (int k = 0; k < 3; ++k) bitsizeint[k] = (unsigned)sizeofint(sizeint[k]);
        bitsize = 0;
    }
    else bitsize = sizeofints(3, sizeint);
    int smallidx = FIRSTIDX;
    while (smallidx < LASTIDX - 1 && magicints[smallidx] < mindiff) smallidx++;
    w.i32(smallidx);
    const int maxidx = std::min(LASTIDX - 1, smallidx + 8), minidx = maxidx - 8;   // often this equals smallidx
    int       smaller = magicints[std::max(FIRSTIDX, smallidx - 1)] / 2, smallnum = magicints[smallidx] / 2;
    unsigned  sizesmall[3] = { (unsigned)magicints[smallidx], (unsigned)magicints[smallidx], (unsigned)magicints[smallidx] };
    const int larger = magicints[maxidx] / 2;

    std::vector<unsigned char> out((size_t)natoms * 3 * 5 + 64, 0);
    BitBuf   b{ out.data(), out.size() };
    int      prevrun = -1, prevcoord[3] = { 0, 0, 0 };
    unsigned tmpcoord[30];
    int      i = 0;
    while (i < natoms)
    {
        int  is_small = 0, is_smaller;
        int* thiscoord = &ip[(size_t)i * 3];
        if (smallidx < maxidx && i >= 1 && std::abs(thiscoord[0] - prevcoord[0]) < larger && std::abs(thiscoord[1] - prevcoord[1]) < larger
            && std::abs(thiscoord[2] - prevcoord[2]) < larger)
            is_smaller = 1;
        else if (smallidx > minidx) is_smaller = -1;
        else is_smaller = 0;
        if (i + 1 < natoms)
        {
            if (std::abs(thiscoord[0] - thiscoord[3]) < smallnum && std::abs(thiscoord[1] - thiscoord[4]) < smallnum
                && std::abs(thiscoord[2] - thiscoord[5]) < smallnum)
            {
                // interchange first with second atom for better compression of water molecules
                for (int d = 0; d < 3; ++d) std::swap(thiscoord[d], thiscoord[d + 3]);
                is_small = 1;
            }
        }
        tmpcoord[0] = (unsigned)(thiscoord[0] - minint[0]);
        tmpcoord[1] = (unsigned)(thiscoord[1] - minint[1]);
        tmpcoord[2] = (unsigned)(thiscoord[2] - minint[2]);
        if (bitsize == 0)
        {
            sendbits(b, (int)bitsizeint[0], tmpcoord[0]);
            sendbits(b, (int)bitsizeint[1], tmpcoord[1]);
            sendbits(b, (int)bitsizeint[2], tmpcoord[2]);
        }
        else sendints(b, 3, bitsize, sizeint, tmpcoord);
        for (int d = 0; d < 3; ++d) prevcoord[d] = thiscoord[d];
        thiscoord += 3;
        i++;
        int run = 0;
        if (is_small == 0 && is_smaller == -1) is_smaller = 0;
        while (is_small && run < 8 * 3)
        {
            const long long d0 = thiscoord[0] - prevcoord[0], d1 = thiscoord[1] - prevcoord[1], d2 = thiscoord[2] - prevcoord[2];
            if (is_smaller == -1 && d0 * d0 + d1 * d1 + d2 * d2 >= (long long)smaller * smaller) is_smaller = 0;
            tmpcoord[run++] = (unsigned)(thiscoord[0] - prevcoord[0] + smallnum);
            tmpcoord[run++] = (unsigned)(thiscoord[1] - prevcoord[1] + smallnum);
            tmpcoord[run++] = (unsigned)(thiscoord[2] - prevcoord[2] + smallnum);
            for (int d = 0; d < 3; ++d) prevcoord[d] = thiscoord[d];
            i++;
            thiscoord += 3;
            is_small = 0;
            if (i < natoms && std::abs(thiscoord[0] - prevcoord[0]) < smallnum && std::abs(thiscoord[1] - prevcoord[1]) < smallnum
                && std::abs(thiscoord[2] - prevcoord[2]) < smallnum)
                is_small = 1;
        }
        if (run != prevrun || is_smaller != 0)
        {
            prevrun = run;
            sendbits(b, 1, 1);   // flag the change in run length
            sendbits(b, 5, (unsigned)(run + is_smaller + 1));
        }
        else sendbits(b, 1, 0);
        for (int k = 0; k < run; k += 3) sendints(b, 3, smallidx, sizesmall, &tmpcoord[k]);
        if (is_smaller != 0)
        {
            smallidx += is_smaller;
            if (is_smaller < 0)
            {
                smallnum = smaller;
                smaller  = magicints[smallidx - 1] / 2;
            }
            else
            {
                smaller  = smallnum;
                smallnum = magicints[smallidx] / 2;
            }
            sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
        }
    }
    size_t nbytes = b.cnt;
    if (b.lastbits != 0) nbytes++;
    w.i32((int)nbytes);
    w.opaque(out.data(), nbytes);
    return true;
}

// ---- frame headers --------------------------------------------------------------------------------------------------------
struct XtcHeader { int natoms, step; float time; float box[9]; };
constexpr size_t kXtcHeaderBytes = 4 * 4 + 9 * 4;

inline bool xtc_parse_header(const unsigned char* p, XtcHeader* h)
{
    if ((int)be32(p) != kXtcMagic) return false;
    h->natoms = (int)be32(p + 4);
    h->step   = (int)be32(p + 8);
    h->time   = be_float(p + 12);
    for (int k = 0; k < 9; ++k) h->box[k] = be_float(p + 16 + 4 * k);
    return true;
}

// bytes of the coordinate block that follows an XTC header, from its first 40 bytes (all that is needed to skip a frame)
inline size_t xtc_coord_block_bytes(const unsigned char* p, int natoms)
{
    if (natoms <= 9) return 4 + (size_t)natoms * 12;
    const size_t nbytes = be32(p + 4 + 4 + 24 + 4);
    return 4 + 4 + 24 + 4 + 4 + ((nbytes + 3) & ~(size_t)3);
}

inline void xtc_write_frame(Writer& w, int natoms, int step, float time, const float box[9], const float* x, float precision)
{
    w.i32(kXtcMagic); w.i32(natoms); w.i32(step); w.f32(time);
    for (int k = 0; k < 9; ++k) w.f32(box[k]);
    xtc_encode_coords(w, natoms, x, precision);
}

struct TrrHeader
{
    int  ir_size, e_size, box_size, vir_size, pres_size, top_size, sym_size, x_size, v_size, f_size, natoms, step, nre;
    bool dbl;
    double t, lambda;
    size_t header_bytes;   // up to and including lambda
};

// p must hold at least 128 bytes
inline bool trr_parse_header(const unsigned char* p, size_t avail, TrrHeader* h)
{
    if (avail < 96 || (int)be32(p) != kTrrMagic) return false;
    size_t off = 4;
    int    slen = (int)be32(p + off); off += 4;           // strlen(version) + 1
    int    n2   = (int)be32(p + off); off += 4;
    if (n2 == slen) { n2 = (int)be32(p + off); off += 4; } // tolerate writers that repeat strlen + 1 before the XDR string
    if (n2 < 0 || n2 > 64) return false;
    off += ((size_t)n2 + 3) & ~(size_t)3;
    int* f[13] = { &h->ir_size, &h->e_size, &h->box_size, &h->vir_size, &h->pres_size, &h->top_size, &h->sym_size,
                   &h->x_size, &h->v_size, &h->f_size, &h->natoms, &h->step, &h->nre };
    if (avail < off + 13 * 4 + 16) return false;
    for (int k = 0; k < 13; ++k) { *f[k] = (int)be32(p + off); off += 4; }
    int real_size = 0;
    if (h->box_size) real_size = h->box_size / 9;
    else if (h->natoms > 0)
    {
        const int any = h->x_size ? h->x_size : (h->v_size ? h->v_size : h->f_size);
        real_size     = any / (h->natoms * 3);
    }
    if (real_size != 4 && real_size != 8) return false;
    h->dbl = real_size == 8;
    // every block is either absent or exactly as large as natoms says (readers index the body by these sizes)
    const long long n3rs = (long long)h->natoms * 3 * real_size;
    for (int blk : { h->x_size, h->v_size, h->f_size })
        if (blk != 0 && blk != n3rs) return false;
    if (h->natoms < 0 || (h->box_size != 0 && h->box_size != 9 * real_size)) return false;
    for (int blk : { h->ir_size, h->e_size, h->vir_size, h->pres_size, h->top_size, h->sym_size })
        if (blk < 0) return false;
    if (h->dbl) { h->t = be_double(p + off); h->lambda = be_double(p + off + 8); off += 16; }
    else { h->t = be_float(p + off); h->lambda = be_float(p + off + 4); off += 8; }
    h->header_bytes = off;
    return true;
}

inline size_t trr_body_bytes(const TrrHeader& h)
{
    return (size_t)h.ir_size + h.e_size + h.box_size + h.vir_size + h.pres_size + h.top_size + h.sym_size + h.x_size + h.v_size + h.f_size;
}

inline void trr_write_frame(Writer& w, int natoms, int step, float time, const float box[9], const float* x, const float* v, const float* f)
{
    static const char version[] = "GMX_trn_file";
    w.i32(kTrrMagic);
    w.i32(13); w.i32(12);                                  // strlen + 1 (checked by readers), then the XDR string: length, bytes
    w.opaque(reinterpret_cast<const unsigned char*>(version), 12);
    const int n3 = natoms * 3 * 4;
    w.i32(0); w.i32(0); w.i32(36); w.i32(0); w.i32(0); w.i32(0); w.i32(0);
    w.i32(x ? n3 : 0); w.i32(v ? n3 : 0); w.i32(f ? n3 : 0);
    w.i32(natoms); w.i32(step); w.i32(0);
    w.f32(time); w.f32(0.f);
    for (int k = 0; k < 9; ++k) w.f32(box[k]);
    for (const float* a : { x, v, f })
        if (a)
            for (int i = 0; i < natoms * 3; ++i) w.f32(a[i]);
}

// reads one real (float or double, big-endian) array of n values into floats
inline void trr_read_reals(const unsigned char* p, bool dbl, size_t n, float* out)
{
    if (dbl) for (size_t i = 0; i < n; ++i) out[i] = (float)be_double(p + 8 * i);
    else for (size_t i = 0; i < n; ++i) out[i] = be_float(p + 4 * i);
}

} // namespace xdrtraj
