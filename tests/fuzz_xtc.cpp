// TEST INFRASTRUCTURE: corrupt-input fuzzer for the stand-in's XTC decoder (gmxstub/xdrtraj.hpp).  Valid coordinate blocks
// are produced by the encoder, then bits are flipped (header fields and stream), the block is truncated, and the decoder is
// run on exact-size heap buffers -- with and without the slack that lets it decode in place -- under AddressSanitizer and
// UBSan (tests/test_trajfmt.py builds it with -fsanitize=address,undefined).  It must either decode or return 0.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include "xdrtraj.hpp"
int main(int argc, char** argv)
{
    const int iters = argc > 1 ? std::atoi(argv[1]) : 20000;
    std::mt19937 rng(12345);
    long ok = 0, rejected = 0;
    for (int iter = 0; iter < iters; ++iter)
    {
        const int natoms = 10 + (int)(rng() % 60);
        std::vector<float> x(natoms * 3);
        const float span = (iter % 3 == 0) ? 3.f : (iter % 3 == 1 ? 800.f : 30000.f);
        float cx = 0, cy = 0, cz = 0;
        for (int i = 0; i < natoms; ++i)
        {
            if (i % 3 == 0) { cx = span * ((rng() % 2000) / 1000.f - 1); cy = span * ((rng() % 2000) / 1000.f - 1); cz = span * ((rng() % 2000) / 1000.f - 1); }
            x[3 * i] = cx + 0.1f * ((rng() % 200) / 100.f - 1); x[3 * i + 1] = cy + 0.1f * ((rng() % 200) / 100.f - 1); x[3 * i + 2] = cz + 0.1f * ((rng() % 200) / 100.f - 1);
        }
        xdrtraj::Writer w;
        if (!xdrtraj::xtc_encode_coords(w, natoms, x.data(), 1000.f)) continue;
        std::vector<unsigned char> blk = w.buf;
        const int nflip = (int)(rng() % 4);
        for (int f = 0; f < nflip; ++f)
        {
            size_t at = (rng() % 3 == 0) ? rng() % std::min<size_t>(44, blk.size()) : rng() % blk.size();
            blk[at] ^= (unsigned char)(1u << (rng() % 8));
        }
        // exact-size heap buffers: ASAN flags any read past `avail`
        const size_t slack = (rng() % 2) ? 0 : 192 + rng() % 64;
        size_t cut = blk.size();
        if (rng() % 5 == 0) cut = rng() % blk.size();     // truncated input
        std::vector<unsigned char> in(cut + slack, 0xAB);
        if (cut) std::memcpy(in.data(), blk.data(), cut);
        std::vector<float> out(natoms * 3);
        const size_t used = xdrtraj::xtc_decode_coords(in.data(), in.size(), natoms, out.data());
        if (used) ok++; else rejected++;
    }
    printf("decoded %ld rejected %ld\n", ok, rejected);
}
