/* Deterministic synthetic trajectory generator; see include/b2a_synth.h. Build with -ffp-contract=off. */
#include "synth_gen.h"

#include <math.h>
#include <stddef.h>

static uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

uint64_t b2a_synth_hash(uint64_t seed, uint64_t frame, uint64_t species, uint64_t mol, uint64_t lane)
{
    uint64_t h = mix64(seed + 0x9e3779b97f4a7c15ull);
    h = mix64(h ^ (frame * 0xd1342543de82ef95ull + 0x632be59bd9b4e019ull));
    h = mix64(h ^ (species * 0xa0761d6478bd642full + 0xe7037ed1a0b428dbull));
    h = mix64(h ^ (mol * 0x8ebc6af09c88c6e3ull + 0x589965cc75374cc3ull));
    h = mix64(h ^ (lane * 0x9e3779b97f4a7c15ull + 0x1d8e4e27c47d124full));
    return h;
}

static double u01(uint64_t h) { return (double)(h >> 40) * (1.0 / 16777216.0); }

int b2a_synth_species(uint64_t seed, int frame, int species, int nmol, int napm, const float* tmpl,
                      const float L[3], float* x_out)
{
    if (nmol < 0 || napm <= 0 || !tmpl || !L || !x_out) return -1;
#pragma omp parallel for schedule(static)
    for (int m = 0; m < nmol; ++m)
    {
        uint64_t lane = 0;
        float    c[3];
        for (int k = 0; k < 3; ++k)
        {
            float u = (float)u01(b2a_synth_hash(seed, (uint64_t)frame, (uint64_t)species, (uint64_t)m, lane++));
            c[k]    = u * L[k];
        }
        /* random unit quaternion by rejection from the 4-cube */
        double q[4], n2;
        do
        {
            n2 = 0.0;
            for (int a = 0; a < 4; ++a)
            {
                q[a] = 2.0 * u01(b2a_synth_hash(seed, (uint64_t)frame, (uint64_t)species, (uint64_t)m, lane++)) - 1.0;
                n2 += q[a] * q[a];
            }
        } while (n2 > 1.0 || n2 < 1e-6);
        double inv = 1.0 / sqrt(n2);
        double w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
        double R[3][3] = {
            { 1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - w * z), 2.0 * (x * z + w * y) },
            { 2.0 * (x * y + w * z), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - w * x) },
            { 2.0 * (x * z - w * y), 2.0 * (y * z + w * x), 1.0 - 2.0 * (x * x + y * y) },
        };
        float* out = x_out + (size_t)m * (size_t)napm * 3;
        for (int j = 0; j < napm; ++j)
        {
            const float* t = tmpl + (size_t)j * 3;
            for (int k = 0; k < 3; ++k)
            {
                double p = (double)c[k] + (R[k][0] * (double)t[0] + R[k][1] * (double)t[1] + R[k][2] * (double)t[2]);
                float  f = (float)p;
                if (f < 0.0f) f += L[k];
                if (f >= L[k]) f -= L[k];
                if (f < 0.0f) f = 0.0f; /* -tiny + L can round to L; keep the [0, L) contract */
                if (f >= L[k]) f = 0.0f;
                out[(size_t)j * 3 + k] = f;
            }
        }
    }
    return 0;
}
