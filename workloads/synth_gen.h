/*
 * synth_gen.h -- deterministic synthetic trajectory generator (host, plain C ABI).
 *
 * the reference ships no data files (*.xtc/*.tpr are git-ignored in the
 * reference, .gitignore:5-17), so every parity test and benchmark runs on synthetic boxes.  The
 * generator is a pure function of (seed, frame, species, molecule): a splitmix64 counter hash mapped
 * to [0,1) by (u >> 40) * 2^-24, molecule centres uniform in the box, rigid templates rotated by a
 * random unit quaternion, atoms wrapped PER ATOM into [0, L) so molecules straddle the faces (this is
 * what exercises the unwrap in GmxHandle::loadPosition / loadPositionCenter, reference
 * include/itp/gmx_bits/gmx_handle.ipp:139-143, 183-191).  Only + - * / sqrt in IEEE double without
 * FMA contraction are used, so the output is bit-reproducible across hosts and compilers.
 */
#ifndef SYNTH_GEN_H
#define SYNTH_GEN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Fill nmol*napm atoms (x_out[nmol*napm][3], float) of one species for one frame.
 * tmpl: [napm][3] rigid template coordinates (nm) relative to the placement centre.
 * L:    orthorhombic box lengths (nm).  Returns 0, or -1 on bad arguments.            */
int b2a_synth_species(uint64_t seed, int frame, int species, int nmol, int napm, const float* tmpl,
                      const float L[3], float* x_out);

/* The raw hash, exported so tests can pin it. */
uint64_t b2a_synth_hash(uint64_t seed, uint64_t frame, uint64_t species, uint64_t mol, uint64_t lane);

#ifdef __cplusplus
}
#endif
#endif
