/*
 * ORACLE / TEST INFRASTRUCTURE ONLY — not part of the product path.
 *
 * Two sources of uniforms for the CPU oracle:
 *  - tape mode: pops a recorded list of the uniforms the reference itself consumed
 *    (oracle/ref_loader.py UniformTape), in the reference's own sequential order — this is
 *    how the oracle is pinned bit-for-bit against the reference's Python with noise ON;
 *  - Philox mode: the counter-based stream layout of include/dynenv_rng.h, so that the CUDA
 *    kernels and the oracle consume identical draws.
 * Philox4x32-10 restated from Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
 * as 1, 2, 3" (SC'11): multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key bumps 0x9E3779B9 /
 * 0xBB67AE85, 10 rounds.
 */
#ifndef ORACLE_RNG_H
#define ORACLE_RNG_H
#include <stdint.h>

typedef struct {
  int mode;            /* 0 = philox, 1 = tape */
  const double* tape;
  int tape_n, tape_pos;
  int tape_underflow;
  uint64_t seed;
  uint32_t env_id;
} OracleRng;

static inline void oracle_philox(uint32_t k0, uint32_t k1, const uint32_t c[4], uint32_t out[4]) {
  uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* uniform k of (episode, tick, site, a, b) */
static inline double oracle_uniform(OracleRng* g, uint32_t episode, uint32_t tick, int site, int a, int b,
                                    int k) {
  if (g->mode == 1) {
    if (g->tape_pos >= g->tape_n) { g->tape_underflow = 1; return 0.0; }
    return g->tape[g->tape_pos++];
  }
  uint32_t ctr[4] = {g->env_id, episode, tick,
                     ((uint32_t)site << 24) | ((uint32_t)(a & 255) << 16) | ((uint32_t)(b & 255) << 8) |
                         (uint32_t)(k / 6)};
  uint32_t o[4];
  oracle_philox((uint32_t)g->seed, (uint32_t)(g->seed >> 32), ctr, o);
  /* 21-bit slot (k % 6) of the 128-bit little-endian block (include/dynenv_rng.h) */
  int bit = 21 * (k % 6);
  int wi = bit >> 5, sh = bit & 31;
  uint64_t two = (uint64_t)o[wi] | ((uint64_t)(wi < 3 ? o[wi + 1] : 0) << 32);
  uint32_t slot = (uint32_t)(two >> sh) & 0x1FFFFFu;
  return (double)slot * (1.0 / 2097152.0);
}

#endif
