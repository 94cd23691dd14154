/*
 * dynenv_b200 — shared device helpers: warp-per-environment conventions, Philox4x32-10
 * counter RNG with the stream layout of include/dynenv_rng.h, small 2-D vector math.
 *
 * Execution model of every simulation kernel in this directory: ONE WARP owns ONE
 * environment.  The environment's state record lives in that warp's slice of shared
 * memory for the whole step (all substeps + all observation ticks); HBM is touched only
 * at the step boundary (coalesced record load/store, action load, observation store).
 * Lane-parallel phases map lanes to bodies / robots / sighted objects; phases that the
 * reference executes as ordered side effects (collision callbacks, falls, penalties,
 * Gauss-Seidel contact impulses) run on lane 0 in the reference's order.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dynenv_b200.h"
#include "../../include/dynenv_rng.h"

#define FULLMASK 0xffffffffu

namespace dynenv {

struct SimCfg {
  int n_envs, n_players, P;
  int obs_type, noise_type, continuous;
  float noise_mag, rand_base;
  int random_init, deterministic_turn, can_fall, use_obs_rewards, auto_reset;
  uint32_t seed_lo, seed_hi;
  uint32_t env_id_offset;
  int sub_begin, sub_end; /* substeps run by one step call: 0..RC_SUBSTEPS (a debugging aid narrows it) */
  int sparse_obs;         /* DYNENV_FLAG_SPARSE_OBS */
};

struct V2 {
  float x, y;
};
__device__ __forceinline__ V2 v2(float x, float y) {
  V2 r;
  r.x = x;
  r.y = y;
  return r;
}
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
__device__ __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
/* one fixed evaluation order, so that equal vectors always have equal lengths */
__device__ __forceinline__ float lensq(V2 a) { return __fmaf_rn(a.x, a.x, __fmul_rn(a.y, a.y)); }
__device__ __forceinline__ float len(V2 a) { return sqrtf(lensq(a)); }
__device__ __forceinline__ V2 rot_cs(V2 a, float c, float s) { return v2(a.x * c - a.y * s, a.x * s + a.y * c); }
/* sin and cos for the bounded angles of the simulation (|x| < ~1e4 rad): three-term Cody-Waite
   reduction by pi/2 and degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4] (~1 ulp).  Unlike
   sincosf there is no huge-argument path, which would be inlined -- local-memory table included --
   at every call site of a kernel that is already instruction-cache bound. */
__device__ __forceinline__ void sincos_b(float x, float* sn, float* cs) {
  float k = rintf(x * 0.636619772367581343f);
  float r = fmaf(k, -1.57079601287841797e+00f, x);
  r = fmaf(k, -3.13916473e-07f, r);
  r = fmaf(k, -5.39030253e-15f, r);
  int q = (int)k;
  float r2 = r * r;
  float ps = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
  float pc = fmaf(fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2, -0.5f), r2, 1.0f);
  float s0 = (q & 1) ? pc : ps, c0 = (q & 1) ? ps : pc;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
}
__device__ __forceinline__ V2 rot(V2 a, float ang) {
  float s, c;
  sincos_b(ang, &s, &c);
  return rot_cs(a, c, s);
}

/* the same, out of line: for rare call sites inside hot loops (instruction-cache footprint) */
__device__ __noinline__ V2 rot_ni(V2 a, float ang) { return rot(a, ang); }

/* Philox4x32-10 (Salmon et al., SC'11) */
__device__ __noinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

/* six 21-bit uniforms of one block (include/dynenv_rng.h) */
struct Rand6 {
  uint32_t w[4];
  __device__ __forceinline__ float u(int s) const {
    uint32_t v;
    switch (s) {
      case 0: v = w[0]; break;
      case 1: v = __funnelshift_r(w[0], w[1], 21); break;
      case 2: v = w[1] >> 10; break;
      case 3: v = __funnelshift_r(w[1], w[2], 31); break;
      case 4: v = __funnelshift_r(w[2], w[3], 20); break;
      default: v = w[3] >> 9; break;
    }
    return (float)(v & 0x1FFFFFu) * (1.0f / 2097152.0f);
  }
};

/* The CTA's dynamic shared memory: [0, CFG_WORDS) a copy of the SimCfg, then one ENV_SMEM_WORDS slice
   per warp.  Everything is addressed through this symbol so that the compiler emits LDS/STS (a pointer
   handed to a non-inlined function would degrade to generic loads). */
extern __shared__ __align__(16) uint32_t g_smem[];
constexpr int CFG_WORDS = 20; /* a multiple of 4: the slices behind it are copied with 128-bit accesses */
static_assert(sizeof(SimCfg) <= CFG_WORDS * 4, "SimCfg must fit its shared-memory slot");

#ifdef DYNENV_BOUNDS_CHECK
/* Debug build (tools/gpu_bounds.sh): every access through SharedF / SharedI and every bit-plane word of the raster is
   checked against the block's dynamic shared memory; violations are counted and read by dynenv_debug_bounds_violations.
   (One environment per block, so the block's allocation is exactly one slice.) */
__device__ unsigned long long g_bounds_violations[4]; /* count, first word index, first block, first thread */
__device__ __forceinline__ void bounds_check(int word) {
  unsigned bytes;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(bytes));
  if (word < 0 || (unsigned)word * 4u >= bytes) {
    if (atomicAdd(&g_bounds_violations[0], 1ull) == 0) {
      g_bounds_violations[1] = (unsigned long long)(long long)word;
      g_bounds_violations[2] = blockIdx.x;
      g_bounds_violations[3] = threadIdx.x;
    }
  }
}
#define DYNENV_BC(word) bounds_check(word)
#else
#define DYNENV_BC(word)
#endif
/* One warp per thread block in every kernel of dynenv_api.cu: the slice starts right behind the
   SimCfg copy, so every shared-memory address is an immediate or lane * stride + immediate.  The translation unit of the
   handed-over environments (rc_slow.cu, DYNENV_MW_SLICE_WORDS defined) runs several warps per block, each on its own
   slice: there the base is a per-warp register. */
#ifdef DYNENV_MW_SLICE_WORDS
struct SharedF {
  static __device__ __forceinline__ int base_() { return CFG_WORDS + (int)(threadIdx.x >> 5) * DYNENV_MW_SLICE_WORDS; }
  __device__ __forceinline__ float& operator[](int k) const { DYNENV_BC(base_() + k); return reinterpret_cast<float*>(g_smem)[base_() + k]; }
  __device__ __forceinline__ float* operator+(int k) const { return reinterpret_cast<float*>(g_smem) + base_() + k; }
};
struct SharedI {
  static __device__ __forceinline__ int base_() { return SharedF::base_(); }
  __device__ __forceinline__ int& operator[](int k) const { DYNENV_BC(base_() + k); return reinterpret_cast<int*>(g_smem)[base_() + k]; }
};
#else
struct SharedF { /* this warp's slice viewed as float */
  static constexpr int base = CFG_WORDS;
  static __device__ __forceinline__ constexpr int base_() { return base; }
  __device__ __forceinline__ float& operator[](int k) const { DYNENV_BC(base + k); return reinterpret_cast<float*>(g_smem)[base + k]; }
  __device__ __forceinline__ float* operator+(int k) const { return reinterpret_cast<float*>(g_smem) + base + k; }
};
struct SharedI { /* ... and as int */
  static constexpr int base = CFG_WORDS;
  static __device__ __forceinline__ constexpr int base_() { return base; }
  __device__ __forceinline__ int& operator[](int k) const { DYNENV_BC(base + k); return reinterpret_cast<int*>(g_smem)[base + k]; }
};
#endif
struct SharedCfg {
  __device__ __forceinline__ const SimCfg* operator->() const { return reinterpret_cast<const SimCfg*>(g_smem); }
  __device__ __forceinline__ const SimCfg& operator*() const { return *reinterpret_cast<const SimCfg*>(g_smem); }
};
/* Two-float coordinates (hi + lo, include/dynenv_b200.h): accumulate d into (hi, lo) without losing
   the bits that do not fit hi (Knuth two-sum + renormalisation) */
__device__ __forceinline__ void two_add(float& hi, float& lo, float d) {
  float s = __fadd_rn(hi, d);
  float bb = __fsub_rn(s, hi);
  float err = __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(d, bb));
  lo = __fadd_rn(lo, err);
  float h2 = __fadd_rn(s, lo);
  lo = __fsub_rn(lo, __fsub_rn(h2, s));
  hi = h2;
}
/* random.randint(lo, hi) from a uniform, as the oracle derives it (oracle/ref_loader.py) */
__device__ __forceinline__ int randint_from(float u, int lo, int hi) { return lo + (int)(u * (float)(hi - lo + 1)); }

}  // namespace dynenv
