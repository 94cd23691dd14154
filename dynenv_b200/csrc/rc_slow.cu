/*
 * dynenv_b200 -- the RoboCup environments that rc_phys_kernel hands over (falls, penalties, touching or cached pairs:
 * 5-10 % of them per block of ten substeps), as a translation unit of its own.
 *
 * Same simulation code as dynenv_api.cu (rc_sim.cuh, rc_contact.cuh, rc_phys2.cuh), compiled with RC_MW: W warps = W
 * environments per thread block, every warp on its own shared-memory slice (the slice base is a per-warp register here,
 * an immediate in the one-warp-per-block kernels of dynenv_api.cu -- hence the second translation unit), meeting at two
 * block barriers per substep.  The rare code these environments run (ordered robot phase, narrow phase, cached-pair
 * bookkeeping, sequential-impulse sweeps, callbacks: 60 KB of SASS in use) does not fit an SM's 32 KB instruction cache; 32
 * independent warps per SM each fetched their own way through it and the kernel issued at 28 % whatever its occupancy
 * (8, 16 or 32 resident warps: profiles/r02_ncu_rc_phys_slow.txt).  Warps that walk the phases together share the
 * fetches.
 *
 * The namespace is renamed so that the host-side shadows of the device tables of rc_sim.cuh do not collide with those of
 * dynenv_api.cu at link time.
 */
#define dynenv dynenv_mw
#define RC_MW 1
/* = rc::ENV_SMEM_WORDS (asserted below): the per-warp slice stride */
#define DYNENV_MW_SLICE_WORDS 1532
#include <cuda_runtime.h>

#include "rc_sim.cuh"
#include "rc_phys2.cuh"
#include "rc_slow.h"

using namespace dynenv;

static_assert(DYNENV_MW_SLICE_WORDS == rc::ENV_SMEM_WORDS, "slice stride");
static_assert(sizeof(SimCfg) == DYNENV_SLOW_CFG_BYTES, "SimCfg crosses the translation units as bytes");

namespace {

constexpr int W = DYNENV_SLOW_WARPS;

/* Work unit = W environments of one bin (the same hand-over substep and phase): the warps of a block start together at
   their environments' hand-over point and nobody idles through the substeps in front of it.  Units are numbered through
   the bins (every block derives the same prefix over the 20 bin lengths), longest first -- the bins handed over inside
   Space.step (contacts: ~5 x the instructions of the others) in ascending substep order, then those handed over for an
   ordered robot phase -- and the blocks of the persistent grid draw them from a counter: ~400 units on 296 resident
   blocks leave a second round, which is then made of the short ones. */
__global__ void __launch_bounds__(32 * W, 32 / W)
rc_phys_slow_kernel(const __grid_constant__ SimCfg cfg, uint32_t* __restrict__ state, const float* __restrict__ actions, int sub_lo,
                    int sub_e, const uint32_t* __restrict__ slow_bins, uint32_t* __restrict__ slow_count) {
  __shared__ uint32_t s_unit;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < sizeof(SimCfg) / 4) g_smem[threadIdx.x] = reinterpret_cast<const uint32_t*>(&cfg)[threadIdx.x];
  __syncthreads();
  /* lane k < 20: units of the k-th bin in processing order and of the bins in front of it */
  const int binK = lane < DYNENV_SLOW_BINS / 2 ? 2 * lane + 1 : 2 * (lane - DYNENV_SLOW_BINS / 2);
  const uint32_t cntK = lane < DYNENV_SLOW_BINS ? slow_count[DYNENV_SLOW_BIN0 + binK] : 0u;
  const uint32_t unitsK = (cntK + W - 1) / W;
  uint32_t incl = unitsK;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(FULLMASK, incl, d);
    if (lane >= d) incl += v;
  }
  const uint32_t total = __shfl_sync(FULLMASK, incl, 31);
  uint32_t* S = g_smem + CFG_WORDS + warp * DYNENV_MW_SLICE_WORDS;
  for (;;) { /* block-uniform trip count: barriers inside */
    if (threadIdx.x == 0) s_unit = atomicAdd(&slow_count[DYNENV_SLOW_NEXT], 1u);
    __syncthreads();
    const uint32_t u = s_unit;
    __syncthreads();
    if (u >= total) break;
    const int slot = __ffs(__ballot_sync(FULLMASK, u < incl)) - 1;
    const int bin = __shfl_sync(FULLMASK, binK, slot);
    const uint32_t k = (u - (__shfl_sync(FULLMASK, incl, slot) - __shfl_sync(FULLMASK, unitsK, slot))) * W + warp;
    const bool valid = k < __shfl_sync(FULLMASK, cntK, slot);
    const int e = valid ? (int)slow_bins[(size_t)bin * cfg.n_envs + k] : 0;
    uint32_t* rec = state + (size_t)e * RC_STATE_WORDS;
    if (valid) {
      const uint4* src = reinterpret_cast<const uint4*>(rec);
      uint4* dst = reinterpret_cast<uint4*>(S);
      for (int w = lane; w < RC_STATE_WORDS / 4; w += 32) dst[w] = src[w];
    }
    __syncwarp();
    rc::Env E;
    E.lane = lane;
    if (lane == 0) E.i[rc::X_ENVID] = (int)(cfg.env_id_offset + (uint32_t)e);
    __syncwarp();
    const int sub_b = sub_lo + (bin >> 1);
    rc::phys_substeps2<true>(E, actions + (size_t)e * cfg.P * 4, sub_b, sub_e, false, bin & 1, sub_b, valid);
    __syncwarp();
    if (valid) {
      const uint4* src = reinterpret_cast<const uint4*>(S);
      uint4* dst = reinterpret_cast<uint4*>(rec);
      for (int w = lane; w < RC_STATE_WORDS / 4; w += 32) dst[w] = src[w];
    }
    __syncwarp();
  }
}

}  // namespace

#ifdef RC_DBG_COUNTERS
extern "C" int dynenv_slow_debug_counters(unsigned long long* out32, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out32, rc::g_dbg, 16 * sizeof(unsigned long long));
  cudaMemcpyFromSymbol(out32 + 16, rc::g_dbg2, 16 * sizeof(unsigned long long));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(rc::g_dbg, z, sizeof(z)); cudaMemcpyToSymbol(rc::g_dbg2, z, sizeof(z)); }
  return 0;
}
#endif

size_t dynenv_slow_smem_bytes(void) { return ((size_t)CFG_WORDS + (size_t)W * DYNENV_MW_SLICE_WORDS) * sizeof(uint32_t); }

cudaError_t dynenv_slow_configure(void) {
  return cudaFuncSetAttribute(rc_phys_slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynenv_slow_smem_bytes());
}

void dynenv_slow_launch(const void* sim_cfg, uint32_t* state, const float* actions, int sub_lo, int sub_e, const uint32_t* slow_bins,
                        uint32_t* slow_count, int sms, cudaStream_t stream) {
  SimCfg cfg;
  memcpy(&cfg, sim_cfg, sizeof(cfg));
  rc_phys_slow_kernel<<<sms * DYNENV_SLOW_BLOCKS_PER_SM, 32 * W, dynenv_slow_smem_bytes(), stream>>>(cfg, state, actions, sub_lo, sub_e, slow_bins, slow_count);
}
