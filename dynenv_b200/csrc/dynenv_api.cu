/*
 * dynenv_b200 — kernels' entry points and the C ABI of include/dynenv_b200.h.
 * No torch types here: device pointers and sizes only.  Every call is asynchronous on the
 * caller's stream; nothing in this file synchronises the device.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "drv_sim.cuh"
#include "rc_sim.cuh"
#include "rc_obs.cuh"
#include "rc_phys2.cuh"
#include "rc_slow.h"
#include "rc_raster.cuh"
#include "rollout.cuh"
#include "arrange.cuh"

using namespace dynenv;

namespace {

/* One warp (= one environment) per thread block, 32 blocks = 32 warps per SM at 64 registers per thread.  Measured
   (65536 environments, whole-episode mean): 4 warps x 8 blocks 18.3 ms / step, 2 x 16 17.6, 1 x 32 15.2; fewer
   resident warps (4 x 6, 4 x 4, 2 x 8) and more registers are slower, 48 warps at 40 registers much slower.  A block
   retires as soon as its environment is done, so a slow environment (contacts, falls) does not hold three others'
   slots -- the kernels are latency / instruction-fetch bound and live on resident warps. */
#ifndef RC_WPC_
#define RC_WPC_ 1
#endif
#ifndef RC_MINB_
#define RC_MINB_ 32
#endif
constexpr int RC_WPC = RC_WPC_; /* environments (warps) per CTA */
/* 32 one-warp blocks per SM: 228 KB of shared memory / 32 - 1 KB reserved per block = 6272 bytes */
static_assert((CFG_WORDS + rc::ENV_SMEM_WORDS) * 4 <= 6272 && (CFG_WORDS + drv::ENV_SMEM_WORDS) * 4 <= 6272, "one more block per SM");

thread_local std::string g_err;

/* Every entry point works on its handle's (or its argument's) device and gives the calling thread its own current device
   back: a process that drives several GPUs (or PyTorch in the same thread) must not find its current device switched
   under it -- neither by a step nor by a destroy called from a garbage collector. */
struct DeviceGuard {
  int prev = -1, want;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) : want(dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != want) err = cudaSetDevice(want);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != want) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return DYNENV_E_CUDA;
}

/* coalesced copy of one environment's state record between HBM and the warp's shared memory */
template <int WORDS = RC_STATE_WORDS>
__device__ __forceinline__ void load_record(uint32_t* S, const uint32_t* g, int lane) {
  const uint4* src = reinterpret_cast<const uint4*>(g);
  uint4* dst = reinterpret_cast<uint4*>(S);
  for (int w = lane; w < WORDS / 4; w += 32) dst[w] = src[w];
  __syncwarp();
}
template <int WORDS = RC_STATE_WORDS>
__device__ __forceinline__ void store_record(uint32_t* g, const uint32_t* S, int lane) {
  __syncwarp();
  const uint4* src = reinterpret_cast<const uint4*>(S);
  uint4* dst = reinterpret_cast<uint4*>(g);
  for (int w = lane; w < WORDS / 4; w += 32) dst[w] = src[w];
}

/* the kernel parameter block -> shared memory, where every function of the simulation reads it */
__device__ __forceinline__ void stage_cfg(const SimCfg& cfg) {
  if (threadIdx.x < sizeof(SimCfg) / 4) g_smem[threadIdx.x] = reinterpret_cast<const uint32_t*>(&cfg)[threadIdx.x];
  __syncthreads();
}

__device__ __forceinline__ rc::Env make_env(int base, const SimCfg& cfg, int e, int lane) {
  rc::Env E;
  static_assert(RC_WPC_ == 1, "the slice base is a compile-time constant: one warp per block");
  (void)base;
  E.lane = lane;
  if (lane == 0) E.i[rc::X_ENVID] = (int)(cfg.env_id_offset + (uint32_t)e);
  __syncwarp();
  return E;
}

/*
 * One env step = RC_NTICK x { rc_phys_kernel (10 substeps; the few environments it hands over finish them in
 * rc_phys_slow_kernel on the side stream), the tick's observation kernels (rc_obsp_kernel<true> field lines,
 * rc_obsp_kernel<false> points; rc_obs_kernel for FULL and IMAGE) } + rc_finish_kernel (dynenv_step below).  rc_phys / rc_finish / rc_reset: grid = ceil(nEnvs / RC_WPC) CTAs
 * of RC_WPC warps, warp w of CTA b owns environment b*RC_WPC + w, whose record it loads into its shared-memory slice
 * and stores back (coalesced).  The phases are separate launches because each one's hot code then fits the SM's
 * instruction cache; fused into one kernel the warps of an SM sit in different phases and the step is bound by
 * instruction fetch.
 */
__global__ void __launch_bounds__(RC_WPC * 32, RC_MINB_)
rc_phys_kernel(const __grid_constant__ SimCfg cfg, uint32_t* __restrict__ state, const float* __restrict__ actions,
               int sub_b, int sub_e, int first, uint32_t* __restrict__ slow_list, uint32_t* __restrict__ slow_count,
               uint8_t* __restrict__ slow_mark, uint32_t* __restrict__ slow_bins) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * RC_WPC + warp;
  stage_cfg(cfg);
  if (e >= cfg.n_envs) return;
  const int base = CFG_WORDS + warp * rc::ENV_SMEM_WORDS;
  uint32_t* S = g_smem + base;
  load_record(S, state + (size_t)e * RC_STATE_WORDS, lane);
  rc::Env E = make_env(base, cfg, e, lane);
  const int handover = rc::phys_substeps2<false>(E, actions + (size_t)e * cfg.P * 4, sub_b, sub_e, first != 0);
  store_record(state + (size_t)e * RC_STATE_WORDS, S, lane);
  if (lane == 0) {
    slow_mark[e] = handover >= 0;
    if (handover >= 0) {
      /* the flat list (walked by the handed-over environments' observation launches) and the same environments binned
         by hand-over point, 2 (sub - sub_b) + phase: rc_slow.cu gives a block environments of one bin */
      slow_list[atomicAdd(slow_count, 1u)] = ((uint32_t)handover << 24) | (uint32_t)e;
      const int bin = handover - 2 * sub_b;
      slow_bins[(size_t)bin * cfg.n_envs + atomicAdd(slow_count + DYNENV_SLOW_BIN0 + bin, 1u)] = (uint32_t)e;
    }
  }
}

/* the environments rc_phys_kernel handed over (falls, penalties, touching or cached pairs: a few percent of them): the
   rest of their substeps, still register-resident between the ordered robot phases (rc_sim.cuh) and the contact halves
   (rc_contact.cuh) they need.  A fixed grid walks the list.  This is the first, one-warp-per-block version, kept for
   comparison (-DRC_SLOW_ONE_WARP): the product launches rc_slow.cu's phase-synchronised multi-warp kernel. */
#ifdef RC_SLOW_ONE_WARP
__global__ void __launch_bounds__(32, 32)
rc_phys_slow_kernel(const __grid_constant__ SimCfg cfg, uint32_t* __restrict__ state, const float* __restrict__ actions, int sub_e,
                    const uint32_t* __restrict__ slow_list, const uint32_t* __restrict__ slow_count) {
  const int lane = threadIdx.x;
  stage_cfg(cfg);
  const uint32_t n = *slow_count;
  uint32_t* S = g_smem + CFG_WORDS;
  for (uint32_t k = blockIdx.x; k < n; k += gridDim.x) {
    const uint32_t item = slow_list[k];
    const int e = (int)(item & 0xFFFFFFu), code = (int)(item >> 24);
    load_record(S, state + (size_t)e * RC_STATE_WORDS, lane);
    rc::Env E = make_env(CFG_WORDS, cfg, e, lane);
    rc::phys_substeps2<true>(E, actions + (size_t)e * cfg.P * 4, code >> 1, sub_e, false, code & 1);
    store_record(state + (size_t)e * RC_STATE_WORDS, S, lane);
    __syncwarp();
  }
}
#endif

/* The observation kernels run twice per tick: on the caller's stream for every environment that rc_phys_kernel finished
   itself (slow_mark[e] == 0; slow_list == nullptr), and on the side stream, behind rc_phys_slow_kernel, for the
   handed-over ones (a fixed grid walks slow_list).  next_env returns the block's next environment or -1. */
__device__ __forceinline__ int next_env(int& cursor, int n_envs, const uint8_t* slow_mark, const uint32_t* slow_list, const uint32_t* slow_count) {
  if (slow_list) {
    const uint32_t n = *slow_count;
    if (cursor < 0) cursor = blockIdx.x; else cursor += gridDim.x;
    return (uint32_t)cursor < n ? (int)(slow_list[cursor] & 0xFFFFFFu) : -1;
  }
  if (cursor >= 0) return -1;
  cursor = blockIdx.x;
  if (cursor >= n_envs || (slow_mark && slow_mark[cursor])) return -1;
  return cursor;
}

/* observation tick t of the step in flight (after substep 10 t + 9) */
template <bool RAW>
__global__ void __launch_bounds__(RC_WPC * 32, RC_MINB_)
rc_obs_kernel(const __grid_constant__ SimCfg cfg, uint32_t* __restrict__ state, int t, int subs_left,
              float* __restrict__ obs, float* __restrict__ full, float* __restrict__ term_obs, float* __restrict__ term_full,
              float* __restrict__ raw, float* __restrict__ term_raw, const uint8_t* __restrict__ slow_mark,
              const uint32_t* __restrict__ slow_list, const uint32_t* __restrict__ slow_count) {
  const int lane = threadIdx.x & 31;
  stage_cfg(cfg);
  const int base = CFG_WORDS;
  uint32_t* S = g_smem + base;
  int cursor = -1;
  for (int e = next_env(cursor, cfg.n_envs, slow_mark, slow_list, slow_count); e >= 0;
       e = next_env(cursor, cfg.n_envs, slow_mark, slow_list, slow_count)) {
    load_record(S, state + (size_t)e * RC_STATE_WORDS, lane);
    rc::Env E = make_env(base, cfg, e, lane);
    const int P = cfg.P;
    const size_t obs_stride = (size_t)RC_NTICK * P * RC_OBS_WORDS, full_stride = (size_t)RC_NTICK * P * RC_FULL_WORDS;
    /* a finishing episode's last observations go to the terminal buffers, its first ones (written by the
       finish kernel after the reset) to `out` */
    const bool finishing = cfg.auto_reset && E.i[RC_S_ELAPSED] + subs_left >= rc::MAXTIME;
    float* o = finishing ? (term_obs ? term_obs + e * obs_stride : nullptr) : (obs ? obs + e * obs_stride : nullptr);
    float* f = finishing ? (term_full ? term_full + e * full_stride : nullptr) : (full ? full + e * full_stride : nullptr);
    const size_t raw_stride = (size_t)RC_NTICK * P * RC_RAW_WORDS;
    float* r = finishing ? (term_raw ? term_raw + e * raw_stride : nullptr) : (raw ? raw + e * raw_stride : nullptr);
    rc::observe_tick<RAW>(E, t, (uint32_t)E.i[RC_S_ELAPSED], o, f, r);
    /* only the processSeens sums changed */
    __syncwarp();
    uint32_t* g = state + (size_t)e * RC_STATE_WORDS;
    for (int w = RC_S_STEP_LSUM + lane; w < RC_S_STEP_END; w += 32) g[w] = S[w];
    __syncwarp();
  }
}

/* PARTIAL observations: the flattened tick of rc_obs.cuh.  No copy of the state in shared memory: the robot table is
   built from the HBM record, only the processSeens sums go back. */
struct GlobalRec {
  const uint32_t* g;
  __device__ __forceinline__ uint32_t operator()(int k) const { return g[k]; }
};
struct SharedRec {
  int base;
  __device__ __forceinline__ uint32_t operator()(int k) const { return g_smem[base + k]; }
};

template <bool LINES>
__global__ void __launch_bounds__(32, 32)
rc_obsp_kernel(const __grid_constant__ SimCfg cfg, uint32_t* state, int t, int subs_left, float* __restrict__ obs,
               float* __restrict__ term_obs, const uint8_t* __restrict__ slow_mark, const uint32_t* __restrict__ slow_list,
               const uint32_t* __restrict__ slow_count) {
  const int lane = threadIdx.x;
  stage_cfg(cfg);
  rc::Env E;
  E.lane = lane;
  int cursor = -1;
  for (int e = next_env(cursor, cfg.n_envs, slow_mark, slow_list, slow_count); e >= 0;
       e = next_env(cursor, cfg.n_envs, slow_mark, slow_list, slow_count)) {
    uint32_t* rec = state + (size_t)e * RC_STATE_WORDS;
    const int elapsed = (int)rec[RC_S_ELAPSED];
    rc::o_prologue(E, GlobalRec{rec}, cfg.env_id_offset + (uint32_t)e, !LINES);
    const size_t obs_stride = (size_t)RC_NTICK * cfg.P * RC_OBS_WORDS;
    const bool finishing = cfg.auto_reset && elapsed + subs_left >= rc::MAXTIME;
    float* o = finishing ? (term_obs ? term_obs + e * obs_stride : nullptr) : (obs ? obs + e * obs_stride : nullptr);
    float* ot = o ? o + (size_t)t * cfg.P * RC_OBS_WORDS : nullptr;
    if (LINES) { /* first launch of the tick: the lines; their counts join the processSeens sum of the landmarks */
      rc::observe_lines(E, (uint32_t)elapsed, ot);
      if (lane < cfg.P) rec[RC_S_STEP_LSUM + lane] += (uint32_t)E.i[rc::O_CNT + 8 * lane + 5];
    } else {     /* second launch: the point objects, false positives, counts */
      rc::observe_partial(E, (uint32_t)elapsed, ot);
      for (int w = lane; w < rc::O_ACC_WORDS; w += 32) rec[RC_S_STEP_LSUM + w] = (uint32_t)E.i[rc::O_ACC + w];
    }
    __syncwarp();
  }
}

/* environment_base.reset for PARTIAL observations: new scene, record to HBM, then the flattened observation ticks over
   the slice the record occupied (they need only the robot table, gathered before the overlay) */
__device__ __forceinline__ void reset_env_partial(const rc::Env E, uint32_t* S, uint32_t* rec, uint32_t env_id, float* obs_env,
                                                  float* info_env, bool write_info) {
  const int lane = E.lane;
  rc::init_scene(E);
  rc::clear_step_scratch(E);
  if (info_env && write_info) {
    rc::fill_info(E, info_env);
    if (lane < RC_MAXP) {
      info_env[RC_I_REWARDS + lane] = 0.f;
      info_env[RC_I_EPR + lane] = 0.f; info_env[RC_I_EPPR + lane] = 0.f; info_env[RC_I_EPOR + lane] = 0.f;
    }
    if (lane == 0) {
      int* ii = reinterpret_cast<int*>(info_env);
      ii[RC_I_DONE] = 0; ii[RC_I_FLAGS] = 0; ii[RC_I_EPG] = 0; ii[RC_I_EPG + 1] = 0;
    }
  }
  store_record(rec, S, lane);
  __syncwarp();
  rc::o_prologue(E, SharedRec{SharedI::base}, env_id, false);
  const int P = E.c->P;
  for (int t = 0; t < RC_NTICK; t++) {
    float* ot = obs_env ? obs_env + (size_t)t * P * RC_OBS_WORDS : nullptr;
    rc::observe_lines(E, DYNENV_TICK_RESETOBS | (uint32_t)t, ot);
    rc::observe_partial(E, DYNENV_TICK_RESETOBS | (uint32_t)t, ot);
  }
}

/* rewards, info, done, episode statistics; finished environments are re-initialised (auto-reset) */
template <bool RAW>
__global__ void __launch_bounds__(RC_WPC * 32, RC_MINB_)
rc_finish_kernel(const __grid_constant__ SimCfg cfg, uint32_t* __restrict__ state, float* __restrict__ obs,
                 float* __restrict__ full, float* __restrict__ info, float* __restrict__ raw) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * RC_WPC + warp;
  stage_cfg(cfg);
  if (e >= cfg.n_envs) return;
  const int base = CFG_WORDS + warp * (RAW ? rc::ENV_SMEM_WORDS_RAW : rc::ENV_SMEM_WORDS);
  uint32_t* S = g_smem + base;
  load_record(S, state + (size_t)e * RC_STATE_WORDS, lane);
  rc::Env E = make_env(base, cfg, e, lane);
  const int P = cfg.P;
  const size_t obs_stride = (size_t)RC_NTICK * P * RC_OBS_WORDS, full_stride = (size_t)RC_NTICK * P * RC_FULL_WORDS;
  float* inf = info ? info + (size_t)e * RC_INFO_WORDS : nullptr;
  rc::finish_step(E, inf);
  if (cfg.auto_reset && E.i[RC_S_ELAPSED] >= rc::MAXTIME) {
    if (!RAW && cfg.obs_type == DYNENV_OBS_PARTIAL) {
      reset_env_partial(E, S, state + (size_t)e * RC_STATE_WORDS, cfg.env_id_offset + (uint32_t)e, obs ? obs + e * obs_stride : nullptr,
                        inf, false);
      return;
    }
    rc::reset_env<RAW>(E, obs ? obs + e * obs_stride : nullptr, full ? full + e * full_stride : nullptr, inf, false,
                       raw ? raw + (size_t)e * RC_NTICK * P * RC_RAW_WORDS : nullptr);
  }
  store_record(state + (size_t)e * RC_STATE_WORDS, S, lane);
}

template <bool RAW>
__global__ void __launch_bounds__(RC_WPC * 32, RC_MINB_)
rc_reset_kernel(const __grid_constant__ SimCfg cfg, uint32_t* __restrict__ state, const uint8_t* __restrict__ mask,
                float* __restrict__ obs, float* __restrict__ full, float* __restrict__ info, float* __restrict__ raw) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * RC_WPC + warp;
  stage_cfg(cfg);
  if (e >= cfg.n_envs) return;
  if (mask && !mask[e]) return;
  const int base = CFG_WORDS + warp * (RAW ? rc::ENV_SMEM_WORDS_RAW : rc::ENV_SMEM_WORDS);
  uint32_t* S = g_smem + base;
  load_record(S, state + (size_t)e * RC_STATE_WORDS, lane);
  rc::Env E = make_env(base, cfg, e, lane);
  const int P = cfg.P;
  const size_t obs_stride = (size_t)RC_NTICK * P * RC_OBS_WORDS, full_stride = (size_t)RC_NTICK * P * RC_FULL_WORDS;
  if (!RAW && cfg.obs_type == DYNENV_OBS_PARTIAL) {
    reset_env_partial(E, S, state + (size_t)e * RC_STATE_WORDS, cfg.env_id_offset + (uint32_t)e, obs ? obs + e * obs_stride : nullptr,
                      info ? info + (size_t)e * RC_INFO_WORDS : nullptr, true);
    return;
  }
  rc::reset_env<RAW>(E, obs ? obs + e * obs_stride : nullptr, full ? full + e * full_stride : nullptr,
                     info ? info + (size_t)e * RC_INFO_WORDS : nullptr, true, raw ? raw + (size_t)e * RC_NTICK * P * RC_RAW_WORDS : nullptr);
  store_record(state + (size_t)e * RC_STATE_WORDS, S, lane);
}

/* ------------------------------------------------------------------------------------------- Driving
 * One env step = drv_phys_kernel (10 substeps) + drv_obs_kernel (the observation of every car, rewards, info,
 * done, auto-reset), same warp-per-environment layout as the RoboCup kernels. */
#ifndef DRV_WPC_
#define DRV_WPC_ 1 /* as for RoboCup: 1 x 32 gives 4.0 ms / step, 4 x 4 (95 registers) 5.6 ms */
#endif
#ifndef DRV_MINB_
#define DRV_MINB_ 32
#endif
constexpr int DRV_WPC = DRV_WPC_;
static_assert(DRV_WPC_ == 1, "the slice base is a compile-time constant: one warp per block");

__device__ __forceinline__ drv::Env make_drv_env(int base, const SimCfg& cfg, int e, int lane) {
  drv::Env E;
  (void)base;
  E.lane = lane;
  if (lane == 0) E.i[drv::X_ENVID] = (int)(cfg.env_id_offset + (uint32_t)e);
  __syncwarp();
  return E;
}

__global__ void __launch_bounds__(DRV_WPC * 32, DRV_MINB_)
drv_phys_kernel(const __grid_constant__ SimCfg cfg, uint32_t* __restrict__ state, const float* __restrict__ actions,
                int sub_b, int sub_e, int first) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * DRV_WPC + warp;
  stage_cfg(cfg);
  if (e >= cfg.n_envs) return;
  const int base = CFG_WORDS + warp * drv::ENV_SMEM_WORDS;
  uint32_t* S = g_smem + base;
  load_record<DRV_STATE_WORDS>(S, state + (size_t)e * DRV_STATE_WORDS, lane);
  drv::Env E = make_drv_env(base, cfg, e, lane);
  drv::phys_substeps(E, actions + (size_t)e * cfg.P * 2, sub_b, sub_e, first != 0);
  store_record<DRV_STATE_WORDS>(state + (size_t)e * DRV_STATE_WORDS, S, lane);
}

/* observation after the last substep (DrivingEnvironment.py:290-294), then rewards / info / done (:299-322);
   a finishing episode's observation goes to the terminal buffer and the environment is re-initialised */
__global__ void __launch_bounds__(DRV_WPC * 32, DRV_MINB_)
drv_obs_kernel(const __grid_constant__ SimCfg cfg, uint32_t* __restrict__ state, int observe, float* __restrict__ obs,
               float* __restrict__ info, float* __restrict__ term_obs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * DRV_WPC + warp;
  stage_cfg(cfg);
  if (e >= cfg.n_envs) return;
  const int base = CFG_WORDS + warp * drv::ENV_SMEM_WORDS;
  uint32_t* S = g_smem + base;
  load_record<DRV_STATE_WORDS>(S, state + (size_t)e * DRV_STATE_WORDS, lane);
  drv::Env E = make_drv_env(base, cfg, e, lane);
  const size_t obs_stride = (size_t)cfg.P * DRV_OBS_WORDS;
  const bool finishing = cfg.auto_reset && E.i[DRV_S_ELAPSED] >= drv::MAXTIME;
  float* o = finishing ? (term_obs ? term_obs + e * obs_stride : nullptr) : (obs ? obs + e * obs_stride : nullptr);
  if (observe) drv::observe(E, (uint32_t)E.i[DRV_S_ELAPSED], o);
  float* inf = info ? info + (size_t)e * DRV_INFO_WORDS : nullptr;
  drv::finish_step(E, inf);
  if (finishing) drv::reset_env(E, obs ? obs + e * obs_stride : nullptr, inf, false);
  store_record<DRV_STATE_WORDS>(state + (size_t)e * DRV_STATE_WORDS, S, lane);
}

__global__ void __launch_bounds__(DRV_WPC * 32, DRV_MINB_)
drv_reset_kernel(const __grid_constant__ SimCfg cfg, uint32_t* __restrict__ state, const uint8_t* __restrict__ mask,
                 float* __restrict__ obs, float* __restrict__ info) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * DRV_WPC + warp;
  stage_cfg(cfg);
  if (e >= cfg.n_envs) return;
  if (mask && !mask[e]) return;
  const int base = CFG_WORDS + warp * drv::ENV_SMEM_WORDS;
  uint32_t* S = g_smem + base;
  load_record<DRV_STATE_WORDS>(S, state + (size_t)e * DRV_STATE_WORDS, lane);
  drv::Env E = make_drv_env(base, cfg, e, lane);
  drv::reset_env(E, obs ? obs + (size_t)e * cfg.P * DRV_OBS_WORDS : nullptr, info ? info + (size_t)e * DRV_INFO_WORDS : nullptr, true);
  store_record<DRV_STATE_WORDS>(state + (size_t)e * DRV_STATE_WORDS, S, lane);
}

}  // namespace

constexpr int MAX_TIMED = 24; /* launches of one step */
/* FP32 FMA throughput probe (dynenv_debug_fp32_peak) */
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters, float a, float b) {
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = (float)(threadIdx.x + k);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++) x[k] = fmaf(x[k], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  if (s == 123.456f) out[0] = s; /* never true: keeps the chains alive */
}

struct DynEnvHandle_s {
  DynEnvConfig cfg;
  SimCfg sim;
  uint32_t* state;
  uint32_t* slow_list;   /* RoboCup: environments handed over to rc_phys_slow_kernel, (2 sub + phase) << 24 | env */
  uint32_t* slow_count;  /* [8][DYNENV_SLOW_CNT_WORDS] per substep block of the step: [0] length of the list, [DYNENV_SLOW_BIN0 + k] of bin k */
  uint32_t* slow_bins;   /* [DYNENV_SLOW_BINS][n_envs] the same environments by hand-over point */
  uint8_t* slow_mark;    /* [n_envs] 1 = handed over in the substep block in flight */
  float* act_f32;        /* [n_envs][agents][action words] dynenv_step_u8's widened actions (allocated on first use) */
  cudaStream_t side;     /* the handed-over environments' substeps and observations run here, beside the others' observations */
  cudaEvent_t ev_fork[8], ev_join[8];
  int sms;
  int state_words;
  int last_launches;
  /* optional per-launch timing (dynenv_debug_timing): events bracketing every launch of the last step */
  int timing;
  cudaEvent_t ev[MAX_TIMED + 1];
  int ev_class[MAX_TIMED]; /* 0 substep block, 1 observation tick, 2 finish, 3 substep block of the handed-over environments */
  int n_timed;
};

extern "C" {

int dynenv_abi_version(void) { return DYNENV_ABI_VERSION; }
const char* dynenv_last_error(void) { return g_err.c_str(); }

int dynenv_create(const DynEnvConfig* cfg, DynEnvHandle* out) {
  if (!cfg || !out) return fail(DYNENV_E_INVALID, "null argument");
  if (cfg->abi_version != DYNENV_ABI_VERSION) return fail(DYNENV_E_INVALID, "ABI version mismatch");
  if (cfg->n_envs <= 0) return fail(DYNENV_E_INVALID, "n_envs must be positive");
  if (cfg->n_players <= 0) return fail(DYNENV_E_INVALID, "n_players must be positive");
  if (cfg->noise_magnitude < 0 || cfg->noise_magnitude > 5)
    return fail(DYNENV_E_INVALID, "Error: The noise magnitude must be between 0 and 5!"); /* environment_base.py:162-164 */
  if (cfg->kind != DYNENV_KIND_ROBOCUP && cfg->kind != DYNENV_KIND_DRIVING) return fail(DYNENV_E_INVALID, "unknown environment kind");
  if (cfg->kind == DYNENV_KIND_DRIVING && cfg->obs_type == DYNENV_OBS_IMAGE)
    return fail(DYNENV_E_INVALID, "Image observation type is not supported for this environment"); /* DrivingEnvironment.py:33-35 */
  if (cfg->obs_type != DYNENV_OBS_FULL && cfg->obs_type != DYNENV_OBS_PARTIAL && cfg->obs_type != DYNENV_OBS_IMAGE)
    return fail(DYNENV_E_INVALID, "unknown observation type");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail(DYNENV_E_NODEVICE, "no CUDA device: dynenv_b200 has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(DYNENV_E_INVALID, "bad device ordinal");
  DeviceGuard guard(cfg->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
  DynEnvHandle_s* h = new DynEnvHandle_s();
  h->cfg = *cfg;
  SimCfg& s = h->sim;
  s.n_envs = cfg->n_envs;
  const bool driving = cfg->kind == DYNENV_KIND_DRIVING;
  const int maxPlayers = driving ? DRV_MAXCARS : 5;
  s.n_players = cfg->n_players > maxPlayers ? maxPlayers : cfg->n_players; /* environment_base.py:56-57 */
  s.P = driving ? s.n_players : 2 * s.n_players;
  s.obs_type = cfg->obs_type;
  s.noise_type = cfg->noise_type;
  s.continuous = cfg->continuous;
  s.noise_mag = cfg->noise_magnitude;
  s.rand_base = 0.01f * cfg->noise_magnitude; /* environment_base.py:170 */
  s.random_init = !!(cfg->flags & DYNENV_FLAG_RANDOM_INIT);
  s.deterministic_turn = !!(cfg->flags & DYNENV_FLAG_DETERMINISTIC_TURN);
  s.can_fall = !!(cfg->flags & DYNENV_FLAG_CAN_FALL);
  s.use_obs_rewards = !!(cfg->flags & DYNENV_FLAG_USE_OBS_REWARDS);
  s.sparse_obs = !!(cfg->flags & DYNENV_FLAG_SPARSE_OBS);
  s.auto_reset = !!(cfg->flags & DYNENV_FLAG_AUTO_RESET);
  s.seed_lo = (uint32_t)cfg->seed;
  s.seed_hi = (uint32_t)(cfg->seed >> 32);
  s.env_id_offset = (uint32_t)cfg->env_id_offset;
  s.sub_begin = 0;
  s.sub_end = driving ? DRV_SUBSTEPS : RC_SUBSTEPS;
  h->state_words = driving ? (int)DRV_STATE_WORDS : (int)RC_STATE_WORDS;
  h->last_launches = 0;
  h->timing = 0;
  h->n_timed = 0;
  size_t bytes = (size_t)cfg->n_envs * h->state_words * sizeof(uint32_t);
  if ((ce = cudaMalloc(&h->state, bytes)) != cudaSuccess) {
    delete h;
    return cuda_fail(ce, "cudaMalloc(state)");
  }
  if ((ce = cudaMemset(h->state, 0, bytes)) != cudaSuccess) {
    cudaFree(h->state);
    delete h;
    return cuda_fail(ce, "cudaMemset(state)");
  }
  h->slow_list = nullptr;
  h->slow_count = nullptr;
  h->slow_bins = nullptr;
  h->slow_mark = nullptr;
  h->act_f32 = nullptr;
  h->side = nullptr;
  h->sms = 148;
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, cfg->device);
  if (!driving) {
    /* the side stream outranks the caller's: its few long blocks must get SM slots while the caller's stream still has
       tens of thousands of short blocks to dispatch, or nothing would overlap */
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cfg->n_envs >= (1 << 24)) {
      cudaFree(h->state);
      delete h;
      return fail(DYNENV_E_INVALID, "at most 2^24 - 1 RoboCup environments per handle");
    }
    if ((ce = cudaMalloc(&h->slow_list, (size_t)cfg->n_envs * sizeof(uint32_t))) != cudaSuccess ||
        (ce = cudaMalloc(&h->slow_count, 8 * DYNENV_SLOW_CNT_WORDS * sizeof(uint32_t))) != cudaSuccess ||
        (ce = cudaMalloc(&h->slow_bins, (size_t)DYNENV_SLOW_BINS * cfg->n_envs * sizeof(uint32_t))) != cudaSuccess ||
        (ce = cudaMalloc(&h->slow_mark, (size_t)cfg->n_envs)) != cudaSuccess ||
        (ce = cudaMemset(h->slow_mark, 0, (size_t)cfg->n_envs)) != cudaSuccess ||
        (ce = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) {
      cudaFree(h->slow_list);
      cudaFree(h->slow_count);
      cudaFree(h->slow_bins);
      cudaFree(h->slow_mark);
      cudaFree(h->state);
      delete h;
      return cuda_fail(ce, "cudaMalloc(slow list)");
    }
    if ((ce = dynenv_slow_configure()) != cudaSuccess) {
      cudaFree(h->slow_list);
      cudaFree(h->slow_count);
      cudaFree(h->slow_bins);
      cudaFree(h->slow_mark);
      cudaFree(h->state);
      cudaStreamDestroy(h->side);
      delete h;
      return cuda_fail(ce, "cudaFuncSetAttribute(rc_phys_slow_kernel)");
    }
    for (int k = 0; k < 8; k++) {
      cudaEventCreateWithFlags(&h->ev_fork[k], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming);
    }
  }
  *out = h;
  return DYNENV_OK;
}

int dynenv_destroy(DynEnvHandle h) {
  if (!h) return DYNENV_OK;
  DeviceGuard guard(h->cfg.device);
  cudaFree(h->state);
  cudaFree(h->slow_list);
  cudaFree(h->slow_count);
  cudaFree(h->slow_bins);
  cudaFree(h->slow_mark);
  cudaFree(h->act_f32);
  if (h->side) {
    cudaStreamDestroy(h->side);
    for (int k = 0; k < 8; k++) { cudaEventDestroy(h->ev_fork[k]); cudaEventDestroy(h->ev_join[k]); }
  }
  if (h->timing)
    for (int k = 0; k <= MAX_TIMED; k++) cudaEventDestroy(h->ev[k]);
  delete h;
  return DYNENV_OK;
}

int dynenv_state_words(DynEnvHandle h) { return h ? h->state_words : DYNENV_E_INVALID; }

int dynenv_reset(DynEnvHandle h, const uint8_t* env_mask_dev, void* out, void* stream) {
  if (!h || !out) return fail(DYNENV_E_INVALID, "null argument");
  DeviceGuard guard(h->cfg.device);
  if (h->cfg.kind == DYNENV_KIND_DRIVING) {
    const DynEnvDrivingOut* d = static_cast<const DynEnvDrivingOut*>(out);
    const int dgrid = (h->sim.n_envs + DRV_WPC - 1) / DRV_WPC;
    const size_t dsmem = ((size_t)CFG_WORDS + (size_t)DRV_WPC * drv::ENV_SMEM_WORDS) * sizeof(uint32_t);
    drv_reset_kernel<<<dgrid, DRV_WPC * 32, dsmem, static_cast<cudaStream_t>(stream)>>>(h->sim, h->state, env_mask_dev, d->obs, d->info);
    h->last_launches = 1;
    cudaError_t de = cudaGetLastError();
    if (de != cudaSuccess) return cuda_fail(de, "drv_reset_kernel launch");
    return DYNENV_OK;
  }
  const DynEnvRoboCupOut* o = static_cast<const DynEnvRoboCupOut*>(out);
  int grid = (h->sim.n_envs + RC_WPC - 1) / RC_WPC;
  size_t smem = ((size_t)CFG_WORDS + (size_t)RC_WPC * rc::ENV_SMEM_WORDS) * sizeof(uint32_t);
  if (h->cfg.obs_type == DYNENV_OBS_IMAGE) {
    smem = ((size_t)CFG_WORDS + (size_t)RC_WPC * rc::ENV_SMEM_WORDS_RAW) * sizeof(uint32_t);
    rc_reset_kernel<true><<<grid, RC_WPC * 32, smem, static_cast<cudaStream_t>(stream)>>>(h->sim, h->state, env_mask_dev, o->obs,
                                                                                           o->full, o->info, o->raw);
  } else
    rc_reset_kernel<false><<<grid, RC_WPC * 32, smem, static_cast<cudaStream_t>(stream)>>>(h->sim, h->state, env_mask_dev, o->obs,
                                                                                            o->full, o->info, nullptr);
  h->last_launches = 1;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "rc_reset_kernel launch");
  return DYNENV_OK;
}

int dynenv_step(DynEnvHandle h, const float* actions_dev, void* out, void* terminal_out, void* stream) {
  if (!h || !out || !actions_dev) return fail(DYNENV_E_INVALID, "null argument");
  DeviceGuard guard(h->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->cfg.kind == DYNENV_KIND_DRIVING) {
    const DynEnvDrivingOut* d = static_cast<const DynEnvDrivingOut*>(out);
    const DynEnvDrivingOut* dt = static_cast<const DynEnvDrivingOut*>(terminal_out);
    const int dgrid = (h->sim.n_envs + DRV_WPC - 1) / DRV_WPC;
    const size_t dsmem = ((size_t)CFG_WORDS + (size_t)DRV_WPC * drv::ENV_SMEM_WORDS) * sizeof(uint32_t);
    h->n_timed = 0;
    if (h->timing) cudaEventRecord(h->ev[0], st);
    drv_phys_kernel<<<dgrid, DRV_WPC * 32, dsmem, st>>>(h->sim, h->state, actions_dev, h->sim.sub_begin, h->sim.sub_end, 1);
    if (h->timing) { cudaEventRecord(h->ev[1], st); h->ev_class[0] = 0; }
    drv_obs_kernel<<<dgrid, DRV_WPC * 32, dsmem, st>>>(h->sim, h->state, h->sim.sub_end == DRV_SUBSTEPS, d->obs, d->info,
                                                       dt ? dt->obs : nullptr);
    if (h->timing) { cudaEventRecord(h->ev[2], st); h->ev_class[1] = 1; h->n_timed = 2; }
    h->last_launches = 2;
    cudaError_t de = cudaGetLastError();
    if (de != cudaSuccess) return cuda_fail(de, "drv step kernels launch");
    return DYNENV_OK;
  }
  const DynEnvRoboCupOut* o = static_cast<const DynEnvRoboCupOut*>(out);
  const DynEnvRoboCupOut* t = static_cast<const DynEnvRoboCupOut*>(terminal_out);
  const int grid = (h->sim.n_envs + RC_WPC - 1) / RC_WPC;
  const size_t smem = ((size_t)CFG_WORDS + (size_t)RC_WPC * rc::ENV_SMEM_WORDS) * sizeof(uint32_t);
  const size_t smem_raw = ((size_t)CFG_WORDS + (size_t)RC_WPC * rc::ENV_SMEM_WORDS_RAW) * sizeof(uint32_t);
  const size_t smem_obsp = ((size_t)CFG_WORDS + rc::OBS_SMEM_WORDS) * sizeof(uint32_t);
  const bool image = h->cfg.obs_type == DYNENV_OBS_IMAGE;
  int launches = 0, nmark = 0;
  h->n_timed = 0;
  auto mark = [&](int cls) { /* an event in front of the first launch (cls < 0), then one behind every timed group of launches */
    if (!h->timing || nmark > MAX_TIMED) return;
    if (cls >= 0) { h->ev_class[nmark - 1] = cls; h->n_timed = nmark; }
    cudaEventRecord(h->ev[nmark++], st);
  };
  int blk = 0;
  cudaMemsetAsync(h->slow_count, 0, 8 * DYNENV_SLOW_CNT_WORDS * sizeof(uint32_t), st);
  mark(-1);
  /* blocks of substeps up to the next observation tick (every 10th substep); each is followed by the few environments
     it handed over to the record-based code */
  for (int b = h->sim.sub_begin; b < h->sim.sub_end;) {
    int e = (b / 10 + 1) * 10;
    if (e > h->sim.sub_end) e = h->sim.sub_end;
    const int bi = blk++ & 7;
    uint32_t* cnt = h->slow_count + bi * DYNENV_SLOW_CNT_WORDS;
    rc_phys_kernel<<<grid, RC_WPC * 32, smem, st>>>(h->sim, h->state, actions_dev, b, e, b == h->sim.sub_begin, h->slow_list, cnt,
                                                     h->slow_mark, h->slow_bins);
    launches++;
    mark(0);
    /* fork: the handed-over environments finish their substeps (and are observed) on the side stream ...
       (dynenv_debug_timing(h, 2): on the caller's stream, one after the other, so that every kernel class can be timed) */
    const bool serial = h->timing == 2;
    cudaStream_t side = serial ? st : h->side;
    if (!serial) {
      cudaEventRecord(h->ev_fork[bi], st);
      cudaStreamWaitEvent(h->side, h->ev_fork[bi], 0);
    }
#ifdef RC_SLOW_ONE_WARP
    rc_phys_slow_kernel<<<h->sms * 32, 32, smem, side>>>(h->sim, h->state, actions_dev, e, h->slow_list, cnt);
#else
    dynenv_slow_launch(&h->sim, h->state, actions_dev, b, e, h->slow_bins, cnt, h->sms, side);
#endif
    if (serial) mark(3);
    launches++;
    if (e % 10 == 0) {
      /* ... while the caller's stream observes all the others */
      const int tick = e / 10 - 1, left = h->sim.sub_end - e;
      for (int pass = 0; pass < 2; pass++) {
        cudaStream_t s2 = pass ? side : st;
        const uint8_t* mk = pass ? nullptr : h->slow_mark;
        const uint32_t* ls = pass ? h->slow_list : nullptr;
        const int g2 = pass ? h->sms * 32 : grid;
        if (image)
          rc_obs_kernel<true><<<g2, 32, smem_raw, s2>>>(h->sim, h->state, tick, left, o->obs, o->full, t ? t->obs : nullptr,
                                                         t ? t->full : nullptr, o->raw, t ? t->raw : nullptr, mk, ls, cnt);
        else if (h->cfg.obs_type == DYNENV_OBS_PARTIAL) {
          rc_obsp_kernel<true><<<g2, 32, smem_obsp, s2>>>(h->sim, h->state, tick, left, o->obs, t ? t->obs : nullptr, mk, ls, cnt);
          rc_obsp_kernel<false><<<g2, 32, smem_obsp, s2>>>(h->sim, h->state, tick, left, o->obs, t ? t->obs : nullptr, mk, ls, cnt);
          launches++;
        }
        else
          rc_obs_kernel<false><<<g2, 32, smem, s2>>>(h->sim, h->state, tick, left, o->obs, o->full, t ? t->obs : nullptr,
                                                      t ? t->full : nullptr, nullptr, nullptr, mk, ls, cnt);
        launches++;
      }
      mark(1);
    }
    /* join */
    if (!serial) {
      cudaEventRecord(h->ev_join[bi], h->side);
      cudaStreamWaitEvent(st, h->ev_join[bi], 0);
      mark(3); /* what the side stream's work took beyond the observations it ran beside */
    }
    b = e;
  }
  if (image) rc_finish_kernel<true><<<grid, RC_WPC * 32, smem_raw, st>>>(h->sim, h->state, o->obs, o->full, o->info, o->raw);
  else rc_finish_kernel<false><<<grid, RC_WPC * 32, smem, st>>>(h->sim, h->state, o->obs, o->full, o->info, nullptr);
  launches++;
  mark(2);
  h->last_launches = launches;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "rc step kernels launch");
  return DYNENV_OK;
}

__global__ void __launch_bounds__(256) act_widen_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; /* four actions per thread */
  if (4 * i + 3 < n) {
    const uchar4 v = reinterpret_cast<const uchar4*>(src)[i];
    reinterpret_cast<float4*>(dst)[i] = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
  } else {
    for (long long k = 4 * i; k < n; k++) dst[k] = (float)src[k];
  }
}

int dynenv_step_u8(DynEnvHandle h, const uint8_t* actions_u8_dev, void* out, void* terminal_out, void* stream) {
  if (!h || !out || !actions_u8_dev) return fail(DYNENV_E_INVALID, "null argument");
  if (h->cfg.continuous) return fail(DYNENV_E_INVALID, "byte actions are for MultiDiscrete action spaces (allowHeadTurn / continuousActions off)");
  const long long n = (long long)h->sim.n_envs * h->sim.P * (h->cfg.kind == DYNENV_KIND_DRIVING ? 2 : 4);
  {
    DeviceGuard guard(h->cfg.device);
    if (!h->act_f32) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone; /* (an allocation would invalidate a capture in progress) */
      if (cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
        return fail(DYNENV_E_INVALID, "dynenv_step_u8: call it once outside stream capture first (its first call allocates the widened action buffer)");
      cudaError_t ce = cudaMalloc(&h->act_f32, (size_t)((n + 3) / 4 * 4) * sizeof(float));
      if (ce != cudaSuccess) return cuda_fail(ce, "cudaMalloc(actions)");
    }
    const long long n4 = (n + 3) / 4;
    act_widen_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(actions_u8_dev, h->act_f32, n);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return cuda_fail(ce, "act_widen_kernel launch");
  }
  const int rc = dynenv_step(h, h->act_f32, out, terminal_out, stream);
  if (rc == DYNENV_OK) h->last_launches += 1;
  return rc;
}

int dynenv_set_seed(DynEnvHandle h, uint64_t seed) {
  if (!h) return fail(DYNENV_E_INVALID, "null handle");
  h->cfg.seed = seed;
  h->sim.seed_lo = (uint32_t)seed;
  h->sim.seed_hi = (uint32_t)(seed >> 32);
  return DYNENV_OK;
}

int dynenv_render_images(int device, const float* raw_dev, int64_t n_records, uint32_t* image_dev, void* stream) {
  if (!raw_dev || !image_dev || n_records < 0) return fail(DYNENV_E_INVALID, "bad argument");
  if (n_records == 0) return DYNENV_OK;
  if (n_records > (int64_t)1 << 27) return fail(DYNENV_E_INVALID, "too many records for one call");
  DeviceGuard guard(device);
  cudaError_t ce = guard.err;
  if (ce != cudaSuccess) return cuda_fail(ce, "cudaSetDevice");
  const size_t smem = (size_t)rcimg::SMEM_WORDS * sizeof(uint32_t);
  static thread_local int configured = -1;
  if (configured != device) {
    ce = cudaFuncSetAttribute(rcimg::rc_raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaFuncSetAttribute(rc_raster_kernel)");
    configured = device;
  }
  rcimg::rc_raster_kernel<<<(unsigned)(8 * n_records), rcimg::THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      raw_dev, (long long)n_records, image_dev);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "rc_raster_kernel launch");
  return DYNENV_OK;
}

int dynenv_rollout_insert(DynEnvHandle h, const void* out, int step, int rollout_size, float* rewards_dev,
                          uint8_t* agent_finished_dev, uint8_t* dones_dev, float* episode_acc_dev, void* stream) {
  if (!h || !out || !rewards_dev || !agent_finished_dev || !dones_dev) return fail(DYNENV_E_INVALID, "null argument");
  if (step < 0 || step >= rollout_size) return fail(DYNENV_E_INVALID, "step outside the rollout");
  DeviceGuard guard(h->cfg.device);
  rollout::InfoLayout L;
  const float* info;
  int A;
  if (h->cfg.kind == DYNENV_KIND_DRIVING) {
    info = static_cast<const DynEnvDrivingOut*>(out)->info;
    A = h->cfg.n_players < DRV_MAXCARS ? h->cfg.n_players : DRV_MAXCARS;
    L = {DRV_INFO_WORDS, DRV_I_REWARDS, DRV_I_DONE, DRV_I_EPR, DRV_I_EPPR, -1, DRV_I_EPG, DRV_I_STATE_CARS, 7};
  } else {
    info = static_cast<const DynEnvRoboCupOut*>(out)->info;
    A = 2 * (h->cfg.n_players < 5 ? h->cfg.n_players : 5);
    L = {RC_INFO_WORDS, RC_I_REWARDS, RC_I_DONE, RC_I_EPR, RC_I_EPPR, RC_I_EPOR, RC_I_EPG, RC_I_STATE_ROBOTS, 6};
  }
  if (!info) return fail(DYNENV_E_INVALID, "out->info is null");
  const long long E = h->cfg.n_envs, n = E * A;
  rollout::insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      info, L, E, A, rewards_dev + (size_t)step * n, agent_finished_dev + (size_t)step * n, dones_dev + (size_t)step * E,
      episode_acc_dev);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "rollout insert_kernel launch");
  return DYNENV_OK;
}

int dynenv_discount_rewards(int device, const float* rewards_dev, const float* final_value_dev, const uint8_t* dones_dev,
                            int agents_per_env, float* out_dev, int rollout_size, int64_t n_columns, float discount,
                            void* stream) {
  if (!rewards_dev || !final_value_dev || !out_dev) return fail(DYNENV_E_INVALID, "null argument");
  if (rollout_size < 0 || n_columns < 0 || agents_per_env <= 0 || n_columns % agents_per_env)
    return fail(DYNENV_E_INVALID, "bad rollout shape");
  if (rollout_size == 0 || n_columns == 0) return DYNENV_OK;
  DeviceGuard guard(device);
  cudaError_t ce = guard.err;
  if (ce != cudaSuccess) return cuda_fail(ce, "cudaSetDevice");
  rollout::discount_kernel<<<(unsigned)((n_columns + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rewards_dev, final_value_dev, dones_dev, agents_per_env, out_dev, rollout_size, (long long)n_columns, discount);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "rollout discount_kernel launch");
  return DYNENV_OK;
}

int dynenv_arrange_inputs(int device, const float* rec_dev, int64_t n_envs, int n_time, int n_players, const DynEnvArrangeLayout* layout,
                          float* const* flat_out, int32_t* counts_out, int32_t* obj_counts_out, int32_t* offsets_out,
                          int32_t* block_sums_out, int32_t* totals_out, void* stream) {
  if (!rec_dev || !layout || !flat_out || !counts_out || !obj_counts_out || !offsets_out || !block_sums_out || !totals_out)
    return fail(DYNENV_E_INVALID, "null argument");
  if (layout->n_types <= 0 || layout->n_types > DYNENV_ARRANGE_MAXTYPES || n_envs <= 0 || n_time <= 0 || n_players <= 0)
    return fail(DYNENV_E_INVALID, "bad arranger shape");
  const long long Q = n_envs * n_players, n = Q * n_time;
  if (n * 64 > 0x7fffffffLL) return fail(DYNENV_E_INVALID, "too many (time, player) pairs for 32-bit row offsets");
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = (int)((n + arrange::SCAN_BLOCK - 1) / arrange::SCAN_BLOCK);
  arrange::FlatPtrs fp;
  for (int k = 0; k < DYNENV_ARRANGE_MAXTYPES; k++) fp.p[k] = k < layout->n_types ? flat_out[k] : nullptr;
  cudaMemsetAsync(totals_out, 0, (layout->n_types + 1) * sizeof(int32_t), st);
  arrange::arr_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rec_dev, *layout, Q, n_time, n_players, counts_out, obj_counts_out,
                                                                           totals_out);
  arrange::arr_scan_kernel<<<dim3(nb, layout->n_types), 256, 0, st>>>(counts_out, n, offsets_out, block_sums_out, nb);
  arrange::arr_scan_sums_kernel<<<layout->n_types, 1024, 0, st>>>(block_sums_out, nb, totals_out);
  arrange::arr_compact_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, st>>>(rec_dev, *layout, Q, n_time, n_players, counts_out, offsets_out,
                                                                                block_sums_out, nb, fp);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "arranger kernels launch");
  return DYNENV_OK;
}

int dynenv_arrange_outputs(int device, const float* const* emb, int n_features, int n_types, int64_t n_envs, int n_time, int n_players,
                           const int32_t* counts, const int32_t* offsets, const int32_t* block_sums, const int32_t* obj_counts,
                           int max_count, float* padded_out, uint8_t* mask_out, void* stream) {
  if (!emb || !counts || !offsets || !block_sums || !obj_counts || !padded_out) return fail(DYNENV_E_INVALID, "null argument");
  if (n_types <= 0 || n_types > DYNENV_ARRANGE_MAXTYPES || n_features <= 0 || max_count <= 0) return fail(DYNENV_E_INVALID, "bad arranger shape");
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
  const long long Q = n_envs * n_players, n = Q * n_time;
  const int nb = (int)((n + arrange::SCAN_BLOCK - 1) / arrange::SCAN_BLOCK);
  arrange::FlatPtrs fp;
  for (int k = 0; k < DYNENV_ARRANGE_MAXTYPES; k++) fp.p[k] = k < n_types ? const_cast<float*>(emb[k]) : nullptr;
  const long long threads = n * max_count;
  arrange::arr_pad_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      fp, n_features, n_types, Q, n_time, counts, offsets, block_sums, nb, obj_counts, max_count, padded_out, mask_out);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "arr_pad_kernel launch");
  return DYNENV_OK;
}

#ifdef DYNENV_BOUNDS_CHECK
__global__ void bounds_selftest_kernel(float* out) { /* one index behind the block's 64 bytes (checked, not accessed) */
  bounds_check(threadIdx.x == 0 ? 1000 : 1);
  out[0] = 1.f;
}
int dynenv_debug_bounds_selftest(void) {
  float* d = nullptr;
  cudaMalloc(&d, 4);
  bounds_selftest_kernel<<<1, 32, 64>>>(d);
  cudaDeviceSynchronize();
  cudaFree(d);
  cudaGetLastError();
  return 0;
}
int dynenv_debug_bounds_violations(unsigned long long* out4) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out4, g_bounds_violations, 4 * sizeof(unsigned long long)) == cudaSuccess ? 0 : DYNENV_E_CUDA;
}
#endif
#ifdef RC_DBG_COUNTERS
int dynenv_debug_counters(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, rc::g_dbg, 16 * sizeof(unsigned long long));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(rc::g_dbg, z, sizeof(z)); }
  return 0;
}
#endif
#ifdef DRV_DBG_COUNTERS
int dynenv_debug_drv_counters(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, drv::g_drv_dbg, 8 * sizeof(unsigned long long));
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(drv::g_drv_dbg, z, sizeof(z)); }
  return 0;
}
#endif
#ifdef RASTER_PROFILE
int dynenv_debug_raster_prof(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, rcimg::g_raster_prof, 16 * sizeof(unsigned long long));
  cudaMemcpyFromSymbol(out16 + 16, rcimg::g_raster_prof2, 8 * sizeof(unsigned long long)); /* out16: 24 words */
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(rcimg::g_raster_prof, z, sizeof(z)); cudaMemcpyToSymbol(rcimg::g_raster_prof2, z, 8 * sizeof(unsigned long long)); }
  return 0;
}
#endif

int dynenv_get_state(DynEnvHandle h, uint32_t* state_dev, void* stream) {
  if (!h || !state_dev) return fail(DYNENV_E_INVALID, "null argument");
  DeviceGuard guard(h->cfg.device);
  cudaError_t ce = cudaMemcpyAsync(state_dev, h->state, (size_t)h->sim.n_envs * h->state_words * sizeof(uint32_t),
                                   cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return cuda_fail(ce, "cudaMemcpyAsync(get_state)");
  return DYNENV_OK;
}

int dynenv_set_state(DynEnvHandle h, const uint32_t* state_dev, void* stream) {
  if (!h || !state_dev) return fail(DYNENV_E_INVALID, "null argument");
  DeviceGuard guard(h->cfg.device);
  cudaError_t ce = cudaMemcpyAsync(h->state, state_dev, (size_t)h->sim.n_envs * h->state_words * sizeof(uint32_t),
                                   cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return cuda_fail(ce, "cudaMemcpyAsync(set_state)");
  return DYNENV_OK;
}

int dynenv_debug_fp32_peak(int device, float* tflops_out) {
  if (!tflops_out) return fail(DYNENV_E_INVALID, "null argument");
  DeviceGuard guard(device);
  cudaError_t ce = guard.err;
  if (ce != cudaSuccess) return cuda_fail(ce, "cudaSetDevice");
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  float* d = nullptr;
  ce = cudaMalloc(&d, 4);
  if (ce != cudaSuccess) return cuda_fail(ce, "cudaMalloc");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8, iters = 4096;
  fma_probe_kernel<<<blocks, 256>>>(d, iters, 0.999f, 0.001f); /* warm-up */
  float best = 0.f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    fma_probe_kernel<<<blocks, 256>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 64.0 * iters * 256.0 * blocks;
    const float tf = (float)(flop / (ms * 1e-3) / 1e12);
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "fma_probe_kernel");
  *tflops_out = best;
  return DYNENV_OK;
}

int dynenv_last_launch_count(DynEnvHandle h) { return h ? h->last_launches : 0; }

int dynenv_debug_timing(DynEnvHandle h, int enable) {
  if (!h) return fail(DYNENV_E_INVALID, "null handle");
  DeviceGuard guard(h->cfg.device);
  if (enable && !h->timing)
    for (int k = 0; k <= MAX_TIMED; k++) cudaEventCreate(&h->ev[k]);
  if (!enable && h->timing)
    for (int k = 0; k <= MAX_TIMED; k++) cudaEventDestroy(h->ev[k]);
  h->timing = enable ? (enable == 2 ? 2 : 1) : 0;
  h->n_timed = 0;
  return DYNENV_OK;
}

int dynenv_last_kernel_ms(DynEnvHandle h, float* ms3, int* launches3) {
  if (!h || !ms3 || !launches3) return fail(DYNENV_E_INVALID, "null argument");
  for (int k = 0; k < 4; k++) { ms3[k] = 0.f; launches3[k] = 0; }
  if (!h->timing || h->n_timed == 0) return DYNENV_OK;
  DeviceGuard guard(h->cfg.device);
  cudaError_t ce = cudaEventSynchronize(h->ev[h->n_timed]);
  if (ce != cudaSuccess) return cuda_fail(ce, "cudaEventSynchronize");
  for (int k = 0; k < h->n_timed; k++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]);
    ms3[h->ev_class[k]] += ms;
    launches3[h->ev_class[k]] += 1;
  }
  return DYNENV_OK;
}

int dynenv_debug_substep_range(DynEnvHandle h, int begin, int end) {
  if (!h || begin < 0 || end > (h->cfg.kind == DYNENV_KIND_DRIVING ? DRV_SUBSTEPS : RC_SUBSTEPS) || begin > end)
    return fail(DYNENV_E_INVALID, "bad substep range");
  h->sim.sub_begin = begin;
  h->sim.sub_end = end;
  return DYNENV_OK;
}

}  // extern "C"
