/* Host-side interface between dynenv_api.cu and rc_slow.cu (the handed-over RoboCup environments' kernel lives in a
   translation unit of its own: several warps per block, per-warp shared-memory base).  Not part of the C ABI. */
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#ifndef DYNENV_SLOW_WARPS
#define DYNENV_SLOW_WARPS 16 /* environments (warps) per block of rc_phys_slow_kernel */
#endif
#ifndef DYNENV_SLOW_BLOCKS_PER_SM
#define DYNENV_SLOW_BLOCKS_PER_SM (32 / DYNENV_SLOW_WARPS) /* resident blocks of the persistent grid per SM */
#endif
#define DYNENV_SLOW_CFG_BYTES 76 /* sizeof(SimCfg), asserted on both sides */
/* hand-over points of one block of ten substeps: bin = 2 (substep - first substep of the block) + phase */
#define DYNENV_SLOW_BINS 20
#define DYNENV_SLOW_BIN0 4       /* counters of a substep block: [0] flat list length, [DYNENV_SLOW_BIN0 + bin] bin lengths */
#define DYNENV_SLOW_NEXT 1       /* ... [DYNENV_SLOW_NEXT] next work unit of rc_phys_slow_kernel's queue */
#define DYNENV_SLOW_CNT_WORDS 32

size_t dynenv_slow_smem_bytes(void);
cudaError_t dynenv_slow_configure(void);
/* the rest of the substeps [sub_lo, sub_e) of the environments in slow_bins[bin][0 .. slow_count[DYNENV_SLOW_BIN0 + bin]),
   bin stride n_envs */
void dynenv_slow_launch(const void* sim_cfg, uint32_t* state, const float* actions, int sub_lo, int sub_e, const uint32_t* slow_bins,
                        uint32_t* slow_count, int sms, cudaStream_t stream);
