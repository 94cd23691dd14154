/*
 * The learner-side re-shuffle of the observations on the device: what InOutArranger (DynEnv/models/models.py:127-274,
 * use :278-307, README.md:228-269) does with nested Python lists, as four streaming kernels over the simulators'
 * padded-with-counts observation records, with no host synchronisation.
 *
 *   rearrange_inputs  (models.py:220-247)  per object type ONE flat array of all sightings ordered by (time step,
 *                     environment x player, row) + counts[type][time][player] + objCounts[time][player] + maxCount
 *       arr_counts_kernel    counts per (type, time, player), objects per (time, player), their maximum
 *       arr_scan_kernel      exclusive prefix sums of the counts in (time, player) order: 1024 per block ...
 *       arr_scan_sums_kernel ... and over the block sums (one block per type); totals[type] = rows of the flat array
 *       arr_compact_kernel   the rows, each to its place (eight lanes per (time, player))
 *   rearrange_outputs (models.py:249-274)  per-type embeddings [n_i][F] -> padded [T][maxCount][players][F] (per player:
 *                     its rows of type 0, then type 1, ...; zeros behind) + padding masks [T][players][maxCount]
 *       arr_pad_kernel
 *
 * Records: rec[E][T][P][rec_words]; player q = e * P + a as in the reference, which chains the environments' agents
 * per time step (models.py:221-223).  HBM-bound streaming work: every record word is read once, every output written once.
 */
#pragma once
#include <cstdint>

#include "../../include/dynenv_b200.h"

namespace dynenv {
namespace arrange {

constexpr int SCAN_BLOCK = 1024;

__global__ void __launch_bounds__(256) arr_counts_kernel(const float* __restrict__ rec, DynEnvArrangeLayout L, long long Q, int T, int P,
                                                         int32_t* __restrict__ counts, int32_t* __restrict__ obj_counts,
                                                         int32_t* __restrict__ totals) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; /* i = t * Q + q */
  const long long n = (long long)T * Q;
  int sum = 0;
  if (i < n) {
    const int t = (int)(i / Q);
    const long long q = i - (long long)t * Q;
    const long long e = q / P;
    const int a = (int)(q - e * P);
    const int32_t* r = reinterpret_cast<const int32_t*>(rec) + ((e * T + t) * P + a) * (long long)L.rec_words;
    for (int k = 0; k < L.n_types; k++) {
      int c = L.count_word[k] < 0 ? 1 : r[L.count_word[k]];
      c = c < 0 ? 0 : (c > L.cap[k] ? L.cap[k] : c);
      counts[(long long)k * n + i] = c;
      sum += c;
    }
    obj_counts[i] = sum;
  }
  /* maxCount: one atomic per warp */
  for (int o = 16; o; o >>= 1) sum = max(sum, __shfl_xor_sync(0xffffffffu, sum, o));
  if ((threadIdx.x & 31) == 0 && sum > 0) atomicMax(&totals[L.n_types], sum);
}

/* exclusive scan of SCAN_BLOCK consecutive counts per block (blockIdx.y = type): local offsets + the block's sum */
__global__ void __launch_bounds__(256) arr_scan_kernel(const int32_t* __restrict__ counts, long long n, int32_t* __restrict__ offsets,
                                                       int32_t* __restrict__ block_sums, int n_blocks) {
  __shared__ int warp_sums[8];
  const int type = blockIdx.y;
  const long long base = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
  const int32_t* c = counts + (long long)type * n;
  int v[4], s = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { v[k] = base + k < n ? c[base + k] : 0; s += v[k]; }
  int incl = s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int wpre = 0;
  for (int w = 0; w < warp; w++) wpre += warp_sums[w];
  int excl = wpre + incl - s;
#pragma unroll
  for (int k = 0; k < 4; k++) { if (base + k < n) offsets[(long long)type * n + base + k] = excl; excl += v[k]; }
  if (threadIdx.x == 255) block_sums[(long long)type * n_blocks + blockIdx.x] = wpre + incl;
}

/* exclusive scan of the block sums (one block per type, in place) and the totals */
__global__ void __launch_bounds__(1024) arr_scan_sums_kernel(int32_t* __restrict__ block_sums, int n_blocks, int32_t* __restrict__ totals) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  int32_t* b = block_sums + (long long)blockIdx.x * n_blocks;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? b[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int wpre = 0;
    for (int w = 0; w < warp; w++) wpre += warp_sums[w];
    const int c0 = carry;
    if (i < n_blocks) b[i] = c0 + wpre + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c0 + wpre + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

struct FlatPtrs {
  float* p[DYNENV_ARRANGE_MAXTYPES];
};

/* eight lanes per (time, player): its rows of every type to their places in the flat arrays.  An observer has ~1.7 rows per
   type (~10 words): a whole warp per (time, player), as in the first version, kept 30 % of its lanes busy behind 18
   dependent loads of counts / offsets; here lane k of a group loads type k's count and row offset, the group shares them
   by shuffle and copies with stride 8. */
static_assert(DYNENV_ARRANGE_MAXTYPES <= 8, "one lane of a group of eight per type");
__global__ void __launch_bounds__(256) arr_compact_kernel(const float* __restrict__ rec, DynEnvArrangeLayout L, long long Q, int T, int P,
                                                          const int32_t* __restrict__ counts, const int32_t* __restrict__ offsets,
                                                          const int32_t* __restrict__ block_sums, int n_blocks, FlatPtrs flat) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; /* i = t * Q + q */
  const int gl = threadIdx.x & 7;
  const long long n = (long long)T * Q;
  const bool valid = i < n;
  const long long ii = valid ? i : n - 1;
  const int t = (int)(ii / Q);
  const long long q = ii - (long long)t * Q;
  const long long e = q / P;
  const int a = (int)(q - e * P);
  const float* r = rec + ((e * T + t) * P + a) * (long long)L.rec_words;
  int c_l = 0, row_l = 0;
  if (valid && gl < L.n_types) {
    c_l = counts[(long long)gl * n + ii];
    row_l = offsets[(long long)gl * n + ii] + block_sums[(long long)gl * n_blocks + (int)(ii / SCAN_BLOCK)];
  }
  for (int k = 0; k < L.n_types; k++) {
    const int c = __shfl_sync(0xffffffffu, c_l, k, 8);
    const int row0 = __shfl_sync(0xffffffffu, row_l, k, 8);
    if (c == 0 || !flat.p[k]) continue;
    const int f = L.feat[k], words = c * f;
    float* dst = flat.p[k] + (long long)row0 * f;
    const float* src = r + L.off[k];
    if (L.stride[k] == f) { /* rows are contiguous in the record */
      for (int w = gl; w < words; w += 8) dst[w] = src[w];
    } else {
      for (int w = gl; w < words; w += 8) {
        const int row = w / f;
        dst[w] = src[row * L.stride[k] + (w - row * f)];
      }
    }
  }
}

/* padded[t][slot][q][F] and mask[t][q][slot]: one thread per (t, slot, q), F floats each */
__global__ void __launch_bounds__(256) arr_pad_kernel(FlatPtrs emb, int F, int n_types, long long Q, int T, const int32_t* __restrict__ counts,
                                                      const int32_t* __restrict__ offsets, const int32_t* __restrict__ block_sums, int n_blocks,
                                                      const int32_t* __restrict__ obj_counts, int max_count, float* __restrict__ padded,
                                                      uint8_t* __restrict__ mask) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; /* ((t * max_count) + slot) * Q + q */
  const long long n = (long long)T * Q;
  if (idx >= (long long)T * max_count * Q) return;
  const long long q = idx % Q;
  const int slot = (int)((idx / Q) % max_count), t = (int)(idx / (Q * max_count));
  const long long i = (long long)t * Q + q;
  float* dst = padded + idx * F;
  const float* src = nullptr;
  int before = 0;
  if (slot < obj_counts[i]) {
    for (int k = 0; k < n_types; k++) {
      if (!emb.p[k]) continue; /* a type without an embedding block takes no slots (models.py:256-262) */
      const int c = counts[(long long)k * n + i];
      if (slot < before + c) {
        const long long row = (long long)offsets[(long long)k * n + i] + block_sums[(long long)k * n_blocks + (int)(i / SCAN_BLOCK)] + (slot - before);
        src = emb.p[k] + row * F;
        break;
      }
      before += c;
    }
  }
  if (mask) mask[((long long)t * Q + q) * max_count + slot] = src ? 0 : 1;
  for (int f = 0; f < F; f++) dst[f] = src ? src[f] : 0.f;
}

}  // namespace arrange
}  // namespace dynenv
