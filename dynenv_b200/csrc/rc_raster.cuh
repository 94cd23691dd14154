/*
 * IMAGE observation raster kernel: one thread block draws the four bit-planes of one camera of one raw detection
 * record (include/dynenv_b200.h RC_R_*, raster_compose.cuh) in shared memory and streams them to HBM.
 *
 * Replaces the drawing half of getAgentVision (RoboCupEnvironment.py:1318-1467) -- cv2.line / polylines / rectangle /
 * circle on 2 x 4 planes of 480 x 640 float64 per robot and tick, 19.7 MB in the reference -- by 307 KB of bit-planes.
 * The kernel is HBM-write bound by construction: 153 600 bytes per block leave through coalesced 128-bit stores,
 * the drawing is atomicOr on shared-memory words (primitives and poly-line segments spread over the threads).
 */
#pragma once
#include "dynenv_common.cuh"
#include "raster_compose.cuh"

namespace dynenv {
namespace rcimg {

constexpr int THREADS = 256;
/* shared-memory layout in 32-bit words */
constexpr int S_PLANES = 0;                                   /* [4][PLANE_WORDS] */
constexpr int S_RAW = S_PLANES + 4 * raster::PLANE_WORDS;     /* [RC_RAW_WORDS]   */
constexpr int S_PRIMS = S_RAW + RC_RAW_WORDS;                 /* Prim[MAX_PRIMS]  */
constexpr int PRIM_WORDS = sizeof(raster::Prim) / 4;
constexpr int S_CONIC = S_PRIMS + raster::MAX_PRIMS * PRIM_WORDS; /* Conic */
constexpr int CONIC_WORDS = (sizeof(raster::Conic) + 3) / 4;
constexpr int S_TMP = S_CONIC + CONIC_WORDS;                  /* x1[480], x2[480], valid[480] of the uncompacted rows */
constexpr int S_PARAMS = (S_TMP + 3 * raster::CONIC_ROWS + 1) / 2 * 2; /* double[8]: conic parameters, centre */
constexpr int S_MISC = S_PARAMS + 16;                         /* [0] number of primitives, [1] conic present */
constexpr int SMEM_WORDS = S_MISC + 4;

__global__ void __launch_bounds__(THREADS, 1)
rc_raster_kernel(const float* __restrict__ raw, long long n_records, uint32_t* __restrict__ images) {
  const int tid = threadIdx.x;
  const long long rec = blockIdx.x >> 1;
  const int cam = blockIdx.x & 1;
  if (rec >= n_records) return;
  uint32_t* planes = g_smem + S_PLANES;
  float* sraw = reinterpret_cast<float*>(g_smem + S_RAW);
  raster::Prim* prims = reinterpret_cast<raster::Prim*>(g_smem + S_PRIMS);
  raster::Conic* conic = reinterpret_cast<raster::Conic*>(g_smem + S_CONIC);
  int* tmp = reinterpret_cast<int*>(g_smem + S_TMP);
  double* params = reinterpret_cast<double*>(g_smem + S_PARAMS);
  int* misc = reinterpret_cast<int*>(g_smem + S_MISC);
  {
    uint4* p4 = reinterpret_cast<uint4*>(planes);
    for (int w = tid; w < raster::PLANE_WORDS; w += THREADS) p4[w] = make_uint4(0, 0, 0, 0); /* 4 planes = PLANE_WORDS uint4 */
    if (tid < RC_RAW_WORDS) sraw[tid] = raw[rec * RC_RAW_WORDS + tid];
  }
  __syncthreads();
  /* primitives: one slot per thread (float64 projections); the conic fit on another warp */
  if (tid < raster::N_SLOTS) {
    raster::Prim p;
    if (!raster::build_prim_slot(sraw, cam, tid, &p)) p.type = -1;
    prims[tid] = p;
  }
  if (tid == 64) {
    int thickness = 0;
    double cx, cy;
    const bool ok = raster::conic_setup(sraw, cam, params, &cx, &cy, &thickness);
    params[6] = cx; params[7] = cy;
    misc[1] = ok ? 1 : 0;
    conic->thickness = thickness;
    conic->n = 0;
  }
  __syncthreads();
  int nprims = raster::N_SLOTS;
  if (misc[1]) {
    for (int y = tid; y < raster::CONIC_ROWS; y += THREADS) {
      int x1 = 0, x2 = 0;
      const bool v = raster::conic_row(params, params[6], params[7], y, &x1, &x2);
      tmp[y] = x1; tmp[raster::CONIC_ROWS + y] = x2; tmp[2 * raster::CONIC_ROWS + y] = v ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) { /* compaction in row order (rare path, 480 steps) */
      int n = 0;
      for (int y = 0; y < raster::CONIC_ROWS; y++)
        if (tmp[2 * raster::CONIC_ROWS + y]) { conic->x1[n] = tmp[y]; conic->x2[n] = tmp[raster::CONIC_ROWS + y]; conic->y[n] = (short)y; n++; }
      conic->n = n;
      misc[0] = raster::conic_closures(*conic, prims, raster::N_SLOTS);
    }
    __syncthreads();
    nprims = misc[0];
  }
  const int nseg = conic->n > 1 ? 2 * (conic->n - 1) : 0;
  for (int w = tid; w < nprims + nseg; w += THREADS) {
    if (w < nprims) {
      if (prims[w].type >= 0) raster::draw_prim(planes, prims[w]);
    } else raster::draw_conic_segment(planes, *conic, w - nprims);
  }
  __syncthreads();
  {
    uint4* dst = reinterpret_cast<uint4*>(images + (size_t)rec * RC_IMG_WORDS + (size_t)cam * 4 * raster::PLANE_WORDS);
    const uint4* src = reinterpret_cast<const uint4*>(planes);
    for (int w = tid; w < raster::PLANE_WORDS; w += THREADS) dst[w] = src[w];
  }
}

}  // namespace rcimg
}  // namespace dynenv
