/*
 * IMAGE observation raster kernel: one thread block draws ONE bit-plane (480 x 640 bits, 38 400 bytes) of one raw
 * detection record (include/dynenv_b200.h RC_R_*, raster_compose.cuh) in shared memory and streams it to HBM with
 * 128-bit stores while the other resident blocks of the SM draw.
 *
 * Replaces the drawing half of getAgentVision (RoboCupEnvironment.py:1318-1467) -- cv2.line / polylines / rectangle /
 * circle on 2 x 4 planes of 480 x 640 float64 per robot and tick, 19.7 MB in the reference -- by 307 KB of bit-planes.
 * HBM-write bound by construction: every plane is written exactly once (zero + store alone: 6.0 TB/s, 92 % of the HBM
 * peak); what keeps the kernel at 15 % of that roof is the drawing.  Drawing: the primitives' rows / pixels in closed
 * form on the lanes (raster_prims.cuh), atomicOr on shared-memory words; large primitives over the whole block, small
 * ones and the centre circle's poly-line segments (32 per unit, one per lane) from a work queue that the block's
 * warps drain.  Four blocks of 256 threads per SM (64 registers).
 *
 * Measured on the bench's IMAGE workload (1024 environments = 409 600 planes, 66 % of them non-empty; clock64 per
 * phase with -DRASTER_PROFILE, tools/gpu_raster_prof.py; ablations -DRASTER_SKIP_DRAW / _SKIP_STORE / _SKIP_CONIC /
 * _NONATOMIC): 16.0 ms, of which set-up + store 3.1 ms, primitives 6 ms, the conic's segments 7 ms; a lines-channel
 * plane with the centre circle takes 170 k cycles, without it 38 k, the other channels 3-7 k.  The store is not the
 * problem: a TMA bulk store (cp.async.bulk shared -> global, -DRASTER_BULK_STORE) gives the same time, a block that
 * draws nothing lives 6 k cycles.  The launch shape does not matter either (256 x 3 blocks per SM 41.0 k env-steps/s,
 * 320 x 3 41.0, 384 x 2 40.5, 512 x 2 39.4, 256 x 4 38.3, 192 x 4 37.9, 128 x 5 35.9 in round 1): the warps are bound
 * by the latency of their own integer chains -- the exact OpenCV fixed-point set-up of every primitive -- i.e. by the
 * work per plane.  The drawing routines are compiled once (noinline): inlined into their call sites the kernel is
 * 207 KB of SASS and 50 % slower (instruction fetch).
 */
#pragma once
#include "dynenv_common.cuh"
#include "raster_compose.cuh"

namespace dynenv {
namespace rcimg {

#ifndef RASTER_THREADS
#define RASTER_THREADS 256
#endif
constexpr int THREADS = RASTER_THREADS, WARPS = THREADS / 32;
static_assert(THREADS >= 64 && THREADS % 32 == 0, "one warp builds primitives while another fits the conic");
/* shared-memory layout in 32-bit words: 45.3 KB (up to five blocks per SM).  The plane doubles as the scratch of the conic's
   uncompacted rows (x1[480], x2[480], valid[480]) before it is zeroed. */
constexpr int PLANE_PRIMS = 40;                               /* most primitives of one channel: 11 lines + 24 crosses + 4 closures */
static_assert(RC_CAP_LINE + RC_CAP_CROSS + RC_CAP_FCROSS + 4 <= PLANE_PRIMS && RC_CAP_ROBOT <= PLANE_PRIMS, "primitive list");
constexpr int S_PLANE = 0;                                    /* [PLANE_WORDS]    */
constexpr int S_RAW = S_PLANE + raster::PLANE_WORDS;          /* [RC_RAW_WORDS]   */
constexpr int S_PRIMS = S_RAW + RC_RAW_WORDS;                 /* Prim[PLANE_PRIMS] */
constexpr int PRIM_WORDS = sizeof(raster::Prim) / 4;
constexpr int S_CONIC = S_PRIMS + PLANE_PRIMS * PRIM_WORDS;   /* Conic */
constexpr int CONIC_WORDS = (sizeof(raster::Conic) + 3) / 4;
constexpr int S_PARAMS = (S_CONIC + CONIC_WORDS + 1) / 2 * 2; /* double[8]: conic parameters, centre */
constexpr int S_MISC = S_PARAMS + 16;                         /* [0] number of primitives, [1] conic present, [2] next work unit */
constexpr int SMEM_WORDS = S_MISC + 4;
static_assert(5 * (SMEM_WORDS * 4 + 1024) <= 228 * 1024, "five blocks per SM");
static_assert(3 * raster::CONIC_ROWS <= raster::PLANE_WORDS, "conic scratch inside the plane");

__device__ __forceinline__ void bulk_store_plane(uint32_t* gdst, const uint32_t* ssrc) {
  /* generic-proxy writes (atomicOr) -> async proxy read; one elected thread issues the copy and keeps the block
     alive until the engine has read the shared memory */
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(raster::PLANE_WORDS * 4)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

#ifdef RASTER_PROFILE
/* dev experiment: cycles per phase, summed over the non-empty planes of a launch.  [0] blocks, [1] setup, [2] conic rows,
   [3] drawing (big primitives + work queue), [4] / [5] planes without / with a conic,
   [7] store, [8] total */
__device__ unsigned long long g_raster_prof[16];
__device__ unsigned long long g_raster_prof2[8]; /* lines channel with / without the conic: planes, total, big-primitive loop, queue */
#define RPROF(x) x
#else
#define RPROF(x)
#endif
#ifndef RASTER_MINB
#define RASTER_MINB 4
#endif
__global__ void __launch_bounds__(THREADS, RASTER_MINB)
rc_raster_kernel(const float* __restrict__ raw, long long n_records, uint32_t* __restrict__ images) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long rec = blockIdx.x >> 3;
  const int plane_id = blockIdx.x & 7, cam = plane_id >> 2, ch = plane_id & 3;
  if (rec >= n_records) return;
  uint32_t* plane = g_smem + S_PLANE;
  float* sraw = reinterpret_cast<float*>(g_smem + S_RAW);
  raster::Prim* prims = reinterpret_cast<raster::Prim*>(g_smem + S_PRIMS);
  raster::Conic* conic = reinterpret_cast<raster::Conic*>(g_smem + S_CONIC);
  int* tmp = reinterpret_cast<int*>(plane);
  double* params = reinterpret_cast<double*>(g_smem + S_PARAMS);
  int* misc = reinterpret_cast<int*>(g_smem + S_MISC);
  RPROF(const long long t0 = clock64();)
  { /* a third of the planes have nothing to draw (no ball in sight, no goal post ...): their list counts say so before
       anything is staged or projected -- zeros straight from registers */
    const int* cnt = reinterpret_cast<const int*>(raw + rec * RC_RAW_WORDS) + RC_R_COUNTS;
    const bool nothing = ch == 0 ? cnt[0] <= 0 : ch == 1 ? cnt[1] <= 0 : ch == 2 ? cnt[2] <= 0
                         : (cnt[3] <= 0 && cnt[4] <= 0 && cnt[5] <= 0 && reinterpret_cast<const int*>(raw + rec * RC_RAW_WORDS)[RC_R_CIRCLE_SEEN] == 0);
    if (nothing) { /* block-uniform */
      uint4* dst = reinterpret_cast<uint4*>(images + (size_t)rec * RC_IMG_WORDS + (size_t)plane_id * raster::PLANE_WORDS);
      for (int w = tid; w < raster::PLANE_WORDS / 4; w += THREADS) dst[w] = make_uint4(0, 0, 0, 0);
      return;
    }
  }
  for (int w = tid; w < RC_RAW_WORDS; w += THREADS) sraw[w] = raw[rec * RC_RAW_WORDS + w];
  if (tid == 0) { misc[0] = 0; misc[1] = 0; misc[2] = 0; conic->n = 0; }
  __syncthreads();
  /* primitives of this plane's channel, compacted (the order does not matter: drawing is a bitwise OR); float64
     projections, one slot per thread; the conic fit on the last warp */
  for (int slot = tid; slot < raster::N_SLOTS; slot += THREADS) {
    raster::Prim p;
    if (raster::slot_plane(slot) == ch && raster::build_prim_slot(sraw, cam, slot, &p)) prims[atomicAdd(&misc[0], 1)] = p;
  }
  if (ch == 3 && tid == THREADS - 32) {
    int thickness = 0;
    double cx, cy;
    const bool ok = raster::conic_setup(sraw, cam, params, &cx, &cy, &thickness);
    params[6] = cx; params[7] = cy;
    misc[1] = ok ? 1 : 0;
    conic->thickness = thickness;
  }
  __syncthreads();
  /* a plane without anything to draw (no ball in sight, no goal post ...) is streamed out as zeros from registers */
  if (misc[0] == 0 && !misc[1]) {
    uint4* dst = reinterpret_cast<uint4*>(images + (size_t)rec * RC_IMG_WORDS + (size_t)plane_id * raster::PLANE_WORDS);
    for (int w = tid; w < raster::PLANE_WORDS / 4; w += THREADS) dst[w] = make_uint4(0, 0, 0, 0);
    return;
  }
  RPROF(const long long t1 = clock64();)
  if (misc[1]) { /* block-uniform */
    for (int y = tid; y < raster::CONIC_ROWS; y += THREADS) {
      int x1 = 0, x2 = 0;
      const bool v = raster::conic_row(params, params[6], params[7], y, &x1, &x2);
      tmp[y] = x1; tmp[raster::CONIC_ROWS + y] = x2; tmp[2 * raster::CONIC_ROWS + y] = v ? 1 : 0;
    }
    __syncthreads();
    if (warp == 0) { /* compaction in row order: ballot prefix over 15 groups of 32 rows */
      int n = 0;
      for (int y0 = 0; y0 < raster::CONIC_ROWS; y0 += 32) {
        const int y = y0 + lane;
        const bool v = tmp[2 * raster::CONIC_ROWS + y] != 0;
        const unsigned bal = __ballot_sync(FULLMASK, v);
        if (v) {
          const int k = n + __popc(bal & ((1u << lane) - 1));
          conic->x1[k] = tmp[y]; conic->x2[k] = tmp[raster::CONIC_ROWS + y]; conic->y[k] = (short)y;
        }
        n += __popc(bal);
      }
      __syncwarp();
      if (lane == 0) {
        conic->n = n;
        misc[0] = raster::conic_closures(*conic, prims, misc[0]);
      }
    }
    __syncthreads();
  }
  {
    uint4* p4 = reinterpret_cast<uint4*>(plane);
    for (int w = tid; w < raster::PLANE_WORDS / 4; w += THREADS) p4[w] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  RPROF(const long long t2 = clock64(); long long t3s = t2;)
  const int nprims = misc[0];
  const int nseg = conic->n > 1 ? 2 * (conic->n - 1) : 0;
  const raster::Lanes L = {lane, 32};
  const raster::Lanes one = {0, 1};
#ifndef RASTER_SKIP_DRAW /* (dev ablation: set-up and store only) */
  /* Everything is a queue of work units that the warps draw from until it is empty: 32 of the conic's ~2 x 479
     one-pixel-high poly-line segments (one lane each; first: they are the long units), then the primitives, one warp each
     with their rows / pixels on its lanes (in list order: the large ones first cost 6 % more, the primitives in front of
     the conic units were slower too).  With a fixed primitive -> warp assignment the block waited
     at its barrier for the warp that had drawn the longest ones (27 k cycles against a mean of 17 k,
     profiles/r01_ncu_rc_raster.txt: stalled_barrier 11.5 per issue); drawing the large primitives with the whole block
     (rows over 256 threads) cost more than it gave: every warp repeats the primitive's set-up -- clipping, slopes, edge
     events: most of its instructions -- for an eighth of the rows (15.7 -> 13.3 ms per 1024 environments without it). */
  {
    RPROF(const long long t3 = clock64(); t3s = t3; if (tid == 0) { int nb = 0; for (int w = 0; w < nprims; w++) nb += raster::prim_is_big(prims[w]) ? 1 : 0;
      atomicAdd(&g_raster_prof[6], (unsigned long long)(t3 - t2)); atomicAdd(&g_raster_prof[13], (unsigned long long)nprims);
      atomicAdd(&g_raster_prof[14], (unsigned long long)nb); atomicAdd(&g_raster_prof[15], (unsigned long long)nseg); })
    const int ncu = (nseg + 31) / 32, nunits = ncu + nprims;
    for (;;) {
      int u = 0;
      if (lane == 0) u = atomicAdd(&misc[2], 1);
      u = __shfl_sync(FULLMASK, u, 0);
      if (u >= nunits) break;
      if (u >= ncu) {
        raster::draw_prim(plane, prims[u - ncu], L);
      } else {
        const int sgm = u * 32 + lane;
#ifndef RASTER_SKIP_CONIC
        if (sgm < nseg) raster::draw_conic_segment(plane, *conic, sgm, one);
#endif
      }
    }
  }
#endif
  __syncthreads();
  RPROF(const long long t5 = clock64();)
#ifdef RASTER_SKIP_STORE /* dev ablation: one word per thread keeps the drawing alive */
  images[(size_t)rec * RC_IMG_WORDS + (size_t)plane_id * raster::PLANE_WORDS + tid] = plane[tid * 37];
#elif defined(RASTER_BULK_STORE)
  if (tid == 0) bulk_store_plane(images + (size_t)rec * RC_IMG_WORDS + (size_t)plane_id * raster::PLANE_WORDS, plane);
#else
  {
    uint4* dst = reinterpret_cast<uint4*>(images + (size_t)rec * RC_IMG_WORDS + (size_t)plane_id * raster::PLANE_WORDS);
    const uint4* src = reinterpret_cast<const uint4*>(plane);
#pragma unroll 5
    for (int w = tid; w < raster::PLANE_WORDS / 4; w += THREADS) dst[w] = src[w];
  }
#endif
  RPROF(if (tid == 0) { const long long t6 = clock64(); atomicAdd(&g_raster_prof[0], 1ull); atomicAdd(&g_raster_prof[1], (unsigned long long)(t1 - t0));
    atomicAdd(&g_raster_prof[2], (unsigned long long)(t2 - t1)); atomicAdd(&g_raster_prof[3], (unsigned long long)(t5 - t2)); atomicAdd(&g_raster_prof[4 + (misc[1] ? 1 : 0)], 1ull);
    atomicAdd(&g_raster_prof[7], (unsigned long long)(t6 - t5)); atomicAdd(&g_raster_prof[8], (unsigned long long)(t6 - t0));
    atomicAdd(&g_raster_prof[9 + (plane_id & 3)], (unsigned long long)(t6 - t0));
    if (ch == 3) { atomicAdd(&g_raster_prof2[misc[1] ? 0 : 4], 1ull); atomicAdd(&g_raster_prof2[misc[1] ? 1 : 5], (unsigned long long)(t6 - t0));
      atomicAdd(&g_raster_prof2[misc[1] ? 2 : 6], (unsigned long long)(t3s - t2)); atomicAdd(&g_raster_prof2[misc[1] ? 3 : 7], (unsigned long long)(t5 - t3s)); } })
}

}  // namespace rcimg
}  // namespace dynenv
