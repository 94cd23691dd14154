/*
 * RoboCup: the contact half of Space.step for the environments the register-resident substeps hand over, in a form
 * that fits the instruction cache.
 *
 * The record-based collision stage of rc_sim.cuh (filter_candidates -> collide_kept -> collide_pair -> narrowphase,
 * 22 KB of code, pairs expanded and collided by lane 0) is replaced by collide_lean: the group pairs the broad phase
 * reported are listed in shared memory, every lane takes one of their shape pairs and runs the narrow phase in
 * registers, and the touching pairs update their cached pair (arbiter) one lane at a time in Chipmunk's ascending (i, j)
 * order.  Everything behind it (ageing, pre-step, warm start, solver sweeps, post-solve callbacks) is rc_sim.cuh's.
 * Same arithmetic as narrowphase / collide_pair; scenes with more than 32 candidate slots (scrums) take the old route.
 */
#pragma once
#include "rc_sim.cuh"

namespace dynenv {
namespace rc {

#ifdef RC_DBG_COUNTERS
__device__ unsigned long long g_dbg2[16];
#endif

/* minimum-penetration axis of two CROSSING core segments (deep inter-penetration: rare, out of line) */
__device__ __noinline__ void seg_seg_cross_axis(V2 a0, V2 a1, V2 b0, V2 b1, V2* nOut, float* dOut) {
  V2 da = a1 - a0, db = b1 - b0;
  float la = len(da), lb = len(db);
  float best = -INFINITY;
  V2 bn = v2(da.y / la, -da.x / la);
#pragma unroll 1
  for (int k = 0; k < 2; k++) {
    V2 ax = k == 0 ? v2(da.y / la, -da.x / la) : v2(-da.y / la, da.x / la);
    float d = fminf(dot(ax, b0), dot(ax, b1)) - fmaxf(dot(ax, a0), dot(ax, a1));
    if (d > best) { best = d; bn = ax; }
  }
#pragma unroll 1
  for (int k = 0; k < 2; k++) {
    V2 ax = k == 0 ? v2(db.y / lb, -db.x / lb) : v2(-db.y / lb, db.x / lb);
    float d = fminf(dot(ax, a0), dot(ax, a1)) - fmaxf(dot(ax, b0), dot(ax, b1));
    if (d > best) { best = d; bn = v2(-ax.x, -ax.y); }
  }
  *nOut = bn;
  *dOut = best;
}

/* narrowphase() of rc_sim.cuh with the manifold in registers: shapes a, b in cpCollide order (circle before segment), in
   the frame of body a (dab = position of b relative to a) */
__device__ __forceinline__ void narrow_lean(const Env E, int a, int b, V2 dab, Manifold& m) {
  V2 ata, atb, bta, btb;
  shape_local(E, a, &ata, &atb);
  shape_local(E, b, &bta, &btb);
  bta = bta + dab;
  btb = btb + dab;
  m.cnt = 0;
  m.n = v2(1, 0);
  m.hash[0] = m.hash[1] = 0;
  if (!is_foot(b)) { /* CircleToCircleQuery (a is a circle too) */
    const float mind = 20.f;
    V2 d = bta - ata;
    float dsq = lensq(d);
    if (dsq < mind * mind) {
      float dist = sqrtf(dsq);
      m.n = dist ? d * (1.f / dist) : v2(1, 0);
      m.p1[0] = ata + m.n * 10.f;
      m.p2[0] = bta + m.n * -10.f;
      m.cnt = 1;
    }
  } else if (!is_foot(a)) { /* CircleToSegment */
    V2 closest = closest_on_seg(ata, bta, btb);
    const float mind = 10.f + ROB_RAD;
    V2 d = closest - ata;
    float dsq = lensq(d);
    if (dsq < mind * mind) {
      float dist = sqrtf(dsq);
      m.n = dist ? d * (1.f / dist) : seg_normal(bta, btb);
      m.p1[0] = ata + m.n * 10.f;
      m.p2[0] = closest + m.n * -ROB_RAD;
      m.cnt = 1;
    }
  } else if (!segments_apart(ata, atb, bta, btb, 2 * ROB_RAD + 0.01f)) { /* SegmentToSegment + SupportEdgeForSegment + ContactPoints */
    V2 n = v2(1, 0);
    float d = INFINITY;
    if (segs_intersect(ata, atb, bta, btb)) seg_seg_cross_axis(ata, atb, bta, btb, &n, &d);
    else { /* closest features: the four end points against the other segment, first minimum wins */
#pragma unroll 1
      for (int k = 0; k < 4; k++) {
        const V2 p = k == 0 ? ata : (k == 1 ? atb : (k == 2 ? bta : btb));
        const V2 c = k < 2 ? closest_on_seg(p, bta, btb) : closest_on_seg(p, ata, atb);
        const V2 dv = k < 2 ? c - p : p - c;
        const float dd = len(dv);
        if (dd < d) { d = dd; if (dd > 0) n = dv * (1.f / dd); }
      }
    }
    const float r1 = ROB_RAD, r2 = ROB_RAD;
    if (d <= r1 + r2) {
      V2 tn1 = seg_normal(ata, atb), tn2 = seg_normal(bta, btb);
      V2 e1a, e1b, e2a, e2b;
      int h1a, h1b, h2a, h2b;
      if (dot(tn1, n) > 0.f) { e1a = ata; h1a = 0; e1b = atb; h1b = 1; }
      else { e1a = atb; h1a = 1; e1b = ata; h1b = 0; }
      if (dot(tn2, v2(-n.x, -n.y)) > 0.f) { e2a = bta; h2a = 0; e2b = btb; h2b = 1; }
      else { e2a = btb; h2a = 1; e2b = bta; h2b = 0; }
      m.n = n;
      float d_e1_a = cross(e1a, n), d_e1_b = cross(e1b, n);
      float d_e2_a = cross(e2a, n), d_e2_b = cross(e2b, n);
      float e1_denom = 1.f / (d_e1_b - d_e1_a + 1.17549435e-38f);
      float e2_denom = 1.f / (d_e2_b - d_e2_a + 1.17549435e-38f);
      int cnt = 0;
      {
        V2 p1 = n * r1 + vlerp(e1a, e1b, clamp01((d_e2_b - d_e1_a) * e1_denom));
        V2 p2 = n * -r2 + vlerp(e2a, e2b, clamp01((d_e1_a - d_e2_a) * e2_denom));
        if (dot(p2 - p1, n) <= 0.f) { m.p1[0] = p1; m.p2[0] = p2; m.hash[0] = 1 + (h1a | (h2b << 1)); cnt = 1; }
      }
      {
        V2 p1 = n * r1 + vlerp(e1a, e1b, clamp01((d_e2_a - d_e1_a) * e1_denom));
        V2 p2 = n * -r2 + vlerp(e2a, e2b, clamp01((d_e1_b - d_e2_a) * e2_denom));
        if (dot(p2 - p1, n) <= 0.f) {
          const int h = 1 + (h1b | (h2a << 1));
          if (cnt) { m.p1[1] = p1; m.p2[1] = p2; m.hash[1] = h; }
          else { m.p1[0] = p1; m.p2[0] = p2; m.hash[0] = h; }
          cnt++;
        }
      }
      m.cnt = cnt;
    }
  }
}

/* the pairs the broad phase reported (own: bit = odd foot of a robot whose feet may touch; g0..g3: bit k of round r =
   this lane's group pair of pairs4) collided: lane = shape pair, arbiters updated in ascending (i, j) order */
__device__ __noinline__ void collide_lean(const Env E, unsigned own, unsigned g0, unsigned g1, unsigned g2, unsigned g3,
                                          const uint32_t pairs4) {
  const int lane = E.lane;
  const unsigned lt = (1u << lane) - 1;
  const int nOwn = __popc(own), nG = __popc(g0) + __popc(g1) + __popc(g2) + __popc(g3);
  if (nOwn + 4 * nG > 32) { /* a scrum: the serial route */
    const int kept = filter_candidates(E, own, g0, g1, g2, g3);
    if (kept) collide_kept(E, kept);
    return;
  }
  /* the list: own-feet pairs (0x100 | odd foot), then the group pairs (a << 4 | b) */
  int* gl = &E.i[X_CAND];
  if ((own >> lane) & 1) gl[__popc(own & lt)] = 0x100 | lane;
  {
    int base = nOwn;
    const unsigned gh[4] = {g0, g1, g2, g3};
#pragma unroll
    for (int rnd = 0; rnd < 4; rnd++) {
      if ((gh[rnd] >> lane) & 1) gl[base + __popc(gh[rnd] & lt)] = (int)((pairs4 >> (8 * rnd)) & 255u);
      base += __popc(gh[rnd]);
    }
  }
  __syncwarp();
  /* slot -> shape pair i < j */
  int i = 0, j = 0;
  bool valid = false;
  if (lane < nOwn) {
    j = gl[lane] & 31; i = j - 1;
    valid = true;
  } else if (lane < nOwn + 4 * nG) {
    const int t = lane - nOwn;
    const int e = gl[nOwn + (t >> 2)];
    const int ga = e >> 4, gb = e & 15, x = (t >> 1) & 1, y = t & 1;
    i = (ga < 10 ? 2 * ga : ga + 10) + x;
    j = (gb < 10 ? 2 * gb : gb + 10) + y;
    valid = (x == 0 || ga < 10) && (y == 0 || gb < 10);
    if (valid) { /* bounding circles of the two shapes */
      const float dx = E.f[X_BCX + i] - E.f[X_BCX + j], dy = E.f[X_BCY + i] - E.f[X_BCY + j];
      const float rr = (i < 20 ? ROB_LEN + ROB_RAD : 10.f) + (j < 20 ? ROB_LEN + ROB_RAD : 10.f) + 0.05f;
      valid = dx * dx + dy * dy < rr * rr;
    }
  }
  int a = i, b = j;
  if (is_foot(a) && !is_foot(b)) { a = j; b = i; } /* cpCollide orders by primitive type */
  Manifold m;
  m.cnt = 0;
  V2 dab = v2(0, 0);
  if (valid) {
    dab = pos_diff(E, b, a);
    narrow_lean(E, a, b, dab, m);
  }
  const unsigned key = (unsigned)(i * 32 + j);
  unsigned tm = __ballot_sync(FULLMASK, m.cnt > 0);
  while (tm) {
    const unsigned kmin = __reduce_min_sync(FULLMASK, ((tm >> lane) & 1) ? key : 0xFFFFFFFFu);
    const int src = __ffs(__ballot_sync(FULLMASK, ((tm >> lane) & 1) && key == kmin)) - 1;
    tm &= ~(1u << src);
    if (lane == src) arbiter_update(E, a, b, dab, m);
    __syncwarp();
  }
}

/* broad phase of an environment handed over in mid-substep (positions integrated, shape geometry cached): the group
   circles and the own-feet test of rc_phys2.cuh, from the record */
__device__ __forceinline__ void record_broadphase(const Env E, const uint32_t pairs4, unsigned& own, unsigned& g0, unsigned& g1,
                                                  unsigned& g2, unsigned& g3) {
  const int lane = E.lane;
  const int P = E.c->P;
  const bool isFoot = lane < 2 * P;
  float bcx = 0, bcy = 0, tax = 0, tay = 0, tbx = 0, tby = 0, px = 0, py = 0;
  if (isFoot || lane == BALL) {
    tax = E.f[RC_S_CAX + lane]; tay = E.f[RC_S_CAY + lane]; tbx = E.f[RC_S_CBX + lane]; tby = E.f[RC_S_CBY + lane];
    bcx = (tax + tbx) * 0.5f; bcy = (tay + tby) * 0.5f;
    px = E.f[RC_S_PX + lane]; py = E.f[RC_S_PY + lane];
  } else if (lane >= POST0 && lane < NSHAPE) {
    const V2 pp = post_pos(lane - POST0);
    bcx = pp.x; bcy = pp.y;
  }
  if (lane < NSHAPE) { E.f[X_BCX + lane] = bcx; E.f[X_BCY + lane] = bcy; }
  const float opx = __shfl_xor_sync(FULLMASK, px, 1), opy = __shfl_xor_sync(FULLMASK, py, 1);
  const float oax = __shfl_xor_sync(FULLMASK, tax, 1), oay = __shfl_xor_sync(FULLMASK, tay, 1);
  const float obx = __shfl_xor_sync(FULLMASK, tbx, 1), oby = __shfl_xor_sync(FULLMASK, tby, 1);
  bool ownFeetApart = false;
  if (lane < 20) {
    if (!(lane & 1)) {
      const float ddx = px - opx, ddy = py - opy;
      E.f[X_GX + (lane >> 1)] = (px + opx) * 0.5f;
      E.f[X_GY + (lane >> 1)] = (py + opy) * 0.5f;
      E.f[X_GR + (lane >> 1)] = 21.65f + 0.5f * sqrtf(ddx * ddx + ddy * ddy);
    } else if (isFoot) {
      const float ex = tbx - tax, ey = tby - tay;
      const float c1 = ex * (oay - tay) - ey * (oax - tax), c2 = ex * (oby - tay) - ey * (obx - tax);
      ownFeetApart = !(c1 > 20.f * (2 * ROB_RAD + 0.1f) && c2 > 20.f * (2 * ROB_RAD + 0.1f));
    }
  } else if (lane < NSHAPE) {
    E.f[X_GX + lane - 10] = bcx; E.f[X_GY + lane - 10] = bcy; E.f[X_GR + lane - 10] = 10.f;
  }
  __syncwarp();
  own = __ballot_sync(FULLMASK, ownFeetApart);
  unsigned gh[4];
#pragma unroll
  for (int rnd = 0; rnd < 4; rnd++) {
    const int a = (pairs4 >> (8 * rnd + 4)) & 15, b = (pairs4 >> (8 * rnd)) & 15;
    const float dx = E.f[X_GX + a] - E.f[X_GX + b], dy = E.f[X_GY + a] - E.f[X_GY + b];
    const float rr = E.f[X_GR + a] + E.f[X_GR + b] + 0.05f;
    gh[rnd] = __ballot_sync(FULLMASK, dx * dx + dy * dy < rr * rr);
  }
  g0 = gh[0]; g1 = gh[1]; g2 = gh[2]; g3 = gh[3];
}

/* the contact half of a substep on the record: positions are integrated, the broad phase has reported own / g0..g3 */
__device__ __noinline__ void contact_half(const Env E, unsigned own, unsigned g0, unsigned g1, unsigned g2, unsigned g3,
                                          const uint32_t pairs4, int prevStepped) {
  const int lane = E.lane;
  const int nArb = E.i[RC_S_NARB];
  __syncwarp();
  if (lane == 0) { E.i[RC_S_STEPPED] = 1; E.i[X_NACTIVE] = 0; E.i[X_TOUCHED] = 0; }
  if (lane < nArb) { /* first-touch pairs of the previous substep become normal */
    const int key = E.i[RC_S_ARBKEY + lane];
    if (((key >> 16) & 255) == RC_ARB_FIRST) E.i[RC_S_ARBKEY + lane] = (key & ~0xff0000) | (RC_ARB_NORMAL << 16);
  }
  __syncwarp();
  if (own | g0 | g1 | g2 | g3) collide_lean(E, own, g0, g1, g2, g3, pairs4);
#ifdef RC_DBG_COUNTERS /* what the contact halves look like: active pairs, kinds of pairs, cached pairs */
  __syncwarp();
  if (lane == 0) {
    const int na = E.i[X_NACTIVE];
    atomicAdd(&g_dbg2[na < 3 ? na : 3], 1ull);
    for (int q = 0; q < na; q++) {
      const int key = E.i[RC_S_ARBKEY + E.i[X_ACTIVE + q]];
      const int a = key & 255, b = (key >> 8) & 255;
      int kind;
      if (a == BALL || b == BALL) kind = (a >= POST0 || b >= POST0) ? 8 : 5;       /* ball-post, ball-foot */
      else if (a >= POST0 || b >= POST0) kind = 7;                                   /* post-foot */
      else kind = (a >> 1) == (b >> 1) ? 4 : 6;                                      /* own feet, feet of two robots */
      atomicAdd(&g_dbg2[kind], 1ull);
    }
    atomicAdd(&g_dbg2[9], (unsigned long long)E.i[RC_S_NARB]);
    atomicAdd(&g_dbg2[10], 1ull);
  }
  __syncwarp();
#endif
  space_step_rest(E, 0, prevStepped);
}

}  // namespace rc
}  // namespace dynenv
