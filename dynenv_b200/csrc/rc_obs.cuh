/*
 * RoboCup PARTIAL observation tick, flattened over (observer x object) and compacted to what is seen.
 *
 * Replaces, for the ten observers of one environment in one go (reference paths under DynEnv/):
 *   getAgentVision (PARTIAL)   RoboCupEnvironment.py:1192-1316,1536-1561
 *   isSeenInArea / isLineInArea / doesInteract / addNoise / addNoiseLine   cutils.py:352-468,546-565,699-821
 *
 * The first version (agent_vision in rc_sim.cuh, still used for the IMAGE records) ran observer after observer and kept
 * 15.8 of 32 lanes busy: a 90-degree field of view sees a third of the 33 point objects, and the five kinds of objects
 * take different branches.  Here:
 *   1. sightings   the warp walks the flat item space, item = 33 * observer + object, 32 items per pass (P + 1 passes):
 *                  isSeenInArea, the cheap proximity tests, and the robots' sightings go to shared memory;
 *   2. queue       every item inside the field of view is pushed onto a shared-memory queue (position in the head
 *                  frame + a packed word); whenever 32 are waiting, one DENSE pass does the occlusion tests, the noise
 *                  draw (one Philox block per item), misclassification, and writes the encoded row;
 *   3. rows        go straight to the HBM record (zero-filled first); the row index inside its list is the list's
 *                  running length + the rank among the lanes of the same (observer, list) (__match_any_sync) -- items
 *                  queue up in the reference's list order, so ranks are list positions;
 *   4. lines       flat over (observer x line), four passes, in a launch of their own (observe_lines / rc_obsl_kernel):
 *                  an SM's instruction cache holds about 32 KB, and a kernel whose hot code is larger issues at half
 *                  the rate (measured: 26 KB hot 81 % issue-active, 33 KB 56 % with a third fewer instructions);
 *   5. the end     false-positive test draws flat over (observer x loop index), false-positive balls flat over the
 *                  robot rows that can have one; counts and the processSeens sums.
 * The kernel reads the few state fields it needs straight from the HBM record: its shared memory holds no copy of the
 * state.  Arithmetic, Philox addressing and list orders are those of agent_vision: identical records for identical draws.
 */
#pragma once
#include "rc_sim.cuh"

namespace dynenv {
namespace rc {

/* shared-memory layout (words from the warp's slice) */
enum {
  O_TAB = 0,                        /* [RC_MAXP][8] robots: x, y, body angle, head angle (body + head), cos, sin, I flag, - */
  OT_X = 0, OT_Y = 1, OT_ANG = 2, OT_HA = 3, OT_HC = 4, OT_HS = 5, OT_FL = 6,
  O_SC = O_TAB + 8 * RC_MAXP,       /* scalars */
  O_BALLX = O_SC, O_BALLY = O_SC + 1, O_OWNED = O_SC + 2, O_CLOSEST = O_SC + 3 /* [2] */, O_ENVID = O_SC + 5,
  O_EPISODE = O_SC + 6,
  O_ROBHAS = O_SC + 8,              /* I [RC_MAXP] slots of the other robots inside the observer's field of view (12 words) */
  O_ROB = O_ROBHAS + 12,            /* [RC_MAXP][RC_MAXP-1][2] their sightings in the observer's head frame */
  O_BP = O_ROB + 18 * RC_MAXP,      /* ball sighting of the pass in flight: x, y, has (4 words) */
  O_RAW = O_BP + 4,                 /* [RC_MAXP][RC_CAP_ROBOT][3] rows of the robot lists: x, y, seen == Normal */
  O_CNT = O_RAW + RC_MAXP * 3 * RC_CAP_ROBOT, /* I [RC_MAXP][8] list lengths */
  O_SEEN = O_CNT + 8 * RC_MAXP,     /* I [RC_MAXP][RC_MAXP] ballsSeen, robotsSeen[9] */
  O_PEND = O_SEEN + RC_MAXP * RC_MAXP, /* [RC_MAXP][6] a ball taken for a cross waits for the crosses: x, y, radius, tag, tag, I flag */
  O_Q = O_PEND + 6 * RC_MAXP,       /* [64][3] queue: x, y, packed */
  O_ACC = O_Q + 192,                /* I processSeens sums: RC_S_STEP_LSUM .. RC_S_STEP_END (112 words) */
  O_ACC_WORDS = RC_S_STEP_END - RC_S_STEP_LSUM,
  OBS_SMEM_WORDS = O_ACC + 112
};
static_assert(O_ACC_WORDS <= 112, "sums");
static_assert((int)OBS_SMEM_WORDS <= (int)ENV_SMEM_WORDS, "the observation scratch fits the state slice it overlays");

__device__ __forceinline__ Rand6 o_rng(const Env E, uint32_t tick, int site, int a, int b, int blk) {
  Rand6 r;
  uint4 w = philox4x32_10(E.c->seed_lo, E.c->seed_hi, (uint32_t)E.i[O_ENVID], (uint32_t)E.i[O_EPISODE], tick,
                          ((uint32_t)site << 24) | ((uint32_t)(a & 255) << 16) | ((uint32_t)(b & 255) << 8) | (uint32_t)blk);
  r.w[0] = w.x; r.w[1] = w.y; r.w[2] = w.z; r.w[3] = w.w;
  return r;
}

/* robot table, scalars and processSeens sums from the environment's record -- rd(k) = word k of it, in HBM or in the
   shared-memory slice this layout is about to overlay (everything is gathered into registers first) */
template <class RD>
__device__ __forceinline__ void o_prologue(const Env E, RD rd, uint32_t env_id, bool load_sums) {
  const int lane = E.lane, P = E.c->P;
  float x = 0, y = 0, a = 0, ha = 0, hs = 0, hc = 1;
  int fl = 0;
  uint32_t sc[5] = {0, 0, 0, 0, 0};
  uint32_t ep = 0;
  if (lane < P) {
    const int L = 2 * lane;
    x = (__uint_as_float(rd(RC_S_PX + L)) + __uint_as_float(rd(RC_S_PX + L + 1))) * 0.5f;   /* Robot.getPos (Robot.py:90-91) */
    y = (__uint_as_float(rd(RC_S_PY + L)) + __uint_as_float(rd(RC_S_PY + L + 1))) * 0.5f;
    a = (__uint_as_float(rd(RC_S_ANG + L)) + __uint_as_float(rd(RC_S_ANG + L + 1))) * 0.5f; /* Robot.getAngle (Robot.py:94-100) */
    ha = a + __uint_as_float(rd(RC_S_HEAD + lane));
    sincos_b(ha, &hs, &hc);
    fl = (rd(RC_S_RFLAGS + lane) & (RC_RF_FALLEN | RC_RF_PENALIZED)) ? 1 : 0;
  } else if (lane == 16) {
    sc[0] = rd(RC_S_PX + BALL); sc[1] = rd(RC_S_PY + BALL); sc[2] = rd(RC_S_BALLOWNED);
    sc[3] = rd(RC_S_CLOSEST); sc[4] = rd(RC_S_CLOSEST + 1);
    ep = rd(RC_S_EPISODE);
  }
  __syncwarp();
  if (lane < P) {
    float* T = E.f + (O_TAB + 8 * lane);
    T[OT_X] = x; T[OT_Y] = y; T[OT_ANG] = a; T[OT_HA] = ha; T[OT_HC] = hc; T[OT_HS] = hs;
    E.i[O_TAB + 8 * lane + OT_FL] = fl;
  } else if (lane == 16) {
    E.i[O_BALLX] = (int)sc[0]; E.i[O_BALLY] = (int)sc[1]; E.i[O_OWNED] = (int)sc[2];
    E.i[O_CLOSEST] = (int)sc[3]; E.i[O_CLOSEST + 1] = (int)sc[4];
    E.i[O_ENVID] = (int)env_id;
    E.i[O_EPISODE] = (int)ep;
  }
  /* (the reset path, whose record lies in the shared memory this layout overlays, passes load_sums = false) */
#pragma unroll 1
  for (int w = lane; w < O_ACC_WORDS; w += 32) E.i[O_ACC + w] = load_sums ? (int)rd(RC_S_STEP_LSUM + w) : 0;
  __syncwarp();
}

/* first word, floats per row and capacity of list t (0 ball .. 5 line) of an observation record: packed constants and a
   shift instead of chains of selects (these sit in the dense passes' inner code) */
static_assert(RC_O_LINE < 512 && RC_F_FCROSS < 16 && RC_CAP_FCROSS < 32, "packed list tables");
__device__ __forceinline__ int list_off(int t) {
  constexpr unsigned long long tab = (unsigned long long)RC_O_BALL | ((unsigned long long)RC_O_ROBOT << 9) | ((unsigned long long)RC_O_POST << 18) |
                                     ((unsigned long long)RC_O_CROSS << 27) | ((unsigned long long)RC_O_FCROSS << 36) | ((unsigned long long)RC_O_LINE << 45);
  return (int)((tab >> (9 * t)) & 511u);
}
__device__ __forceinline__ int list_feat(int t) {
  constexpr unsigned tab = RC_F_BALL | (RC_F_ROBOT << 4) | (RC_F_POST << 8) | (RC_F_CROSS << 12) | (RC_F_FCROSS << 16) | (RC_F_LINE << 20);
  return (int)((tab >> (4 * t)) & 15u);
}
__device__ __forceinline__ int list_cap(int t) {
  constexpr unsigned tab = RC_CAP_BALL | (RC_CAP_ROBOT << 5) | (RC_CAP_POST << 10) | (RC_CAP_CROSS << 15) | (RC_CAP_FCROSS << 20) | (RC_CAP_LINE << 25);
  return (int)((tab >> (5 * t)) & 31u);
}

/* one dense pass over up to 32 queued sightings: occlusion / proximity by the robots in view (RoboCupEnvironment.py:
   1227-1237), noise, false negatives and misclassification (cutils.py:417-468), list membership and the encoded row
   (:1254-1276, :1537-1548) */
__device__ __noinline__ void o_process(const Env E, uint32_t tick, int head, int n, float* __restrict__ obs_tick) {
  const int lane = E.lane;
  const float magn = E.c->noise_mag, randB = E.c->rand_base;
  const bool realistic = E.c->noise_type == DYNENV_NOISE_REALISTIC;
  const float maxVis0 = (W_ * 0.4f) * (W_ * 0.4f), maxVis1 = (W_ * 0.8f) * (W_ * 0.8f);
  const unsigned lt = (1u << lane) - 1;
  const bool live = lane < n;
  const int q = O_Q + 3 * ((head + lane) & 63);
  V2 pp = v2(E.f[q], E.f[q + 1]);
  const uint32_t pk = live ? (uint32_t)E.i[q + 2] : 0u;
  const int ob = pk & 15, kind = (pk >> 4) & 7, idx = (pk >> 7) & 31;
  int seen = live ? (int)((pk >> 12) & 3) : S_NO;
  int inter = (pk >> 14) & 3;
  const float* T = E.f + (O_TAB + 8 * ob);
  const float headAngle = T[OT_HA];
  const int teamI = team_of(E, ob);
  const float team = (float)teamI;
  float radius = PEN_RADIUS, maxDist = maxVis0, ex0 = 0, ex1 = 0, ex2 = 0;
  if (kind == 0) {
    radius = BALL_RADIUS * 2; ex0 = (float)(E.i[O_OWNED] * teamI);
  } else if (kind == 1) {
    const int r = idx < ob ? idx : idx + 1;
    radius = ROB_TOTAL; maxDist = maxVis1;
    ex0 = E.f[O_TAB + 8 * r + OT_ANG] - headAngle;
    ex1 = (float)(teamI * team_of(E, r));
    ex2 = (float)E.i[O_TAB + 8 * ob + OT_FL]; /* the OBSERVER's fallen|penalized flag (reference behaviour) */
  } else {
    const int mi = idx + (kind == 2 ? 0 : (kind == 3 ? 4 : 7));
    const float4 mk = *reinterpret_cast<const float4*>(g_marks[mi]);
    ex0 = mk.z; ex1 = mk.w;
    if (kind == 2) { radius = POST_RADIUS; maxDist = maxVis1; }
    if (kind == 4) ex2 = -headAngle;
  }
  /* cutils.doesInteract(robot, object, 35, canOcclude) for the observer's robots in view */
  {
    unsigned m = live ? (unsigned)E.i[O_ROBHAS + ob] : 0u;
    const int robBase = O_ROB + 18 * ob;
    const float R2 = (ROB_TOTAL * 2) * (ROB_TOTAL * 2);
    const float lp = lensq(pp);
    while (m) {
      const int i = __ffs(m) - 1;
      m &= m - 1;
      const float ax = E.f[robBase + 2 * i], ay = E.f[robBase + 2 * i + 1];
      if (!(kind == 1 && i == idx)) {
        const float dx = ax - pp.x, dy = ay - pp.y;
        const float cr = ax * pp.y - ay * pp.x, la = ax * ax + ay * ay;
        int t = (dx * dx + dy * dy < R2) ? I_NEARBY : I_NO;
        if (cr * cr < R2 * la && la < lp) t = I_OCCLUDE;
        inter = max(inter, t);
      }
    }
  }
  if (inter == I_OCCLUDE) seen = S_NO;
  {
    const Rand6 ru = o_rng(E, tick, RC_SITE_OBS_BALL + kind, ob, idx, 0);
    if (seen != S_NO) {
      V2 noiseVec = v2((ru.u(0) - 0.5f) * magn, (ru.u(1) - 0.5f) * magn);
      if (!realistic) {
        if (ru.u(2) < randB) seen = S_NO;
        pp = pp + noiseVec;
        radius *= (1.f - (ru.u(4) - 0.5f) * 0.2f);
        if (kind == 4) ex2 += (ru.u(5) - 0.5f) * magn * 0.1f;
      } else {
        float range = 0.25f + 3.75f * lensq(pp) * (1.f / maxDist);
        float mult = range;
        if (inter == I_NEARBY) mult = range * 2;
        if (seen == S_DISTANT) mult = range * 3;
        else if (seen == S_PARTIAL) mult = range * 4;
        /* |newPos| - |p| without cancellation: (2 p.d + d.d) / (|newPos| + |p|) */
        V2 dsp = v2(noiseVec.x * mult / 4, noiseVec.y * mult / 4);
        V2 newPos = pp + dsp;
        float lsum = fast_len(newPos) + fast_len(pp);
        float diff = lsum > 0 ? __fdividef(2.f * dot(pp, dsp) + dot(dsp, dsp), lsum) : 0.f;
        if (ru.u(2) < randB * mult) seen = S_NO;
        if ((kind == 0 || kind == 3) && ru.u(3) < randB * mult / 2) seen = S_MISCLASS;
        pp = newPos;
        radius *= 1.f + (ru.u(4) * 0.1f * diff);
        if (kind == 4) ex2 += (ru.u(5) - 0.5f) * magn * mult * (1.f / 180.f);
      }
    }
  }
  /* list membership; misclassified balls and crosses change lists (:1258-1263) */
  int list = kind;
  float t1 = ex0, t2 = ex1;
  bool out = seen != S_NO;
  if (seen == S_MISCLASS) {
    if (kind == 0) { /* a ball taken for a cross gets random tags and joins the cross list behind the crosses: at the end */
      Rand6 rm = o_rng(E, tick, RC_SITE_OBS_MISCROSS, ob, 0, 0);
      float* pd = E.f + (O_PEND + 6 * ob);
      pd[0] = pp.x; pd[1] = pp.y; pd[2] = radius;
      pd[3] = (float)randint_from(rm.u(0), -1, 1); pd[4] = (float)randint_from(rm.u(1), -1, 1);
      E.i[O_PEND + 6 * ob + 5] = 1;
      out = false;
    } else { list = 0; ex0 = 0.f; }
  }
  /* row = the list's length so far + the rank among this pass' lanes of the same (observer, list) */
  const unsigned key = out ? (unsigned)(ob * 8 + list) : (0x100u + lane);
  const unsigned grp = __match_any_sync(FULLMASK, key);
  const int cntAt = O_CNT + 8 * ob + list;
  const int row = out ? E.i[cntAt] + __popc(grp & lt) : 0;
  __syncwarp();
  if (out && (grp & lt) == 0) E.i[cntAt] += __popc(grp);
  if (live) { /* seen flags (:1254-1255): robots by slot, ball strictly seen */
    if (kind == 1) E.i[O_SEEN + RC_MAXP * ob + 1 + idx] = seen != S_NO ? 1 : 0;
    if (kind == 0) E.i[O_SEEN + RC_MAXP * ob] = (seen != S_NO && seen != S_MISCLASS) ? 1 : 0;
  }
  if (out) {
    /* every row starts (a, b, size): normalised position (ball, robot) or polar (cutils.py:308-315) */
    const float l2 = lensq(pp);
    const float inv = l2 > 0 ? rsqrtf(l2) : 0.f;
    const float dist = l2 * inv;
    const bool polar = list >= 2;
    float w0 = polar ? (dist * STDNORM - 0.5f) * 2.f : pp.x * STDNORM2;
    float w1 = polar ? (dist > 0 ? pp.x * team * inv : 1.f) : pp.y * STDNORM2;
    float w2 = polar ? (dist > 0 ? pp.y * team * inv : 0.f) : (radius - (list == 0 ? BALL_RADIUS * 2 : ROB_TOTAL)) * 2.f;
    float w3, w4, w5 = 0.f, w6 = 0.f, w7 = 0.f;
    float s = 0.f, c = 1.f;
    if (list == 1 || list == 4) sincos_b(list == 1 ? ex0 : ex2, &s, &c);
    if (list == 0) { w3 = ex0; w4 = (ob == E.i[O_CLOSEST] || ob == E.i[O_CLOSEST + 1]) ? 1.f : 0.f; }
    else if (list == 1) { w3 = c; w4 = s; w5 = ex1; w6 = ex2; }
    else { w3 = (radius - (list == 2 ? POST_RADIUS : PEN_RADIUS)) * 2.f; w4 = t1 * team; w5 = t2 * team; w6 = c; w7 = -s; }
    const int feat = list_feat(list);
    if (obs_tick) {
      float* dst = obs_tick + (size_t)ob * RC_OBS_WORDS + list_off(list) + row * feat;
      dst[0] = w0; dst[1] = w1; dst[2] = w2; dst[3] = w3; dst[4] = w4;
      if (feat > 5) dst[5] = w5;
      if (feat > 6) dst[6] = w6;
      if (feat > 7) dst[7] = w7;
    }
    if (list == 1) {
      float* rw = E.f + (O_RAW + ob * 3 * RC_CAP_ROBOT + 3 * row);
      rw[0] = pp.x; rw[1] = pp.y; rw[2] = seen == S_NORMAL ? 1.f : 0.f;
    }
  }
  __syncwarp();
}

/* The lines of every observer -- isLineInArea (cutils.py:751-821), addNoiseLine (:352-378), normalizeLine (:333-347) --
   flat over item = 11 * observer + line.  A launch of its own (rc_obsl_kernel) in front of the point objects': together
   the two would not fit the instruction cache.  Also zero-fills the tick's records (rows behind the counts
   stay zero; not with DYNENV_FLAG_SPARSE_OBS) and leaves each observer's line count in the record (RC_O_COUNTS + 5) and in O_CNT (the caller adds it to the
   processSeens sum). */
__device__ __noinline__ void observe_lines(const Env E, uint32_t tick, float* __restrict__ obs_tick) {
  const int lane = E.lane;
  const int P = E.c->P;
  const float magn = E.c->noise_mag, randB = E.c->rand_base;
  const bool realistic = E.c->noise_type == DYNENV_NOISE_REALISTIC;
  const float maxVis1 = (W_ * 0.8f) * (W_ * 0.8f);
  const unsigned lt = (1u << lane) - 1;
  if (obs_tick && !E.c->sparse_obs) { /* (DYNENV_FLAG_SPARSE_OBS: rows behind the counts keep what the buffer held) */
    float4* dst = reinterpret_cast<float4*>(obs_tick);
#pragma unroll 2
    for (int w = lane; w < P * (RC_OBS_WORDS / 4); w += 32) dst[w] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (lane < RC_MAXP) E.i[O_CNT + 8 * lane + 5] = 0;
  __syncwarp();
#pragma unroll 1
  for (int base = 0; base < 11 * P; base += 32) {
    const int item = base + lane;
    const bool valid = item < 11 * P;
    const int ob = valid ? item / 11 : P - 1, j = valid ? item - 11 * ob : 0;
    const float* T = E.f + (O_TAB + 8 * ob);
    const V2 pos = v2(T[OT_X], T[OT_Y]);
    const float hc = T[OT_HC], hs = T[OT_HS];
    const float qq = 0.70710678118654752f;
    const V2 vec1 = v2((hc - hs) * qq, (hs + hc) * qq), vec2 = v2((hc + hs) * qq, (hs - hc) * qq);
    int seen = S_NO;
    V2 lp1 = v2(0, 0), lp2 = v2(0, 0), ldir = v2(0, 0);
    float tx = 0, ty = 0;
    if (valid) {
      const float* ln = g_lines[j];
      V2 p1 = v2(ln[0], ln[1]) - pos, p2 = v2(ln[2], ln[3]) - pos;
      tx = ln[4]; ty = ln[5];
      float dist11 = cross(vec1, p1), dist12 = cross(vec1, p2);
      float dist21 = cross(vec2, p1), dist22 = cross(vec2, p2);
      if (!(dist11 > 0 && dist12 > 0) && !(dist21 < 0 && dist22 < 0)) {
        seen = S_NORMAL;
        V2 D = p2 - p1;
        float t1 = 0.f, t2 = 1.f; /* line parameters of the clipped end points */
        if (!(dist11 <= 0 && dist21 >= 0)) {
          float inter1 = cross(p1, vec1) / (cross(vec1, D) + 1e-7f);
          float inter2 = cross(p1, vec2) / (cross(vec2, D) + 1e-7f);
          t1 = (inter1 < 1 && inter2 < 1) ? fmaxf(inter1, inter2) : fminf(inter1, inter2);
          seen = S_PARTIAL;
        }
        if (!(dist12 <= 0 && dist22 >= 0)) {
          float inter1 = cross(p2, vec1) / -cross(vec1, D);
          float inter2 = cross(p2, vec2) / -cross(vec2, D);
          t2 = 1.f - ((inter1 < 1 && inter2 < 1) ? fmaxf(inter1, inter2) : fminf(inter1, inter2));
          seen = S_PARTIAL;
        }
        V2 pt1 = p1 + D * t1, pt2 = p2 - D * (1.f - t2);
        if (lensq(pt1) > maxVis1 || lensq(pt2) > maxVis1) seen = S_DISTANT;
        lp1 = rot_cs(pt1, hc, -hs);
        lp2 = rot_cs(pt2, hc, -hs);
        /* lp2 - lp1 from the line's own direction: a clipped piece can be a fraction of a pixel long, which the
           difference of two float32 end points (~300 px) cannot resolve */
        ldir = rot_cs(D, hc, -hs) * (t2 - t1);
        if (lp1.x < 0 || lp2.x < 0) seen = S_NO;
      }
    }
    if (__any_sync(FULLMASK, seen != S_NO)) {
      const Rand6 ru = o_rng(E, tick, RC_SITE_OBS_LINE, ob, j, 0);
      if (seen != S_NO) {
        V2 n1 = v2((ru.u(0) - 0.5f) * magn, (ru.u(1) - 0.5f) * magn);
        V2 n2 = v2((ru.u(2) - 0.5f) * magn, (ru.u(3) - 0.5f) * magn);
        if (!realistic) {
          if (ru.u(4) < randB) seen = S_NO;
          lp1 = lp1 + n1;
          ldir = ldir + (n2 - n1);
        } else {
          float m1 = 0.25f + 3.75f * lensq(lp1) * (1.f / maxVis1);
          float m2 = 0.25f + 3.75f * lensq(lp2) * (1.f / maxVis1);
          if (ru.u(4) < randB * ((m1 + m2) * 0.5f)) seen = S_NO;
          V2 a1 = v2(n1.x * m1 / 2, n1.y * m1 / 2), a2 = v2(n2.x * m2 / 2, n2.y * m2 / 2);
          lp1 = lp1 + a1;
          ldir = ldir + (a2 - a1);
        }
      }
    }
    const bool out = seen != S_NO;
    const unsigned grp = __match_any_sync(FULLMASK, out ? (unsigned)ob : (0x100u + lane));
    const int cntAt = O_CNT + 8 * ob + 5;
    const int row = out ? E.i[cntAt] + __popc(grp & lt) : 0;
    __syncwarp();
    if (out && (grp & lt) == 0) E.i[cntAt] += __popc(grp);
    if (out && obs_tick) {
      float* dst = obs_tick + (size_t)ob * RC_OBS_WORDS + RC_O_LINE + row * RC_F_LINE;
      float ld = fast_len(ldir); /* ldir = lp2 - lp1; |lp2 x lp1| = |ldir x lp1| */
      float ild = ld > 0 ? __fdividef(1.f, ld) : 0.f;
      dst[0] = (__fdividef(fabsf(cross(ldir, lp1)), ld + 1e-7f) * STDNORM - 0.5f) * 2.f;
      dst[1] = ld > 0 ? ldir.x * ild : 1.f;
      dst[2] = ld > 0 ? ldir.y * ild : 0.f;
      dst[3] = tx; dst[4] = ty;
    }
    __syncwarp();
  }
  if (lane < P && obs_tick) reinterpret_cast<int*>(obs_tick)[(size_t)lane * RC_OBS_WORDS + RC_O_COUNTS + 5] = E.i[O_CNT + 8 * lane + 5];
  __syncwarp();
}

/* one observation tick of every agent (o_prologue has run); obs_tick = this tick's [P][RC_OBS_WORDS] or null */
__device__ __noinline__ void observe_partial(const Env E, uint32_t tick, float* __restrict__ obs_tick) {
  const int lane = E.lane;
  const int P = E.c->P;
  const float magn = E.c->noise_mag, randB = E.c->rand_base;
  const bool realistic = E.c->noise_type == DYNENV_NOISE_REALISTIC;
  /* _setup_vision(0.4, 0.8): environment_base.py:161-172 */
  const float maxVis0 = (W_ * 0.4f) * (W_ * 0.4f), maxVis1 = (W_ * 0.8f) * (W_ * 0.8f);
  const unsigned lt = (1u << lane) - 1;
#pragma unroll 1
  for (int w = lane; w < 8 * RC_MAXP + RC_MAXP * RC_MAXP + 6 * RC_MAXP; w += 32) E.i[O_CNT + w] = 0; /* O_CNT, O_SEEN, O_PEND */
  __syncwarp();
  /* the lines were listed by observe_lines (an earlier launch): their counts */
  if (lane < P && obs_tick) E.i[O_CNT + 8 * lane + 5] = __ldcg(reinterpret_cast<const int*>(obs_tick) + (size_t)lane * RC_OBS_WORDS + RC_O_COUNTS + 5);
  __syncwarp();
  /* ---------------------------------------------------------------- point objects: sightings, pass p = the last p
     objects (line crosses) of observer p - 1 on lanes [0, p), the first 32 - p of observer p on lanes [p, 32) */
  const float ballX = E.f[O_BALLX], ballY = E.f[O_BALLY];
  int qhead = 0, qn = 0;
#pragma unroll 1
  for (int p = 0; p <= P; p++) {
    const bool inA = lane < p;
    const bool valid = inA || p < P;
    int ob = inA ? p - 1 : p;
    if (ob >= P) ob = P - 1;
    const int obj = inA ? lane + 33 - p : lane - p;
    const float* T = E.f + (O_TAB + 8 * ob);
    const V2 pos = v2(T[OT_X], T[OT_Y]);
    const float hc = T[OT_HC], hs = T[OT_HS];
    /* FoV edge directions: (1,0) rotated by headAngle +- pi/4 */
    const float qq = 0.70710678118654752f;
    const V2 vec1 = v2((hc - hs) * qq, (hs + hc) * qq), vec2 = v2((hc + hs) * qq, (hs - hc) * qq);
    int kind = -1, idx = 0; /* 0 ball 1 robot 2 post 3 cross 4 line cross */
    float radius = PEN_RADIUS, maxDist = maxVis0;
    V2 pt = v2(0, 0);
    if (valid) {
      if (obj == 0) {
        kind = 0; pt = v2(ballX, ballY); radius = BALL_RADIUS * 2;
      } else if (obj < 10) {
        idx = obj - 1;
        if (idx < P - 1) {
          kind = 1;
          const int r = idx < ob ? idx : idx + 1;
          pt = v2(E.f[O_TAB + 8 * r + OT_X], E.f[O_TAB + 8 * r + OT_Y]); radius = ROB_TOTAL; maxDist = maxVis1;
        }
      } else {
        const int mi = obj - 10;
        kind = mi < 4 ? 2 : (mi < 7 ? 3 : 4);
        idx = mi - (mi < 4 ? 0 : (mi < 7 ? 4 : 7));
        const float2 mk = *reinterpret_cast<const float2*>(g_marks[mi]);
        pt = v2(mk.x, mk.y);
        if (kind == 2) { radius = POST_RADIUS; maxDist = maxVis1; }
      }
    }
    /* isSeenInArea, allowPartial (cutils.py:699-747) */
    int seen = S_NO;
    bool has = false;
    V2 pp = v2(0, 0);
    if (kind >= 0) {
      pt = pt - pos;
      const float dist1 = cross(vec1, pt), dist2 = cross(vec2, pt);
      if (dist1 < radius && dist2 > -radius) {
        seen = (dist1 < -radius && dist2 > radius) ? (lensq(pt) < maxDist ? S_NORMAL : S_DISTANT) : S_PARTIAL;
        pp = rot_cs(pt, hc, -hs);
        has = true;
      }
    }
    /* the robots' sightings for the occlusion tests of the dense passes; the ball's for the proximity tests right here */
    if (kind == 1) { E.f[O_ROB + 18 * ob + 2 * idx] = pp.x; E.f[O_ROB + 18 * ob + 2 * idx + 1] = pp.y; }
    if (kind == 0) { E.f[O_BP] = pp.x; E.f[O_BP + 1] = pp.y; E.f[O_BP + 2] = has ? 1.f : 0.f; }
    const unsigned robHasNow = (__ballot_sync(FULLMASK, kind == 1 && has) >> (p + 1)) & 0x1FFu;
    if (lane == 0 && p < P) E.i[O_ROBHAS + p] = (int)robHasNow;
    __syncwarp();
    /* ball near a post (tested on the posts' lanes), penalty cross near the ball: both live in the head segment */
    bool nearby = false, postNearBall = false;
    if (has && (kind == 2 || kind == 3) && E.f[O_BP + 2] != 0.f) {
      const float dx = pp.x - E.f[O_BP], dy = pp.y - E.f[O_BP + 1];
      const float thr = kind == 2 ? BALL_RADIUS * 8 : BALL_RADIUS * 4;
      const bool nearB = dx * dx + dy * dy < thr * thr;
      if (kind == 3) nearby = nearB;
      else postNearBall = nearB;
    }
    const unsigned pnb = __ballot_sync(FULLMASK, postNearBall);
    if (kind == 0 && has && pnb) nearby = true;
    /* queue what is inside the field of view */
    const unsigned pm = __ballot_sync(FULLMASK, has);
    if (has) {
      const int q = O_Q + 3 * ((qhead + qn + __popc(pm & lt)) & 63);
      E.f[q] = pp.x; E.f[q + 1] = pp.y;
      E.i[q + 2] = ob | (kind << 4) | (idx << 7) | (seen << 12) | ((nearby ? I_NEARBY : I_NO) << 14);
    }
    qn += __popc(pm);
    __syncwarp();
    if (qn >= 32) {
      o_process(E, tick, qhead, 32, obs_tick);
      qhead = (qhead + 32) & 63;
      qn -= 32;
    }
  }
  if (qn > 0) o_process(E, tick, qhead, qn, obs_tick);
  /* ---------------------------------------------------------------- the end, lane = observer: the ball that was
     taken for a cross, numLandMarks (before false positives) */
  if (lane < P) {
    const int cnt = O_CNT + 8 * lane;
    if (E.i[O_PEND + 6 * lane + 5]) {
      const float* pd = E.f + (O_PEND + 6 * lane);
      const float team = (float)team_of(E, lane);
      const int at = E.i[cnt + 3];
      if (obs_tick) { /* the polar row of o_process */
        float* dst = obs_tick + (size_t)lane * RC_OBS_WORDS + RC_O_CROSS + at * RC_F_CROSS;
        const float l2 = lensq(v2(pd[0], pd[1]));
        const float inv = l2 > 0 ? rsqrtf(l2) : 0.f;
        const float dist = l2 * inv;
        dst[0] = (dist * STDNORM - 0.5f) * 2.f;
        dst[1] = dist > 0 ? pd[0] * team * inv : 1.f;
        dst[2] = dist > 0 ? pd[1] * team * inv : 0.f;
        dst[3] = (pd[2] - PEN_RADIUS) * 2.f;
        dst[4] = pd[3] * team; dst[5] = pd[4] * team;
      }
      E.i[cnt + 3] = at + 1;
    }
    E.i[cnt + 6] = E.i[cnt + 4] + E.i[cnt + 5] + E.i[cnt + 3] + E.i[cnt + 2];
  }
  __syncwarp();
  /* random false positives (:1279-1305): ten draws per observer, item = 10 * observer + loop index; the hits build
     their rows in parallel and append them in item (= loop) order; a false line cross is INSERTED at len(crossDets)
     (reference behaviour), the warp shifting the rows behind it */
#pragma unroll 1
  for (int base = 0; obs_tick && base < 10 * P; base += 32) {
    const int item = base + lane;
    const bool valid = item < 10 * P;
    const int ob = valid ? item / 10 : 0, i = valid ? item - 10 * ob : 0;
    const Rand6 ru = o_rng(E, tick, RC_SITE_OBS_FP, ob, i, 0);
    const bool hit = valid && ru.u(0) < randB;
    if (!__any_sync(FULLMASK, hit)) continue;
    const Rand6 r2 = o_rng(E, tick, RC_SITE_OBS_FP, ob, i, 1); /* k = 6, 7 */
    const int c = randint_from(ru.u(1), 0, 5);
    const float d = ru.u(2) * sqrtf(maxVis1);
    float sa, ca, s2, c2;
    sincos_b(ru.u(3) * 2 * ROB_FOV - ROB_FOV, &sa, &ca);
    sincos_b(c == 1 ? (ru.u(5) - 0.5f) * 2 * PI_F : r2.u(1) * PI_F * 2, &s2, &c2);
    const V2 fp = v2(d * ca, d * sa);
    const float sz = 1.f - 0.4f * (ru.u(4) - 0.5f);
    const float team = (float)team_of(E, ob);
    const float rad0 = c == 2 ? POST_RADIUS : PEN_RADIUS;
    const float dist = len(fp), inv = dist > 0 ? 1.f / dist : 0.f; /* cutils.py:308-315 */
    const bool polar = c >= 2;
    const float w0 = polar ? (dist * STDNORM - 0.5f) / 0.5f : fp.x * STDNORM2;
    const float w1 = polar ? (dist > 0 ? fp.x * team * inv : 1.f) : fp.y * STDNORM2;
    const float w2 = polar ? (dist > 0 ? fp.y * team * inv : 0.f) : (c == 0 ? (BALL_RADIUS * 2 * sz - BALL_RADIUS * 2) : (ROB_TOTAL * sz - ROB_TOTAL)) * 2.f;
    const float w3 = polar ? (rad0 * sz - rad0) * 2.f : (c == 0 ? 0.f : c2);
    const float w4 = polar ? (float)randint_from(ru.u(5), -1, 1) * team
                           : (c == 0 ? ((ob == E.i[O_CLOSEST] || ob == E.i[O_CLOSEST + 1]) ? 1.f : 0.f) : s2);
    const float w5 = polar ? (float)randint_from(r2.u(0), -1, 1) * team : (r2.u(0) > 0.5f ? -1.f : 1.f);
    const float w6 = polar ? c2 : (r2.u(1) > 0.9f ? 1.f : 0.f);
    const float w7 = -s2;
    float* rec = obs_tick + (size_t)ob * RC_OBS_WORDS;
    const int pre3 = E.i[O_CNT + 8 * ob + 3];                       /* len(crossDets) before this pass' false crosses */
    const unsigned m3 = __ballot_sync(FULLMASK, hit && c == 3);
    { /* classes 0..3: appended to their lists */
      const bool app = hit && c <= 3;
      const unsigned grp = __match_any_sync(FULLMASK, app ? (unsigned)(ob * 8 + c) : (0x100u + lane));
      const int cntAt = O_CNT + 8 * ob + c, cap = list_cap(c);
      const int at = app ? E.i[cntAt] + __popc(grp & lt) : 0;
      __syncwarp();
      if (app && (grp & lt) == 0) E.i[cntAt] = min(E.i[cntAt] + __popc(grp), cap);
      if (app && at < cap) {
        const int feat = list_feat(c);
        float* dst = rec + list_off(c) + at * feat;
        dst[0] = w0; dst[1] = w1; dst[2] = w2; dst[3] = w3; dst[4] = w4;
        if (feat > 5) dst[5] = w5;
        if (feat > 6) dst[6] = w6;
        if (c == 1) {
          float* rw = E.f + (O_RAW + ob * 3 * RC_CAP_ROBOT + 3 * at);
          rw[0] = fp.x; rw[1] = fp.y; rw[2] = 1.f;
        }
      }
      __syncwarp();
    }
    unsigned m4 = __ballot_sync(FULLMASK, hit && c == 4);
    while (m4) { /* class 4: list.insert(len(crossDets), row) into the line-cross list */
      const int src = __ffs(m4) - 1;
      m4 &= m4 - 1;
      const int ob4 = __shfl_sync(FULLMASK, ob, src);
      const int n4 = E.i[O_CNT + 8 * ob4 + 4];
      /* len(crossDets) when the loop reaches this index: the false crosses of lower loop indices count, later ones do not */
      const unsigned same = __ballot_sync(FULLMASK, ob == ob4);
      const int len3 = min(__shfl_sync(FULLMASK, pre3, src) + __popc(m3 & same & ((1u << src) - 1)), (int)RC_CAP_CROSS);
      const int at4 = min(len3, n4);
      if (at4 < RC_CAP_FCROSS) {
        float* fc = obs_tick + (size_t)ob4 * RC_OBS_WORDS + RC_O_FCROSS;
        const int last = n4 < RC_CAP_FCROSS ? n4 : RC_CAP_FCROSS - 1;
        const int words = (last - at4) * RC_F_FCROSS; /* <= 136: the rows at4 .. last - 1 move up by one */
        float mv[5];
#pragma unroll
        for (int k = 0; k < 5; k++) mv[k] = lane + 32 * k < words ? __ldcg(fc + at4 * RC_F_FCROSS + lane + 32 * k) : 0.f;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 5; k++) if (lane + 32 * k < words) fc[(at4 + 1) * RC_F_FCROSS + lane + 32 * k] = mv[k];
        if (lane == src) {
          float* dst = fc + at4 * RC_F_FCROSS;
          dst[0] = w0; dst[1] = w1; dst[2] = w2; dst[3] = w3; dst[4] = w4; dst[5] = w5; dst[6] = w6; dst[7] = w7;
        }
        if (lane == 0 && n4 < RC_CAP_FCROSS) E.i[O_CNT + 8 * ob4 + 4] = n4 + 1;
      }
      __syncwarp();
    }
  }
  /* false-positive balls near robots (:1308-1316): every row of a robot list that is a Normal sighting within 250 can
     have one; those rows are queued (observer, row) and drawn in dense passes */
  if (realistic && obs_tick) {
    qhead = 0; qn = 0;
#pragma unroll 1
    for (int base = 0; base < RC_CAP_ROBOT * P; base += 32) {
      const int item = base + lane;
      const int ob = item / RC_CAP_ROBOT, row = item - RC_CAP_ROBOT * ob;
      bool cand = false;
      if (item < RC_CAP_ROBOT * P && row < E.i[O_CNT + 8 * ob + 1]) {
        const float* rw = E.f + (O_RAW + ob * 3 * RC_CAP_ROBOT + 3 * row);
        cand = rw[2] != 0.f && lensq(v2(rw[0], rw[1])) < 250.f * 250.f;
      }
      const unsigned cm = __ballot_sync(FULLMASK, cand);
      if (cand) E.i[O_Q + ((qn + __popc(cm & lt)) & 127)] = ob | (row << 4);
      qn += __popc(cm);
    }
    __syncwarp();
#pragma unroll 1
    for (int base = 0; base < qn; base += 32) {
      const bool live = base + lane < qn;
      const int pk = live ? E.i[O_Q + base + lane] : 0;
      const int ob = pk & 15, row = pk >> 4;
      const Rand6 ru = o_rng(E, tick, RC_SITE_OBS_FPBALL, ob, row, 0);
      const bool create = live && ru.u(0) < randB * 10;
      const unsigned grp = __match_any_sync(FULLMASK, create ? (unsigned)ob : (0x100u + lane));
      const int cntAt = O_CNT + 8 * ob;
      const int at = create ? E.i[cntAt] + __popc(grp & lt) : 0;
      __syncwarp();
      if (create && (grp & lt) == 0) E.i[cntAt] = min(E.i[cntAt] + __popc(grp), (int)RC_CAP_BALL);
      if (create && at < RC_CAP_BALL && obs_tick) {
        const float* rw = E.f + (O_RAW + ob * 3 * RC_CAP_ROBOT + 3 * row);
        const float bx = rw[0] + (2 * ru.u(2) - 1.f) * ROB_TOTAL;
        const float by = rw[1] + (2 * ru.u(3) - 1.f) * ROB_TOTAL;
        const float brad = BALL_RADIUS * 2 * (1.f - 0.4f * (ru.u(4) - 0.5f));
        float* dst = obs_tick + (size_t)ob * RC_OBS_WORDS + RC_O_BALL + at * RC_F_BALL;
        dst[0] = bx * STDNORM2; dst[1] = by * STDNORM2; dst[2] = (brad - BALL_RADIUS * 2) * 2.f; dst[3] = 0.f;
        dst[4] = (ob == E.i[O_CLOSEST] || ob == E.i[O_CLOSEST + 1]) ? 1.f : 0.f;
      }
      __syncwarp();
    }
  }
  /* counts, numLandMarks, seen flags; running sums for processSeens.  Two lanes per observer: lane ob writes words 0 .. 8 of
     the record's count block (6 counts, numLandMarks, ballsSeen, robotsSeen[0]), lane 16 + ob words 9 .. 16 (robotsSeen[1 .. 8]) */
  {
    const int ob = lane & 15, hi = lane >> 4;
    if (ob < P) {
      int* rc = obs_tick ? reinterpret_cast<int*>(obs_tick) + (size_t)ob * RC_OBS_WORDS + RC_O_COUNTS : nullptr;
      const int cb = O_CNT + 8 * ob, sb = O_SEEN + RC_MAXP * ob;
      if (!hi) {
        int c[7];
#pragma unroll
        for (int k = 0; k < 7; k++) c[k] = E.i[cb + k];
        const int bs = E.i[sb], r0 = E.i[sb + 1];
        if (rc) {
#pragma unroll
          for (int k = 0; k < 7; k++) rc[k] = c[k];
          rc[7] = bs; rc[8] = r0;
        }
        E.i[O_ACC + (RC_S_STEP_LSUM - RC_S_STEP_LSUM) + ob] += c[6] - c[5]; /* observe_lines added the lines */
        E.i[O_ACC + (RC_S_STEP_BSUM - RC_S_STEP_LSUM) + ob] += bs;
        E.i[O_ACC + (RC_S_STEP_RSEEN - RC_S_STEP_LSUM) + 9 * ob] += r0;
      } else {
#pragma unroll
        for (int j = 1; j < 9; j++) {
          const int v = E.i[sb + 1 + j];
          if (rc) rc[8 + j] = v;
          E.i[O_ACC + (RC_S_STEP_RSEEN - RC_S_STEP_LSUM) + 9 * ob + j] += v;
        }
      }
    }
  }
  __syncwarp();
}

}  // namespace rc
}  // namespace dynenv
