/*
 * RoboCup PARTIAL observation tick, flattened over (observer x object).
 *
 * Replaces, for the ten observers of one environment in one go (reference paths under DynEnv/):
 *   getAgentVision (PARTIAL)   RoboCupEnvironment.py:1192-1316,1536-1561
 *   isSeenInArea / isLineInArea / doesInteract / addNoise / addNoiseLine   cutils.py:352-468,546-565,699-821
 *
 * The first version (agent_vision in rc_sim.cuh, still used for the IMAGE records) ran observer after observer:
 * one 32-lane pass over the 33 point objects, a second pass for the 33rd, eleven lanes of lines, and a Philox
 * evaluation of which a third of the lanes was used -- 15.8 of 32 lanes active on average.  Here the warp walks the
 * flat item space instead:
 *   point objects   item = 33 * observer + object, 32 items per pass: pass p holds the last p objects (all line
 *                   crosses) of observer p - 1 on lanes [0, p) and the first 32 - p objects of observer p on lanes
 *                   [p, 32); P + 1 passes instead of 2 P
 *   lines           two observers per pass (22 lanes)
 *   per observer    false-positive draws, false-positive balls, counts, the coalesced store of the record
 * Rows are compacted with segmented ballots (each lane ranks itself inside its observer's lane range; the only list
 * that straddles a pass boundary is the line-cross list, which carries its count over).  Two staged records are in
 * flight (observer parity); the kernel reads the few state fields it needs straight from the HBM record, so its
 * shared memory holds no copy of the state.
 *
 * Arithmetic, Philox addressing and list orders are those of agent_vision: identical records for identical draws.
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
  O_NL = O_SC + 8,                  /* I [RC_MAXP] lines seen per observer (12 words) */
  O_ROB = O_NL + 12,                /* [2][RC_MAXP-1][2] sightings of the other robots in the observer's head frame */
  O_BP = O_ROB + 36,                /* [2][16] ball x, y, has; posts [4] x, y, has */
  O_RAW = O_BP + 32,                /* [2][RC_CAP_ROBOT][3] rows of the robot list: x, y, seen == Normal */
  O_CNT = O_RAW + 2 * 3 * RC_CAP_ROBOT, /* I [2][8] list lengths */
  O_LBUF = O_CNT + 16,              /* [RC_CAP_LINE * RC_F_LINE] line rows of the odd observer of a pair (56 words) */
  O_ACC = O_LBUF + 56,              /* I processSeens sums: RC_S_STEP_LSUM .. RC_S_STEP_END (112 words) */
  O_ACC_WORDS = RC_S_STEP_END - RC_S_STEP_LSUM,
  O_STAGE = O_ACC + 112,            /* [2][RC_OBS_WORDS] */
  OBS_SMEM_WORDS = O_STAGE + 2 * RC_OBS_WORDS
};
static_assert(O_STAGE % 4 == 0 && O_ACC_WORDS <= 112, "stage alignment");
static_assert(OBS_SMEM_WORDS <= ENV_SMEM_WORDS, "the flattened observation scratch fits the state slice it overlays");

__device__ __forceinline__ Rand6 o_rng(const Env E, uint32_t tick, int site, int a, int b, int blk) {
  Rand6 r;
  uint4 w = philox4x32_10(E.c->seed_lo, E.c->seed_hi, (uint32_t)E.i[O_ENVID], (uint32_t)E.i[O_EPISODE], tick,
                          ((uint32_t)site << 24) | ((uint32_t)(a & 255) << 16) | ((uint32_t)(b & 255) << 8) | (uint32_t)blk);
  r.w[0] = w.x; r.w[1] = w.y; r.w[2] = w.z; r.w[3] = w.w;
  return r;
}

/* robot table, scalars and processSeens sums from the environment's record -- rd(k) = word k of it, in HBM or in the
   shared-memory slice this layout is about to overlay (everything is gathered into registers first); both stages cleared */
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
  uint32_t acc[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int w = lane + 32 * k;
    acc[k] = (load_sums && w < O_ACC_WORDS) ? rd(RC_S_STEP_LSUM + w) : 0u;
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
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int w = lane + 32 * k;
    if (w < O_ACC_WORDS) E.i[O_ACC + w] = (int)acc[k];
  }
  for (int w = lane; w < 2 * RC_OBS_WORDS / 4; w += 32) reinterpret_cast<float4*>(E.f + O_STAGE)[w] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
}

/* insert a row into one list of a staged record (Python list.insert semantics, capped) */
__device__ __noinline__ void o_stage_insert(const Env E, int par, int type, int at, const float* row) {
  const int off[6] = {RC_O_BALL, RC_O_ROBOT, RC_O_POST, RC_O_CROSS, RC_O_FCROSS, RC_O_LINE};
  const int cap[6] = {RC_CAP_BALL, RC_CAP_ROBOT, RC_CAP_POST, RC_CAP_CROSS, RC_CAP_FCROSS, RC_CAP_LINE};
  const int feat[6] = {RC_F_BALL, RC_F_ROBOT, RC_F_POST, RC_F_CROSS, RC_F_FCROSS, RC_F_LINE};
  int n = E.i[O_CNT + 8 * par + type];
  int c = cap[type], f = feat[type];
  if (at > n) at = n;
  if (at >= c) return;
  float* base = E.f + (O_STAGE + par * RC_OBS_WORDS + off[type]);
  int last = n < c ? n : c - 1;
  for (int k = last; k > at; k--)
    for (int w = 0; w < f; w++) base[k * f + w] = base[(k - 1) * f + w];
  for (int w = 0; w < f; w++) base[at * f + w] = row[w];
  if (n < c) E.i[O_CNT + 8 * par + type] = n + 1;
}

/* random false positives (RoboCupEnvironment.py:1279-1305): `hits` = the loop indices whose first draw fired; rare,
   inserted in loop order by lane 0 */
__device__ __noinline__ void o_false_positives(const Env E, int ag, uint32_t tick, unsigned hits, float team, float closest) {
  const int par = ag & 1;
  const float maxVis1 = (W_ * 0.8f) * (W_ * 0.8f);
  const int cnt = O_CNT + 8 * par;
  while (hits) {
    int i = __ffs(hits) - 1;
    hits &= hits - 1;
    if (E.lane == 0) {
      Rand6 ru = o_rng(E, tick, RC_SITE_OBS_FP, ag, i, 0);
      Rand6 r2 = o_rng(E, tick, RC_SITE_OBS_FP, ag, i, 1); /* k = 6, 7 */
      int c = randint_from(ru.u(1), 0, 5);
      float d = ru.u(2) * sqrtf(maxVis1);
      float a = ru.u(3) * 2 * ROB_FOV - ROB_FOV;
      V2 fp = rot(v2(d, 0), a);
      float sz = 1.f - 0.4f * (ru.u(4) - 0.5f);
      float row[8];
      if (c == 0) {
        row[0] = fp.x * STDNORM2; row[1] = fp.y * STDNORM2; row[2] = (BALL_RADIUS * 2 * sz - BALL_RADIUS * 2) * 2.f;
        row[3] = 0.f; row[4] = closest;
        o_stage_insert(E, par, 0, E.i[cnt + 0], row);
      } else if (c == 1) {
        float s, cs;
        sincos_b((ru.u(5) - 0.5f) * 2 * PI_F, &s, &cs);
        row[0] = fp.x * STDNORM2; row[1] = fp.y * STDNORM2; row[2] = (ROB_TOTAL * sz - ROB_TOTAL) * 2.f;
        row[3] = cs; row[4] = s;
        row[5] = r2.u(0) > 0.5f ? -1.f : 1.f;
        row[6] = r2.u(1) > 0.9f ? 1.f : 0.f;
        int at = E.i[cnt + 1];
        if (at < RC_CAP_ROBOT) {
          float* rw = E.f + (O_RAW + par * 3 * RC_CAP_ROBOT + 3 * at);
          rw[0] = fp.x; rw[1] = fp.y; rw[2] = 1.f;
        }
        o_stage_insert(E, par, 1, at, row);
      } else if (c <= 4) {
        float t1 = (float)randint_from(ru.u(5), -1, 1), t2 = (float)randint_from(r2.u(0), -1, 1);
        float rad0 = c == 2 ? POST_RADIUS : PEN_RADIUS;
        polar_row(row, fp, rad0 * sz, rad0, t1, t2, team);
        if (c == 4) {
          float s, cs;
          sincos_b(r2.u(1) * PI_F * 2, &s, &cs);
          row[6] = cs; row[7] = -s;
          o_stage_insert(E, par, 4, E.i[cnt + 3], row); /* inserted at len(crossDets) (reference behaviour) */
        } else o_stage_insert(E, par, c, E.i[cnt + c], row);
      }
    }
    __syncwarp();
  }
}

/* lines of observers q (lanes 0..10) and q + 1 (lanes 11..21): isLineInArea (cutils.py:751-821), addNoiseLine (:352-378),
   normalizeLine (:333-347).  The even observer's rows go into its (free) stage, the odd one's into O_LBUF: its stage
   still belongs to observer q - 1 */
__device__ __noinline__ void o_lines_pair(const Env E, int q, uint32_t tick) {
  const int lane = E.lane, P = E.c->P;
  const float magn = E.c->noise_mag, randB = E.c->rand_base;
  const bool realistic = E.c->noise_type == DYNENV_NOISE_REALISTIC;
  const float maxVis1 = (W_ * 0.8f) * (W_ * 0.8f);
  const int second = lane >= 11 ? 1 : 0;
  const int ob = q + second, j = lane - 11 * second;
  const bool valid = lane < 22 && ob < P;
  const int obc = ob < P ? ob : P - 1;
  const float* T = E.f + (O_TAB + 8 * obc);
  const V2 pos = v2(T[OT_X], T[OT_Y]);
  const float hc = T[OT_HC], hs = T[OT_HS];
  const float qq = 0.70710678118654752f;
  const V2 vec1 = v2((hc - hs) * qq, (hs + hc) * qq), vec2 = v2((hc + hs) * qq, (hs - hc) * qq);
  const Rand6 ru = o_rng(E, tick, RC_SITE_OBS_LINE, obc, j, 0);
  int seen = S_NO;
  V2 lp1 = v2(0, 0), ldir = v2(0, 0);
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
      V2 lp2 = rot_cs(pt2, hc, -hs);
      /* lp2 - lp1 from the line's own direction: a clipped piece can be a fraction of a pixel long, which the
         difference of two float32 end points (~300 px) cannot resolve */
      ldir = rot_cs(D, hc, -hs) * (t2 - t1);
      if (lp1.x < 0 || lp2.x < 0) seen = S_NO;
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
  }
  const unsigned mLine = __ballot_sync(FULLMASK, seen != S_NO);
  const unsigned seg = second ? (0x7FFu << 11) : 0x7FFu;
  if (seen != S_NO) {
    const int row = __popc(mLine & seg & ((1u << lane) - 1));
    float* dst = second ? E.f + (O_LBUF + row * RC_F_LINE) : E.f + (O_STAGE + RC_O_LINE + row * RC_F_LINE);
    float ld = fast_len(ldir); /* ldir = lp2 - lp1; |lp2 x lp1| = |ldir x lp1| */
    float ild = ld > 0 ? __fdividef(1.f, ld) : 0.f;
    dst[0] = (__fdividef(fabsf(cross(ldir, lp1)), ld + 1e-7f) * STDNORM - 0.5f) * 2.f;
    dst[1] = ld > 0 ? ldir.x * ild : 1.f;
    dst[2] = ld > 0 ? ldir.y * ild : 0.f;
    dst[3] = tx; dst[4] = ty;
  }
  if (lane == 0) E.i[O_NL + q] = __popc(mLine & 0x7FFu);
  if (lane == 11 && q + 1 < P) E.i[O_NL + q + 1] = __popc(mLine & (0x7FFu << 11));
  __syncwarp();
}

/* what follows the sightings of observer ag: false positives, false-positive balls, counts, the processSeens sums,
   and the record's way to HBM (the stage is cleared for observer ag + 2) */
__device__ __forceinline__ void o_finalize(const Env E, int ag, uint32_t tick, int nb, int nr, int np, int ncr, int nf,
                                           float* __restrict__ obs_tick) {
  const int lane = E.lane;
  const int par = ag & 1;
  const float randB = E.c->rand_base;
  const bool realistic = E.c->noise_type == DYNENV_NOISE_REALISTIC;
  const unsigned lt = (1u << lane) - 1;
  const int stg = O_STAGE + par * RC_OBS_WORDS, cnt = O_CNT + 8 * par, rawb = O_RAW + par * 3 * RC_CAP_ROBOT;
  const float team = (float)team_of(E, ag);
  const float closest = (ag == E.i[O_CLOSEST] || ag == E.i[O_CLOSEST + 1]) ? 1.f : 0.f;
  const int nl = E.i[O_NL + ag];
  if (lane == 0) {
    E.i[stg + RC_O_NLANDMARKS] = nf + nl + ncr + np;
    E.i[cnt + 0] = nb; E.i[cnt + 1] = nr; E.i[cnt + 2] = np;
    E.i[cnt + 3] = ncr; E.i[cnt + 4] = nf; E.i[cnt + 5] = nl;
  }
  /* one Philox evaluation serves two consumers: lanes 11..20 the test draw of false-positive loop index lane - 11,
     lanes 21..31 the false-positive ball of robot row lane - 21 (include/dynenv_rng.h) */
  const Rand6 rmix = o_rng(E, tick, lane < 21 ? RC_SITE_OBS_FP : RC_SITE_OBS_FPBALL, ag, lane < 21 ? lane - 11 : lane - 21, 0);
  __syncwarp();
  {
    unsigned hits = __ballot_sync(FULLMASK, lane >= 11 && lane < 21 && rmix.u(0) < randB) >> 11;
    if (hits) o_false_positives(E, ag, tick, hits, team, closest);
  }
  /* false-positive balls near robots (:1308-1316): lane 21 + k = row k of the robot list (a 12th row, which needs
     three false-positive robots, takes a second round with its own draw) */
  if (realistic) {
    const int nrNow = E.i[cnt + 1];
    for (int row0 = 0; row0 < nrNow; row0 += 11) {
      const int row = row0 + lane - 21;
      bool create = false;
      float bx = 0, by = 0, brad = 0;
      if (lane >= 21 && row < nrNow && E.f[rawb + 3 * row + 2] != 0.f) {
        Rand6 ru = rmix;
        if (row0) ru = o_rng(E, tick, RC_SITE_OBS_FPBALL, ag, row, 0);
        V2 rp = v2(E.f[rawb + 3 * row], E.f[rawb + 3 * row + 1]);
        if (ru.u(0) < randB * 10 && lensq(rp) < 250.f * 250.f) {
          create = true;
          bx = rp.x + (2 * ru.u(2) - 1.f) * ROB_TOTAL;
          by = rp.y + (2 * ru.u(3) - 1.f) * ROB_TOTAL;
          brad = BALL_RADIUS * 2 * (1.f - 0.4f * (ru.u(4) - 0.5f));
        }
      }
      unsigned cm = __ballot_sync(FULLMASK, create);
      if (cm) {
        int nbNow = E.i[cnt + 0];
        __syncwarp();
        int at = nbNow + __popc(cm & lt);
        if (create && at < RC_CAP_BALL) {
          float* dst = E.f + (stg + RC_O_BALL + at * RC_F_BALL);
          dst[0] = bx * STDNORM2; dst[1] = by * STDNORM2; dst[2] = (brad - BALL_RADIUS * 2) * 2.f; dst[3] = 0.f; dst[4] = closest;
        }
        if (lane == 0) E.i[cnt + 0] = min(nbNow + __popc(cm), (int)RC_CAP_BALL);
        __syncwarp();
      }
    }
  }
  if (lane < 6) E.i[stg + RC_O_COUNTS + lane] = E.i[cnt + lane];
  /* running sums for processSeens */
  if (lane == 0) {
    E.i[O_ACC + (RC_S_STEP_LSUM - RC_S_STEP_LSUM) + ag] += E.i[stg + RC_O_NLANDMARKS];
    E.i[O_ACC + (RC_S_STEP_BSUM - RC_S_STEP_LSUM) + ag] += E.i[stg + RC_O_BALLSEEN];
  }
  if (lane >= 1 && lane < 10) E.i[O_ACC + (RC_S_STEP_RSEEN - RC_S_STEP_LSUM) + 9 * ag + lane - 1] += E.i[stg + RC_O_ROBSEEN + lane - 1];
  __syncwarp();
  float4* src = reinterpret_cast<float4*>(E.f + stg);
  if (obs_tick) {
    float4* dst = reinterpret_cast<float4*>(obs_tick + (size_t)ag * RC_OBS_WORDS);
    for (int w = lane; w < RC_OBS_WORDS / 4; w += 32) { dst[w] = src[w]; src[w] = make_float4(0.f, 0.f, 0.f, 0.f); }
  } else {
    for (int w = lane; w < RC_OBS_WORDS / 4; w += 32) src[w] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
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
  const float ballX = E.f[O_BALLX], ballY = E.f[O_BALLY];
  const int owned = E.i[O_OWNED];
  unsigned robHasPrev = 0;  /* slots of observer p - 1's robots inside its field of view */
  int nfCarry = 0;          /* line crosses observer p - 1 listed in pass p - 1 */
  int c_nb = 0, c_nr = 0, c_np = 0, c_ncr = 0; /* its other list lengths */
#pragma unroll 1
  for (int p = 0; p <= P; p++) {
    if (!(p & 1) && p < P) o_lines_pair(E, p, tick);
    const bool inA = lane < p;
    const bool valid = inA || p < P;
    int ob = inA ? p - 1 : p;
    if (ob >= P) ob = P - 1;
    const int par = ob & 1;
    const int obj = inA ? lane + 33 - p : lane - p;
    const float* T = E.f + (O_TAB + 8 * ob);
    const V2 pos = v2(T[OT_X], T[OT_Y]);
    const float headAngle = T[OT_HA], hc = T[OT_HC], hs = T[OT_HS];
    const int teamI = team_of(E, ob);
    const float team = (float)teamI;
    const float closest = (ob == E.i[O_CLOSEST] || ob == E.i[O_CLOSEST + 1]) ? 1.f : 0.f;
    /* FoV edge directions: (1,0) rotated by headAngle +- pi/4 */
    const float q = 0.70710678118654752f;
    const V2 vec1 = v2((hc - hs) * q, (hs + hc) * q), vec2 = v2((hc + hs) * q, (hs - hc) * q);
    const int stg = O_STAGE + par * RC_OBS_WORDS;

    int kind = -1, idx = 0; /* 0 ball 1 robot 2 post 3 cross 4 line cross */
    float radius = PEN_RADIUS, maxDist = maxVis0, ex0 = 0, ex1 = 0, ex2 = 0;
    V2 pt = v2(0, 0);
    if (valid) {
      if (obj == 0) {
        kind = 0; pt = v2(ballX, ballY); radius = BALL_RADIUS * 2; ex0 = (float)(owned * teamI);
      } else if (obj < 10) {
        idx = obj - 1;
        if (idx < P - 1) {
          kind = 1;
          const int r = idx < ob ? idx : idx + 1;
          const float* R = E.f + (O_TAB + 8 * r);
          pt = v2(R[OT_X], R[OT_Y]); radius = ROB_TOTAL; maxDist = maxVis1;
          ex0 = R[OT_ANG] - headAngle;
          ex1 = (float)(teamI * team_of(E, r));
          ex2 = (float)E.i[O_TAB + 8 * ob + OT_FL]; /* the OBSERVER's fallen|penalized flag (reference behaviour) */
        }
      } else {
        const int mi = obj - 10;
        kind = mi < 4 ? 2 : (mi < 7 ? 3 : 4);
        idx = mi - (mi < 4 ? 0 : (mi < 7 ? 4 : 7));
        const float4 mk = *reinterpret_cast<const float4*>(g_marks[mi]);
        pt = v2(mk.x, mk.y); ex0 = mk.z; ex1 = mk.w;
        if (kind == 2) { radius = POST_RADIUS; maxDist = maxVis1; }
        if (kind == 4) ex2 = -headAngle;
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
    /* publish the sightings the interaction tests read (kinds 0..2 only occur in the head segment) */
    const int robBase = O_ROB + par * 18, bpBase = O_BP + par * 16;
    if (kind == 1) { E.f[robBase + 2 * idx] = pp.x; E.f[robBase + 2 * idx + 1] = pp.y; }
    if (kind == 0) { E.f[bpBase] = pp.x; E.f[bpBase + 1] = pp.y; E.f[bpBase + 2] = has ? 1.f : 0.f; }
    const unsigned robHasNow = (__ballot_sync(FULLMASK, kind == 1 && has) >> (p + 1)) & 0x1FFu;
    __syncwarp();
    /* interactions (RoboCupEnvironment.py:1227-1237): occlusion / proximity by the robots in view */
    int inter = I_NO;
    {
      unsigned m = inA ? robHasPrev : robHasNow;
      if (!has) m = 0;
      const float R2 = (ROB_TOTAL * 2) * (ROB_TOTAL * 2);
      const float lp = lensq(pp);
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        const float ax = E.f[robBase + 2 * i], ay = E.f[robBase + 2 * i + 1];
        if (!(kind == 1 && i == idx)) { /* cutils.doesInteract(robot, object, 35, canOcclude) */
          const float dx = ax - pp.x, dy = ay - pp.y;
          const float cr = ax * pp.y - ay * pp.x, la = ax * ax + ay * ay;
          int t = (dx * dx + dy * dy < R2) ? I_NEARBY : I_NO;
          if (cr * cr < R2 * la && la < lp) t = I_OCCLUDE;
          inter = max(inter, t);
        }
      }
      /* ball near a post (tested on the posts' lanes), penalty cross near the ball */
      bool postNearBall = false;
      if (has && (kind == 2 || kind == 3) && E.f[bpBase + 2] != 0.f) {
        const float dx = pp.x - E.f[bpBase], dy = pp.y - E.f[bpBase + 1];
        const float thr = kind == 2 ? BALL_RADIUS * 8 : BALL_RADIUS * 4;
        const bool nearB = dx * dx + dy * dy < thr * thr;
        if (kind == 3) { if (nearB) inter = max(inter, I_NEARBY); }
        else postNearBall = nearB;
      }
      const unsigned pnb = __ballot_sync(FULLMASK, postNearBall);
      if (kind == 0 && has && pnb) inter = max(inter, I_NEARBY);
    }
    /* noise, false negatives, misclassification (cutils.py:417-468) */
    if (inter == I_OCCLUDE) seen = S_NO;
    const bool draw = seen != S_NO;
    if (__any_sync(FULLMASK, draw)) {
      const Rand6 ru = o_rng(E, tick, RC_SITE_OBS_BALL + (kind < 0 ? 0 : kind), ob, idx, 0);
      if (draw) {
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
    /* list membership (RoboCupEnvironment.py:1254-1276); every list but the line crosses lives in the head segment */
    const bool okStrict = seen != S_NO && seen != S_MISCLASS;
    const unsigned mBall = __ballot_sync(FULLMASK, kind == 0 && okStrict);
    const unsigned mBallMis = __ballot_sync(FULLMASK, kind == 0 && seen == S_MISCLASS);
    const unsigned mRob = __ballot_sync(FULLMASK, kind == 1 && seen != S_NO);
    const unsigned mPost = __ballot_sync(FULLMASK, kind == 2 && seen != S_NO);
    const unsigned mCross = __ballot_sync(FULLMASK, kind == 3 && okStrict);
    const unsigned mCrossMis = __ballot_sync(FULLMASK, kind == 3 && seen == S_MISCLASS);
    const unsigned mFc = __ballot_sync(FULLMASK, kind == 4 && okStrict);
    const unsigned segA = (1u << p) - 1;
    /* misclassified balls and crosses change lists (:1258-1263) */
    int lkind = kind;
    float t1 = ex0, t2 = ex1;
    if (seen == S_MISCLASS) {
      if (kind == 0) { /* a ball taken for a cross gets random tags */
        Rand6 rm = o_rng(E, tick, RC_SITE_OBS_MISCROSS, ob, 0, 0);
        t1 = (float)randint_from(rm.u(0), -1, 1);
        t2 = (float)randint_from(rm.u(1), -1, 1);
        lkind = 3;
      } else { lkind = 0; ex0 = 0.f; }
    }
    if (seen != S_NO) {
      if (lkind <= 1) { /* ball / robot rows: normalised position and size */
        int rrow;
        float* dst;
        if (lkind == 0) {
          rrow = kind == 0 ? 0 : __popc(mBall) + __popc(mCrossMis & lt);
          dst = E.f + (stg + RC_O_BALL + rrow * RC_F_BALL);
          dst[2] = (radius - BALL_RADIUS * 2) * 2.f; dst[3] = ex0; dst[4] = closest;
        } else {
          rrow = __popc(mRob & lt);
          dst = E.f + (stg + RC_O_ROBOT + rrow * RC_F_ROBOT);
          float s, c;
          sincos_b(ex0, &s, &c);
          dst[2] = (radius - ROB_TOTAL) * 2.f; dst[3] = c; dst[4] = s; dst[5] = ex1; dst[6] = ex2;
          float* rw = E.f + (O_RAW + par * 3 * RC_CAP_ROBOT + 3 * rrow);
          rw[0] = pp.x; rw[1] = pp.y; rw[2] = seen == S_NORMAL ? 1.f : 0.f;
        }
        dst[0] = pp.x * STDNORM2; dst[1] = pp.y * STDNORM2;
      } else { /* posts, crosses, line crosses: polar rows (cutils.py:308-315) */
        int rrow, off, feat = 6;
        if (lkind == 2) { rrow = __popc(mPost & lt); off = RC_O_POST; }
        else if (lkind == 3) { rrow = kind == 0 ? __popc(mCross) : __popc(mCross & lt); off = RC_O_CROSS; }
        else { rrow = inA ? nfCarry + __popc(mFc & lt) : __popc(mFc & lt & ~segA); off = RC_O_FCROSS; feat = RC_F_FCROSS; }
        float* dst = E.f + (stg + off + rrow * feat);
        float l2 = lensq(pp);
        float inv = l2 > 0 ? rsqrtf(l2) : 0.f;
        float dist = l2 * inv;
        dst[0] = (dist * STDNORM - 0.5f) * 2.f;
        dst[1] = dist > 0 ? pp.x * team * inv : 1.f;
        dst[2] = dist > 0 ? pp.y * team * inv : 0.f;
        dst[3] = (radius - (lkind == 2 ? POST_RADIUS : PEN_RADIUS)) * 2.f;
        dst[4] = t1 * team; dst[5] = t2 * team;
        if (lkind == 4) {
          float s, c;
          sincos_b(ex2, &s, &c);
          dst[6] = c; dst[7] = -s;
        }
      }
    }
    /* seen flags (:1254-1255): robots by slot, ball strictly seen */
    if (kind == 1) E.i[stg + RC_O_ROBSEEN + idx] = seen != S_NO ? 1 : 0;
    if (kind == 0) E.i[stg + RC_O_BALLSEEN] = okStrict ? 1 : 0;
    __syncwarp();
    if (p > 0) o_finalize(E, p - 1, tick, c_nb, c_nr, c_np, c_ncr, nfCarry + __popc(mFc & segA), obs_tick);
    c_nb = __popc(mBall) + __popc(mCrossMis);
    c_nr = __popc(mRob);
    c_np = __popc(mPost);
    c_ncr = __popc(mCross) + __popc(mBallMis);
    nfCarry = __popc(mFc & ~segA);
    robHasPrev = robHasNow;
    if (!(p & 1) && p + 1 < P) { /* stage 1 is free now: the lines of observer p + 1 move in */
      const int used = E.i[O_NL + p + 1] * RC_F_LINE; /* rows behind the count stay zero */
      for (int w = lane; w < used; w += 32) E.f[O_STAGE + RC_OBS_WORDS + RC_O_LINE + w] = E.f[O_LBUF + w];
      __syncwarp();
    }
  }
}

}  // namespace rc
}  // namespace dynenv
