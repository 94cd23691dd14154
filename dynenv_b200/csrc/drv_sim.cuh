/*
 * Driving rollout hot path, one warp per environment (see dynenv_common.cuh).
 *
 * Replaces, for a batch of environments (reference paths under DynEnv/):
 *   DrivingEnvironment.step            DrivingEnvironment.py:248-322
 *   processAction / Car.accelerate,turn,crash      :357-373, Car.py:55-117
 *   tick(car) / Road.isPointOnRoad     :376-426, Road.py:74-97
 *   move(pedestrian) / isOffRoad / isOut           :429-524
 *   collision callbacks carCrash / pedHit / carHit / ignore_collision   :587-683
 *   pymunk Space.step (Chipmunk2D 7.0 semantics) called at :278,
 *     with cutils.friction_car / friction_car_crashed / friction_pedestrian_dead   cutils.py:78-140
 *   getAgentVision (PARTIAL)           :750-977 + cutils.py:380-413,479-542,569-696, Road.py:36-71
 *   getFullState / get_full_obs        :686-747, :121-124
 *   scene set-up / reset               :20-115,527-584, environment_base.py:205-224
 *
 * All arithmetic is float32 (positions are two-float numbers, car angles float64: include/dynenv_b200.h).
 * Shared-memory words [0, DRV_STATE_WORDS) are the persistent state record; scratch follows.
 * Shape ids: buildings 0..3, cars 4..13, pedestrians 14..33, obstacles 34..53 (the reference's insertion order).
 */
#pragma once
#include "dynenv_common.cuh"

namespace dynenv {
namespace drv {

constexpr float PI_F = 3.14159265358979323846f;
constexpr float W_ = 1700.f, H_ = 1000.f;
constexpr int MAXTIME = 6000;
constexpr float DT = 0.01f;
constexpr float BUILD_W = 400.f, BUILD_H = 225.f, OBST_HALF = 10.f, PED_R = 5.f; /* :100-105, Obstacle.py, Pedestrian.py:16 */
constexpr float PED_MINV = 1.f / 90.f;
constexpr float ROAD_W = 35.f;
constexpr float MAXVIS0 = (W_ * 0.4f) * (W_ * 0.4f), MAXVIS1 = (W_ * 0.6f) * (W_ * 0.6f); /* _setup_vision(0.4, 0.6) */
constexpr float ANGLE_NOISE = PI_F / 180.f; /* cutils.py:203 */
constexpr float NORMX = 10.f / W_, NORMY = 10.f / H_, NORMW = 1.f / 7.5f, NORMH = 1.f / 15.f;
constexpr float SNORMX = 0.5f / (W_ + 100.f), SNORMY = 0.5f / (H_ + 100.f), SNORMW = 1.f / 15.f, SNORMH = 1.f / 25.f;
constexpr int NCAND = 32; /* bounding-box overlapping pairs per substep: one per lane */
constexpr int NOCC = 34; /* possible occluders of one observer: cars 0..9, obstacles 10..29, buildings 30..33 */
constexpr int NRESET_BLOCKS = 48;

enum { S_NO = 0, S_PARTIAL = 1, S_DISTANT = 2, S_NORMAL = 3, S_MISCLASS = 4 };
enum { I_NO = 0, I_NEARBY = 1, I_OCCLUDE = 2 };

/* ---- scratch layout (words after the state record) */
enum {
  X_REW = DRV_S_STEP_REW, X_POSREW = DRV_S_STEP_POSREW, X_TEAMREW = DRV_S_STEP_TEAMREW, X_DIAG = DRV_S_STEP_DIAG,
  X_TOUCHED = DRV_STATE_WORDS, /* I bit k: cached pair k collided this substep */
  X_NACTIVE = X_TOUCHED + 1,
  X_ACTIVE = X_NACTIVE + 1,    /* I[DRV_MAXACTIVE] */
  X_ENVID = X_ACTIVE + DRV_MAXACTIVE,
  X_UNION = (X_ENVID + 1 + 3) / 4 * 4,
  /* physics view */
  X_ANX = X_UNION,             /* F[16] contact normal of the active pair */
  X_ANY = X_ANX + DRV_MAXACTIVE,
  X_CR1X = X_ANY + DRV_MAXACTIVE, /* per contact [2 * DRV_MAXACTIVE] */
  X_CR1Y = X_CR1X + 2 * DRV_MAXACTIVE,
  X_CR2X = X_CR1Y + 2 * DRV_MAXACTIVE,
  X_CR2Y = X_CR2X + 2 * DRV_MAXACTIVE,
  X_CNMASS = X_CR2Y + 2 * DRV_MAXACTIVE,
  X_CBIAS = X_CNMASS + 2 * DRV_MAXACTIVE,
  X_CBOUNCE = X_CBIAS + 2 * DRV_MAXACTIVE,
  X_CJBIAS = X_CBOUNCE + 2 * DRV_MAXACTIVE,
  X_CJN = X_CJBIAS + 2 * DRV_MAXACTIVE,
  X_CC = X_CJN + 2 * DRV_MAXACTIVE,  /* F[10] cos / sin of the car angles, bounding-box half extents */
  X_CS = X_CC + DRV_MAXCARS,
  X_CEX = X_CS + DRV_MAXCARS,
  X_CEY = X_CEX + DRV_MAXCARS,
  X_BPMASK = X_CEY + DRV_MAXCARS,    /* I[64] per shape: cars whose bounding box overlaps it */
  X_NCAND = X_BPMASK + 64,
  X_CAND = X_NCAND + 1,              /* I[NCAND] lo * 64 + hi */
  X_CBOX = (X_CAND + NCAND + 3) / 4 * 4, /* F[10][4] per car: centre and bounding-box half extents, one 128-bit load */
  X_CMASS = X_CBOX + 4 * DRV_MAXCARS, /* F[10] mass, 1 / mass, 1 / moment of every car: once per launch (phys_substeps) */
  X_CMINV = X_CMASS + DRV_MAXCARS,
  X_CIINV = X_CMINV + DRV_MAXCARS,
  X_CC0 = X_CIINV + DRV_MAXCARS,      /* F[10] cos(road 0 angle - car angle); with X_CC / X_CS refreshed when the angle changes */
  X_PHYS_END = X_CC0 + DRV_MAXCARS,
  /* vision view */
  V_ANG2 = X_UNION,            /* F[NOCC] per occluder: angle of its centre, min / max corner angle ... */
  V_AMN = V_ANG2 + NOCC,
  V_AMX = V_AMN + NOCC,
  V_Q1X = V_AMX + NOCC,        /* ... the corners with the min / max angle and the closest one */
  V_Q1Y = V_Q1X + NOCC,
  V_Q2X = V_Q1Y + NOCC,
  V_Q2Y = V_Q2X + NOCC,
  V_QMX = V_Q2Y + NOCC,
  V_QMY = V_QMX + NOCC,
  V_PX = V_QMY + NOCC,         /* centre in the observer frame */
  V_PY = V_PX + NOCC,
  V_OK = V_PY + NOCC,          /* I bit 0: seen and in the list, bit 1: the closest corner is the min or max one */
  V_STAGE = (V_OK + NOCC + 3) / 4 * 4, /* [DRV_OBS_WORDS] the observation record being built */
  V_END = V_STAGE + DRV_OBS_WORDS,
  /* reset view: Philox blocks of the scene draws */
  R_BLOCKS = X_UNION,          /* U[NRESET_BLOCKS][4] */
  R_END = R_BLOCKS + 4 * NRESET_BLOCKS,
  ENV_SMEM_WORDS_A = X_PHYS_END > V_END ? X_PHYS_END : V_END,
  ENV_SMEM_WORDS = ((ENV_SMEM_WORDS_A > R_END ? ENV_SMEM_WORDS_A : R_END) + 3) / 4 * 4
};

struct Env {
  SharedF f;
  SharedI i;
  SharedCfg c;
  int lane;
};

/* ---- tables (Car.py:9-13) */
__device__ __forceinline__ float car_mass(int t) { return t == 0 ? 1200.f : (t == 1 ? 1800.f : (t == 2 ? 3500.f : 5000.f)); }
__device__ __forceinline__ float car_width(int t) { return 5.f + (float)t; }
__device__ __forceinline__ float car_length(int t) { return 10.f + 5.f * (float)t; }
__device__ __forceinline__ float car_power(int t) { return 3.f + (float)(t & 1); }
__device__ __forceinline__ float car_iinv(int t) {
  float h = car_length(t), w = car_width(t);
  return 1.f / (car_mass(t) * (4.f * h * h + 4.f * w * w) / 12.f); /* moment_for_poly of the box */
}
__device__ __forceinline__ V2 build_pos(int k) { return v2(k < 2 ? 365.f : 1385.f, (k & 1) ? 800.f : 200.f); }

__device__ __forceinline__ bool is_car(int s) { return s >= DRV_SHAPE_CAR0 && s < DRV_SHAPE_PED0; }
__device__ __forceinline__ bool is_ped(int s) { return s >= DRV_SHAPE_PED0 && s < DRV_SHAPE_OBS0; }

__device__ __forceinline__ Rand6 rng_block(const Env E, uint32_t tick, int site, int a, int b, int blk) {
  Rand6 r;
  uint4 w = philox4x32_10(E.c->seed_lo, E.c->seed_hi, (uint32_t)E.i[X_ENVID], (uint32_t)E.i[DRV_S_EPISODE], tick,
                          ((uint32_t)site << 24) | ((uint32_t)(a & 255) << 16) | ((uint32_t)(b & 255) << 8) | (uint32_t)blk);
  r.w[0] = w.x; r.w[1] = w.y; r.w[2] = w.z; r.w[3] = w.w;
  return r;
}

/* Car angles are float64 numbers (words DRV_S_CANG = high, DRV_S_CANGL = low 32 bits): cars start at atan2 of axis
   directions and steer by multiples of 2 degrees, so they spend most of their time EXACTLY at the float64 value of a
   multiple of pi/2 or one rounding away from it, where the sign of the vanishing sine / cosine -- double(pi/2) is
   6.12e-17 short of pi/2 -- decides lane sides in the reference (Road.py:44-52,88-91).  Holding the angle in float64
   and updating it with the reference's own operations reproduces those signs; everything continuous stays float32. */
__device__ __forceinline__ double load_ang(const Env E, int i) { return __hiloint2double(E.i[DRV_S_CANG + i], E.i[DRV_S_CANGL + i]); }
__device__ __forceinline__ void store_ang(const Env E, int i, double a) {
  E.i[DRV_S_CANG + i] = __double2hiint(a);
  E.i[DRV_S_CANGL + i] = __double2loint(a);
}
constexpr double PIO2_D = 1.5707963267948966;
/* sin / cos (float32 accuracy) of a float64 angle, |x| < ~1e4: Cody-Waite reduction by the double-double pi/2 in
   float64 (exact sign and magnitude of the residual next to multiples of pi/2), float32 minimax polynomials */
__device__ __forceinline__ void sincos_d(double x, float* sn, float* cs) {
  double k = rint(x * 0.63661977236758138);
  double rd = fma(k, -1.5707963267948966, x);
  rd = fma(k, -6.123233995736766e-17, rd);
  int q = (int)k;
  float r = (float)rd;
  float r2 = r * r;
  float ps = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
  float pc = fmaf(fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2, -0.5f), r2, 1.0f);
  float s0 = (q & 1) ? pc : ps, c0 = (q & 1) ? ps : pc;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
}

/* ---- two-float positions of the shapes */
__device__ __forceinline__ void shape_pos2(const Env E, int s, float* x, float* xl, float* y, float* yl) {
  if (is_car(s)) {
    int i = s - DRV_SHAPE_CAR0;
    *x = E.f[DRV_S_CPX + i]; *xl = E.f[DRV_S_CPXL + i]; *y = E.f[DRV_S_CPY + i]; *yl = E.f[DRV_S_CPYL + i];
  } else if (is_ped(s)) {
    int i = s - DRV_SHAPE_PED0;
    *x = E.f[DRV_S_PPX + i]; *xl = E.f[DRV_S_PPXL + i]; *y = E.f[DRV_S_PPY + i]; *yl = E.f[DRV_S_PPYL + i];
  } else if (s < DRV_NBUILD) {
    V2 p = build_pos(s);
    *x = p.x; *y = p.y; *xl = 0.f; *yl = 0.f;
  } else {
    *x = E.f[DRV_S_OPX + s - DRV_SHAPE_OBS0]; *y = E.f[DRV_S_OPY + s - DRV_SHAPE_OBS0]; *xl = 0.f; *yl = 0.f;
  }
}
__device__ __forceinline__ V2 pos_diff(const Env E, int b, int a) { /* position of b relative to a */
  float bx, bxl, by, byl, ax, axl, ay, ayl;
  shape_pos2(E, b, &bx, &bxl, &by, &byl);
  shape_pos2(E, a, &ax, &axl, &ay, &ayl);
  return v2(__fadd_rn(__fsub_rn(bx, ax), __fsub_rn(bxl, axl)), __fadd_rn(__fsub_rn(by, ay), __fsub_rn(byl, ayl)));
}
__device__ __forceinline__ float body_minv(const Env E, int s) {
  if (is_car(s)) return E.f[X_CMINV + s - DRV_SHAPE_CAR0];
  return is_ped(s) ? PED_MINV : 0.f;
}
__device__ __forceinline__ float body_iinv(const Env E, int s) {
  if (is_car(s)) return E.f[X_CIINV + s - DRV_SHAPE_CAR0];
  return 0.f; /* pedestrians: contact anchors are parallel to the normal, their spin never changes */
}

/* ---- roads (Road.py; DrivingEnvironment.py:110-115): road 0 runs along y at x = 875 (2 lanes each way, length 1000),
   road 1 along x at y = 500 (1 lane each way, length 1750), lane width 35.  LanePosition: AtGoal 0, InRightLane 1,
   InOpposingLane 2, OverRoad 3, OffRoad 4.  (px, pxl, py, pyl) two-float point, crel = cos(road.angle - angle). */
__device__ __forceinline__ int point_on_road(int road, float px, float pxl, float py, float pyl, float crel) {
  /* Road.py:74-97 */
  float dist = road == 0 ? __fsub_rn(__fsub_rn(875.f, px), pxl) : __fadd_rn(__fsub_rn(py, 500.f), pyl);
  float lim = road == 0 ? 2 * ROAD_W + 5.f : ROAD_W + 5.f;
  if (fabsf(dist) >= lim) return 4;
  float dirDist = road == 0 ? py : px;
  float length = road == 0 ? 1000.f : 1750.f;
  int pos = 3;
  if (dirDist >= -10.f && dirDist <= length + 10.f) {
    float rel = crel * dist;
    pos = rel < 0.f ? 1 : 2;
  }
  return pos;
}
__device__ __forceinline__ bool off_road(float px, float pxl, float py, float pyl) { /* :509-520 */
  int a = point_on_road(0, px, pxl, py, pyl, 1.f), b = point_on_road(1, px, pxl, py, pyl, 1.f);
  return (a < b ? a : b) >= 3;
}
__device__ __forceinline__ bool is_out(float x, float y) { return x <= 0.f || y <= 0.f || x >= W_ || y >= H_; } /* :523-524 */
__device__ __forceinline__ V2 road_dir(int r) { return r == 0 ? v2(0.f, 1.f) : v2(1.f, 0.f); }
__device__ __forceinline__ V2 road_normal(int r) { return r == 0 ? v2(-1.f, 0.f) : v2(0.f, 1.f); }
__device__ __forceinline__ V2 road_p0(int r) { return r == 0 ? v2(875.f, 0.f) : v2(0.f, 500.f); }
__device__ __forceinline__ V2 road_p1(int r) { return r == 0 ? v2(875.f, 1000.f) : v2(1750.f, 500.f); }
__device__ __forceinline__ int road_lanes(int r) { return r == 0 ? 2 : 1; }

/* cos / sin of car i's angle and cos(road 0's angle - angle), cached in the launch's scratch: the angle changes only
   when the car steers or spins, the double-double reductions are needed only then */
__device__ __forceinline__ void refresh_trig(const Env E, int i, double an) {
  float sn, cs, s0, c0;
  sincos_d(an, &sn, &cs);
  sincos_d(__dsub_rn(PIO2_D, an), &s0, &c0); /* road 0: cos(road.angle - angle) with the reference's rounding */
  E.f[X_CC + i] = cs; E.f[X_CS + i] = sn; E.f[X_CC0 + i] = c0;
}

/* ================================================================= per-substep game logic */
/* processAction (substep 0) + tick for car `lane` (DrivingEnvironment.py:357-426, Car.py:55-117) */
__device__ __noinline__ void cars_phase(const Env E, int sub, const float* actions_env) {
  const int i = E.lane;
  if (i >= E.c->P) return;
  float px = E.f[DRV_S_CPX + i], pxl = E.f[DRV_S_CPXL + i], py = E.f[DRV_S_CPY + i], pyl = E.f[DRV_S_CPYL + i];
  double an = load_ang(E, i);
  float vx = E.f[DRV_S_CVX + i], vy = E.f[DRV_S_CVY + i];
  int fl = E.i[DRV_S_CFLAGS + i];
  const int type = E.i[DRV_S_CTYPE + i];
  float rew = E.f[X_REW + i], posrew = E.f[X_POSREW + i];
  float sn = E.f[X_CS + i], cs = E.f[X_CC + i];
  if (sub % 10 == 0) {
    float a0 = actions_env[2 * i], a1 = actions_env[2 * i + 1];
    const bool cont = E.c->continuous != 0;
    float acc = cont ? a0 : a0 - 1.f, steer = cont ? a1 : (a1 - 1.f) * 2.f;
    if (!(fabsf(acc) <= 3.f) || !(fabsf(steer) <= 3.f)) atomicOr(&E.i[X_DIAG], RC_IF_BAD_ACTION); /* DrivingEnvironment.py:364-367 */
    float dx = E.f[DRV_S_CDIRX + i], dy = E.f[DRV_S_CDIRY + i];
    if (!(fl & DRV_CF_FINISHED)) { /* Car.accelerate */
      float power = acc;
      float moveDir = vx * dx + vy * dy;
      bool apply = true;
      if (cont) {
        if (acc * moveDir < 0.f) power *= 2.f;
        else if (acc < 0.f) power *= 0.75f;
      } else {
        if (acc < 0.f) power = acc * 0.75f;
        if (acc == 0.f) power = moveDir == 0.f ? 0.f : (moveDir > 0.f ? -2.f : 2.f);
        else if (acc < 0.f && moveDir > 0.f) apply = false;
        else if (acc > 0.f && moveDir < 0.f) apply = false;
      }
      if (apply) {
        float p = car_power(type) * power;
        vx += p * cs; vy += p * sn;
        if (acc == 0.f && (vx * dx + vy * dy) * moveDir < 0.f) { vx = 0.f; vy = 0.f; }
      }
      if (steer != 0.f) { /* Car.turn */
        float rot = steer * (PI_F / 180.f);
        float rs, rc_;
        sincos_b(rot, &rs, &rc_);
        an += (double)steer * (3.14159265358979323846 / 180);
        V2 d = rot_cs(v2(dx, dy), rc_, rs);
        V2 v = rot_cs(v2(vx, vy), rc_, rs);
        E.f[DRV_S_CDIRX + i] = d.x; E.f[DRV_S_CDIRY + i] = d.y;
        vx = v.x; vy = v.y;
        refresh_trig(E, i, an);
        sn = E.f[X_CS + i]; cs = E.f[X_CC + i];
        store_ang(E, i, an);
      }
    }
  }
  /* tick */
  const float c0 = E.f[X_CC0 + i];
  int l0 = point_on_road(0, px, pxl, py, pyl, c0), l1 = point_on_road(1, px, pxl, py, pyl, cs);
  int lane = l0 < l1 ? l0 : l1;
  float gx = E.f[DRV_S_CGOALX + i], gy = E.f[DRV_S_CGOALY + i];
  float qx = E.f[DRV_S_CPREVX + i], qy = E.f[DRV_S_CPREVY + i];
  /* |prev - goal| - |pos - goal| without the cancellation of two ~1000 px lengths */
  V2 d0 = v2(qx - gx, qy - gy), d1 = v2(__fadd_rn(__fsub_rn(px, gx), pxl), __fadd_rn(__fsub_rn(py, gy), pyl));
  float len0 = len(d0), len1 = len(d1);
  V2 dq = v2(__fsub_rn(__fsub_rn(qx, px), pxl), __fsub_rn(__fsub_rn(qy, py), pyl));
  float den = len0 + len1;
  float diff = den > 0.f ? dot(dq, d0 + d1) / den : 0.f;
  const bool finished0 = fl & DRV_CF_FINISHED;
  if (!finished0) { rew += diff / 50.f; posrew += fmaxf(0.f, diff / 50.f); }
  qx = px; qy = py;
  const float speed = sqrtf(vx * vx + vy * vy);
  if (lane >= 3) {
    if (!finished0) {
      if (lane == 3 && len1 < 100.f) {
        lane = 0;
        fl |= DRV_CF_FINISHED;
        float bonus = (float)(MAXTIME - E.i[DRV_S_ELAPSED]) / 100.f;
        rew += bonus; posrew += bonus;
      } else {
        fl |= DRV_CF_FINISHED | DRV_CF_CRASHED;
        rew -= speed / 5.f;
      }
    }
  } else if (lane == 2) {
    if (!finished0) rew -= speed / 10000.f;
  }
  if (qx >= W_ + 50.f || qx <= -50.f || qy >= H_ + 50.f || qy <= -50.f) { vx = 0.f; vy = 0.f; }
  E.f[DRV_S_CVX + i] = vx; E.f[DRV_S_CVY + i] = vy;
  E.f[DRV_S_CPREVX + i] = qx; E.f[DRV_S_CPREVY + i] = qy;
  E.i[DRV_S_CFLAGS + i] = fl;
  E.i[DRV_S_CLANE + i] = lane;
  E.f[X_REW + i] = rew; E.f[X_POSREW + i] = posrew;
}

/* move(pedestrian) for pedestrian lane - 10 (DrivingEnvironment.py:429-506) */
__device__ __noinline__ void peds_phase(const Env E) {
  const int i = E.lane - 10;
  if (i < 0 || i >= E.i[DRV_S_NPED]) return;
  int fl = E.i[DRV_S_PFLAGS + i];
  if (fl & DRV_PF_DEAD) return;
  float px = E.f[DRV_S_PPX + i], pxl = E.f[DRV_S_PPXL + i], py = E.f[DRV_S_PPY + i], pyl = E.f[DRV_S_PPYL + i];
  const bool offRoad = off_road(px, pxl, py, pyl), out = is_out(px, py);
  const int road = (fl & DRV_PF_ROAD) ? 1 : 0;
  float moving = E.f[DRV_S_PMOVING + i];
  float vx = E.f[DRV_S_PVX + i], vy = E.f[DRV_S_PVY + i];
  if (moving > 0.f) {
    moving = fmaxf(0.f, moving - 10.f);
    if (fl & DRV_PF_CROSSING) {
      if (!(fl & DRV_PF_BEGINCROSSING) && offRoad) { moving = 0.f; fl &= ~DRV_PF_CROSSING; vx = 0.f; vy = 0.f; }
      else if ((fl & DRV_PF_BEGINCROSSING) && !offRoad) fl &= ~DRV_PF_BEGINCROSSING;
    }
    if (out) { moving = 0.f; vx = 0.f; vy = 0.f; }
  } else {
    if (!(fl & DRV_PF_CROSSING)) {
      Rand6 r = rng_block(E, (uint32_t)E.i[DRV_S_ELAPSED], DRV_SITE_PEDMOVE, i, 0, 0);
      moving = (float)randint_from(r.u(0), 5000, 30000);
      int speed = randint_from(r.u(1), -2, 2);
      V2 dir = road_dir(road);
      if (!offRoad) {
        fl = (fl | DRV_PF_CROSSING) & ~DRV_PF_BEGINCROSSING;
        if (speed == 0) speed = 2;
      } else if (out) {
        if (is_out(px + dir.x, py + dir.y)) dir = v2(-dir.x, -dir.y);
      } else if (r.u(2) < 0.05f) {
        fl |= DRV_PF_CROSSING | DRV_PF_BEGINCROSSING;
        V2 n = road_normal(road);
        dir = (fl & DRV_PF_SIDE) ? n : v2(-n.x, -n.y);
        fl ^= DRV_PF_SIDE;
        speed = randint_from(r.u(3), 1, 2);
      }
      float sp = (float)(E.i[DRV_S_PSPEED + i] * speed);
      vx = sp * dir.x; vy = sp * dir.y;
    } else if (offRoad) fl &= ~(DRV_PF_CROSSING | DRV_PF_BEGINCROSSING);
  }
  E.f[DRV_S_PMOVING + i] = moving;
  E.f[DRV_S_PVX + i] = vx; E.f[DRV_S_PVY + i] = vy;
  E.i[DRV_S_PFLAGS + i] = fl;
}

/* ================================================================= collision callbacks (serial, all `begin`) */
__device__ __forceinline__ void car_crash(const Env E, int i) { E.i[DRV_S_CFLAGS + i] |= DRV_CF_FINISHED | DRV_CF_CRASHED; }
__device__ __forceinline__ V2 car_vel(const Env E, int i) { return v2(E.f[DRV_S_CVX + i], E.f[DRV_S_CVY + i]); }
/* cos(angle(dp) - angle(v)) */
__device__ __forceinline__ float cos_between(V2 dp, V2 v) { return dot(dp, v) / (len(dp) * len(v)); }

__device__ void cb_car_crash(const Env E, int i1, int i2) {
  /* DrivingEnvironment.py:591-637 */
  V2 v1 = car_vel(E, i1), v2_ = car_vel(E, i2);
  float v1l = len(v1) / 5.f, v2l = len(v2_) / 5.f;
  bool c1 = E.i[DRV_S_CFLAGS + i1] & DRV_CF_CRASHED, c2 = E.i[DRV_S_CFLAGS + i2] & DRV_CF_CRASHED;
  int l1 = E.i[DRV_S_CLANE + i1], l2 = E.i[DRV_S_CLANE + i2];
  float r1 = 0.f, r2 = 0.f;
  if (!c1) r1 -= v1l;
  if (!c2) r2 -= v2l;
  if (l1 != 1 && !c1) r1 -= v1l;
  if (l2 != 1 && !c2) r2 -= v2l;
  if (l1 == 1 && l2 == 1) {
    V2 dp = pos_diff(E, DRV_SHAPE_CAR0 + i1, DRV_SHAPE_CAR0 + i2);
    if (len(v1) > 1.f && cos_between(dp, v1) < -0.4f && !c1) r1 -= v1l;
    if (len(v2_) > 1.f && cos_between(dp, v2_) > 0.4f && !c2) r2 -= v2l;
  }
  E.f[X_REW + i1] += r1;
  E.f[X_REW + i2] += r2;
  car_crash(E, i1);
  car_crash(E, i2);
}
__device__ bool cb_ped_hit(const Env E, int ci, int pi) {
  /* :640-667; false = the pair is ignored */
  V2 v1 = car_vel(E, ci);
  float v1l = len(v1);
  if (!(v1l > 1.f)) return false;
  E.f[DRV_S_PMOVING + pi] = 0.f; E.f[DRV_S_PVX + pi] = 0.f; E.f[DRV_S_PVY + pi] = 0.f; /* Pedestrian.die */
  E.i[DRV_S_PFLAGS + pi] |= DRV_PF_DEAD;
  V2 dp = pos_diff(E, DRV_SHAPE_CAR0 + ci, DRV_SHAPE_PED0 + pi);
  if (cos_between(dp, v1) < -0.4f && !(E.i[DRV_S_CFLAGS + ci] & DRV_CF_FINISHED)) {
    car_crash(E, ci);
    E.f[X_REW + ci] -= v1l / 5.f;
  }
  return true;
}
__device__ void cb_car_hit(const Env E, int ci) {
  /* :670-683 */
  if (!(E.i[DRV_S_CFLAGS + ci] & DRV_CF_FINISHED)) E.f[X_REW + ci] -= len(car_vel(E, ci)) / 5.f;
  car_crash(E, ci);
}

/* ================================================================= narrow phase helpers */
__device__ __forceinline__ float clamp01(float t) { return fminf(fmaxf(t, 0.f), 1.f); }
__device__ __forceinline__ V2 vlerp(V2 a, V2 b, float t) { return a * (1.f - t) + b * t; }
__device__ __forceinline__ V2 closest_on_seg(V2 p, V2 a, V2 b) {
  V2 d = b - a;
  float t = clamp01(dot(d, p - a) / lensq(d));
  return a + d * t;
}

/* ================================================================= space.step */
struct BodyV {
  float vx, vy, w, vbx, vby, wb;
};
__device__ __forceinline__ BodyV load_body(const Env E, int s) {
  BodyV b;
  if (is_car(s)) {
    int i = s - DRV_SHAPE_CAR0;
    b.vx = E.f[DRV_S_CVX + i]; b.vy = E.f[DRV_S_CVY + i]; b.w = E.f[DRV_S_CW + i];
    b.vbx = E.f[DRV_S_CVBX + i]; b.vby = E.f[DRV_S_CVBY + i]; b.wb = E.f[DRV_S_CWB + i];
  } else if (is_ped(s)) {
    int i = s - DRV_SHAPE_PED0;
    b.vx = E.f[DRV_S_PVX + i]; b.vy = E.f[DRV_S_PVY + i]; b.w = 0.f;
    b.vbx = E.f[DRV_S_PVBX + i]; b.vby = E.f[DRV_S_PVBY + i]; b.wb = 0.f;
  } else {
    b.vx = b.vy = b.w = b.vbx = b.vby = b.wb = 0.f;
  }
  return b;
}
__device__ __forceinline__ void store_body(const Env E, int s, const BodyV& b) {
  if (is_car(s)) {
    int i = s - DRV_SHAPE_CAR0;
    E.f[DRV_S_CVX + i] = b.vx; E.f[DRV_S_CVY + i] = b.vy; E.f[DRV_S_CW + i] = b.w;
    E.f[DRV_S_CVBX + i] = b.vbx; E.f[DRV_S_CVBY + i] = b.vby; E.f[DRV_S_CWB + i] = b.wb;
  } else if (is_ped(s)) {
    int i = s - DRV_SHAPE_PED0;
    E.f[DRV_S_PVX + i] = b.vx; E.f[DRV_S_PVY + i] = b.vy;
    E.f[DRV_S_PVBX + i] = b.vbx; E.f[DRV_S_PVBY + i] = b.vby;
  }
}

/* cpArbiterApplyImpulse for every active contact, in order (Gauss-Seidel): lane 0.  Shape friction is 0
   everywhere, so the tangential impulse is clamped to 0 and not computed. */
__device__ __noinline__ void contacts_iteration(const Env E) {
  int nact = E.i[X_NACTIVE];
  for (int q = 0; q < nact; q++) {
    int k = E.i[X_ACTIVE + q];
    int key = E.i[DRV_S_ARBKEY + k];
    int a = key & 255, b = (key >> 8) & 255;
    int ncon = E.i[DRV_S_ARBCON + k] & 255;
    V2 n = v2(E.f[X_ANX + q], E.f[X_ANY + q]);
    float ami = body_minv(E, a), aii = body_iinv(E, a), bmi = body_minv(E, b), bii = body_iinv(E, b);
    BodyV A = load_body(E, a), B = load_body(E, b);
    for (int c = 0; c < ncon; c++) {
      int ci = 2 * q + c;
      V2 r1 = v2(E.f[X_CR1X + ci], E.f[X_CR1Y + ci]), r2 = v2(E.f[X_CR2X + ci], E.f[X_CR2Y + ci]);
      float nMass = E.f[X_CNMASS + ci];
      V2 vb1 = v2(A.vbx - r1.y * A.wb, A.vby + r1.x * A.wb);
      V2 vb2 = v2(B.vbx - r2.y * B.wb, B.vby + r2.x * B.wb);
      V2 vr = v2((B.vx - r2.y * B.w) - (A.vx - r1.y * A.w), (B.vy + r2.x * B.w) - (A.vy + r1.x * A.w));
      float vbn = dot(vb2 - vb1, n);
      float vrn = dot(vr, n);
      float jbn = (E.f[X_CBIAS + ci] - vbn) * nMass;
      float jbnOld = E.f[X_CJBIAS + ci];
      float jBias = fmaxf(jbnOld + jbn, 0.f);
      float jn = -(E.f[X_CBOUNCE + ci] + vrn) * nMass;
      float jnOld = E.f[X_CJN + ci];
      float jnAcc = fmaxf(jnOld + jn, 0.f);
      E.f[X_CJBIAS + ci] = jBias;
      E.f[X_CJN + ci] = jnAcc;
      V2 jb = n * (jBias - jbnOld);
      A.vbx -= jb.x * ami; A.vby -= jb.y * ami; A.wb -= aii * cross(r1, jb);
      B.vbx += jb.x * bmi; B.vby += jb.y * bmi; B.wb += bii * cross(r2, jb);
      V2 j = n * (jnAcc - jnOld);
      A.vx -= j.x * ami; A.vy -= j.y * ami; A.w -= aii * cross(r1, j);
      B.vx += j.x * bmi; B.vy += j.y * bmi; B.w += bii * cross(r2, j);
    }
    store_body(E, a, A);
    store_body(E, b, B);
  }
}

/* cached pairs that did not collide this substep: age, expiry after 3 substeps; lane = cached pair */
__device__ __noinline__ void age_cached_pairs(const Env E) {
  const int lane = E.lane;
  int nArb = E.i[DRV_S_NARB];
  unsigned touched = (unsigned)E.i[X_TOUCHED];
  int key = 0, con = 0;
  float jn0 = 0, jt0 = 0, jn1 = 0, jt1 = 0;
  bool have = lane < nArb;
  bool keep = false;
  if (have) {
    key = E.i[DRV_S_ARBKEY + lane]; con = E.i[DRV_S_ARBCON + lane];
    jn0 = E.f[DRV_S_ARBJN0 + lane]; jt0 = E.f[DRV_S_ARBJT0 + lane];
    jn1 = E.f[DRV_S_ARBJN1 + lane]; jt1 = E.f[DRV_S_ARBJT1 + lane];
    keep = true;
    if (!((touched >> lane) & 1)) {
      int age = ((key >> 24) & 255) + 1;
      key = (key & 0xffff) | (RC_ARB_CACHED << 16) | (age << 24);
      keep = age < 3;
    }
  }
  unsigned keepmask = __ballot_sync(FULLMASK, keep);
  __syncwarp();
  int newIdx = __popc(keepmask & ((1u << lane) - 1));
  if (keep) {
    E.i[DRV_S_ARBKEY + newIdx] = key; E.i[DRV_S_ARBCON + newIdx] = con;
    E.f[DRV_S_ARBJN0 + newIdx] = jn0; E.f[DRV_S_ARBJT0 + newIdx] = jt0;
    E.f[DRV_S_ARBJN1 + newIdx] = jn1; E.f[DRV_S_ARBJT1 + newIdx] = jt1;
  }
  int nact = E.i[X_NACTIVE];
  int myActive = lane < nact ? E.i[X_ACTIVE + lane] : 0;
  int remapped = __shfl_sync(FULLMASK, newIdx, myActive);
  __syncwarp();
  if (lane < nact) E.i[X_ACTIVE + lane] = remapped;
  if (lane == 0) E.i[DRV_S_NARB] = __popc(keepmask);
  if (have && lane >= __popc(keepmask)) { /* records stay canonical: unused slots are zero */
    E.i[DRV_S_ARBKEY + lane] = 0; E.i[DRV_S_ARBCON + lane] = 0;
    E.f[DRV_S_ARBJN0 + lane] = 0.f; E.f[DRV_S_ARBJT0 + lane] = 0.f;
    E.f[DRV_S_ARBJN1 + lane] = 0.f; E.f[DRV_S_ARBJT1 + lane] = 0.f;
  }
  __syncwarp();
}

/* cached contact impulses re-applied in pair order (cpArbiterApplyCachedImpulse): lane 0 */
__device__ __noinline__ void warm_start_contacts(const Env E, float dt_coef) {
  if (E.lane == 0) {
    const int nact = E.i[X_NACTIVE];
    for (int q = 0; q < nact; q++) {
      int k = E.i[X_ACTIVE + q];
      int key = E.i[DRV_S_ARBKEY + k];
      if (((key >> 16) & 255) == RC_ARB_FIRST) continue;
      int a = key & 255, b = (key >> 8) & 255;
      int ncon = E.i[DRV_S_ARBCON + k] & 255;
      V2 n = v2(E.f[X_ANX + q], E.f[X_ANY + q]);
      BodyV A = load_body(E, a), B = load_body(E, b);
      float ami = body_minv(E, a), aii = body_iinv(E, a), bmi = body_minv(E, b), bii = body_iinv(E, b);
      for (int c = 0; c < ncon; c++) {
        int ci = 2 * q + c;
        V2 j = n * (E.f[X_CJN + ci] * dt_coef);
        V2 r1 = v2(E.f[X_CR1X + ci], E.f[X_CR1Y + ci]), r2 = v2(E.f[X_CR2X + ci], E.f[X_CR2Y + ci]);
        A.vx -= j.x * ami; A.vy -= j.y * ami; A.w -= aii * cross(r1, j);
        B.vx += j.x * bmi; B.vy += j.y * bmi; B.w += bii * cross(r2, j);
      }
      store_body(E, a, A);
      store_body(E, b, B);
    }
  }
  __syncwarp();
}

/* shape pairs whose bounding boxes overlap (X_BPMASK) -> X_CAND, in Chipmunk's ascending (i, j) insertion order */
__device__ __forceinline__ void build_candidates(const Env E) {
  const int lane = E.lane;
  const int N = E.c->P;
  int m0 = E.i[X_BPMASK + lane], m1 = E.i[X_BPMASK + 32 + lane];
  if (lane == 0) E.i[X_NCAND] = 0;
  __syncwarp();
  /* buildings first (lo = 0..3), then per car lo = 4 + i every shape above it */
  for (int s = 0; s < DRV_NBUILD; s++) {
    int m = __shfl_sync(FULLMASK, m0, s);
    if (lane == 0) {
      int n = E.i[X_NCAND];
      while (m) {
        int i = __ffs(m) - 1;
        m &= m - 1;
        if (n < NCAND) E.i[X_CAND + n++] = s * 64 + DRV_SHAPE_CAR0 + i;
        else E.i[X_DIAG] |= RC_IF_ACTIVE_OVERFLOW;
      }
      E.i[X_NCAND] = n;
    }
  }
  for (int i = 0; i < N; i++) {
    const int lo = DRV_SHAPE_CAR0 + i;
    unsigned b0 = __ballot_sync(FULLMASK, ((m0 >> i) & 1) && lane > lo);
    unsigned b1 = __ballot_sync(FULLMASK, (m1 >> i) & 1);
    if (lane == 0 && (b0 | b1)) {
      int n = E.i[X_NCAND];
      while (b0) {
        int hi = __ffs(b0) - 1;
        b0 &= b0 - 1;
        if (n < NCAND) E.i[X_CAND + n++] = lo * 64 + hi;
        else E.i[X_DIAG] |= RC_IF_ACTIVE_OVERFLOW;
      }
      while (b1) {
        int hi = 32 + __ffs(b1) - 1;
        b1 &= b1 - 1;
        if (n < NCAND) E.i[X_CAND + n++] = lo * 64 + hi;
        else E.i[X_DIAG] |= RC_IF_ACTIVE_OVERFLOW;
      }
      E.i[X_NCAND] = n;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void friction(float m, float fr, float rfr, float& x, float& y, float& theta) {
  /* cutils.apply_friction (cutils.py:102-140) after the default velocity update (no forces, damping 1) */
  float factor = fr * m, rotFactor = rfr * m;
  float length = 1.f / (fabsf(x) + fabsf(y) + 1e-5f);
  float a0 = x * factor * length, a1 = y * factor * length;
  x = fabsf(x) < factor ? 0.f : x - a0;
  y = fabsf(y) < factor ? 0.f : y - a1;
  theta = fabsf(theta) < rotFactor ? 0.f : theta - (theta > 0.f ? rotFactor : -rotFactor);
}

/* ---- narrow phase in registers: one candidate pair per lane, everything unrolled (no local memory) */
struct BoxR { /* vertices CCW (-h,-w), (h,-w), (h,w), (-h,w); plane normal i belongs to edge v[i-1] -> v[i] */
  V2 v0, v1, v2, v3;
  float c, s;
};
__device__ __forceinline__ BoxR make_boxr(V2 ctr, float c, float s, float hx, float hy) {
  BoxR b;
  b.v0 = v2(ctr.x + -hx * c - -hy * s, ctr.y + -hx * s + -hy * c);
  b.v1 = v2(ctr.x + hx * c - -hy * s, ctr.y + hx * s + -hy * c);
  b.v2 = v2(ctr.x + hx * c - hy * s, ctr.y + hx * s + hy * c);
  b.v3 = v2(ctr.x + -hx * c - hy * s, ctr.y + -hx * s + hy * c);
  b.c = c; b.s = s;
  return b;
}
__device__ __forceinline__ V2 bvert(const BoxR& b, int i) { return i == 0 ? b.v0 : (i == 1 ? b.v1 : (i == 2 ? b.v2 : b.v3)); }
__device__ __forceinline__ V2 bnorm(const BoxR& b, int i) {
  return i == 0 ? v2(-b.c, -b.s) : (i == 1 ? v2(b.s, -b.c) : (i == 2 ? v2(b.c, b.s) : v2(-b.s, b.c)));
}
__device__ __forceinline__ BoxR shape_boxr(const Env E, int s, V2 ctr) {
  if (is_car(s)) {
    int i = s - DRV_SHAPE_CAR0, t = E.i[DRV_S_CTYPE + i];
    return make_boxr(ctr, E.f[X_CC + i], E.f[X_CS + i], car_length(t), car_width(t));
  }
  if (s < DRV_NBUILD) return make_boxr(ctr, 1.f, 0.f, BUILD_W, BUILD_H);
  return make_boxr(ctr, 1.f, 0.f, OBST_HALF, OBST_HALF);
}
struct ManR {
  float nx, ny;
  V2 p1a, p2a, p1b, p2b; /* contact 0 / 1 */
  int h0, h1, cnt;
};
__device__ __forceinline__ void sat_axis(V2 ax, const BoxR& own, const BoxR& other, float sign, float& best, V2& bn) {
  float mn = fminf(fminf(dot(ax, other.v0), dot(ax, other.v1)), fminf(dot(ax, other.v2), dot(ax, other.v3)));
  float mx = fmaxf(fmaxf(dot(ax, own.v0), dot(ax, own.v1)), fmaxf(dot(ax, own.v2), dot(ax, own.v3)));
  if (mn - mx > best) { best = mn - mx; bn = v2(ax.x * sign, ax.y * sign); }
}
__device__ __forceinline__ void support_edge_r(const BoxR& s, V2 n, V2* pa, int* ha, V2* pb, int* hb) {
  int i1 = 0;
  float best = dot(s.v0, n), d;
  d = dot(s.v1, n); if (d > best) { best = d; i1 = 1; }
  d = dot(s.v2, n); if (d > best) { best = d; i1 = 2; }
  d = dot(s.v3, n); if (d > best) { best = d; i1 = 3; }
  int i0 = (i1 + 3) & 3, i2 = (i1 + 1) & 3;
  if (dot(n, bnorm(s, i1)) > dot(n, bnorm(s, i2))) { *pa = bvert(s, i0); *ha = i0; *pb = bvert(s, i1); *hb = i1; }
  else { *pa = bvert(s, i1); *ha = i1; *pb = bvert(s, i2); *hb = i2; }
}
/* box - box: minimum-penetration axis among both boxes' edge normals, then Chipmunk's ContactPoints */
__device__ __forceinline__ void box_box_r(const BoxR& A, const BoxR& B, ManR* m) {
  float best = -3.0e38f;
  V2 bn = v2(1.f, 0.f);
  sat_axis(bnorm(A, 1), A, B, 1.f, best, bn); sat_axis(bnorm(A, 2), A, B, 1.f, best, bn);
  sat_axis(bnorm(A, 3), A, B, 1.f, best, bn); sat_axis(bnorm(A, 0), A, B, 1.f, best, bn);
  sat_axis(bnorm(B, 1), B, A, -1.f, best, bn); sat_axis(bnorm(B, 2), B, A, -1.f, best, bn);
  sat_axis(bnorm(B, 3), B, A, -1.f, best, bn); sat_axis(bnorm(B, 0), B, A, -1.f, best, bn);
  m->cnt = 0;
  if (!(best <= 0.f)) return;
  V2 n = bn;
  m->nx = n.x; m->ny = n.y;
  V2 e1a, e1b, e2a, e2b;
  int h1a, h1b, h2a, h2b;
  support_edge_r(A, n, &e1a, &h1a, &e1b, &h1b);
  support_edge_r(B, v2(-n.x, -n.y), &e2a, &h2a, &e2b, &h2b);
  float d_e1_a = cross(e1a, n), d_e1_b = cross(e1b, n), d_e2_a = cross(e2a, n), d_e2_b = cross(e2b, n);
  float e1_denom = 1.f / (d_e1_b - d_e1_a + 1.1754944e-38f), e2_denom = 1.f / (d_e2_b - d_e2_a + 1.1754944e-38f);
  int cnt = 0;
  {
    V2 p1 = vlerp(e1a, e1b, clamp01((d_e2_b - d_e1_a) * e1_denom));
    V2 p2 = vlerp(e2a, e2b, clamp01((d_e1_a - d_e2_a) * e2_denom));
    if (dot(p2 - p1, n) <= 0.f) { m->p1a = p1; m->p2a = p2; m->h0 = 1 + h1a + 4 * h2b; cnt = 1; }
  }
  {
    V2 p1 = vlerp(e1a, e1b, clamp01((d_e2_a - d_e1_a) * e1_denom));
    V2 p2 = vlerp(e2a, e2b, clamp01((d_e1_b - d_e2_a) * e2_denom));
    if (dot(p2 - p1, n) <= 0.f) {
      int h = 1 + h1b + 4 * h2a;
      if (cnt == 0) { m->p1a = p1; m->p2a = p2; m->h0 = h; }
      else { m->p1b = p1; m->p2b = p2; m->h1 = h; }
      cnt++;
    }
  }
  m->cnt = cnt;
}
/* closest point of edge a -> b (plane normal nrm) to the origin; inside the edge it is the foot of the perpendicular,
   taken along the plane normal: for an axis-aligned box the off-axis component then vanishes like in float64 instead
   of carrying the rounding of a + d t (1e-7, which a fast car turns into a spurious impulse and spin) */
__device__ __forceinline__ V2 edge_closest(V2 a, V2 b, V2 nrm) {
  V2 dv = b - a;
  float t = -dot(dv, a) / lensq(dv);
  return t <= 0.f ? a : (t >= 1.f ? b : nrm * dot(nrm, a));
}
/* circle of radius cr at the origin (a) - box (b): CircleToPoly */
__device__ __forceinline__ void circle_box_r(float cr, const BoxR& P, ManR* m) {
  /* inside: the centre is on the inner side of every edge */
  bool inside = !(cross(P.v1 - P.v0, v2(-P.v0.x, -P.v0.y)) < 0.f) && !(cross(P.v2 - P.v1, v2(-P.v1.x, -P.v1.y)) < 0.f) &&
                !(cross(P.v3 - P.v2, v2(-P.v2.x, -P.v2.y)) < 0.f) && !(cross(P.v0 - P.v3, v2(-P.v3.x, -P.v3.y)) < 0.f);
  float d;
  V2 pb, n;
  if (inside) {
    float best = -dot(P.v0, bnorm(P, 0)), q;
    int bi = 0;
    q = -dot(P.v1, bnorm(P, 1)); if (q > best) { best = q; bi = 1; }
    q = -dot(P.v2, bnorm(P, 2)); if (q > best) { best = q; bi = 2; }
    q = -dot(P.v3, bnorm(P, 3)); if (q > best) { best = q; bi = 3; }
    V2 nb = bnorm(P, bi);
    n = v2(-nb.x, -nb.y);
    pb = closest_on_seg(v2(0.f, 0.f), bvert(P, (bi + 3) & 3), bvert(P, bi));
    d = best;
  } else {
    V2 q = edge_closest(P.v3, P.v0, bnorm(P, 0));
    d = len(q); pb = q;
    float dd;
    q = edge_closest(P.v0, P.v1, bnorm(P, 1)); dd = len(q); if (dd < d) { d = dd; pb = q; }
    q = edge_closest(P.v1, P.v2, bnorm(P, 2)); dd = len(q); if (dd < d) { d = dd; pb = q; }
    q = edge_closest(P.v2, P.v3, bnorm(P, 3)); dd = len(q); if (dd < d) { d = dd; pb = q; }
    n = d > 0.f ? v2(pb.x / d, pb.y / d) : v2(1.f, 0.f);
  }
  m->cnt = 0;
  if (d <= cr) {
    m->nx = n.x; m->ny = n.y;
    m->p1a = n * cr;
    m->p2a = pb;
    m->h0 = 0;
    m->cnt = 1;
  }
}

__device__ __forceinline__ void friction_phase(const Env E, bool isCar, bool isPed) {
  /* velocity update through the friction callbacks: cars (cutils.py:78-85), dead pedestrians (:88-90); live
     pedestrians have the default update (nothing to do without forces) */
  const int lane = E.lane;
  if (isCar) {
    const int i = lane;
    float x = E.f[DRV_S_CVX + i], y = E.f[DRV_S_CVY + i], th = E.f[DRV_S_CW + i];
    const bool fin = E.i[DRV_S_CFLAGS + i] & DRV_CF_FINISHED;
    friction(E.f[X_CMASS + i], fin ? 5e-4f : 5e-5f, fin ? 2e-5f : 1e-5f, x, y, th);
    E.f[DRV_S_CVX + i] = x; E.f[DRV_S_CVY + i] = y; E.f[DRV_S_CW + i] = th;
  } else if (isPed && (E.i[DRV_S_PFLAGS + lane - 10] & DRV_PF_DEAD)) {
    const int i = lane - 10;
    float x = E.f[DRV_S_PVX + i], y = E.f[DRV_S_PVY + i], th = 0.f;
    friction(90.f, 5e-2f, 2e-4f, x, y, th);
    E.f[DRV_S_PVX + i] = x; E.f[DRV_S_PVY + i] = y;
  }
  __syncwarp();
}

/* The collision half of cpSpaceStep for an environment with bounding-box overlapping pairs or cached pairs:
   narrow phase (lane = candidate pair), cpArbiterUpdate, begin callbacks (serial, in pair order), ageing of the
   cached pairs, pre-step, the friction velocity update, warm start and the 10 solver iterations.  Pairs that share
   no dynamic body are independent systems and iterate in the registers of their lane; otherwise the contacts are
   solved in pair order on lane 0 (Gauss-Seidel like the reference). */
/* keepMask: cached pairs of bodies that nothing can reach in this substep -- neither collided nor aged */
__device__ __noinline__ void contacts_substep(const Env E, int prevStepped, bool haveCand, bool isCar, bool isPed, unsigned keepMask) {
  const int lane = E.lane;
  const unsigned lt = (1u << lane) - 1;
  int ncand = 0;
  if (haveCand) {
    build_candidates(E);
    ncand = E.i[X_NCAND];
  }
  /* ---- narrow phase */
  ManR m;
  m.cnt = 0; m.nx = m.ny = 0.f; m.h0 = m.h1 = 0;
  m.p1a = m.p2a = m.p1b = m.p2b = v2(0.f, 0.f);
  int a = 0, b = 0;
  V2 dab = v2(0.f, 0.f);
  if (lane < ncand) {
    const int key = E.i[X_CAND + lane];
    a = key >> 6; b = key & 63;
    if (!is_ped(a) && is_ped(b)) { int t = a; a = b; b = t; } /* cpCollide orders by primitive type: circle first */
    dab = pos_diff(E, b, a);
    BoxR B = shape_boxr(E, b, dab);
    if (is_ped(a)) circle_box_r(PED_R, B, &m);
    else {
      BoxR A = shape_boxr(E, a, v2(0.f, 0.f));
      box_box_r(A, B, &m);
    }
  }
  bool hit = m.cnt > 0;
  const int nArb0 = E.i[DRV_S_NARB];
  if (!__any_sync(FULLMASK, hit) && nArb0 == 0) {
    friction_phase(E, isCar, isPed);
    return;
  }
  /* ---- cpArbiterUpdate: find or create the cached pair */
  const int key16 = a | (b << 8);
  int k = -1;
  if (hit)
    for (int kk = 0; kk < nArb0; kk++)
      if ((E.i[DRV_S_ARBKEY + kk] & 0xffff) == key16) { k = kk; break; }
  const bool isNew = hit && k < 0;
  const unsigned newMask = __ballot_sync(FULLMASK, isNew);
  if (isNew) {
    k = nArb0 + __popc(newMask & lt);
    if (k >= DRV_MAXARB) { hit = false; E.i[X_DIAG] |= RC_IF_ARB_OVERFLOW; }
  }
  int state = RC_ARB_FIRST, oldCon = 0;
  float ojn0 = 0.f, ojn1 = 0.f;
  if (hit && !isNew) {
    state = (E.i[DRV_S_ARBKEY + k] >> 16) & 255;
    if (state == RC_ARB_CACHED) state = RC_ARB_FIRST;
    oldCon = E.i[DRV_S_ARBCON + k];
    ojn0 = E.f[DRV_S_ARBJN0 + k]; ojn1 = E.f[DRV_S_ARBJN1 + k];
  }
  __syncwarp();
  if (lane == 0) E.i[DRV_S_NARB] = min(nArb0 + __popc(newMask), DRV_MAXARB);
  /* ---- begin callbacks of first-touch pairs, in pair order (rare) */
  unsigned firstMask = __ballot_sync(FULLMASK, hit && state == RC_ARB_FIRST);
  unsigned ignoreMask = 0;
  while (firstMask) {
    const int src = __ffs(firstMask) - 1;
    firstMask &= firstMask - 1;
    const int aa = __shfl_sync(FULLMASK, a, src), bb = __shfl_sync(FULLMASK, b, src);
    int ign = 0;
    if (lane == 0) {
      if (is_car(aa) && is_car(bb)) cb_car_crash(E, aa - DRV_SHAPE_CAR0, bb - DRV_SHAPE_CAR0);
      else if (is_ped(aa)) ign = cb_ped_hit(E, bb - DRV_SHAPE_CAR0, aa - DRV_SHAPE_PED0) ? 0 : 1;
      else cb_car_hit(E, (is_car(aa) ? aa : bb) - DRV_SHAPE_CAR0);
    }
    ign = __shfl_sync(FULLMASK, ign, 0);
    if (ign) ignoreMask |= 1u << src;
    __syncwarp();
  }
  if (hit && ((ignoreMask >> lane) & 1)) state = RC_ARB_IGNORE;
  /* ---- cached-pair records, active list (pair order = lane order) */
  bool act = hit && state != RC_ARB_IGNORE;
  unsigned actMask = __ballot_sync(FULLMASK, act);
  int q = __popc(actMask & lt);
  if (act && q >= DRV_MAXACTIVE) { act = false; E.i[X_DIAG] |= RC_IF_ACTIVE_OVERFLOW; }
  actMask = __ballot_sync(FULLMASK, act);
  const int nact = __popc(actMask);
  const unsigned touched = __reduce_or_sync(FULLMASK, hit ? (1u << k) : 0u) | keepMask;
  float jn0 = 0.f, jn1 = 0.f;
  if (hit) {
    E.i[DRV_S_ARBKEY + k] = key16 | (state << 16); /* age = 0 */
    if (act) {
      const int oldN = oldCon & 255, oh0 = (oldCon >> 8) & 255, oh1 = (oldCon >> 16) & 255;
      if (oldN > 0 && m.h0 == oh0) jn0 = ojn0;
      if (oldN > 1 && m.h0 == oh1) jn0 = ojn1;
      if (m.cnt > 1) {
        if (oldN > 0 && m.h1 == oh0) jn1 = ojn0;
        if (oldN > 1 && m.h1 == oh1) jn1 = ojn1;
      }
      E.i[DRV_S_ARBCON + k] = m.cnt | (m.h0 << 8) | ((m.cnt > 1 ? m.h1 : 0) << 16);
      E.i[X_ACTIVE + q] = k;
    } else {
      E.i[DRV_S_ARBCON + k] = 0;
      E.f[DRV_S_ARBJN0 + k] = 0.f; E.f[DRV_S_ARBJN1 + k] = 0.f;
    }
    if (isNew) { E.f[DRV_S_ARBJT0 + k] = 0.f; E.f[DRV_S_ARBJT1 + k] = 0.f; }
  }
  if (lane == 0) { E.i[X_TOUCHED] = (int)touched; E.i[X_NACTIVE] = nact; }
  __syncwarp();
  age_cached_pairs(E);
  if (nact == 0) {
    friction_phase(E, isCar, isPed);
    return;
  }
  if (act) k = E.i[X_ACTIVE + q];
  /* ---- cpArbiterPreStep (velocities before the friction update) */
  const float slop = 0.1f;
  const float biasCoef = 6.1259606640e-02f; /* 1 - collisionBias^dt, collisionBias = 0.9^60 */
  const V2 n = v2(m.nx, m.ny);
  const V2 r1a = m.p1a, r2a = m.p2a - dab, r1b = m.p1b, r2b = m.p2b - dab;
  float ami = 0.f, aii = 0.f, bmi = 0.f, bii = 0.f;
  float nMass0 = 0.f, nMass1 = 0.f, bias0 = 0.f, bias1 = 0.f, bounce0 = 0.f, bounce1 = 0.f;
  if (act) {
    ami = body_minv(E, a); aii = body_iinv(E, a); bmi = body_minv(E, b); bii = body_iinv(E, b);
    BodyV A = load_body(E, a), B = load_body(E, b);
    {
      float r1cn = cross(r1a, n), r2cn = cross(r2a, n);
      nMass0 = 1.f / (ami + aii * r1cn * r1cn + bmi + bii * r2cn * r2cn);
      bias0 = -biasCoef * fminf(0.f, dot(r2a - r1a + dab, n) + slop) / DT;
      V2 vr = v2((B.vx - r2a.y * B.w) - (A.vx - r1a.y * A.w), (B.vy + r2a.x * B.w) - (A.vy + r1a.x * A.w));
      bounce0 = dot(vr, n) * (0.05f * 0.05f); /* every shape: elasticity 0.05 */
    }
    if (m.cnt > 1) {
      float r1cn = cross(r1b, n), r2cn = cross(r2b, n);
      nMass1 = 1.f / (ami + aii * r1cn * r1cn + bmi + bii * r2cn * r2cn);
      bias1 = -biasCoef * fminf(0.f, dot(r2b - r1b + dab, n) + slop) / DT;
      V2 vr = v2((B.vx - r2b.y * B.w) - (A.vx - r1b.y * A.w), (B.vy + r2b.x * B.w) - (A.vy + r1b.x * A.w));
      bounce1 = dot(vr, n) * (0.05f * 0.05f);
    }
  }
  /* do two active pairs share a dynamic body? */
  bool conflict = false;
  {
    const bool aDyn = a >= DRV_SHAPE_CAR0 && a < DRV_SHAPE_OBS0, bDyn = b >= DRV_SHAPE_CAR0 && b < DRV_SHAPE_OBS0;
    unsigned mm = actMask;
    while (mm) {
      const int src = __ffs(mm) - 1;
      mm &= mm - 1;
      const int aj = __shfl_sync(FULLMASK, a, src), bj = __shfl_sync(FULLMASK, b, src);
      if (act && src != lane && ((aDyn && (a == aj || a == bj)) || (bDyn && (b == aj || b == bj)))) conflict = true;
    }
    conflict = __any_sync(FULLMASK, conflict);
  }
  __syncwarp();
  friction_phase(E, isCar, isPed);
  const float dt_coef = prevStepped ? 1.f : 0.f;
  if (!conflict) {
    if (act) {
      BodyV A = load_body(E, a), B = load_body(E, b);
      if (state != RC_ARB_FIRST) { /* cpArbiterApplyCachedImpulse */
        V2 j = n * (jn0 * dt_coef);
        A.vx -= j.x * ami; A.vy -= j.y * ami; A.w -= aii * cross(r1a, j);
        B.vx += j.x * bmi; B.vy += j.y * bmi; B.w += bii * cross(r2a, j);
        if (m.cnt > 1) {
          j = n * (jn1 * dt_coef);
          A.vx -= j.x * ami; A.vy -= j.y * ami; A.w -= aii * cross(r1b, j);
          B.vx += j.x * bmi; B.vy += j.y * bmi; B.w += bii * cross(r2b, j);
        }
      }
      float jb0 = 0.f, jb1 = 0.f;
#pragma unroll 1
      for (int it = 0; it < 10; it++) { /* cpArbiterApplyImpulse; shape friction is 0: no tangential impulse */
#pragma unroll
        for (int c = 0; c < 2; c++) {
          if (c == 1 && m.cnt < 2) break;
          const V2 r1 = c ? r1b : r1a, r2 = c ? r2b : r2a;
          const float nMass = c ? nMass1 : nMass0, bias = c ? bias1 : bias0, bounce = c ? bounce1 : bounce0;
          float& jBiasAcc = c ? jb1 : jb0;
          float& jnAccR = c ? jn1 : jn0;
          V2 vb1 = v2(A.vbx - r1.y * A.wb, A.vby + r1.x * A.wb);
          V2 vb2 = v2(B.vbx - r2.y * B.wb, B.vby + r2.x * B.wb);
          V2 vr = v2((B.vx - r2.y * B.w) - (A.vx - r1.y * A.w), (B.vy + r2.x * B.w) - (A.vy + r1.x * A.w));
          float vbn = dot(vb2 - vb1, n), vrn = dot(vr, n);
          float jbn = (bias - vbn) * nMass, jbnOld = jBiasAcc;
          jBiasAcc = fmaxf(jbnOld + jbn, 0.f);
          float jn = -(bounce + vrn) * nMass, jnOld = jnAccR;
          jnAccR = fmaxf(jnOld + jn, 0.f);
          V2 jb = n * (jBiasAcc - jbnOld);
          A.vbx -= jb.x * ami; A.vby -= jb.y * ami; A.wb -= aii * cross(r1, jb);
          B.vbx += jb.x * bmi; B.vby += jb.y * bmi; B.wb += bii * cross(r2, jb);
          V2 j = n * (jnAccR - jnOld);
          A.vx -= j.x * ami; A.vy -= j.y * ami; A.w -= aii * cross(r1, j);
          B.vx += j.x * bmi; B.vy += j.y * bmi; B.w += bii * cross(r2, j);
        }
      }
      store_body(E, a, A);
      store_body(E, b, B);
    }
  } else {
    /* shared bodies: the reference's Gauss-Seidel order on lane 0 */
    if (act) {
      E.f[X_ANX + q] = n.x; E.f[X_ANY + q] = n.y;
      const int c0 = 2 * q, c1 = 2 * q + 1;
      E.f[X_CR1X + c0] = r1a.x; E.f[X_CR1Y + c0] = r1a.y; E.f[X_CR2X + c0] = r2a.x; E.f[X_CR2Y + c0] = r2a.y;
      E.f[X_CR1X + c1] = r1b.x; E.f[X_CR1Y + c1] = r1b.y; E.f[X_CR2X + c1] = r2b.x; E.f[X_CR2Y + c1] = r2b.y;
      E.f[X_CNMASS + c0] = nMass0; E.f[X_CNMASS + c1] = nMass1;
      E.f[X_CBIAS + c0] = bias0; E.f[X_CBIAS + c1] = bias1;
      E.f[X_CBOUNCE + c0] = bounce0; E.f[X_CBOUNCE + c1] = bounce1;
      E.f[X_CJBIAS + c0] = 0.f; E.f[X_CJBIAS + c1] = 0.f;
      E.f[X_CJN + c0] = jn0; E.f[X_CJN + c1] = jn1;
    }
    __syncwarp();
    warm_start_contacts(E, dt_coef);
    if (lane == 0)
      for (int it = 0; it < 10; it++) contacts_iteration(E);
    __syncwarp();
    if (act) { jn0 = E.f[X_CJN + 2 * q]; jn1 = E.f[X_CJN + 2 * q + 1]; }
  }
  if (act) { /* accumulated impulses back into the cached pairs */
    E.f[DRV_S_ARBJN0 + k] = jn0;
    E.f[DRV_S_ARBJN1 + k] = m.cnt > 1 ? jn1 : 0.f;
  }
  __syncwarp();
}

#ifdef DRV_DBG_COUNTERS
__device__ unsigned long long g_drv_dbg[8]; /* substeps, with a contact half, with cached pairs, environment-launches with a contact half */
#define X_DBGFLAG (ENV_SMEM_WORDS - 1)
#endif
__device__ __noinline__ void space_step(const Env E) {
  /* cpSpaceStep (Chipmunk2D 7.0.x), dt = 0.01; lanes 0..9 = cars, lanes 10..29 = pedestrians */
  const int lane = E.lane;
  const int N = E.c->P;
  const int nPed = E.i[DRV_S_NPED], nObs = E.i[DRV_S_NOBS];
  const int prevStepped = E.i[DRV_S_STEPPED];
  const int nArb = E.i[DRV_S_NARB];
  __syncwarp();
  if (lane == 0) {
    E.i[DRV_S_STEPPED] = 1;
    E.i[X_NACTIVE] = 0;
    E.i[X_TOUCHED] = 0;
  }
  if (lane < nArb) { /* first-touch pairs of the previous substep become normal */
    int key = E.i[DRV_S_ARBKEY + lane];
    if (((key >> 16) & 255) == RC_ARB_FIRST) E.i[DRV_S_ARBKEY + lane] = (key & ~0xff0000) | (RC_ARB_NORMAL << 16);
  }
  /* integrate positions (cpBodyUpdatePosition); bounding boxes of this lane's two shapes in registers:
     round 0 = shape `lane` (buildings 0..3, cars 4..13, pedestrians 0..17), round 1 = shape 32 + lane */
  const bool isCar = lane < N, isPed = lane >= 10 && lane - 10 < nPed;
  bool atRest = false; /* this lane's body did not move in this substep and has no velocity */
  if (isCar) {
    const int i = lane;
    float px = E.f[DRV_S_CPX + i], pxl = E.f[DRV_S_CPXL + i], py = E.f[DRV_S_CPY + i], pyl = E.f[DRV_S_CPYL + i];
    double an = load_ang(E, i);
    two_add(px, pxl, __fmul_rn(E.f[DRV_S_CVX + i] + E.f[DRV_S_CVBX + i], DT));
    two_add(py, pyl, __fmul_rn(E.f[DRV_S_CVY + i] + E.f[DRV_S_CVBY + i], DT));
    const float wsum = E.f[DRV_S_CW + i] + E.f[DRV_S_CWB + i];
    if (wsum != 0.f) { an += (double)wsum * 0.01; store_ang(E, i, an); refresh_trig(E, i, an); }
    atRest = E.f[DRV_S_CVX + i] == 0.f && E.f[DRV_S_CVY + i] == 0.f && E.f[DRV_S_CW + i] == 0.f && E.f[DRV_S_CVBX + i] == 0.f &&
             E.f[DRV_S_CVBY + i] == 0.f && E.f[DRV_S_CWB + i] == 0.f;
    E.f[DRV_S_CPX + i] = px; E.f[DRV_S_CPXL + i] = pxl; E.f[DRV_S_CPY + i] = py; E.f[DRV_S_CPYL + i] = pyl;
    E.f[DRV_S_CVBX + i] = 0.f; E.f[DRV_S_CVBY + i] = 0.f; E.f[DRV_S_CWB + i] = 0.f;
    const float sn = E.f[X_CS + i], cs = E.f[X_CC + i];
    const int t = E.i[DRV_S_CTYPE + i];
    const float h = car_length(t), w = car_width(t);
    const float bex = fabsf(h * cs) + fabsf(w * sn), bey = fabsf(h * sn) + fabsf(w * cs);
    E.f[X_CEX + i] = bex;
    E.f[X_CEY + i] = bey;
    *reinterpret_cast<float4*>(&E.f[X_CBOX + 4 * i]) = make_float4(px, py, bex, bey);
  } else if (isPed) {
    const int i = lane - 10;
    float px = E.f[DRV_S_PPX + i], pxl = E.f[DRV_S_PPXL + i], py = E.f[DRV_S_PPY + i], pyl = E.f[DRV_S_PPYL + i];
    two_add(px, pxl, __fmul_rn(E.f[DRV_S_PVX + i] + E.f[DRV_S_PVBX + i], DT));
    two_add(py, pyl, __fmul_rn(E.f[DRV_S_PVY + i] + E.f[DRV_S_PVBY + i], DT));
    atRest = E.f[DRV_S_PVX + i] == 0.f && E.f[DRV_S_PVY + i] == 0.f && E.f[DRV_S_PVBX + i] == 0.f && E.f[DRV_S_PVBY + i] == 0.f;
    E.f[DRV_S_PPX + i] = px; E.f[DRV_S_PPXL + i] = pxl; E.f[DRV_S_PPY + i] = py; E.f[DRV_S_PPYL + i] = pyl;
    E.f[DRV_S_PVBX + i] = 0.f; E.f[DRV_S_PVBY + i] = 0.f;
  }
  __syncwarp();
  /* broad phase: every shape against the N cars (Chipmunk's bounding-box test; the margin covers the float32
     rounding of the absolute coordinates, the narrow phase decides exactly) */
  {
    /* Bit per dynamic body (cars 0..9, pedestrians 10..29).  A pair of bodies at rest is exactly where it was in the
       previous substep and its contacts carry no impulse; unless an impulse can reach it through a chain of overlapping
       pairs from a body that moves (or from a cached pair that still stores an impulse, is ageing or has just begun),
       colliding and solving it again changes nothing.  `dirty` is the closure of the moving bodies under "overlaps or
       is cached with"; only pairs with a dirty body go through the collision half of the substep, the cached pairs of
       clean bodies are kept as they are. */
#ifdef DRV_NO_QUIET_SKIP /* dev check (tools/gpu_drv_quiet_check.py): collide and solve every pair, every substep */
    unsigned dirty = __ballot_sync(FULLMASK, isCar || isPed);
#else
    /* (the first substep after a reset or a state injection collides everything: overlapping bodies at rest have no
       cached pair yet) */
    unsigned dirty = __ballot_sync(FULLMASK, (isCar || isPed) && (!atRest || !prevStepped));
#endif
    unsigned pairBodies = 0; /* lane < nArb: the dynamic bodies of cached pair `lane` */
    {
      unsigned seed = 0;
      if (lane < nArb) {
        const int key = E.i[DRV_S_ARBKEY + lane];
        const int a = key & 255, b = (key >> 8) & 255;
        if (a >= DRV_SHAPE_CAR0 && a < DRV_SHAPE_OBS0) pairBodies |= 1u << (a - DRV_SHAPE_CAR0);
        if (b >= DRV_SHAPE_CAR0 && b < DRV_SHAPE_OBS0) pairBodies |= 1u << (b - DRV_SHAPE_CAR0);
        const int st = (key >> 16) & 0xffff; /* state | age << 8 */
        const bool quiet = (st == RC_ARB_NORMAL || st == RC_ARB_IGNORE) && E.f[DRV_S_ARBJN0 + lane] == 0.f && E.f[DRV_S_ARBJN1 + lane] == 0.f;
        if (!quiet) seed = pairBodies;
      }
      if (nArb) dirty |= __reduce_or_sync(FULLMASK, seed);
    }
    int mask[2];
    unsigned own[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int s = lane + 32 * r;
      float cx = 0.f, cy = 0.f, ex = -1e9f, ey = -1e9f; /* negative extents: no overlap */
      int selfCar = 99;
      if (s < DRV_NBUILD) { V2 p = build_pos(s); cx = p.x; cy = p.y; ex = BUILD_W; ey = BUILD_H; }
      else if (s < DRV_SHAPE_PED0) {
        int j = s - DRV_SHAPE_CAR0;
        if (j < N) { cx = E.f[DRV_S_CPX + j]; cy = E.f[DRV_S_CPY + j]; ex = E.f[X_CEX + j]; ey = E.f[X_CEY + j]; selfCar = j; }
      } else if (s < DRV_SHAPE_OBS0) {
        int j = s - DRV_SHAPE_PED0;
        if (j < nPed) { cx = E.f[DRV_S_PPX + j]; cy = E.f[DRV_S_PPY + j]; ex = PED_R; ey = PED_R; }
      } else if (s < DRV_NSHAPE) {
        int j = s - DRV_SHAPE_OBS0;
        if (j < nObs) { cx = E.f[DRV_S_OPX + j]; cy = E.f[DRV_S_OPY + j]; ex = OBST_HALF; ey = OBST_HALF; }
      }
      int m = 0;
#pragma unroll
      for (int i = 0; i < DRV_MAXCARS; i++) {
        if (i < N) { /* warp-uniform */
          const float4 b = *reinterpret_cast<const float4*>(&E.f[X_CBOX + 4 * i]);
          const bool ov = fabsf(b.x - cx) <= b.z + ex + 0.01f && fabsf(b.y - cy) <= b.w + ey + 0.01f;
          if (ov && i < selfCar) m |= 1 << i; /* selfCar = 99 for everything that is not a car */
        }
      }
      mask[r] = m;
      own[r] = (s >= DRV_SHAPE_CAR0 && s < DRV_SHAPE_OBS0) ? 1u << (s - DRV_SHAPE_CAR0) : 0u; /* cars 4..13 -> 0..9, pedestrians 14..33 -> 10..29 */
    }
    const bool anyPair = __any_sync(FULLMASK, (mask[0] | mask[1]) != 0) || nArb;
    if (anyPair && dirty) {
      for (;;) { /* closure: a few rounds at most (chains of touching bodies are short) */
        unsigned add = 0;
#pragma unroll
        for (int r = 0; r < 2; r++)
          if (mask[r] && ((own[r] & dirty) || ((unsigned)mask[r] & dirty))) add |= own[r] | (unsigned)mask[r];
        if (pairBodies & dirty) add |= pairBodies;
        const unsigned nd = dirty | __reduce_or_sync(FULLMASK, add);
        if (nd == dirty) break;
        dirty = nd;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (!(own[r] & dirty)) mask[r] &= (int)(dirty & 0x3ffu); /* pairs without a dirty body drop out */
    const bool haveCand = __any_sync(FULLMASK, (mask[0] | mask[1]) != 0);
    const unsigned keepMask = __ballot_sync(FULLMASK, lane < nArb && !(pairBodies & dirty)); /* cached pairs left alone */
    const bool dirtyCached = __any_sync(FULLMASK, lane < nArb && (pairBodies & dirty) != 0);
#ifdef DRV_DBG_COUNTERS
    if (lane == 0) { atomicAdd(&g_drv_dbg[0], 1ull); if (haveCand || dirtyCached) atomicAdd(&g_drv_dbg[1], 1ull); if (nArb) atomicAdd(&g_drv_dbg[2], 1ull);
      if ((haveCand || dirtyCached) && !E.i[X_DBGFLAG]) { E.i[X_DBGFLAG] = 1; atomicAdd(&g_drv_dbg[3], 1ull); } }
#endif
    if (haveCand || dirtyCached) {
      if (haveCand) {
        E.i[X_BPMASK + lane] = mask[0];
        E.i[X_BPMASK + 32 + lane] = mask[1];
        __syncwarp();
      }
      contacts_substep(E, prevStepped, haveCand, isCar, isPed, keepMask);
    } else friction_phase(E, isCar, isPed);
  }
}

__device__ __forceinline__ void phys_substeps(const Env E, const float* actions_env, int sub_b, int sub_e, bool first) {
  const int lane = E.lane;
  if (first) { /* DrivingEnvironment.py:251-256 */
    if (lane < DRV_MAXCARS) { E.f[X_REW + lane] = 0.f; E.f[X_POSREW + lane] = 0.f; }
    if (lane == 0) { E.f[X_TEAMREW] = 0.f; E.i[X_DIAG] = 0; }
  }
  if (lane < DRV_MAXCARS) { /* Car.py:9-13: mass by type; moment_for_poly of the box */
    const int t = E.i[DRV_S_CTYPE + lane];
    E.f[X_CMASS + lane] = car_mass(t);
    E.f[X_CMINV + lane] = 1.f / car_mass(t);
    E.f[X_CIINV + lane] = car_iinv(t);
    refresh_trig(E, lane, load_ang(E, lane));
  }
  __syncwarp();
#ifdef DRV_DBG_COUNTERS
  if (lane == 0) E.i[X_DBGFLAG] = 0;
  __syncwarp();
#endif
  const int N = E.c->P;
#pragma unroll 1
  for (int sub = sub_b; sub < sub_e; sub++) {
    cars_phase(E, sub, actions_env);
    peds_phase(E);
    __syncwarp();
    space_step(E);
    /* :280-287 */
    bool fin = lane < N && (E.i[DRV_S_CFLAGS + lane] & (DRV_CF_FINISHED | DRV_CF_CRASHED)) == DRV_CF_FINISHED;
    unsigned all = __ballot_sync(FULLMASK, fin);
    if (lane == 0) {
      int el = E.i[DRV_S_ELAPSED] + 1;
      E.i[DRV_S_ELAPSED] = el;
      if (!E.i[DRV_S_ALLFIN] && all == ((1u << N) - 1)) {
        E.i[DRV_S_ALLFIN] = 1;
        E.f[X_TEAMREW] += (float)(MAXTIME - el) / 100.f;
      }
    }
    __syncwarp();
  }
}

/* ================================================================= observations */
__device__ __forceinline__ float normalize_(float pt, float nf, float mean) { return (pt * nf - mean) * 2.f; } /* cutils.py:318-323 */
__device__ __forceinline__ float wrap_pi(float a) {
  if (a > PI_F) a -= 2.f * PI_F;
  if (a < -PI_F) a += 2.f * PI_F;
  return a;
}

/* atan2 to ~2e-7 rad in ~25 instructions (atan2f costs ~75): octant reduction, then Cephes' atanf scheme -- a second
   reduction to |t| <= tan(pi/8) and a 4-term odd polynomial */
__device__ __forceinline__ float fast_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = mx > 0.f ? __fdividef(mn, mx) : 0.f;
  const bool big = a > 0.41421356237309503f;
  const float t = big ? __fdividef(a - 1.f, a + 1.f) : a;
  const float z = t * t;
  float r = fmaf(fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f), z * t, t);
  if (big) r += 0.78539816339744831f;
  if (ay > ax) r = 1.57079632679489662f - r;
  if (x < 0.f) r = 3.14159265358979324f - r;
  return copysignf(r, y);
}

/* cutils.doesInteractPoly (cutils.py:643-696, canOcclude = True) of a point-like sighting at p1 (angle ang1 in the
   observer frame) against occluder o, whose view-block data the observer's pass has put into shared memory */
__device__ __forceinline__ int interact_poly(const Env E, float p1x, float p1y, float ang1, int o, float radius) {
  int ret = I_NO;
  if (radius > 0.f) {
    float dx = E.f[V_PX + o] - p1x, dy = E.f[V_PY + o] - p1y;
    if (__fmaf_rn(dx, dx, __fmul_rn(dy, dy)) < radius) ret = I_NEARBY;
  }
  float pAngle = ang1 - E.f[V_ANG2 + o];
  if (pAngle > PI_F) pAngle -= 2.f * PI_F;
  else if (pAngle < -PI_F) pAngle += 2.f * PI_F;
  if (pAngle > E.f[V_AMN + o] && pAngle < E.f[V_AMX + o]) {
    V2 p1 = v2(p1x, p1y);
    V2 q1 = v2(E.f[V_Q1X + o], E.f[V_Q1Y + o]), q2 = v2(E.f[V_Q2X + o], E.f[V_Q2Y + o]);
    if (E.i[V_OK + o] & 2) {
      if (cross(q2 - q1, p1 - q1) < 0.f) ret = I_OCCLUDE;
    } else {
      V2 qm = v2(E.f[V_QMX + o], E.f[V_QMY + o]);
      if (cross(q2 - qm, p1 - qm) < 0.f && cross(qm - q1, p1 - q1) < 0.f) ret = I_OCCLUDE;
    }
  }
  return ret;
}

/* getViewBlockAngle part of doesInteractPoly for an occluder with centre p (observer frame) and axis-aligned
   corners tr + (+-hx, +-hy) (NOT rotated: the reference discards its rotated() results) */
__device__ __forceinline__ void occluder_store(const Env E, int o, V2 p, V2 tr, float hx, float hy, float ang2, bool ok) {
  const float sx[4] = {1.f, -1.f, -1.f, 1.f}, sy[4] = {1.f, 1.f, -1.f, -1.f};
  float ang[4], dist[4];
  V2 cr[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    cr[k] = v2(tr.x + sx[k] * hx, tr.y + sy[k] * hy);
    ang[k] = wrap_pi(fast_atan2(cr[k].y, cr[k].x) - ang2);
    dist[k] = lensq(cr[k]);
  }
  int mn = 0, mx = 0, cl = 0;
#pragma unroll
  for (int k = 1; k < 4; k++) {
    if (ang[k] < ang[mn]) mn = k;
    if (ang[k] > ang[mx]) mx = k;
    if (dist[k] < dist[cl]) cl = k;
  }
  float amn = ang[0], amx = ang[0];
  V2 q1 = cr[0], q2 = cr[0], qm = cr[0];
#pragma unroll
  for (int k = 1; k < 4; k++) {
    if (mn == k) { amn = ang[k]; q1 = cr[k]; }
    if (mx == k) { amx = ang[k]; q2 = cr[k]; }
    if (cl == k) qm = cr[k];
  }
  E.f[V_ANG2 + o] = ang2; E.f[V_AMN + o] = amn; E.f[V_AMX + o] = amx;
  E.f[V_Q1X + o] = q1.x; E.f[V_Q1Y + o] = q1.y; E.f[V_Q2X + o] = q2.x; E.f[V_Q2Y + o] = q2.y;
  E.f[V_QMX + o] = qm.x; E.f[V_QMY + o] = qm.y;
  E.f[V_PX + o] = p.x; E.f[V_PY + o] = p.y;
  E.i[V_OK + o] = (ok ? 1 : 0) | ((cl == mn || cl == mx) ? 2 : 0);
}

struct Det { /* one sighting (cutils.isSeenInRadius result) held by a lane */
  int seen;
  float px, py, c, s, w, h, fin;
};

/* cutils.addNoiseRect (cutils.py:479-542); the corner updates are not restated: nothing reads them afterwards */
__device__ __forceinline__ void noise_rect(const Env E, const Rand6& r, Det& d, int interaction, float maxDist, bool misClass) {
  if (!d.seen) return;
  const float magn = E.c->noise_mag, rand_ = E.c->rand_base;
  float nx = (r.u(0) - 0.5f) * magn, ny = (r.u(1) - 0.5f) * magn;
  float angleDiff;
  if (E.c->noise_type == DYNENV_NOISE_RANDOM) {
    if (r.u(2) < rand_) { d.seen = S_NO; return; }
    angleDiff = (r.u(4) - 0.5f) * magn * ANGLE_NOISE;
  } else {
    float range = 0.25f + 3.75f * sqrtf(d.px * d.px + d.py * d.py) / maxDist;
    float mult = range;
    if (interaction == I_NEARBY) mult = range * 2.f;
    if (d.seen == S_DISTANT) mult = range * 3.f;
    else if (d.seen == S_PARTIAL) mult = range * 4.f;
    nx *= mult; ny *= mult;
    if (r.u(2) < rand_ * mult) { d.seen = S_NO; return; }
    if (misClass && r.u(3) < rand_ * mult / 2.f) d.seen = S_MISCLASS;
    angleDiff = (r.u(4) - 0.5f) * magn * ANGLE_NOISE * 0.25f;
  }
  float sd, cd;
  sincos_b(angleDiff, &sd, &cd);
  float c = d.c * cd - d.s * sd, s = d.s * cd + d.c * sd; /* cos / sin of (atan2(s, c) + angleDiff) */
  d.c = c; d.s = s;
  d.px += nx; d.py += ny;
}

/* rows of one list group into the staged record: lanes with `valid` append their row after `count` rows
   (list order = lane order), rows beyond the capacity are dropped; returns the new row count */
__device__ __forceinline__ int emit_rows(const Env E, int lane, bool valid, int off, int stride, int cap, int count,
                                         const float* row, int nf) {
  unsigned bal = __ballot_sync(FULLMASK, valid);
  int idx = count + __popc(bal & ((1u << lane) - 1));
  if (valid && idx < cap)
    for (int k = 0; k < nf; k++) E.f[V_STAGE + off + idx * stride + k] = row[k];
  return count + __popc(bal);
}

/* getAgentVision of car `ag` (DrivingEnvironment.py:750-977) into the staged record.
   Lane roles -- slot A: car `lane` (lanes 0..9) or obstacle `lane - 10` (lanes 10..29);
   slot B: pedestrian `lane` (0..19), building `lane - 20` (20..23), the observer itself (24), lane rows (25..30);
   false positives: loop index `lane` (0..9). */
__device__ __noinline__ void agent_vision(const Env E, int ag, uint32_t tick) {
  const int lane = E.lane;
  const int N = E.c->P, nPed = E.i[DRV_S_NPED], nObs = E.i[DRV_S_NOBS];
  const bool noisy = E.c->noise_mag != 0.f;
  const float ox = E.f[DRV_S_CPX + ag], oxl = E.f[DRV_S_CPXL + ag], oy = E.f[DRV_S_CPY + ag], oyl = E.f[DRV_S_CPYL + ag];
  const double oa = load_ang(E, ag);
  float sa, ca;
  sincos_d(oa, &sa, &ca);
  for (int w = lane; w < DRV_OBS_WORDS; w += 32) E.f[V_STAGE + w] = 0.f;
  /* ---- visibility (cutils.isSeenInRadius) and view-block data */
  Det A, B;
  A.seen = S_NO; B.seen = S_NO;
  A.px = A.py = A.c = A.s = A.w = A.h = A.fin = 0.f;
  B = A;
  float angA = 0.f, angB = 0.f;
  bool listA = false, listB = false;
  int nFree; /* slot-A lanes without a car / obstacle: the first four of them take the buildings' view-block data */
  {
    bool have = false;
    float x = 0, xl = 0, y = 0, yl = 0, hx = OBST_HALF, hy = OBST_HALF;
    if (lane < 10) {
      if (lane < N && lane != ag) {
        have = true;
        x = E.f[DRV_S_CPX + lane]; xl = E.f[DRV_S_CPXL + lane]; y = E.f[DRV_S_CPY + lane]; yl = E.f[DRV_S_CPYL + lane];
        int t = E.i[DRV_S_CTYPE + lane];
        hx = car_length(t); hy = car_width(t);
        A.fin = (E.i[DRV_S_CFLAGS + lane] & DRV_CF_FINISHED) ? 1.f : 0.f;
      }
    } else if (lane < 30) {
      if (lane - 10 < nObs) { have = true; x = E.f[DRV_S_OPX + lane - 10]; y = E.f[DRV_S_OPY + lane - 10]; }
    }
    const unsigned freeMask = __ballot_sync(FULLMASK, !have);
    nFree = __popc(freeMask);
    const int bld = have ? -1 : __popc(freeMask & ((1u << lane) - 1));
    const bool isBuild = !have && bld < DRV_NBUILD;
    if (isBuild) { V2 bp = build_pos(bld); x = bp.x; y = bp.y; hx = BUILD_W; hy = BUILD_H; }
    if (have || isBuild) {
      V2 tr = v2(__fadd_rn(__fsub_rn(x, ox), __fsub_rn(xl, oxl)), __fadd_rn(__fsub_rn(y, oy), __fsub_rn(yl, oyl)));
      float dist = lensq(tr);
      const bool seen = isBuild || dist <= MAXVIS0; /* buildings: isSeenInRadius with maxDist 2e7 */
      V2 p = v2(tr.x * ca + tr.y * sa, -tr.x * sa + tr.y * ca);
      const float ang = fast_atan2(p.y, p.x);
      if (have && seen) {
        A.seen = dist <= MAXVIS1 ? S_NORMAL : S_DISTANT;
        A.px = p.x; A.py = p.y;
        if (lane < 10) sincos_d(__dsub_rn(load_ang(E, lane), oa), &A.s, &A.c);
        else { A.c = ca; A.s = -sa; }
        A.w = hy; A.h = hx;
        angA = ang;
        listA = true;
      }
      occluder_store(E, isBuild ? 30 + bld : lane, p, tr, hx, hy, ang, seen);
    }
    if (!have && lane < 30) E.i[V_OK + lane] = 0;
  }
  if (lane < 20) {
    if (lane < nPed) {
      V2 tr = v2(__fadd_rn(__fsub_rn(E.f[DRV_S_PPX + lane], ox), __fsub_rn(E.f[DRV_S_PPXL + lane], oxl)),
                 __fadd_rn(__fsub_rn(E.f[DRV_S_PPY + lane], oy), __fsub_rn(E.f[DRV_S_PPYL + lane], oyl)));
      float dist = lensq(tr);
      if (dist <= MAXVIS0) {
        B.seen = dist <= MAXVIS1 ? S_NORMAL : S_DISTANT;
        B.px = tr.x * ca + tr.y * sa; B.py = -tr.x * sa + tr.y * ca;
        B.c = ca; B.s = -sa;
        angB = fast_atan2(B.py, B.px);
        listB = true;
      }
    }
  }
  if (nFree < DRV_NBUILD && lane >= 20 + nFree && lane < 24) { /* 10 cars and 20 obstacles: the remaining buildings */
    V2 bp = build_pos(lane - 20);
    V2 tr = v2(__fsub_rn(__fsub_rn(bp.x, ox), oxl), __fsub_rn(__fsub_rn(bp.y, oy), oyl));
    V2 p = v2(tr.x * ca + tr.y * sa, -tr.x * sa + tr.y * ca);
    occluder_store(E, 30 + lane - 20, p, tr, BUILD_W, BUILD_H, fast_atan2(p.y, p.x), true);
  }
  __syncwarp();
  /* ---- occlusion by buildings (:785-792): occluded sightings leave their lists */
  if (listA) {
    int m = 0;
    for (int b = 0; b < 4; b++) m = max(m, interact_poly(E, A.px, A.py, angA, 30 + b, 0.f));
    if (m == I_OCCLUDE) { listA = false; A.seen = S_NO; }
  }
  if (listB) {
    int m = 0;
    for (int b = 0; b < 4; b++) m = max(m, interact_poly(E, B.px, B.py, angB, 30 + b, 0.f));
    if (m == I_OCCLUDE) { listB = false; B.seen = S_NO; }
  }
  __syncwarp();
  if (lane < 30 && !listA) E.i[V_OK + lane] &= ~1;
  __syncwarp();
  /* ---- pedestrians hidden by / near cars and obstacles (:795-805); `max` of two lists is lexicographic */
  int pedInter = I_NO;
  {
    int carPed = 0, obsPed = 0;
    /* the cars / obstacles still in their lists: a warp-uniform set, visited bit by bit */
    const unsigned inList = __ballot_sync(FULLMASK, lane < 30 && (lane < 10 ? lane < N : lane < 10 + nObs) && (E.i[V_OK + lane] & 1));
    if (listB) {
      for (unsigned m = inList & 0x3FFu; m; m &= m - 1)
        carPed = max(carPed, interact_poly(E, B.px, B.py, angB, __ffs(m) - 1, 400.f));
      for (unsigned m = inList >> 10; m; m &= m - 1)
        obsPed = max(obsPed, interact_poly(E, B.px, B.py, angB, 10 + __ffs(m) - 1, 400.f));
    }
    unsigned differ = __ballot_sync(FULLMASK, listB && carPed != obsPed);
    unsigned obsWins = __ballot_sync(FULLMASK, listB && obsPed > carPed);
    bool useObs = differ && ((obsWins >> (__ffs(differ) - 1)) & 1);
    pedInter = useObs ? obsPed : carPed;
    if (listB && pedInter == I_OCCLUDE) B.seen = S_NO;
  }
  /* ---- lane rows (Road.getCarLaneDistances, Road.py:36-71) and the observer itself, on the slot-B lanes 24..30 */
  float laneDist = 0.f, laneType = 0.f;
  if (lane == 24) {
    B.seen = S_NORMAL; B.px = ox + oxl; B.py = oy + oyl; B.c = ca; B.s = sa;
  } else if (lane >= 25 && lane < 31) {
    const int road = lane < 29 ? 0 : 1, k = lane < 29 ? lane - 25 : lane - 29, nL = road_lanes(road);
    float dist = (road == 0 ? __fsub_rn(__fsub_rn(875.f, ox), oxl) : __fadd_rn(__fsub_rn(oy, 500.f), oyl)) / ROAD_W;
    if (!(fabsf(dist) > 10.f)) {
      float c = ca, s = -sa; /* cos / sin of (road.angle - carAngle), road 1: angle 0 */
      if (road == 0) sincos_d(__dsub_rn(PIO2_D, oa), &s, &c);
      float distMult = 1.f, flip = 1.f;
      if (c >= 0.f) { flip = -1.f; c = -c; s = -s; distMult = -1.f; }
      B.seen = S_NORMAL;
      laneDist = ((dist + 0.5f) + (float)(k - nL)) * ROAD_W * 0.1f * distMult;
      B.c = c; B.s = s;
      laneType = (k < nL ? 1.f : -1.f) * flip;
    }
  }
  /* ---- noise (:808-818) */
  if (noisy) {
    if (__any_sync(FULLMASK, listA)) {
      Rand6 r;
      if (listA) r = rng_block(E, tick, lane < 10 ? DRV_SITE_OBS_CAR : DRV_SITE_OBS_OBST, ag, lane < 10 ? lane : lane - 10, 0);
      if (listA) noise_rect(E, r, A, I_NO, MAXVIS1, true);
    }
    {
      const bool isLaneRow = lane >= 25 && lane < 31 && B.seen;
      const bool any = listB || lane == 24 || isLaneRow;
      Rand6 r;
      if (any) {
        int site = lane < 20 ? DRV_SITE_OBS_PED : (lane == 24 ? DRV_SITE_OBS_SELF : DRV_SITE_OBS_LANE);
        int b = lane < 20 ? lane : (lane == 24 ? 0 : lane - 25);
        r = rng_block(E, tick, site, ag, b, 0);
      }
      if (listB) noise_rect(E, r, B, pedInter, MAXVIS0, false);
      else if (lane == 24) noise_rect(E, r, B, I_NO, MAXVIS1, false);
      else if (isLaneRow) { /* cutils.addNoiseLane (cutils.py:380-413) */
        const float magn = E.c->noise_mag, rand_ = E.c->rand_base;
        float distNoise = (r.u(0) - 0.5f) * magn, angleDiff = (r.u(1) - 0.5f) * magn, da;
        if (E.c->noise_type == DYNENV_NOISE_RANDOM) {
          if (r.u(2) < rand_) B.seen = S_NO;
          laneDist *= distNoise;
          da = ANGLE_NOISE * angleDiff;
        } else {
          float m1 = 0.25f + 3.75f * laneDist * laneDist / MAXVIS1;
          if (r.u(2) < rand_ * m1) B.seen = S_NO;
          laneDist += distNoise * m1;
          da = ANGLE_NOISE * m1 / 5.f * angleDiff;
        }
        float sd, cd;
        sincos_b(da, &sd, &cd);
        float c = B.c * cd - B.s * sd, s = B.s * cd + B.c * sd;
        B.c = c; B.s = s;
      }
    }
  }
  /* ---- random false positives (:828-874): loop index = lane */
  int fpClass = -1;
  Det F;
  F.seen = S_NORMAL; F.px = F.py = F.c = F.s = F.w = F.h = F.fin = 0.f;
  float fpLaneDist = 0.f, fpLaneType = 0.f;
  if (noisy && lane < 10) {
    Rand6 r0 = rng_block(E, tick, DRV_SITE_OBS_FP, ag, lane, 0);
    if (r0.u(0) < E.c->rand_base) {
      fpClass = randint_from(r0.u(1), 0, 5);
      float d = r0.u(2) * MAXVIS1, a1 = r0.u(3) * 2.f * PI_F, angle = r0.u(4) * 2.f * PI_F;
      float s1, c1;
      sincos_b(a1, &s1, &c1);
      F.px = d * c1; F.py = d * s1;
      Rand6 r1 = rng_block(E, tick, DRV_SITE_OBS_FP, ag, lane, 1);
      if (fpClass <= 1) {
        F.w = r0.u(5) * 5.f + 5.f;
        F.h = r1.u(0) * 10.f + 5.f;
        sincos_b(angle, &F.s, &F.c);
      } else if (fpClass == 3) {
        float a = (r0.u(5) - 0.5f) * PI_F * 2.f;
        fpLaneDist = floorf(r1.u(0) * W_ / 2.f); /* `random() * W // 2` */
        sincos_b(a, &F.s, &F.c);
        fpLaneType = (float)randint_from(r1.u(1), -1, 1);
      }
    }
  }
  /* ---- lists in the reference's order, final filters and encoding (:821-826, :877-977) */
  float row[7];
  /* cars: real sightings, obstacles misclassified as cars, false positives */
  const unsigned balListCars = __ballot_sync(FULLMASK, lane < 10 && listA);
  const unsigned balMisObst = __ballot_sync(FULLMASK, lane >= 10 && lane < 30 && listA && A.seen == S_MISCLASS);
  const unsigned balFpCar = __ballot_sync(FULLMASK, fpClass == 0);
  row[0] = normalize_(A.px, NORMX, 0.f); row[1] = normalize_(A.py, NORMY, 0.f); row[2] = A.c; row[3] = A.s;
  row[4] = normalize_(A.w, NORMW, 0.5f); row[5] = normalize_(A.h, NORMH, 0.5f); row[6] = lane < 10 ? A.fin : 0.f;
  int cnt = 0;
  cnt = emit_rows(E, lane, lane < 10 && listA && A.seen != S_NO && A.seen != S_MISCLASS, DRV_O_CAR, DRV_F_CAR, DRV_CAP_CAR, cnt, row, 7);
  cnt = emit_rows(E, lane, lane >= 10 && lane < 30 && listA && A.seen == S_MISCLASS, DRV_O_CAR, DRV_F_CAR, DRV_CAP_CAR, cnt, row, 7);
  float frow[7];
  frow[0] = normalize_(F.px, NORMX, 0.f); frow[1] = normalize_(F.py, NORMY, 0.f); frow[2] = F.c; frow[3] = F.s;
  frow[4] = normalize_(F.w, NORMW, 0.5f); frow[5] = normalize_(F.h, NORMH, 0.5f); frow[6] = 0.f;
  cnt = emit_rows(E, lane, fpClass == 0, DRV_O_CAR, DRV_F_CAR, DRV_CAP_CAR, cnt, frow, 7);
  const int nCarRows = min(cnt, DRV_CAP_CAR);
  /* obstacles: real sightings, cars misclassified as obstacles, false positives */
  cnt = 0;
  cnt = emit_rows(E, lane, lane >= 10 && lane < 30 && listA && A.seen != S_NO && A.seen != S_MISCLASS, DRV_O_OBS, DRV_F_OBS, DRV_CAP_OBS, cnt, row, 6);
  cnt = emit_rows(E, lane, lane < 10 && listA && A.seen == S_MISCLASS, DRV_O_OBS, DRV_F_OBS, DRV_CAP_OBS, cnt, row, 6);
  cnt = emit_rows(E, lane, fpClass == 1, DRV_O_OBS, DRV_F_OBS, DRV_CAP_OBS, cnt, frow, 6);
  const int nObsRows = min(cnt, DRV_CAP_OBS);
  /* pedestrians: real sightings, false positives, false positives near cars (REALISTIC, :877-882) */
  cnt = 0;
  row[0] = normalize_(B.px, NORMX, 0.f); row[1] = normalize_(B.py, NORMY, 0.f);
  cnt = emit_rows(E, lane, lane < 20 && listB && B.seen != S_NO, DRV_O_PED, DRV_F_PED, DRV_CAP_PED, cnt, row, 2);
  cnt = emit_rows(E, lane, fpClass == 2, DRV_O_PED, DRV_F_PED, DRV_CAP_PED, cnt, frow, 2);
  if (noisy && E.c->noise_type == DYNENV_NOISE_REALISTIC) {
    /* rows of the car list in list order: group 1 = lanes 0..9, group 2 = lanes 10..29, group 3 = false positives */
    const int n1 = __popc(balListCars), n2 = __popc(balMisObst);
    const unsigned lt = (1u << lane) - 1;
#pragma unroll 1
    for (int g = 0; g < 3; g++) {
      bool cand;
      int idx;
      float cx, cy;
      if (g == 0) { cand = lane < 10 && listA && A.seen == S_NORMAL; idx = __popc(balListCars & lt); cx = A.px; cy = A.py; }
      else if (g == 1) { cand = lane >= 10 && lane < 30 && listA && A.seen == S_MISCLASS; idx = n1 + __popc(balMisObst & lt); cx = A.px; cy = A.py; }
      else { cand = fpClass == 0; idx = n1 + n2 + __popc(balFpCar & lt); cx = F.px; cy = F.py; }
      bool hit = false;
      if (__any_sync(FULLMASK, cand)) {
        if (cand) {
          Rand6 r = rng_block(E, tick, DRV_SITE_OBS_FPPED, ag, idx, 0);
          if (r.u(0) < E.c->rand_base * 10.f && sqrtf(cx * cx + cy * cy) < 250.f) {
            hit = true;
            row[0] = normalize_(cx + (2.f * r.u(1) - 1.f) * 10.f, NORMX, 0.f);
            row[1] = normalize_(cy + (2.f * r.u(2) - 1.f) * 10.f, NORMY, 0.f);
          }
        }
      }
      cnt = emit_rows(E, lane, hit, DRV_O_PED, DRV_F_PED, DRV_CAP_PED, cnt, row, 2);
    }
  }
  const int nPedRows = min(cnt, DRV_CAP_PED);
  /* lanes */
  cnt = 0;
  row[0] = laneDist; row[1] = B.c; row[2] = B.s; row[3] = laneType;
  cnt = emit_rows(E, lane, lane >= 25 && lane < 31 && B.seen != S_NO, DRV_O_LANE, DRV_F_LANE_FULL, DRV_CAP_LANE, cnt, row, 4);
  frow[0] = fpLaneDist; frow[1] = F.c; frow[2] = F.s; frow[3] = fpLaneType;
  cnt = emit_rows(E, lane, fpClass == 3, DRV_O_LANE, DRV_F_LANE_FULL, DRV_CAP_LANE, cnt, frow, 4);
  const int nLaneRows = min(cnt, DRV_CAP_LANE);
  if (lane == 24) {
    const int t = E.i[DRV_S_CTYPE + ag];
    float* o = E.f + (V_STAGE + DRV_O_SELF);
    o[0] = normalize_(B.px, NORMX, 5.f); o[1] = normalize_(B.py, NORMY, 5.f); o[2] = B.c; o[3] = B.s;
    o[4] = normalize_(car_width(t), NORMW, 0.5f); o[5] = normalize_(car_length(t), NORMH, 0.5f);
    o[6] = normalize_(E.f[DRV_S_CGOALX + ag], NORMX, 5.f); o[7] = normalize_(E.f[DRV_S_CGOALY + ag], NORMY, 5.f);
    o[8] = (E.i[DRV_S_CFLAGS + ag] & DRV_CF_FINISHED) ? 1.f : 0.f;
    int* ci = &E.i[V_STAGE + DRV_O_COUNTS];
    ci[0] = nCarRows; ci[1] = nObsRows; ci[2] = nPedRows; ci[3] = nLaneRows;
  }
  __syncwarp();
}

/* getFullState rows (DrivingEnvironment.py:686-747) */
__device__ __forceinline__ void car_row(const Env E, int c, float* r) {
  const int t = E.i[DRV_S_CTYPE + c];
  float sn, cs;
  sincos_d(load_ang(E, c), &sn, &cs);
  r[0] = normalize_(E.f[DRV_S_CPX + c] + E.f[DRV_S_CPXL + c], SNORMX, 0.f);
  r[1] = normalize_(E.f[DRV_S_CPY + c] + E.f[DRV_S_CPYL + c], SNORMY, 0.f);
  r[2] = cs; r[3] = sn;
  r[4] = normalize_(car_width(t), SNORMW, 0.5f); r[5] = normalize_(car_length(t), SNORMH, 0.5f);
  r[6] = (E.i[DRV_S_CFLAGS + c] & DRV_CF_FINISHED) ? 1.f : 0.f;
}

/* get_full_obs entry of car `ag` (:121-124) into the staged record */
__device__ __noinline__ void full_obs_agent(const Env E, int ag) {
  const int lane = E.lane;
  const int N = E.c->P, nPed = E.i[DRV_S_NPED], nObs = E.i[DRV_S_NOBS];
  for (int w = lane; w < DRV_OBS_WORDS; w += 32) E.f[V_STAGE + w] = 0.f;
  __syncwarp();
  float r7[7];
  if (lane < N) {
    car_row(E, lane, r7);
    if (lane == ag) {
      float* o = E.f + (V_STAGE + DRV_O_SELF);
      for (int k = 0; k < 6; k++) o[k] = r7[k];
      o[6] = normalize_(E.f[DRV_S_CGOALX + ag], SNORMX, 0.f); o[7] = normalize_(E.f[DRV_S_CGOALY + ag], SNORMY, 0.f);
      o[8] = r7[6];
    } else {
      float* o = E.f + (V_STAGE + DRV_O_CAR + (lane < ag ? lane : lane - 1) * DRV_F_CAR);
      for (int k = 0; k < 7; k++) o[k] = r7[k];
    }
  }
  if (lane < nObs) {
    float* o = E.f + (V_STAGE + DRV_O_OBS + lane * DRV_F_OBS);
    o[0] = normalize_(E.f[DRV_S_OPX + lane], SNORMX, 0.f); o[1] = normalize_(E.f[DRV_S_OPY + lane], SNORMY, 0.f);
    o[2] = normalize_(OBST_HALF, SNORMW, 0.5f); o[3] = normalize_(OBST_HALF, SNORMH, 0.5f);
  }
  if (lane < nPed) {
    float* o = E.f + (V_STAGE + DRV_O_PED + lane * DRV_F_PED);
    o[0] = normalize_(E.f[DRV_S_PPX + lane] + E.f[DRV_S_PPXL + lane], SNORMX, 0.f);
    o[1] = normalize_(E.f[DRV_S_PPY + lane] + E.f[DRV_S_PPYL + lane], SNORMY, 0.f);
  }
  if (lane < 8) {
    /* `l.Lanes[i - l.nLanes]` indexes the 2*nLanes+1 lane lines with a NEGATIVE Python index: row i holds the
       geometry of lane offset (i - nLanes) mod (2*nLanes+1) - nLanes under the type tag of i (:690-696) */
    const int road = lane < 5 ? 0 : 1, nL = road_lanes(road), i = (lane < 5 ? lane : lane - 5) - nL;
    const int cntL = 2 * nL + 1;
    const int t = i - nL; /* in [-2 nL, 0]: Python's modulo is one wrap */
    const int off = (t < 0 ? t + cntL : t) - nL;
    V2 nrm = road_normal(road), p0 = road_p0(road), p1 = road_p1(road);
    V2 l0 = p0 + nrm * ((float)off * ROAD_W), l1 = p1 + nrm * ((float)off * ROAD_W);
    float* o = E.f + (V_STAGE + DRV_O_LANE + lane * DRV_F_LANE_FULL);
    o[0] = normalize_(l0.x, SNORMX, 0.f); o[1] = normalize_(l0.y, SNORMY, 0.f);
    o[2] = normalize_(l1.x, SNORMX, 0.f); o[3] = normalize_(l1.y, SNORMY, 0.f);
    o[4] = (i == nL || i == -nL) ? 1.f : (i == 0 ? -1.f : 0.f);
  }
  if (lane == 0) {
    int* ci = &E.i[V_STAGE + DRV_O_COUNTS];
    ci[0] = N - 1; ci[1] = nObs; ci[2] = nPed; ci[3] = 8;
  }
  __syncwarp();
}

/* one observation of every car (`tick` addresses the noise draws) to obs_env[N][DRV_OBS_WORDS] */
__device__ __noinline__ void observe(const Env E, uint32_t tick, float* obs_env) {
  const int lane = E.lane;
  const int N = E.c->P;
#pragma unroll 1
  for (int ag = 0; ag < N; ag++) {
    if (E.c->obs_type == DYNENV_OBS_FULL) full_obs_agent(E, ag);
    else agent_vision(E, ag, tick);
    if (obs_env) {
      float4* dst = reinterpret_cast<float4*>(obs_env + (size_t)ag * DRV_OBS_WORDS);
      const float4* src = reinterpret_cast<const float4*>(E.f + V_STAGE);
      for (int w = lane; w < DRV_OBS_WORDS / 4; w += 32) dst[w] = src[w];
    }
    __syncwarp();
  }
}

__device__ __noinline__ void fill_info(const Env E, float* info_env) {
  /* info['Full State'] = getFullState() (:686-718) */
  if (!info_env) return;
  const int lane = E.lane;
  const int N = E.c->P, nPed = E.i[DRV_S_NPED], nObs = E.i[DRV_S_NOBS];
  for (int w = DRV_I_STATE_CARS + lane; w < DRV_INFO_WORDS; w += 32) info_env[w] = 0.f;
  __syncwarp();
  if (lane < N) {
    float r7[7];
    car_row(E, lane, r7);
    for (int k = 0; k < 7; k++) info_env[DRV_I_STATE_CARS + lane * 7 + k] = r7[k];
    info_env[DRV_I_GOALS + 2 * lane] = normalize_(E.f[DRV_S_CGOALX + lane], SNORMX, 0.f);
    info_env[DRV_I_GOALS + 2 * lane + 1] = normalize_(E.f[DRV_S_CGOALY + lane], SNORMY, 0.f);
  }
  if (lane < nObs) {
    float* o = info_env + DRV_I_STATE_OBS + lane * 4;
    o[0] = normalize_(E.f[DRV_S_OPX + lane], SNORMX, 0.f); o[1] = normalize_(E.f[DRV_S_OPY + lane], SNORMY, 0.f);
    o[2] = normalize_(OBST_HALF, SNORMW, 0.5f); o[3] = normalize_(OBST_HALF, SNORMH, 0.5f);
  }
  if (lane < nPed) {
    float* o = info_env + DRV_I_STATE_PED + lane * 2;
    o[0] = normalize_(E.f[DRV_S_PPX + lane] + E.f[DRV_S_PPXL + lane], SNORMX, 0.f);
    o[1] = normalize_(E.f[DRV_S_PPY + lane] + E.f[DRV_S_PPYL + lane], SNORMY, 0.f);
  }
  if (lane == 0) {
    reinterpret_cast<int*>(info_env)[DRV_I_NOBS] = nObs;
    reinterpret_cast<int*>(info_env)[DRV_I_NPED] = nPed;
  }
}

/* rewards, episode accumulators, info, done (DrivingEnvironment.py:299-322) */
__device__ __noinline__ void finish_step(const Env E, float* info_env) {
  const int lane = E.lane;
  const int N = E.c->P;
  const bool done = E.i[DRV_S_ELAPSED] >= MAXTIME;
  float rew = 0.f, epr = 0.f, eppr = 0.f;
  if (lane < N) {
    const float team = E.f[X_TEAMREW];
    rew = E.f[X_REW + lane] + team;
    epr = E.f[DRV_S_EPR + lane] + rew;
    eppr = E.f[DRV_S_EPPR + lane] + E.f[X_POSREW + lane] + fmaxf(0.f, team);
    E.f[DRV_S_EPR + lane] = epr;
    E.f[DRV_S_EPPR + lane] = eppr;
  }
  const int fl = lane < N ? E.i[DRV_S_CFLAGS + lane] : 0;
  const unsigned okMask = __ballot_sync(FULLMASK, (fl & (DRV_CF_FINISHED | DRV_CF_CRASHED)) == DRV_CF_FINISHED);
  const unsigned crMask = __ballot_sync(FULLMASK, (fl & DRV_CF_CRASHED) != 0);
  if (info_env) {
    int* ii = reinterpret_cast<int*>(info_env);
    if (lane < DRV_MAXCARS) {
      info_env[DRV_I_REWARDS + lane] = rew;
      info_env[DRV_I_EPR + lane] = done ? epr : 0.f;
      info_env[DRV_I_EPPR + lane] = done ? eppr : 0.f;
    }
    if (lane == 0) {
      ii[DRV_I_DONE] = done ? 1 : 0;
      ii[DRV_I_EPG] = done ? __popc(okMask) : 0;
      ii[DRV_I_EPG + 1] = done ? __popc(crMask) : 0;
      ii[DRV_I_FLAGS] = E.i[X_DIAG];
    }
  }
  fill_info(E, info_env);
  __syncwarp();
}

/* ================================================================= reset (scene set-up) */
__device__ __forceinline__ float reset_u(const Env E, int k) {
  Rand6 r;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&E.i[R_BLOCKS + 4 * (k / 6)]);
  r.w[0] = w[0]; r.w[1] = w[1]; r.w[2] = w[2]; r.w[3] = w[3];
  return r.u(k % 6);
}
__device__ __forceinline__ V2 walk_spot(int road, int side, float length, float width) {
  /* Road.getWalkSpot (Road.py:117-125) */
  V2 nrm = road_normal(road), p0 = road_p0(road), p1 = road_p1(road);
  float off = (float)(road_lanes(road) + 1) * ROAD_W * (side ? -1.f : 1.f);
  V2 w0 = p0 + nrm * off, w1 = p1 + nrm * off;
  V2 c = w0 + (w1 - w0) * length;
  return c + nrm * (width * ROAD_W * (side ? 1.f : -1.f));
}

/* DrivingEnvironment.__init__ / _setup_scene (:20-115, 527-584): a new episode's scene from the RESET stream */
__device__ __noinline__ void init_scene(const Env E) {
  const int lane = E.lane;
  const int N = E.c->P;
  const int episode = E.i[DRV_S_EPISODE] + 1;
  __syncwarp();
  for (int w = lane; w < DRV_STATE_WORDS; w += 32) E.i[w] = 0;
  __syncwarp();
  if (lane == 0) E.i[DRV_S_EPISODE] = episode;
  __syncwarp();
  for (int b = lane; b < NRESET_BLOCKS; b += 32) {
    Rand6 r = rng_block(E, DYNENV_TICK_RESET, DRV_SITE_RESET, 0, 0, b);
    E.i[R_BLOCKS + 4 * b] = (int)r.w[0]; E.i[R_BLOCKS + 4 * b + 1] = (int)r.w[1];
    E.i[R_BLOCKS + 4 * b + 2] = (int)r.w[2]; E.i[R_BLOCKS + 4 * b + 3] = (int)r.w[3];
  }
  __syncwarp();
  if (lane == 0) {
    int k = 0;
    int roadSel[DRV_MAXCARS], endSel[DRV_MAXCARS], teams[DRV_MAXCARS], types[DRV_MAXCARS];
    for (int i = 0; i < N; i++) roadSel[i] = (int)(reset_u(E, k++) * 2.f);
    for (int i = 0; i < N; i++) endSel[i] = (int)(reset_u(E, k++) * 2.f);
    for (int i = 0; i < N; i++) teams[i] = (int)(reset_u(E, k++) * 3.f);
    for (int i = 0; i < N; i++) types[i] = (int)(reset_u(E, k++) * 4.f);
    int perm[30];
    for (int i = 0; i < 30; i++) perm[i] = i;
    for (int i = 29; i > 0; i--) {
      int j = (int)(reset_u(E, k++) * (float)(i + 1));
      int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
    for (int i = 0; i < N; i++) {
      int spotID = perm[i];
      const int roadID = spotID < 20 ? 0 : 1;
      spotID -= roadID ? 20 : 0;
      const int laneID = spotID / 5, spot = spotID % 5;
      /* Road.getSpot (Road.py:100-114) */
      const int nL = road_lanes(roadID);
      const int end = laneID >= nL ? 1 : 0;
      V2 dir = road_dir(roadID), nrm = road_normal(roadID);
      V2 pos = end ? road_p1(roadID) : road_p0(roadID);
      V2 spotDir = (end ? v2(-dir.x, -dir.y) : dir) * 90.f;
      V2 laneDir = (end ? nrm : v2(-nrm.x, -nrm.y)) * ROAD_W;
      float ln = (float)(end ? laneID - nL : laneID) + 0.5f;
      V2 p = pos + laneDir * ln + spotDir * (float)spot;
      /* atan2(spotDir): (0, 90) -> pi/2, (-0, -90) -> -pi/2, (90, 0) -> 0, (-90, -0) -> -pi */
      const double ang = roadID == 0 ? (end ? -PIO2_D : PIO2_D) : (end ? -2 * PIO2_D : 0.0);
      E.i[DRV_S_CTYPE + i] = types[i];
      E.i[DRV_S_CTEAM + i] = teams[i];
      V2 goal = endSel[i] ? road_p1(roadSel[i]) : road_p0(roadSel[i]);
      E.f[DRV_S_CGOALX + i] = goal.x; E.f[DRV_S_CGOALY + i] = goal.y;
      E.f[DRV_S_CPX + i] = p.x; E.f[DRV_S_CPY + i] = p.y;
      store_ang(E, i, ang);
      float sn, cs;
      sincos_d(ang, &sn, &cs);
      E.f[DRV_S_CDIRX + i] = cs; E.f[DRV_S_CDIRY + i] = sn;
      E.f[DRV_S_CPREVX + i] = p.x; E.f[DRV_S_CPREVY + i] = p.y;
      E.i[DRV_S_CLANE + i] = 4;
    }
    const int nPed = 10 + (int)(reset_u(E, k++) * 11.f);
    E.i[DRV_S_NPED] = nPed;
    {
      const int kr = k, ks = k + nPed, kl = k + 2 * nPed, kw = k + 3 * nPed;
      k += 4 * nPed;
      for (int i = 0; i < nPed; i++) {
        int road = (int)(reset_u(E, kr + i) * 2.f), side = (int)(reset_u(E, ks + i) * 2.f);
        V2 c = walk_spot(road, side, reset_u(E, kl + i), reset_u(E, kw + i) / 2.f + 0.25f);
        E.f[DRV_S_PPX + i] = c.x; E.f[DRV_S_PPY + i] = c.y;
        E.i[DRV_S_PFLAGS + i] = (side ? DRV_PF_SIDE : 0) | (road ? DRV_PF_ROAD : 0);
        E.i[DRV_S_PSPEED + i] = 3 + (int)(reset_u(E, k++) * 4.f);
      }
    }
    {
      const int n = 10 + (int)(reset_u(E, k++) * 11.f);
      const int kr = k, ks = k + n, kl = k + 2 * n, kw = k + 3 * n;
      int nObs = 0;
      for (int i = 0; i < n; i++) {
        int road = (int)(reset_u(E, kr + i) * 2.f), side = (int)(reset_u(E, ks + i) * 2.f);
        V2 c = walk_spot(road, side, reset_u(E, kl + i), reset_u(E, kw + i) / 2.f + 0.25f);
        if (off_road(c.x, 0.f, c.y, 0.f)) { E.f[DRV_S_OPX + nObs] = c.x; E.f[DRV_S_OPY + nObs] = c.y; nObs++; }
      }
      E.i[DRV_S_NOBS] = nObs;
    }
  }
  __syncwarp();
}

/* reset + the initial observation (environment_base.py:205-224: nTimeSteps = 1 copy) */
__device__ __noinline__ void reset_env(const Env E, float* obs_env, float* info_env, bool write_info) {
  init_scene(E);
  observe(E, DYNENV_TICK_RESETOBS, obs_env);
  if (write_info && info_env) {
    const int lane = E.lane;
    for (int w = lane; w < DRV_I_STATE_CARS; w += 32) info_env[w] = 0.f;
    fill_info(E, info_env);
  }
  __syncwarp();
}

}  // namespace drv
}  // namespace dynenv
