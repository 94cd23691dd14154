/*
 * RoboCup rollout hot path, one warp per environment (see dynenv_common.cuh).
 *
 * Replaces, for a batch of environments (reference paths under DynEnv/):
 *   RoboCupEnvironment.step            RoboCupEnvironment.py:446-524
 *   processAction / Robot.step,turn,kick,turnHead   :527-581, Robot.py:103-138
 *   tick                               :862-1007
 *   isBallOutOfField / ballFreeKickProcess          :622-732, :600-619
 *   fall / penalize / getFreePenaltySpot            :735-859
 *   collision callbacks                :1010-1146
 *   pymunk Space.step (Chipmunk2D 7.0 semantics) called at :482,
 *     with cutils.friction_robot / friction_ball    cutils.py:93-140
 *   getAgentVision (PARTIAL)           :1192-1316,1536-1561  + cutils.py:352-468,546-565,699-821
 *   getFullState                       :1149-1189
 *   processSeens                       :1563-1575
 *   scene set-up / reset               :42-63,239-336, environment_base.py:205-224
 *
 * All arithmetic is float32 (the reference computes in float64 and casts its observations
 * to float32).  Shared-memory words [0, RC_STATE_WORDS) are the persistent state record of
 * include/dynenv_b200.h; scratch follows.
 */
#pragma once
#include "dynenv_common.cuh"

namespace dynenv {
namespace rc {

/* ---- field constants (RoboCupEnvironment.py:26-29,73-93; Robot.py:13-20) */
constexpr float PI_F = 3.14159265358979323846f;
constexpr float W_ = 1040.f, H_ = 740.f, SIDE = 70.f, LINEW = 5.f;
constexpr float PEN_RADIUS = 5.f, PEN_LENGTH = 60.f, PEN_WIDTH = 110.f, PEN_DIST = 130.f;
constexpr float CC_RADIUS = 75.f, GOAL_W = 80.f, POST_RADIUS = 5.f, BALL_RADIUS = 5.f;
constexpr float FIELD_W = 900.f, FIELD_H = 600.f;
constexpr float ROB_LEN = 10.f, ROB_RAD = 7.5f, ROB_TOTAL = 17.5f, ROB_FOV = PI_F / 4;
constexpr float ROB_VEL = 50.f, ROB_ANGVEL = 20.f, ROB_MASS = 4000.f;
constexpr float HEAD_MAX = 2 * PI_F / 3;
constexpr int MAXTIME = 12000;
constexpr int BALL = 20, POST0 = 21, NSHAPE = 25;
constexpr float FOOT_MINV = 1.f / 4000.f;
constexpr float FOOT_IINV = 1.f / (4000.f * ((35.f * 35.f + 4.f * 7.5f * 7.5f) / 12.f + 100.f)); /* Robot.py:34 */
constexpr float BALL_MINV = 0.1f, BALL_IINV = 1.f / 500.f;                                       /* Ball.py:8-10 */
constexpr float DT = 0.01f;
constexpr float STDNORM2 = 2.f * (2.f / 1040.f); /* normalize(): pt * standardNorm * 2 */
constexpr float STDNORM = 2.f / 1040.f;

enum { S_NO = 0, S_PARTIAL = 1, S_DISTANT = 2, S_NORMAL = 3, S_MISCLASS = 4 };
enum { I_NO = 0, I_NEARBY = 1, I_OCCLUDE = 2 };

constexpr int NSHAPE_ = 25, RC_MAXCAND = 64;
/* ---- scratch layout (words after the state record) */
enum {
  /* accumulators of the step in flight live in the record (they cross kernel launches) */
  X_REW = RC_S_STEP_REW, X_POSREW = RC_S_STEP_POSREW, X_TEAMREW = RC_S_STEP_TEAMREW, X_DIAG = RC_S_STEP_DIAG,
  X_LSUM = RC_S_STEP_LSUM, X_BSUM = RC_S_STEP_BSUM, X_RSEEN = RC_S_STEP_RSEEN,
  X_CURREW = RC_STATE_WORDS,   /* F[2] currReward of isBallOutOfField */
  X_OLDBX = X_CURREW + 2,      /* F ball position before the rules moved it */
  X_OLDBY = X_OLDBX + 1,
  X_TOUCHED = X_OLDBY + 1,     /* I bit k: cached pair k collided this substep */
  X_NACTIVE = X_TOUCHED + 1,   /* I */
  X_ACTIVE = X_NACTIVE + 1,    /* I[RC_MAXACTIVE] index of the cached pair */
  X_ENVID = X_ACTIVE + RC_MAXACTIVE, /* I global environment id (RNG counter word 0) */
  X_UNION = (X_ENVID + 1 + 3) / 4 * 4,
  /* physics view of the union: active pairs and their contacts */
  X_ANX = X_UNION,             /* F[16] normal */
  X_ANY = X_ANX + RC_MAXACTIVE,
  X_AE = X_ANY + RC_MAXACTIVE, /* F restitution */
  X_AU = X_AE + RC_MAXACTIVE,  /* F friction    */
  X_CR1X = X_AU + RC_MAXACTIVE, /* per contact [2*RC_MAXACTIVE] */
  X_CR1Y = X_CR1X + 2 * RC_MAXACTIVE,
  X_CR2X = X_CR1Y + 2 * RC_MAXACTIVE,
  X_CR2Y = X_CR2X + 2 * RC_MAXACTIVE,
  X_CNMASS = X_CR2Y + 2 * RC_MAXACTIVE,
  X_CTMASS = X_CNMASS + 2 * RC_MAXACTIVE,
  X_CBIAS = X_CTMASS + 2 * RC_MAXACTIVE,
  X_CBOUNCE = X_CBIAS + 2 * RC_MAXACTIVE,
  X_CJBIAS = X_CBOUNCE + 2 * RC_MAXACTIVE,
  X_CJN = X_CJBIAS + 2 * RC_MAXACTIVE,
  X_CJT = X_CJN + 2 * RC_MAXACTIVE,
  X_PHYS_END = X_CJT + 2 * RC_MAXACTIVE,
  /* broad phase: bounding circles of the shapes and of the groups (robot = both feet), candidate pairs */
  X_BCX = X_PHYS_END,          /* F[25] */
  X_BCY = X_BCX + NSHAPE_,     /* F[25] */
  X_GX = X_BCY + NSHAPE_,      /* F[16]: 15 groups + a far-away dummy for unused pair slots */
  X_GY = X_GX + 16,
  X_GR = X_GY + 16,
  X_NCAND = X_GR + 16,         /* I */
  X_CAND = X_NCAND + 1,        /* U16[RC_MAXCAND] i * 32 + j, two per word */
  X_BP_END = X_CAND + RC_MAXCAND / 2,
  /* vision view of the union */
  V_ROB = X_UNION,             /* F[9][3] other robots' sightings: x, y, has */
  V_BALL = V_ROB + 27,         /* F[3] */
  V_POST = V_BALL + 3,         /* F[4][3] */
  V_RAW = V_POST + 12,         /* F[RC_CAP_ROBOT][3] rows of the robot list: x, y, seen == Normal */
  V_CNT = V_RAW + 3 * RC_CAP_ROBOT, /* I[6] list lengths */
  V_STAGE = (V_CNT + 6 + 3) / 4 * 4, /* [RC_OBS_WORDS] the observation record being built */
  V_END = V_STAGE + RC_OBS_WORDS,
  /* CFG_WORDS + ENV_SMEM_WORDS <= 1568 words: 32 one-warp blocks per SM (228 KB / 32 - 1 KB reserved per block) */
  ENV_SMEM_WORDS = ((X_BP_END > V_END ? X_BP_END : V_END) + 3) / 4 * 4,
  /* IMAGE observations: the raw detection record (RC_R_*) is staged behind the slice; only the kernels instantiated
     for IMAGE mode allocate these words (per-warp stride ENV_SMEM_WORDS_RAW) */
  V_RAWREC = ENV_SMEM_WORDS,
  ENV_SMEM_WORDS_RAW = ENV_SMEM_WORDS + RC_RAW_WORDS
};

/* ---- static scene tables (RoboCupEnvironment.py:153-227,296-302) */
/* point landmarks seen by getAgentVision: 4 posts, 3 penalty crosses, 16 line crosses: x, y, tag1, tag2 */
__device__ __align__(16) const float g_marks[23][4] = {
    {70, 450, 1, -0.27f},   {70, 290, 1, 0.27f},    {970, 450, -1, -0.27f}, {970, 290, -1, 0.27f},
    {520, 370, 0, 0},       {200, 370, 1, 0},       {840, 370, -1, 0},
    {70, 70, 1, 1},         {70, 670, 1, -1},       {970, 70, -1, 1},       {970, 670, -1, -1},
    {520, 70, 0, 1},        {520, 670, 0, -1},      {520, 220, 0, 0.5f},    {520, 520, 0, -0.5f},
    {70, 260, 1, 0.37f},    {70, 480, 1, -0.37f},   {130, 260, 0.87f, 0.37f}, {130, 480, 0.87f, -0.37f},
    {970, 260, -1, 0.37f},  {970, 480, -1, -0.37f}, {910, 260, -0.87f, 0.37f}, {910, 480, -0.87f, -0.37f}};
/* 11 field lines: x1, y1, x2, y2, tagx, tagy */
__device__ const float g_lines[11][6] = {
    {70, 70, 70, 670, 1, 0},        {970, 70, 970, 670, -1, 0},      {70, 70, 970, 70, 0, 1},
    {70, 670, 970, 670, 0, -1},     {520, 70, 520, 670, 0, 0},       {70, 260, 130, 260, 1, 0.37f},
    {70, 480, 130, 480, 1, -0.37f}, {130, 260, 130, 480, 0.87f, 0},  {910, 260, 970, 260, -1, 0.37f},
    {910, 480, 970, 480, -1, -0.37f}, {910, 260, 910, 480, -0.87f, 0}};

/* broad phase over groups: robots 0..9, ball 10, posts 11..14; all pairs (a < b) with a dynamic (99) */
__device__ const unsigned char g_pairs[99][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {0, 7}, {0, 8}, {0, 9}, {0, 10}, {0, 11}, {0, 12}, {0, 13}, {0, 14}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {1, 7}, {1, 8}, {1, 9}, {1, 10}, {1, 11}, {1, 12}, {1, 13}, {1, 14}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {2, 9}, {2, 10}, {2, 11}, {2, 12}, {2, 13}, {2, 14}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {3, 9}, {3, 10}, {3, 11}, {3, 12}, {3, 13}, {3, 14}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {4, 9}, {4, 10}, {4, 11}, {4, 12}, {4, 13}, {4, 14}, {5, 6}, {5, 7}, {5, 8}, {5, 9}, {5, 10}, {5, 11}, {5, 12}, {5, 13}, {5, 14}, {6, 7}, {6, 8}, {6, 9}, {6, 10}, {6, 11}, {6, 12}, {6, 13}, {6, 14}, {7, 8}, {7, 9}, {7, 10}, {7, 11}, {7, 12}, {7, 13}, {7, 14}, {8, 9}, {8, 10}, {8, 11}, {8, 12}, {8, 13}, {8, 14}, {9, 10}, {9, 11}, {9, 12}, {9, 13}, {9, 14}, {10, 11}, {10, 12}, {10, 13}, {10, 14}};

struct Env {
  SharedF f;
  SharedI i;
  SharedCfg c;
  int lane;
};

/* one Philox block of this environment's stream (include/dynenv_rng.h) */
__device__ __forceinline__ Rand6 rng_block(const Env E, uint32_t tick, int site, int a, int b, int blk);

__device__ __forceinline__ bool is_foot(int s) { return s < 20; }
__device__ __forceinline__ float shape_e(int s) { return s < 20 ? 0.3f : (s == BALL ? 0.98f : 0.95f); }
__device__ __forceinline__ float shape_u(int s) { return s < 20 ? 2.5f : (s == BALL ? 3.0f : 0.f); }
__device__ __forceinline__ float body_minv(int s) { return s < 20 ? FOOT_MINV : (s == BALL ? BALL_MINV : 0.f); }
__device__ __forceinline__ float body_iinv(int s) { return s < 20 ? FOOT_IINV : (s == BALL ? BALL_IINV : 0.f); }
__device__ __forceinline__ float body_m(int s) { return s < 20 ? ROB_MASS : 10.f; }
__device__ __forceinline__ V2 post_pos(int k) { return v2(k < 2 ? SIDE : W_ - SIDE, (k & 1) ? H_ / 2 - GOAL_W : H_ / 2 + GOAL_W); }
__device__ __forceinline__ int team_of(const Env E, int r) { return r < E.c->n_players ? 1 : -1; }
__device__ __forceinline__ bool shape_active(const Env E, int s) { return s >= 20 || s < 2 * E.c->P; }

/* position of body b relative to body a, from both words: exact where it matters (close bodies) */
__device__ __forceinline__ V2 pos_diff(const Env E, int b, int a) {
  float bx, by, bxl = 0.f, byl = 0.f, ax, ay, axl = 0.f, ayl = 0.f;
  if (b < RC_NBODY) { bx = E.f[RC_S_PX + b]; by = E.f[RC_S_PY + b]; bxl = E.f[RC_S_PXL + b]; byl = E.f[RC_S_PYL + b]; }
  else { V2 p = post_pos(b - POST0); bx = p.x; by = p.y; }
  if (a < RC_NBODY) { ax = E.f[RC_S_PX + a]; ay = E.f[RC_S_PY + a]; axl = E.f[RC_S_PXL + a]; ayl = E.f[RC_S_PYL + a]; }
  else { V2 p = post_pos(a - POST0); ax = p.x; ay = p.y; }
  return v2(__fadd_rn(__fsub_rn(bx, ax), __fsub_rn(bxl, axl)), __fadd_rn(__fsub_rn(by, ay), __fsub_rn(byl, ayl)));
}
__device__ __forceinline__ void set_pose(const Env E, int s, float x, float y, float ang) {
  E.f[RC_S_PX + s] = x; E.f[RC_S_PY + s] = y; E.f[RC_S_ANG + s] = ang;
  E.f[RC_S_PXL + s] = 0.f; E.f[RC_S_PYL + s] = 0.f; E.f[RC_S_ANGL + s] = 0.f;
}
__device__ __forceinline__ V2 robot_pos(const Env E, int r) { /* Robot.getPos (Robot.py:90-91) */
  return v2((E.f[RC_S_PX + 2 * r] + E.f[RC_S_PX + 2 * r + 1]) * 0.5f,
            (E.f[RC_S_PY + 2 * r] + E.f[RC_S_PY + 2 * r + 1]) * 0.5f);
}
__device__ __forceinline__ float robot_angle(const Env E, int r) { /* Robot.getAngle (Robot.py:94-100) */
  return (E.f[RC_S_ANG + 2 * r] + E.f[RC_S_ANG + 2 * r + 1]) * 0.5f;
}
__device__ __forceinline__ V2 ball_pos(const Env E) { return v2(E.f[RC_S_PX + BALL], E.f[RC_S_PY + BALL]); }

__device__ __forceinline__ Rand6 rng_block(const Env E, uint32_t tick, int site, int a, int b, int blk) {
  Rand6 r;
  uint4 w = philox4x32_10(E.c->seed_lo, E.c->seed_hi, (uint32_t)E.i[X_ENVID], (uint32_t)E.i[RC_S_EPISODE], tick,
                          ((uint32_t)site << 24) | ((uint32_t)(a & 255) << 16) | ((uint32_t)(b & 255) << 8) | (uint32_t)blk);
  r.w[0] = w.x; r.w[1] = w.y; r.w[2] = w.z; r.w[3] = w.w;
  return r;
}
__device__ __forceinline__ float rnd(const Env E, uint32_t tick, int site, int a, int b, int k) {
  return rng_block(E, tick, site, a, b, k / 6).u(k % 6);
}

/* world geometry of a shape from its body (cpShapeCacheBB): feet are segments (-10,+-10)-(10,+-10) */
__device__ __forceinline__ void shape_world(int s, float px, float py, float ang, V2* ta, V2* tb) {
  if (is_foot(s)) {
    float ly = (s & 1) ? -ROB_LEN : ROB_LEN; /* Robot.py:24-27 */
    float sn, cs;
    sincos_b(ang, &sn, &cs);
    *ta = v2(px - ROB_LEN * cs - ly * sn, py - ROB_LEN * sn + ly * cs);
    *tb = v2(px + ROB_LEN * cs - ly * sn, py + ROB_LEN * sn + ly * cs);
  } else {
    *ta = v2(px, py);
    *tb = *ta;
  }
}
__device__ __forceinline__ void shape_cache_store(const Env E, int s) {
  V2 ta, tb;
  shape_world(s, E.f[RC_S_PX + s], E.f[RC_S_PY + s], E.f[RC_S_ANG + s], &ta, &tb);
  E.f[RC_S_CAX + s] = ta.x;
  E.f[RC_S_CAY + s] = ta.y;
  E.f[RC_S_CBX + s] = tb.x;
  E.f[RC_S_CBY + s] = tb.y;
}
__device__ __forceinline__ void shape_geom(const Env E, int s, V2* ta, V2* tb) { /* cached geometry */
  if (s < RC_NBODY) {
    *ta = v2(E.f[RC_S_CAX + s], E.f[RC_S_CAY + s]);
    *tb = v2(E.f[RC_S_CBX + s], E.f[RC_S_CBY + s]);
  } else {
    *ta = post_pos(s - POST0);
    *tb = *ta;
  }
}
/* segment normal: rperp(normalize(b - a)) */
__device__ __forceinline__ V2 seg_normal(V2 ta, V2 tb) {
  V2 d = tb - ta;
  float il = 1.f / len(d);
  return v2(d.y * il, -d.x * il);
}

__device__ __forceinline__ V2 closest_on_seg(V2 p, V2 a, V2 b) {
  V2 d = b - a;
  float t = dot(d, p - a) / lensq(d);
  t = fminf(fmaxf(t, 0.f), 1.f);
  return a + d * t;
}

/* ================================================================= game rules (serial code) */
__device__ void penalize(const Env E, int r);

__device__ float point_query_dist(const Env E, int s, V2 p) { /* cpShapePointQuery distance */
  V2 ta, tb;
  shape_geom(E, s, &ta, &tb);
  if (is_foot(s)) return len(p - closest_on_seg(p, ta, tb)) - ROB_RAD;
  return len(p - ta) - 10.f;
}

__device__ __forceinline__ void last_kicked_push(const Env E, int r) {
  int n = E.i[RC_S_NLASTK] < 4 ? E.i[RC_S_NLASTK] : 3;
  for (int k = n; k > 0; k--) E.i[RC_S_LASTK + k] = E.i[RC_S_LASTK + k - 1];
  E.i[RC_S_LASTK] = r;
  E.i[RC_S_NLASTK] = n + 1;
}
__device__ __forceinline__ bool in_last_kicked(const Env E, int r) {
  bool f = false;
  for (int k = 0; k < E.i[RC_S_NLASTK]; k++) f |= (E.i[RC_S_LASTK + k] == r);
  return f;
}

__device__ __noinline__ void fall(const Env E, int r, int punish) {
  /* RoboCupEnvironment.py:735-791 */
  V2 pos = robot_pos(E, r);
  if (punish) E.f[X_REW + r] -= 2.f;
  /* space.point_query(pos, 40, filter) on the geometry cached at the last collision phase; the
     inf/NaN forces the reference applies to static posts are never integrated: skipped */
#pragma unroll 1
  for (int s = 0; s <= BALL; s++) {
    if (!shape_active(E, s)) continue;
    if (s == 2 * r || s == 2 * r + 1) continue;
    if (!(point_query_dist(E, s, pos) < 40.f)) continue;
    float force = ROB_VEL * ROB_MASS * body_m(s) / 50.f;
    V2 bp = v2(E.f[RC_S_PX + s], E.f[RC_S_PY + s]);
    V2 dp = pos - bp;
    float l = len(dp);
    dp = v2(-dp.x * force / l, -dp.y * force / l);
    E.f[RC_S_FX + s] += dp.x;
    E.f[RC_S_FY + s] += dp.y;
    E.f[RC_S_TQ + s] += (pos.x - bp.x) * dp.y - (pos.y - bp.y) * dp.x;
    if (s == BALL) {
      if (E.i[RC_S_NLASTK] && !in_last_kicked(E, r)) last_kicked_push(E, r);
      if (E.i[RC_S_BALLOWNED] != 0) {
        E.f[RC_S_GRACE] = 0.f;
        E.f[RC_S_BALLFREE] = 0.f;
        E.i[RC_S_BALLOWNED] = 0;
      }
    }
  }
  E.i[RC_S_RFLAGS + r] |= RC_RF_FALLEN;
  E.i[RC_S_FALLCNT + r] += 1;
  E.f[RC_S_FALLT + r] = 4000.f;
  if (E.i[RC_S_FALLCNT + r] > 2) penalize(E, r);
}

__device__ __noinline__ void free_penalty_spot(const Env E, int r, V2* spot, float* angle) {
  /* RoboCupEnvironment.py:793-821; spots :139-151 */
  int team = team_of(E, r);
  float y = E.f[RC_S_PY + BALL];
  *angle = y < H_ / 2 ? -PI_F / 2 : PI_F / 2;
  float sy = y > H_ / 2 ? SIDE : H_ - SIDE;
  V2 sel = v2(0, 0);
  bool found = false;
#pragma unroll 1
  for (int k = 0; k < 7 && !found; k++) {
    float sx = team > 0 ? SIDE + (k + 1) * ROB_TOTAL * 3 : W_ - SIDE - (k + 1) * ROB_TOTAL * 3;
    V2 sp = v2(sx, sy);
    if (k == 0) sel = sp;
    bool ok = true;
#pragma unroll 1
    for (int o = 0; o < E.c->P; o++) {
      if (o == r) continue;
      if (len(sp - robot_pos(E, o)) < ROB_TOTAL * 3) { ok = false; break; }
    }
    if (ok) { sel = sp; found = true; }
  }
  *spot = sel;
}

__device__ __noinline__ void penalize(const Env E, int r) {
  /* RoboCupEnvironment.py:824-859 */
  int teamIdx = team_of(E, r) > 0 ? 0 : 1;
  int fl = E.i[RC_S_RFLAGS + r] | RC_RF_PENALIZED;
  E.f[RC_S_PENT + r] = (float)E.i[RC_S_PENALTIMES + teamIdx];
  E.f[X_REW + r] -= (float)E.i[RC_S_PENALTIMES + teamIdx] / 2000.f;
  E.i[RC_S_PENALTIMES + teamIdx] += 10000;
  V2 pos;
  float ang;
  free_penalty_spot(E, r, &pos, &ang);
  for (int k = 0; k < 2; k++) {
    int b = 2 * r + k;
    set_pose(E, b, pos.x, pos.y, ang);
    E.f[RC_S_VX + b] = 0.f; E.f[RC_S_VY + b] = 0.f; E.f[RC_S_W + b] = 0.f;
  }
  if ((fl & RC_RF_KICKING) && (fl & RC_RF_JOINT_REMOVED)) {
    fl &= ~(RC_RF_KICKING | RC_RF_JOINT_REMOVED);
    fl |= RC_RF_JOINT_ADD_PENDING; /* space.add(joint): deferred while the space is locked */
  }
  E.i[RC_S_RFLAGS + r] = fl;
}

__device__ __forceinline__ void joint_add_now(const Env E, int r) {
  int fl = E.i[RC_S_RFLAGS + r];
  if (fl & RC_RF_JOINT_ADD_PENDING) E.i[RC_S_RFLAGS + r] = fl & ~(RC_RF_JOINT_ADD_PENDING | RC_RF_JOINT_ABSENT);
}

__device__ __noinline__ void process_action(const Env E, int r, float4 act, const Rand6& ru) {
  /* RoboCupEnvironment.py:527-581, Robot.py:103-138 */
  float move = act.x, turn = act.y, kick = act.z, head = act.w;
  int team = team_of(E, r);
  if (E.c->deterministic_turn) head = -3.f * team;
  if (!E.c->continuous) head -= 3.f;
  int fl = E.i[RC_S_RFLAGS + r];
  bool canMove = !(fl & (RC_RF_PENALIZED | RC_RF_KICKING | RC_RF_FALLEN));
  int L = 2 * r;
  if (move > 0 && canMove) {
    float u = E.c->can_fall ? ru.u(0) : 0.f;
    if (u > 0.999f) { fall(E, r, 0); return; }
    int dir = (int)move - 1;
    E.f[RC_S_MOVET + r] = 500.f;
    V2 v = dir == 0 ? v2(0, 2 * ROB_VEL) : dir == 1 ? v2(0, -2 * ROB_VEL) : dir == 2 ? v2(2.5f * ROB_VEL, 0) : v2(-2 * ROB_VEL, 0);
    v = rot(v, E.f[RC_S_ANG + L]);
    E.f[RC_S_VX + L] = v.x;
    E.f[RC_S_VY + L] = v.y;
  }
  if (turn > 0 && canMove) {
    float u = E.c->can_fall ? ru.u(1) : 0.f;
    if (u > 0.999f) { fall(E, r, 0); return; }
    E.f[RC_S_MOVET + r] = 500.f;
    E.f[RC_S_W + L] += ((int)turn - 1) ? ROB_ANGVEL : -ROB_ANGVEL;
  }
  if (head != 0) { /* Robot.turnHead */
    E.f[RC_S_HEADMOV + r] = head * PI_F / 720.f;
    E.f[RC_S_MOVET + r] = 500.f;
  }
  if (kick > 0 && move == 0 && turn == 0 && canMove) {
    float u = E.c->can_fall ? ru.u(2) : 0.f;
    if (u > 0.99f) { fall(E, r, 0); return; }
    int foot = (int)kick - 1;
    int fb = 2 * r + (foot ? 1 : 0);
    E.f[RC_S_INITX + r] = E.f[RC_S_PX + fb]; E.f[RC_S_INITXL + r] = E.f[RC_S_PXL + fb];
    E.f[RC_S_INITY + r] = E.f[RC_S_PY + fb]; E.f[RC_S_INITYL + r] = E.f[RC_S_PYL + fb];
    E.i[RC_S_RFLAGS + r] = (E.i[RC_S_RFLAGS + r] & ~RC_RF_FOOT) | RC_RF_KICKING | (foot ? RC_RF_FOOT : 0);
    E.f[RC_S_MOVET + r] = 1000.f;
  }
}

__device__ __noinline__ void tick_robot(const Env E, int r) {
  /* RoboCupEnvironment.py:862-1007 */
  const float time = 10.f;
  int L = 2 * r, R = 2 * r + 1;
  int fl = E.i[RC_S_RFLAGS + r];
  float mt = E.f[RC_S_MOVET + r];
  if (mt > 0) {
    mt -= time;
    float hm = E.f[RC_S_HEADMOV + r];
    if (hm != 0) E.f[RC_S_HEAD + r] = fmaxf(-HEAD_MAX, fminf(HEAD_MAX, E.f[RC_S_HEAD + r] + hm));
    if (fl & RC_RF_KICKING) {
      int foot = (fl & RC_RF_FOOT) ? R : L;
      if (mt + time > 500 && mt <= 500) {
        if (!(fl & RC_RF_JOINT_REMOVED)) fl |= RC_RF_JOINT_REMOVED | RC_RF_JOINT_ABSENT;
        V2 v = rot(v2(ROB_VEL * 3.f, 0), E.f[RC_S_ANG + foot]);
        E.f[RC_S_VX + foot] = v.x;
        E.f[RC_S_VY + foot] = v.y;
      }
      if (mt + time > 400 && mt <= 400) {
        V2 v = rot(v2(ROB_VEL * 2.5f, 0), E.f[RC_S_ANG + foot]);
        E.f[RC_S_VX + foot] = -v.x;
        E.f[RC_S_VY + foot] = -v.y;
      } else if (mt <= 300) {
        E.f[RC_S_VX + foot] = 0.f;
        E.f[RC_S_VY + foot] = 0.f;
        fl &= ~RC_RF_KICKING;
        E.f[RC_S_PX + foot] = E.f[RC_S_INITX + r]; E.f[RC_S_PXL + foot] = E.f[RC_S_INITXL + r];
        E.f[RC_S_PY + foot] = E.f[RC_S_INITY + r]; E.f[RC_S_PYL + foot] = E.f[RC_S_INITYL + r];
        if (fl & RC_RF_JOINT_REMOVED) fl &= ~(RC_RF_JOINT_REMOVED | RC_RF_JOINT_ABSENT);
      }
    }
    if (mt <= 0) {
      mt = 0;
      E.f[RC_S_HEADMOV + r] = 0.f;
      E.f[RC_S_VX + L] = 0.f; E.f[RC_S_VY + L] = 0.f; E.f[RC_S_W + L] = 0.f;
      E.f[RC_S_VX + R] = 0.f; E.f[RC_S_VY + R] = 0.f; E.f[RC_S_W + R] = 0.f;
    }
    E.f[RC_S_MOVET + r] = mt;
  }
  E.i[RC_S_RFLAGS + r] = fl;
  if (fl & RC_RF_FALLEN) {
    float ft = E.f[RC_S_FALLT + r] - time;
    E.f[RC_S_FALLT + r] = ft;
    if (ft < 0) {
      float u = rnd(E, (uint32_t)E.i[RC_S_ELAPSED], RC_SITE_REFALL, r, 0, 0);
      if (u > 0.9f && !(fl & RC_RF_PENALIZED) && E.c->can_fall) { fall(E, r, 0); return; }
      fl &= ~RC_RF_FALLEN;
      E.i[RC_S_RFLAGS + r] = fl;
      E.i[RC_S_FALLCNT + r] = 0;
    }
  }
  if (fl & RC_RF_PENALIZED) {
    float pt = E.f[RC_S_PENT + r] - time;
    if (pt <= 0) {
      pt = 0;
      fl &= ~(RC_RF_PENALIZED | RC_RF_FALLEN);
      E.i[RC_S_RFLAGS + r] = fl;
      E.i[RC_S_FALLCNT + r] = 0;
      V2 pos;
      float ang;
      free_penalty_spot(E, r, &pos, &ang);
      set_pose(E, L, pos.x, pos.y, ang);
      set_pose(E, R, pos.x, pos.y, ang);
    }
    E.f[RC_S_PENT + r] = pt;
  } else {
    int teamIdx = team_of(E, r) > 0 ? 0 : 1;
    V2 pos = robot_pos(E, r);
    float robX = teamIdx ? W_ - pos.x : pos.x;
    const float penX = SIDE + PEN_LENGTH + LINEW / 2;
    int nd = E.i[RC_S_NDEF + teamIdx];
    int d0 = E.i[RC_S_DEFENDERS + 2 * teamIdx], d1 = E.i[RC_S_DEFENDERS + 2 * teamIdx + 1];
    bool inDef = (nd > 0 && d0 == r) || (nd > 1 && d1 == r);
    if (robX < penX && pos.y > (H_ / 2 - PEN_WIDTH) && pos.y < (H_ / 2 + PEN_WIDTH)) {
      if (!inDef) {
        if (nd >= 2) penalize(E, r);
        else {
          E.i[RC_S_DEFENDERS + 2 * teamIdx + nd] = r;
          E.i[RC_S_NDEF + teamIdx] = nd + 1;
        }
      }
    } else if (inDef) { /* list.remove: first occurrence, order kept */
      if (d0 == r) { E.i[RC_S_DEFENDERS + 2 * teamIdx] = d1; }
      E.i[RC_S_DEFENDERS + 2 * teamIdx + 1] = 0;
      if (nd == 1) E.i[RC_S_DEFENDERS + 2 * teamIdx] = 0;
      E.i[RC_S_NDEF + teamIdx] = nd - 1;
    }
  }
  V2 pos = robot_pos(E, r);
  if (pos.y < 0 || pos.x < 0 || pos.y > H_ || pos.x > W_) penalize(E, r);
  V2 prev = v2(E.f[RC_S_PREVX + r], E.f[RC_S_PREVY + r]);
  if (pos.x != prev.x || pos.y != prev.y) {
    if ((r == E.i[RC_S_CLOSEST] || r == E.i[RC_S_CLOSEST + 1]) && !(E.i[RC_S_RFLAGS + r] & RC_RF_PENALIZED)) {
      V2 bp = ball_pos(E);
      float diff = len(pos - bp) - len(prev - bp);
      E.f[X_REW + r] -= diff * 0.05f;
      E.f[X_POSREW + r] += fmaxf(0.f, -diff * 0.05f);
    }
    E.f[RC_S_PREVX + r] = pos.x;
    E.f[RC_S_PREVY + r] = pos.y;
  }
}

/* Does the per-robot phase of robot r touch anything but r's own state this substep?  Conservative:
   falls (random draws of processAction, re-fall after getting up), penalties (3rd fall, third
   defender, leaving the field), the penalty re-entry (reads the other robots) and changes of the
   defenders lists.  When no lane says yes, the ten robots tick in parallel, else in robot order. */
__device__ __noinline__ bool robot_phase_needs_order(const Env E, int r, int sub, float4 act, const Rand6& ru) {
  int fl = E.i[RC_S_RFLAGS + r];
  bool need = false;
  if (sub == 0 && E.c->can_fall && !(fl & (RC_RF_PENALIZED | RC_RF_KICKING | RC_RF_FALLEN)))
    need = (act.x > 0 && ru.u(0) > 0.999f) || (act.y > 0 && ru.u(1) > 0.999f) || (act.z > 0 && ru.u(2) > 0.99f);
  if ((fl & RC_RF_FALLEN) && E.f[RC_S_FALLT + r] - 10.f < 0) need = true;
  if (fl & RC_RF_PENALIZED) {
    if (E.f[RC_S_PENT + r] - 10.f <= 0) need = true;
  } else {
    int teamIdx = team_of(E, r) > 0 ? 0 : 1;
    int nd = E.i[RC_S_NDEF + teamIdx];
    bool inDef = (nd > 0 && E.i[RC_S_DEFENDERS + 2 * teamIdx] == r) || (nd > 1 && E.i[RC_S_DEFENDERS + 2 * teamIdx + 1] == r);
    const float penX = SIDE + PEN_LENGTH + LINEW / 2;
    /* position now, and with the kicking foot put back (the only own move of tick) */
    for (int v = 0; v < 2; v++) {
      V2 pos = robot_pos(E, r);
      if (v == 1) {
        if (!(fl & RC_RF_KICKING)) break;
        int foot = 2 * r + ((fl & RC_RF_FOOT) ? 1 : 0), other = foot ^ 1;
        pos = v2((E.f[RC_S_INITX + r] + E.f[RC_S_PX + other]) * 0.5f, (E.f[RC_S_INITY + r] + E.f[RC_S_PY + other]) * 0.5f);
      }
      float robX = teamIdx ? W_ - pos.x : pos.x;
      bool inBox = robX < penX && pos.y > (H_ / 2 - PEN_WIDTH) && pos.y < (H_ / 2 + PEN_WIDTH);
      if (inBox != inDef) need = true;
      if (pos.y < 0 || pos.x < 0 || pos.y > H_ || pos.x > W_) need = true;
    }
  }
  return need;
}

/* the range checks of processAction (RoboCupEnvironment.py:543-550) for the batched path: true = the reference would raise */
__device__ __forceinline__ bool bad_action(const Env E, float4 a) {
  bool bad = !(a.x >= 0.f && a.x <= 4.f) || !(a.y >= 0.f && a.y <= 2.f) || !(a.z >= 0.f && a.z <= 2.f);
  if (E.c->continuous) bad |= !(fabsf(a.w) <= 3.f);
  else bad |= !(a.w >= 0.f && a.w <= 6.f);
  return bad;
}

__device__ __noinline__ void robots_phase(const Env E, int sub, const float* actions) {
  int lane = E.lane;
  int P = E.c->P;
  float4 act = make_float4(0, 0, 0, 0);
  Rand6 ru;
  ru.w[0] = ru.w[1] = ru.w[2] = ru.w[3] = 0;
  bool mine = lane < P;
  if (sub == 0 && mine) {
    act = *reinterpret_cast<const float4*>(actions + 4 * lane);
    if (E.c->can_fall) ru = rng_block(E, (uint32_t)E.i[RC_S_ELAPSED], RC_SITE_ACTION, lane, 0, 0);
    if (bad_action(E, act)) atomicOr(&E.i[X_DIAG], RC_IF_BAD_ACTION);
  }
  bool need = mine && robot_phase_needs_order(E, lane, sub, act, ru);
  unsigned ordered = __ballot_sync(FULLMASK, need);
  int rounds = ordered ? P : 1;
  for (int k = 0; k < rounds; k++) {
    bool now = mine && (ordered ? lane == k : true);
    if (now) {
      if (sub == 0) {
        process_action(E, lane, act, ru);
        joint_add_now(E, lane);
      }
      tick_robot(E, lane);
      joint_add_now(E, lane);
    }
    __syncwarp();
  }
}

/* ================================================================= collision callbacks (serial) */
__device__ void cb_robot_pushing_det(const Env E, int sa, int sb) {
  /* begin, (Robot, Robot): RoboCupEnvironment.py:1010-1035 */
  int r1 = sa / 2, r2 = sb / 2;
  V2 v1 = v2(E.f[RC_S_VX + sa], E.f[RC_S_VY + sa]);
  V2 v2_ = v2(E.f[RC_S_VX + sb], E.f[RC_S_VY + sb]);
  V2 dp = robot_pos(E, r1) - robot_pos(E, r2);
  /* cos(angle(dp) - angle(v)) = dp.v / (|dp| |v|) */
  float ld = len(dp);
  float l1 = len(v1), l2 = len(v2_);
  bool p1 = l1 > 1 && (ld > 0 ? dot(dp, v1) / (ld * l1) : v1.x / l1) < -0.4f;
  bool p2 = l2 > 1 && (ld > 0 ? dot(dp, v2_) / (ld * l2) : v2_.x / l2) > 0.4f;
  int f1 = E.i[RC_S_RFLAGS + r1];
  f1 = (f1 & ~RC_RF_MIGHTPUSH) | (p1 ? RC_RF_MIGHTPUSH : 0) | RC_RF_TOUCHING;
  E.i[RC_S_RFLAGS + r1] = f1;
  int f2 = E.i[RC_S_RFLAGS + r2];
  f2 = (f2 & ~RC_RF_MIGHTPUSH) | (p2 ? RC_RF_MIGHTPUSH : 0) | RC_RF_TOUCHING;
  E.i[RC_S_RFLAGS + r2] = f2;
  E.i[RC_S_TOUCHCNT + r1] = 0;
  E.i[RC_S_TOUCHCNT + r2] = 0;
}

__device__ void cb_robot_collision(const Env E, int sa, int sb) {
  /* post_solve, (Robot, Robot): RoboCupEnvironment.py:1038-1081 */
  if (!E.c->can_fall) return;
  int i1 = sa / 2, i2 = sb / 2;
  if (i1 == i2) return;
  int* f1 = &E.i[RC_S_RFLAGS + i1];
  int* f2 = &E.i[RC_S_RFLAGS + i2];
  if (!(*f1 & (RC_RF_FALLEN | RC_RF_PENALIZED))) E.i[RC_S_TOUCHCNT + i1] += 1;
  if (!(*f2 & (RC_RF_FALLEN | RC_RF_PENALIZED))) E.i[RC_S_TOUCHCNT + i2] += 1;
  /* thresholds 0.9999^n resp. 0.99995^n as exp(n ln t) */
  const float lnNormal = -1.000050003e-04f, lnPushing = -5.000125004e-05f;
  Rand6 ru = rng_block(E, (uint32_t)E.i[RC_S_ELAPSED], RC_SITE_ROBCOL, sa, sb, 0);
  if (ru.u(0) > expf(((*f1 & RC_RF_MIGHTPUSH) ? lnPushing : lnNormal) * (float)E.i[RC_S_TOUCHCNT + i1]) &&
      !(*f1 & RC_RF_FALLEN)) {
    fall(E, i1, (*f1 & RC_RF_MIGHTPUSH) ? 1 : 0);
    E.i[RC_S_TOUCHCNT + i1] = 0;
  }
  if (ru.u(1) > expf(((*f2 & RC_RF_MIGHTPUSH) ? lnPushing : lnNormal) * (float)E.i[RC_S_TOUCHCNT + i2]) &&
      !(*f2 & RC_RF_FALLEN)) {
    fall(E, i2, (*f2 & RC_RF_MIGHTPUSH) ? 1 : 0);
    E.i[RC_S_TOUCHCNT + i2] = 0;
  }
  bool diffTeam = team_of(E, i1) != team_of(E, i2);
  if ((*f1 & RC_RF_MIGHTPUSH) && !(*f2 & RC_RF_MIGHTPUSH) && (*f2 & RC_RF_FALLEN) && diffTeam) {
    penalize(E, i1);
    E.i[RC_S_TOUCHCNT + i1] = 0;
  } else if ((*f2 & RC_RF_MIGHTPUSH) && !(*f1 & RC_RF_MIGHTPUSH) && (*f1 & RC_RF_FALLEN) && diffTeam) {
    penalize(E, i2);
    E.i[RC_S_TOUCHCNT + i2] = 0;
  }
}

__device__ void cb_separate(const Env E, int sa, int sb) {
  /* separate, (Robot, Robot) and (Robot, Goalpost): RoboCupEnvironment.py:1084-1094 */
  if (is_foot(sa)) { E.i[RC_S_RFLAGS + sa / 2] &= ~(RC_RF_TOUCHING | RC_RF_MIGHTPUSH); E.i[RC_S_TOUCHCNT + sa / 2] = 0; }
  if (is_foot(sb)) { E.i[RC_S_RFLAGS + sb / 2] &= ~(RC_RF_TOUCHING | RC_RF_MIGHTPUSH); E.i[RC_S_TOUCHCNT + sb / 2] = 0; }
}

__device__ void cb_goalpost_collision(const Env E, int foot, int post) {
  /* post_solve, (Robot, Goalpost): RoboCupEnvironment.py:1097-1123 */
  if (!E.c->can_fall) return;
  int r = foot / 2;
  int fl = E.i[RC_S_RFLAGS + r];
  if (fl & RC_RF_FALLEN) { E.i[RC_S_TOUCHCNT + r] = 0; return; }
  if (!(fl & RC_RF_TOUCHING)) { E.i[RC_S_RFLAGS + r] = fl | RC_RF_TOUCHING; E.i[RC_S_TOUCHCNT + r] = 0; }
  E.i[RC_S_TOUCHCNT + r] += 1;
  float thresh = expf(-2.000200027e-04f * (float)E.i[RC_S_TOUCHCNT + r]); /* 0.9998^n */
  float u = rnd(E, (uint32_t)E.i[RC_S_ELAPSED], RC_SITE_POSTCOL, post, foot, 0);
  if (u > thresh) fall(E, r, 1);
}

__device__ void cb_ball_collision(const Env E, int foot) {
  /* begin, (Robot, Ball): RoboCupEnvironment.py:1126-1146 */
  int r = foot / 2;
  int owned = E.i[RC_S_BALLOWNED];
  if (owned != 0) {
    if (team_of(E, r) != owned && !(E.i[RC_S_RFLAGS + r] & RC_RF_PENALIZED) && E.c->can_fall) penalize(E, r);
    else { E.i[RC_S_BALLOWNED] = 0; E.f[RC_S_GRACE] = 0.f; E.f[RC_S_BALLFREE] = 0.f; }
  }
  last_kicked_push(E, r);
}

/* ================================================================= narrow phase (serial) */
struct Manifold {
  int cnt;
  V2 n;
  V2 p1[2], p2[2];
  int hash[2];
};

__device__ bool segs_intersect(V2 p1, V2 p2, V2 p3, V2 p4) {
  float d1 = cross(p4 - p3, p1 - p3), d2 = cross(p4 - p3, p2 - p3);
  float d3 = cross(p2 - p1, p3 - p1), d4 = cross(p2 - p1, p4 - p1);
  return (d1 * d2 < 0) && (d3 * d4 < 0);
}

/* closest features of two core segments: separated -> distance and direction; crossing -> the
   minimum-penetration axis among the edge normals (what GJK/EPA returns for two segments) */
__device__ void seg_seg_closest(V2 a0, V2 a1, V2 b0, V2 b1, V2* nOut, float* dOut) {
  if (segs_intersect(a0, a1, b0, b1)) {
    V2 da = a1 - a0, db = b1 - b0;
    float la = len(da), lb = len(db);
    float best = -INFINITY;
    V2 bn = v2(da.y / la, -da.x / la);
    for (int k = 0; k < 2; k++) {
      V2 ax = k == 0 ? v2(da.y / la, -da.x / la) : v2(-da.y / la, da.x / la);
      float d = fminf(dot(ax, b0), dot(ax, b1)) - fmaxf(dot(ax, a0), dot(ax, a1));
      if (d > best) { best = d; bn = ax; }
    }
    for (int k = 0; k < 2; k++) {
      V2 ax = k == 0 ? v2(db.y / lb, -db.x / lb) : v2(-db.y / lb, db.x / lb);
      float d = fminf(dot(ax, a0), dot(ax, a1)) - fmaxf(dot(ax, b0), dot(ax, b1));
      if (d > best) { best = d; bn = v2(-ax.x, -ax.y); }
    }
    *nOut = bn;
    *dOut = best;
    return;
  }
  float best = INFINITY;
  V2 bn = v2(1, 0);
  for (int k = 0; k < 2; k++) {
    V2 p = k == 0 ? a0 : a1;
    V2 dv = closest_on_seg(p, b0, b1) - p;
    float d = len(dv);
    if (d < best) { best = d; if (d > 0) bn = dv * (1.f / d); }
  }
  for (int k = 0; k < 2; k++) {
    V2 p = k == 0 ? b0 : b1;
    V2 dv = p - closest_on_seg(p, a0, a1);
    float d = len(dv);
    if (d < best) { best = d; if (d > 0) bn = dv * (1.f / d); }
  }
  *nOut = bn;
  *dOut = best;
}

__device__ __forceinline__ float clamp01(float t) { return fminf(fmaxf(t, 0.f), 1.f); }

/* cheap exact test: are two non-crossing-or-crossing segments farther apart than `dist`?  (closest
   points of two segments, Ericson, Real-Time Collision Detection 5.1.9); used to skip the manifold
   construction for foot pairs that are near but not touching */
__device__ bool segments_apart(V2 p1, V2 q1, V2 p2, V2 q2, float dist) {
  V2 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
  float c = dot(d1, r), b = dot(d1, d2);
  float denom = a * e - b * b;
  float s = denom > 1e-6f ? clamp01((b * f - c * e) / denom) : 0.f;
  float t = (b * s + f) / e;
  if (t < 0.f) { t = 0.f; s = clamp01(-c / a); }
  else if (t > 1.f) { t = 1.f; s = clamp01((b - c) / a); }
  V2 dd = (p1 + d1 * s) - (p2 + d2 * t);
  return dot(dd, dd) > dist * dist;
}
__device__ __forceinline__ V2 vlerp(V2 a, V2 b, float t) { return a * (1.f - t) + b * t; }

/* Shape geometry relative to its own body: feet are the segments (-10, +-10)-(10, +-10) (Robot.py:24-27) */
__device__ __forceinline__ void shape_local(const Env E, int s, V2* oa, V2* ob) {
  if (is_foot(s)) {
    float ly = (s & 1) ? -ROB_LEN : ROB_LEN;
    float sn, cs;
    sincos_b(E.f[RC_S_ANG + s], &sn, &cs);
    *oa = v2(-ROB_LEN * cs - ly * sn, -ROB_LEN * sn + ly * cs);
    *ob = v2(ROB_LEN * cs - ly * sn, ROB_LEN * sn + ly * cs);
  } else {
    *oa = v2(0, 0);
    *ob = v2(0, 0);
  }
}

/* Narrow phase in the frame of body a (origin = a's position, dab = position of b relative to a from
   the two-float coordinates), so that penetration depths of ~0.1 px are not lost in float32
   coordinates of ~500 px.  Contact points come out relative to a's position. */
__device__ __noinline__ void narrowphase(const Env E, int a, int b, V2 dab, Manifold* m) {
  V2 ata, atb, bta, btb;
  shape_local(E, a, &ata, &atb);
  shape_local(E, b, &bta, &btb);
  bta = bta + dab;
  btb = btb + dab;
  m->cnt = 0;
  if (!is_foot(a) && !is_foot(b)) { /* CircleToCircleQuery */
    float mind = 20.f;
    V2 d = bta - ata;
    float dsq = lensq(d);
    if (dsq < mind * mind) {
      float dist = sqrtf(dsq);
      m->n = dist ? d * (1.f / dist) : v2(1, 0);
      m->p1[0] = ata + m->n * 10.f;
      m->p2[0] = bta + m->n * -10.f;
      m->hash[0] = 0;
      m->cnt = 1;
    }
  } else if (!is_foot(a)) { /* CircleToSegment */
    V2 closest = closest_on_seg(ata, bta, btb);
    float mind = 10.f + ROB_RAD;
    V2 d = closest - ata;
    float dsq = lensq(d);
    if (dsq < mind * mind) {
      float dist = sqrtf(dsq);
      m->n = dist ? d * (1.f / dist) : seg_normal(bta, btb);
      m->p1[0] = ata + m->n * 10.f;
      m->p2[0] = closest + m->n * -ROB_RAD;
      m->hash[0] = 0;
      m->cnt = 1;
    }
  } else { /* SegmentToSegment + SupportEdgeForSegment + ContactPoints */
    if (segments_apart(ata, atb, bta, btb, 2 * ROB_RAD + 0.01f)) return;
    V2 n;
    float d;
    seg_seg_closest(ata, atb, bta, btb, &n, &d);
    const float r1 = ROB_RAD, r2 = ROB_RAD;
    if (!(d <= r1 + r2)) return;
    V2 tn1 = seg_normal(ata, atb), tn2 = seg_normal(bta, btb);
    V2 e1a, e1b, e2a, e2b;
    int h1a, h1b, h2a, h2b;
    if (dot(tn1, n) > 0.f) { e1a = ata; h1a = 0; e1b = atb; h1b = 1; }
    else { e1a = atb; h1a = 1; e1b = ata; h1b = 0; }
    if (dot(tn2, v2(-n.x, -n.y)) > 0.f) { e2a = bta; h2a = 0; e2b = btb; h2b = 1; }
    else { e2a = btb; h2a = 1; e2b = bta; h2b = 0; }
    m->n = n;
    float d_e1_a = cross(e1a, n), d_e1_b = cross(e1b, n);
    float d_e2_a = cross(e2a, n), d_e2_b = cross(e2b, n);
    float e1_denom = 1.f / (d_e1_b - d_e1_a + 1.17549435e-38f);
    float e2_denom = 1.f / (d_e2_b - d_e2_a + 1.17549435e-38f);
    int cnt = 0;
    {
      V2 p1 = n * r1 + vlerp(e1a, e1b, clamp01((d_e2_b - d_e1_a) * e1_denom));
      V2 p2 = n * -r2 + vlerp(e2a, e2b, clamp01((d_e1_a - d_e2_a) * e2_denom));
      if (dot(p2 - p1, n) <= 0.f) { m->p1[cnt] = p1; m->p2[cnt] = p2; m->hash[cnt] = 1 + (h1a | (h2b << 1)); cnt++; }
    }
    {
      V2 p1 = n * r1 + vlerp(e1a, e1b, clamp01((d_e2_a - d_e1_a) * e1_denom));
      V2 p2 = n * -r2 + vlerp(e2a, e2b, clamp01((d_e1_b - d_e2_a) * e2_denom));
      if (dot(p2 - p1, n) <= 0.f) { m->p1[cnt] = p1; m->p2[cnt] = p2; m->hash[cnt] = 1 + (h1b | (h2a << 1)); cnt++; }
    }
    m->cnt = cnt;
  }
}

/* cpArbiterUpdate + begin callbacks of a touching pair (a, b in cpCollide order, dab = position of b relative to a,
   m = its contact manifold): one lane, pairs in ascending (i, j) order */
__device__ __noinline__ void arbiter_update(const Env E, int a, int b, V2 dab, const Manifold& m) {
  int nArb = E.i[RC_S_NARB];
  int key16 = a | (b << 8);
  int k = 0;
#pragma unroll 1
  for (; k < nArb; k++)
    if ((E.i[RC_S_ARBKEY + k] & 0xffff) == key16) break;
  if (k == nArb) {
    if (nArb >= RC_MAXARB) { E.i[X_DIAG] |= RC_IF_ARB_OVERFLOW; return; }
    E.i[RC_S_NARB] = nArb + 1;
    E.i[RC_S_ARBKEY + k] = key16 | (RC_ARB_FIRST << 16);
    E.i[RC_S_ARBCON + k] = 0;
  }
  int nact = E.i[X_NACTIVE];
  if (nact >= RC_MAXACTIVE) { E.i[X_DIAG] |= RC_IF_ACTIVE_OVERFLOW; return; }
  /* contacts: anchors relative to the bodies, impulses carried over by feature hash */
  int con = E.i[RC_S_ARBCON + k];
  int oldN = con & 255, oh0 = (con >> 8) & 255, oh1 = (con >> 16) & 255;
  float ojn0 = E.f[RC_S_ARBJN0 + k], ojt0 = E.f[RC_S_ARBJT0 + k], ojn1 = E.f[RC_S_ARBJN1 + k], ojt1 = E.f[RC_S_ARBJT1 + k];
#pragma unroll
  for (int c = 0; c < 2; c++) {
    int ci = 2 * nact + c;
    float jn = 0, jt = 0;
    if (c < m.cnt) {
      if (oldN > 0 && m.hash[c] == oh0) { jn = ojn0; jt = ojt0; }
      if (oldN > 1 && m.hash[c] == oh1) { jn = ojn1; jt = ojt1; }
      E.f[X_CR1X + ci] = m.p1[c].x; E.f[X_CR1Y + ci] = m.p1[c].y;
      E.f[X_CR2X + ci] = m.p2[c].x - dab.x; E.f[X_CR2Y + ci] = m.p2[c].y - dab.y;
    }
    E.f[X_CJN + ci] = jn;
    E.f[X_CJT + ci] = jt;
  }
  E.i[RC_S_ARBCON + k] = m.cnt | (m.hash[0] << 8) | ((m.cnt > 1 ? m.hash[1] : 0) << 16);
  E.f[X_ANX + nact] = m.n.x;
  E.f[X_ANY + nact] = m.n.y;
  E.f[X_AE + nact] = shape_e(a) * shape_e(b);
  E.f[X_AU + nact] = shape_u(a) * shape_u(b);
  int state = (E.i[RC_S_ARBKEY + k] >> 16) & 255;
  if (state == RC_ARB_CACHED) state = RC_ARB_FIRST;
  E.i[RC_S_ARBKEY + k] = key16 | (state << 16); /* age = 0 */
  if (state == RC_ARB_FIRST) { /* begin callbacks (all accept) */
    if (is_foot(a) && is_foot(b)) cb_robot_pushing_det(E, a, b);
    else if (a == BALL && is_foot(b)) cb_ball_collision(E, b);
  }
  E.i[X_ACTIVE + nact] = k;
  E.i[X_NACTIVE] = nact + 1;
  E.i[X_TOUCHED] |= 1 << k;
}

/* cpSpaceCollideShapes + cpArbiterUpdate + begin callbacks for an AABB-overlapping pair i < j */
__device__ __noinline__ void collide_pair(const Env E, int i, int j) {
  int a = i, b = j;
  if (is_foot(a) && !is_foot(b)) { a = j; b = i; } /* cpCollide orders by primitive type */
  Manifold m;
  V2 dab = pos_diff(E, b, a);
  narrowphase(E, a, b, dab, &m);
  if (m.cnt == 0) return;
  arbiter_update(E, a, b, dab, m);
}

/* ================================================================= space.step */
struct BodyVel {
  float vx, vy, w;
};

__device__ __forceinline__ void load_body(const Env E, int s, float* vx, float* vy, float* w, float* vbx, float* vby, float* wb) {
  if (s < RC_NBODY) {
    *vx = E.f[RC_S_VX + s]; *vy = E.f[RC_S_VY + s]; *w = E.f[RC_S_W + s];
    *vbx = E.f[RC_S_VBX + s]; *vby = E.f[RC_S_VBY + s]; *wb = E.f[RC_S_WB + s];
  } else {
    *vx = *vy = *w = *vbx = *vby = *wb = 0.f;
  }
}

/* cpArbiterApplyImpulse for every active contact, in order (Gauss-Seidel): lane 0 */
__device__ __noinline__ void contacts_iteration(const Env E) {
  int nact = E.i[X_NACTIVE];
#pragma unroll 1
  for (int q = 0; q < nact; q++) {
    int key = E.i[RC_S_ARBKEY + E.i[X_ACTIVE + q]];
    int a = key & 255, b = (key >> 8) & 255;
    int ncon = E.i[RC_S_ARBCON + E.i[X_ACTIVE + q]] & 255;
    V2 n = v2(E.f[X_ANX + q], E.f[X_ANY + q]);
    float u = E.f[X_AU + q];
    float ami = body_minv(a), aii = body_iinv(a), bmi = body_minv(b), bii = body_iinv(b);
    float avx, avy, aw, avbx, avby, awb, bvx, bvy, bw, bvbx, bvby, bwb;
    load_body(E, a, &avx, &avy, &aw, &avbx, &avby, &awb);
    load_body(E, b, &bvx, &bvy, &bw, &bvbx, &bvby, &bwb);
#pragma unroll 1
    for (int c = 0; c < ncon; c++) {
      int ci = 2 * q + c;
      V2 r1 = v2(E.f[X_CR1X + ci], E.f[X_CR1Y + ci]), r2 = v2(E.f[X_CR2X + ci], E.f[X_CR2Y + ci]);
      float nMass = E.f[X_CNMASS + ci], tMass = E.f[X_CTMASS + ci];
      V2 vb1 = v2(avbx - r1.y * awb, avby + r1.x * awb);
      V2 vb2 = v2(bvbx - r2.y * bwb, bvby + r2.x * bwb);
      V2 vr = v2((bvx - r2.y * bw) - (avx - r1.y * aw), (bvy + r2.x * bw) - (avy + r1.x * aw));
      float vbn = dot(vb2 - vb1, n);
      float vrn = dot(vr, n);
      float vrt = dot(vr, v2(-n.y, n.x));
      float jbn = (E.f[X_CBIAS + ci] - vbn) * nMass;
      float jbnOld = E.f[X_CJBIAS + ci];
      float jBias = fmaxf(jbnOld + jbn, 0.f);
      float jn = -(E.f[X_CBOUNCE + ci] + vrn) * nMass;
      float jnOld = E.f[X_CJN + ci];
      float jnAcc = fmaxf(jnOld + jn, 0.f);
      float jtMax = u * jnAcc;
      float jt = -vrt * tMass;
      float jtOld = E.f[X_CJT + ci];
      float jtAcc = fminf(fmaxf(jtOld + jt, -jtMax), jtMax);
      E.f[X_CJBIAS + ci] = jBias;
      E.f[X_CJN + ci] = jnAcc;
      E.f[X_CJT + ci] = jtAcc;
      /* bias impulse */
      V2 jb = n * (jBias - jbnOld);
      avbx -= jb.x * ami; avby -= jb.y * ami; awb -= aii * cross(r1, jb);
      bvbx += jb.x * bmi; bvby += jb.y * bmi; bwb += bii * cross(r2, jb);
      float dn = jnAcc - jnOld, dtt = jtAcc - jtOld;
      V2 j = v2(n.x * dn - n.y * dtt, n.x * dtt + n.y * dn);
      avx -= j.x * ami; avy -= j.y * ami; aw -= aii * cross(r1, j);
      bvx += j.x * bmi; bvy += j.y * bmi; bw += bii * cross(r2, j);
    }
    if (a < RC_NBODY) {
      E.f[RC_S_VX + a] = avx; E.f[RC_S_VY + a] = avy; E.f[RC_S_W + a] = aw;
      E.f[RC_S_VBX + a] = avbx; E.f[RC_S_VBY + a] = avby; E.f[RC_S_WB + a] = awb;
    }
    if (b < RC_NBODY) {
      E.f[RC_S_VX + b] = bvx; E.f[RC_S_VY + b] = bvy; E.f[RC_S_W + b] = bw;
      E.f[RC_S_VBX + b] = bvbx; E.f[RC_S_VBY + b] = bvby; E.f[RC_S_WB + b] = bwb;
    }
  }
}

/* ---- cold parts of space.step (pairs in contact): out of line, so that the contact-free substep --
   the common case -- stays a short, instruction-cache resident piece of code */

/* cached pairs that did not collide this substep: age, separate callbacks, expiry after 3 substeps
   (cpSpaceArbiterSetFilter); lane = cached pair */
__device__ __noinline__ void age_cached_pairs(const Env E) {
  const int lane = E.lane;
  int nArb = E.i[RC_S_NARB];
  unsigned touched = (unsigned)E.i[X_TOUCHED];
  int key = 0, con = 0;
  float jn0 = 0, jt0 = 0, jn1 = 0, jt1 = 0;
  bool have = lane < nArb;
  bool keep = false, sep = false;
  if (have) {
    key = E.i[RC_S_ARBKEY + lane]; con = E.i[RC_S_ARBCON + lane];
    jn0 = E.f[RC_S_ARBJN0 + lane]; jt0 = E.f[RC_S_ARBJT0 + lane];
    jn1 = E.f[RC_S_ARBJN1 + lane]; jt1 = E.f[RC_S_ARBJT1 + lane];
    keep = true;
    if (!((touched >> lane) & 1)) {
      int age = ((key >> 24) & 255) + 1;
      int state = (key >> 16) & 255;
      int a = key & 255, b = (key >> 8) & 255;
      if (state != RC_ARB_CACHED) {
        state = RC_ARB_CACHED;
        sep = !(a == BALL); /* (Ball, Robot) and (Ball, Goalpost) have no separate handler */
      }
      key = a | (b << 8) | (state << 16) | (age << 24);
      keep = age < 3;
    }
  }
  unsigned sepmask = __ballot_sync(FULLMASK, sep);
  unsigned keepmask = __ballot_sync(FULLMASK, keep);
  if (sepmask && lane == 0) {
    unsigned mk = sepmask;
    while (mk) {
      int k = __ffs(mk) - 1;
      mk &= mk - 1;
      int kk = E.i[RC_S_ARBKEY + k];
      cb_separate(E, kk & 255, (kk >> 8) & 255);
    }
  }
  __syncwarp();
  int newIdx = __popc(keepmask & ((1u << lane) - 1));
  if (keep) {
    E.i[RC_S_ARBKEY + newIdx] = key; E.i[RC_S_ARBCON + newIdx] = con;
    E.f[RC_S_ARBJN0 + newIdx] = jn0; E.f[RC_S_ARBJT0 + newIdx] = jt0;
    E.f[RC_S_ARBJN1 + newIdx] = jn1; E.f[RC_S_ARBJT1 + newIdx] = jt1;
  }
  int nact = E.i[X_NACTIVE];
  int myActive = lane < nact ? E.i[X_ACTIVE + lane] : 0;
  int remapped = __shfl_sync(FULLMASK, newIdx, myActive);
  __syncwarp();
  if (lane < nact) E.i[X_ACTIVE + lane] = remapped;
  if (lane == 0) E.i[RC_S_NARB] = __popc(keepmask);
  if (have && lane >= __popc(keepmask)) { /* records stay canonical: unused slots are zero */
    E.i[RC_S_ARBKEY + lane] = 0; E.i[RC_S_ARBCON + lane] = 0;
    E.f[RC_S_ARBJN0 + lane] = 0.f; E.f[RC_S_ARBJT0 + lane] = 0.f;
    E.f[RC_S_ARBJN1 + lane] = 0.f; E.f[RC_S_ARBJT1 + lane] = 0.f;
  }
  __syncwarp();
}

/* cpArbiterPreStep of the active contacts: lane = 2 * pair + contact */
__device__ __noinline__ void prestep_contacts(const Env E, float biasCoef) {
  const int lane = E.lane;
  const int nact = E.i[X_NACTIVE];
  const float slop = 0.1f;
  int q = lane >> 1, c = lane & 1;
  if (q < nact) {
    int k = E.i[X_ACTIVE + q];
    int key = E.i[RC_S_ARBKEY + k];
    int ncon = E.i[RC_S_ARBCON + k] & 255;
    if (c < ncon) {
      int a = key & 255, b = (key >> 8) & 255;
      V2 n = v2(E.f[X_ANX + q], E.f[X_ANY + q]);
      V2 r1 = v2(E.f[X_CR1X + lane], E.f[X_CR1Y + lane]), r2 = v2(E.f[X_CR2X + lane], E.f[X_CR2Y + lane]);
      float ami = body_minv(a), aii = body_iinv(a), bmi = body_minv(b), bii = body_iinv(b);
      float r1cn = cross(r1, n), r2cn = cross(r2, n);
      V2 t = v2(-n.y, n.x);
      float r1ct = cross(r1, t), r2ct = cross(r2, t);
      E.f[X_CNMASS + lane] = 1.f / (ami + aii * r1cn * r1cn + bmi + bii * r2cn * r2cn);
      E.f[X_CTMASS + lane] = 1.f / (ami + aii * r1ct * r1ct + bmi + bii * r2ct * r2ct);
      V2 bd = pos_diff(E, b, a);
      float dist = dot(r2 - r1 + bd, n);
      E.f[X_CBIAS + lane] = -biasCoef * fminf(0.f, dist + slop) / DT;
      E.f[X_CJBIAS + lane] = 0.f;
      float avx, avy, aw, bvx, bvy, bw, d0, d1, d2;
      load_body(E, a, &avx, &avy, &aw, &d0, &d1, &d2);
      load_body(E, b, &bvx, &bvy, &bw, &d0, &d1, &d2);
      V2 vr = v2((bvx - r2.y * bw) - (avx - r1.y * aw), (bvy + r2.x * bw) - (avy + r1.x * aw));
      E.f[X_CBOUNCE + lane] = dot(vr, n) * E.f[X_AE + q];
    }
  }
  __syncwarp();
}

/* cached contact impulses re-applied in pair order (cpArbiterApplyCachedImpulse): lane 0 */
__device__ __noinline__ void warm_start_contacts(const Env E, float dt_coef) {
  if (E.lane == 0) {
    const int nact = E.i[X_NACTIVE];
#pragma unroll 1
    for (int q = 0; q < nact; q++) {
      int k = E.i[X_ACTIVE + q];
      int key = E.i[RC_S_ARBKEY + k];
      if (((key >> 16) & 255) == RC_ARB_FIRST) continue;
      int a = key & 255, b = (key >> 8) & 255;
      int ncon = E.i[RC_S_ARBCON + k] & 255;
      V2 n = v2(E.f[X_ANX + q], E.f[X_ANY + q]);
#pragma unroll 1
      for (int c = 0; c < ncon; c++) {
        int ci = 2 * q + c;
        float jn = E.f[X_CJN + ci], jt = E.f[X_CJT + ci];
        V2 j = v2(n.x * jn - n.y * jt, n.x * jt + n.y * jn) * dt_coef;
        V2 r1 = v2(E.f[X_CR1X + ci], E.f[X_CR1Y + ci]), r2 = v2(E.f[X_CR2X + ci], E.f[X_CR2Y + ci]);
        if (a < RC_NBODY) {
          E.f[RC_S_VX + a] -= j.x * body_minv(a); E.f[RC_S_VY + a] -= j.y * body_minv(a);
          E.f[RC_S_W + a] -= body_iinv(a) * cross(r1, j);
        }
        if (b < RC_NBODY) {
          E.f[RC_S_VX + b] += j.x * body_minv(b); E.f[RC_S_VY + b] += j.y * body_minv(b);
          E.f[RC_S_W + b] += body_iinv(b) * cross(r2, j);
        }
      }
    }
  }
  __syncwarp();
}

/* accumulated impulses back into the cached pairs, post-solve callbacks in pair order, deferred
   space.add(joint) */
__device__ __noinline__ void post_solve(const Env E) {
  const int lane = E.lane;
  const int nact = E.i[X_NACTIVE];
  if (lane < nact) {
    int k = E.i[X_ACTIVE + lane];
    E.f[RC_S_ARBJN0 + k] = E.f[X_CJN + 2 * lane]; E.f[RC_S_ARBJT0 + k] = E.f[X_CJT + 2 * lane];
    E.f[RC_S_ARBJN1 + k] = E.f[X_CJN + 2 * lane + 1]; E.f[RC_S_ARBJT1 + k] = E.f[X_CJT + 2 * lane + 1];
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll 1
    for (int q = 0; q < nact; q++) {
      int key = E.i[RC_S_ARBKEY + E.i[X_ACTIVE + q]];
      int a = key & 255, b = (key >> 8) & 255;
      if (is_foot(a) && is_foot(b)) cb_robot_collision(E, a, b);
      else if (a >= POST0 && is_foot(b)) cb_goalpost_collision(E, b, a);
    }
  }
  __syncwarp();
  if (lane < E.c->P) joint_add_now(E, lane);
  __syncwarp();
}

/* do two shapes (cached world geometry) come within their radii + margin?  lane-parallel filter in front
   of the serial narrow phase; the margin covers the float32 rounding of the absolute coordinates */
__device__ bool shapes_touch(const Env E, int i, int j) {
  V2 ia, ib, ja, jb;
  shape_geom(E, i, &ia, &ib);
  shape_geom(E, j, &ja, &jb);
  const float m = 0.02f;
  if (is_foot(i) && is_foot(j)) return !segments_apart(ia, ib, ja, jb, 2 * ROB_RAD + m);
  if (!is_foot(i) && !is_foot(j)) return lensq(ia - ja) < (20.f + m) * (20.f + m);
  V2 c = is_foot(i) ? ja : ia;
  V2 sa = is_foot(i) ? ia : ja, sb = is_foot(i) ? ib : jb;
  return lensq(c - closest_on_seg(c, sa, sb)) < (10.f + ROB_RAD + m) * (10.f + ROB_RAD + m);
}

/* group pairs that came close (bit k of g[k / 32] = pair g_pairs[k]) and own-feet pairs (bit = odd foot):
   expanded to shape pairs (lane 0) and tested exactly in parallel (lane = candidate) on the cached shape geometry and
   the bounding-circle centres X_BCX / X_BCY; returns how many touch (their keys stay in X_CAND).  Reads nothing else
   of the record: the register-resident substep calls it without flushing. */
__device__ __noinline__ int filter_candidates(const Env E, unsigned own, unsigned g0, unsigned g1, unsigned g2, unsigned g3) {
  const int lane = E.lane;
  unsigned short* cand = reinterpret_cast<unsigned short*>(g_smem + E.i.base_() + X_CAND);
  if (lane == 0) {
    int n = 0;
    while (own) {
      int j = __ffs(own) - 1;
      own &= own - 1;
      cand[n++] = (j - 1) * 32 + j;
    }
    unsigned g[4] = {g0, g1, g2, g3};
#pragma unroll 1
    for (int w = 0; w < 4; w++) {
      unsigned m = g[w];
      while (m) {
        int k = 32 * w + __ffs(m) - 1;
        m &= m - 1;
        int a = g_pairs[k][0], b = g_pairs[k][1];
        int a0 = a < 10 ? 2 * a : a + 10, na = a < 10 ? 2 : 1;
        int b0 = b < 10 ? 2 * b : b + 10, nb = b < 10 ? 2 : 1;
#pragma unroll 1
        for (int x = 0; x < na; x++)
#pragma unroll 1
          for (int y = 0; y < nb; y++) {
            int i = a0 + x, j = b0 + y;
            float dx = E.f[X_BCX + i] - E.f[X_BCX + j], dy = E.f[X_BCY + i] - E.f[X_BCY + j];
            float rr = (i < 20 ? ROB_LEN + ROB_RAD : 10.f) + (j < 20 ? ROB_LEN + ROB_RAD : 10.f) + 0.05f;
            if (dx * dx + dy * dy < rr * rr) { /* bounding circles of the two shapes */
              if (n < RC_MAXCAND) cand[n++] = i * 32 + j;
              else E.i[X_DIAG] |= RC_IF_ACTIVE_OVERFLOW;
            }
          }
      }
    }
    E.i[X_NCAND] = n;
  }
  __syncwarp();
  const int n = E.i[X_NCAND];
  int kept = 0;
  for (int base = 0; base < n; base += 32) {
    int c = base + lane;
    int key = c < n ? cand[c] : 0;
    bool touch = c < n && shapes_touch(E, key >> 5, key & 31);
    unsigned km = __ballot_sync(FULLMASK, touch);
    __syncwarp();
    if (touch) cand[kept + __popc(km & ((1u << lane) - 1))] = key; /* kept + rank <= c: no overwrite of unread keys */
    kept += __popc(km);
    __syncwarp();
  }
  return kept;
}

/* the touching candidates (filter_candidates left them in X_CAND) collided in ascending (i, j) order: lane 0 */
__device__ __noinline__ void collide_kept(const Env E, int kept) {
  unsigned short* cand = reinterpret_cast<unsigned short*>(g_smem + E.i.base_() + X_CAND);
  if (E.lane == 0) {
    for (int x = 1; x < kept; x++) { /* insertion sort: tiny */
      int key = cand[x], y = x - 1;
      while (y >= 0 && cand[y] > key) { cand[y + 1] = cand[y]; y--; }
      cand[y + 1] = key;
    }
    for (int x = 0; x < kept; x++) collide_pair(E, cand[x] >> 5, cand[x] & 31);
  }
  __syncwarp();
}

/* this lane's group pairs of the four broad-phase rounds, one byte each (a << 4 | b); pairs with a
   robot that does not exist point at the far-away dummy group */
__device__ __forceinline__ uint32_t broadphase_pairs(const Env E) {
  uint32_t packed = 0;
  const int P = E.c->P;
#pragma unroll
  for (int rnd = 0; rnd < 4; rnd++) {
    int k = E.lane + 32 * rnd;
    uint32_t byte = 0xFE; /* (dummy 15, post 14) */
    if (k < 99) {
      int a = g_pairs[k][0], b = g_pairs[k][1];
      if ((a >= 10 || a < P) && (b >= 10 || b < P)) byte = (uint32_t)(a << 4 | b);
    }
    packed |= byte << (8 * rnd);
  }
  if (E.lane == 0) { E.f[X_GX + 15] = 1e18f; E.f[X_GY + 15] = 1e18f; E.f[X_GR + 15] = 0.f; }
  return packed;
}

/* cpSpaceStep after the positions are integrated and the broad phase has run: narrow phase of the candidate pairs,
   cached-pair ageing, friction callbacks, warm start, solver sweeps, post-solve.  Reads and writes the record in shared
   memory (the register-resident substep of rc_phys2.cuh enters here whenever a pair comes close or is cached). */
__device__ __noinline__ void space_step_rest(const Env E, int kept, int prevStepped) {
  const int lane = E.lane;
  const int P = E.c->P;
  const float dt = DT;
  const bool active = lane < NSHAPE && shape_active(E, lane);
  if (kept) collide_kept(E, kept);
  if (E.i[RC_S_NARB]) age_cached_pairs(E);
  const int nact = E.i[X_NACTIVE];
  const float biasCoef = 6.1259606640e-02f; /* 1 - collisionBias^dt, collisionBias = 0.9^60 */
  if (nact) prestep_contacts(E, biasCoef);
  /* joints pre-step (lane = robot): pivot with both anchors at the centres of gravity, rotary limit 0..0 */
  const int L = 2 * lane, R = 2 * lane + 1;
  const bool isRobot = lane < P;
  float pivBx = 0, pivBy = 0, rotBias = 0, pivJx = 0, pivJy = 0, rotJ = 0;
  bool pivOn = false;
  const float pivK = 1.f / (2 * FOOT_MINV), rotISum = 1.f / (2 * FOOT_IINV);
  if (isRobot) {
    pivOn = !(E.i[RC_S_RFLAGS + lane] & RC_RF_JOINT_ABSENT);
    pivJx = E.f[RC_S_PIVJX + lane]; pivJy = E.f[RC_S_PIVJY + lane]; rotJ = E.f[RC_S_ROTJ + lane];
    if (pivOn) {
      const float c = -2.2762779044e+00f; /* -(1 - errorBias^dt) / dt, joint.error_bias = 0.1 (Robot.py:59) */
      V2 d = pos_diff(E, R, L);
      pivBx = d.x * c;
      pivBy = d.y * c;
    }
    float dist = __fadd_rn(__fsub_rn(E.f[RC_S_ANG + R], E.f[RC_S_ANG + L]), __fsub_rn(E.f[RC_S_ANGL + R], E.f[RC_S_ANGL + L]));
    float pdist = -dist; /* min = max = 0 */
    rotBias = -biasCoef * pdist / dt;
    if (rotBias == 0) rotJ = 0;
  }
  /* velocity update through the friction callbacks (cutils.py:93-140), lane = body */
  if (lane <= BALL && active) {
    float minv = body_minv(lane), iinv = body_iinv(lane), m = body_m(lane);
    float friction = lane == BALL ? 2.8e-2f : 1e-3f, rotFriction = lane == BALL ? 1e-3f : 1e-2f;
    float spin = lane == BALL ? 5e-2f : 0.f;
    float x = E.f[RC_S_VX + lane] + (E.f[RC_S_FX + lane] * minv) * dt;
    float y = E.f[RC_S_VY + lane] + (E.f[RC_S_FY + lane] * minv) * dt;
    float theta = E.f[RC_S_W + lane] + E.f[RC_S_TQ + lane] * iinv * dt;
    E.f[RC_S_FX + lane] = 0.f; E.f[RC_S_FY + lane] = 0.f; E.f[RC_S_TQ + lane] = 0.f;
    float factor = friction * m, rotFactor = rotFriction * m;
    float length = 1.f / (fabsf(x) + fabsf(y) + 1e-5f);
    float a0 = x * factor * length, a1 = y * factor * length;
    a0 += a1 * spin * theta;
    a1 -= a0 * spin * theta;
    x = fabsf(x) < factor ? 0.f : x - a0;
    y = fabsf(y) < factor ? 0.f : y - a1;
    theta = fabsf(theta) < rotFactor ? 0.f : theta - (theta > 0 ? rotFactor : -rotFactor);
    E.f[RC_S_VX + lane] = x; E.f[RC_S_VY + lane] = y; E.f[RC_S_W + lane] = theta;
  }
  __syncwarp();
  /* warm start: cached contact impulses in pair order, then the joints */
  const float dt_coef = prevStepped ? 1.f : 0.f;
  if (nact) warm_start_contacts(E, dt_coef);
  float lvx = 0, lvy = 0, lw = 0, rvx = 0, rvy = 0, rw = 0;
  if (isRobot) {
    lvx = E.f[RC_S_VX + L]; lvy = E.f[RC_S_VY + L]; lw = E.f[RC_S_W + L];
    rvx = E.f[RC_S_VX + R]; rvy = E.f[RC_S_VY + R]; rw = E.f[RC_S_W + R];
    if (pivOn) {
      float jx = pivJx * dt_coef, jy = pivJy * dt_coef;
      lvx -= jx * FOOT_MINV; lvy -= jy * FOOT_MINV;
      rvx += jx * FOOT_MINV; rvy += jy * FOOT_MINV;
    }
    float j = rotJ * dt_coef;
    lw -= j * FOOT_IINV;
    rw += j * FOOT_IINV;
  }
  /* solver: 10 iterations of {all contacts in order, then all joints}; without contacts the
     joints of different robots are independent and iterate in registers */
  /* Without contacts every joint is an independent one-constraint system whose first sweep already
     solves it exactly (K is the exact inverse); sweeps 3..10 of the reference only move the last bit,
     so two sweeps are run.  With contacts all ten sweeps are needed and run. */
  const int sweeps = nact ? 10 : 2;
#pragma unroll 1
  for (int it = 0; it < sweeps; it++) {
    if (nact) {
      if (isRobot) {
        E.f[RC_S_VX + L] = lvx; E.f[RC_S_VY + L] = lvy; E.f[RC_S_W + L] = lw;
        E.f[RC_S_VX + R] = rvx; E.f[RC_S_VY + R] = rvy; E.f[RC_S_W + R] = rw;
      }
      __syncwarp();
      if (lane == 0) contacts_iteration(E);
      __syncwarp();
      if (isRobot) {
        lvx = E.f[RC_S_VX + L]; lvy = E.f[RC_S_VY + L]; lw = E.f[RC_S_W + L];
        rvx = E.f[RC_S_VX + R]; rvy = E.f[RC_S_VY + R]; rw = E.f[RC_S_W + R];
      }
    }
    if (pivOn) { /* cpPivotJoint applyImpulse, r1 = r2 = 0 */
      float jx = pivK * (pivBx - (rvx - lvx)), jy = pivK * (pivBy - (rvy - lvy));
      pivJx += jx; pivJy += jy;
      lvx -= jx * FOOT_MINV; lvy -= jy * FOOT_MINV;
      rvx += jx * FOOT_MINV; rvy += jy * FOOT_MINV;
    }
    if (rotBias != 0) { /* cpRotaryLimitJoint applyImpulse */
      float j = -(rotBias + (rw - lw)) * rotISum;
      float jOld = rotJ;
      rotJ = rotBias < 0 ? fmaxf(jOld + j, 0.f) : fminf(jOld + j, 0.f);
      j = rotJ - jOld;
      lw -= j * FOOT_IINV;
      rw += j * FOOT_IINV;
    }
  }
  if (isRobot) {
    E.f[RC_S_VX + L] = lvx; E.f[RC_S_VY + L] = lvy; E.f[RC_S_W + L] = lw;
    E.f[RC_S_VX + R] = rvx; E.f[RC_S_VY + R] = rvy; E.f[RC_S_W + R] = rw;
    E.f[RC_S_PIVJX + lane] = pivJx; E.f[RC_S_PIVJY + lane] = pivJy; E.f[RC_S_ROTJ + lane] = rotJ;
  }
  __syncwarp();
  if (nact) post_solve(E);
}

/* ================================================================= observations */
/* getFullState(agent): RoboCupEnvironment.py:1162-1188 -> one RC_FULL_WORDS record, lanes cooperate */
__device__ __noinline__ void full_state_agent(const Env E, int ag, float* dst /* global or shared */) {
  int lane = E.lane;
  int P = E.c->P;
  float team = (float)team_of(E, ag);
  if (lane == 0) {
    V2 bp = ball_pos(E);
    dst[RC_U_BALL + 0] = (bp.x - W_ / 2) * STDNORM * team;
    dst[RC_U_BALL + 1] = (bp.y - H_ / 2) * STDNORM * team;
    dst[RC_U_BALL + 2] = (float)(E.i[RC_S_BALLOWNED] * team_of(E, ag));
    dst[RC_U_BALL + 3] = (ag == E.i[RC_S_CLOSEST] || ag == E.i[RC_S_CLOSEST + 1]) ? 1.f : 0.f;
    V2 p = robot_pos(E, ag);
    float a = robot_angle(E, ag), h = E.f[RC_S_HEAD + ag];
    float s, c;
    dst[RC_U_SELF + 0] = (p.x - W_ / 2) * STDNORM * team;
    dst[RC_U_SELF + 1] = (p.y - H_ / 2) * STDNORM * team;
    sincos_b(a + h, &s, &c);
    dst[RC_U_SELF + 2] = c; dst[RC_U_SELF + 3] = s;
    sincos_b(h, &s, &c);
    dst[RC_U_SELF + 4] = c; dst[RC_U_SELF + 5] = s;
    dst[RC_U_SELF + 6] = team;
    dst[RC_U_SELF + 7] = (E.i[RC_S_RFLAGS + ag] & (RC_RF_FALLEN | RC_RF_PENALIZED)) ? 1.f : 0.f;
    dst[RC_U_OTHERS + 54] = 0.f; dst[RC_U_OTHERS + 55] = 0.f;
  }
  if (lane >= 1 && lane < RC_MAXP) {
    int k = lane - 1;
    float* o = dst + RC_U_OTHERS + 6 * k;
    if (k < P - 1) {
      int r = k < ag ? k : k + 1;
      V2 q = robot_pos(E, r);
      float s, c;
      sincos_b(robot_angle(E, r), &s, &c);
      o[0] = (q.x - W_ / 2) * STDNORM * team;
      o[1] = (q.y - H_ / 2) * STDNORM * team;
      o[2] = c; o[3] = s;
      o[4] = (float)(team_of(E, r) * team_of(E, ag));
      o[5] = (E.i[RC_S_RFLAGS + r] & (RC_RF_FALLEN | RC_RF_PENALIZED)) ? 1.f : 0.f;
    } else {
      o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.f;
    }
  }
}

/* insert a row into one list of the staged record (Python list.insert semantics, capped) */
__device__ __noinline__ void stage_insert(const Env E, int type, int at, const float* row) {
  const int off[6] = {RC_O_BALL, RC_O_ROBOT, RC_O_POST, RC_O_CROSS, RC_O_FCROSS, RC_O_LINE};
  const int cap[6] = {RC_CAP_BALL, RC_CAP_ROBOT, RC_CAP_POST, RC_CAP_CROSS, RC_CAP_FCROSS, RC_CAP_LINE};
  const int feat[6] = {RC_F_BALL, RC_F_ROBOT, RC_F_POST, RC_F_CROSS, RC_F_FCROSS, RC_F_LINE};
  int n = E.i[V_CNT + type];
  int c = cap[type], f = feat[type];
  if (at > n) at = n;
  if (at >= c) return;
  float* base = E.f + V_STAGE + off[type];
  int last = n < c ? n : c - 1;
  for (int k = last; k > at; k--)
    for (int w = 0; w < f; w++) base[k * f + w] = base[(k - 1) * f + w];
  for (int w = 0; w < f; w++) base[at * f + w] = row[w];
  if (n < c) E.i[V_CNT + type] = n + 1;
}

__device__ __noinline__ void polar_row(float* row, V2 p, float radius, float sizeMean, float t1, float t2, float team) {
  /* cutils.py:308-315: (scale(dist), cos, sin of atan2(y*team, x*team), size, tags * team) */
  float dist = len(p);
  float inv = dist > 0 ? 1.f / dist : 0.f;
  row[0] = (dist * STDNORM - 0.5f) / 0.5f;
  row[1] = dist > 0 ? p.x * team * inv : 1.f;
  row[2] = dist > 0 ? p.y * team * inv : 0.f;
  row[3] = (radius - sizeMean) * 2.f;
  row[4] = t1 * team;
  row[5] = t2 * team;
}

/* random false positives (RoboCupEnvironment.py:1279-1305): `hits` = the loop indices whose first draw
   fired; they are rare and inserted in loop order.  Out of line: keeps the hot vision code small. */
__device__ __noinline__ void false_positives(const Env E, int ag, uint32_t tick, unsigned hits, float team) {
  const int lane = E.lane;
  const float maxVis1 = (W_ * 0.8f) * (W_ * 0.8f);
  while (hits) {
    int i = __ffs(hits) - 1;
    hits &= hits - 1;
    if (lane == 0) {
      Rand6 ru = rng_block(E, tick, RC_SITE_OBS_FP, ag, i, 0);
      Rand6 r2 = rng_block(E, tick, RC_SITE_OBS_FP, ag, i, 1); /* k = 6, 7 */
      int c = randint_from(ru.u(1), 0, 5);
      float d = ru.u(2) * sqrtf(maxVis1);
      float a = ru.u(3) * 2 * ROB_FOV - ROB_FOV;
      V2 fp = rot(v2(d, 0), a);
      float sz = 1.f - 0.4f * (ru.u(4) - 0.5f);
      float closest = (ag == E.i[RC_S_CLOSEST] || ag == E.i[RC_S_CLOSEST + 1]) ? 1.f : 0.f;
      float row[8];
      if (c == 0) {
        row[0] = fp.x * STDNORM2; row[1] = fp.y * STDNORM2; row[2] = (BALL_RADIUS * 2 * sz - BALL_RADIUS * 2) * 2.f;
        row[3] = 0.f; row[4] = closest;
        stage_insert(E, 0, E.i[V_CNT + 0], row);
      } else if (c == 1) {
        float s, cs;
        sincos_b((ru.u(5) - 0.5f) * 2 * PI_F, &s, &cs);
        row[0] = fp.x * STDNORM2; row[1] = fp.y * STDNORM2; row[2] = (ROB_TOTAL * sz - ROB_TOTAL) * 2.f;
        row[3] = cs; row[4] = s;
        row[5] = r2.u(0) > 0.5f ? -1.f : 1.f;
        row[6] = r2.u(1) > 0.9f ? 1.f : 0.f;
        int at = E.i[V_CNT + 1];
        if (at < RC_CAP_ROBOT) { E.f[V_RAW + 3 * at] = fp.x; E.f[V_RAW + 3 * at + 1] = fp.y; E.f[V_RAW + 3 * at + 2] = 1.f; }
        stage_insert(E, 1, at, row);
      } else if (c <= 4) {
        float t1 = (float)randint_from(ru.u(5), -1, 1), t2 = (float)randint_from(r2.u(0), -1, 1);
        float rad0 = c == 2 ? POST_RADIUS : PEN_RADIUS;
        polar_row(row, fp, rad0 * sz, rad0, t1, t2, team);
        if (c == 4) {
          float s, cs;
          sincos_b(r2.u(1) * PI_F * 2, &s, &cs);
          row[6] = cs; row[7] = -s;
          stage_insert(E, 4, E.i[V_CNT + 3], row); /* inserted at len(crossDets) (reference behaviour) */
        } else stage_insert(E, c, E.i[V_CNT + c], row);
      }
    }
    __syncwarp();
  }
}

/* observation encoders: 2-ulp reciprocal square root / division are ample for values that the reference
   rounds to float32 anyway, and save the IEEE sequences' slow-path checks */
__device__ __forceinline__ float fast_len(V2 a) {
  float l2 = lensq(a);
  return l2 > 0.f ? l2 * rsqrtf(l2) : 0.f;
}

/* getAgentVision of one observer (PARTIAL): builds the record in shared memory (V_STAGE).
   Point objects: lane = object (ball, 9 robot slots, 4 posts, 3 penalty crosses, 16 line crosses = 33,
   so a second, usually empty, pass handles the last line cross); lines: lane = line. */
template <bool RAW>
__device__ __noinline__ void agent_vision(const Env E, int ag, uint32_t tick) {
  const int lane = E.lane;
  const int P = E.c->P;
  const float magn = E.c->noise_mag, randB = E.c->rand_base;
  const bool realistic = E.c->noise_type == DYNENV_NOISE_REALISTIC;
  const int teamI = team_of(E, ag);
  const float team = (float)teamI;
  const V2 pos = robot_pos(E, ag);
  const float headAngle = robot_angle(E, ag) + E.f[RC_S_HEAD + ag];
  float hs, hc;
  sincos_b(headAngle, &hs, &hc);
  /* FoV edge directions: (1,0) rotated by headAngle +- pi/4 */
  const float q = 0.70710678118654752f;
  const V2 vec1 = v2((hc - hs) * q, (hs + hc) * q), vec2 = v2((hc + hs) * q, (hs - hc) * q);
  /* _setup_vision(0.4, 0.8): environment_base.py:161-172 */
  const float maxVis0 = (W_ * 0.4f) * (W_ * 0.4f), maxVis1 = (W_ * 0.8f) * (W_ * 0.8f);
  const float agFlag = (E.i[RC_S_RFLAGS + ag] & (RC_RF_FALLEN | RC_RF_PENALIZED)) ? 1.f : 0.f;
  const float closest = (ag == E.i[RC_S_CLOSEST] || ag == E.i[RC_S_CLOSEST + 1]) ? 1.f : 0.f;
  const unsigned lt = (1u << lane) - 1;
  for (int w = lane; w < RC_OBS_WORDS / 4; w += 32) reinterpret_cast<float4*>(E.f + V_STAGE)[w] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();

  int nb = 0, nr = 0, np = 0, ncr = 0, nf = 0;
  unsigned robHas = 0; /* lanes 1..9 of pass 0 whose robot is inside the field of view */
  for (int pass = 0; pass < 2; pass++) {
    const int obj = pass * 32 + lane;
    int kind = -1, idx = 0; /* 0 ball 1 robot 2 post 3 cross 4 line cross */
    float radius = PEN_RADIUS, maxDist = maxVis0, ex0 = 0, ex1 = 0, ex2 = 0;
    V2 pt = v2(0, 0);
    if (obj == 0) {
      kind = 0; pt = ball_pos(E); radius = BALL_RADIUS * 2; ex0 = (float)(E.i[RC_S_BALLOWNED] * teamI);
    } else if (obj < 10) {
      idx = obj - 1;
      if (idx < P - 1) {
        kind = 1;
        int r = idx < ag ? idx : idx + 1;
        pt = robot_pos(E, r); radius = ROB_TOTAL; maxDist = maxVis1;
        ex0 = robot_angle(E, r) - headAngle;
        ex1 = (float)(teamI * team_of(E, r));
        ex2 = agFlag; /* the OBSERVER's fallen|penalized flag (reference behaviour) */
      }
    } else if (obj < 33) {
      int mi = obj - 10;
      kind = mi < 4 ? 2 : (mi < 7 ? 3 : 4);
      idx = mi - (mi < 4 ? 0 : (mi < 7 ? 4 : 7));
      float4 mk = *reinterpret_cast<const float4*>(g_marks[mi]);
      pt = v2(mk.x, mk.y); ex0 = mk.z; ex1 = mk.w;
      if (kind == 2) { radius = POST_RADIUS; maxDist = maxVis1; }
      if (kind == 4) ex2 = -headAngle;
    }
    /* isSeenInArea, allowPartial (cutils.py:699-747) */
    int seen = S_NO;
    bool has = false;
    V2 p = v2(0, 0);
    if (kind >= 0) {
      pt = pt - pos;
      float dist1 = cross(vec1, pt), dist2 = cross(vec2, pt);
      if (dist1 < radius && dist2 > -radius) {
        seen = (dist1 < -radius && dist2 > radius) ? (lensq(pt) < maxDist ? S_NORMAL : S_DISTANT) : S_PARTIAL;
        p = rot_cs(pt, hc, -hs);
        has = true;
      }
    }
    if (pass == 1 && !__any_sync(FULLMASK, has)) break;
    if (pass == 0) { /* publish the sightings the interaction tests read */
      robHas = __ballot_sync(FULLMASK, kind == 1 && has);
      if (kind == 1) { E.f[V_ROB + 3 * idx] = p.x; E.f[V_ROB + 3 * idx + 1] = p.y; }
      if (kind == 0) { E.f[V_BALL] = p.x; E.f[V_BALL + 1] = p.y; E.f[V_BALL + 2] = has ? 1.f : 0.f; }
      if (kind == 2) { E.f[V_POST + 3 * idx] = p.x; E.f[V_POST + 3 * idx + 1] = p.y; E.f[V_POST + 3 * idx + 2] = has ? 1.f : 0.f; }
      __syncwarp();
    }
    /* interactions (RoboCupEnvironment.py:1227-1237): occlusion / proximity by the robots in view */
    int inter = I_NO;
    {
      unsigned m = robHas;
      const float R2 = (ROB_TOTAL * 2) * (ROB_TOTAL * 2);
      const float lp = lensq(p);
      while (m) {
        int i = __ffs(m) - 2; /* lane -> robot slot */
        m &= m - 1;
        float ax = E.f[V_ROB + 3 * i], ay = E.f[V_ROB + 3 * i + 1];
        if (has && !(kind == 1 && i == idx)) { /* cutils.doesInteract(robot, object, 35, canOcclude) */
          float dx = ax - p.x, dy = ay - p.y;
          float cr = ax * p.y - ay * p.x, la = ax * ax + ay * ay;
          int t = (dx * dx + dy * dy < R2) ? I_NEARBY : I_NO;
          if (cr * cr < R2 * la && la < lp) t = I_OCCLUDE;
          inter = max(inter, t);
        }
      }
      if (has && kind == 0) { /* ball near a post */
        for (int k = 0; k < 4; k++) {
          float dx = p.x - E.f[V_POST + 3 * k], dy = p.y - E.f[V_POST + 3 * k + 1];
          if (E.f[V_POST + 3 * k + 2] != 0.f && dx * dx + dy * dy < (BALL_RADIUS * 8) * (BALL_RADIUS * 8)) inter = max(inter, I_NEARBY);
        }
      } else if (has && kind == 3) { /* penalty cross near the ball */
        float dx = p.x - E.f[V_BALL], dy = p.y - E.f[V_BALL + 1];
        if (E.f[V_BALL + 2] != 0.f && dx * dx + dy * dy < (BALL_RADIUS * 4) * (BALL_RADIUS * 4)) inter = max(inter, I_NEARBY);
      }
    }
    /* noise, false negatives, misclassification (cutils.py:417-468) */
    if (inter == I_OCCLUDE) seen = S_NO;
    else if (seen != S_NO) {
      const int site = RC_SITE_OBS_BALL + kind;
      Rand6 ru = rng_block(E, tick, site, ag, idx, 0);
      V2 noiseVec = v2((ru.u(0) - 0.5f) * magn, (ru.u(1) - 0.5f) * magn);
      if (!realistic) {
        if (ru.u(2) < randB) seen = S_NO;
        p = p + noiseVec;
        radius *= (1.f - (ru.u(4) - 0.5f) * 0.2f);
        if (kind == 4) ex2 += (ru.u(5) - 0.5f) * magn * 0.1f;
      } else {
        float range = 0.25f + 3.75f * lensq(p) * (1.f / maxDist);
        float mult = range;
        if (inter == I_NEARBY) mult = range * 2;
        if (seen == S_DISTANT) mult = range * 3;
        else if (seen == S_PARTIAL) mult = range * 4;
        /* |newPos| - |p| without cancellation: (2 p.d + d.d) / (|newPos| + |p|) */
        V2 dsp = v2(noiseVec.x * mult / 4, noiseVec.y * mult / 4);
        V2 newPos = p + dsp;
        float lsum = fast_len(newPos) + fast_len(p);
        float diff = lsum > 0 ? __fdividef(2.f * dot(p, dsp) + dot(dsp, dsp), lsum) : 0.f;
        if (ru.u(2) < randB * mult) seen = S_NO;
        if ((kind == 0 || kind == 3) && ru.u(3) < randB * mult / 2) seen = S_MISCLASS;
        p = newPos;
        radius *= 1.f + (ru.u(4) * 0.1f * diff);
        if (kind == 4) ex2 += (ru.u(5) - 0.5f) * magn * mult * (1.f / 180.f);
      }
    }
    /* list membership (RoboCupEnvironment.py:1254-1276) */
    const bool okStrict = seen != S_NO && seen != S_MISCLASS;
    unsigned mBall = __ballot_sync(FULLMASK, kind == 0 && okStrict);
    unsigned mBallMis = __ballot_sync(FULLMASK, kind == 0 && seen == S_MISCLASS);
    unsigned mRob = __ballot_sync(FULLMASK, kind == 1 && seen != S_NO);
    unsigned mPost = __ballot_sync(FULLMASK, kind == 2 && seen != S_NO);
    unsigned mCross = __ballot_sync(FULLMASK, kind == 3 && okStrict);
    unsigned mCrossMis = __ballot_sync(FULLMASK, kind == 3 && seen == S_MISCLASS);
    unsigned mFc = __ballot_sync(FULLMASK, kind == 4 && okStrict);
    /* misclassified balls and crosses change lists (:1258-1263) */
    int lkind = kind;
    float t1 = ex0, t2 = ex1;
    if (seen == S_MISCLASS) {
      if (kind == 0) { /* a ball taken for a cross gets random tags */
        Rand6 rm = rng_block(E, tick, RC_SITE_OBS_MISCROSS, ag, 0, 0);
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
          dst = E.f + (V_STAGE + RC_O_BALL + rrow * RC_F_BALL);
          dst[2] = (radius - BALL_RADIUS * 2) * 2.f; dst[3] = ex0; dst[4] = closest;
        } else {
          rrow = __popc(mRob & lt);
          dst = E.f + (V_STAGE + RC_O_ROBOT + rrow * RC_F_ROBOT);
          float s, c;
          sincos_b(ex0, &s, &c);
          dst[2] = (radius - ROB_TOTAL) * 2.f; dst[3] = c; dst[4] = s; dst[5] = ex1; dst[6] = ex2;
          E.f[V_RAW + 3 * rrow] = p.x; E.f[V_RAW + 3 * rrow + 1] = p.y; E.f[V_RAW + 3 * rrow + 2] = seen == S_NORMAL ? 1.f : 0.f;
        }
        dst[0] = p.x * STDNORM2; dst[1] = p.y * STDNORM2;
      } else { /* posts, crosses, line crosses: polar rows (cutils.py:308-315) */
        int rrow, off, feat = 6;
        if (lkind == 2) { rrow = __popc(mPost & lt); off = RC_O_POST; }
        else if (lkind == 3) { rrow = kind == 0 ? __popc(mCross) : __popc(mCross & lt); off = RC_O_CROSS; }
        else { rrow = nf + __popc(mFc & lt); off = RC_O_FCROSS; feat = RC_F_FCROSS; }
        float* dst = E.f + (V_STAGE + off + rrow * feat);
        float l2 = lensq(p);
        float inv = l2 > 0 ? rsqrtf(l2) : 0.f;
        float dist = l2 * inv;
        dst[0] = (dist * STDNORM - 0.5f) * 2.f;
        dst[1] = dist > 0 ? p.x * team * inv : 1.f;
        dst[2] = dist > 0 ? p.y * team * inv : 0.f;
        dst[3] = (radius - (lkind == 2 ? POST_RADIUS : PEN_RADIUS)) * 2.f;
        dst[4] = t1 * team; dst[5] = t2 * team;
        if (lkind == 4) {
          float s, c;
          sincos_b(ex2, &s, &c);
          dst[6] = c; dst[7] = -s;
        }
      }
    }
    if (pass == 0) {
      nb = __popc(mBall) + __popc(mCrossMis);
      nr = __popc(mRob);
      np = __popc(mPost);
      ncr = __popc(mCross) + __popc(mBallMis);
      /* seen flags (:1254-1255): robots by slot, ball strictly seen */
      if (kind == 1) E.i[V_STAGE + RC_O_ROBSEEN + idx] = seen != S_NO ? 1 : 0;
      if (lane == 0) E.i[V_STAGE + RC_O_BALLSEEN] = okStrict ? 1 : 0;
    }
    nf += __popc(mFc);
  }
  /* one Philox evaluation serves three consumers: lanes 0..10 the noise of line `lane`, lanes 11..20 the
     test draw of false-positive loop index lane - 11, lanes 21..31 the false-positive ball of robot row
     lane - 21 (every draw is addressed by its consumer, include/dynenv_rng.h) */
  const Rand6 rmix = rng_block(E, tick, lane < 11 ? RC_SITE_OBS_LINE : (lane < 21 ? RC_SITE_OBS_FP : RC_SITE_OBS_FPBALL), ag,
                               lane < 11 ? lane : (lane < 21 ? lane - 11 : lane - 21), 0);
  /* lines: isLineInArea (cutils.py:751-821), addNoiseLine (:352-378), normalizeLine (:333-347) */
  int nl;
  {
    int seen = S_NO;
    V2 lp1 = v2(0, 0), ldir = v2(0, 0);
    float tx = 0, ty = 0;
    if (lane < 11) {
      const float* ln = g_lines[lane];
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
        /* lp2 - lp1 from the line's own direction: a clipped piece can be a fraction of a pixel long,
           which the difference of two float32 end points (~300 px) cannot resolve */
        ldir = rot_cs(D, hc, -hs) * (t2 - t1);
        if (lp1.x < 0 || lp2.x < 0) seen = S_NO;
        if (seen != S_NO) {
          const Rand6& ru = rmix;
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
    unsigned mLine = __ballot_sync(FULLMASK, seen != S_NO);
    nl = __popc(mLine);
    if (seen != S_NO) {
      float* dst = E.f + (V_STAGE + RC_O_LINE + __popc(mLine & lt) * RC_F_LINE);
      float ld = fast_len(ldir); /* ldir = lp2 - lp1; |lp2 x lp1| = |ldir x lp1| */
      float ild = ld > 0 ? __fdividef(1.f, ld) : 0.f;
      dst[0] = (__fdividef(fabsf(cross(ldir, lp1)), ld + 1e-7f) * STDNORM - 0.5f) * 2.f;
      dst[1] = ld > 0 ? ldir.x * ild : 1.f;
      dst[2] = ld > 0 ? ldir.y * ild : 0.f;
      dst[3] = tx; dst[4] = ty;
      if (RAW) { /* the end points the image branch draws (RoboCupEnvironment.py:1324-1341) */
        float* r = E.f + (V_RAWREC + RC_R_LINE + __popc(mLine & lt) * 4);
        r[0] = lp1.x; r[1] = lp1.y; r[2] = lp1.x + ldir.x; r[3] = lp1.y + ldir.y;
      }
    }
  }
  if (RAW) {
    /* the centre circle: isSeenInArea without partial sightings (cutils.py:699-747) and addNoise (:417-468); the
       PARTIAL observation does not contain it, the image does (RoboCupEnvironment.py:1224,1251,1343-1393) */
    if (lane == 31) {
      const float radius0 = CC_RADIUS * 2;
      V2 pt = v2(520.f, 370.f) - pos;
      const float dist1 = cross(vec1, pt), dist2 = cross(vec2, pt);
      int seen = S_NO;
      float radius = radius0;
      V2 p = v2(0, 0);
      if (dist1 < radius0 && dist2 > -radius0) {
        if (dist1 < -radius0 && dist2 > radius0) seen = lensq(pt) < maxVis1 ? S_NORMAL : S_DISTANT;
        else {
          float b1 = -2.f * (vec1.x * pt.x - vec1.y * pt.y), b2 = -2.f * (vec2.x * pt.x - vec2.y * pt.y);
          float c = dot(pt, pt) - radius0 * radius0;
          float sqr1 = b1 * b1 - 4.f * c, sqr2 = b2 * b2 - 4.f * c;
          bool found = false;
          if (sqr1 >= 0) {
            float sq = sqrtf(sqr1);
            if ((-b1 + sq) > 0 || (-b1 - sq) > 0) { seen = S_PARTIAL; found = true; }
          }
          if (!found && sqr2 >= 0) {
            float sq = sqrtf(sqr2);
            if ((-b2 + sq) > 0 || (-b2 - sq) > 0) seen = S_PARTIAL;
          }
        }
        p = rot_cs(pt, hc, -hs);
      }
      if (seen != S_NO) {
        Rand6 ru = rng_block(E, tick, RC_SITE_OBS_CIRCLE, ag, 0, 0);
        V2 noiseVec = v2((ru.u(0) - 0.5f) * magn, (ru.u(1) - 0.5f) * magn);
        if (!realistic) {
          if (ru.u(2) < randB) seen = S_NO;
          p = p + noiseVec;
          radius *= (1.f - (ru.u(4) - 0.5f) * 0.2f);
        } else {
          float range = 0.25f + 3.75f * lensq(p) * (1.f / maxVis1);
          float mult = range;
          if (seen == S_DISTANT) mult = range * 3;
          else if (seen == S_PARTIAL) mult = range * 4;
          V2 dsp = v2(noiseVec.x * mult / 4, noiseVec.y * mult / 4);
          V2 newPos = p + dsp;
          float lsum = len(newPos) + len(p);
          float diff = lsum > 0 ? (2.f * dot(p, dsp) + dot(dsp, dsp)) / lsum : 0.f;
          if (ru.u(2) < randB * mult) seen = S_NO;
          p = newPos;
          radius *= 1.f + (ru.u(4) * 0.1f * diff);
        }
      }
      E.i[V_RAWREC + RC_R_CIRCLE_SEEN] = seen != S_NO ? 1 : 0;
      E.f[V_RAWREC + RC_R_CIRCLE] = seen != S_NO ? p.x : 0.f;
      E.f[V_RAWREC + RC_R_CIRCLE + 1] = seen != S_NO ? p.y : 0.f;
      E.f[V_RAWREC + RC_R_CIRCLE + 2] = seen != S_NO ? radius : 0.f;
    }
  }
  if (lane == 0) {
    E.i[V_STAGE + RC_O_NLANDMARKS] = nf + nl + ncr + np;
    E.i[V_CNT + 0] = nb; E.i[V_CNT + 1] = nr; E.i[V_CNT + 2] = np;
    E.i[V_CNT + 3] = ncr; E.i[V_CNT + 4] = nf; E.i[V_CNT + 5] = nl;
  }
  __syncwarp();
  /* random false positives (:1279-1305): ten draws, lane = loop index */
  {
    unsigned hits = __ballot_sync(FULLMASK, lane >= 11 && lane < 21 && rmix.u(0) < randB) >> 11;
    if (hits) false_positives(E, ag, tick, hits, team);
  }
  /* false-positive balls near robots (:1308-1316): lane 21 + k = row k of the robot list (a 12th row, which
     needs three false-positive robots, takes a second round with its own draw) */
  if (realistic) {
    const int nrNow = E.i[V_CNT + 1];
    for (int row0 = 0; row0 < nrNow; row0 += 11) {
      const int row = row0 + lane - 21;
      bool create = false;
      float bx = 0, by = 0, brad = 0;
      if (lane >= 21 && row < nrNow && E.f[V_RAW + 3 * row + 2] != 0.f) {
        Rand6 ru = rmix;
        if (row0) ru = rng_block(E, tick, RC_SITE_OBS_FPBALL, ag, row, 0);
        V2 rp = v2(E.f[V_RAW + 3 * row], E.f[V_RAW + 3 * row + 1]);
        if (ru.u(0) < randB * 10 && lensq(rp) < 250.f * 250.f) {
          create = true;
          bx = rp.x + (2 * ru.u(2) - 1.f) * ROB_TOTAL;
          by = rp.y + (2 * ru.u(3) - 1.f) * ROB_TOTAL;
          brad = BALL_RADIUS * 2 * (1.f - 0.4f * (ru.u(4) - 0.5f));
        }
      }
      unsigned cm = __ballot_sync(FULLMASK, create);
      if (cm) {
        int nbNow = E.i[V_CNT + 0];
        __syncwarp();
        int at = nbNow + __popc(cm & lt);
        if (create && at < RC_CAP_BALL) {
          float* dst = E.f + (V_STAGE + RC_O_BALL + at * RC_F_BALL);
          dst[0] = bx * STDNORM2; dst[1] = by * STDNORM2; dst[2] = (brad - BALL_RADIUS * 2) * 2.f; dst[3] = 0.f; dst[4] = closest;
        }
        if (lane == 0) E.i[V_CNT + 0] = min(nbNow + __popc(cm), (int)RC_CAP_BALL);
        __syncwarp();
      }
    }
  }
  if (lane < 6) E.i[V_STAGE + RC_O_COUNTS + lane] = E.i[V_CNT + lane];
  /* running sums for processSeens */
  if (lane == 0) {
    E.i[X_LSUM + ag] += E.i[V_STAGE + RC_O_NLANDMARKS];
    E.i[X_BSUM + ag] += E.i[V_STAGE + RC_O_BALLSEEN];
  }
  if (lane >= 1 && lane < 10) E.i[X_RSEEN + 9 * ag + lane - 1] += E.i[V_STAGE + RC_O_ROBSEEN + lane - 1];
  __syncwarp();
  if (RAW) {
    /* positions and sizes of the point detections back from the encoded rows (the encoders are affine / polar maps,
       RoboCupEnvironment.py:1537-1548, cutils.py:295-315): row k of every list, false positives included */
    for (int slot = lane; slot < RC_CAP_BALL + RC_CAP_ROBOT + RC_CAP_POST + RC_CAP_CROSS + RC_CAP_FCROSS; slot += 32) {
      int type, k = slot;
      if (k < RC_CAP_BALL) type = 0;
      else if ((k -= RC_CAP_BALL) < RC_CAP_ROBOT) type = 1;
      else if ((k -= RC_CAP_ROBOT) < RC_CAP_POST) type = 2;
      else if ((k -= RC_CAP_POST) < RC_CAP_CROSS) type = 3;
      else { k -= RC_CAP_CROSS; type = 4; }
      const int off = type == 0 ? RC_O_BALL : (type == 1 ? RC_O_ROBOT : (type == 2 ? RC_O_POST : (type == 3 ? RC_O_CROSS : RC_O_FCROSS)));
      const int feat = type == 0 ? RC_F_BALL : (type == 1 ? RC_F_ROBOT : (type == 4 ? RC_F_FCROSS : 6));
      const int roff = type == 0 ? RC_R_BALL : (type == 1 ? RC_R_ROBOT : (type == 2 ? RC_R_POST : (type == 3 ? RC_R_CROSS : RC_R_FCROSS)));
      const float* row = E.f + (V_STAGE + off + k * feat);
      float x = 0.f, y = 0.f, r = 0.f;
      if (k < E.i[V_CNT + type]) {
        if (type <= 1) {
          x = row[0] * (1.f / STDNORM2); y = row[1] * (1.f / STDNORM2);
          r = row[2] * 0.5f + (type == 0 ? BALL_RADIUS * 2 : ROB_TOTAL);
        } else {
          float dist = (row[0] * 0.5f + 0.5f) * (1.f / STDNORM);
          x = row[1] * team * dist; y = row[2] * team * dist;
          r = row[3] * 0.5f + (type == 2 ? POST_RADIUS : PEN_RADIUS);
        }
      }
      float* dst = E.f + (V_RAWREC + roff + 3 * k);
      dst[0] = x; dst[1] = y; dst[2] = r;
    }
    if (lane < RC_NTYPES) E.i[V_RAWREC + RC_R_COUNTS + lane] = E.i[V_CNT + lane];
    {
      const int nl = E.i[V_CNT + 5]; /* unused line rows stay zero */
      for (int w = lane; w < RC_CAP_LINE * 4; w += 32)
        if (w >= nl * 4) E.f[V_RAWREC + RC_R_LINE + w] = 0.f;
      for (int w = RC_RAW_WORDS_RAW + lane; w < RC_RAW_WORDS; w += 32) E.f[V_RAWREC + w] = 0.f;
    }
    __syncwarp();
  }
}

/* one observation tick of every agent; records go straight to HBM with coalesced stores */
template <bool RAW>
__device__ __noinline__ void observe_tick(const Env E, int t, uint32_t tick, float* obs_env, float* full_env, float* raw_env = nullptr) {
  const int P = E.c->P;
  const int lane = E.lane;
  if (E.c->obs_type == DYNENV_OBS_FULL) {
    if (full_env) {
      for (int ag = 0; ag < P; ag++) full_state_agent(E, ag, full_env + ((size_t)t * P + ag) * RC_FULL_WORDS);
    }
    return;
  }
  for (int ag = 0; ag < P; ag++) {
    agent_vision<RAW>(E, ag, tick);
    if (RAW && raw_env) {
      float4* dst = reinterpret_cast<float4*>(raw_env + ((size_t)t * P + ag) * RC_RAW_WORDS);
      const float4* src = reinterpret_cast<const float4*>(E.f + V_RAWREC);
      for (int w = lane; w < RC_RAW_WORDS / 4; w += 32) dst[w] = src[w];
    }
    if (obs_env) {
      float4* dst = reinterpret_cast<float4*>(obs_env + ((size_t)t * P + ag) * RC_OBS_WORDS);
      const float4* src = reinterpret_cast<const float4*>(E.f + V_STAGE);
      for (int w = lane; w < RC_OBS_WORDS / 4; w += 32) dst[w] = src[w];
    }
    __syncwarp();
  }
}

/* ================================================================= reset (scene set-up) */
__device__ __noinline__ void init_scene(const Env E) {
  /* RoboCupEnvironment.__init__ + _setup_scene (RoboCupEnvironment.py:42-63,239-336); lane 0 draws the
     scene, every lane clears the record first */
  const int lane = E.lane;
  int episode = E.i[RC_S_EPISODE] + 1;
  __syncwarp();
  for (int w = lane; w < RC_STATE_WORDS; w += 32) E.i[w] = 0;
  __syncwarp();
  if (lane == 0) {
    const SimCfg& c = *E.c;
    E.i[RC_S_EPISODE] = episode;
    E.i[RC_S_BALLOWNED] = 1;
    E.f[RC_S_BALLFREE] = 9999.f;
    E.i[RC_S_PENALTIMES] = 20000;
    E.i[RC_S_PENALTIMES + 1] = 20000;
    int k = 0;
    Rand6 blk;
    auto U = [&]() {
      if (k % 6 == 0) blk = rng_block(E, DYNENV_TICK_RESET, RC_SITE_RESET, 0, 0, k / 6);
      float u = blk.u(k % 6);
      k++;
      return u;
    };
    float sx[2][5], sy[2][5];
    if (c.random_init) {
      const float xL[7] = {SIDE + 10, SIDE + 50, SIDE + 250, SIDE + 450, SIDE + 650, SIDE + 850, SIDE + 890};
      const float yL[3] = {SIDE + 20, SIDE + 300, SIDE + 580};
      float rx[10], ry[10];
      int n = 0;
      for (int i = 0; i < 6; i++) {
        bool edge = (i == 0 || i == 5);
        int ny = edge ? 1 : 2;
        for (int j = 0; j < ny; j++) {
          float yB = edge ? yL[0] : yL[j], yE = edge ? yL[2] : yL[j + 1];
          float ux = U(), uy = U();
          rx[n] = xL[i] + ux * (xL[i + 1] - xL[i]);
          ry[n] = yB + uy * (yE - yB);
          n++;
        }
      }
      sx[0][0] = rx[0]; sy[0][0] = ry[0];
      sx[1][0] = rx[9]; sy[1][0] = ry[9];
      int perm[8];
      for (int i = 0; i < 8; i++) perm[i] = i;
      for (int i = 7; i > 0; i--) { int j = (int)(U() * (float)(i + 1)); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
      for (int i = 0; i < 8; i++) { sx[i < 4 ? 0 : 1][1 + (i & 3)] = rx[perm[i] + 1]; sy[i < 4 ? 0 : 1][1 + (i & 3)] = ry[perm[i] + 1]; }
    } else {
      const float cX = W_ / 2;
      float u0, u1;
      u0 = U(); u1 = U(); sx[0][0] = cX - (BALL_RADIUS * 2 + ROB_TOTAL) - u0 * 50; sy[0][0] = H_ / 2 + (u1 - 0.5f) * 25;
      u0 = U(); u1 = U(); sx[0][1] = cX - (ROB_TOTAL + LINEW * 2) - u0 * 50; sy[0][1] = SIDE + FIELD_H / 4 + (u1 - 0.5f) * 50;
      u0 = U(); u1 = U(); sx[0][2] = cX - (ROB_TOTAL + LINEW * 2) - u0 * 50; sy[0][2] = SIDE + 3 * FIELD_H / 4 + (u1 - 0.5f) * 50;
      u0 = U(); u1 = U(); sx[0][3] = cX - (FIELD_W / 4) - (u0 - 0.5f) * 50; sy[0][3] = SIDE + FIELD_H / 2 + (u1 - 0.5f) * 50;
      u0 = U(); sx[0][4] = SIDE + 20; sy[0][4] = H_ / 2 + (u0 - 0.5f) * 50;
      u0 = U(); u1 = U(); sx[1][0] = cX + (CC_RADIUS * 2 + ROB_TOTAL + LINEW / 2) + u0 * 50; sy[1][0] = H_ / 2 + (u1 - 0.5f) * 50;
      u0 = U(); u1 = U(); sx[1][1] = cX + (ROB_TOTAL + LINEW / 2 + CC_RADIUS) + u0 * 50; sy[1][1] = SIDE + FIELD_H / 4 + (u1 - 0.5f) * 50;
      u0 = U(); u1 = U(); sx[1][2] = cX + (ROB_TOTAL + LINEW / 2 + CC_RADIUS) + u0 * 50; sy[1][2] = SIDE + 3 * FIELD_H / 4 + (u1 - 0.5f) * 50;
      u0 = U(); u1 = U(); sx[1][3] = cX + (SIDE + FIELD_W / 4) + u0 * 50; sy[1][3] = SIDE + FIELD_H / 2 + (u1 - 0.5f) * 50;
      u0 = U(); sx[1][4] = W_ - (SIDE + 20); sy[1][4] = H_ / 2 + (u0 - 0.5f) * 50;
    }
    int ids[2][5];
    for (int t = 0; t < 2; t++) { /* np.random.permutation(maxPlayers) */
      for (int i = 0; i < 5; i++) ids[t][i] = i;
      for (int i = 4; i > 0; i--) { int j = (int)(U() * (float)(i + 1)); int tmp = ids[t][i]; ids[t][i] = ids[t][j]; ids[t][j] = tmp; }
    }
    for (int r = 0; r < c.P; r++) {
      int t = r < c.n_players ? 0 : 1;
      int i = r - t * c.n_players;
      float px = sx[t][ids[t][i]], py = sy[t][ids[t][i]];
      float ang = t == 0 ? 0.f : PI_F;
      for (int f = 0; f < 2; f++) {
        int b = 2 * r + f;
        E.f[RC_S_PX + b] = px; E.f[RC_S_PY + b] = py; E.f[RC_S_ANG + b] = ang;
        shape_cache_store(E, b);
      }
      E.f[RC_S_PREVX + r] = px; E.f[RC_S_PREVY + r] = py;
      if (c.deterministic_turn) E.f[RC_S_HEAD + r] = (t == 0 ? 1.f : -1.f) * HEAD_MAX;
    }
    float bx = 520.f, by = 370.f;
    if (c.random_init) {
      bx = U() * FIELD_W + SIDE;
      by = U() * FIELD_H + SIDE;
      int owned = U() > 0.4f ? 1 : 0;
      if (owned != 0 && U() > 0.5f) owned = -owned;
      E.i[RC_S_BALLOWNED] = owned;
    }
    E.f[RC_S_PX + BALL] = bx; E.f[RC_S_PY + BALL] = by;
    shape_cache_store(E, BALL);
    E.f[RC_S_BPREVX] = bx; E.f[RC_S_BPREVY] = by;
  }
  __syncwarp();
}

/* info['Full State'] + info['Recon States'] (RoboCupEnvironment.py:511-512,1151-1161) */
__device__ __noinline__ void fill_info(const Env E, float* info_env) {
  const int lane = E.lane;
  const int P = E.c->P;
  if (lane < P) {
    V2 p = robot_pos(E, lane);
    float s, c;
    sincos_b(robot_angle(E, lane), &s, &c);
    float* o = info_env + RC_I_STATE_ROBOTS + 6 * lane;
    o[0] = p.x * STDNORM2; o[1] = p.y * STDNORM2; o[2] = c; o[3] = s;
    o[4] = (float)team_of(E, lane);
    o[5] = (E.i[RC_S_RFLAGS + lane] & (RC_RF_FALLEN | RC_RF_PENALIZED)) ? 1.f : 0.f;
  }
  if (lane == 0) {
    V2 bp = ball_pos(E);
    info_env[RC_I_STATE_BALL] = bp.x * STDNORM2;
    info_env[RC_I_STATE_BALL + 1] = bp.y * STDNORM2;
    info_env[RC_I_STATE_BALL + 2] = (float)E.i[RC_S_BALLOWNED];
  }
  for (int ag = 0; ag < P; ag++) full_state_agent(E, ag, info_env + RC_I_RECON + ag * RC_FULL_WORDS);
}

__device__ __forceinline__ void clear_step_scratch(const Env E) {
  for (int w = RC_S_STEP_REW + E.lane; w < RC_S_STEP_END; w += 32) E.i[w] = 0;
  __syncwarp();
}

__device__ __forceinline__ void finish_step(const Env E, float* info_env) {
  const int lane = E.lane;
  const int P = E.c->P;
  /* processSeens (:1563-1575) and reward assembly (:499-509) */
  if (lane < P) {
    float obsRew = 0.f;
    if (E.c->obs_type != DYNENV_OBS_FULL && E.c->use_obs_rewards) {
      float lSeens = fmaxf(0.f, fminf(3.f, (float)E.i[X_LSUM + lane] / (float)RC_NTICK));
      float rs = 0.f;
      for (int k = 0; k < P - 1; k++) rs += fmaxf(0.f, fminf(2.f, (float)E.i[X_RSEEN + 9 * lane + k]));
      float bSeens = fmaxf(0.f, fminf(3.f, (float)E.i[X_BSUM + lane]));
      obsRew = 0.0025f * (rs + lSeens) + 0.01f * bSeens;
    }
    float teamRew = E.f[X_TEAMREW + (lane < E.c->n_players ? 0 : 1)];
    float rew = E.f[X_REW + lane] + teamRew + obsRew;
    float posRew = E.f[X_POSREW + lane] + fmaxf(0.f, teamRew) + fmaxf(0.f, obsRew);
    E.f[RC_S_EPR + lane] += rew;
    E.f[RC_S_EPPR + lane] += posRew;
    E.f[RC_S_EPOR + lane] += obsRew;
    if (info_env) info_env[RC_I_REWARDS + lane] = rew;
  } else if (lane < RC_MAXP && info_env) info_env[RC_I_REWARDS + lane] = 0.f;
  __syncwarp();
  if (info_env) {
    fill_info(E, info_env);
    bool done = E.i[RC_S_ELAPSED] >= MAXTIME;
    if (lane == 0) {
      reinterpret_cast<int*>(info_env)[RC_I_DONE] = done ? 1 : 0;
      reinterpret_cast<int*>(info_env)[RC_I_FLAGS] = E.i[X_DIAG];
      reinterpret_cast<int*>(info_env)[RC_I_EPG] = done ? E.i[RC_S_GOALS] : 0;
      reinterpret_cast<int*>(info_env)[RC_I_EPG + 1] = done ? E.i[RC_S_GOALS + 1] : 0;
    }
    if (lane < RC_MAXP) {
      bool v = done && lane < P;
      info_env[RC_I_EPR + lane] = v ? E.f[RC_S_EPR + lane] : 0.f;
      info_env[RC_I_EPPR + lane] = v ? E.f[RC_S_EPPR + lane] : 0.f;
      info_env[RC_I_EPOR + lane] = v ? E.f[RC_S_EPOR + lane] : 0.f;
    }
  }
  __syncwarp();
}

/* environment_base.reset: new scene + nTimeSteps copies of the initial observation (:205-224) */
template <bool RAW>
__device__ __noinline__ void reset_env(const Env E, float* obs_env, float* full_env, float* info_env, bool write_info, float* raw_env = nullptr) {
  init_scene(E);
  clear_step_scratch(E);
  for (int t = 0; t < RC_NTICK; t++) observe_tick<RAW>(E, t, DYNENV_TICK_RESETOBS | (uint32_t)t, obs_env, full_env, raw_env);
  if (info_env && write_info) {
    const int lane = E.lane;
    fill_info(E, info_env);
    if (lane < RC_MAXP) {
      info_env[RC_I_REWARDS + lane] = 0.f;
      info_env[RC_I_EPR + lane] = 0.f; info_env[RC_I_EPPR + lane] = 0.f; info_env[RC_I_EPOR + lane] = 0.f;
    }
    if (lane == 0) {
      int* ii = reinterpret_cast<int*>(info_env);
      ii[RC_I_DONE] = 0; ii[RC_I_FLAGS] = 0; ii[RC_I_EPG] = 0; ii[RC_I_EPG + 1] = 0;
    }
  }
  __syncwarp();
}

}  // namespace rc
}  // namespace dynenv
