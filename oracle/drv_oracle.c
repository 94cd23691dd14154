/*
 * ORACLE / TEST INFRASTRUCTURE ONLY — CPU restatement of DynEnv's Driving rollout hot path.
 *
 * Sequential, double precision, one environment per DrvEnv; every block cites the reference file:line it
 * restates (paths under DynEnv/).  The rigid-body part restates the Chipmunk2D 7.0.x subset that pymunk 5.x
 * exposes to the reference (box-box and circle-box narrow phase, arbiter cache with begin callbacks,
 * sequential-impulse contacts); that dependency is NOT in /root/reference: PARITY UNPINNED at the pymunk
 * boundary.  Everything above it (rules, rewards, vision, noise, encoders) is pinned against the reference's
 * own Python run on oracle/shim (tests/golden/driving_*.npz, tests/test_oracle_golden.py).
 *
 * Deviation from the reference, on purpose: continuousActions=True raises
 * UnboundLocalError in the reference (DrivingEnvironment.py:360-368); here acc = action[0], steer = action[1]
 * as its README describes.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may call this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dynenv_b200.h"
#include "../include/dynenv_rng.h"
#include "oracle_rng.h"

#define PI 3.14159265358979323846
#define W_ 1700.0
#define H_ 1000.0
#define MAXTIME 6000
#define NCAR DRV_MAXCARS
#define NPEDMAX DRV_MAXPED
#define NOBSMAX DRV_MAXOBS
#define ANGLE_NOISE (PI / 180.0) /* cutils.py:203 */
#define LISTCAP 40

typedef struct { double x, y; } V2;
static inline V2 v2(double x, double y) { V2 r = {x, y}; return r; }
static inline V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 vsub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 vmul(V2 a, double s) { return v2(a.x * s, a.y * s); }
static inline double vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline double vcross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline double vlen(V2 a) { return sqrt(a.x * a.x + a.y * a.y); }
static inline double vlensq(V2 a) { return a.x * a.x + a.y * a.y; }
static inline V2 vrot(V2 a, double ang) {
  double c = cos(ang), s = sin(ang);
  return v2(a.x * c - a.y * s, a.x * s + a.y * c);
}

/* Car.py:9-13 */
static const double CAR_MASS[4] = {1200, 1800, 3500, 5000};
static const double CAR_WIDTH[4] = {5, 6, 7, 8};
static const double CAR_LENGTH[4] = {10, 15, 20, 25};
static const double CAR_POWER[4] = {3, 4, 3, 4};

typedef struct {
  V2 p0, p1, dir, normal;
  double length, width, angle;
  int nLanes;
} Road;

typedef struct {
  double px, py, a, vx, vy, w, vbx, vby, wb;
  double m, m_inv, i_inv;
} Body;

typedef struct {
  Body b;
  V2 dir, prev, goal;
  int type, team, finished, crashed, lane;
  double width, height; /* half extents: width along local y, height along local x */
} Car;

typedef struct {
  Body b;
  double moving;
  int speed, dead, crossing, beginCrossing, side, road;
} Ped;

typedef struct {
  V2 r1, r2;
  int hash;
  double jnAcc, jtAcc, jBias, nMass, tMass, bias, bounce;
} Contact;

typedef struct {
  int a, b, state, age, ncon, touched;
  Contact con[2];
  V2 n;
  double e, u;
} Arbiter;

typedef struct {
  double self[DRV_F_SELF];
  double car[DRV_CAP_CAR][DRV_F_CAR];
  double obs[DRV_CAP_OBS][DRV_F_OBS];
  double ped[DRV_CAP_PED][DRV_F_PED];
  double lane[DRV_CAP_LANE][DRV_F_LANE_FULL];
  int32_t counts[4];
} DrvAgentObs;

typedef struct DrvStepOut {
  DrvAgentObs agent[NCAR];
  double rewards[NCAR];
  int32_t done;
  double episode_r[NCAR], episode_p_r[NCAR];
  int32_t episode_g[2];
  double state_cars[NCAR][7];
  double state_obs[NOBSMAX][4];
  double state_ped[NPEDMAX][2];
  int32_t nobs, nped;
} DrvStepOut;

struct DrvEnv {
  int N, obsType, noiseType, continuous;
  double noiseMag, randBase;
  OracleRng rng;
  Road roads[2];
  Car car[NCAR];
  Ped ped[NPEDMAX];
  V2 obst[NOBSMAX];
  int elapsed, episode, allFinished, nPed, nObs, stepped;
  Arbiter arb[DRV_MAXARB];
  int nArb, active[DRV_MAXARB], nActive, arbOverflow;
  double epR[NCAR], epPR[NCAR];
  double carRewards[NCAR], carPosRewards[NCAR], teamReward;
  int subBegin, subEnd;
};
typedef struct DrvEnv DrvEnv;

static const V2 BUILD_POS[4] = {{365.0, 200.0}, {365.0, 800.0}, {1385.0, 200.0}, {1385.0, 800.0}}; /* :100-105 */
#define BUILD_W 400.0
#define BUILD_H 225.0
#define OBST_HALF 10.0
#define PED_R 5.0 /* Pedestrian.py:16: radius * 2 */

static inline double rnd(DrvEnv* e, uint32_t tick, int site, int a, int b, int k) {
  return oracle_uniform(&e->rng, (uint32_t)e->episode, tick, site, a, b, k);
}
static inline int rnd_int(DrvEnv* e, uint32_t tick, int site, int a, int b, int k, int lo, int hi) {
  return lo + (int)(rnd(e, tick, site, a, b, k) * (hi - lo + 1));
}

/* ---------------------------------------------------------------- roads (Road.py) */
static void road_init(Road* r, int nLanes, double width, V2 p0, V2 p1) {
  /* Road.py:11-34 */
  r->nLanes = nLanes; r->width = width; r->p0 = p0; r->p1 = p1;
  V2 d = vsub(p1, p0);
  r->length = vlen(d);
  r->dir = v2(d.x / r->length, d.y / r->length);
  r->normal = vrot(r->dir, PI / 2);
  r->angle = atan2(r->dir.y, r->dir.x);
}
static int is_point_on_road(const Road* r, V2 point, double angle) {
  /* Road.py:74-97; LanePosition: AtGoal 0, InRightLane 1, InOpposingLane 2, OverRoad 3, OffRoad 4 */
  V2 pt = vsub(point, r->p0);
  double dist = vcross(r->dir, pt);
  if (fabs(dist) >= r->nLanes * r->width + 5) return 4;
  int pos = 3;
  double dirDist = vdot(r->dir, pt);
  if (dirDist >= -10 && dirDist <= r->length + 10) {
    double relAngle = cos(r->angle - angle) * dist;
    pos = relAngle < 0 ? 1 : 2;
  }
  return pos;
}
static int is_off_road(const DrvEnv* e, V2 p) { /* DrivingEnvironment.py:509-520 */
  int pos = 4;
  for (int k = 0; k < 2; k++) { int q = is_point_on_road(&e->roads[k], p, 0); if (q < pos) pos = q; }
  return pos >= 3;
}
static int is_out(V2 p) { return p.x <= 0 || p.y <= 0 || p.x >= W_ || p.y >= H_; } /* :523-524 */

/* ---------------------------------------------------------------- bodies and shapes */
static void body_set_mass(Body* b, double m, double inertia) {
  b->m = m; b->m_inv = 1.0 / m; b->i_inv = 1.0 / inertia;
}
static Body* shape_body(DrvEnv* e, int s, Body* stat) {
  if (s >= DRV_SHAPE_CAR0 && s < DRV_SHAPE_PED0) return &e->car[s - DRV_SHAPE_CAR0].b;
  if (s >= DRV_SHAPE_PED0 && s < DRV_SHAPE_OBS0) return &e->ped[s - DRV_SHAPE_PED0].b;
  memset(stat, 0, sizeof(*stat));
  V2 p = s < DRV_NBUILD ? BUILD_POS[s] : e->obst[s - DRV_SHAPE_OBS0];
  stat->px = p.x; stat->py = p.y; stat->m = INFINITY;
  return stat;
}
static int is_car(int s) { return s >= DRV_SHAPE_CAR0 && s < DRV_SHAPE_PED0; }
static int is_ped(int s) { return s >= DRV_SHAPE_PED0 && s < DRV_SHAPE_OBS0; }

/* world vertices (CCW: (-h,-w), (h,-w), (h,w), (-h,w)) and plane normals (plane i: edge v[i-1] -> v[i]) of a box */
typedef struct { V2 v[4], n[4]; double l, b, r, t; } Box;
static void box_world(DrvEnv* e, int s, Box* bx) {
  double hx, hy, px, py, ang = 0;
  if (is_car(s)) {
    Car* c = &e->car[s - DRV_SHAPE_CAR0];
    hx = c->height; hy = c->width; px = c->b.px; py = c->b.py; ang = c->b.a;
  } else if (s < DRV_NBUILD) { hx = BUILD_W; hy = BUILD_H; px = BUILD_POS[s].x; py = BUILD_POS[s].y; }
  else { hx = OBST_HALF; hy = OBST_HALF; px = e->obst[s - DRV_SHAPE_OBS0].x; py = e->obst[s - DRV_SHAPE_OBS0].y; }
  const double lx[4] = {-hx, hx, hx, -hx}, ly[4] = {-hy, -hy, hy, hy};
  const double nx[4] = {-1, 0, 1, 0}, ny[4] = {0, -1, 0, 1};
  double c = cos(ang), sn = sin(ang);
  bx->l = bx->b = INFINITY; bx->r = bx->t = -INFINITY;
  for (int i = 0; i < 4; i++) {
    bx->v[i] = v2(px + lx[i] * c - ly[i] * sn, py + lx[i] * sn + ly[i] * c);
    bx->n[i] = v2(nx[i] * c - ny[i] * sn, nx[i] * sn + ny[i] * c);
    bx->l = fmin(bx->l, bx->v[i].x); bx->r = fmax(bx->r, bx->v[i].x);
    bx->b = fmin(bx->b, bx->v[i].y); bx->t = fmax(bx->t, bx->v[i].y);
  }
}

static V2 closest_on_seg(V2 p, V2 a, V2 b) {
  V2 d = vsub(b, a);
  double t = vdot(d, vsub(p, a)) / vlensq(d);
  t = t < 0 ? 0 : (t > 1 ? 1 : t);
  return vadd(a, vmul(d, t));
}
static inline double clamp01(double t) { return t < 0 ? 0 : (t > 1 ? 1 : t); }
static inline V2 vlerp(V2 a, V2 b, double t) { return vadd(vmul(a, 1.0 - t), vmul(b, t)); }

/* SupportEdgeForPoly: the edge of the box that faces direction n; vertex indices are the feature hashes */
static void support_edge(const Box* s, V2 n, V2* pa, int* ha, V2* pb, int* hb) {
  int i1 = 0;
  double best = -INFINITY;
  for (int i = 0; i < 4; i++) { double d = vdot(s->v[i], n); if (d > best) { best = d; i1 = i; } }
  int i0 = (i1 + 3) % 4, i2 = (i1 + 1) % 4;
  if (vdot(n, s->n[i1]) > vdot(n, s->n[i2])) { *pa = s->v[i0]; *ha = i0; *pb = s->v[i1]; *hb = i1; }
  else { *pa = s->v[i1]; *ha = i1; *pb = s->v[i2]; *hb = i2; }
}

/* box - box: minimum-penetration axis among both boxes' edge normals (what GJK/EPA returns for overlapping
   convex polygons; separated polygons of radius 0 make no contact), then ContactPoints (cpCollision.c) */
static int collide_box_box(const Box* A, const Box* B, V2* nOut, Contact* out) {
  double best = -INFINITY;
  V2 bn = v2(1, 0);
  for (int i = 0; i < 4; i++) { /* edge i -> i+1 has plane normal n[i+1] */
    V2 ax = A->n[(i + 1) % 4];
    double mn = INFINITY, mx = -INFINITY;
    for (int q = 0; q < 4; q++) { mn = fmin(mn, vdot(ax, B->v[q])); mx = fmax(mx, vdot(ax, A->v[q])); }
    if (mn - mx > best) { best = mn - mx; bn = ax; }
  }
  for (int i = 0; i < 4; i++) {
    V2 ax = B->n[(i + 1) % 4];
    double mn = INFINITY, mx = -INFINITY;
    for (int q = 0; q < 4; q++) { mn = fmin(mn, vdot(ax, A->v[q])); mx = fmax(mx, vdot(ax, B->v[q])); }
    if (mn - mx > best) { best = mn - mx; bn = v2(-ax.x, -ax.y); }
  }
  if (!(best <= 0.0)) return 0;
  V2 n = bn;
  *nOut = n;
  V2 e1a, e1b, e2a, e2b;
  int h1a, h1b, h2a, h2b;
  support_edge(A, n, &e1a, &h1a, &e1b, &h1b);
  support_edge(B, v2(-n.x, -n.y), &e2a, &h2a, &e2b, &h2b);
  double d_e1_a = vcross(e1a, n), d_e1_b = vcross(e1b, n), d_e2_a = vcross(e2a, n), d_e2_b = vcross(e2b, n);
  double e1_denom = 1.0 / (d_e1_b - d_e1_a + 2.2250738585072014e-308);
  double e2_denom = 1.0 / (d_e2_b - d_e2_a + 2.2250738585072014e-308);
  int cnt = 0;
  {
    V2 p1 = vlerp(e1a, e1b, clamp01((d_e2_b - d_e1_a) * e1_denom));
    V2 p2 = vlerp(e2a, e2b, clamp01((d_e1_a - d_e2_a) * e2_denom));
    if (vdot(vsub(p2, p1), n) <= 0.0) { out[cnt].r1 = p1; out[cnt].r2 = p2; out[cnt].hash = 1 + h1a + 4 * h2b; cnt++; }
  }
  {
    V2 p1 = vlerp(e1a, e1b, clamp01((d_e2_a - d_e1_a) * e1_denom));
    V2 p2 = vlerp(e2a, e2b, clamp01((d_e1_b - d_e2_a) * e2_denom));
    if (vdot(vsub(p2, p1), n) <= 0.0) { out[cnt].r1 = p1; out[cnt].r2 = p2; out[cnt].hash = 1 + h1b + 4 * h2a; cnt++; }
  }
  return cnt;
}

/* circle (a) - box (b): CircleToPoly */
static int collide_circle_box(V2 c, double cr, const Box* P, V2* nOut, Contact* out) {
  int inside = 1;
  for (int i = 0; i < 4; i++) {
    V2 a = P->v[i], b = P->v[(i + 1) % 4];
    if ((b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x) < 0) { inside = 0; break; }
  }
  double d;
  V2 pb, n;
  if (inside) {
    double best = -INFINITY;
    int bi = 0;
    for (int i = 0; i < 4; i++) { double q = vdot(vsub(c, P->v[i]), P->n[i]); if (q > best) { best = q; bi = i; } }
    n = v2(-P->n[bi].x, -P->n[bi].y);
    pb = closest_on_seg(c, P->v[(bi + 3) % 4], P->v[bi]);
    d = best;
  } else {
    d = INFINITY;
    pb = c;
    for (int i = 0; i < 4; i++) {
      V2 q = closest_on_seg(c, P->v[(i + 3) % 4], P->v[i]);
      double dd = vlen(vsub(q, c));
      if (dd < d) { d = dd; pb = q; }
    }
    n = d > 0 ? v2((pb.x - c.x) / d, (pb.y - c.y) / d) : v2(1, 0);
  }
  if (d <= cr) {
    *nOut = n;
    out[0].r1 = vadd(c, vmul(n, cr));
    out[0].r2 = pb;
    out[0].hash = 0;
    return 1;
  }
  return 0;
}

/* ---------------------------------------------------------------- car actions (Car.py:55-117) */
static void car_accelerate(Car* c, double dir, int continuous) {
  if (c->finished) return;
  double power = dir;
  double moveDir = c->b.vx * c->dir.x + c->b.vy * c->dir.y;
  if (continuous) {
    if (dir * moveDir < 0) power *= 2;
    else if (dir < 0) power *= 0.75;
  } else {
    if (dir < 0) power = dir * 0.75;
    if (dir == 0) power = moveDir == 0 ? 0 : (moveDir > 0 ? -2 : 2);
    else if (dir < 0 && moveDir > 0) return;
    else if (dir > 0 && moveDir < 0) return;
  }
  V2 v = vrot(v2(CAR_POWER[c->type] * power, 0), c->b.a);
  c->b.vx += v.x; c->b.vy += v.y;
  if (dir == 0 && (c->b.vx * c->dir.x + c->b.vy * c->dir.y) * moveDir < 0) { c->b.vx = 0; c->b.vy = 0; }
}
static void car_turn(Car* c, double dir) {
  if (c->finished) return;
  double rot = dir * (PI / 180);
  c->b.a += rot;
  c->dir = vrot(c->dir, rot);
  V2 v = vrot(v2(c->b.vx, c->b.vy), rot);
  c->b.vx = v.x; c->b.vy = v.y;
}
static void car_crash(Car* c) { c->finished = 1; c->crashed = 1; }

static void process_action(DrvEnv* e, int i, const double* action) {
  /* DrivingEnvironment.py:357-373 */
  double acc, steer;
  if (!e->continuous) { acc = action[0] - 1; steer = (action[1] - 1) * 2; }
  else { acc = action[0]; steer = action[1]; }
  car_accelerate(&e->car[i], acc, e->continuous);
  if (steer != 0) car_turn(&e->car[i], steer);
}

static void tick_car(DrvEnv* e, int i) {
  /* DrivingEnvironment.py:376-426 */
  Car* c = &e->car[i];
  V2 pos = v2(c->b.px, c->b.py);
  c->lane = 4;
  for (int k = 0; k < 2; k++) { int q = is_point_on_road(&e->roads[k], pos, c->b.a); if (q < c->lane) c->lane = q; }
  double diff = vlen(vsub(c->prev, c->goal)) - vlen(vsub(pos, c->goal));
  if (!c->finished) { e->carRewards[i] += diff / 50; e->carPosRewards[i] += fmax(0.0, diff / 50); }
  c->prev = pos;
  if (c->lane >= 3) {
    if (!c->finished) {
      if (c->lane == 3 && vlen(vsub(pos, c->goal)) < 100.0) {
        c->lane = 0;
        c->finished = 1;
        e->carRewards[i] += (MAXTIME - e->elapsed) / 100.0;
        e->carPosRewards[i] += (MAXTIME - e->elapsed) / 100.0;
      } else {
        car_crash(c);
        e->carRewards[i] -= sqrt(c->b.vx * c->b.vx + c->b.vy * c->b.vy) / 5;
      }
    }
  } else if (c->lane == 2) {
    if (!c->finished) e->carRewards[i] -= sqrt(c->b.vx * c->b.vx + c->b.vy * c->b.vy) / 10000;
  }
  /* the position clamps act on a temporary Vec2d in the reference: only the velocity is zeroed */
  if (c->prev.x >= W_ + 50 || c->prev.x <= -50 || c->prev.y >= H_ + 50 || c->prev.y <= -50) { c->b.vx = 0; c->b.vy = 0; }
}

static void move_ped(DrvEnv* e, int i) {
  /* DrivingEnvironment.py:429-506 */
  Ped* p = &e->ped[i];
  if (p->dead) return;
  V2 pos = v2(p->b.px, p->b.py);
  int offRoad = is_off_road(e, pos), out = is_out(pos);
  const Road* rd = &e->roads[p->road];
  if (p->moving > 0) {
    p->moving = fmax(0.0, p->moving - 10.0);
    if (p->crossing) {
      if (!p->beginCrossing && offRoad) { p->moving = 0; p->crossing = 0; p->b.vx = 0; p->b.vy = 0; }
      else if (p->beginCrossing && !offRoad) p->beginCrossing = 0;
    }
    if (out) { p->moving = 0; p->b.vx = 0; p->b.vy = 0; }
  } else {
    if (!p->crossing) {
      uint32_t tick = (uint32_t)e->elapsed;
      p->moving = rnd_int(e, tick, DRV_SITE_PEDMOVE, i, 0, 0, 5000, 30000);
      int speed = rnd_int(e, tick, DRV_SITE_PEDMOVE, i, 0, 1, -2, 2);
      V2 dir = rd->dir;
      if (!offRoad) {
        p->crossing = 1; p->beginCrossing = 0;
        if (speed == 0) speed = 2;
      } else if (out) {
        dir = is_out(vadd(pos, rd->dir)) ? v2(-rd->dir.x, -rd->dir.y) : rd->dir;
      } else if (rnd(e, tick, DRV_SITE_PEDMOVE, i, 0, 2) < 0.05) {
        p->crossing = 1; p->beginCrossing = 1;
        dir = p->side ? rd->normal : v2(-rd->normal.x, -rd->normal.y);
        p->side = p->side ? 0 : 1;
        speed = rnd_int(e, tick, DRV_SITE_PEDMOVE, i, 0, 3, 1, 2);
      }
      p->b.vx = p->speed * dir.x * speed; p->b.vy = p->speed * dir.y * speed;
    } else if (offRoad) { p->crossing = 0; p->beginCrossing = 0; }
  }
}

/* ---------------------------------------------------------------- collision callbacks (all `begin`) */
static double vangle(V2 v) { return atan2(v.y, v.x); }

static void cb_car_crash(DrvEnv* e, int i1, int i2) {
  /* DrivingEnvironment.py:591-637 */
  Car* c1 = &e->car[i1];
  Car* c2 = &e->car[i2];
  V2 v1 = v2(c1->b.vx, c1->b.vy), v2_ = v2(c2->b.vx, c2->b.vy);
  double v1l = vlen(v1) / 5, v2l = vlen(v2_) / 5;
  if (!c1->crashed) e->carRewards[i1] -= v1l;
  if (!c2->crashed) e->carRewards[i2] -= v2l;
  if (c1->lane != 1 && !c1->crashed) e->carRewards[i1] -= v1l;
  if (c2->lane != 1 && !c2->crashed) e->carRewards[i2] -= v2l;
  if (c1->lane == 1 && c2->lane == 1) {
    V2 dp = vsub(v2(c1->b.px, c1->b.py), v2(c2->b.px, c2->b.py));
    if (vlen(v1) > 1 && cos(vangle(dp) - vangle(v1)) < -0.4 && !c1->crashed) e->carRewards[i1] -= v1l;
    if (vlen(v2_) > 1 && cos(vangle(dp) - vangle(v2_)) > 0.4 && !c2->crashed) e->carRewards[i2] -= v2l;
  }
  car_crash(c1);
  car_crash(c2);
}
static int cb_ped_hit(DrvEnv* e, int ci, int pi) {
  /* :640-667; returns 0 when the pair is to be ignored */
  Car* c = &e->car[ci];
  Ped* p = &e->ped[pi];
  V2 v1 = v2(c->b.vx, c->b.vy);
  double v1l = vlen(v1);
  if (!(v1l > 1)) return 0;
  p->moving = 0; p->b.vx = 0; p->b.vy = 0; p->dead = 1; /* Pedestrian.die */
  V2 dp = vsub(v2(c->b.px, c->b.py), v2(p->b.px, p->b.py));
  if (cos(vangle(dp) - vangle(v1)) < -0.4 && !c->finished) {
    car_crash(c);
    e->carRewards[ci] -= v1l / 5;
  }
  return 1;
}
static void cb_car_hit(DrvEnv* e, int ci) {
  /* :670-683 */
  Car* c = &e->car[ci];
  if (!c->finished) e->carRewards[ci] -= sqrt(c->b.vx * c->b.vx + c->b.vy * c->b.vy) / 5;
  car_crash(c);
}

/* ---------------------------------------------------------------- space.step */
static inline double k_scalar(const Body* a, const Body* b, V2 r1, V2 r2, V2 n) {
  double r1cn = vcross(r1, n), r2cn = vcross(r2, n);
  return a->m_inv + a->i_inv * r1cn * r1cn + b->m_inv + b->i_inv * r2cn * r2cn;
}
static inline void apply_impulses(Body* a, Body* b, V2 r1, V2 r2, V2 j) {
  a->vx -= j.x * a->m_inv; a->vy -= j.y * a->m_inv; a->w += a->i_inv * (r1.x * -j.y - r1.y * -j.x);
  b->vx += j.x * b->m_inv; b->vy += j.y * b->m_inv; b->w += b->i_inv * (r2.x * j.y - r2.y * j.x);
}
static inline void apply_bias_impulses(Body* a, Body* b, V2 r1, V2 r2, V2 j) {
  a->vbx -= j.x * a->m_inv; a->vby -= j.y * a->m_inv; a->wb += a->i_inv * (r1.x * -j.y - r1.y * -j.x);
  b->vbx += j.x * b->m_inv; b->vby += j.y * b->m_inv; b->wb += b->i_inv * (r2.x * j.y - r2.y * j.x);
}
static void apply_friction(Body* b, double friction, double rotFriction) {
  /* cutils.py:102-140 after the default velocity update (no forces, gravity 0, damping 1) */
  double factor = friction * b->m, rotFactor = rotFriction * b->m;
  double x = b->vx, y = b->vy, theta = b->w;
  double length = 1.0 / (fabs(x) + fabs(y) + 1e-5);
  double a0 = x * factor * length, a1 = y * factor * length;
  if (fabs(x) < factor) x = 0; else x -= a0;
  if (fabs(y) < factor) y = 0; else y -= a1;
  if (fabs(theta) < rotFactor) theta = 0; else theta -= theta > 0 ? rotFactor : -rotFactor;
  b->vx = x; b->vy = y; b->w = theta;
}

static Arbiter* arb_find(DrvEnv* e, int a, int b) {
  for (int k = 0; k < e->nArb; k++) if (e->arb[k].a == a && e->arb[k].b == b) return &e->arb[k];
  return NULL;
}

static void collide_pair(DrvEnv* e, int i, int j, const Box* boxes, const int* haveBox) {
  /* cpSpaceCollideShapes for shapes i < j (insertion order), at least one of them a car */
  (void)haveBox;
  int a = i, b = j;
  if (!is_ped(a) && is_ped(b)) { a = j; b = i; } /* cpCollide orders by primitive type: circle first */
  Contact con[2];
  memset(con, 0, sizeof(con));
  V2 n = v2(0, 0);
  int cnt;
  if (is_ped(a)) {
    const Ped* p = &e->ped[a - DRV_SHAPE_PED0];
    const Box* B = &boxes[b];
    if (!(p->b.px - PED_R <= B->r && B->l <= p->b.px + PED_R && p->b.py - PED_R <= B->t && B->b <= p->b.py + PED_R)) return;
    cnt = collide_circle_box(v2(p->b.px, p->b.py), PED_R, B, &n, con);
  } else {
    const Box* A = &boxes[a];
    const Box* B = &boxes[b];
    if (!(A->l <= B->r && B->l <= A->r && A->b <= B->t && B->b <= A->t)) return;
    cnt = collide_box_box(A, B, &n, con);
  }
  if (cnt == 0) return;
  Arbiter* arb = arb_find(e, a, b);
  if (!arb) {
    if (e->nArb >= DRV_MAXARB) { e->arbOverflow = 1; return; }
    arb = &e->arb[e->nArb++];
    memset(arb, 0, sizeof(*arb));
    arb->a = a; arb->b = b; arb->state = RC_ARB_FIRST;
  }
  Body sa, sb;
  Body* ba = shape_body(e, a, &sa);
  Body* bb = shape_body(e, b, &sb);
  for (int c = 0; c < cnt; c++) { /* cpArbiterUpdate */
    con[c].r1 = vsub(con[c].r1, v2(ba->px, ba->py));
    con[c].r2 = vsub(con[c].r2, v2(bb->px, bb->py));
    con[c].jnAcc = con[c].jtAcc = 0;
    for (int o = 0; o < arb->ncon; o++)
      if (con[c].hash == arb->con[o].hash) { con[c].jnAcc = arb->con[o].jnAcc; con[c].jtAcc = arb->con[o].jtAcc; }
  }
  arb->ncon = cnt;
  for (int c = 0; c < cnt; c++) arb->con[c] = con[c];
  arb->n = n;
  arb->e = 0.05 * 0.05; /* every shape: elasticity 0.05 */
  arb->u = 0.0;         /* default shape friction 0 */
  if (arb->state == RC_ARB_CACHED) arb->state = RC_ARB_FIRST;
  if (arb->state == RC_ARB_FIRST) { /* begin callbacks */
    if (is_car(a) && is_car(b)) cb_car_crash(e, a - DRV_SHAPE_CAR0, b - DRV_SHAPE_CAR0);
    else if (is_ped(a)) { if (!cb_ped_hit(e, b - DRV_SHAPE_CAR0, a - DRV_SHAPE_PED0)) arb->state = RC_ARB_IGNORE; }
    else cb_car_hit(e, (is_car(a) ? a : b) - DRV_SHAPE_CAR0);
  }
  if (arb->state != RC_ARB_IGNORE) e->active[e->nActive++] = (int)(arb - e->arb);
  else arb->ncon = 0;
  arb->age = 0;
  arb->touched = 1;
}

static void space_step(DrvEnv* e, double dt) {
  int prevStepped = e->stepped;
  e->stepped = 1;
  for (int k = 0; k < e->nArb; k++) {
    if (e->arb[k].state == RC_ARB_FIRST) e->arb[k].state = RC_ARB_NORMAL;
    e->arb[k].touched = 0;
  }
  e->nActive = 0;
  /* integrate positions */
  for (int i = 0; i < e->N; i++) {
    Body* b = &e->car[i].b;
    b->px += (b->vx + b->vbx) * dt; b->py += (b->vy + b->vby) * dt; b->a += (b->w + b->wb) * dt;
    b->vbx = b->vby = b->wb = 0;
  }
  for (int i = 0; i < e->nPed; i++) {
    Body* b = &e->ped[i].b;
    b->px += (b->vx + b->vbx) * dt; b->py += (b->vy + b->vby) * dt;
    b->vbx = b->vby = b->wb = 0;
  }
  /* world geometry of every box */
  static __thread Box boxes[DRV_NSHAPE];
  int haveBox[DRV_NSHAPE];
  memset(haveBox, 0, sizeof(haveBox));
  for (int s = 0; s < DRV_NBUILD; s++) { box_world(e, s, &boxes[s]); haveBox[s] = 1; }
  for (int i = 0; i < e->N; i++) { box_world(e, DRV_SHAPE_CAR0 + i, &boxes[DRV_SHAPE_CAR0 + i]); haveBox[DRV_SHAPE_CAR0 + i] = 1; }
  for (int i = 0; i < e->nObs; i++) { box_world(e, DRV_SHAPE_OBS0 + i, &boxes[DRV_SHAPE_OBS0 + i]); haveBox[DRV_SHAPE_OBS0 + i] = 1; }
  /* collide, pairs in ascending (i, j) insertion order; only pairs with a car can make contacts (the
     handlers ignore pedestrian-pedestrian and pedestrian-obstacle pairs: :67-68) */
  for (int s = 0; s < DRV_NBUILD; s++)
    for (int i = 0; i < e->N; i++) collide_pair(e, s, DRV_SHAPE_CAR0 + i, boxes, haveBox);
  for (int i = 0; i < e->N; i++) {
    for (int j = i + 1; j < e->N; j++) collide_pair(e, DRV_SHAPE_CAR0 + i, DRV_SHAPE_CAR0 + j, boxes, haveBox);
    for (int j = 0; j < e->nPed; j++) collide_pair(e, DRV_SHAPE_CAR0 + i, DRV_SHAPE_PED0 + j, boxes, haveBox);
    for (int j = 0; j < e->nObs; j++) collide_pair(e, DRV_SHAPE_CAR0 + i, DRV_SHAPE_OBS0 + j, boxes, haveBox);
  }
  /* cached pairs that did not collide: age, expiry after 3 substeps (no separate handlers here) */
  {
    int w = 0;
    for (int k = 0; k < e->nArb; k++) {
      Arbiter* arb = &e->arb[k];
      if (!arb->touched) {
        arb->age += 1;
        if (arb->state != RC_ARB_CACHED) arb->state = RC_ARB_CACHED;
        if (arb->age >= 3) continue;
      }
      if (w != k) {
        for (int q = 0; q < e->nActive; q++) if (e->active[q] == k) e->active[q] = w;
        e->arb[w] = *arb;
      }
      w++;
    }
    e->nArb = w;
  }
  /* pre-step */
  const double slop = 0.1;
  const double biasCoef = 1.0 - pow(pow(1.0 - 0.1, 60.0), dt);
  for (int q = 0; q < e->nActive; q++) {
    Arbiter* arb = &e->arb[e->active[q]];
    Body sa, sb;
    Body* a = shape_body(e, arb->a, &sa);
    Body* b = shape_body(e, arb->b, &sb);
    V2 n = arb->n;
    V2 bd = v2(b->px - a->px, b->py - a->py);
    for (int c = 0; c < arb->ncon; c++) {
      Contact* con = &arb->con[c];
      con->nMass = 1.0 / k_scalar(a, b, con->r1, con->r2, n);
      con->tMass = 1.0 / k_scalar(a, b, con->r1, con->r2, v2(-n.y, n.x));
      double dist = vdot(vadd(vsub(con->r2, con->r1), bd), n);
      con->bias = -biasCoef * fmin(0.0, dist + slop) / dt;
      con->jBias = 0;
      V2 vr = v2((b->vx - con->r2.y * b->w) - (a->vx - con->r1.y * a->w), (b->vy + con->r2.x * b->w) - (a->vy + con->r1.x * a->w));
      con->bounce = vdot(vr, n) * arb->e;
    }
  }
  /* velocities: friction callbacks of cars (cutils.py:78-85) and dead pedestrians (:88-90); live
     pedestrians have the default update (nothing to do without forces) */
  for (int i = 0; i < e->N; i++) {
    if (e->car[i].finished) apply_friction(&e->car[i].b, 5e-4, 2e-5); else apply_friction(&e->car[i].b, 5e-5, 1e-5);
  }
  for (int i = 0; i < e->nPed; i++) if (e->ped[i].dead) apply_friction(&e->ped[i].b, 5e-2, 2e-4);
  /* warm start */
  double dt_coef = prevStepped ? 1.0 : 0.0;
  for (int q = 0; q < e->nActive; q++) {
    Arbiter* arb = &e->arb[e->active[q]];
    if (arb->state == RC_ARB_FIRST) continue;
    Body sa, sb;
    Body* a = shape_body(e, arb->a, &sa);
    Body* b = shape_body(e, arb->b, &sb);
    V2 n = arb->n;
    for (int c = 0; c < arb->ncon; c++) {
      Contact* con = &arb->con[c];
      V2 j = v2(n.x * con->jnAcc - n.y * con->jtAcc, n.x * con->jtAcc + n.y * con->jnAcc);
      apply_impulses(a, b, con->r1, con->r2, vmul(j, dt_coef));
    }
  }
  /* solver */
  for (int it = 0; it < 10; it++) {
    for (int q = 0; q < e->nActive; q++) {
      Arbiter* arb = &e->arb[e->active[q]];
      Body sa, sb;
      Body* a = shape_body(e, arb->a, &sa);
      Body* b = shape_body(e, arb->b, &sb);
      V2 n = arb->n;
      for (int c = 0; c < arb->ncon; c++) {
        Contact* con = &arb->con[c];
        V2 r1 = con->r1, r2 = con->r2;
        V2 vb1 = v2(a->vbx - r1.y * a->wb, a->vby + r1.x * a->wb);
        V2 vb2 = v2(b->vbx - r2.y * b->wb, b->vby + r2.x * b->wb);
        V2 vr = v2((b->vx - r2.y * b->w) - (a->vx - r1.y * a->w), (b->vy + r2.x * b->w) - (a->vy + r1.x * a->w));
        double vbn = vdot(vsub(vb2, vb1), n), vrn = vdot(vr, n), vrt = vdot(vr, v2(-n.y, n.x));
        double jbn = (con->bias - vbn) * con->nMass, jbnOld = con->jBias;
        con->jBias = fmax(jbnOld + jbn, 0.0);
        double jn = -(con->bounce + vrn) * con->nMass, jnOld = con->jnAcc;
        con->jnAcc = fmax(jnOld + jn, 0.0);
        double jtMax = arb->u * con->jnAcc, jt = -vrt * con->tMass, jtOld = con->jtAcc;
        con->jtAcc = fmin(fmax(jtOld + jt, -jtMax), jtMax);
        apply_bias_impulses(a, b, r1, r2, vmul(n, con->jBias - jbnOld));
        double dn = con->jnAcc - jnOld, dtt = con->jtAcc - jtOld;
        apply_impulses(a, b, r1, r2, v2(n.x * dn - n.y * dtt, n.x * dtt + n.y * dn));
      }
    }
  }
}

/* ---------------------------------------------------------------- vision (cutils.py) */
enum { S_NO = 0, S_PARTIAL = 1, S_DISTANT = 2, S_NORMAL = 3, S_MISCLASS = 4 };
enum { I_NO = 0, I_NEARBY = 1, I_OCCLUDE = 2 };

typedef struct {
  int seen;
  V2 p;        /* centre in the observer frame */
  double c, s; /* cos / sin of the relative angle */
  int hasCorners;
  V2 corner[4]; /* relative to the observer, NOT rotated (reference behaviour: the rotated() results are discarded) */
  double w, h;
  int finished;
  int src;     /* entity index (RNG address) */
} Det;

static Det is_seen_in_radius(V2 point, const V2* corners, double angle, V2 obsPt, double obsAngle, double maxDist,
                             double distantDist) {
  /* cutils.py:569-596 */
  Det d;
  memset(&d, 0, sizeof(d));
  V2 tr = vsub(point, obsPt);
  double dist = vlensq(tr);
  if (dist <= maxDist) {
    d.seen = S_DISTANT;
    if (dist <= distantDist) d.seen = S_NORMAL;
    if (corners) {
      d.hasCorners = 1;
      for (int k = 0; k < 4; k++) d.corner[k] = vadd(vsub(corners[k], point), tr);
    }
    d.p = vrot(tr, -obsAngle);
    double ta = angle - obsAngle;
    d.c = cos(ta); d.s = sin(ta);
  }
  return d;
}

static int does_interact_poly(const Det* e1, const Det* e2, double radius) {
  /* cutils.py:643-696 (canOcclude = True at every call site) */
  int ret = I_NO;
  if (e1->seen == S_NO || e2->seen == S_NO) return ret;
  V2 p1 = e1->p, p2 = e2->p;
  if (radius > 0 && vlensq(vsub(p2, p1)) < radius) ret = I_NEARBY;
  double angle2 = vangle(p2);
  double ang[4], dist[4];
  for (int k = 0; k < 4; k++) {
    ang[k] = vangle(e2->corner[k]) - angle2;
    dist[k] = vlensq(e2->corner[k]);
    if (ang[k] > PI) ang[k] -= 2 * PI;
    if (ang[k] < -PI) ang[k] += 2 * PI;
  }
  int mn = 0, mx = 0, cl = 0;
  for (int k = 1; k < 4; k++) {
    if (ang[k] < ang[mn]) mn = k;
    if (ang[k] > ang[mx]) mx = k;
    if (dist[k] < dist[cl]) cl = k;
  }
  V2 q1 = e2->corner[mn], q2 = e2->corner[mx], qm = e2->corner[cl];
  double pAngle = vangle(p1) - angle2;
  if (pAngle > PI) pAngle -= 2 * PI; else if (pAngle < -PI) pAngle += 2 * PI;
  if (pAngle > ang[mn] && pAngle < ang[mx]) {
    if (cl == mn || cl == mx) {
      if (vcross(vsub(q2, q1), vsub(p1, q1)) < 0) ret = I_OCCLUDE;
    } else if (vcross(vsub(q2, qm), vsub(p1, qm)) < 0 && vcross(vsub(qm, q1), vsub(p1, q1)) < 0) ret = I_OCCLUDE;
  }
  return ret;
}

typedef struct { DrvEnv* e; uint32_t tick; int obs; } NoiseCtx;

static void add_noise_rect(NoiseCtx* nc, int site, Det* obj, int interaction, double maxDist, int misClass) {
  /* cutils.py:479-542; the corner updates are not restated: nothing reads the corners afterwards */
  DrvEnv* e = nc->e;
  double magn = e->noiseMag, rand_ = e->randBase;
  if (!obj->seen) return;
  int idx = obj->src;
  double u0 = rnd(e, nc->tick, site, nc->obs, idx, 0), u1 = rnd(e, nc->tick, site, nc->obs, idx, 1);
  V2 noiseVec = v2((u0 - 0.5) * magn, (u1 - 0.5) * magn);
  if (e->noiseType == DYNENV_NOISE_RANDOM) {
    if (rnd(e, nc->tick, site, nc->obs, idx, 2) < rand_) obj->seen = S_NO;
    else {
      V2 newPos = vadd(obj->p, noiseVec);
      double angleDiff = (rnd(e, nc->tick, site, nc->obs, idx, 4) - 0.5) * magn * ANGLE_NOISE;
      double ang = atan2(obj->s, obj->c) + angleDiff;
      obj->c = cos(ang); obj->s = sin(ang);
      obj->p = newPos;
    }
  } else {
    int st = obj->seen;
    double range = 0.25 + 3.75 * vlen(obj->p) / maxDist;
    double mult = range;
    if (interaction == I_NEARBY) mult = range * 2;
    if (st == S_DISTANT) mult = range * 3;
    else if (st == S_PARTIAL) mult = range * 4;
    V2 newPos = vadd(obj->p, vmul(noiseVec, mult));
    if (rnd(e, nc->tick, site, nc->obs, idx, 2) < rand_ * mult) { obj->seen = S_NO; return; }
    if (misClass && rnd(e, nc->tick, site, nc->obs, idx, 3) < rand_ * mult / 2) obj->seen = S_MISCLASS;
    double angleDiff = (rnd(e, nc->tick, site, nc->obs, idx, 4) - 0.5) * magn * ANGLE_NOISE * 0.25;
    double ang = atan2(obj->s, obj->c) + angleDiff;
    obj->c = cos(ang); obj->s = sin(ang);
    obj->p = newPos;
  }
}

typedef struct { int seen; double dist, c, s, type; int src; } LaneDet;

static void add_noise_lane(NoiseCtx* nc, LaneDet* obj, double maxDist) {
  /* cutils.py:380-413 */
  DrvEnv* e = nc->e;
  double magn = e->noiseMag, rand_ = e->randBase;
  if (!obj->seen) return;
  double distNoise = (rnd(e, nc->tick, DRV_SITE_OBS_LANE, nc->obs, obj->src, 0) - 0.5) * magn;
  double angleDiff = (rnd(e, nc->tick, DRV_SITE_OBS_LANE, nc->obs, obj->src, 1) - 0.5) * magn;
  if (e->noiseType == DYNENV_NOISE_RANDOM) {
    if (rnd(e, nc->tick, DRV_SITE_OBS_LANE, nc->obs, obj->src, 2) < rand_) obj->seen = S_NO;
    obj->dist *= distNoise;
    double ang = atan2(obj->s, obj->c) + ANGLE_NOISE * angleDiff;
    obj->c = cos(ang); obj->s = sin(ang);
  } else {
    double m1 = 0.25 + 3.75 * obj->dist * obj->dist / maxDist;
    if (rnd(e, nc->tick, DRV_SITE_OBS_LANE, nc->obs, obj->src, 2) < rand_ * m1) obj->seen = S_NO;
    obj->dist += distNoise * m1;
    double ang = atan2(obj->s, obj->c) + ANGLE_NOISE * m1 / 5 * angleDiff;
    obj->c = cos(ang); obj->s = sin(ang);
  }
}

static int lane_distances(const Road* r, V2 carPos, double carAngle, LaneDet* out, int src0) {
  /* Road.getCarLaneDistances (Road.py:36-71): 2*nLanes rows, or one NoSighting row (dropped by the caller) */
  V2 pt = vsub(carPos, r->p0);
  double dist = vcross(r->dir, pt) / r->width;
  if (fabs(dist) > 10) return 0;
  double a = r->angle - carAngle, c = cos(a), s = sin(a), distMult = 1, flip = 1;
  if (c >= 0) { flip = -1; c = -c; s = -s; distMult = -1; }
  int n = 2 * r->nLanes;
  for (int k = 0; k < n; k++) {
    out[k].seen = S_NORMAL;
    out[k].dist = ((dist + 0.5) + (k - r->nLanes)) * r->width * 0.1 * distMult;
    out[k].c = c; out[k].s = s;
    out[k].type = (k < r->nLanes ? 1 : -1) * flip;
    out[k].src = src0 + k;
  }
  return n;
}

static inline double normalize_(double pt, double nf, double mean) { return ((pt * nf) - mean) * 2; } /* cutils.py:318-323 */
#define NORMX (10.0 / W_)
#define NORMY (10.0 / H_)
#define NORMW (1.0 / 7.5)
#define NORMH (1.0 / 15)
#define SNORMX (0.5 / (W_ + 100))
#define SNORMY (0.5 / (H_ + 100))
#define SNORMW (1.0 / 15)
#define SNORMH (1.0 / 25)

static void list_append(Det* l, int* n, int cap, Det d) { if (*n < cap) l[(*n)++] = d; }

static void agent_vision(DrvEnv* e, int ag, uint32_t tick, DrvAgentObs* out) {
  /* DrivingEnvironment.py:750-977 */
  memset(out, 0, sizeof(*out));
  Car* agent = &e->car[ag];
  double ang = agent->b.a;
  V2 apos = v2(agent->b.px, agent->b.py);
  const double maxVis0 = (W_ * 0.4) * (W_ * 0.4), maxVis1 = (W_ * 0.6) * (W_ * 0.6); /* _setup_vision(0.4, 0.6) */
  NoiseCtx nc = {e, tick, ag};
  Det selfDet;
  memset(&selfDet, 0, sizeof(selfDet));
  selfDet.seen = S_NORMAL; selfDet.p = apos; selfDet.c = cos(ang); selfDet.s = sin(ang); selfDet.src = 0;
  Det cars[LISTCAP], peds[LISTCAP], obss[LISTCAP], builds[4];
  int ncar = 0, nped = 0, nobs = 0;
  for (int c = 0; c < e->N; c++) {
    if (c == ag) continue;
    Car* o = &e->car[c];
    V2 pt = v2(o->b.px, o->b.py);
    /* Car.getPoints ignores the car's angle: axis-aligned corners (Car.py:124-126) */
    V2 cor[4] = {v2(pt.x + o->height, pt.y + o->width), v2(pt.x - o->height, pt.y + o->width),
                 v2(pt.x - o->height, pt.y - o->width), v2(pt.x + o->height, pt.y - o->width)};
    Det d = is_seen_in_radius(pt, cor, o->b.a, apos, ang, maxVis0, maxVis1);
    d.w = o->width; d.h = o->height; d.finished = o->finished; d.src = c;
    if (d.seen != S_NO) list_append(cars, &ncar, LISTCAP, d);
  }
  for (int k = 0; k < e->nObs; k++) {
    V2 pt = e->obst[k];
    V2 cor[4] = {v2(pt.x + OBST_HALF, pt.y + OBST_HALF), v2(pt.x - OBST_HALF, pt.y + OBST_HALF),
                 v2(pt.x - OBST_HALF, pt.y - OBST_HALF), v2(pt.x + OBST_HALF, pt.y - OBST_HALF)};
    Det d = is_seen_in_radius(pt, cor, 0, apos, ang, maxVis0, maxVis1);
    d.w = OBST_HALF; d.h = OBST_HALF; d.src = k;
    if (d.seen != S_NO) list_append(obss, &nobs, LISTCAP, d);
  }
  for (int k = 0; k < 4; k++) {
    V2 pt = BUILD_POS[k];
    V2 cor[4] = {v2(pt.x + BUILD_W, pt.y + BUILD_H), v2(pt.x - BUILD_W, pt.y + BUILD_H),
                 v2(pt.x - BUILD_W, pt.y - BUILD_H), v2(pt.x + BUILD_W, pt.y - BUILD_H)};
    builds[k] = is_seen_in_radius(pt, cor, 0, apos, ang, 20000000, 20000000);
  }
  for (int k = 0; k < e->nPed; k++) {
    Det d = is_seen_in_radius(v2(e->ped[k].b.px, e->ped[k].b.py), NULL, 0, apos, ang, maxVis0, maxVis1);
    d.src = k;
    if (d.seen != S_NO) list_append(peds, &nped, LISTCAP, d);
  }
  LaneDet lanes[LISTCAP];
  int nlane = 0;
  nlane += lane_distances(&e->roads[0], apos, ang, lanes + nlane, 0);
  nlane += lane_distances(&e->roads[1], apos, ang, lanes + nlane, 4);
  /* occlusion by buildings (:785-792) */
  {
    int w = 0;
    for (int i = 0; i < ncar; i++) { int m = 0; for (int b = 0; b < 4; b++) { int t = does_interact_poly(&cars[i], &builds[b], 0); if (t > m) m = t; } if (m != I_OCCLUDE) cars[w++] = cars[i]; }
    ncar = w; w = 0;
    for (int i = 0; i < nped; i++) { int m = 0; for (int b = 0; b < 4; b++) { int t = does_interact_poly(&peds[i], &builds[b], 0); if (t > m) m = t; } if (m != I_OCCLUDE) peds[w++] = peds[i]; }
    nped = w; w = 0;
    for (int i = 0; i < nobs; i++) { int m = 0; for (int b = 0; b < 4; b++) { int t = does_interact_poly(&obss[i], &builds[b], 0); if (t > m) m = t; } if (m != I_OCCLUDE) obss[w++] = obss[i]; }
    nobs = w;
  }
  /* pedestrians hidden by / near cars and obstacles (:795-805); `max` of two lists is lexicographic */
  int carPed[LISTCAP], obsPed[LISTCAP], pedInter[LISTCAP];
  for (int i = 0; i < nped; i++) {
    int m = 0;
    for (int c = 0; c < ncar; c++) { int t = does_interact_poly(&peds[i], &cars[c], 400); if (t > m) m = t; }
    carPed[i] = m;
    m = 0;
    for (int o = 0; o < nobs; o++) { int t = does_interact_poly(&peds[i], &obss[o], 400); if (t > m) m = t; }
    obsPed[i] = m;
  }
  {
    int useObs = 0;
    for (int i = 0; i < nped; i++) if (carPed[i] != obsPed[i]) { useObs = obsPed[i] > carPed[i]; break; }
    for (int i = 0; i < nped; i++) pedInter[i] = useObs ? obsPed[i] : carPed[i];
  }
  for (int i = 0; i < nped; i++) if (pedInter[i] == I_OCCLUDE) peds[i].seen = S_NO;
  /* noise (:808-818) */
  add_noise_rect(&nc, DRV_SITE_OBS_SELF, &selfDet, I_NO, maxVis1, 0);
  for (int i = 0; i < ncar; i++) add_noise_rect(&nc, DRV_SITE_OBS_CAR, &cars[i], I_NO, maxVis1, 1);
  for (int i = 0; i < nped; i++) add_noise_rect(&nc, DRV_SITE_OBS_PED, &peds[i], pedInter[i], maxVis0, 0);
  for (int i = 0; i < nobs; i++) add_noise_rect(&nc, DRV_SITE_OBS_OBST, &obss[i], I_NO, maxVis1, 1);
  for (int i = 0; i < nlane; i++) add_noise_lane(&nc, &lanes[i], maxVis1);
  /* cars <-> obstacles misclassification (:821-826) */
  { int n0 = ncar; for (int i = 0; i < n0; i++) if (cars[i].seen == S_MISCLASS) { Det d = cars[i]; d.seen = S_NORMAL; list_append(obss, &nobs, LISTCAP, d); } }
  { int n0 = nobs; for (int i = 0; i < n0; i++) if (obss[i].seen == S_MISCLASS) { Det d = obss[i]; d.seen = S_NORMAL; d.finished = 0; list_append(cars, &ncar, LISTCAP, d); } }
  /* random false positives (:828-874) */
  for (int i = 0; i < 10; i++) {
    if (rnd(e, tick, DRV_SITE_OBS_FP, ag, i, 0) < e->randBase) {
      int c = rnd_int(e, tick, DRV_SITE_OBS_FP, ag, i, 1, 0, 5);
      double d = rnd(e, tick, DRV_SITE_OBS_FP, ag, i, 2) * maxVis1;
      double a1 = rnd(e, tick, DRV_SITE_OBS_FP, ag, i, 3) * 2 * PI;
      V2 pos = vrot(v2(d, 0), a1);
      double angle = rnd(e, tick, DRV_SITE_OBS_FP, ag, i, 4) * 2 * PI;
      if (c <= 1) {
        double w = rnd(e, tick, DRV_SITE_OBS_FP, ag, i, 5) * 5 + 5;
        double h = rnd(e, tick, DRV_SITE_OBS_FP, ag, i, 6) * 10 + 5;
        Det dt;
        memset(&dt, 0, sizeof(dt));
        dt.seen = S_NORMAL; dt.p = pos; dt.c = cos(angle); dt.s = sin(angle); dt.w = w; dt.h = h;
        if (c == 0) list_append(cars, &ncar, LISTCAP, dt); else list_append(obss, &nobs, LISTCAP, dt);
      } else if (c == 2) {
        Det dt;
        memset(&dt, 0, sizeof(dt));
        dt.seen = S_NORMAL; dt.p = pos;
        list_append(peds, &nped, LISTCAP, dt);
      } else if (c == 3) {
        double a = (rnd(e, tick, DRV_SITE_OBS_FP, ag, i, 5) - 0.5) * PI * 2;
        double dist = floor(rnd(e, tick, DRV_SITE_OBS_FP, ag, i, 6) * W_ / 2); /* `random() * W // 2` */
        if (nlane < LISTCAP) {
          LaneDet* l = &lanes[nlane++];
          l->seen = S_NORMAL; l->dist = dist; l->c = cos(a); l->s = sin(a);
          l->type = rnd_int(e, tick, DRV_SITE_OBS_FP, ag, i, 7, -1, 1);
        }
      }
    }
  }
  /* false-positive pedestrians near cars (:877-882) */
  if (e->noiseType == DYNENV_NOISE_REALISTIC) {
    for (int i = 0; i < ncar; i++) {
      if (cars[i].seen != S_NORMAL) continue;
      if (!(rnd(e, tick, DRV_SITE_OBS_FPPED, ag, i, 0) < e->randBase * 10)) continue;
      if (!(vlen(cars[i].p) < 250)) continue;
      double ox = 2 * rnd(e, tick, DRV_SITE_OBS_FPPED, ag, i, 1) - 1.0;
      double oy = 2 * rnd(e, tick, DRV_SITE_OBS_FPPED, ag, i, 2) - 1.0;
      Det dt;
      memset(&dt, 0, sizeof(dt));
      dt.seen = S_NORMAL; dt.p = v2(cars[i].p.x + ox * 10, cars[i].p.y + oy * 10);
      list_append(peds, &nped, LISTCAP, dt);
    }
  }
  /* final filters and encoding (:885-977) */
  out->self[0] = normalize_(selfDet.p.x, NORMX, 5.0); out->self[1] = normalize_(selfDet.p.y, NORMY, 5.0);
  out->self[2] = selfDet.c; out->self[3] = selfDet.s;
  out->self[4] = normalize_(agent->width, NORMW, 0.5); out->self[5] = normalize_(agent->height, NORMH, 0.5);
  out->self[6] = normalize_(agent->goal.x, NORMX, 5.0); out->self[7] = normalize_(agent->goal.y, NORMY, 5.0);
  out->self[8] = agent->finished;
  int k = 0;
  for (int i = 0; i < ncar && k < DRV_CAP_CAR; i++) {
    if (cars[i].seen == S_NO || cars[i].seen == S_MISCLASS) continue;
    double* r = out->car[k++];
    r[0] = normalize_(cars[i].p.x, NORMX, 0); r[1] = normalize_(cars[i].p.y, NORMY, 0); r[2] = cars[i].c; r[3] = cars[i].s;
    r[4] = normalize_(cars[i].w, NORMW, 0.5); r[5] = normalize_(cars[i].h, NORMH, 0.5); r[6] = cars[i].finished;
  }
  out->counts[0] = k; k = 0;
  for (int i = 0; i < nobs && k < DRV_CAP_OBS; i++) {
    if (obss[i].seen == S_NO || obss[i].seen == S_MISCLASS) continue;
    double* r = out->obs[k++];
    r[0] = normalize_(obss[i].p.x, NORMX, 0); r[1] = normalize_(obss[i].p.y, NORMY, 0); r[2] = obss[i].c; r[3] = obss[i].s;
    r[4] = normalize_(obss[i].w, NORMW, 0.5); r[5] = normalize_(obss[i].h, NORMH, 0.5);
  }
  out->counts[1] = k; k = 0;
  for (int i = 0; i < nped && k < DRV_CAP_PED; i++) {
    if (peds[i].seen == S_NO) continue;
    double* r = out->ped[k++];
    r[0] = normalize_(peds[i].p.x, NORMX, 0); r[1] = normalize_(peds[i].p.y, NORMY, 0);
  }
  out->counts[2] = k; k = 0;
  for (int i = 0; i < nlane && k < DRV_CAP_LANE; i++) {
    if (lanes[i].seen == S_NO) continue;
    double* r = out->lane[k++];
    r[0] = lanes[i].dist; r[1] = lanes[i].c; r[2] = lanes[i].s; r[3] = lanes[i].type;
  }
  out->counts[3] = k;
}

/* getFullState pieces (DrivingEnvironment.py:686-747) */
static void car_row(const Car* c, double* r) {
  r[0] = normalize_(c->b.px, SNORMX, 0); r[1] = normalize_(c->b.py, SNORMY, 0); r[2] = cos(c->b.a); r[3] = sin(c->b.a);
  r[4] = normalize_(c->width, SNORMW, 0.5); r[5] = normalize_(c->height, SNORMH, 0.5); r[6] = c->finished;
}
static void full_obs_agent(DrvEnv* e, int ag, DrvAgentObs* out) {
  /* get_full_obs (:121-124): ([others, obstacles, pedestrians], [self, lanes], (1,1,1)) */
  memset(out, 0, sizeof(*out));
  Car* a = &e->car[ag];
  double r7[7];
  car_row(a, r7);
  for (int k = 0; k < 6; k++) out->self[k] = r7[k];
  out->self[6] = normalize_(a->goal.x, SNORMX, 0); out->self[7] = normalize_(a->goal.y, SNORMY, 0); out->self[8] = a->finished;
  int k = 0;
  for (int c = 0; c < e->N; c++) if (c != ag) car_row(&e->car[c], out->car[k++]);
  out->counts[0] = k;
  for (int o = 0; o < e->nObs; o++) {
    out->obs[o][0] = normalize_(e->obst[o].x, SNORMX, 0); out->obs[o][1] = normalize_(e->obst[o].y, SNORMY, 0);
    out->obs[o][2] = normalize_(OBST_HALF, SNORMW, 0.5); out->obs[o][3] = normalize_(OBST_HALF, SNORMH, 0.5);
  }
  out->counts[1] = e->nObs;
  for (int p = 0; p < e->nPed; p++) {
    out->ped[p][0] = normalize_(e->ped[p].b.px, SNORMX, 0); out->ped[p][1] = normalize_(e->ped[p].b.py, SNORMY, 0);
  }
  out->counts[2] = e->nPed;
  k = 0;
  for (int rd = 0; rd < 2; rd++) {
    const Road* r = &e->roads[rd];
    for (int i = -r->nLanes; i <= r->nLanes; i++) {
      /* `l.Lanes[i - l.nLanes]` indexes the 2*nLanes+1 lane lines with a NEGATIVE Python index, so row i holds
         the geometry of lane offset (i - nLanes) mod (2*nLanes+1) - nLanes, under the type tag of i (:690-696) */
      int cnt = 2 * r->nLanes + 1;
      int off = (((i - r->nLanes) % cnt) + cnt) % cnt - r->nLanes;
      V2 l0 = vadd(r->p0, vmul(r->normal, off * r->width)), l1 = vadd(r->p1, vmul(r->normal, off * r->width));
      double* row = out->lane[k++];
      row[0] = normalize_(l0.x, SNORMX, 0); row[1] = normalize_(l0.y, SNORMY, 0);
      row[2] = normalize_(l1.x, SNORMX, 0); row[3] = normalize_(l1.y, SNORMY, 0);
      row[4] = abs(i) == r->nLanes ? 1 : (i == 0 ? -1 : 0);
    }
  }
  out->counts[3] = k;
}

/* ---------------------------------------------------------------- public API (ctypes) */
int drv_stepout_size(void) { return (int)sizeof(DrvStepOut); }

static void roads_init(DrvEnv* e) {
  road_init(&e->roads[0], 2, 35, v2(875, 0), v2(875, 1000)); /* :110-115 */
  road_init(&e->roads[1], 1, 35, v2(0, 500), v2(1750, 500));
}
static void car_set_type(Car* c, int type) {
  c->type = type; c->width = CAR_WIDTH[type]; c->height = CAR_LENGTH[type];
  /* moment_for_poly of the box (+-height, +-width) */
  body_set_mass(&c->b, CAR_MASS[type], CAR_MASS[type] * (4 * c->height * c->height + 4 * c->width * c->width) / 12.0);
}

DrvEnv* drv_create(int nPlayers, int obsType, int noiseType, double noiseMag, int continuous, unsigned flags,
                   uint64_t seed, uint32_t env_id) {
  (void)flags;
  DrvEnv* e = (DrvEnv*)calloc(1, sizeof(DrvEnv));
  e->N = nPlayers > NCAR ? NCAR : nPlayers;
  e->obsType = obsType; e->noiseType = noiseType; e->noiseMag = noiseMag; e->continuous = continuous;
  e->randBase = 0.01 * noiseMag;
  e->rng.mode = 0; e->rng.seed = seed; e->rng.env_id = env_id;
  roads_init(e);
  return e;
}
void drv_destroy(DrvEnv* e) { free(e); }
void drv_set_tape(DrvEnv* e, const double* tape, int n) { e->rng.mode = tape ? 1 : 0; e->rng.tape = tape; e->rng.tape_n = n; e->rng.tape_pos = 0; e->rng.tape_underflow = 0; }
int drv_tape_pos(DrvEnv* e) { return e->rng.tape_pos; }
int drv_tape_underflow(DrvEnv* e) { return e->rng.tape_underflow; }
int drv_arb_overflow(DrvEnv* e) { return e->arbOverflow; }
void drv_debug_substep_range(DrvEnv* e, int begin, int end) { e->subBegin = begin; e->subEnd = end; }

static V2 walk_spot(const Road* r, int side, double length, double width) {
  /* Road.getWalkSpot (Road.py:117-125) */
  double off = (r->nLanes + 1) * r->width * (side ? -1 : 1);
  V2 w0 = vadd(r->p0, vmul(r->normal, off)), w1 = vadd(r->p1, vmul(r->normal, off));
  V2 c = vadd(w0, vmul(vsub(w1, w0), length));
  return vadd(c, vmul(r->normal, width * r->width * (side ? 1 : -1)));
}

void drv_reset(DrvEnv* e) {
  /* DrivingEnvironment.__init__ / _setup_scene (:20-115, 527-584) */
  int episode = e->episode + 1;
  OracleRng g = e->rng;
  int N = e->N, oT = e->obsType, nT = e->noiseType, cA = e->continuous;
  double nM = e->noiseMag;
  memset(e, 0, sizeof(*e));
  e->rng = g; e->N = N; e->obsType = oT; e->noiseType = nT; e->continuous = cA; e->noiseMag = nM; e->randBase = 0.01 * nM;
  e->episode = episode;
  roads_init(e);
  int k = 0;
#define U() rnd(e, DYNENV_TICK_RESET, DRV_SITE_RESET, 0, 0, k++)
  int roadSel[NCAR], endSel[NCAR], teams[NCAR], types[NCAR];
  for (int i = 0; i < N; i++) roadSel[i] = (int)(U() * 2);
  for (int i = 0; i < N; i++) endSel[i] = (int)(U() * 2);
  for (int i = 0; i < N; i++) teams[i] = (int)(U() * 3);
  for (int i = 0; i < N; i++) types[i] = (int)(U() * 4);
  int perm[30];
  for (int i = 0; i < 30; i++) perm[i] = i;
  for (int i = 29; i > 0; i--) { int j = (int)(U() * (i + 1)); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
  for (int i = 0; i < N; i++) {
    int spotID = perm[i];
    int roadID = spotID < 20 ? 0 : 1;
    spotID -= roadID ? 20 : 0;
    int laneID = spotID / 5, spot = spotID % 5;
    const Road* r = &e->roads[roadID];
    /* Road.getSpot (Road.py:100-114) */
    int end = laneID >= r->nLanes ? 1 : 0;
    V2 pos = end ? r->p1 : r->p0;
    V2 spotDir = vmul(end ? v2(-r->dir.x, -r->dir.y) : r->dir, 90.0);
    V2 laneDir = vmul(end ? r->normal : v2(-r->normal.x, -r->normal.y), r->width);
    double lane = (end ? laneID - r->nLanes : laneID) + 0.5;
    V2 p = vadd(vadd(pos, vmul(laneDir, lane)), vmul(spotDir, spot));
    double angle = atan2(spotDir.y, spotDir.x);
    Car* c = &e->car[i];
    car_set_type(c, types[i]);
    c->team = teams[i];
    c->goal = endSel[i] ? e->roads[roadSel[i]].p1 : e->roads[roadSel[i]].p0;
    c->b.px = p.x; c->b.py = p.y; c->b.a = angle;
    c->dir = vrot(v2(1, 0), angle);
    c->prev = p;
    c->lane = 4;
  }
  e->nPed = 10 + (int)(U() * 11);
  {
    int roadIds[NPEDMAX], sideIds[NPEDMAX];
    double lenOffs[NPEDMAX], widthOffs[NPEDMAX];
    for (int i = 0; i < e->nPed; i++) roadIds[i] = (int)(U() * 2);
    for (int i = 0; i < e->nPed; i++) sideIds[i] = (int)(U() * 2);
    for (int i = 0; i < e->nPed; i++) lenOffs[i] = U();
    for (int i = 0; i < e->nPed; i++) widthOffs[i] = U() / 2 + 0.25;
    for (int i = 0; i < e->nPed; i++) {
      Ped* p = &e->ped[i];
      V2 c = walk_spot(&e->roads[roadIds[i]], sideIds[i], lenOffs[i], widthOffs[i]);
      body_set_mass(&p->b, 90, 90 * (PED_R * PED_R) / 2.0);
      p->b.px = c.x; p->b.py = c.y;
      p->road = roadIds[i]; p->side = sideIds[i];
      p->speed = 3 + (int)(U() * 4);
    }
  }
  {
    int n = 10 + (int)(U() * 11);
    int roadIds[NOBSMAX], sideIds[NOBSMAX];
    double lenOffs[NOBSMAX], widthOffs[NOBSMAX];
    for (int i = 0; i < n; i++) roadIds[i] = (int)(U() * 2);
    for (int i = 0; i < n; i++) sideIds[i] = (int)(U() * 2);
    for (int i = 0; i < n; i++) lenOffs[i] = U();
    for (int i = 0; i < n; i++) widthOffs[i] = U() / 2 + 0.25;
    e->nObs = 0;
    for (int i = 0; i < n; i++) {
      V2 c = walk_spot(&e->roads[roadIds[i]], sideIds[i], lenOffs[i], widthOffs[i]);
      if (is_off_road(e, c)) e->obst[e->nObs++] = c;
    }
  }
#undef U
}

static void fill_obs(DrvEnv* e, uint32_t tick, DrvStepOut* out) {
  for (int i = 0; i < e->N; i++) {
    if (e->obsType == DYNENV_OBS_FULL) full_obs_agent(e, i, &out->agent[i]);
    else agent_vision(e, i, tick, &out->agent[i]);
  }
}
static void fill_info(DrvEnv* e, DrvStepOut* out) {
  for (int i = 0; i < e->N; i++) car_row(&e->car[i], out->state_cars[i]);
  for (int o = 0; o < e->nObs; o++) {
    out->state_obs[o][0] = normalize_(e->obst[o].x, SNORMX, 0); out->state_obs[o][1] = normalize_(e->obst[o].y, SNORMY, 0);
    out->state_obs[o][2] = normalize_(OBST_HALF, SNORMW, 0.5); out->state_obs[o][3] = normalize_(OBST_HALF, SNORMH, 0.5);
  }
  for (int p = 0; p < e->nPed; p++) {
    out->state_ped[p][0] = normalize_(e->ped[p].b.px, SNORMX, 0); out->state_ped[p][1] = normalize_(e->ped[p].b.py, SNORMY, 0);
  }
  out->nobs = e->nObs; out->nped = e->nPed;
}

void drv_reset_obs(DrvEnv* e, DrvStepOut* out) {
  memset(out, 0, sizeof(*out));
  fill_obs(e, DYNENV_TICK_RESETOBS, out);
  fill_info(e, out);
}

void drv_step(DrvEnv* e, const double* actions /*[N][2]*/, DrvStepOut* out) {
  /* DrivingEnvironment.step (:248-322) */
  memset(out, 0, sizeof(*out));
  e->teamReward = 0;
  for (int i = 0; i < NCAR; i++) { e->carRewards[i] = 0; e->carPosRewards[i] = 0; }
  for (int i = e->subBegin; i < (e->subEnd ? e->subEnd : DRV_SUBSTEPS); i++) {
    for (int c = 0; c < e->N; c++) {
      if (i % 10 == 0) process_action(e, c, actions + 2 * c);
      tick_car(e, c);
    }
    for (int p = 0; p < e->nPed; p++) move_ped(e, p);
    space_step(e, 1.0 / 100.0);
    e->elapsed += 1;
    int allFin = 1;
    for (int c = 0; c < e->N; c++) allFin &= (e->car[c].finished && !e->car[c].crashed);
    if (!e->allFinished && allFin) { e->allFinished = 1; e->teamReward += (MAXTIME - e->elapsed) / 100.0; }
    if (i % 10 == 9) fill_obs(e, (uint32_t)e->elapsed, out);
  }
  for (int c = 0; c < e->N; c++) {
    e->carRewards[c] += e->teamReward;
    e->epR[c] += e->carRewards[c];
    e->carPosRewards[c] += fmax(0.0, e->teamReward);
    e->epPR[c] += e->carPosRewards[c];
    out->rewards[c] = e->carRewards[c];
  }
  fill_info(e, out);
  if (e->elapsed >= MAXTIME) {
    out->done = 1;
    for (int c = 0; c < e->N; c++) {
      out->episode_r[c] = e->epR[c]; out->episode_p_r[c] = e->epPR[c];
      out->episode_g[0] += e->car[c].finished && !e->car[c].crashed;
      out->episode_g[1] += e->car[c].crashed;
    }
  }
}

/* state record <-> DrvEnv; words are doubles holding the float / int value of each 32-bit word */
void drv_store_state(const DrvEnv* e, double* w) {
  for (int i = 0; i < DRV_STATE_WORDS; i++) w[i] = 0;
  for (int i = 0; i < NCAR; i++) {
    const Car* c = &e->car[i];
    w[DRV_S_CPX + i] = c->b.px; w[DRV_S_CPY + i] = c->b.py; w[DRV_S_CANG + i] = c->b.a;
    w[DRV_S_CVX + i] = c->b.vx; w[DRV_S_CVY + i] = c->b.vy; w[DRV_S_CW + i] = c->b.w;
    w[DRV_S_CVBX + i] = c->b.vbx; w[DRV_S_CVBY + i] = c->b.vby; w[DRV_S_CWB + i] = c->b.wb;
    w[DRV_S_CDIRX + i] = c->dir.x; w[DRV_S_CDIRY + i] = c->dir.y;
    w[DRV_S_CPREVX + i] = c->prev.x; w[DRV_S_CPREVY + i] = c->prev.y;
    w[DRV_S_CGOALX + i] = c->goal.x; w[DRV_S_CGOALY + i] = c->goal.y;
    w[DRV_S_CTYPE + i] = c->type; w[DRV_S_CTEAM + i] = c->team;
    w[DRV_S_CFLAGS + i] = (c->finished ? DRV_CF_FINISHED : 0) | (c->crashed ? DRV_CF_CRASHED : 0);
    w[DRV_S_CLANE + i] = c->lane;
    w[DRV_S_EPR + i] = e->epR[i]; w[DRV_S_EPPR + i] = e->epPR[i];
  }
  for (int i = 0; i < NPEDMAX; i++) {
    const Ped* p = &e->ped[i];
    w[DRV_S_PPX + i] = p->b.px; w[DRV_S_PPY + i] = p->b.py; w[DRV_S_PVX + i] = p->b.vx; w[DRV_S_PVY + i] = p->b.vy;
    w[DRV_S_PVBX + i] = p->b.vbx; w[DRV_S_PVBY + i] = p->b.vby;
    w[DRV_S_PMOVING + i] = p->moving; w[DRV_S_PSPEED + i] = p->speed;
    w[DRV_S_PFLAGS + i] = (p->dead ? DRV_PF_DEAD : 0) | (p->crossing ? DRV_PF_CROSSING : 0) |
                          (p->beginCrossing ? DRV_PF_BEGINCROSSING : 0) | (p->side ? DRV_PF_SIDE : 0) | (p->road ? DRV_PF_ROAD : 0);
  }
  for (int i = 0; i < NOBSMAX; i++) { w[DRV_S_OPX + i] = e->obst[i].x; w[DRV_S_OPY + i] = e->obst[i].y; }
  w[DRV_S_ELAPSED] = e->elapsed; w[DRV_S_EPISODE] = e->episode; w[DRV_S_ALLFIN] = e->allFinished;
  w[DRV_S_NPED] = e->nPed; w[DRV_S_NOBS] = e->nObs; w[DRV_S_STEPPED] = e->stepped;
  w[DRV_S_NARB] = e->nArb;
  for (int k = 0; k < e->nArb; k++) {
    const Arbiter* a = &e->arb[k];
    w[DRV_S_ARBKEY + k] = a->a | (a->b << 8) | (a->state << 16) | (a->age << 24);
    int h0 = a->ncon > 0 ? a->con[0].hash : 0, h1 = a->ncon > 1 ? a->con[1].hash : 0;
    w[DRV_S_ARBCON + k] = a->ncon | (h0 << 8) | (h1 << 16);
    w[DRV_S_ARBJN0 + k] = a->ncon > 0 ? a->con[0].jnAcc : 0; w[DRV_S_ARBJT0 + k] = a->ncon > 0 ? a->con[0].jtAcc : 0;
    w[DRV_S_ARBJN1 + k] = a->ncon > 1 ? a->con[1].jnAcc : 0; w[DRV_S_ARBJT1 + k] = a->ncon > 1 ? a->con[1].jtAcc : 0;
  }
}

void drv_load_state(DrvEnv* e, const double* w) {
  roads_init(e);
  for (int i = 0; i < NCAR; i++) {
    Car* c = &e->car[i];
    car_set_type(c, (int)w[DRV_S_CTYPE + i]);
    c->b.px = w[DRV_S_CPX + i] + w[DRV_S_CPXL + i]; c->b.py = w[DRV_S_CPY + i] + w[DRV_S_CPYL + i];
    c->b.a = w[DRV_S_CANG + i] + w[DRV_S_CANGL + i];
    c->b.vx = w[DRV_S_CVX + i]; c->b.vy = w[DRV_S_CVY + i]; c->b.w = w[DRV_S_CW + i];
    c->b.vbx = w[DRV_S_CVBX + i]; c->b.vby = w[DRV_S_CVBY + i]; c->b.wb = w[DRV_S_CWB + i];
    c->dir = v2(w[DRV_S_CDIRX + i], w[DRV_S_CDIRY + i]);
    c->prev = v2(w[DRV_S_CPREVX + i], w[DRV_S_CPREVY + i]);
    c->goal = v2(w[DRV_S_CGOALX + i], w[DRV_S_CGOALY + i]);
    c->team = (int)w[DRV_S_CTEAM + i];
    int fl = (int)w[DRV_S_CFLAGS + i];
    c->finished = !!(fl & DRV_CF_FINISHED); c->crashed = !!(fl & DRV_CF_CRASHED);
    c->lane = (int)w[DRV_S_CLANE + i];
    e->epR[i] = w[DRV_S_EPR + i]; e->epPR[i] = w[DRV_S_EPPR + i];
  }
  for (int i = 0; i < NPEDMAX; i++) {
    Ped* p = &e->ped[i];
    body_set_mass(&p->b, 90, 90 * (PED_R * PED_R) / 2.0);
    p->b.px = w[DRV_S_PPX + i] + w[DRV_S_PPXL + i]; p->b.py = w[DRV_S_PPY + i] + w[DRV_S_PPYL + i];
    p->b.vx = w[DRV_S_PVX + i]; p->b.vy = w[DRV_S_PVY + i]; p->b.vbx = w[DRV_S_PVBX + i]; p->b.vby = w[DRV_S_PVBY + i];
    p->moving = w[DRV_S_PMOVING + i]; p->speed = (int)w[DRV_S_PSPEED + i];
    int fl = (int)w[DRV_S_PFLAGS + i];
    p->dead = !!(fl & DRV_PF_DEAD); p->crossing = !!(fl & DRV_PF_CROSSING); p->beginCrossing = !!(fl & DRV_PF_BEGINCROSSING);
    p->side = !!(fl & DRV_PF_SIDE); p->road = !!(fl & DRV_PF_ROAD);
  }
  for (int i = 0; i < NOBSMAX; i++) e->obst[i] = v2(w[DRV_S_OPX + i], w[DRV_S_OPY + i]);
  e->elapsed = (int)w[DRV_S_ELAPSED]; e->episode = (int)w[DRV_S_EPISODE]; e->allFinished = (int)w[DRV_S_ALLFIN];
  e->nPed = (int)w[DRV_S_NPED]; e->nObs = (int)w[DRV_S_NOBS]; e->stepped = (int)w[DRV_S_STEPPED];
  e->nArb = (int)w[DRV_S_NARB];
  for (int k = 0; k < e->nArb; k++) {
    Arbiter* a = &e->arb[k];
    memset(a, 0, sizeof(*a));
    int key = (int)w[DRV_S_ARBKEY + k], cn = (int)w[DRV_S_ARBCON + k];
    a->a = key & 255; a->b = (key >> 8) & 255; a->state = (key >> 16) & 255; a->age = (key >> 24) & 255;
    a->ncon = cn & 255; a->con[0].hash = (cn >> 8) & 255; a->con[1].hash = (cn >> 16) & 255;
    a->con[0].jnAcc = w[DRV_S_ARBJN0 + k]; a->con[0].jtAcc = w[DRV_S_ARBJT0 + k];
    a->con[1].jnAcc = w[DRV_S_ARBJN1 + k]; a->con[1].jtAcc = w[DRV_S_ARBJT1 + k];
  }
}
