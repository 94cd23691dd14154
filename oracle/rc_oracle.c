/*
 * ORACLE / TEST INFRASTRUCTURE ONLY — CPU restatement of DynEnv's RoboCup rollout hot path.
 *
 * Sequential, double precision, one environment per RcEnv.  It follows the reference
 * (szemenyeim/DynEnv) function by function; every block cites the reference file:line it
 * restates (paths under DynEnv/).  The rigid-body part restates the Chipmunk2D 7.0.x
 * subset that pymunk 5.x exposes to the reference (cpSpaceStep.c, cpArbiter.c,
 * cpCollision.c, cpPivotJoint.c, cpRotaryLimitJoint.c, cpBody.c); that dependency is NOT in
 * /root/reference and not installable here: PARITY UNPINNED at the pymunk boundary.
 * Everything above that boundary (rules, rewards, vision, noise, encoders) is pinned against
 * the reference's own Python run on oracle/shim (tests/golden, tests/test_oracle_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may call this.  The product path (dynenv_b200/csrc) never does.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dynenv_b200.h"
#include "../include/dynenv_rng.h"
#include "oracle_rng.h"

#define PI 3.14159265358979323846

/* ---- field constants (RoboCupEnvironment.py:26-29,73-93; Robot.py:13-20) */
#define W_ 1040.0
#define H_ 740.0
#define SIDE 70.0
#define LINEW 5.0
#define PEN_RADIUS 5.0
#define PEN_LENGTH 60.0
#define PEN_WIDTH 110.0
#define PEN_DIST 130.0
#define CC_RADIUS 75.0
#define GOAL_W 80.0
#define POST_RADIUS 5.0
#define BALL_RADIUS 5.0
#define FIELD_W 900.0
#define FIELD_H 600.0
#define ROB_LEN 10.0
#define ROB_RAD 7.5
#define ROB_TOTAL 17.5
#define ROB_FOV (PI / 4)
#define ROB_VEL 50.0
#define ROB_ANGVEL 20.0
#define ROB_MASS 4000.0
#define HEAD_MAX (2 * PI / 3)
#define MAXTIME 12000
#define NSHAPE 25 /* feet 0..19, ball 20, posts 21..24 */
#define BALL 20
#define POST0 21
#define LISTCAP 24

typedef struct { double x, y; } V2;
static inline V2 v2(double x, double y) { V2 r = {x, y}; return r; }
static inline V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 vsub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 vmul(V2 a, double s) { return v2(a.x * s, a.y * s); }
static inline double vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline double vcross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline double vlen(V2 a) { return sqrt(a.x * a.x + a.y * a.y); }
static inline double vlensq(V2 a) { return a.x * a.x + a.y * a.y; }
static inline V2 vrot(V2 a, double ang) { /* Vec2d.rotated */
  double c = cos(ang), s = sin(ang);
  return v2(a.x * c - a.y * s, a.x * s + a.y * c);
}

typedef struct {
  double px, py, a, vx, vy, w, fx, fy, t, vbx, vby, wb;
  double m_inv, i_inv, m;
} Body;

typedef struct {
  V2 ta, tb, tn;   /* cached world geometry (circle: ta = centre) */
  double l, b, r, t; /* cached bounding box */
} ShapeCache;

typedef struct {
  V2 p1, p2;       /* absolute contact points on a and b (before conversion) */
  V2 r1, r2;
  int hash;
  double jnAcc, jtAcc, jBias, nMass, tMass, bias, bounce;
} Contact;

typedef struct {
  int a, b;      /* shape ids, ordered by primitive type (circle first) then insertion */
  int state, age;
  int ncon;
  Contact con[2];
  V2 n;
  double e, u;
  int touched;   /* collided in the current substep */
} Arbiter;

typedef struct {
  double headAngle, headMoving, moveTime, fallTime, penalTime;
  V2 initPos, prevPos;
  V2 pivJ; double rotJ;
  /* joint pre-step scratch */
  V2 pivBias; double rotBias, rotISum, pivK;
  int kicking, foot, fallen, penalized, touching, mightPush, jointRemoved, jointAbsent, jointAddPending;
  int fallCntr, touchCntr;
  int team, id;
} Robot;

typedef struct {
  int nb, nr, np, nc, nf, nl;
  double ball[LISTCAP][5];
  double rob[LISTCAP][7];
  double post[LISTCAP][6];
  double cross[LISTCAP][6];
  double fcross[LISTCAP][8];
  double line[LISTCAP][5];
} ObsLists;

struct RcEnv {
  /* configuration */
  int nPlayers, P, obsType, noiseType, allowHeadTurn;
  double noiseMag, randBase;
  int randomInit, deterministicTurn, canFall, useObsRewards;
  OracleRng rng;
  /* state */
  Body body[NSHAPE];
  ShapeCache sc[NSHAPE];
  Robot rob[RC_MAXP];
  V2 ballPrev;
  int lastKicked[4], nLastKicked;
  int elapsed, episode, ballOwned;
  double ballFreeCntr, gracePeriod;
  int goals[2], closestID[2], defenders[2][2], nDef[2], penalTimes[2];
  int stepped;
  Arbiter arb[RC_MAXARB];
  int nArb;
  double epR[RC_MAXP], epPR[RC_MAXP], epOR[RC_MAXP];
  /* per-step scratch */
  double teamRewards[2], obsRewards[RC_MAXP], robotRewards[RC_MAXP], robotPosRewards[RC_MAXP];
  int active[RC_MAXARB], nActive;
  int arbOverflow;
  int subBegin, subEnd; /* debugging aid: substeps run by rc_step (0 = unset -> 0..RC_SUBSTEPS) */
  struct RcRawOut* rawOut; /* optional: raw detection lists of every observation (IMAGE raster oracle) */
};
typedef struct RcEnv RcEnv;

/* ---------------------------------------------------------------- static scene data */
typedef struct { V2 p1, p2; double tx, ty; } FieldLine;
typedef struct { V2 p; double t1, t2, t3; } FieldMark;
static FieldLine g_lines[11];
static FieldMark g_pcross[3], g_lcross[16];
static struct { V2 p; double side, dir; } g_posts[4];
static V2 g_penaltySpots[2][14];
static int g_scene_ready = 0;

static void scene_init(void) {
  if (g_scene_ready) return;
  /* RoboCupEnvironment.py:153-195 */
  int i = 0;
  g_lines[i++] = (FieldLine){{SIDE, SIDE}, {SIDE, H_ - SIDE}, 1, 0};
  g_lines[i++] = (FieldLine){{W_ - SIDE, SIDE}, {W_ - SIDE, H_ - SIDE}, -1, 0};
  g_lines[i++] = (FieldLine){{SIDE, SIDE}, {W_ - SIDE, SIDE}, 0, 1};
  g_lines[i++] = (FieldLine){{SIDE, H_ - SIDE}, {W_ - SIDE, H_ - SIDE}, 0, -1};
  g_lines[i++] = (FieldLine){{W_ / 2, SIDE}, {W_ / 2, H_ - SIDE}, 0, 0};
  g_lines[i++] = (FieldLine){{SIDE, H_ / 2 - PEN_WIDTH}, {SIDE + PEN_LENGTH, H_ / 2 - PEN_WIDTH}, 1, 0.37};
  g_lines[i++] = (FieldLine){{SIDE, H_ / 2 + PEN_WIDTH}, {SIDE + PEN_LENGTH, H_ / 2 + PEN_WIDTH}, 1, -0.37};
  g_lines[i++] = (FieldLine){{SIDE + PEN_LENGTH, H_ / 2 - PEN_WIDTH}, {SIDE + PEN_LENGTH, H_ / 2 + PEN_WIDTH}, 0.87, 0};
  g_lines[i++] = (FieldLine){{W_ - SIDE - PEN_LENGTH, H_ / 2 - PEN_WIDTH}, {W_ - SIDE, H_ / 2 - PEN_WIDTH}, -1, 0.37};
  g_lines[i++] = (FieldLine){{W_ - SIDE - PEN_LENGTH, H_ / 2 + PEN_WIDTH}, {W_ - SIDE, H_ / 2 + PEN_WIDTH}, -1, -0.37};
  g_lines[i++] = (FieldLine){{W_ - SIDE - PEN_LENGTH, H_ / 2 - PEN_WIDTH}, {W_ - SIDE - PEN_LENGTH, H_ / 2 + PEN_WIDTH}, -0.87, 0};
  /* :199-203 */
  g_pcross[0] = (FieldMark){{W_ / 2, H_ / 2}, 0, 0, 0};
  g_pcross[1] = (FieldMark){{SIDE + PEN_DIST, H_ / 2}, 1, 0, 0};
  g_pcross[2] = (FieldMark){{W_ - (SIDE + PEN_DIST), H_ / 2}, -1, 0, 0};
  /* :205-227 */
  i = 0;
  g_lcross[i++] = (FieldMark){{SIDE, SIDE}, 1, 1, 0};
  g_lcross[i++] = (FieldMark){{SIDE, H_ - SIDE}, 1, -1, 0};
  g_lcross[i++] = (FieldMark){{W_ - SIDE, SIDE}, -1, 1, 0};
  g_lcross[i++] = (FieldMark){{W_ - SIDE, H_ - SIDE}, -1, -1, 0};
  g_lcross[i++] = (FieldMark){{W_ / 2, SIDE}, 0, 1, 0};
  g_lcross[i++] = (FieldMark){{W_ / 2, H_ - SIDE}, 0, -1, 0};
  g_lcross[i++] = (FieldMark){{W_ / 2, H_ / 2 - CC_RADIUS * 2}, 0, 0.5, 0};
  g_lcross[i++] = (FieldMark){{W_ / 2, H_ / 2 + CC_RADIUS * 2}, 0, -0.5, 0};
  g_lcross[i++] = (FieldMark){{SIDE, H_ / 2 - PEN_WIDTH}, 1, 0.37, 0};
  g_lcross[i++] = (FieldMark){{SIDE, H_ / 2 + PEN_WIDTH}, 1, -0.37, 0};
  g_lcross[i++] = (FieldMark){{SIDE + PEN_LENGTH, H_ / 2 - PEN_WIDTH}, 0.87, 0.37, 0};
  g_lcross[i++] = (FieldMark){{SIDE + PEN_LENGTH, H_ / 2 + PEN_WIDTH}, 0.87, -0.37, 0};
  g_lcross[i++] = (FieldMark){{W_ - SIDE, H_ / 2 - PEN_WIDTH}, -1, 0.37, 0};
  g_lcross[i++] = (FieldMark){{W_ - SIDE, H_ / 2 + PEN_WIDTH}, -1, -0.37, 0};
  g_lcross[i++] = (FieldMark){{W_ - SIDE - PEN_LENGTH, H_ / 2 - PEN_WIDTH}, -0.87, 0.37, 0};
  g_lcross[i++] = (FieldMark){{W_ - SIDE - PEN_LENGTH, H_ / 2 + PEN_WIDTH}, -0.87, -0.37, 0};
  /* :296-302 */
  g_posts[0].p = v2(SIDE, H_ / 2 + GOAL_W); g_posts[0].side = 1; g_posts[0].dir = -0.27;
  g_posts[1].p = v2(SIDE, H_ / 2 - GOAL_W); g_posts[1].side = 1; g_posts[1].dir = 0.27;
  g_posts[2].p = v2(W_ - SIDE, H_ / 2 + GOAL_W); g_posts[2].side = -1; g_posts[2].dir = -0.27;
  g_posts[3].p = v2(W_ - SIDE, H_ / 2 - GOAL_W); g_posts[3].side = -1; g_posts[3].dir = 0.27;
  /* :139-151 */
  for (int k = 0; k < 7; k++) {
    g_penaltySpots[0][k] = v2(SIDE + (k + 1) * ROB_TOTAL * 3, SIDE);
    g_penaltySpots[0][7 + k] = v2(SIDE + (k + 1) * ROB_TOTAL * 3, H_ - SIDE);
    g_penaltySpots[1][k] = v2(W_ - SIDE - (k + 1) * ROB_TOTAL * 3, SIDE);
    g_penaltySpots[1][7 + k] = v2(W_ - SIDE - (k + 1) * ROB_TOTAL * 3, H_ - SIDE);
  }
  g_scene_ready = 1;
}

/* ---------------------------------------------------------------- helpers */
static inline int is_foot(int s) { return s < 20; }
static inline double shape_radius(int s) { return s < 20 ? ROB_RAD : 10.0; /* Ball.py:13, Goalpost.py:11: 2x nominal */ }
static inline double shape_e(int s) { return s < 20 ? 0.3 : (s == BALL ? 0.98 : 0.95); }
static inline double shape_u(int s) { return s < 20 ? 2.5 : (s == BALL ? 3.0 : 0.0); }

static inline double rnd(RcEnv* e, uint32_t tick, int site, int a, int b, int k) {
  return oracle_uniform(&e->rng, (uint32_t)e->episode, tick, site, a, b, k);
}
static inline int rnd_int(RcEnv* e, uint32_t tick, int site, int a, int b, int k, int lo, int hi) {
  /* random.randint(lo, hi) as derived by UniformTape */
  return lo + (int)(rnd(e, tick, site, a, b, k) * (hi - lo + 1));
}

static V2 robot_pos(const RcEnv* e, int r) { /* Robot.getPos (Robot.py:90-91) */
  const Body* L = &e->body[2 * r];
  const Body* R = &e->body[2 * r + 1];
  return v2((L->px + R->px) / 2.0, (L->py + R->py) / 2.0);
}
static double robot_angle(const RcEnv* e, int r) { /* Robot.getAngle (Robot.py:94-100) */
  return (e->body[2 * r].a + e->body[2 * r + 1].a) / 2.0;
}
static V2 ball_pos(const RcEnv* e) { return v2(e->body[BALL].px, e->body[BALL].py); }

static void shape_cache_update(RcEnv* e, int s) { /* cpShapeCacheBB */
  Body* b = &e->body[s];
  ShapeCache* c = &e->sc[s];
  double r = shape_radius(s);
  if (is_foot(s)) {
    double ly = (s & 1) ? -ROB_LEN : ROB_LEN; /* Robot.py:24-27: left foot y=+10, right y=-10 */
    double cs = cos(b->a), sn = sin(b->a);
    c->ta = v2(b->px + (-ROB_LEN) * cs - ly * sn, b->py + (-ROB_LEN) * sn + ly * cs);
    c->tb = v2(b->px + (ROB_LEN)*cs - ly * sn, b->py + (ROB_LEN)*sn + ly * cs);
    /* seg->n = rperp(normalize(b - a)) = (0,-1) local */
    c->tn = v2(0 * cs - (-1.0) * sn, 0 * sn + (-1.0) * cs);
    c->l = fmin(c->ta.x, c->tb.x) - r;
    c->r = fmax(c->ta.x, c->tb.x) + r;
    c->b = fmin(c->ta.y, c->tb.y) - r;
    c->t = fmax(c->ta.y, c->tb.y) + r;
  } else {
    c->ta = v2(b->px, b->py);
    c->tb = c->ta;
    c->tn = v2(0, 0);
    c->l = b->px - r; c->r = b->px + r; c->b = b->py - r; c->t = b->py + r;
  }
}

/* ---------------------------------------------------------------- narrow phase */
static V2 closest_on_seg(V2 p, V2 a, V2 b) {
  V2 d = vsub(b, a);
  double t = vdot(d, vsub(p, a)) / vlensq(d);
  t = t < 0 ? 0 : (t > 1 ? 1 : t);
  return vadd(a, vmul(d, t));
}

static int collide_circle_circle(V2 p1, V2 p2, double r1, double r2, V2* n, Contact* out) {
  /* CircleToCircleQuery (cpCollision.c) */
  double mind = r1 + r2;
  V2 d = vsub(p2, p1);
  double dsq = vlensq(d);
  if (dsq < mind * mind) {
    double dist = sqrt(dsq);
    *n = dist ? vmul(d, 1.0 / dist) : v2(1, 0);
    out[0].p1 = vadd(p1, vmul(*n, r1));
    out[0].p2 = vadd(p2, vmul(*n, -r2));
    out[0].hash = 0;
    return 1;
  }
  return 0;
}

static int collide_circle_segment(const ShapeCache* c, double cr, const ShapeCache* s, double sr, V2* n,
                                  Contact* out) {
  /* CircleToSegment (cpCollision.c); end-cap tangents are zero => never rejected */
  V2 sd = vsub(s->tb, s->ta);
  double t = vdot(sd, vsub(c->ta, s->ta)) / vlensq(sd);
  t = t < 0 ? 0 : (t > 1 ? 1 : t);
  V2 closest = vadd(s->ta, vmul(sd, t));
  double mind = cr + sr;
  V2 d = vsub(closest, c->ta);
  double dsq = vlensq(d);
  if (dsq < mind * mind) {
    double dist = sqrt(dsq);
    *n = dist ? vmul(d, 1.0 / dist) : s->tn;
    out[0].p1 = vadd(c->ta, vmul(*n, cr));
    out[0].p2 = vadd(closest, vmul(*n, -sr));
    out[0].hash = 0;
    return 1;
  }
  return 0;
}

static int segs_intersect(V2 p1, V2 p2, V2 p3, V2 p4) {
  double d1 = vcross(vsub(p4, p3), vsub(p1, p3));
  double d2 = vcross(vsub(p4, p3), vsub(p2, p3));
  double d3 = vcross(vsub(p2, p1), vsub(p3, p1));
  double d4 = vcross(vsub(p2, p1), vsub(p4, p1));
  return (d1 * d2 < 0) && (d3 * d4 < 0);
}

/* ClosestPoints (n, d) of GJK/EPA for two core segments: separated -> distance and direction of
   the closest feature pair; crossing -> minimum-penetration axis among the four edge normals of
   the Minkowski parallelogram. */
static void seg_seg_closest(V2 a0, V2 a1, V2 b0, V2 b1, V2* nOut, double* dOut) {
  if (segs_intersect(a0, a1, b0, b1)) {
    V2 da = vsub(a1, a0), db = vsub(b1, b0);
    double la = vlen(da), lb = vlen(db);
    V2 axA[2] = {v2(da.y / la, -da.x / la), v2(-da.y / la, da.x / la)};
    V2 axB[2] = {v2(db.y / lb, -db.x / lb), v2(-db.y / lb, db.x / lb)};
    double best = -INFINITY;
    V2 bn = axA[0];
    for (int k = 0; k < 2; k++) { /* candidate n = +axis of A */
      double mnB = fmin(vdot(axA[k], b0), vdot(axA[k], b1));
      double mxA = fmax(vdot(axA[k], a0), vdot(axA[k], a1));
      double d = mnB - mxA;
      if (d > best) { best = d; bn = axA[k]; }
    }
    for (int k = 0; k < 2; k++) { /* candidate n = -axis of B */
      double mnA = fmin(vdot(axB[k], a0), vdot(axB[k], a1));
      double mxB = fmax(vdot(axB[k], b0), vdot(axB[k], b1));
      double d = mnA - mxB;
      if (d > best) { best = d; bn = v2(-axB[k].x, -axB[k].y); }
    }
    *nOut = bn; *dOut = best;
    return;
  }
  double best = INFINITY;
  V2 bn = v2(1, 0);
  V2 A[2] = {a0, a1}, B[2] = {b0, b1};
  for (int k = 0; k < 2; k++) { /* vertex of A against segment B: vector p -> c points A -> B */
    V2 c = closest_on_seg(A[k], b0, b1);
    V2 dv = vsub(c, A[k]);
    double d = vlen(dv);
    if (d < best) { best = d; if (d > 0) bn = vmul(dv, 1.0 / d); }
  }
  for (int k = 0; k < 2; k++) { /* vertex of B against segment A: c -> p points A -> B */
    V2 c = closest_on_seg(B[k], a0, a1);
    V2 dv = vsub(B[k], c);
    double d = vlen(dv);
    if (d < best) { best = d; if (d > 0) bn = vmul(dv, 1.0 / d); }
  }
  *nOut = bn; *dOut = best;
}

static inline double clamp01(double t) { return t < 0 ? 0 : (t > 1 ? 1 : t); }
static inline V2 vlerp(V2 a, V2 b, double t) { return vadd(vmul(a, 1.0 - t), vmul(b, t)); }

static int collide_segment_segment(const ShapeCache* s1, const ShapeCache* s2, V2* nOut, Contact* out) {
  /* SegmentToSegment + SupportEdgeForSegment + ContactPoints (cpCollision.c) */
  V2 n; double d;
  seg_seg_closest(s1->ta, s1->tb, s2->ta, s2->tb, &n, &d);
  double r1 = ROB_RAD, r2 = ROB_RAD;
  if (!(d <= r1 + r2)) return 0;
  V2 e1a, e1b, e2a, e2b; int h1a, h1b, h2a, h2b;
  if (vdot(s1->tn, n) > 0.0) { e1a = s1->ta; h1a = 0; e1b = s1->tb; h1b = 1; }
  else { e1a = s1->tb; h1a = 1; e1b = s1->ta; h1b = 0; }
  V2 nn = v2(-n.x, -n.y);
  if (vdot(s2->tn, nn) > 0.0) { e2a = s2->ta; h2a = 0; e2b = s2->tb; h2b = 1; }
  else { e2a = s2->tb; h2a = 1; e2b = s2->ta; h2b = 0; }
  *nOut = n;
  int cnt = 0;
  double d_e1_a = vcross(e1a, n), d_e1_b = vcross(e1b, n);
  double d_e2_a = vcross(e2a, n), d_e2_b = vcross(e2b, n);
  double e1_denom = 1.0 / (d_e1_b - d_e1_a + 2.2250738585072014e-308);
  double e2_denom = 1.0 / (d_e2_b - d_e2_a + 2.2250738585072014e-308);
  {
    V2 p1 = vadd(vmul(n, r1), vlerp(e1a, e1b, clamp01((d_e2_b - d_e1_a) * e1_denom)));
    V2 p2 = vadd(vmul(n, -r2), vlerp(e2a, e2b, clamp01((d_e1_a - d_e2_a) * e2_denom)));
    if (vdot(vsub(p2, p1), n) <= 0.0) {
      out[cnt].p1 = p1; out[cnt].p2 = p2; out[cnt].hash = 1 + (h1a | (h2b << 1)); cnt++;
    }
  }
  {
    V2 p1 = vadd(vmul(n, r1), vlerp(e1a, e1b, clamp01((d_e2_a - d_e1_a) * e1_denom)));
    V2 p2 = vadd(vmul(n, -r2), vlerp(e2a, e2b, clamp01((d_e1_b - d_e2_a) * e2_denom)));
    if (vdot(vsub(p2, p1), n) <= 0.0) {
      out[cnt].p1 = p1; out[cnt].p2 = p2; out[cnt].hash = 1 + (h1b | (h2a << 1)); cnt++;
    }
  }
  return cnt;
}

/* ---------------------------------------------------------------- game rules */
static void penalize(RcEnv* e, int r);

static double point_query_dist(const RcEnv* e, int s, V2 p) { /* cpShapePointQuery distance */
  const ShapeCache* c = &e->sc[s];
  if (is_foot(s)) return vlen(vsub(p, closest_on_seg(p, c->ta, c->tb))) - ROB_RAD;
  return vlen(vsub(p, c->ta)) - 10.0;
}

static void fall(RcEnv* e, int r, int punish) {
  /* RoboCupEnvironment.py:735-791 */
  Robot* rb = &e->rob[r];
  V2 pos = robot_pos(e, r);
  if (punish) e->robotRewards[r] -= 2;
  /* space.point_query(pos, 40, filter): every shape whose surface is closer than 40 px, using
     the shape geometry cached at the last collision phase.  Static posts receive inf/NaN
     forces in the reference, which Chipmunk never integrates: skipped here. */
  for (int s = 0; s <= BALL; s++) {
    if (s < 20 && s >= 2 * e->P) continue;
    if (s == 2 * r || s == 2 * r + 1) continue;
    if (!(point_query_dist(e, s, pos) < 40.0)) continue;
    Body* ob = &e->body[s];
    double force = ROB_VEL * ROB_MASS * ob->m / 50.0;
    V2 dp = vsub(pos, v2(ob->px, ob->py));
    double len = vlen(dp);
    dp = v2(-dp.x * force / len, -dp.y * force / len);
    /* apply_force_at_world_point(dp, pos) */
    ob->fx += dp.x; ob->fy += dp.y;
    ob->t += (pos.x - ob->px) * dp.y - (pos.y - ob->py) * dp.x;
    if (s == BALL) {
      int found = 0;
      for (int k = 0; k < e->nLastKicked; k++) found |= (e->lastKicked[k] == r);
      if (e->nLastKicked && !found) {
        int n = e->nLastKicked < 4 ? e->nLastKicked : 3;
        for (int k = n; k > 0; k--) e->lastKicked[k] = e->lastKicked[k - 1];
        e->lastKicked[0] = r;
        e->nLastKicked = n + 1;
      }
      if (e->ballOwned != 0) { e->gracePeriod = 0; e->ballFreeCntr = 0; e->ballOwned = 0; }
    }
  }
  rb->fallen = 1;
  rb->fallCntr += 1;
  rb->fallTime = 4000;
  if (rb->fallCntr > 2) penalize(e, r);
}

static void free_penalty_spot(const RcEnv* e, int r, V2* spot, double* angle) {
  /* RoboCupEnvironment.py:793-821 */
  const V2* spots = e->rob[r].team > 0 ? g_penaltySpots[0] : g_penaltySpots[1];
  double y = e->body[BALL].py;
  *angle = y < H_ / 2 ? -PI / 2 : PI / 2;
  const V2* avail = y > H_ / 2 ? spots : spots + 7;
  V2 sel = avail[0];
  for (int k = 0; k < 7; k++) {
    int ok = 1;
    for (int o = 0; o < e->P; o++) {
      if (o == r) continue;
      if (vlen(vsub(avail[k], robot_pos(e, o))) < ROB_TOTAL * 3) { ok = 0; break; }
    }
    if (ok) { sel = avail[k]; break; }
  }
  *spot = sel;
}

static void penalize(RcEnv* e, int r) {
  /* RoboCupEnvironment.py:824-859 */
  Robot* rb = &e->rob[r];
  rb->penalized = 1;
  int teamIdx = rb->team > 0 ? 0 : 1;
  rb->penalTime = e->penalTimes[teamIdx];
  e->robotRewards[r] -= e->penalTimes[teamIdx] / 2000.0;
  e->penalTimes[teamIdx] += 10000;
  V2 pos; double ang;
  free_penalty_spot(e, r, &pos, &ang);
  for (int f = 0; f < 2; f++) {
    Body* b = &e->body[2 * r + f];
    b->px = pos.x; b->py = pos.y; b->a = ang;
    b->vx = 0; b->vy = 0; b->w = 0;
  }
  if (rb->kicking && rb->jointRemoved) {
    rb->kicking = 0;
    rb->jointRemoved = 0;
    /* space.add(joint): immediate outside the step, deferred to the end of the step when
       issued from a collision callback (pymunk Space.add while locked) */
    rb->jointAddPending = 1;
  }
}

static void joint_add_now(RcEnv* e) {
  for (int r = 0; r < e->P; r++)
    if (e->rob[r].jointAddPending) { e->rob[r].jointAddPending = 0; e->rob[r].jointAbsent = 0; }
}

static void process_action(RcEnv* e, int r, const double* action) {
  /* RoboCupEnvironment.py:527-581, Robot.py:103-138 */
  Robot* rb = &e->rob[r];
  double move = action[0], turn = action[1], kick = action[2], head = action[3];
  if (e->deterministicTurn) head = -3 * rb->team;
  if (!e->allowHeadTurn) head -= 3;
  int canMove = !(rb->penalized || rb->kicking || rb->fallen);
  uint32_t tick = (uint32_t)e->elapsed;
  if (move > 0 && canMove) {
    double u = e->canFall ? rnd(e, tick, RC_SITE_ACTION, r, 0, 0) : 0;
    if (u > 0.999) { fall(e, r, 0); return; }
    /* Robot.step(move - 1) */
    int dir = (int)move - 1;
    rb->moveTime = 500;
    V2 v = dir == 0 ? v2(0, 2 * ROB_VEL) : dir == 1 ? v2(0, -2 * ROB_VEL) : dir == 2 ? v2(2.5 * ROB_VEL, 0)
                                                                                      : v2(-2 * ROB_VEL, 0);
    Body* L = &e->body[2 * r];
    v = vrot(v, L->a);
    L->vx = v.x; L->vy = v.y;
  }
  if (turn > 0 && canMove) {
    double u = e->canFall ? rnd(e, tick, RC_SITE_ACTION, r, 0, 1) : 0;
    if (u > 0.999) { fall(e, r, 0); return; }
    /* Robot.turn(turn - 1) */
    rb->moveTime = 500;
    e->body[2 * r].w += ((int)turn - 1) ? ROB_ANGVEL : -ROB_ANGVEL;
  }
  if (head != 0) { /* Robot.turnHead */
    rb->headMoving = head * PI / 720;
    rb->moveTime = 500;
  }
  if (kick > 0 && move == 0 && turn == 0 && canMove) {
    double u = e->canFall ? rnd(e, tick, RC_SITE_ACTION, r, 0, 2) : 0;
    if (u > 0.99) { fall(e, r, 0); return; }
    /* Robot.kick(kick - 1) */
    rb->foot = (int)kick - 1;
    Body* fb = &e->body[2 * r + (rb->foot ? 1 : 0)];
    rb->initPos = v2(fb->px, fb->py);
    rb->kicking = 1;
    rb->moveTime = 1000;
  }
}

static int in_list2(const int* l, int n, int v) {
  for (int k = 0; k < n; k++) if (l[k] == v) return 1;
  return 0;
}

static void tick_robot(RcEnv* e, int r) {
  /* RoboCupEnvironment.py:862-1007 */
  Robot* rb = &e->rob[r];
  const double time = 10.0;
  Body* L = &e->body[2 * r];
  Body* R = &e->body[2 * r + 1];
  if (rb->moveTime > 0) {
    rb->moveTime -= time;
    if (rb->headMoving != 0) {
      rb->headAngle += rb->headMoving;
      rb->headAngle = fmax(-HEAD_MAX, fmin(HEAD_MAX, rb->headAngle));
    }
    if (rb->kicking) {
      Body* foot = rb->foot ? R : L;
      if (rb->moveTime + time > 500 && rb->moveTime <= 500) {
        if (!rb->jointRemoved) { rb->jointRemoved = 1; rb->jointAbsent = 1; }
        V2 v = vrot(v2(ROB_VEL * 3.0, 0), foot->a);
        foot->vx = v.x; foot->vy = v.y;
      }
      if (rb->moveTime + time > 400 && rb->moveTime <= 400) {
        V2 v = vrot(v2(ROB_VEL * 2.5, 0), foot->a);
        foot->vx = -v.x; foot->vy = -v.y;
      } else if (rb->moveTime <= 300) {
        foot->vx = 0; foot->vy = 0;
        rb->kicking = 0;
        foot->px = rb->initPos.x; foot->py = rb->initPos.y;
        if (rb->jointRemoved) { rb->jointRemoved = 0; rb->jointAbsent = 0; }
      }
    }
    if (rb->moveTime <= 0) {
      rb->moveTime = 0;
      rb->headMoving = 0;
      L->vx = 0; L->vy = 0; L->w = 0;
      R->vx = 0; R->vy = 0; R->w = 0;
    }
  }
  if (rb->fallen) {
    rb->fallTime -= time;
    if (rb->fallTime < 0) {
      double u = rnd(e, (uint32_t)e->elapsed, RC_SITE_REFALL, r, 0, 0);
      if (u > 0.9 && !rb->penalized && e->canFall) { fall(e, r, 0); return; }
      rb->fallen = 0;
      rb->fallCntr = 0;
    }
  }
  if (rb->penalized) {
    rb->penalTime -= time;
    if (rb->penalTime <= 0) {
      rb->penalTime = 0; rb->penalized = 0; rb->fallCntr = 0; rb->fallen = 0;
      V2 pos; double ang;
      free_penalty_spot(e, r, &pos, &ang);
      L->px = pos.x; L->py = pos.y; L->a = ang;
      R->px = pos.x; R->py = pos.y; R->a = ang;
    }
  } else {
    int teamIdx = rb->team > 0 ? 0 : 1;
    V2 pos = robot_pos(e, r);
    double robX = teamIdx ? W_ - pos.x : pos.x;
    double penX = SIDE + PEN_LENGTH + LINEW / 2;
    int inDef = in_list2(e->defenders[teamIdx], e->nDef[teamIdx], r);
    if (robX < penX && pos.y > (H_ / 2 - PEN_WIDTH) && pos.y < (H_ / 2 + PEN_WIDTH)) {
      if (!inDef) {
        if (e->nDef[teamIdx] >= 2) penalize(e, r);
        else e->defenders[teamIdx][e->nDef[teamIdx]++] = r;
      }
    } else if (inDef) { /* list.remove: first occurrence, order kept */
      int n = e->nDef[teamIdx], w = 0, removed = 0;
      for (int q = 0; q < n; q++) {
        if (!removed && e->defenders[teamIdx][q] == r) { removed = 1; continue; }
        e->defenders[teamIdx][w++] = e->defenders[teamIdx][q];
      }
      e->nDef[teamIdx] = w;
    }
  }
  V2 pos = robot_pos(e, r);
  if (pos.y < 0 || pos.x < 0 || pos.y > H_ || pos.x > W_) penalize(e, r);
  if (pos.x != rb->prevPos.x || pos.y != rb->prevPos.y) {
    if ((r == e->closestID[0] || r == e->closestID[1]) && !rb->penalized) {
      V2 bp = ball_pos(e);
      double diff = vlen(vsub(pos, bp)) - vlen(vsub(rb->prevPos, bp));
      e->robotRewards[r] -= diff * 0.05;
      e->robotPosRewards[r] += fmax(0.0, -diff * 0.05);
    }
    rb->prevPos = pos;
  }
}

static void ball_free_kick(RcEnv* e, int team) {
  /* RoboCupEnvironment.py:600-619 */
  if (team == 0) {
    const double time = 10.0;
    if (e->gracePeriod > 0) {
      e->gracePeriod -= time;
      if (e->gracePeriod < 0) { e->gracePeriod = 0; e->ballFreeCntr = 9999; }
    } else if (e->ballFreeCntr > 0) {
      e->ballFreeCntr -= time;
      if (e->ballFreeCntr < 0) { e->ballFreeCntr = 0; e->ballOwned = 0; }
    }
  } else {
    e->ballOwned = team;
    e->gracePeriod = 14999;
    e->ballFreeCntr = 0;
  }
}

static void ball_out_of_field(RcEnv* e) {
  /* RoboCupEnvironment.py:622-732 */
  int finished = 0;
  Body* bb = &e->body[BALL];
  V2 pos = v2(bb->px, bb->py); /* a copy: stays at the old place after the ball is moved */
  double currReward[2] = {0, 0};
  double outMin = SIDE - BALL_RADIUS;
  double outMaxX = W_ - SIDE + BALL_RADIUS;
  double outMaxY = H_ - SIDE + BALL_RADIUS;
  int team = 0;
  if (pos.y < outMin || pos.x < outMin || pos.y > outMaxY || pos.x > outMaxX) {
    double x = W_ / 2, y = H_ / 2;
    team = e->nLastKicked ? e->rob[e->lastKicked[0]].team : 1;
    if (pos.y < outMin || pos.y > outMaxY) {
      x = team < 0 ? pos.x + 50 : pos.x - 50;
      y = pos.y < outMin ? outMin + BALL_RADIUS : outMaxY - BALL_RADIUS;
    } else {
      if (pos.y < H_ / 2 + GOAL_W && pos.y > H_ / 2 - GOAL_W) {
        finished = 1;
        if (pos.x < outMin) { currReward[0] += -25; currReward[1] += 25; e->goals[1] += 1; }
        else { currReward[0] += 25; currReward[1] += -25; e->goals[0] += 1; }
      } else {
        if (pos.x < outMin) {
          if (team < 0) x = SIDE + PEN_LENGTH;
          else { x = SIDE; y = pos.y < H_ / 2 ? SIDE : H_ - SIDE; }
        } else {
          if (team > 0) x = W_ - (SIDE + PEN_LENGTH);
          else { x = W_ - SIDE; y = pos.y < H_ / 2 ? SIDE : H_ - SIDE; }
        }
      }
    }
    bb->px = x; bb->py = y; bb->vx = 0; bb->vy = 0; bb->w = 0;
  }
  ball_free_kick(e, -team);
  if (!finished) {
    currReward[0] += (bb->px - e->ballPrev.x) / 20;
    currReward[1] -= (bb->px - e->ballPrev.x) / 20;
  }
  e->ballPrev = v2(bb->px, bb->py);
  double disc = 1.0;
  for (int i = 0; i < e->nLastKicked; i++) {
    int id = e->lastKicked[i];
    double rew = (id < e->nPlayers ? currReward[0] : currReward[1]) * disc;
    e->robotRewards[id] += rew;
    e->robotPosRewards[id] += fmax(0.0, rew);
    disc *= 0.5;
  }
  for (int r = 0; r < e->P; r++) {
    int cond1 = (r == e->closestID[0] || r == e->closestID[1]);
    int cond2 = vlen(vsub(robot_pos(e, r), pos)) < 150;
    if (cond1 || cond2) {
      if (!in_list2(e->lastKicked, e->nLastKicked, r))
        e->robotRewards[r] += fmin(0.0, (r < e->nPlayers ? currReward[0] : currReward[1]) * 0.5);
    }
  }
  e->teamRewards[0] += currReward[0] * 0.1;
  e->teamRewards[1] += currReward[1] * 0.1;
  V2 bp = ball_pos(e);
  for (int t = 0; t < 2; t++) {
    int best = 0; double bd = INFINITY;
    for (int k = 0; k < e->nPlayers; k++) {
      double d = vlensq(vsub(bp, robot_pos(e, t * e->nPlayers + k)));
      if (d < bd) { bd = d; best = k; }
    }
    e->closestID[t] = t * e->nPlayers + best;
  }
}

/* ---------------------------------------------------------------- collision callbacks */
static double vangle(V2 v) { return atan2(v.y, v.x); } /* cutils.angle (cutils.py:600-601) */

static void cb_robot_pushing_det(RcEnv* e, int sa, int sb) {
  /* begin, (Robot, Robot): RoboCupEnvironment.py:1010-1035 */
  int r1 = sa / 2, r2 = sb / 2;
  V2 v1 = v2(e->body[sa].vx, e->body[sa].vy);
  V2 v2_ = v2(e->body[sb].vx, e->body[sb].vy);
  V2 dp = vsub(robot_pos(e, r1), robot_pos(e, r2));
  e->rob[r1].mightPush = vlen(v1) > 1 && cos(vangle(dp) - vangle(v1)) < -0.4;
  e->rob[r2].mightPush = vlen(v2_) > 1 && cos(vangle(dp) - vangle(v2_)) > 0.4;
  e->rob[r1].touching = 1; e->rob[r2].touching = 1;
  e->rob[r1].touchCntr = 0; e->rob[r2].touchCntr = 0;
}

static void cb_robot_collision(RcEnv* e, int sa, int sb) {
  /* post_solve, (Robot, Robot): RoboCupEnvironment.py:1038-1081 */
  if (!e->canFall) return;
  int i1 = sa / 2, i2 = sb / 2;
  if (i1 == i2) return;
  Robot* r1 = &e->rob[i1];
  Robot* r2 = &e->rob[i2];
  if (!(r1->fallen || r1->penalized)) r1->touchCntr += 1;
  if (!(r2->fallen || r2->penalized)) r2->touchCntr += 1;
  const double normalThresh = 0.9999, pushingThresh = 0.99995;
  uint32_t tick = (uint32_t)e->elapsed;
  double u = rnd(e, tick, RC_SITE_ROBCOL, sa, sb, 0);
  if (u > pow(r1->mightPush ? pushingThresh : normalThresh, r1->touchCntr) && !r1->fallen) {
    fall(e, i1, r1->mightPush);
    r1->touchCntr = 0;
  }
  u = rnd(e, tick, RC_SITE_ROBCOL, sa, sb, 1);
  if (u > pow(r2->mightPush ? pushingThresh : normalThresh, r2->touchCntr) && !r2->fallen) {
    fall(e, i2, r2->mightPush);
    r2->touchCntr = 0;
  }
  if (r1->mightPush && !r2->mightPush && r2->fallen && r1->team != r2->team) {
    penalize(e, i1); r1->touchCntr = 0;
  } else if (r2->mightPush && !r1->mightPush && r1->fallen && r1->team != r2->team) {
    penalize(e, i2); r2->touchCntr = 0;
  }
}

static void cb_separate(RcEnv* e, int sa, int sb) {
  /* separate, (Robot, Robot) and (Robot, Goalpost): RoboCupEnvironment.py:1084-1094 */
  int s[2] = {sa, sb};
  for (int k = 0; k < 2; k++)
    if (is_foot(s[k])) {
      Robot* rb = &e->rob[s[k] / 2];
      rb->touching = 0; rb->mightPush = 0; rb->touchCntr = 0;
    }
}

static void cb_goalpost_collision(RcEnv* e, int foot, int post) {
  /* post_solve, (Robot, Goalpost): RoboCupEnvironment.py:1097-1123 */
  if (!e->canFall) return;
  int r = foot / 2;
  Robot* rb = &e->rob[r];
  if (rb->fallen) { rb->touchCntr = 0; return; }
  if (!rb->touching) { rb->touching = 1; rb->touchCntr = 0; }
  rb->touchCntr += 1;
  double thresh = pow(0.9998, rb->touchCntr);
  double u = rnd(e, (uint32_t)e->elapsed, RC_SITE_POSTCOL, post, foot, 0);
  if (u > thresh) fall(e, r, 1);
}

static void cb_ball_collision(RcEnv* e, int foot) {
  /* begin, (Robot, Ball): RoboCupEnvironment.py:1126-1146 */
  int r = foot / 2;
  Robot* rb = &e->rob[r];
  if (e->ballOwned != 0) {
    if (rb->team != e->ballOwned && !rb->penalized && e->canFall) penalize(e, r);
    else { e->ballOwned = 0; e->gracePeriod = 0; e->ballFreeCntr = 0; }
  }
  int n = e->nLastKicked < 4 ? e->nLastKicked : 3;
  for (int k = n; k > 0; k--) e->lastKicked[k] = e->lastKicked[k - 1];
  e->lastKicked[0] = r;
  e->nLastKicked = n + 1;
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

static void apply_friction(Body* b, double friction, double rotFriction, double spin, double dt) {
  /* cutils.py:102-140 on top of cpBodyUpdateVelocity (gravity 0, damping 1) */
  b->vx = b->vx + (b->fx * b->m_inv) * dt;
  b->vy = b->vy + (b->fy * b->m_inv) * dt;
  b->w = b->w + b->t * b->i_inv * dt;
  b->fx = 0; b->fy = 0; b->t = 0;
  double factor = friction * b->m, rotFactor = rotFriction * b->m;
  double x = b->vx, y = b->vy;
  double length = 1.0 / (fabs(x) + fabs(y) + 1e-5);
  double theta = b->w;
  double a0 = x * factor * length, a1 = y * factor * length;
  a0 += a1 * spin * theta;
  a1 -= a0 * spin * theta;
  if (fabs(x) < factor) x = 0; else x -= a0;
  if (fabs(y) < factor) y = 0; else y -= a1;
  if (fabs(theta) < rotFactor) theta = 0; else theta -= theta > 0 ? rotFactor : -rotFactor;
  b->vx = x; b->vy = y; b->w = theta;
}

static int shape_active(const RcEnv* e, int s) { return s >= 20 || s < 2 * e->P; }

static Arbiter* arb_find(RcEnv* e, int a, int b) {
  for (int k = 0; k < e->nArb; k++) if (e->arb[k].a == a && e->arb[k].b == b) return &e->arb[k];
  return NULL;
}

static void collide_pair(RcEnv* e, int i, int j) {
  /* cpSpaceCollideShapes for insertion-ordered shapes i < j */
  const ShapeCache* A = &e->sc[i];
  const ShapeCache* B = &e->sc[j];
  if (!(A->l <= B->r && B->l <= A->r && A->b <= B->t && B->b <= A->t)) return;
  /* cpCollide: order by primitive type (circle < segment) */
  int a = i, b = j;
  if (is_foot(a) && !is_foot(b)) { a = j; b = i; }
  Contact con[2];
  V2 n = v2(0, 0);
  int cnt;
  if (!is_foot(a) && !is_foot(b)) cnt = collide_circle_circle(e->sc[a].ta, e->sc[b].ta, 10.0, 10.0, &n, con);
  else if (!is_foot(a)) cnt = collide_circle_segment(&e->sc[a], 10.0, &e->sc[b], ROB_RAD, &n, con);
  else cnt = collide_segment_segment(&e->sc[a], &e->sc[b], &n, con);
  if (cnt == 0) return;
  Arbiter* arb = arb_find(e, a, b);
  if (!arb) {
    if (e->nArb >= RC_MAXARB) { e->arbOverflow = 1; return; }
    arb = &e->arb[e->nArb++];
    memset(arb, 0, sizeof(*arb));
    arb->a = a; arb->b = b; arb->state = RC_ARB_FIRST;
  }
  /* cpArbiterUpdate */
  Body* ba = &e->body[a];
  Body* bb = &e->body[b];
  for (int c = 0; c < cnt; c++) {
    con[c].r1 = vsub(con[c].p1, v2(ba->px, ba->py));
    con[c].r2 = vsub(con[c].p2, v2(bb->px, bb->py));
    con[c].jnAcc = con[c].jtAcc = 0;
    for (int o = 0; o < arb->ncon; o++)
      if (con[c].hash == arb->con[o].hash) { con[c].jnAcc = arb->con[o].jnAcc; con[c].jtAcc = arb->con[o].jtAcc; }
  }
  arb->ncon = cnt;
  for (int c = 0; c < cnt; c++) arb->con[c] = con[c];
  arb->n = n;
  arb->e = shape_e(a) * shape_e(b);
  arb->u = shape_u(a) * shape_u(b);
  if (arb->state == RC_ARB_CACHED) arb->state = RC_ARB_FIRST;
  if (arb->state == RC_ARB_FIRST) { /* begin callbacks (all return True) */
    if (is_foot(a) && is_foot(b)) cb_robot_pushing_det(e, a, b);
    else if (a == BALL && is_foot(b)) cb_ball_collision(e, b);
  }
  e->active[e->nActive++] = (int)(arb - e->arb);
  arb->age = 0;
  arb->touched = 1;
}

static void space_step(RcEnv* e, double dt) {
  /* cpSpaceStep (Chipmunk2D 7.0.x) */
  int prevStepped = e->stepped;
  e->stepped = 1;
  for (int k = 0; k < e->nArb; k++) {
    if (e->arb[k].state == RC_ARB_FIRST) e->arb[k].state = RC_ARB_NORMAL;
    e->arb[k].touched = 0;
  }
  e->nActive = 0;
  /* integrate positions (cpBodyUpdatePosition) */
  for (int s = 0; s <= BALL; s++) {
    if (!shape_active(e, s)) continue;
    Body* b = &e->body[s];
    b->px += (b->vx + b->vbx) * dt;
    b->py += (b->vy + b->vby) * dt;
    b->a += (b->w + b->wb) * dt;
    b->vbx = 0; b->vby = 0; b->wb = 0;
    shape_cache_update(e, s);
  }
  /* collide, pairs in ascending (i, j) insertion order; static-static skipped */
  for (int i = 0; i <= BALL; i++) {
    if (!shape_active(e, i)) continue;
    for (int j = i + 1; j < NSHAPE; j++) {
      if (!shape_active(e, j)) continue;
      collide_pair(e, i, j);
    }
  }
  /* cached arbiters: age, separate callbacks, expiry (cpSpaceArbiterSetFilter) */
  {
    int w = 0;
    for (int k = 0; k < e->nArb; k++) {
      Arbiter* arb = &e->arb[k];
      if (!arb->touched) {
        arb->age += 1;
        if (arb->state != RC_ARB_CACHED) {
          arb->state = RC_ARB_CACHED;
          if (!(arb->a == BALL && is_foot(arb->b)) && !(arb->a == BALL && arb->b >= POST0))
            cb_separate(e, arb->a, arb->b);
        }
        if (arb->age >= 3) continue; /* collisionPersistence: dropped */
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
    Body* a = &e->body[arb->a];
    Body* b = &e->body[arb->b];
    V2 n = arb->n;
    V2 bd = v2(b->px - a->px, b->py - a->py);
    for (int c = 0; c < arb->ncon; c++) {
      Contact* con = &arb->con[c];
      con->nMass = 1.0 / k_scalar(a, b, con->r1, con->r2, n);
      con->tMass = 1.0 / k_scalar(a, b, con->r1, con->r2, v2(-n.y, n.x));
      double dist = vdot(vadd(vsub(con->r2, con->r1), bd), n);
      con->bias = -biasCoef * fmin(0.0, dist + slop) / dt;
      con->jBias = 0;
      V2 vr = v2((b->vx - con->r2.y * b->w) - (a->vx - con->r1.y * a->w),
                 (b->vy + con->r2.x * b->w) - (a->vy + con->r1.x * a->w));
      con->bounce = vdot(vr, n) * arb->e;
    }
  }
  if (getenv("RC_ORACLE_DEBUG"))
    for (int q = 0; q < e->nActive; q++) {
      Arbiter* arb = &e->arb[e->active[q]];
      for (int c = 0; c < arb->ncon; c++) {
        Contact* con = &arb->con[c];
        printf("ORC q%d (%d,%d) st%d c%d n=(%.7f,%.7f) r1=(%.6f,%.6f) r2=(%.6f,%.6f) nMass=%.4f tMass=%.4f bias=%.6f bounce=%.6f jn=%.3f jt=%.3f\n",
               q, arb->a, arb->b, arb->state, c, arb->n.x, arb->n.y, con->r1.x, con->r1.y, con->r2.x, con->r2.y, con->nMass, con->tMass,
               con->bias, con->bounce, con->jnAcc, con->jtAcc);
      }
    }
  for (int r = 0; r < e->P; r++) {
    Robot* rb = &e->rob[r];
    Body* a = &e->body[2 * r];
    Body* b = &e->body[2 * r + 1];
    if (!rb->jointAbsent) { /* cpPivotJoint preStep, anchors at both centres of gravity */
      rb->pivK = 1.0 / (a->m_inv + b->m_inv);
      double c = -(1.0 - pow(0.1, dt)) / dt; /* joint.error_bias = 0.1 (Robot.py:59) */
      rb->pivBias = v2((b->px - a->px) * c, (b->py - a->py) * c);
    }
    { /* cpRotaryLimitJoint preStep, min = max = 0 */
      double dist = b->a - a->a;
      double pdist = 0;
      if (dist > 0) pdist = 0 - dist; else if (dist < 0) pdist = 0 - dist;
      rb->rotISum = 1.0 / (a->i_inv + b->i_inv);
      rb->rotBias = -biasCoef * pdist / dt;
      if (rb->rotBias == 0) rb->rotJ = 0;
    }
  }
  /* integrate velocities through the friction callbacks (cutils.py:93-99) */
  for (int s = 0; s < 2 * e->P; s++) apply_friction(&e->body[s], 1e-3, 1e-2, 0.0, dt);
  apply_friction(&e->body[BALL], 2.8e-2, 1e-3, 5e-2, dt);
  /* warm start */
  double dt_coef = prevStepped ? 1.0 : 0.0;
  for (int q = 0; q < e->nActive; q++) {
    Arbiter* arb = &e->arb[e->active[q]];
    if (arb->state == RC_ARB_FIRST) continue;
    Body* a = &e->body[arb->a];
    Body* b = &e->body[arb->b];
    V2 n = arb->n;
    for (int c = 0; c < arb->ncon; c++) {
      Contact* con = &arb->con[c];
      V2 j = v2(n.x * con->jnAcc - n.y * con->jtAcc, n.x * con->jtAcc + n.y * con->jnAcc);
      apply_impulses(a, b, con->r1, con->r2, vmul(j, dt_coef));
    }
  }
  for (int r = 0; r < e->P; r++) {
    Robot* rb = &e->rob[r];
    Body* a = &e->body[2 * r];
    Body* b = &e->body[2 * r + 1];
    if (!rb->jointAbsent) apply_impulses(a, b, v2(0, 0), v2(0, 0), vmul(rb->pivJ, dt_coef));
    double j = rb->rotJ * dt_coef;
    a->w -= j * a->i_inv; b->w += j * b->i_inv;
  }
  /* solver: 10 iterations of {all arbiters, then all constraints} */
  for (int it = 0; it < 10; it++) {
    if (getenv("RC_ORACLE_DEBUG") && it < 2) {
      for (int b = 4; b < 10; b++) printf("ORC it%d pre-contacts body %d v=(%.6f,%.6f) w=%.6f\n", it, b, e->body[b].vx, e->body[b].vy, e->body[b].w);
      for (int r = 2; r < 5; r++) printf("ORC it%d robot %d pivOn %d pivB=(%.6f,%.6f) rotBias=%.6f pivJ=(%.3f,%.3f) rotJ=%.3f\n", it, r, !e->rob[r].jointAbsent, e->rob[r].pivBias.x, e->rob[r].pivBias.y, e->rob[r].rotBias, e->rob[r].pivJ.x, e->rob[r].pivJ.y, e->rob[r].rotJ);
    }
    for (int q = 0; q < e->nActive; q++) {
      Arbiter* arb = &e->arb[e->active[q]];
      Body* a = &e->body[arb->a];
      Body* b = &e->body[arb->b];
      V2 n = arb->n;
      for (int c = 0; c < arb->ncon; c++) { /* cpArbiterApplyImpulse */
        Contact* con = &arb->con[c];
        V2 r1 = con->r1, r2 = con->r2;
        V2 vb1 = v2(a->vbx - r1.y * a->wb, a->vby + r1.x * a->wb);
        V2 vb2 = v2(b->vbx - r2.y * b->wb, b->vby + r2.x * b->wb);
        V2 vr = v2((b->vx - r2.y * b->w) - (a->vx - r1.y * a->w), (b->vy + r2.x * b->w) - (a->vy + r1.x * a->w));
        double vbn = vdot(vsub(vb2, vb1), n);
        double vrn = vdot(vr, n);
        double vrt = vdot(vr, v2(-n.y, n.x));
        double jbn = (con->bias - vbn) * con->nMass;
        double jbnOld = con->jBias;
        con->jBias = fmax(jbnOld + jbn, 0.0);
        double jn = -(con->bounce + vrn) * con->nMass;
        double jnOld = con->jnAcc;
        con->jnAcc = fmax(jnOld + jn, 0.0);
        double jtMax = arb->u * con->jnAcc;
        double jt = -vrt * con->tMass;
        double jtOld = con->jtAcc;
        con->jtAcc = fmin(fmax(jtOld + jt, -jtMax), jtMax);
        apply_bias_impulses(a, b, r1, r2, vmul(n, con->jBias - jbnOld));
        double dn = con->jnAcc - jnOld, dtt = con->jtAcc - jtOld;
        apply_impulses(a, b, r1, r2, v2(n.x * dn - n.y * dtt, n.x * dtt + n.y * dn));
      }
    }
    for (int r = 0; r < e->P; r++) {
      Robot* rb = &e->rob[r];
      Body* a = &e->body[2 * r];
      Body* b = &e->body[2 * r + 1];
      if (!rb->jointAbsent) { /* cpPivotJoint applyImpulse with r1 = r2 = 0 */
        V2 vr = v2(b->vx - a->vx, b->vy - a->vy);
        V2 j = v2(rb->pivK * (rb->pivBias.x - vr.x), rb->pivK * (rb->pivBias.y - vr.y));
        rb->pivJ = vadd(rb->pivJ, j);
        apply_impulses(a, b, v2(0, 0), v2(0, 0), j);
      }
      if (rb->rotBias != 0) { /* cpRotaryLimitJoint applyImpulse */
        double wr = b->w - a->w;
        double j = -(rb->rotBias + wr) * rb->rotISum;
        double jOld = rb->rotJ;
        if (rb->rotBias < 0) rb->rotJ = fmax(jOld + j, 0.0); else rb->rotJ = fmin(jOld + j, 0.0);
        j = rb->rotJ - jOld;
        a->w -= j * a->i_inv; b->w += j * b->i_inv;
      }
    }
  }
  /* post-solve callbacks, active arbiters in order */
  for (int q = 0; q < e->nActive; q++) {
    Arbiter* arb = &e->arb[e->active[q]];
    if (is_foot(arb->a) && is_foot(arb->b)) cb_robot_collision(e, arb->a, arb->b);
    else if (arb->a >= POST0 && is_foot(arb->b)) cb_goalpost_collision(e, arb->b, arb->a);
  }
  /* deferred space.add(joint) issued from callbacks */
  joint_add_now(e);
}

/* ---------------------------------------------------------------- vision (cutils.py) */
enum { S_NO = 0, S_PARTIAL = 1, S_DISTANT = 2, S_NORMAL = 3, S_MISCLASS = 4 };
enum { I_NO = 0, I_NEARBY = 1, I_OCCLUDE = 2 };

typedef struct { int seen; int has; V2 p; double radius; double ex[4]; } Sight; /* [type, rotPt|None, radius, extras] */
typedef struct { int seen; int has; V2 p1, p2; double tx, ty; } LineSight;

static Sight is_seen_in_area(V2 point, V2 dir1, V2 dir2, double maxDist, double angle, double radius,
                             int allowPartial) {
  /* cutils.py:699-747 */
  Sight s;
  memset(&s, 0, sizeof(s));
  s.seen = S_NO; s.has = 0; s.radius = radius;
  double dist1 = vcross(dir1, point);
  double dist2 = vcross(dir2, point);
  if (dist1 < radius && dist2 > -radius) {
    if (dist1 < -radius && dist2 > radius) {
      s.seen = vlensq(point) < maxDist ? S_NORMAL : S_DISTANT;
    } else if (allowPartial) {
      s.seen = S_PARTIAL;
    } else {
      double b1 = -2 * (dir1.x * point.x - dir1.y * point.y);
      double b2 = -2 * (dir2.x * point.x - dir2.y * point.y);
      double c = vdot(point, point) - radius * radius;
      double sqr1 = b1 * b1 - 4 * c, sqr2 = b2 * b2 - 4 * c;
      int found = 0;
      if (sqr1 >= 0) {
        double sq = sqrt(sqr1);
        if ((-b1 + sq) > 0 || (-b1 - sq) > 0) { s.seen = S_PARTIAL; found = 1; }
      }
      if (!found && sqr2 >= 0) {
        double sq = sqrt(sqr2);
        if ((-b2 + sq) > 0 || (-b2 - sq) > 0) s.seen = S_PARTIAL;
      }
    }
    s.p = vrot(point, -angle);
    s.has = 1;
  }
  return s;
}

static LineSight is_line_in_area(V2 p1, V2 p2, V2 dir1, V2 dir2, double maxDist, double angle) {
  /* cutils.py:751-821 */
  LineSight s;
  memset(&s, 0, sizeof(s));
  s.seen = S_NO;
  double dist11 = vcross(dir1, p1), dist12 = vcross(dir1, p2);
  if (!(dist11 > 0 && dist12 > 0)) {
    double dist21 = vcross(dir2, p1), dist22 = vcross(dir2, p2);
    if (!(dist21 < 0 && dist22 < 0)) {
      s.seen = S_NORMAL;
      V2 pt1, pt2;
      if (dist11 <= 0 && dist21 >= 0) pt1 = p1;
      else {
        double inter1 = vcross(p1, dir1) / (vcross(dir1, vsub(p2, p1)) + 1e-7);
        double inter2 = vcross(p1, dir2) / (vcross(dir2, vsub(p2, p1)) + 1e-7);
        double inter = (inter1 < 1 && inter2 < 1) ? fmax(inter1, inter2) : fmin(inter1, inter2);
        pt1 = vadd(p1, vmul(vsub(p2, p1), inter));
        s.seen = S_PARTIAL;
      }
      if (dist12 <= 0 && dist22 >= 0) pt2 = p2;
      else {
        double inter1 = vcross(p2, dir1) / vcross(dir1, vsub(p1, p2));
        double inter2 = vcross(p2, dir2) / vcross(dir2, vsub(p1, p2));
        double inter = (inter1 < 1 && inter2 < 1) ? fmax(inter1, inter2) : fmin(inter1, inter2);
        pt2 = vadd(p2, vmul(vsub(p1, p2), inter));
        s.seen = S_PARTIAL;
      }
      if (vlensq(pt1) > maxDist || vlensq(pt2) > maxDist) s.seen = S_DISTANT;
      pt1 = vrot(pt1, -angle);
      pt2 = vrot(pt2, -angle);
      if (pt1.x < 0 || pt2.x < 0) s.seen = S_NO;
      s.p1 = pt1; s.p2 = pt2; s.has = 1;
    }
  }
  return s;
}

static int does_interact(const Sight* o1, const Sight* o2, double radius, int canOcclude) {
  /* cutils.py:546-565 (obj = rotPt or None) */
  if (!o1->has || !o2->has) return I_NO;
  int type = I_NO;
  if (vlen(vsub(o1->p, o2->p)) < radius) type = I_NEARBY;
  if (canOcclude) {
    double dist = vcross(o1->p, o2->p) / vlen(o1->p);
    if (fabs(dist) < radius && vlensq(o1->p) < vlensq(o2->p)) type = I_OCCLUDE;
  }
  return type;
}

typedef struct { RcEnv* e; uint32_t tick; int obs; } NoiseCtx;

static void add_noise(NoiseCtx* nc, int site, int idx, Sight* obj, int interaction, double maxDist, int misClass,
                      int angleNoise) {
  /* cutils.py:417-468; obj->ex[2] is obj[5] (the angle) for field crosses */
  RcEnv* e = nc->e;
  double magn = e->noiseMag, rand_ = e->randBase;
  if (interaction == I_OCCLUDE) { obj->seen = S_NO; return; }
  if (obj->seen) {
    double ux = rnd(e, nc->tick, site, nc->obs, idx, 0);
    double uy = rnd(e, nc->tick, site, nc->obs, idx, 1);
    V2 noiseVec = v2((ux - 0.5) * magn, (uy - 0.5) * magn);
    if (e->noiseType == DYNENV_NOISE_RANDOM) {
      if (rnd(e, nc->tick, site, nc->obs, idx, 2) < rand_) obj->seen = S_NO;
      obj->p = vadd(obj->p, noiseVec);
      obj->radius *= (1 - (rnd(e, nc->tick, site, nc->obs, idx, 4) - 0.5) * 0.2);
      if (angleNoise) obj->ex[2] += (rnd(e, nc->tick, site, nc->obs, idx, 5) - 0.5) * magn / 10;
    } else {
      int st = obj->seen;
      double range = 0.25 + 3.75 * vlensq(obj->p) / maxDist;
      double mult = range;
      if (interaction == I_NEARBY) mult = range * 2;
      if (st == S_DISTANT) mult = range * 3;
      else if (st == S_PARTIAL) mult = range * 4;
      V2 newPos = vadd(obj->p, v2(noiseVec.x * mult / 4, noiseVec.y * mult / 4));
      double diff = vlen(newPos) - vlen(obj->p);
      if (rnd(e, nc->tick, site, nc->obs, idx, 2) < rand_ * mult) st = S_NO;
      if (misClass && rnd(e, nc->tick, site, nc->obs, idx, 3) < rand_ * mult / 2) st = S_MISCLASS;
      obj->seen = st;
      obj->p = newPos;
      obj->radius *= 1 + (rnd(e, nc->tick, site, nc->obs, idx, 4) * 0.1 * diff);
      if (angleNoise) obj->ex[2] += (rnd(e, nc->tick, site, nc->obs, idx, 5) - 0.5) * magn * mult / 180;
    }
  }
}

static void add_noise_line(NoiseCtx* nc, int idx, LineSight* obj, double maxDist) {
  /* cutils.py:352-378 */
  RcEnv* e = nc->e;
  double magn = e->noiseMag, rand_ = e->randBase;
  if (obj->seen) {
    double u0 = rnd(e, nc->tick, RC_SITE_OBS_LINE, nc->obs, idx, 0);
    double u1 = rnd(e, nc->tick, RC_SITE_OBS_LINE, nc->obs, idx, 1);
    double u2 = rnd(e, nc->tick, RC_SITE_OBS_LINE, nc->obs, idx, 2);
    double u3 = rnd(e, nc->tick, RC_SITE_OBS_LINE, nc->obs, idx, 3);
    V2 n1 = v2((u0 - 0.5) * magn, (u1 - 0.5) * magn);
    V2 n2 = v2((u2 - 0.5) * magn, (u3 - 0.5) * magn);
    if (e->noiseType == DYNENV_NOISE_RANDOM) {
      if (rnd(e, nc->tick, RC_SITE_OBS_LINE, nc->obs, idx, 4) < rand_) obj->seen = S_NO;
      obj->p1 = vadd(obj->p1, n1);
      obj->p2 = vadd(obj->p2, n2);
    } else {
      double m1 = 0.25 + 3.75 * vlensq(obj->p1) / maxDist;
      double m2 = 0.25 + 3.75 * vlensq(obj->p2) / maxDist;
      double m = (m1 + m2) * 0.5;
      if (getenv("RC_ORACLE_DEBUG") && nc->obs == 8 && e->elapsed % 50 == 0)
        printf("ORC line %d lp1=(%.5f,%.5f) lp2=(%.5f,%.5f) m1=%.6f m2=%.6f n1=(%.6f,%.6f) n2=(%.6f,%.6f) u4=%.6f\n", idx, obj->p1.x, obj->p1.y, obj->p2.x, obj->p2.y, m1, m2, n1.x, n1.y, n2.x, n2.y, rnd(e, nc->tick, RC_SITE_OBS_LINE, nc->obs, idx, 4));
      if (rnd(e, nc->tick, RC_SITE_OBS_LINE, nc->obs, idx, 4) < rand_ * m) obj->seen = S_NO;
      obj->p1 = vadd(obj->p1, v2(n1.x * m1 / 2, n1.y * m1 / 2));
      obj->p2 = vadd(obj->p2, v2(n2.x * m2 / 2, n2.y * m2 / 2));
    }
  }
}

static inline double normalize_(double pt, double nf) { return ((pt * nf) - 0) * 2 * 1; } /* cutils.py:318-323 */
static inline double norm_after_scale(double pt, double nf, double mean, double team) { return (pt - mean) * nf * team; }
static inline double scale_(double v, double nf) { return ((v * nf) - 0.5) / 0.5; } /* cutils.py:295-296 */
#define STDNORM (2.0 / W_)
#define SIZENORM (10.0 / PEN_RADIUS)

static void to_polar(double* out, V2 p, double radius, double sizeMean, double t1, double t2, double team) {
  /* cutils.py:308-315 */
  double dist = sqrt(p.x * p.x + p.y * p.y);
  double ang = atan2(p.y * team, p.x * team);
  out[0] = scale_(dist, STDNORM);
  out[1] = cos(ang);
  out[2] = sin(ang);
  out[3] = (radius - sizeMean) * SIZENORM;
  out[4] = t1 * team;
  out[5] = t2 * team;
}

/* internal lists of sightings after noise; capped lists: inserts beyond LISTCAP are dropped */
typedef struct { V2 p; double radius, a, b, c; int seen; } Det; /* generic detection */

static void list_insert(Det* l, int* n, int cap, int at, Det d) {
  if (at > *n) at = *n;
  if (at >= cap) return;
  int last = *n < cap ? *n : cap - 1;
  for (int k = last; k > at; k--) l[k] = l[k - 1];
  l[at] = d;
  if (*n < cap) (*n)++;
}

/* the detection lists of one getAgentVision call as the IMAGE branch draws them (RoboCupEnvironment.py:1318-1467):
   positions and sizes in the observer's head frame, before the float32 encoders */
typedef struct RcRawVision {
  double ball[RC_CAP_BALL][3], robot[RC_CAP_ROBOT][3], post[RC_CAP_POST][3], cross[RC_CAP_CROSS][3],
      fcross[RC_CAP_FCROSS][3]; /* x, y, size */
  double line[RC_CAP_LINE][4];  /* x1, y1, x2, y2 */
  double circle[4];             /* seen (SightingType != NoSighting), x, y, radius */
  int32_t counts[RC_NTYPES];
  int32_t pad[2];
} RcRawVision;
typedef struct RcRawOut { RcRawVision v[RC_NTICK][RC_MAXP]; } RcRawOut;

typedef struct {
  int counts[RC_NTYPES];
  int numLandMarks;
  int robotsSeen[RC_MAXP - 1];
  int ballsSeen;
  double ball[RC_CAP_BALL][RC_F_BALL];
  double robot[RC_CAP_ROBOT][RC_F_ROBOT];
  double post[RC_CAP_POST][RC_F_POST];
  double cross[RC_CAP_CROSS][RC_F_CROSS];
  double fcross[RC_CAP_FCROSS][RC_F_FCROSS];
  double line[RC_CAP_LINE][RC_F_LINE];
} AgentObs;

static void agent_vision(RcEnv* e, int ag, uint32_t tick, AgentObs* out, RcRawVision* raw) {
  /* RoboCupEnvironment.py:1192-1316,1536-1561 */
  memset(out, 0, sizeof(*out));
  Robot* agent = &e->rob[ag];
  V2 pos = robot_pos(e, ag);
  double angle = robot_angle(e, ag);
  double headAngle = angle + agent->headAngle;
  double angle1 = headAngle + ROB_FOV, angle2 = headAngle - ROB_FOV;
  V2 vec1 = vrot(v2(1, 0), angle1), vec2 = vrot(v2(1, 0), angle2);
  /* _setup_vision(0.4, 0.8): environment_base.py:161-172 */
  double maxVis0 = (W_ * 0.4) * (W_ * 0.4), maxVis1 = (W_ * 0.8) * (W_ * 0.8);
  int P = e->P;
  NoiseCtx nc = {e, tick, ag};

  Sight ballDet = is_seen_in_area(vsub(ball_pos(e), pos), vec1, vec2, maxVis0, headAngle, BALL_RADIUS * 2, 1);
  ballDet.ex[0] = e->ballOwned * agent->team;
  Sight robDets[RC_MAXP]; int robIdx[RC_MAXP]; int nRob = 0;
  for (int r = 0; r < P; r++) {
    if (r == ag) continue;
    Sight s = is_seen_in_area(vsub(robot_pos(e, r), pos), vec1, vec2, maxVis1, headAngle, ROB_TOTAL, 1);
    s.ex[0] = robot_angle(e, r) - headAngle;
    s.ex[1] = agent->team * e->rob[r].team;
    s.ex[2] = (agent->fallen || agent->penalized) ? 1 : 0; /* the OBSERVER's flag (quirk) */
    robIdx[nRob] = r;
    robDets[nRob++] = s;
  }
  Sight goalDets[4];
  for (int k = 0; k < 4; k++) {
    goalDets[k] = is_seen_in_area(vsub(g_posts[k].p, pos), vec1, vec2, maxVis1, headAngle, POST_RADIUS, 1);
    goalDets[k].ex[0] = g_posts[k].side; goalDets[k].ex[1] = g_posts[k].dir;
  }
  Sight crossDets[3];
  for (int k = 0; k < 3; k++) {
    crossDets[k] = is_seen_in_area(vsub(g_pcross[k].p, pos), vec1, vec2, maxVis0, headAngle, PEN_RADIUS, 1);
    crossDets[k].ex[0] = g_pcross[k].t1; crossDets[k].ex[1] = g_pcross[k].t2;
  }
  Sight fcDets[16];
  for (int k = 0; k < 16; k++) {
    fcDets[k] = is_seen_in_area(vsub(g_lcross[k].p, pos), vec1, vec2, maxVis0, headAngle, PEN_RADIUS, 1);
    fcDets[k].ex[0] = g_lcross[k].t1; fcDets[k].ex[1] = g_lcross[k].t2; fcDets[k].ex[2] = g_lcross[k].t3 - headAngle;
  }
  LineSight lineDets[11];
  for (int k = 0; k < 11; k++) {
    lineDets[k] = is_line_in_area(vsub(g_lines[k].p1, pos), vsub(g_lines[k].p2, pos), vec1, vec2, maxVis1, headAngle);
    lineDets[k].tx = g_lines[k].tx; lineDets[k].ty = g_lines[k].ty;
  }
  Sight circleDet = is_seen_in_area(vsub(v2(520, 370), pos), vec1, vec2, maxVis1, headAngle, CC_RADIUS * 2, 0);

  /* interactions (:1227-1237) */
  int robRobInter[RC_MAXP];
  if (e->nPlayers > 1) {
    for (int j = 0; j < nRob; j++) {
      int m = 0;
      /* `rob1 != rob2` compares the sighting LISTS by value (quirk A6): identical sightings
         (e.g. two unseen robots, [NoSighting, None, 17.5, ...] with equal extras) are skipped too */
      for (int i = 0; i < nRob; i++) {
        const Sight* a = &robDets[i];
        const Sight* b = &robDets[j];
        int equal = (i == j) ||
                    (a->seen == b->seen && a->has == b->has && (!a->has || (a->p.x == b->p.x && a->p.y == b->p.y)) &&
                     a->radius == b->radius && a->ex[0] == b->ex[0] && a->ex[1] == b->ex[1] && a->ex[2] == b->ex[2]);
        if (equal) continue;
        int t = does_interact(a, b, ROB_TOTAL * 2, 1);
        if (t > m) m = t;
      }
      robRobInter[j] = m;
    }
  } else robRobInter[0] = 0;
  int robBallInter = 0;
  for (int i = 0; i < nRob; i++) { int t = does_interact(&robDets[i], &ballDet, ROB_TOTAL * 2, 1); if (t > robBallInter) robBallInter = t; }
  int robPostInter[4], robCrossInter[3], robFCInter[16];
  for (int k = 0; k < 4; k++) { int m = 0; for (int i = 0; i < nRob; i++) { int t = does_interact(&robDets[i], &goalDets[k], ROB_TOTAL * 2, 1); if (t > m) m = t; } robPostInter[k] = m; }
  for (int k = 0; k < 3; k++) { int m = 0; for (int i = 0; i < nRob; i++) { int t = does_interact(&robDets[i], &crossDets[k], ROB_TOTAL * 2, 1); if (t > m) m = t; } robCrossInter[k] = m; }
  for (int k = 0; k < 16; k++) { int m = 0; for (int i = 0; i < nRob; i++) { int t = does_interact(&robDets[i], &fcDets[k], ROB_TOTAL * 2, 1); if (t > m) m = t; } robFCInter[k] = m; }
  int ballPostInter = 0;
  for (int k = 0; k < 4; k++) { int t = does_interact(&ballDet, &goalDets[k], BALL_RADIUS * 8, 0); if (t > ballPostInter) ballPostInter = t; }
  int ballCrossInter[3];
  for (int k = 0; k < 3; k++) ballCrossInter[k] = does_interact(&ballDet, &crossDets[k], BALL_RADIUS * 4, 0);

  /* noise, false negatives, misclassification (:1240-1252) */
  add_noise(&nc, RC_SITE_OBS_BALL, 0, &ballDet, robBallInter > ballPostInter ? robBallInter : ballPostInter, maxVis0, 1, 0);
  for (int i = 0; i < nRob; i++) add_noise(&nc, RC_SITE_OBS_ROBOT, i, &robDets[i], robRobInter[i], maxVis1, 0, 0);
  for (int k = 0; k < 4; k++) add_noise(&nc, RC_SITE_OBS_POST, k, &goalDets[k], robPostInter[k], maxVis1, 0, 0);
  for (int k = 0; k < 3; k++)
    add_noise(&nc, RC_SITE_OBS_CROSS, k, &crossDets[k], robCrossInter[k] > ballCrossInter[k] ? robCrossInter[k] : ballCrossInter[k], maxVis0, 1, 0);
  for (int k = 0; k < 16; k++) add_noise(&nc, RC_SITE_OBS_FCROSS, k, &fcDets[k], robFCInter[k], maxVis0, 0, 1);
  add_noise(&nc, RC_SITE_OBS_CIRCLE, 0, &circleDet, 0, maxVis1, 0, 0);
  for (int k = 0; k < 11; k++) add_noise_line(&nc, k, &lineDets[k], maxVis1);

  for (int i = 0; i < nRob; i++) out->robotsSeen[i] = robDets[i].seen != S_NO;
  out->ballsSeen = (ballDet.seen != S_NO && ballDet.seen != S_MISCLASS);

  /* build the lists; misclassified ball <-> cross swaps (:1258-1274) */
  Det balls[LISTCAP], robs[LISTCAP], posts[LISTCAP], crosses[LISTCAP], fcs[LISTCAP];
  int nb = 0, nr = 0, np = 0, ncr = 0, nf = 0;
  Det crossAll[8]; int nCrossAll = 0;
  for (int k = 0; k < 3; k++) {
    Det d = {crossDets[k].p, crossDets[k].radius, crossDets[k].ex[0], crossDets[k].ex[1], 0, crossDets[k].seen};
    crossAll[nCrossAll++] = d;
  }
  Det ballAll[8]; int nBallAll = 0;
  { Det d = {ballDet.p, ballDet.radius, ballDet.ex[0], 0, 0, ballDet.seen}; ballAll[nBallAll++] = d; }
  if (ballDet.seen == S_MISCLASS) {
    Det d = {ballDet.p, ballDet.radius, 0, 0, 0, S_NORMAL};
    d.a = rnd_int(e, tick, RC_SITE_OBS_MISCROSS, ag, 0, 0, -1, 1);
    d.b = rnd_int(e, tick, RC_SITE_OBS_MISCROSS, ag, 0, 1, -1, 1);
    crossAll[nCrossAll++] = d;
  }
  { int nc0 = nCrossAll;
    for (int k = 0; k < nc0; k++)
      if (crossAll[k].seen == S_MISCLASS) { Det d = {crossAll[k].p, crossAll[k].radius, 0, 0, 0, S_NORMAL}; ballAll[nBallAll++] = d; } }
  for (int k = 0; k < nBallAll; k++) if (ballAll[k].seen != S_NO && ballAll[k].seen != S_MISCLASS) list_insert(balls, &nb, RC_CAP_BALL, nb, ballAll[k]);
  for (int i = 0; i < nRob; i++) if (robDets[i].seen != S_NO) { Det d = {robDets[i].p, robDets[i].radius, robDets[i].ex[0], robDets[i].ex[1], robDets[i].ex[2], robDets[i].seen}; list_insert(robs, &nr, RC_CAP_ROBOT, nr, d); }
  for (int k = 0; k < 4; k++) if (goalDets[k].seen != S_NO) { Det d = {goalDets[k].p, goalDets[k].radius, goalDets[k].ex[0], goalDets[k].ex[1], 0, goalDets[k].seen}; list_insert(posts, &np, RC_CAP_POST, np, d); }
  for (int k = 0; k < nCrossAll; k++) if (crossAll[k].seen != S_NO && crossAll[k].seen != S_MISCLASS) list_insert(crosses, &ncr, RC_CAP_CROSS, ncr, crossAll[k]);
  for (int k = 0; k < 16; k++) if (fcDets[k].seen != S_NO && fcDets[k].seen != S_MISCLASS) { Det d = {fcDets[k].p, fcDets[k].radius, fcDets[k].ex[0], fcDets[k].ex[1], fcDets[k].ex[2], fcDets[k].seen}; list_insert(fcs, &nf, RC_CAP_FCROSS, nf, d); }
  int nl = 0; int lineKeep[11];
  for (int k = 0; k < 11; k++) if (lineDets[k].seen != S_NO) lineKeep[nl++] = k;
  out->numLandMarks = nf + nl + ncr + np;

  /* random false positives (:1279-1305) */
  for (int i = 0; i < 10; i++) {
    if (rnd(e, tick, RC_SITE_OBS_FP, ag, i, 0) < e->randBase) {
      int c = rnd_int(e, tick, RC_SITE_OBS_FP, ag, i, 1, 0, 5);
      double d = rnd(e, tick, RC_SITE_OBS_FP, ag, i, 2) * sqrt(maxVis1);
      double a = rnd(e, tick, RC_SITE_OBS_FP, ag, i, 3) * 2 * ROB_FOV - ROB_FOV;
      V2 fp = vrot(v2(d, 0), a);
      if (c == 0) {
        Det dt = {fp, BALL_RADIUS * 2 * (1 - 0.4 * (rnd(e, tick, RC_SITE_OBS_FP, ag, i, 4) - 0.5)), 0, 0, 0, S_NORMAL};
        list_insert(balls, &nb, RC_CAP_BALL, nb, dt);
      } else if (c == 1) {
        Det dt = {fp, 0, 0, 0, 0, S_NORMAL};
        dt.radius = ROB_TOTAL * (1 - 0.4 * (rnd(e, tick, RC_SITE_OBS_FP, ag, i, 4) - 0.5));
        dt.a = (rnd(e, tick, RC_SITE_OBS_FP, ag, i, 5) - 0.5) * 2 * PI;
        dt.b = (rnd(e, tick, RC_SITE_OBS_FP, ag, i, 6) > 0.5) ? -1 : 1;
        dt.c = (rnd(e, tick, RC_SITE_OBS_FP, ag, i, 7) > 0.9) ? 1 : 0;
        list_insert(robs, &nr, RC_CAP_ROBOT, nr, dt);
      } else if (c == 2) {
        Det dt = {fp, 0, 0, 0, 0, S_NORMAL};
        dt.radius = POST_RADIUS * (1 - 0.4 * (rnd(e, tick, RC_SITE_OBS_FP, ag, i, 4) - 0.5));
        dt.a = rnd_int(e, tick, RC_SITE_OBS_FP, ag, i, 5, -1, 1);
        dt.b = rnd_int(e, tick, RC_SITE_OBS_FP, ag, i, 6, -1, 1);
        list_insert(posts, &np, RC_CAP_POST, np, dt);
      } else if (c == 3) {
        Det dt = {fp, 0, 0, 0, 0, S_NORMAL};
        dt.radius = PEN_RADIUS * (1 - 0.4 * (rnd(e, tick, RC_SITE_OBS_FP, ag, i, 4) - 0.5));
        dt.a = rnd_int(e, tick, RC_SITE_OBS_FP, ag, i, 5, -1, 1);
        dt.b = rnd_int(e, tick, RC_SITE_OBS_FP, ag, i, 6, -1, 1);
        list_insert(crosses, &ncr, RC_CAP_CROSS, ncr, dt);
      } else if (c == 4) {
        Det dt = {fp, 0, 0, 0, 0, S_NORMAL};
        dt.radius = PEN_RADIUS * (1 - 0.4 * (rnd(e, tick, RC_SITE_OBS_FP, ag, i, 4) - 0.5));
        dt.a = rnd_int(e, tick, RC_SITE_OBS_FP, ag, i, 5, -1, 1);
        dt.b = rnd_int(e, tick, RC_SITE_OBS_FP, ag, i, 6, -1, 1);
        dt.c = rnd(e, tick, RC_SITE_OBS_FP, ag, i, 7) * PI * 2;
        list_insert(fcs, &nf, RC_CAP_FCROSS, ncr, dt); /* inserted at len(crossDets) (quirk A8) */
      }
    }
  }
  /* false-positive balls near robots (:1308-1316) */
  if (e->noiseType == DYNENV_NOISE_REALISTIC) {
    for (int i = 0; i < nr; i++) {
      if (robs[i].seen != S_NORMAL) continue;
      if (!(rnd(e, tick, RC_SITE_OBS_FPBALL, ag, i, 0) < e->randBase * 10)) continue;
      if (!(vlen(robs[i].p) < 250)) continue;
      if (rnd(e, tick, RC_SITE_OBS_FPBALL, ag, i, 1) < e->randBase * 8) robs[i].seen = S_NO;
      double ox = 2 * rnd(e, tick, RC_SITE_OBS_FPBALL, ag, i, 2) - 1.0;
      double oy = 2 * rnd(e, tick, RC_SITE_OBS_FPBALL, ag, i, 3) - 1.0;
      Det dt = {v2(robs[i].p.x + ox * ROB_TOTAL, robs[i].p.y + oy * ROB_TOTAL), 0, 0, 0, 0, S_NORMAL};
      dt.radius = BALL_RADIUS * 2 * (1 - 0.4 * (rnd(e, tick, RC_SITE_OBS_FPBALL, ag, i, 4) - 0.5));
      list_insert(balls, &nb, RC_CAP_BALL, nb, dt);
    }
  }
  /* encode (:1537-1561) */
  int isClosest = (ag == e->closestID[0] || ag == e->closestID[1]);
  for (int k = 0; k < nb; k++) {
    out->ball[k][0] = normalize_(balls[k].p.x, STDNORM);
    out->ball[k][1] = normalize_(balls[k].p.y, STDNORM);
    out->ball[k][2] = norm_after_scale(balls[k].radius, SIZENORM, BALL_RADIUS * 2, 1);
    out->ball[k][3] = balls[k].a;
    out->ball[k][4] = isClosest;
  }
  for (int k = 0; k < nr; k++) {
    out->robot[k][0] = normalize_(robs[k].p.x, STDNORM);
    out->robot[k][1] = normalize_(robs[k].p.y, STDNORM);
    out->robot[k][2] = norm_after_scale(robs[k].radius, SIZENORM, ROB_TOTAL, 1);
    out->robot[k][3] = cos(robs[k].a);
    out->robot[k][4] = sin(robs[k].a);
    out->robot[k][5] = robs[k].b;
    out->robot[k][6] = robs[k].c;
  }
  for (int k = 0; k < np; k++) to_polar(out->post[k], posts[k].p, posts[k].radius, POST_RADIUS, posts[k].a, posts[k].b, agent->team);
  for (int k = 0; k < ncr; k++) to_polar(out->cross[k], crosses[k].p, crosses[k].radius, PEN_RADIUS, crosses[k].a, crosses[k].b, agent->team);
  for (int k = 0; k < nf; k++) {
    to_polar(out->fcross[k], fcs[k].p, fcs[k].radius, PEN_RADIUS, fcs[k].a, fcs[k].b, agent->team);
    out->fcross[k][6] = cos(fcs[k].c);
    out->fcross[k][7] = -sin(fcs[k].c);
  }
  for (int q = 0; q < nl; q++) { /* normalizeLine: cutils.py:333-347 */
    LineSight* ln = &lineDets[lineKeep[q]];
    V2 diff = vsub(ln->p2, ln->p1);
    double dist = fabs(ln->p2.x * ln->p1.y - ln->p2.y * ln->p1.x) / (vlen(diff) + 1e-7);
    double ang = atan2(diff.y, diff.x);
    out->line[q][0] = scale_(dist, STDNORM);
    out->line[q][1] = cos(ang);
    out->line[q][2] = sin(ang);
    out->line[q][3] = ln->tx;
    out->line[q][4] = ln->ty;
  }
  out->counts[0] = nb; out->counts[1] = nr; out->counts[2] = np; out->counts[3] = ncr; out->counts[4] = nf; out->counts[5] = nl;
  if (raw) {
    memset(raw, 0, sizeof(*raw));
    for (int k = 0; k < nb; k++) { raw->ball[k][0] = balls[k].p.x; raw->ball[k][1] = balls[k].p.y; raw->ball[k][2] = balls[k].radius; }
    for (int k = 0; k < nr; k++) { raw->robot[k][0] = robs[k].p.x; raw->robot[k][1] = robs[k].p.y; raw->robot[k][2] = robs[k].radius; }
    for (int k = 0; k < np; k++) { raw->post[k][0] = posts[k].p.x; raw->post[k][1] = posts[k].p.y; raw->post[k][2] = posts[k].radius; }
    for (int k = 0; k < ncr; k++) { raw->cross[k][0] = crosses[k].p.x; raw->cross[k][1] = crosses[k].p.y; raw->cross[k][2] = crosses[k].radius; }
    for (int k = 0; k < nf; k++) { raw->fcross[k][0] = fcs[k].p.x; raw->fcross[k][1] = fcs[k].p.y; raw->fcross[k][2] = fcs[k].radius; }
    for (int q = 0; q < nl; q++) {
      const LineSight* ln = &lineDets[lineKeep[q]];
      raw->line[q][0] = ln->p1.x; raw->line[q][1] = ln->p1.y; raw->line[q][2] = ln->p2.x; raw->line[q][3] = ln->p2.y;
    }
    raw->circle[0] = circleDet.seen != S_NO;
    if (circleDet.seen != S_NO) { raw->circle[1] = circleDet.p.x; raw->circle[2] = circleDet.p.y; raw->circle[3] = circleDet.radius; }
    for (int k = 0; k < RC_NTYPES; k++) raw->counts[k] = out->counts[k];
  }
}

/* getFullState(agent): RoboCupEnvironment.py:1162-1188 */
static void full_state_agent(const RcEnv* e, int ag, double* ball4, double* self8, double* others /*[MAXP-1][6]*/) {
  const Robot* agent = &e->rob[ag];
  double team = agent->team;
  V2 bp = ball_pos(e);
  ball4[0] = norm_after_scale(bp.x, STDNORM, W_ / 2, team);
  ball4[1] = norm_after_scale(bp.y, STDNORM, H_ / 2, team);
  ball4[2] = e->ballOwned * agent->team;
  ball4[3] = (ag == e->closestID[0] || ag == e->closestID[1]);
  V2 p = robot_pos(e, ag);
  double a = robot_angle(e, ag);
  self8[0] = norm_after_scale(p.x, STDNORM, W_ / 2, team);
  self8[1] = norm_after_scale(p.y, STDNORM, H_ / 2, team);
  self8[2] = cos(a + agent->headAngle);
  self8[3] = sin(a + agent->headAngle);
  self8[4] = cos(agent->headAngle);
  self8[5] = sin(agent->headAngle);
  self8[6] = agent->team;
  self8[7] = (agent->fallen || agent->penalized) ? 1 : 0;
  int k = 0;
  for (int r = 0; r < e->P; r++) {
    if (r == ag) continue;
    V2 q = robot_pos(e, r);
    double ra = robot_angle(e, r);
    double* o = others + 6 * k++;
    o[0] = norm_after_scale(q.x, STDNORM, W_ / 2, team);
    o[1] = norm_after_scale(q.y, STDNORM, H_ / 2, team);
    o[2] = cos(ra);
    o[3] = sin(ra);
    o[4] = e->rob[r].team * agent->team;
    o[5] = (e->rob[r].fallen || e->rob[r].penalized) ? 1 : 0;
  }
}

/* ---------------------------------------------------------------- public API (ctypes) */
typedef struct RcStepOut {
  double ball[RC_NTICK][RC_MAXP][RC_CAP_BALL][RC_F_BALL];
  double robot[RC_NTICK][RC_MAXP][RC_CAP_ROBOT][RC_F_ROBOT];
  double post[RC_NTICK][RC_MAXP][RC_CAP_POST][RC_F_POST];
  double cross[RC_NTICK][RC_MAXP][RC_CAP_CROSS][RC_F_CROSS];
  double fcross[RC_NTICK][RC_MAXP][RC_CAP_FCROSS][RC_F_FCROSS];
  double line[RC_NTICK][RC_MAXP][RC_CAP_LINE][RC_F_LINE];
  int32_t counts[RC_NTICK][RC_MAXP][RC_NTYPES];
  int32_t num_landmarks[RC_NTICK][RC_MAXP];
  int32_t robots_seen[RC_NTICK][RC_MAXP][RC_MAXP - 1];
  int32_t balls_seen[RC_NTICK][RC_MAXP];
  double full_ball[RC_NTICK][RC_MAXP][4];
  double full_self[RC_NTICK][RC_MAXP][8];
  double full_others[RC_NTICK][RC_MAXP][RC_MAXP - 1][6];
  double recon_ball[RC_MAXP][4];
  double recon_self[RC_MAXP][8];
  double recon_others[RC_MAXP][RC_MAXP - 1][6];
  double state_robots[RC_MAXP][6];
  double state_ball[3];
  double rewards[RC_MAXP];
  int32_t done;
  double episode_r[RC_MAXP], episode_p_r[RC_MAXP], episode_o_r[RC_MAXP];
  int32_t episode_g[2];
} RcStepOut;

int rc_stepout_size(void) { return (int)sizeof(RcStepOut); }
int rc_rawout_size(void) { return (int)sizeof(RcRawOut); }
void rc_set_raw_out(RcEnv* e, RcRawOut* out) { e->rawOut = out; }

RcEnv* rc_create(int nPlayers, int obsType, int noiseType, double noiseMag, int allowHeadTurn, unsigned flags,
                 uint64_t seed, uint32_t env_id) {
  scene_init();
  RcEnv* e = (RcEnv*)calloc(1, sizeof(RcEnv));
  e->nPlayers = nPlayers > 5 ? 5 : nPlayers; /* environment_base.py:56-57 */
  e->P = 2 * e->nPlayers;
  e->obsType = obsType; e->noiseType = noiseType; e->noiseMag = noiseMag; e->allowHeadTurn = allowHeadTurn;
  e->randBase = 0.01 * noiseMag; /* environment_base.py:170 */
  e->randomInit = !!(flags & DYNENV_FLAG_RANDOM_INIT);
  e->deterministicTurn = !!(flags & DYNENV_FLAG_DETERMINISTIC_TURN);
  e->canFall = !!(flags & DYNENV_FLAG_CAN_FALL);
  e->useObsRewards = !!(flags & DYNENV_FLAG_USE_OBS_REWARDS);
  e->rng.mode = 0; e->rng.seed = seed; e->rng.env_id = env_id;
  return e;
}
void rc_destroy(RcEnv* e) { free(e); }
void rc_set_tape(RcEnv* e, const double* tape, int n) { e->rng.mode = tape ? 1 : 0; e->rng.tape = tape; e->rng.tape_n = n; e->rng.tape_pos = 0; e->rng.tape_underflow = 0; }
int rc_tape_pos(RcEnv* e) { return e->rng.tape_pos; }
int rc_tape_underflow(RcEnv* e) { return e->rng.tape_underflow; }
int rc_arb_overflow(RcEnv* e) { return e->arbOverflow; }
void rc_debug_substep_range(RcEnv* e, int begin, int end) { e->subBegin = begin; e->subEnd = end; }

static void init_body_consts(RcEnv* e) {
  for (int s = 0; s < NSHAPE; s++) {
    Body* b = &e->body[s];
    if (s < 20) { /* Robot.py:34-36: moment_for_segment(4000, (-10,+-10), (10,+-10), 7.5) */
      b->m = ROB_MASS;
      double length = 20.0 + 2 * ROB_RAD;
      double inertia = ROB_MASS * ((length * length + 4 * ROB_RAD * ROB_RAD) / 12.0 + 100.0);
      b->m_inv = 1.0 / ROB_MASS; b->i_inv = 1.0 / inertia;
    } else if (s == BALL) { /* Ball.py:8-10: moment_for_circle(10, 0, 10) */
      b->m = 10; b->m_inv = 0.1; b->i_inv = 1.0 / (10 * 0.5 * 100.0);
    } else {
      b->m = INFINITY; b->m_inv = 0; b->i_inv = 0;
    }
  }
  for (int k = 0; k < 4; k++) {
    Body* b = &e->body[POST0 + k];
    b->px = g_posts[k].p.x; b->py = g_posts[k].p.y;
    shape_cache_update(e, POST0 + k);
  }
  for (int r = 0; r < RC_MAXP; r++) { e->rob[r].id = r; e->rob[r].team = r < e->nPlayers ? 1 : -1; }
}

void rc_reset(RcEnv* e) {
  /* RoboCupEnvironment.__init__ + _setup_scene (RoboCupEnvironment.py:42-63,239-336) */
  int episode = e->episode + 1;
  OracleRng g = e->rng;
  int nP = e->nPlayers, oT = e->obsType, nT = e->noiseType, aH = e->allowHeadTurn;
  double nM = e->noiseMag;
  int f0 = e->randomInit, f1 = e->deterministicTurn, f2 = e->canFall, f3 = e->useObsRewards;
  struct RcRawOut* rawOut = e->rawOut;
  memset(e, 0, sizeof(*e));
  e->rawOut = rawOut;
  e->rng = g; e->nPlayers = nP; e->P = 2 * nP; e->obsType = oT; e->noiseType = nT; e->allowHeadTurn = aH;
  e->noiseMag = nM; e->randBase = 0.01 * nM;
  e->randomInit = f0; e->deterministicTurn = f1; e->canFall = f2; e->useObsRewards = f3;
  e->episode = episode;
  init_body_consts(e);
  e->ballOwned = 1; e->ballFreeCntr = 9999; e->gracePeriod = 0;
  e->penalTimes[0] = e->penalTimes[1] = 20000;
  int k = 0;
#define U() rnd(e, DYNENV_TICK_RESET, RC_SITE_RESET, 0, 0, k++)
  V2 spots[2][5];
  if (e->randomInit) {
    double xL[7] = {SIDE + 10, SIDE + 50, SIDE + 250, SIDE + 450, SIDE + 650, SIDE + 850, SIDE + 890};
    double yL[3] = {SIDE + 20, SIDE + 300, SIDE + 580};
    V2 rs[10]; int n = 0;
    for (int i = 0; i < 6; i++) {
      int edge = (i == 0 || i == 5);
      int ny = edge ? 1 : 2;
      for (int j = 0; j < ny; j++) {
        double yB = edge ? yL[0] : yL[j], yE = edge ? yL[2] : yL[j + 1];
        double ux = U(); double uy = U();
        rs[n++] = v2(xL[i] + ux * (xL[i + 1] - xL[i]), yB + uy * (yE - yB));
      }
    }
    spots[0][0] = rs[0]; spots[1][0] = rs[9];
    int perm[8]; for (int i = 0; i < 8; i++) perm[i] = i;
    for (int i = 7; i > 0; i--) { int j = (int)(U() * (i + 1)); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
    for (int i = 0; i < 8; i++) spots[i < 4 ? 0 : 1][1 + (i & 3)] = rs[perm[i] + 1];
  } else {
    double cX = W_ / 2;
    double u0, u1;
    u0 = U(); u1 = U(); spots[0][0] = v2(cX - (BALL_RADIUS * 2 + ROB_TOTAL) - u0 * 50, H_ / 2 + (u1 - 0.5) * 25);
    u0 = U(); u1 = U(); spots[0][1] = v2(cX - (ROB_TOTAL + LINEW * 2) - u0 * 50, SIDE + FIELD_H / 4 + (u1 - 0.5) * 50);
    u0 = U(); u1 = U(); spots[0][2] = v2(cX - (ROB_TOTAL + LINEW * 2) - u0 * 50, SIDE + 3 * FIELD_H / 4 + (u1 - 0.5) * 50);
    u0 = U(); u1 = U(); spots[0][3] = v2(cX - (FIELD_W / 4) - (u0 - 0.5) * 50, SIDE + FIELD_H / 2 + (u1 - 0.5) * 50);
    u0 = U(); spots[0][4] = v2(SIDE + 20, H_ / 2 + (u0 - 0.5) * 50);
    u0 = U(); u1 = U(); spots[1][0] = v2(cX + (CC_RADIUS * 2 + ROB_TOTAL + LINEW / 2) + u0 * 50, H_ / 2 + (u1 - 0.5) * 50);
    u0 = U(); u1 = U(); spots[1][1] = v2(cX + (ROB_TOTAL + LINEW / 2 + CC_RADIUS) + u0 * 50, SIDE + FIELD_H / 4 + (u1 - 0.5) * 50);
    u0 = U(); u1 = U(); spots[1][2] = v2(cX + (ROB_TOTAL + LINEW / 2 + CC_RADIUS) + u0 * 50, SIDE + 3 * FIELD_H / 4 + (u1 - 0.5) * 50);
    u0 = U(); u1 = U(); spots[1][3] = v2(cX + (SIDE + FIELD_W / 4) + u0 * 50, SIDE + FIELD_H / 2 + (u1 - 0.5) * 50);
    u0 = U(); spots[1][4] = v2(W_ - (SIDE + 20), H_ / 2 + (u0 - 0.5) * 50);
  }
  int ids[2][5];
  for (int t = 0; t < 2; t++) { /* np.random.permutation(maxPlayers) */
    for (int i = 0; i < 5; i++) ids[t][i] = i;
    for (int i = 4; i > 0; i--) { int j = (int)(U() * (i + 1)); int tmp = ids[t][i]; ids[t][i] = ids[t][j]; ids[t][j] = tmp; }
  }
  for (int r = 0; r < e->P; r++) {
    int t = r < e->nPlayers ? 0 : 1;
    int i = r - t * e->nPlayers;
    V2 p = spots[t][ids[t][i]];
    double ang = t == 0 ? 0 : PI;
    for (int f = 0; f < 2; f++) {
      Body* b = &e->body[2 * r + f];
      b->px = p.x; b->py = p.y; b->a = ang;
      shape_cache_update(e, 2 * r + f);
    }
    e->rob[r].prevPos = p;
    if (e->deterministicTurn) e->rob[r].headAngle = e->rob[r].team * HEAD_MAX;
  }
  double bx = 520, by = 370;
  if (e->randomInit) {
    bx = U() * FIELD_W + SIDE;
    by = U() * FIELD_H + SIDE;
    e->ballOwned = U() > 0.4 ? 1 : 0;
    if (e->ballOwned != 0 && U() > 0.5) e->ballOwned *= -1;
  }
#undef U
  e->body[BALL].px = bx; e->body[BALL].py = by;
  shape_cache_update(e, BALL);
  e->ballPrev = v2(bx, by);
}

/* state record <-> RcEnv; words are doubles holding the float / int value of each 32-bit word */
void rc_store_state(const RcEnv* e, double* w) {
  for (int i = 0; i < RC_STATE_WORDS; i++) w[i] = 0;
  for (int s = 0; s < RC_NBODY; s++) {
    const Body* b = &e->body[s];
    w[RC_S_PX + s] = b->px; w[RC_S_PY + s] = b->py; w[RC_S_ANG + s] = b->a;
    w[RC_S_VX + s] = b->vx; w[RC_S_VY + s] = b->vy; w[RC_S_W + s] = b->w;
    w[RC_S_FX + s] = b->fx; w[RC_S_FY + s] = b->fy; w[RC_S_TQ + s] = b->t;
    w[RC_S_VBX + s] = b->vbx; w[RC_S_VBY + s] = b->vby; w[RC_S_WB + s] = b->wb;
    w[RC_S_CAX + s] = e->sc[s].ta.x; w[RC_S_CAY + s] = e->sc[s].ta.y;
    w[RC_S_CBX + s] = e->sc[s].tb.x; w[RC_S_CBY + s] = e->sc[s].tb.y;
  }
  for (int r = 0; r < RC_MAXP; r++) {
    const Robot* rb = &e->rob[r];
    w[RC_S_HEAD + r] = rb->headAngle; w[RC_S_HEADMOV + r] = rb->headMoving; w[RC_S_MOVET + r] = rb->moveTime;
    w[RC_S_INITX + r] = rb->initPos.x; w[RC_S_INITY + r] = rb->initPos.y;
    w[RC_S_FALLT + r] = rb->fallTime; w[RC_S_PENT + r] = rb->penalTime;
    w[RC_S_PREVX + r] = rb->prevPos.x; w[RC_S_PREVY + r] = rb->prevPos.y;
    w[RC_S_PIVJX + r] = rb->pivJ.x; w[RC_S_PIVJY + r] = rb->pivJ.y; w[RC_S_ROTJ + r] = rb->rotJ;
    int fl = (rb->kicking ? RC_RF_KICKING : 0) | (rb->foot ? RC_RF_FOOT : 0) | (rb->fallen ? RC_RF_FALLEN : 0) |
             (rb->penalized ? RC_RF_PENALIZED : 0) | (rb->touching ? RC_RF_TOUCHING : 0) |
             (rb->mightPush ? RC_RF_MIGHTPUSH : 0) | (rb->jointRemoved ? RC_RF_JOINT_REMOVED : 0) |
             (rb->jointAbsent ? RC_RF_JOINT_ABSENT : 0) | (rb->jointAddPending ? RC_RF_JOINT_ADD_PENDING : 0);
    w[RC_S_RFLAGS + r] = fl; w[RC_S_FALLCNT + r] = rb->fallCntr; w[RC_S_TOUCHCNT + r] = rb->touchCntr;
    w[RC_S_EPR + r] = e->epR[r]; w[RC_S_EPPR + r] = e->epPR[r]; w[RC_S_EPOR + r] = e->epOR[r];
  }
  w[RC_S_BPREVX] = e->ballPrev.x; w[RC_S_BPREVY] = e->ballPrev.y;
  for (int k = 0; k < 4; k++) w[RC_S_LASTK + k] = e->lastKicked[k];
  w[RC_S_NLASTK] = e->nLastKicked;
  w[RC_S_ELAPSED] = e->elapsed; w[RC_S_EPISODE] = e->episode; w[RC_S_BALLOWNED] = e->ballOwned;
  w[RC_S_BALLFREE] = e->ballFreeCntr; w[RC_S_GRACE] = e->gracePeriod;
  for (int t = 0; t < 2; t++) {
    w[RC_S_GOALS + t] = e->goals[t]; w[RC_S_CLOSEST + t] = e->closestID[t];
    w[RC_S_DEFENDERS + 2 * t] = e->nDef[t] > 0 ? e->defenders[t][0] : 0;
    w[RC_S_DEFENDERS + 2 * t + 1] = e->nDef[t] > 1 ? e->defenders[t][1] : 0;
    w[RC_S_NDEF + t] = e->nDef[t]; w[RC_S_PENALTIMES + t] = e->penalTimes[t];
  }
  w[RC_S_STEPPED] = e->stepped;
  w[RC_S_NARB] = e->nArb;
  for (int k = 0; k < e->nArb; k++) {
    const Arbiter* a = &e->arb[k];
    w[RC_S_ARBKEY + k] = a->a | (a->b << 8) | (a->state << 16) | (a->age << 24);
    int h0 = a->ncon > 0 ? a->con[0].hash : 0, h1 = a->ncon > 1 ? a->con[1].hash : 0;
    w[RC_S_ARBCON + k] = a->ncon | (h0 << 8) | (h1 << 16);
    w[RC_S_ARBJN0 + k] = a->ncon > 0 ? a->con[0].jnAcc : 0; w[RC_S_ARBJT0 + k] = a->ncon > 0 ? a->con[0].jtAcc : 0;
    w[RC_S_ARBJN1 + k] = a->ncon > 1 ? a->con[1].jnAcc : 0; w[RC_S_ARBJT1 + k] = a->ncon > 1 ? a->con[1].jtAcc : 0;
  }
}

void rc_load_state(RcEnv* e, const double* w) {
  init_body_consts(e);
  for (int s = 0; s < RC_NBODY; s++) {
    Body* b = &e->body[s];
    /* two-float coordinates of the float32 record: hi + lo (the oracle itself stores lo = 0) */
    b->px = w[RC_S_PX + s] + w[RC_S_PXL + s]; b->py = w[RC_S_PY + s] + w[RC_S_PYL + s];
    b->a = w[RC_S_ANG + s] + w[RC_S_ANGL + s];
    b->vx = w[RC_S_VX + s]; b->vy = w[RC_S_VY + s]; b->w = w[RC_S_W + s];
    b->fx = w[RC_S_FX + s]; b->fy = w[RC_S_FY + s]; b->t = w[RC_S_TQ + s];
    b->vbx = w[RC_S_VBX + s]; b->vby = w[RC_S_VBY + s]; b->wb = w[RC_S_WB + s];
    e->sc[s].ta = v2(w[RC_S_CAX + s], w[RC_S_CAY + s]);
    e->sc[s].tb = v2(w[RC_S_CBX + s], w[RC_S_CBY + s]);
  }
  for (int r = 0; r < RC_MAXP; r++) {
    Robot* rb = &e->rob[r];
    rb->headAngle = w[RC_S_HEAD + r]; rb->headMoving = w[RC_S_HEADMOV + r]; rb->moveTime = w[RC_S_MOVET + r];
    rb->initPos = v2(w[RC_S_INITX + r] + w[RC_S_INITXL + r], w[RC_S_INITY + r] + w[RC_S_INITYL + r]);
    rb->fallTime = w[RC_S_FALLT + r]; rb->penalTime = w[RC_S_PENT + r];
    rb->prevPos = v2(w[RC_S_PREVX + r], w[RC_S_PREVY + r]);
    rb->pivJ = v2(w[RC_S_PIVJX + r], w[RC_S_PIVJY + r]); rb->rotJ = w[RC_S_ROTJ + r];
    int fl = (int)w[RC_S_RFLAGS + r];
    rb->kicking = !!(fl & RC_RF_KICKING); rb->foot = !!(fl & RC_RF_FOOT); rb->fallen = !!(fl & RC_RF_FALLEN);
    rb->penalized = !!(fl & RC_RF_PENALIZED); rb->touching = !!(fl & RC_RF_TOUCHING);
    rb->mightPush = !!(fl & RC_RF_MIGHTPUSH); rb->jointRemoved = !!(fl & RC_RF_JOINT_REMOVED);
    rb->jointAbsent = !!(fl & RC_RF_JOINT_ABSENT); rb->jointAddPending = !!(fl & RC_RF_JOINT_ADD_PENDING);
    rb->fallCntr = (int)w[RC_S_FALLCNT + r]; rb->touchCntr = (int)w[RC_S_TOUCHCNT + r];
    e->epR[r] = w[RC_S_EPR + r]; e->epPR[r] = w[RC_S_EPPR + r]; e->epOR[r] = w[RC_S_EPOR + r];
  }
  e->ballPrev = v2(w[RC_S_BPREVX], w[RC_S_BPREVY]);
  for (int k = 0; k < 4; k++) e->lastKicked[k] = (int)w[RC_S_LASTK + k];
  e->nLastKicked = (int)w[RC_S_NLASTK];
  e->elapsed = (int)w[RC_S_ELAPSED]; e->episode = (int)w[RC_S_EPISODE]; e->ballOwned = (int)w[RC_S_BALLOWNED];
  e->ballFreeCntr = w[RC_S_BALLFREE]; e->gracePeriod = w[RC_S_GRACE];
  for (int t = 0; t < 2; t++) {
    e->goals[t] = (int)w[RC_S_GOALS + t]; e->closestID[t] = (int)w[RC_S_CLOSEST + t];
    e->defenders[t][0] = (int)w[RC_S_DEFENDERS + 2 * t]; e->defenders[t][1] = (int)w[RC_S_DEFENDERS + 2 * t + 1];
    e->nDef[t] = (int)w[RC_S_NDEF + t]; e->penalTimes[t] = (int)w[RC_S_PENALTIMES + t];
  }
  e->stepped = (int)w[RC_S_STEPPED];
  e->nArb = (int)w[RC_S_NARB];
  for (int k = 0; k < e->nArb; k++) {
    Arbiter* a = &e->arb[k];
    memset(a, 0, sizeof(*a));
    int key = (int)w[RC_S_ARBKEY + k], cn = (int)w[RC_S_ARBCON + k];
    a->a = key & 255; a->b = (key >> 8) & 255; a->state = (key >> 16) & 255; a->age = (key >> 24) & 255;
    a->ncon = cn & 255; a->con[0].hash = (cn >> 8) & 255; a->con[1].hash = (cn >> 16) & 255;
    a->con[0].jnAcc = w[RC_S_ARBJN0 + k]; a->con[0].jtAcc = w[RC_S_ARBJT0 + k];
    a->con[1].jnAcc = w[RC_S_ARBJN1 + k]; a->con[1].jtAcc = w[RC_S_ARBJT1 + k];
  }
}

static void fill_obs_tick(RcEnv* e, int t, uint32_t tick, RcStepOut* out) {
  for (int r = 0; r < e->P; r++) {
    if (e->obsType == DYNENV_OBS_FULL) {
      full_state_agent(e, r, out->full_ball[t][r], out->full_self[t][r], &out->full_others[t][r][0][0]);
    } else {
      AgentObs ao;
      agent_vision(e, r, tick, &ao, e->rawOut ? &e->rawOut->v[t][r] : NULL);
      memcpy(out->ball[t][r], ao.ball, sizeof(ao.ball));
      memcpy(out->robot[t][r], ao.robot, sizeof(ao.robot));
      memcpy(out->post[t][r], ao.post, sizeof(ao.post));
      memcpy(out->cross[t][r], ao.cross, sizeof(ao.cross));
      memcpy(out->fcross[t][r], ao.fcross, sizeof(ao.fcross));
      memcpy(out->line[t][r], ao.line, sizeof(ao.line));
      for (int k = 0; k < RC_NTYPES; k++) out->counts[t][r][k] = ao.counts[k];
      out->num_landmarks[t][r] = ao.numLandMarks;
      for (int k = 0; k < RC_MAXP - 1; k++) out->robots_seen[t][r][k] = ao.robotsSeen[k];
      out->balls_seen[t][r] = ao.ballsSeen;
    }
  }
}

static void fill_info(RcEnv* e, RcStepOut* out) {
  /* getFullState(): RoboCupEnvironment.py:1151-1161 */
  for (int r = 0; r < e->P; r++) {
    V2 p = robot_pos(e, r);
    double a = robot_angle(e, r);
    out->state_robots[r][0] = normalize_(p.x, STDNORM);
    out->state_robots[r][1] = normalize_(p.y, STDNORM);
    out->state_robots[r][2] = cos(a);
    out->state_robots[r][3] = sin(a);
    out->state_robots[r][4] = e->rob[r].team;
    out->state_robots[r][5] = (e->rob[r].fallen || e->rob[r].penalized) ? 1 : 0;
    full_state_agent(e, r, out->recon_ball[r], out->recon_self[r], &out->recon_others[r][0][0]);
  }
  V2 bp = ball_pos(e);
  out->state_ball[0] = normalize_(bp.x, STDNORM);
  out->state_ball[1] = normalize_(bp.y, STDNORM);
  out->state_ball[2] = e->ballOwned;
}

/* environment_base.reset: the nTimeSteps initial observations (environment_base.py:216-224) */
void rc_reset_obs(RcEnv* e, RcStepOut* out) {
  memset(out, 0, sizeof(*out));
  for (int t = 0; t < RC_NTICK; t++) fill_obs_tick(e, t, DYNENV_TICK_RESETOBS | (uint32_t)t, out);
  fill_info(e, out);
}

void rc_step(RcEnv* e, const double* actions /*[P][4]*/, RcStepOut* out) {
  /* RoboCupEnvironment.step: RoboCupEnvironment.py:446-524 */
  memset(out, 0, sizeof(*out));
  e->teamRewards[0] = e->teamRewards[1] = 0;
  for (int r = 0; r < RC_MAXP; r++) { e->obsRewards[r] = 0; e->robotRewards[r] = 0; e->robotPosRewards[r] = 0; }
  for (int i = e->subBegin; i < (e->subEnd ? e->subEnd : RC_SUBSTEPS); i++) {
    for (int r = 0; r < e->P; r++) {
      if (i == 0) { process_action(e, r, actions + 4 * r); joint_add_now(e); }
      tick_robot(e, r);
      joint_add_now(e); /* outside the locked step, space.add is immediate */
    }
    ball_out_of_field(e);
    space_step(e, 1.0 / 100.0);
    e->elapsed += 1;
    if (i % 10 == 9) fill_obs_tick(e, i / 10, (uint32_t)e->elapsed, out);
  }
  /* processSeens: RoboCupEnvironment.py:1563-1575 */
  if (e->obsType != DYNENV_OBS_FULL) {
    for (int r = 0; r < e->P; r++) {
      double lsum = 0, bsum = 0, rs = 0;
      for (int t = 0; t < RC_NTICK; t++) { lsum += out->num_landmarks[t][r]; bsum += out->balls_seen[t][r]; }
      double lSeens = fmax(0.0, fmin(3.0, lsum / RC_NTICK));
      for (int k = 0; k < e->P - 1; k++) {
        double c = 0;
        for (int t = 0; t < RC_NTICK; t++) c += out->robots_seen[t][r][k];
        rs += fmax(0.0, fmin(2.0, c));
      }
      double bSeens = fmax(0.0, fmin(3.0, bsum));
      if (e->useObsRewards) e->obsRewards[r] += (0.0025 * (rs + lSeens) + 0.01 * bSeens);
    }
  }
  for (int r = 0; r < e->P; r++) {
    int t = r < e->nPlayers ? 0 : 1;
    e->robotRewards[r] += e->teamRewards[t];
    e->robotRewards[r] += e->obsRewards[r];
    e->epR[r] += e->robotRewards[r];
    e->robotPosRewards[r] += fmax(0.0, e->teamRewards[t]);
    e->robotPosRewards[r] += fmax(0.0, e->obsRewards[r]);
    e->epPR[r] += e->robotPosRewards[r];
    e->epOR[r] += e->obsRewards[r];
    out->rewards[r] = e->robotRewards[r];
  }
  fill_info(e, out);
  if (e->elapsed >= MAXTIME) {
    out->done = 1;
    for (int r = 0; r < e->P; r++) { out->episode_r[r] = e->epR[r]; out->episode_p_r[r] = e->epPR[r]; out->episode_o_r[r] = e->epOR[r]; }
    out->episode_g[0] = e->goals[0]; out->episode_g[1] = e->goals[1];
  }
}
