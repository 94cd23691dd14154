/*
 * Composition of the IMAGE observation of one robot and one camera from a raw detection record (include/dynenv_b200.h,
 * RC_R_*): which primitives to draw where.  Restates, for a record,
 *   RoboCupEnvironment.py:1318-1467   the drawing calls of getAgentVision's image branch
 *   cutils.py:166-200                 pin-hole camera matrices (top / bottom camera of the robot's head)
 *   cutils.py:206-222 projectPoints, :226-257 getConicPoints, :261-284 estimateConic (the conic through eight
 *                                     projected circle points; evaluated in closed form, see conic_setup)
 * in float64 like the reference (the projected coordinates are truncated to integer pixels; float32 would move
 * one coordinate in ~1e3 across a pixel boundary).  The pixel work is raster_prims.cuh.
 *
 * Host + device: the CUDA kernel (rc_raster.cuh) spreads the work items of one image over a thread block, the CPU
 * test (tests/test_raster_prims.py) runs them serially.  Where the reference itself would raise (a projected
 * coordinate that does not fit an int, a line thickness outside 1..32767) the primitive is skipped.
 */
#pragma once
#include "../../include/dynenv_b200.h"
#include "raster_prims.cuh"

namespace raster {

/* topTr / bottomTr of cutils.py:177-200: A [R|t]^-1 with focal 543.6, principal point (319.5, 239.5), pitch
   0.0209 + 0.25 at height 58.364 (top) and 0.6929 + 0.25 at height 53.774 (bottom); rows x, y, w over [X, Y, Z, 1] */
RP_FN void project(int cam, double X, double Y, double Z, double* u, double* v) {
  const double top[12] = {543.60000000000002, -85.497791791539967, 307.84797806510034, 3182.6176409012332,
                          0, -587.86504541553177, 85.298876772176541, 33809.365805102643,
                          0, -0.26759872235223775, 0.96353044777809183, 9.9612445724608243};
  const double bot[12] = {543.60000000000002, -258.56019545653214, 187.68823971011045, 12952.04888690959,
                          0, -513.15334559704377, -299.22375223661794, 29111.671653727328,
                          0, -0.80926508750088311, 0.58744362976560394, 40.538494168731113};
  const double* m = cam ? bot : top;
  const double w = m[8] * X + m[9] * Y + m[10] * Z + m[11];
  *u = (m[0] * X + m[1] * Y + m[2] * Z + m[3]) / w;
  *v = (m[4] * X + m[5] * Y + m[6] * Z + m[7]) / w;
}
RP_FN double rp_sqrt(double x) {
#ifdef __CUDA_ARCH__
  return sqrt(x);
#else
  return __builtin_sqrt(x);
#endif
}
RP_FN double rp_fabs(double x) { return x < 0 ? -x : x; }
RP_FN bool to_int(double v, int* out) { /* Python int(): truncation; false where cv2 could not take the value */
  if (!(rp_fabs(v) < 1.0e9)) return false;
  *out = (int)v;
  return true;
}
RP_FN int ceil_int(double v) { /* math.ceil of a non-negative finite value, saturated */
  if (!(v < 1.0e9)) return 1000000000;
  int k = (int)v;
  return (double)k < v ? k + 1 : k;
}

enum { PRIM_LINE = 0, PRIM_RECT = 1, PRIM_CIRCLE = 2 };
struct Prim {
  int type, plane, x0, y0, x1, y1, t; /* LINE: end points, thickness t; RECT: corners; CIRCLE: centre (x0, y0), radius t */
};
constexpr int MAX_PRIMS = 72; /* 11 lines + 12 robots + 6 posts + 24 crosses + 8 balls = 61 slots + 4 conic closures */
constexpr int CONIC_ROWS = IMG_H;

/* the conic of the centre circle in one camera: x of both branches for the image rows with a real solution */
struct Conic {
  int n;         /* rows with a solution (compacted, ascending y) */
  int thickness;
  int x1[CONIC_ROWS], x2[CONIC_ROWS];
  short y[CONIC_ROWS];
};

/* a primitive that can cover a large part of the image: the raster kernel queues these in front of the small ones */
RP_FN bool prim_is_big(const Prim& p) {
  if (p.type == PRIM_LINE) return p.t >= 24;
  if (p.type == PRIM_CIRCLE) return p.t >= 40;
  const long long w = (long long)p.x1 - p.x0, h = (long long)p.y1 - p.y0;
  return (w < 0 ? -w : w) * (h < 0 ? -h : h) >= 64 * 64;
}
/* `plane`: the bit-plane the primitive belongs to (p.plane) */
RP_FN void draw_prim(uint32_t* plane, const Prim& p, Lanes L) {
  if (p.type == PRIM_LINE) cv_line(plane, p.x0, p.y0, p.x1, p.y1, p.t, L);
  else if (p.type == PRIM_RECT) cv_rectangle_filled(plane, p.x0, p.y0, p.x1, p.y1, L);
  else circle_filled(plane, p.x0, p.y0, p.t, L);
}
/* segment s (0-based) of the two poly-lines of the conic: branch s / (n - 1), points i - 1 -> i with i = s % (n - 1) + 1 */
RP_FN void draw_conic_segment(uint32_t* plane3, const Conic& c, int s, Lanes L) {
  const int b = s / (c.n - 1), i = s % (c.n - 1) + 1;
  const int* x = b ? c.x2 : c.x1;
  cv_polyline_segment(plane3, x[i - 1], c.y[i - 1], x[i], c.y[i], c.thickness, i, L);
}

RP_FN double round_half_even(double v) {
#ifdef __CUDA_ARCH__
  return rint(v);
#else
  return __builtin_rint(v);
#endif
}
RP_FN int to_int32_like_numpy(double v) { /* float64 -> int32 cast of an out-of-range / NaN value: INT_MIN (cvttsd2si) */
  return rp_fabs(v) < 2147483648.0 ? (int)v : (int)0x80000000;
}
/* getConicPoints, row y (cutils.py:226-257) */
RP_FN bool conic_row(const double* params, double cx, double cy, int y, int* x1, int* x2) {
  const double yc = (double)y - cy;
  const double a = params[0], a4 = 4 * a, overa = 1.0 / (2 * a);
  const double b = yc * params[2] + params[3];
  const double c = yc * (yc * params[1] + params[4]) + params[5];
  const double sqr = b * b - a4 * c;
  if (!(sqr >= 0)) return false;
  const double sq = rp_sqrt(sqr);
  *x1 = to_int32_like_numpy(round_half_even((-b + sq) * overa + cx));
  *x2 = to_int32_like_numpy(round_half_even((-b - sq) * overa + cx));
  return true;
}

/* the 3 x 3 homography of the ground plane (Y = 0) into the image of camera `cam`: columns X, Z, 1 of the projection */
RP_FN void ground_homography(int cam, double H[3][3]) {
  /* H = [m0 m2 m3; m4 m6 m7; m8 m10 m11] of the matrices in project() */
  const double top[12] = {543.60000000000002, -85.497791791539967, 307.84797806510034, 3182.6176409012332,
                          0, -587.86504541553177, 85.298876772176541, 33809.365805102643,
                          0, -0.26759872235223775, 0.96353044777809183, 9.9612445724608243};
  const double bot[12] = {543.60000000000002, -258.56019545653214, 187.68823971011045, 12952.04888690959,
                          0, -513.15334559704377, -299.22375223661794, 29111.671653727328,
                          0, -0.80926508750088311, 0.58744362976560394, 40.538494168731113};
  const double* m = cam ? bot : top;
  for (int r = 0; r < 3; r++) { H[r][0] = m[4 * r]; H[r][1] = m[4 * r + 2]; H[r][2] = m[4 * r + 3]; }
}

/* the centre circle (RoboCupEnvironment.py:1343-1393): conic parameters of camera `cam`; returns false when nothing
   can be drawn.  cx, cy: projected centre.  The reference fits the conic through the projections of eight points
   of the circle (estimateConic: null vector of the 8 x 6 design matrix by SVD, cutils.py:261-284).  Those points
   lie exactly on the image of the circle under the ground-plane homography H, so the fitted conic IS
   H^-T C H^-1 (C: the circle's matrix) up to scale and rounding; it is evaluated in closed form here (the scale and
   sign of the parameter vector do not matter: they swap the two branches, which are drawn alike). */
RP_FN bool conic_setup(const float* raw, int cam, double* params, double* cx, double* cy, int* thickness) {
  if (reinterpret_cast<const int*>(raw)[RC_R_CIRCLE_SEEN] == 0) return false;
  const double X = raw[RC_R_CIRCLE], Y = raw[RC_R_CIRCLE + 1], R = raw[RC_R_CIRCLE + 2];
  project(cam, -Y, 0.0, X, cx, cy);
  double H[3][3];
  ground_homography(cam, H);
  /* G = adj(H) (= det H * H^-1): image point (u, v, 1) -> ground point up to scale */
  double G[3][3];
  G[0][0] = H[1][1] * H[2][2] - H[1][2] * H[2][1]; G[0][1] = H[0][2] * H[2][1] - H[0][1] * H[2][2]; G[0][2] = H[0][1] * H[1][2] - H[0][2] * H[1][1];
  G[1][0] = H[1][2] * H[2][0] - H[1][0] * H[2][2]; G[1][1] = H[0][0] * H[2][2] - H[0][2] * H[2][0]; G[1][2] = H[0][2] * H[1][0] - H[0][0] * H[1][2];
  G[2][0] = H[1][0] * H[2][1] - H[1][1] * H[2][0]; G[2][1] = H[0][1] * H[2][0] - H[0][0] * H[2][1]; G[2][2] = H[0][0] * H[1][1] - H[0][1] * H[1][0];
  /* centred image coordinates: (u, v, 1) = T (x, y, 1); ground point g = G T (x, y, 1) */
  double GT[3][3];
  for (int r = 0; r < 3; r++) { GT[r][0] = G[r][0]; GT[r][1] = G[r][1]; GT[r][2] = G[r][0] * *cx + G[r][1] * *cy + G[r][2]; }
  /* circle (a, b, R) in ground coordinates g = (gx, gz, gw): (gx - a gw)^2 + (gz - b gw)^2 - R^2 gw^2 = 0 */
  const double a = -Y, b = X;
  double P[3], Q[3]; /* gx - a gw and gz - b gw as linear forms in (x, y, 1) */
  for (int c = 0; c < 3; c++) { P[c] = GT[0][c] - a * GT[2][c]; Q[c] = GT[1][c] - b * GT[2][c]; }
  const double* Wf = GT[2];
  const double R2 = R * R;
  /* A x^2 + B y^2 + C xy + D x + E y + F */
  params[0] = P[0] * P[0] + Q[0] * Q[0] - R2 * Wf[0] * Wf[0];
  params[1] = P[1] * P[1] + Q[1] * Q[1] - R2 * Wf[1] * Wf[1];
  params[2] = 2 * (P[0] * P[1] + Q[0] * Q[1] - R2 * Wf[0] * Wf[1]);
  params[3] = 2 * (P[0] * P[2] + Q[0] * Q[2] - R2 * Wf[0] * Wf[2]);
  params[4] = 2 * (P[1] * P[2] + Q[1] * Q[2] - R2 * Wf[1] * Wf[2]);
  params[5] = P[2] * P[2] + Q[2] * Q[2] - R2 * Wf[2] * Wf[2];
  const double len = rp_sqrt(X * X + Y * Y);
  if (!(len < 1.0e9)) return false;
  int q = (int)(len / (cam ? 20.0 : 40.0));
  const int cap = cam ? 29 : 14;
  q = q < 0 ? 0 : (q > cap ? cap : q);
  *thickness = (cam ? 30 : 15) - q;
  return true;
}

/* Primitive slots of one camera: one per possible row of every list, so that a thread can build slot k on its own:
   lines 0..10, robots 11..22, posts 23..28, crosses 29..34, line crosses 35..52, balls 53..60; the conic's closing
   lines are appended behind them by conic_closures.  Returns false for an empty slot. */
constexpr int SLOT_ROBOT = RC_CAP_LINE, SLOT_POST = SLOT_ROBOT + RC_CAP_ROBOT, SLOT_CROSS = SLOT_POST + RC_CAP_POST,
              SLOT_FCROSS = SLOT_CROSS + RC_CAP_CROSS, SLOT_BALL = SLOT_FCROSS + RC_CAP_FCROSS, N_SLOTS = SLOT_BALL + RC_CAP_BALL;
static_assert(N_SLOTS + 4 <= MAX_PRIMS, "primitive slots");
RP_FN int slot_plane(int slot) { /* channel 0 balls, 1 robots, 2 goal posts, 3 lines / crosses (and the centre circle) */
  return slot < SLOT_ROBOT ? 3 : (slot < SLOT_POST ? 1 : (slot < SLOT_CROSS ? 2 : (slot < SLOT_BALL ? 3 : 0)));
}
RP_FN bool build_prim_slot(const float* raw, int cam, int slot, Prim* out) {
  const int* cnt = reinterpret_cast<const int*>(raw) + RC_R_COUNTS;
  double u[4], v[4];
  Prim p;
  p.x0 = p.y0 = p.x1 = p.y1 = p.t = 0;
  if (slot < SLOT_ROBOT) { /* lines (:1324-1341) */
    const int q = slot;
    if (q >= cnt[5]) return false;
    const double x1 = raw[RC_R_LINE + 4 * q], y1 = raw[RC_R_LINE + 4 * q + 1], x2 = raw[RC_R_LINE + 4 * q + 2], y2 = raw[RC_R_LINE + 4 * q + 3];
    project(cam, (-y1 - y2) / 2, 0.0, (x1 + x2) / 2, &u[0], &v[0]);
    project(cam, (-y1 - y2) / 2 + 5.0 / 2, 0.0, (x1 + x2) / 2, &u[1], &v[1]);
    project(cam, -y1, 0.0, x1, &u[2], &v[2]);
    project(cam, -y2, 0.0, x2, &u[3], &v[3]);
    const int rad = ceil_int(rp_sqrt((u[0] - u[1]) * (u[0] - u[1]) + (v[0] - v[1]) * (v[0] - v[1])));
    p.type = PRIM_LINE; p.plane = 3; p.t = rad;
    if (!(rad >= 1 && rad <= 32767 && to_int(u[2], &p.x0) && to_int(v[2], &p.y0) && to_int(u[3], &p.x1) && to_int(v[3], &p.y1))) return false;
  } else if (slot < SLOT_POST) { /* robots: filled rectangles (:1395-1408) */
    const int k = slot - SLOT_ROBOT;
    if (k >= cnt[1]) return false;
    const double x = raw[RC_R_ROBOT + 3 * k], y = raw[RC_R_ROBOT + 3 * k + 1], r = raw[RC_R_ROBOT + 3 * k + 2];
    project(cam, -y - r, 0.0, x, &u[0], &v[0]);
    project(cam, -y + r, 58.0, x, &u[1], &v[1]);
    p.type = PRIM_RECT; p.plane = 1;
    if (!(to_int(u[0], &p.x0) && to_int(v[0], &p.y0) && to_int(u[1], &p.x1) && to_int(v[1], &p.y1))) return false;
  } else if (slot < SLOT_CROSS) { /* goal posts: thick lines from the ground to 80 (:1410-1425) */
    const int k = slot - SLOT_POST;
    if (k >= cnt[2]) return false;
    const double x = raw[RC_R_POST + 3 * k], y = raw[RC_R_POST + 3 * k + 1], r = raw[RC_R_POST + 3 * k + 2];
    project(cam, -y, 0.0, x, &u[0], &v[0]);
    project(cam, -y + r / 2, 0.0, x, &u[1], &v[1]);
    project(cam, -y, 80.0, x, &u[2], &v[2]);
    const int rad = ceil_int(rp_sqrt((u[0] - u[1]) * (u[0] - u[1]) + (v[0] - v[1]) * (v[0] - v[1])));
    p.type = PRIM_LINE; p.plane = 2; p.t = rad;
    if (!(rad >= 1 && rad <= 32767 && to_int(u[0], &p.x0) && to_int(v[0], &p.y0) && to_int(u[2], &p.x1) && to_int(v[2], &p.y1))) return false;
  } else if (slot < N_SLOTS) { /* crosses, line crosses (:1427-1451) and balls (:1453-1466): filled circles */
    const int list = slot < SLOT_FCROSS ? 0 : (slot < SLOT_BALL ? 1 : 2);
    const int k = slot - (list == 0 ? SLOT_CROSS : (list == 1 ? SLOT_FCROSS : SLOT_BALL));
    const int off = list == 0 ? RC_R_CROSS : (list == 1 ? RC_R_FCROSS : RC_R_BALL);
    if (k >= (list == 0 ? cnt[3] : (list == 1 ? cnt[4] : cnt[0]))) return false;
    const double x = raw[off + 3 * k], y = raw[off + 3 * k + 1], r = raw[off + 3 * k + 2];
    const double h = list == 2 ? r / 2 : 0.0;
    project(cam, -y, h, x, &u[0], &v[0]);
    project(cam, -y + r / 2, h, x, &u[1], &v[1]);
    const int rad = ceil_int(rp_sqrt((u[0] - u[1]) * (u[0] - u[1]) + (v[0] - v[1]) * (v[0] - v[1])));
    p.type = PRIM_CIRCLE; p.plane = list == 2 ? 0 : 3; p.t = rad;
    if (!(rad <= 1000000 && to_int(u[0], &p.x0) && to_int(v[0], &p.y0))) return false;
  } else return false;
  *out = p;
  return true;
}

/* the lines that close the conic at its first / last row unless they lie on the image border (:1381-1393) */
RP_FN int conic_closures(const Conic& c, Prim* prims, int n) {
  if (c.n <= 0) return n;
  Prim p;
  p.type = PRIM_LINE; p.plane = 3; p.t = c.thickness;
  if (c.y[0] != 0 && n < MAX_PRIMS) { p.x0 = c.x1[0]; p.y0 = c.y[0]; p.x1 = c.x2[0]; p.y1 = c.y[0]; prims[n++] = p; }
  if (c.y[c.n - 1] < IMG_H - 1 && n < MAX_PRIMS) { p.x0 = c.x1[c.n - 1]; p.y0 = c.y[c.n - 1]; p.x1 = c.x2[c.n - 1]; p.y1 = c.y[c.n - 1]; prims[n++] = p; }
  return n;
}

#ifndef __CUDA_ARCH__
/* serial path (host): the four planes of one camera */
inline void render_camera_host(const float* raw, int cam, uint32_t* planes /* [4][PLANE_WORDS], zeroed */) {
  Prim prims[MAX_PRIMS];
  for (int k = 0; k < N_SLOTS; k++)
    if (!build_prim_slot(raw, cam, k, &prims[k])) prims[k].type = -1;
  int n = N_SLOTS;
  static thread_local Conic conic;
  conic.n = 0;
  double params[6], cx, cy;
  if (conic_setup(raw, cam, params, &cx, &cy, &conic.thickness)) {
    for (int y = 0; y < CONIC_ROWS; y++) {
      int x1, x2;
      if (conic_row(params, cx, cy, y, &x1, &x2)) { conic.x1[conic.n] = x1; conic.x2[conic.n] = x2; conic.y[conic.n] = (short)y; conic.n++; }
    }
    n = conic_closures(conic, prims, n);
  }
  const Lanes one = {0, 1};
  for (int k = 0; k < n; k++)
    if (prims[k].type >= 0) draw_prim(planes + prims[k].plane * PLANE_WORDS, prims[k], one);
  if (conic.n > 1)
    for (int s = 0; s < 2 * (conic.n - 1); s++) draw_conic_segment(planes + 3 * PLANE_WORDS, conic, s, one);
}
#endif

}  // namespace raster
