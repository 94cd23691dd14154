/*
 * Rasterisation primitives of the IMAGE observation (RoboCupEnvironment.py:1318-1467): the pixel sets of
 * cv2.line / cv2.polylines / cv2.rectangle(filled) / cv2.circle(filled) with lineType LINE_8, shift 0, on a
 * 480 x 640 single-bit plane (20 uint32 words per row, bit x % 32 of word x / 32 = pixel x).
 *
 * The reference calls OpenCV (an external dependency, imgproc/src/drawing.cpp of OpenCV 4.x); what is restated here is
 * its published integer algorithms -- clipLine, the 8-connected line iterator, the fixed-point Line2, FillConvexPoly,
 * ThickLine with round caps, PolyLine, and the midpoint Circle -- so that the planes are bit-identical to cv2's
 * (tests/test_raster_prims.py compares them with the installed cv2 on random primitives, including far-off-image
 * coordinates).  Host + device code: the CUDA raster kernel draws into shared memory with atomicOr, the CPU test
 * compiles the same functions with g++.
 */
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RP_FN __host__ __device__ inline
#define RP_BIG __host__ __device__ __noinline__ /* one copy of the larger routines per kernel: the raster kernel is instruction-fetch bound */
#else
#define RP_FN inline
#define RP_BIG inline
#endif

namespace raster {

constexpr int IMG_W = 640, IMG_H = 480, WORDS_PER_ROW = IMG_W / 32, PLANE_WORDS = IMG_H * WORDS_PER_ROW;
constexpr int XY_SHIFT = 16;
constexpr int64_t XY_ONE = (int64_t)1 << XY_SHIFT;

/* cv::Point2l.  OpenCV computes in 64-bit integers so that end points anywhere in int range work; the routines below are
   templates over the integer type and run in 32 bits whenever every coordinate they can produce fits (a primitive whose 16.16
   fixed-point coordinates stay below 2^29, i.e. within +-8191 px: all but the far-off-image ones) -- the same operations
   on the same values, half the integer instructions on a GPU without 64-bit integer ALUs. */
template <class I>
struct P2T {
  I x, y;
};
typedef P2T<int64_t> P2;
typedef P2T<int32_t> P2s;
constexpr int SMALL_PX = 8191; /* |coordinate| <= SMALL_PX px: 16.16 fixed point fits 32 bits with headroom for sums of two */

RP_FN void or_word(uint32_t* p, uint32_t m) {
#ifdef __CUDA_ARCH__
#ifdef DYNENV_BOUNDS_CHECK
  { /* the plane is the first PLANE_WORDS words of the block's shared memory (rc_raster.cuh) */
    const long long w = (long long)(p - reinterpret_cast<uint32_t*>(dynenv::g_smem));
    dynenv::bounds_check(w < 0 || w >= PLANE_WORDS ? -1 : (int)w);
  }
#endif
#ifdef RASTER_NONATOMIC /* dev ablation (racy: wrong pixels) */
  *p |= m;
#else
  atomicOr(p, m); /* (testing the word first -- bits only ever get set -- is slower: the load queues behind the atomics) */
#endif
#else
  *p |= m;
#endif
}
/* ICV_HLINE on row y for columns x1..x2 inclusive, both already inside the image (callers clip like OpenCV) */
RP_FN void hline(uint32_t* plane, int y, int x1, int x2) {
  if (x1 > x2) return;
  uint32_t* row = plane + y * WORDS_PER_ROW;
  const int w1 = x1 >> 5, w2 = x2 >> 5;
  const uint32_t m1 = 0xffffffffu << (x1 & 31), m2 = 0xffffffffu >> (31 - (x2 & 31));
  if (w1 == w2) { or_word(row + w1, m1 & m2); return; }
  or_word(row + w1, m1);
  for (int w = w1 + 1; w < w2; w++) or_word(row + w, 0xffffffffu);
  or_word(row + w2, m2);
}
template <class I>
RP_FN void put_point(uint32_t* plane, I x, I y) {
  if (x >= 0 && x < IMG_W && y >= 0 && y < IMG_H) or_word(plane + (int)y * WORDS_PER_ROW + ((int)x >> 5), 1u << ((int)x & 31));
}

/* a / b of C (truncation towards zero) for b > 0 and |a| < 2^53: on the device a float64 division and one correction
   instead of the ~100-instruction 64-bit integer division routine (the quotient of the correctly rounded division can
   only exceed the exact one by 1) */
RP_FN int64_t div_trunc(int64_t a, int64_t b) {
#ifdef __CUDA_ARCH__
  const bool neg = a < 0;
  const int64_t ua = neg ? -a : a;
  int64_t q = (int64_t)((double)ua / (double)b);
  if (q * b > ua) q--;
  return neg ? -q : q;
#else
  return a / b;
#endif
}

/* cv::clipLine(Size2l, Point2l&, Point2l&) */
template <class I>
RP_FN bool clip_line(I width, I height, P2T<I>& pt1, P2T<I>& pt2) {
  const I right = width - 1, bottom = height - 1;
  if (width <= 0 || height <= 0) return false;
  I &x1 = pt1.x, &y1 = pt1.y, &x2 = pt2.x, &y2 = pt2.y;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    I a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (I)((double)(a - y1) * (x2 - x1) / (y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (I)((double)(a - y2) * (x2 - x1) / (y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (I)((double)(a - x1) * (y2 - y1) / (x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (I)((double)(a - x2) * (y2 - y1) / (x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

/* Every drawing routine takes the calling lane and the number of cooperating lanes: its set-up (clipping, slopes,
   edge tables -- a few dozen integer operations) is computed redundantly by every lane, its pixels / rows are spread
   over the lanes in closed form.  The host build passes {0, 1}; the CUDA raster kernel a warp's {lane, 32}. */
struct Lanes {
  int lane, n;
};

/* Line(): cv::LineIterator, 8-connected, left to right, after clipping to the image.  Pixel i lies i steps along the
   major axis and m_i steps along the minor one, where m_i counts the earlier iterations whose error term was negative:
   m_i = max(0, ceil((2 dy i - dx) / (2 dx))) for the iterator's err_0 = dx - 2 dy, err += -2 dy (+ 2 dx when negative). */
RP_BIG void line_thin(uint32_t* plane, P2 pt1, P2 pt2, Lanes L) {
  if ((uint64_t)pt1.x >= (uint64_t)IMG_W || (uint64_t)pt2.x >= (uint64_t)IMG_W || (uint64_t)pt1.y >= (uint64_t)IMG_H ||
      (uint64_t)pt2.y >= (uint64_t)IMG_H) {
    if (!clip_line<int64_t>(IMG_W, IMG_H, pt1, pt2)) return;
  }
  int delta_x = 1, delta_y = 1;
  int dx = (int)(pt2.x - pt1.x), dy = (int)(pt2.y - pt1.y);
  if (dx < 0) { dx = -dx; dy = -dy; pt1 = pt2; } /* leftToRight */
  if (dy < 0) { dy = -dy; delta_y = -1; }
  const bool vert = dy > dx;
  if (vert) { int t = dx; dx = dy; dy = t; t = delta_x; delta_x = delta_y; delta_y = t; }
  /* major step (delta_x along the major axis), minor step delta_y; `vert` swaps which image axis is which */
  const int count = dx + 1;
  const int x0 = (int)pt1.x, y0 = (int)pt1.y;
  for (int i = L.lane; i < count; i += L.n) {
    const int num = 2 * dy * i - dx;
    const int m = num > 0 ? (num + 2 * dx - 1) / (2 * dx) : 0;
    const int x = vert ? x0 + m * delta_y : x0 + i * delta_x;
    const int y = vert ? y0 + i * delta_x : y0 + m * delta_y;
    put_point(plane, x, y);
  }
}

/* Line2(): line between fixed-point (XY_SHIFT) end points */
template <class I>
RP_BIG void line2(uint32_t* plane, P2T<I> pt1, P2T<I> pt2, Lanes L) {
  const I ONE = (I)XY_ONE;
  if (!clip_line<I>((I)IMG_W << XY_SHIFT, (I)IMG_H << XY_SHIFT, pt1, pt2)) return;
  I dx = pt2.x - pt1.x, dy = pt2.y - pt1.y;
  I j = dx < 0 ? -1 : 0;
  const I ax = (dx ^ j) - j;
  I i = dy < 0 ? -1 : 0;
  const I ay = (dy ^ i) - i;
  I x_step, y_step;
  int ecount;
  if (ax > ay) {
    dy = (dy ^ j) - j;
    pt1.x ^= pt2.x & j; pt2.x ^= pt1.x & j; pt1.x ^= pt2.x & j;
    pt1.y ^= pt2.y & j; pt2.y ^= pt1.y & j; pt1.y ^= pt2.y & j;
    x_step = ONE;
    y_step = (I)div_trunc((int64_t)dy << XY_SHIFT, (int64_t)(ax | 1)); /* |dy| <= ax: the quotient fits 17 bits */
    ecount = (int)((pt2.x - pt1.x) >> XY_SHIFT);
  } else {
    dx = (dx ^ i) - i;
    pt1.x ^= pt2.x & i; pt2.x ^= pt1.x & i; pt1.x ^= pt2.x & i;
    pt1.y ^= pt2.y & i; pt2.y ^= pt1.y & i; pt1.y ^= pt2.y & i;
    x_step = (I)div_trunc((int64_t)dx << XY_SHIFT, (int64_t)(ay | 1));
    y_step = ONE;
    ecount = (int)((pt2.y - pt1.y) >> XY_SHIFT);
  }
  pt1.x += (ONE >> 1);
  pt1.y += (ONE >> 1);
  if (L.lane == 0) put_point<I>(plane, (pt2.x + (ONE >> 1)) >> XY_SHIFT, (pt2.y + (ONE >> 1)) >> XY_SHIFT);
  if (L.n == 1) {
    /* one lane walks the whole line (the conic's poly-line segments, one per lane): its pixels are neighbours, and those of
       one 32-pixel word go out as ONE OR.  The raster kernel is bound by its shared-memory atomics -- behind the drawing
       barrier a block waits ~30 k cycles for them to drain, ~4 cycles per lane-level atomic and SM -- and half of them were
       the single pixels of these edges. */
    int cw = -1;
    uint32_t cm = 0;
    const bool horiz = ax > ay;
    const I xs = pt1.x >> XY_SHIFT, ys = pt1.y >> XY_SHIFT;
    for (int k = 0; k <= ecount; k++) {
      const I x = horiz ? xs + k : (pt1.x + k * x_step) >> XY_SHIFT, y = horiz ? (pt1.y + k * y_step) >> XY_SHIFT : ys + k;
      if (x >= 0 && x < IMG_W && y >= 0 && y < IMG_H) {
        const int w = (int)y * WORDS_PER_ROW + ((int)x >> 5);
        if (w != cw) {
          if (cm) or_word(plane + cw, cm);
          cw = w;
          cm = 0;
        }
        cm |= 1u << ((int)x & 31);
      }
    }
    if (cm) or_word(plane + cw, cm);
  } else { /* several lanes share the line: pixel k on lane k % n (one loop for both major axes: the kernel is short of instruction cache) */
    const bool horiz = ax > ay;
    const I xs = pt1.x >> XY_SHIFT, ys = pt1.y >> XY_SHIFT;
    for (int k = L.lane; k <= ecount; k += L.n)
      put_point<I>(plane, horiz ? xs + k : (pt1.x + k * x_step) >> XY_SHIFT, horiz ? (pt1.y + k * y_step) >> XY_SHIFT : ys + k);
  }
}

/* vertex k (k in 0..3) of a quadrilateral held in registers (no indexed array: local memory is what the raster kernel
   cannot afford, its blocks leave the SM almost no L1) */
template <class I>
struct QuadT {
  P2T<I> a, b, c, d;
};
typedef QuadT<int64_t> Quad;
template <class I>
RP_FN P2T<I> qget(const QuadT<I>& q, int k) { return k == 0 ? q.a : (k == 1 ? q.b : (k == 2 ? q.c : q.d)); }

/* one edge cursor of FillConvexPoly: up to three re-seeding events (row, x, dx) */
template <class I>
struct EdgeEvents {
  int n, y0, y1, y2;
  I x0, x1, x2, dx0, dx1, dx2;
};
template <class I>
RP_FN void ev_push(EdgeEvents<I>& e, int y, I x, I dx) {
  if (e.n == 0) { e.y0 = y; e.x0 = x; e.dx0 = dx; }
  else if (e.n == 1) { e.y1 = y; e.x1 = x; e.dx1 = dx; }
  else { e.y2 = y; e.x2 = x; e.dx2 = dx; }
  e.n++;
}
template <class I>
RP_FN I ev_x_at(const EdgeEvents<I>& e, int yy) { /* the cursor's x in row yy: last event at or above yy */
  if (e.n > 2 && e.y2 <= yy) return e.x2 + e.dx2 * (yy - e.y2);
  if (e.n > 1 && e.y1 <= yy) return e.x1 + e.dx1 * (yy - e.y1);
  return e.x0 + e.dx0 * (yy - e.y0);
}

/* FillConvexPoly() for 4 vertices, line_type LINE_8; `shift` = 0 (integer vertices) or XY_SHIFT.  OpenCV walks the
   rows with two edge cursors (x += dx per row) that it re-seeds at vertex rows; here the re-seeding events are found
   first (row, x, dx per cursor: the same control flow, jumping from vertex row to vertex row), then every row's span
   follows in closed form and the rows are spread over the lanes. */
template <class I>
RP_BIG void fill_convex4(uint32_t* plane, const QuadT<I> v, int shift, Lanes L) {
  const int npts = 4;
  int e_idx0, e_idx1, e_ye0, e_ye1; /* the two cursors: current vertex, end row */
  EdgeEvents<I> ev0, ev1;
  ev0.n = ev1.n = 0;
  ev0.y0 = ev0.y1 = ev0.y2 = ev1.y0 = ev1.y1 = ev1.y2 = 0;
  ev0.x0 = ev0.x1 = ev0.x2 = ev0.dx0 = ev0.dx1 = ev0.dx2 = 0;
  ev1.x0 = ev1.x1 = ev1.x2 = ev1.dx0 = ev1.dx1 = ev1.dx2 = 0;
  const int delta = 1 << shift >> 1;
  int i, y, imin = 0;
  int edges = npts;
  I xmin, xmax, ymin, ymax;
  const int delta1 = (int)(XY_ONE >> 1);
  P2T<I> p0 = v.d;
  p0.x <<= XY_SHIFT - shift;
  p0.y <<= XY_SHIFT - shift;
  xmin = xmax = v.a.x;
  ymin = ymax = v.a.y;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (i = 0; i < npts; i++) {
    P2T<I> p = qget(v, i);
    if (p.y < ymin) { ymin = p.y; imin = i; }
    ymax = ymax > p.y ? ymax : p.y;
    xmax = xmax > p.x ? xmax : p.x;
    xmin = xmin < p.x ? xmin : p.x;
    p.x <<= XY_SHIFT - shift;
    p.y <<= XY_SHIFT - shift;
    if (shift == 0) {
      P2 a, b;
      a.x = p0.x >> XY_SHIFT; a.y = p0.y >> XY_SHIFT; b.x = p.x >> XY_SHIFT; b.y = p.y >> XY_SHIFT;
      line_thin(plane, a, b, L);
    } else line2<I>(plane, p0, p, L);
    p0 = p;
  }
  xmin = (xmin + delta) >> shift;
  xmax = (xmax + delta) >> shift;
  ymin = (ymin + delta) >> shift;
  ymax = (ymax + delta) >> shift;
  if ((int)xmax < 0 || (int)ymax < 0 || (int)xmin >= IMG_W || (int)ymin >= IMG_H) return;
  ymax = ymax < IMG_H - 1 ? ymax : IMG_H - 1;
  e_idx0 = e_idx1 = imin;
  e_ye0 = e_ye1 = y = (int)ymin;
  int yend;
  for (;;) { /* row y: a cursor whose edge ends here takes the next edge that reaches below y */
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (i = 0; i < 2; i++) {
      const int ye = i ? e_ye1 : e_ye0;
      if (y >= ye) {
        int idx0 = i ? e_idx1 : e_idx0;
        const int di = i ? npts - 1 : 1;
        int idx = idx0 + di;
        if (idx >= npts) idx -= npts;
        int ty = 0;
        for (; edges-- > 0;) {
          const P2T<I> pv = qget(v, idx);
          ty = (int)((pv.y + delta) >> shift);
          if (ty > y) {
            I xs = qget(v, idx0).x, xe = pv.x;
            if (shift != XY_SHIFT) { xs <<= XY_SHIFT - shift; xe <<= XY_SHIFT - shift; }
            const I dxe = (I)div_trunc((int64_t)(xe - xs) * 2 + (ty - y), 2 * (int64_t)(ty - y));
            if (i) { e_ye1 = ty; e_idx1 = idx; ev_push<I>(ev1, y, xs, dxe); }
            else { e_ye0 = ty; e_idx0 = idx; ev_push<I>(ev0, y, xs, dxe); }
            break;
          }
          idx0 = idx;
          idx += di;
          if (idx >= npts) idx -= npts;
        }
      }
    }
    if (edges < 0) { yend = y - 1; break; }
    const int ynext = e_ye0 < e_ye1 ? e_ye0 : e_ye1;
    if (ynext > (int)ymax) { yend = (int)ymax; break; }
    y = ynext;
  }
  if (ev0.n == 0 || ev1.n == 0) return;
  const int ystart = (int)ymin > 0 ? (int)ymin : 0;
  for (int yy = ystart + L.lane; yy <= yend; yy += L.n) {
    const I xa = ev_x_at<I>(ev0, yy), xb = ev_x_at<I>(ev1, yy);
    const I xl = xa > xb ? xb : xa, xr = xa > xb ? xa : xb;
    int xx1 = (int)((xl + delta1) >> XY_SHIFT);
    int xx2 = (int)((xr + delta1) >> XY_SHIFT);
    if (xx2 >= 0 && xx1 < IMG_W) {
      if (xx1 < 0) xx1 = 0;
      if (xx2 >= IMG_W) xx2 = IMG_W - 1;
      hline(plane, yy, xx1, xx2);
    }
  }
}

/* Circle() with fill: Bresenham circle, horizontal spans.  The (dx, dy) sequence is walked by every lane; iteration k
   draws its four spans on lane k % n. */
RP_BIG void circle_filled(uint32_t* plane, int cx, int cy, int radius, Lanes L) {
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  int it = 0;
  while (dx >= dy) {
    /* the walk's next state first: OpenCV draws the rows cy -+ dx in EVERY iteration, each time a little wider (dy grows
       while dx stands); the spans of one row nest, so only the widest -- the last before dx changes or the walk ends --
       is drawn here: the same pixels with half the shared-memory writes (what the raster kernel is bound by) */
    const int dy1 = dy + 1;
    int err1 = err + plus;
    const int mask = (err1 <= 0) - 1;
    err1 -= minus & mask;
    const int dx1 = dx + mask;
    const bool widest = mask != 0 || dx1 < dy1;
    if (it % L.n == L.lane) {
      const int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
      int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
      if (x11 < IMG_W && x12 >= 0 && y21 < IMG_H && y22 >= 0) {
        x11 = x11 > 0 ? x11 : 0;
        x12 = x12 < IMG_W - 1 ? x12 : IMG_W - 1;
        if ((unsigned)y11 < (unsigned)IMG_H) hline(plane, y11, x11, x12);
        if (dy && (unsigned)y12 < (unsigned)IMG_H) hline(plane, y12, x11, x12);
        if (widest && x21 < IMG_W && x22 >= 0) {
          x21 = x21 > 0 ? x21 : 0;
          x22 = x22 < IMG_W - 1 ? x22 : IMG_W - 1;
          if ((unsigned)y21 < (unsigned)IMG_H) hline(plane, y21, x21, x22);
          if (dx && (unsigned)y22 < (unsigned)IMG_H) hline(plane, y22, x21, x22);
        }
      }
    }
    it++;
    dy = dy1;
    err = err1;
    plus += 2;
    dx = dx1;
    minus -= mask & 2;
  }
}

RP_FN int64_t cv_round(double v) { /* cvRound: round half to even (lrint in the default rounding mode) */
#ifdef __CUDA_ARCH__
  return (int64_t)__double2ll_rn(v);
#else
  return (int64_t)__builtin_llrint(v);
#endif
}

/* ThickLine() with integer end points; flags bit 0 / 1: round cap at p0 / p1 */
RP_BIG void thick_line(uint32_t* plane, P2 p0, P2 p1, int thickness, int flags, Lanes L) {
  const double INV_XY_ONE = 1. / XY_ONE;
  if (thickness <= 1) { /* shift == 0: the end points are integers */
    line_thin(plane, p0, p1, L);
    return;
  }
  { /* OpenCV 4.13 first clips the centre line to the image grown by `thickness` on every side */
    const int64_t m = thickness;
    p0.x += m; p0.y += m; p1.x += m; p1.y += m;
    if (!clip_line<int64_t>(IMG_W + 2 * m, IMG_H + 2 * m, p0, p1)) return;
    p0.x -= m; p0.y -= m; p1.x -= m; p1.y -= m;
  }
  p0.x <<= XY_SHIFT; p0.y <<= XY_SHIFT; p1.x <<= XY_SHIFT; p1.y <<= XY_SHIFT;
  Quad pt;
  P2 dp;
  dp.x = dp.y = 0;
  double dx = (p0.x - p1.x) * INV_XY_ONE, dy = (p1.y - p0.y) * INV_XY_ONE;
  double r = dx * dx + dy * dy;
  const int oddThickness = thickness & 1;
  const bool small = thickness <= 2048; /* the clipped end points lie within the image grown by the thickness: every corner within +-3712 px */
  thickness <<= XY_SHIFT - 1;
  if ((r < 0 ? -r : r) > 2.220446049250313e-16) {
#ifdef __CUDA_ARCH__
    r = (thickness + oddThickness * XY_ONE * 0.5) / sqrt(r);
#else
    r = (thickness + oddThickness * XY_ONE * 0.5) / __builtin_sqrt(r);
#endif
    dp.x = cv_round(dy * r);
    dp.y = cv_round(dx * r);
    pt.a.x = p0.x + dp.x; pt.a.y = p0.y + dp.y;
    pt.b.x = p0.x - dp.x; pt.b.y = p0.y - dp.y;
    pt.c.x = p1.x - dp.x; pt.c.y = p1.y - dp.y;
    pt.d.x = p1.x + dp.x; pt.d.y = p1.y + dp.y;
    if (small) {
      QuadT<int32_t> q;
      q.a.x = (int32_t)pt.a.x; q.a.y = (int32_t)pt.a.y; q.b.x = (int32_t)pt.b.x; q.b.y = (int32_t)pt.b.y;
      q.c.x = (int32_t)pt.c.x; q.c.y = (int32_t)pt.c.y; q.d.x = (int32_t)pt.d.x; q.d.y = (int32_t)pt.d.y;
      fill_convex4<int32_t>(plane, q, XY_SHIFT, L);
    } else fill_convex4<int64_t>(plane, pt, XY_SHIFT, L);
  }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int i = 0; i < 2; i++) {
    if (flags & (i + 1)) {
      const int cx = (int)((p0.x + (XY_ONE >> 1)) >> XY_SHIFT), cy = (int)((p0.y + (XY_ONE >> 1)) >> XY_SHIFT);
      circle_filled(plane, cx, cy, (int)((thickness + (XY_ONE >> 1)) >> XY_SHIFT), L);
    }
    p0 = p1;
  }
}

/* cv2.line(img, p0, p1, 1, thickness) */
RP_FN void cv_line(uint32_t* plane, int x0, int y0, int x1, int y1, int thickness, Lanes L) {
  P2 a, b;
  a.x = x0; a.y = y0; b.x = x1; b.y = y1;
  thick_line(plane, a, b, thickness, 3, L);
}
/* cv2.rectangle(img, p0, p1, 1, -1) */
RP_FN void cv_rectangle_filled(uint32_t* plane, int x0, int y0, int x1, int y1, Lanes L) {
  const bool small = x0 >= -SMALL_PX && x0 <= SMALL_PX && x1 >= -SMALL_PX && x1 <= SMALL_PX && y0 >= -SMALL_PX && y0 <= SMALL_PX &&
                     y1 >= -SMALL_PX && y1 <= SMALL_PX;
  if (small) {
    QuadT<int32_t> q;
    q.a.x = x0; q.a.y = y0; q.b.x = x1; q.b.y = y0; q.c.x = x1; q.c.y = y1; q.d.x = x0; q.d.y = y1;
    fill_convex4<int32_t>(plane, q, 0, L);
    return;
  }
  Quad pt;
  pt.a.x = x0; pt.a.y = y0; pt.b.x = x1; pt.b.y = y0; pt.c.x = x1; pt.c.y = y1; pt.d.x = x0; pt.d.y = y1;
  fill_convex4<int64_t>(plane, pt, 0, L);
}
/* segment i (1 <= i < count) of cv2.polylines(img, [pts], False, 1, thickness): PolyLine() caps the first segment at
   both ends and every later one at its far end */
RP_FN void cv_polyline_segment(uint32_t* plane, int xa, int ya, int xb, int yb, int thickness, int i, Lanes L) {
  P2 a, b;
  a.x = xa; a.y = ya; b.x = xb; b.y = yb;
  thick_line(plane, a, b, thickness, i == 1 ? 3 : 2, L);
}

}  // namespace raster
