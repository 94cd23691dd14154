// CPU build of dynenv_b200/csrc/raster_prims.cuh for tests/test_raster_prims.py (compared with the installed cv2)
#include "../../dynenv_b200/csrc/raster_compose.cuh"
extern "C" {
// one raw record (float32 [RC_RAW_WORDS]) -> the 8 planes of the IMAGE observation (zeroed by the caller)
void rp_render(const float* raw, uint32_t* planes) {
  raster::render_camera_host(raw, 0, planes);
  raster::render_camera_host(raw, 1, planes + 4 * raster::PLANE_WORDS);
}
// `lanes` > 1 emulates the cooperating lanes of a warp one after the other (same pixels, any order)
static int g_lanes = 1;
void rp_set_lanes(int n) { g_lanes = n; }
#define FOR_LANES for (raster::Lanes L = {0, g_lanes}; L.lane < g_lanes; L.lane++)
void rp_line(uint32_t* plane, int x0, int y0, int x1, int y1, int thickness) { FOR_LANES raster::cv_line(plane, x0, y0, x1, y1, thickness, L); }
void rp_rect(uint32_t* plane, int x0, int y0, int x1, int y1) { FOR_LANES raster::cv_rectangle_filled(plane, x0, y0, x1, y1, L); }
void rp_circle(uint32_t* plane, int cx, int cy, int r) { FOR_LANES raster::circle_filled(plane, cx, cy, r, L); }
// one segment of a poly-line: i = 1 caps both ends, i >= 2 the far end only (PolyLine)
void rp_segment(uint32_t* plane, int xa, int ya, int xb, int yb, int thickness, int i) {
  FOR_LANES raster::cv_polyline_segment(plane, xa, ya, xb, yb, thickness, i, L);
}
void rp_polyline(uint32_t* plane, const int* pts, int n, int thickness) {
  for (int i = 1; i < n; i++) FOR_LANES raster::cv_polyline_segment(plane, pts[2 * i - 2], pts[2 * i - 1], pts[2 * i], pts[2 * i + 1], thickness, i, L);
}
}
