// hrl_sensors.cuh - task-layer sensors shared by the fused step kernels and the stand-alone
// parity kernels.  Citations relative to /root/reference/hrl_pybullet_envs.
#pragma once
#include "hrl_math.cuh"

#define HRL_PI_D 3.14159265358979323846
#define HRL_PI_F 3.14159265358979323846f

// The config carries spans as float32; the reference's defaults are the float64 constants pi
// and 2*pi (ant_gather_env.py:22, ant_maze_bullet_env.py:23), so those two are snapped back.
__host__ __device__ __forceinline__ double snap_span(float s) {
  if (s == (float)(2.0 * HRL_PI_D)) return 2.0 * HRL_PI_D;
  if (s == (float)HRL_PI_D) return HRL_PI_D;
  return (double)s;
}

// Gather sector sensor for ONE item (envs/gather/ant_gather_env.py:143-162, twin
// gather_base.py:134-152).  Returns the bin or -1; *d2_out = squared xy distance in double
// (ant_gather_env.py:198-200), identical bit-for-bit to the reference's float64 value for
// float32-representable inputs (explicit _rn ops: no FMA contraction).
// Angle/bin: float32 fast path; when the float result is within 2e-4 of any decision
// boundary (bin edge, +-half span) the reference's exact float64 sequence is evaluated
// instead, which makes the bin index bit-exact (SURVEY.md hard part 5).
// Angle -> bin part, out of line on purpose: one copy instead of one per item keeps the
// once-per-step task layer small (its cost is instruction fetch, not arithmetic).
__device__ __noinline__ int gather_angle_bin(double dxd, double dyd, float yaw, int n_bins, float span) {
  float dx = (float)dxd, dy = (float)dyd;
  const float two_pi = 2.0f * HRL_PI_F;
  float a = atan2f(dy, dx) - yaw;  // :148
  // Python's `% (2 pi)` then `> pi -> -= 2 pi` (:151-153) == wrap to (-pi, pi].  Nearest-multiple
  // reduction handles any |a| (the gimbal branch of the Bullet yaw reaches +-2 pi); the two forms only
  // differ at |a| = pi, which is either outside the half span or "risky" and re-evaluated exactly below.
  a = fmaf(-two_pi, rintf(a * (1.0f / two_pi)), a);
  float half = 0.5f * span, res = span / (float)n_bins;
  float t = (a + half) / res;
  bool risky = fabsf(fabsf(a) - half) < 2e-4f || fabsf(t - rintf(t)) < 2e-4f || !(fabsf(a) < 1e30f);
  if (!risky) {
    if (fabsf(a) > half) return -1;  // :159
    int b = (int)t;                  // :161
    return b > n_bins - 1 ? n_bins - 1 : b;
  }
  // exact path: the reference's float64 expression sequence
  double ad = atan2(dyd, dxd) - (double)yaw;
  ad = fmod(ad, 2.0 * HRL_PI_D);
  if (ad < 0.0) ad += 2.0 * HRL_PI_D;
  if (ad > HRL_PI_D) ad -= 2.0 * HRL_PI_D;
  if (ad < -HRL_PI_D) ad += 2.0 * HRL_PI_D;
  double spand = snap_span(span);
  double halfd = spand * 0.5, resd = spand / (double)n_bins;
  if (fabs(ad) > halfd) return -1;
  int b = (int)((ad + halfd) / resd);
  // the reference raises IndexError at exactly +half span; clamp instead
  return b > n_bins - 1 ? n_bins - 1 : b;
}
__device__ __forceinline__ int gather_item_bin(float rx, float ry, float yaw, float ox, float oy, int n_bins,
                                               float sensor_range, float span, double* d2_out) {
  double dxd = __dsub_rn((double)ox, (double)rx), dyd = __dsub_rn((double)oy, (double)ry);
  double d2 = __dadd_rn(__dmul_rn(dxd, dxd), __dmul_rn(dyd, dyd));
  *d2_out = d2;
  if (d2 > (double)sensor_range) return -1;  // :145 (squared distance vs unsquared range, kept)
  return gather_angle_bin(dxd, dyd, yaw, n_bins, span);
}

// intersection_utils.py:93-104 (tie order 1,4,2,3)
__device__ __forceinline__ int quadrant_d(double x, double y) {
  if (x >= 0 && y >= 0) return 1;
  if (x >= 0 && y <= 0) return 4;
  if (x <= 0 && y >= 0) return 2;
  if (x <= 0 && y <= 0) return 3;
  return 0;
}

// One ray of the wall lidar, sizeable_enclosed_scene.py:63-97, evaluated in float64 with the
// reference's determinant formula (intersection_utils.py:84-90); ray and bounds are infinite
// lines.  bounds: n_lines x (x1,y1,x2,y2).
__device__ __forceinline__ float lidar_ray(int i, int n_bins, float span_f, float range_f, int n_lines,
                                           const float* __restrict__ bounds, float pxf, float pyf, float yawf) {
  const double span = snap_span(span_f), range = (double)range_f, px = (double)pxf, py = (double)pyf, yaw = (double)yawf;
  double ang;
  if (span == 2.0 * HRL_PI_D) ang = HRL_PI_D / 2 + yaw + ((double)(i + 1) / (double)n_bins) * span;  // :68-69
  else ang = HRL_PI_D / 2 + yaw + ((double)i / (double)(n_bins - 1)) * span;                           // :70-71
  double sn, cs;
  sincos(ang, &sn, &cs);
  double x2 = __dadd_rn(px, __dmul_rn(range, cs)), y2 = __dadd_rn(py, __dmul_rn(range, sn));
  int sq = quadrant_d(__dsub_rn(x2, px), __dsub_rn(y2, py));
  double best = 0.0;
  // fp32 pre-filter: with the ray written as P + t u (u = (cs, sn)), a bound line is hit at t = num / cr.  A line that
  // is certainly farther than `range`, or certainly behind the ray (t < 0 with both ray components away from 0, i.e. the
  // hit lies in the opposite quadrant), cannot contribute and is dropped; every other line - in particular every
  // knife edge (near-parallel, t ~ 0, axis-aligned ray) - goes through the reference's float64 expression sequence
  // below, so the result is bit-identical to evaluating all lines exactly.  The survivors are walked through a bit
  // mask so that the lanes of a warp spend their float64 passes on DIFFERENT lines at the same time (trip count =
  // the largest survivor count in the warp, typically 2-3 of 7) instead of idling through each other's lines.
  const float csf = (float)cs, snf = (float)sn;
  const bool clean_dir = fabsf(csf) > 1e-3f && fabsf(snf) > 1e-3f;
  unsigned live = 0;
  for (int l = 0; l < n_lines; l++) {
    const float x3f = bounds[4 * l], y3f = bounds[4 * l + 1], x34f = x3f - bounds[4 * l + 2], y34f = y3f - bounds[4 * l + 3];
    const float af = x3f - pxf, bf = y3f - pyf;
    const float n1 = af * y34f, n2 = bf * x34f, c1 = csf * y34f, c2 = snf * x34f;
    const float num = n1 - n2, cr = c1 - c2;
    const float en = 1e-6f * (fabsf(n1) + fabsf(n2)), ec = 1e-6f * (fabsf(c1) + fabsf(c2));
    bool drop = false;
    if (fabsf(cr) > 8.f * ec) {
      drop = fabsf(num) - en > range_f * 1.0001f * (fabsf(cr) + ec);                                                         // |t| > range
      drop = drop || (clean_dir && fabsf(num) > 8.f * en && (num < 0.f) != (cr < 0.f) && fabsf(num) > 1e-3f * fabsf(cr));  // t < 0
    }
    if (!drop) live |= 1u << l;
  }
  // the nearest accepted hit wins (1 - dist / range is monotone in dist, and so is the square root): the survivors only
  // track the smallest squared distance; one sqrt and one division per ray, after the loop, give the same double as
  // max over lines of 1 - sqrt(dd) / range.  (A hit with sqrt(dd) == range exactly yields 0 and never beats `best`,
  // so testing dd against range^2 cannot change the result either.)
  const double x12 = __dsub_rn(px, x2), y12 = __dsub_rn(py, y2), c12 = __dsub_rn(__dmul_rn(px, y2), __dmul_rn(py, x2));
  const double range2 = __dmul_rn(range, range);
  double best_dd = -1.0;
  for (; live; live &= live - 1) {
    const int l = __ffs(live) - 1;
    double x3 = bounds[4 * l], y3 = bounds[4 * l + 1], x4 = bounds[4 * l + 2], y4 = bounds[4 * l + 3];
    double x34 = __dsub_rn(x3, x4), y34 = __dsub_rn(y3, y4);
    double d = __dsub_rn(__dmul_rn(x12, y34), __dmul_rn(y12, x34));
    if (d == 0.0) continue;
    double c34 = __dsub_rn(__dmul_rn(x3, y4), __dmul_rn(y3, x4));
    double ix = __ddiv_rn(__dsub_rn(__dmul_rn(c12, x34), __dmul_rn(x12, c34)), d);
    double iy = __ddiv_rn(__dsub_rn(__dmul_rn(c12, y34), __dmul_rn(y12, c34)), d);
    double ddx = __dsub_rn(px, ix), ddy = __dsub_rn(py, iy);
    double dd = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
    if (dd > range2 && sqrt(dd) > range) continue;   // (the sqrt only runs in the sliver where the two tests could differ)
    if (sq != quadrant_d(__dsub_rn(ix, px), __dsub_rn(iy, py))) continue;
    if (best_dd < 0.0 || dd < best_dd) best_dd = dd;
  }
  if (best_dd >= 0.0) {
    const double val = 1.0 - sqrt(best_dd) / range;
    if (val > best) best = val;
  }
  return (float)best;
}

// intersection_utils.py:14-71 (orientation test + collinear cases), float64 like the reference
__device__ __forceinline__ int orient_d(double px, double py, double qx, double qy, double rx, double ry) {
  const double val = __dsub_rn(__dmul_rn(__dsub_rn(qy, py), __dsub_rn(rx, qx)), __dmul_rn(__dsub_rn(qx, px), __dsub_rn(ry, qy)));
  return val > 0 ? 1 : (val < 0 ? 2 : 0);
}
__device__ __forceinline__ bool on_seg_d(double px, double py, double qx, double qy, double rx, double ry) {
  return qx <= fmax(px, rx) && qx >= fmin(px, rx) && qy <= fmax(py, ry) && qy >= fmin(py, ry);
}
__device__ __forceinline__ bool segment_intersection_d(double p1x, double p1y, double q1x, double q1y, double p2x, double p2y,
                                                       double q2x, double q2y) {
  const int o1 = orient_d(p1x, p1y, q1x, q1y, p2x, p2y), o2 = orient_d(p1x, p1y, q1x, q1y, q2x, q2y);
  const int o3 = orient_d(p2x, p2y, q2x, q2y, p1x, p1y), o4 = orient_d(p2x, p2y, q2x, q2y, q1x, q1y);
  if (o1 != o2 && o3 != o4) return true;
  if (o1 == 0 && on_seg_d(p1x, p1y, p2x, p2y, q1x, q1y)) return true;
  if (o2 == 0 && on_seg_d(p1x, p1y, q2x, q2y, q1x, q1y)) return true;
  if (o3 == 0 && on_seg_d(p2x, p2y, p1x, p1y, q2x, q2y)) return true;
  if (o4 == 0 && on_seg_d(p2x, p2y, q1x, q1y, q2x, q2y)) return true;
  return false;
}

// Maze goal sector sensor, ant_maze_bullet_env.py:135-178 (sense_target=True): one reading
// 1 - walk_target_dist / range in the bin of the goal direction, nothing when the goal is out of
// range, outside the span, or hidden behind one of the 3 box_bounds segments (maze_scene.py:19-21).
// Evaluated in float64 with the reference's expression sequence (rare, non-default path).
__device__ __noinline__ int maze_target_bin(int n_bins, float span_f, float range_f, int n_box, const float* __restrict__ box,
                                            float rx, float ry, float yaw, float tx, float ty, float wtd) {
  if (n_bins <= 0 || (double)wtd > (double)range_f) return -1;
  for (int l = 0; l < n_box; l++)
    if (segment_intersection_d(rx, ry, tx, ty, box[4 * l], box[4 * l + 1], box[4 * l + 2], box[4 * l + 3])) return -1;
  double ad = atan2(__dsub_rn((double)ty, (double)ry), __dsub_rn((double)tx, (double)rx)) - (double)yaw;
  ad = fmod(ad, 2.0 * HRL_PI_D);
  if (ad < 0.0) ad += 2.0 * HRL_PI_D;
  if (ad > HRL_PI_D) ad -= 2.0 * HRL_PI_D;
  if (ad < -HRL_PI_D) ad += 2.0 * HRL_PI_D;
  const double spand = snap_span(span_f), halfd = spand * 0.5, resd = spand / (double)n_bins;
  if (fabs(ad) > halfd) return -1;
  const int b = (int)((ad + halfd) / resd);
  return b > n_bins - 1 ? n_bins - 1 : b;
}

// Bullet getEulerFromQuaternion (SURVEY.md A.3 "Queries").  Also returns cos/sin of the yaw: outside
// the gimbal branch they follow from the atan2 arguments with one rsqrt (no sincos evaluation).
__device__ __forceinline__ void quat_to_rpy(float x, float y, float z, float w, float& roll, float& pitch, float& yaw,
                                            float& cyaw, float& syaw) {
  float sarg = -2.f * (x * z - w * y);
  if (sarg <= -0.99999f) { pitch = -0.5f * HRL_PI_F; roll = 0.f; yaw = 2.f * atan2f(x, -y); sincosf(yaw, &syaw, &cyaw); }
  else if (sarg >= 0.99999f) { pitch = 0.5f * HRL_PI_F; roll = 0.f; yaw = 2.f * atan2f(-x, y); sincosf(yaw, &syaw, &cyaw); }
  else {
    pitch = asinf(sarg);
    roll = atan2f(2.f * (y * z + w * x), w * w - x * x - y * y + z * z);
    const float ys = 2.f * (x * y + w * z), yc = w * w + x * x - y * y - z * z;
    yaw = atan2f(ys, yc);
    const float h2 = ys * ys + yc * yc;
    if (h2 > 0.f) { const float ih = rsqrt_ftz(h2); cyaw = yc * ih; syaw = ys * ih; }
    else { cyaw = 1.f; syaw = 0.f; }  // atan2(0, 0) = 0
  }
}
