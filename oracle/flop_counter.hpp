// flop_counter.hpp - a counting stand-in for `real` (oracle/hrl_oracle.c compiled as C++ with -DHRLO_COUNT).
// TEST / MEASUREMENT INFRASTRUCTURE: tools/count_oracle_flops.py builds libhrl_oracle_count.so with it and reads the
// counters to fit the coefficients of the algorithmic FLOP model of hrl_pybullet_envs_b200/roofline.py (SURVEY.md 8(d):
// "replace the coefficients by an instrumented operation count of the oracle").  Every arithmetic operator and math
// call on a `real` bumps a counter; comparisons, copies and conversions are free.
#pragma once
#include <cmath>
#include <cstdint>

struct hrlo_flop_counts {
  uint64_t add, mul, div, sqrt_, trig, fabs_;
};
extern "C" hrlo_flop_counts g_hrlo_flops;

struct creal {
  double v;
  creal() = default;
  creal(double x) : v(x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return (float)v; }
  explicit operator int() const { return (int)v; }
};
inline creal operator+(creal a, creal b) { g_hrlo_flops.add++; return creal(a.v + b.v); }
inline creal operator-(creal a, creal b) { g_hrlo_flops.add++; return creal(a.v - b.v); }
inline creal operator*(creal a, creal b) { g_hrlo_flops.mul++; return creal(a.v * b.v); }
inline creal operator/(creal a, creal b) { g_hrlo_flops.div++; return creal(a.v / b.v); }
inline creal operator-(creal a) { return creal(-a.v); }
inline creal& operator+=(creal& a, creal b) { g_hrlo_flops.add++; a.v += b.v; return a; }
inline creal& operator-=(creal& a, creal b) { g_hrlo_flops.add++; a.v -= b.v; return a; }
inline creal& operator*=(creal& a, creal b) { g_hrlo_flops.mul++; a.v *= b.v; return a; }
inline creal& operator/=(creal& a, creal b) { g_hrlo_flops.div++; a.v /= b.v; return a; }
inline bool operator<(creal a, creal b) { return a.v < b.v; }
inline bool operator>(creal a, creal b) { return a.v > b.v; }
inline bool operator<=(creal a, creal b) { return a.v <= b.v; }
inline bool operator>=(creal a, creal b) { return a.v >= b.v; }
inline bool operator==(creal a, creal b) { return a.v == b.v; }
inline bool operator!=(creal a, creal b) { return a.v != b.v; }
inline creal c_sqrt(creal a) { g_hrlo_flops.sqrt_++; return creal(std::sqrt(a.v)); }
inline creal c_sin(creal a) { g_hrlo_flops.trig++; return creal(std::sin(a.v)); }
inline creal c_cos(creal a) { g_hrlo_flops.trig++; return creal(std::cos(a.v)); }
inline creal c_atan2(creal a, creal b) { g_hrlo_flops.trig++; return creal(std::atan2(a.v, b.v)); }
inline creal c_asin(creal a) { g_hrlo_flops.trig++; return creal(std::asin(a.v)); }
inline creal c_fabs(creal a) { g_hrlo_flops.fabs_++; return creal(std::fabs(a.v)); }
inline creal c_floor(creal a) { return creal(std::floor(a.v)); }
inline bool isfinite(creal a) { return std::isfinite(a.v); }
inline bool isnan(creal a) { return std::isnan(a.v); }
