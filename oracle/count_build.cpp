// count_build.cpp - the oracle compiled as C++ with the counting `real` of flop_counter.hpp (TEST INFRASTRUCTURE).
// Exports the oracle's usual C API plus hrlo_flop_counts_get / _reset.  Built by `make count` into
// _build/libhrl_oracle_count.so; used only by tools/count_oracle_flops.py.
#define HRLO_COUNT
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "flop_counter.hpp"
extern "C" {
hrlo_flop_counts g_hrlo_flops = {0, 0, 0, 0, 0, 0};
#include "hrl_oracle.c"
void hrlo_flop_counts_get(uint64_t out[6]) {
  out[0] = g_hrlo_flops.add; out[1] = g_hrlo_flops.mul; out[2] = g_hrlo_flops.div;
  out[3] = g_hrlo_flops.sqrt_; out[4] = g_hrlo_flops.trig; out[5] = g_hrlo_flops.fabs_;
}
void hrlo_flop_counts_reset(void) { g_hrlo_flops = hrlo_flop_counts{0, 0, 0, 0, 0, 0}; }
}
