// Host-only part of the C ABI: the synthetic-problem generator (sim_generator.cpp) and the default parameters, linked into
// libasc_host.so WITHOUT any CUDA code.  bench.py --impl reference and the CPU tests that only need problems load this library,
// so the reference arm never maps the product library libasc_b200.so.
#include <cstring>

#include "../../include/asc.h"

extern "C" {

#include "default_params.inl"

const char* asc_last_error(void) { return "libasc_host.so: host-side generator only (no compute entry points)"; }
const char* asc_version(void) { return "arm_slam_calib_b200 host library (generator + defaults)"; }

}  // extern "C"
