// Lookup of the lane-per-agent kernel instantiations (gmpe_coop.cuh).  They are compiled in
// their own translation units -- built in parallel by build.py -- because every (scenario, N)
// pair is a separate kernel: N = 2 .. 10 for every scenario (worlds with N == L agents /
// landmarks and no obstacles or walls), and N = 11 .. 16 for the rot-inv family, which only this
// kernel serves (gmpe_coop_wide_f32.cu / _f64.cu).
#pragma once
#include <cstddef>

namespace gmpe {
// false: no instantiation for this (scenario, N); fn / smem_bytes / warps untouched.
// warps: warps per CTA the instantiation was compiled for (its launch bound)
bool coop_lookup_f32(int scenario, int num_agents, const void** fn, size_t* smem_bytes, int* warps);
bool coop_lookup_f64(int scenario, int num_agents, const void** fn, size_t* smem_bytes, int* warps);
bool coop_lookup_wide_f32(int scenario, int num_agents, const void** fn, size_t* smem_bytes, int* warps);
bool coop_lookup_wide_f64(int scenario, int num_agents, const void** fn, size_t* smem_bytes, int* warps);
constexpr int COOP_MIN_AGENTS = 2, COOP_MAX_AGENTS = 10, COOP_MAX_AGENTS_ROT_INV = 16;
}  // namespace gmpe
