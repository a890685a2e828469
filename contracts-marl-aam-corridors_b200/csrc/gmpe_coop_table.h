// Lookup of the lane-per-agent kernel instantiations (gmpe_coop.cuh).  They are compiled in
// their own translation units -- one per precision, built in parallel by build.py -- because
// every (scenario, N) pair is a separate kernel: N = 2 .. 10 for every scenario (worlds with
// N == L agents / landmarks and no obstacles or walls).
#pragma once
#include <cstddef>

namespace gmpe {
// false: no instantiation for this (scenario, N); fn / smem_bytes untouched
bool coop_lookup_f32(int scenario, int num_agents, const void** fn, size_t* smem_bytes);
bool coop_lookup_f64(int scenario, int num_agents, const void** fn, size_t* smem_bytes);
constexpr int COOP_MIN_AGENTS = 2, COOP_MAX_AGENTS = 10;
}  // namespace gmpe
