// Instantiations of gmpe_step_coop_kernel<float, SCEN, N> (see gmpe_coop_table.h).
#include "gmpe_coop.cuh"
#include "gmpe_coop_table.h"

namespace gmpe {
namespace {
using Real = float;

template <int SCEN, int N>
void bind(const void** fn, size_t* smem, int* warps) {
  *fn = (const void*)gmpe_step_coop_kernel<Real, SCEN, N>;
  *smem = (size_t)CoopTraits<Real, SCEN, N>::WARPS * CoopTraits<Real, SCEN, N>::SLAB_BYTES;
  *warps = CoopTraits<Real, SCEN, N>::WARPS;
}

template <int SCEN>
bool bind_n(int n, const void** fn, size_t* smem, int* warps) {
  switch (n) {
    case 2: bind<SCEN, 2>(fn, smem, warps); return true;
    case 3: bind<SCEN, 3>(fn, smem, warps); return true;
    case 4: bind<SCEN, 4>(fn, smem, warps); return true;
    case 5: bind<SCEN, 5>(fn, smem, warps); return true;
    case 6: bind<SCEN, 6>(fn, smem, warps); return true;
    case 7: bind<SCEN, 7>(fn, smem, warps); return true;
    case 8: bind<SCEN, 8>(fn, smem, warps); return true;
    case 9: bind<SCEN, 9>(fn, smem, warps); return true;
    case 10: bind<SCEN, 10>(fn, smem, warps); return true;
    default: return false;
  }
}
}  // namespace

bool coop_lookup_f32(int scenario, int n, const void** fn, size_t* smem, int* warps) {
  if (n > COOP_MAX_AGENTS) return coop_lookup_wide_f32(scenario, n, fn, smem, warps);
  switch (scenario) {
    case GMPE_SCEN_JULY_TUBE: return bind_n<GMPE_SCEN_JULY_TUBE>(n, fn, smem, warps);
    case GMPE_SCEN_NAV_GRAPH: return bind_n<GMPE_SCEN_NAV_GRAPH>(n, fn, smem, warps);
    case GMPE_SCEN_THREE_PHASE: return bind_n<GMPE_SCEN_THREE_PHASE>(n, fn, smem, warps);
    case GMPE_SCEN_TWO_PHASE: return bind_n<GMPE_SCEN_TWO_PHASE>(n, fn, smem, warps);
    case GMPE_SCEN_CORRIDOR_ROT_INV: return bind_n<GMPE_SCEN_CORRIDOR_ROT_INV>(n, fn, smem, warps);
    default: return false;
  }
}
}  // namespace gmpe
