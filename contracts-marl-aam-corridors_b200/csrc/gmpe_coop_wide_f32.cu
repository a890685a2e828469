// Instantiations of gmpe_step_coop_kernel<float, SCEN, N> for the wide worlds of the rot-inv
// family, COOP_MAX_AGENTS < N <= COOP_MAX_AGENTS_ROT_INV (see gmpe_coop_table.h): still one lane per
// agent (E = 2N <= 32 keeps a row of the graph in one 32-bit mask), two worlds per warp up to N = 16.
#include "gmpe_coop.cuh"
#include "gmpe_coop_table.h"

namespace gmpe {
namespace {
using Real = float;

template <int SCEN, int N>
void bind(const void** fn, size_t* smem, int* warps) {
  using T = CoopTraits<Real, SCEN, N>;
  static_assert(T::E <= 32, "a row of the graph must fit a 32-bit mask");
  *fn = (const void*)gmpe_step_coop_kernel<Real, SCEN, N>;
  *smem = (size_t)T::WARPS * T::SLAB_BYTES;
  *warps = T::WARPS;
}

template <int SCEN>
bool bind_n(int n, const void** fn, size_t* smem, int* warps) {
  switch (n) {
    case 11: bind<SCEN, 11>(fn, smem, warps); return true;
    case 12: bind<SCEN, 12>(fn, smem, warps); return true;
    case 13: bind<SCEN, 13>(fn, smem, warps); return true;
    case 14: bind<SCEN, 14>(fn, smem, warps); return true;
    case 15: bind<SCEN, 15>(fn, smem, warps); return true;
    case 16: bind<SCEN, 16>(fn, smem, warps); return true;
    default: return false;
  }
}
}  // namespace

bool coop_lookup_wide_f32(int scenario, int n, const void** fn, size_t* smem, int* warps) {
  switch (scenario) {
    case GMPE_SCEN_THREE_PHASE: return bind_n<GMPE_SCEN_THREE_PHASE>(n, fn, smem, warps);
    case GMPE_SCEN_TWO_PHASE: return bind_n<GMPE_SCEN_TWO_PHASE>(n, fn, smem, warps);
    case GMPE_SCEN_CORRIDOR_ROT_INV: return bind_n<GMPE_SCEN_CORRIDOR_ROT_INV>(n, fn, smem, warps);
    default: return false;
  }
}
}  // namespace gmpe
