// gmpe_place.cuh -- episode reset on the device: the uniform-draw source and the placement of
// one small world in registers (lane-per-agent kernel, gmpe_coop.cuh).
//
// random_scenario + setup_tube_params (july:440-613, three_phase_graph.py:416-587), the landmark
// formations of multiagent/custom_scenarios/utils.py (set_landmarks_in_point :165-193,
// set_landmarks_in_line :77-131, set_landmarks_in_circle :231-262); navigation placement is
// defined-by-build.  Draw order and arithmetic are those of place_world
// (gmpe_step.cuh) with O = W = 0.
#pragma once
#include "gmpe_step.cuh"

namespace gmpe {

template <typename Real, int N>
struct Placed {
  typename RT<Real>::V4 s[N];
  typename RT<Real>::V2 lm[N];
  Real wf[12];
  bool overflow;  // a rejection loop hit GMPE_PLACE_MAX_ATTEMPTS and kept its last draw
};

template <typename Real, int SCEN, int N>
__device__ __noinline__ Placed<Real, N> place_small(const GmpeConfig& c, const DrawSrc& ds) {
  Placed<Real, N> o;
  o.overflow = false;
  uint32_t idx = 0;
  const double ws = c.world_size;
  const bool heading = c.dynamics != GMPE_DYN_DOUBLE_INTEGRATOR;
  const Real sep = Real(c.separation_dist);
#pragma unroll
  for (int q = 0; q < 12; ++q) o.wf[q] = Real(0);
  if (SCEN == GMPE_SCEN_JULY_TUBE || SCEN == GMPE_SCEN_THREE_PHASE || SCEN == GMPE_SCEN_TWO_PHASE || SCEN == GMPE_SCEN_CORRIDOR_ROT_INV) {
    // three_phase_graph.py:416-587 differs from july:440-613 by one extra draw (tube length) and
    // two spawn constants
    constexpr bool sc = SCEN == GMPE_SCEN_CORRIDOR_ROT_INV;  // sc:441-622: fixed tube length
    constexpr bool tp = SCEN == GMPE_SCEN_THREE_PHASE || SCEN == GMPE_SCEN_TWO_PHASE || sc;  // tp2:416-587 is identical
    const double jit = tp ? 0.3 : 0.2, spread = tp ? 3.0 : 5.0;
    (void)ds.u01(idx++);  // wall_length draw (july:368), value unused
    const double width = fmax(3 * c.entity_size * 2.5, ws * 0.15);
    const double angle = ds.uniform(idx++, -1.5707963267948966, 1.5707963267948966);
    double tl = ws * 0.8;
    if (tp && !sc) tl += ds.uniform(idx++, -ws * 0.3, ws * 0.1);  // tp:503
    double ca, sa;
    sincos(angle, &sa, &ca);
    const double ent0 = sa * (tl / 4), ent1 = ca * (tl / 4), ext0 = sa * (-tl / 4), ext1 = ca * (-tl / 4);
    const double Lh = sqrt((ext0 - ent0) * (ext0 - ent0) + (ext1 - ent1) * (ext1 - ent1)) + 1e-9;
    const double e0 = (ext0 - ent0) / Lh, e1 = (ext1 - ent1) / Lh;
    o.wf[GMPE_WF_ENT_X] = Real(ent0); o.wf[GMPE_WF_ENT_Y] = Real(ent1);
    o.wf[GMPE_WF_EXIT_X] = Real(ext0); o.wf[GMPE_WF_EXIT_Y] = Real(ext1);
    o.wf[GMPE_WF_E_X] = Real(e0); o.wf[GMPE_WF_E_Y] = Real(e1);
    o.wf[GMPE_WF_N_X] = Real((float)(-e1)); o.wf[GMPE_WF_N_Y] = Real((float)e0);
    o.wf[GMPE_WF_L] = Real(Lh); o.wf[GMPE_WF_HALF_W] = Real(width * 0.5);
    o.wf[GMPE_WF_WIDTH] = Real(width); o.wf[GMPE_WF_ANGLE] = Real(angle);
#pragma unroll
    for (int k = 0; k < N; ++k) {  // july:452-486, rejection against the agents placed so far
      Real px, py;
      bool bad;
      int tries = 0;
      do {
        const double jx = jit * ds.uniform(idx, -ws, ws);
        const double jy = jit * ds.uniform(idx + 1, -ws, ws);
        idx += 2;
        const double d = (ws + k) / spread;
        px = Real(ent0 + d * sa + jx); py = Real(ent1 + d * ca + jy);
        bad = false;
#pragma unroll
        for (int a = 0; a < k; ++a) bad = bad || (norm2(o.s[a].x - px, o.s[a].y - py) < sep);
        if (bad && ++tries >= GMPE_PLACE_MAX_ATTEMPTS) { o.overflow = true; bad = false; }
      } while (bad);
      Real th = Real(0), v = Real(0);
      if (heading) { th = wrap_heading(Real(ds.uniform(idx++, 0.0, 6.283185307179586))); v = Real(c.v_min); }
      o.s[k] = RT<Real>::mk4(px, py, th, v);
    }
#pragma unroll
    for (int l = 0; l < N; ++l) {
      double lx, ly;
      formation_landmark(c, l, N, ext0, ext1, sa, ca, lx, ly);
      o.lm[l] = RT<Real>::mk2(Real(lx), Real(ly));
    }
  } else {
    const double h = ws / 2;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      Real px, py;
      bool bad;
      int tries = 0;
      do {
        px = Real(0.9 * ds.uniform(idx, -h, h));
        py = Real(0.9 * ds.uniform(idx + 1, -h, h));
        idx += 2;
        bad = false;
#pragma unroll
        for (int a = 0; a < k; ++a) bad = bad || (norm2(o.s[a].x - px, o.s[a].y - py) < sep);
        if (bad && ++tries >= GMPE_PLACE_MAX_ATTEMPTS) { o.overflow = true; bad = false; }
      } while (bad);
      Real th = Real(0), v = Real(0);
      if (heading) { th = wrap_heading(Real(ds.uniform(idx++, 0.0, 6.283185307179586))); v = Real(c.v_min); }
      o.s[k] = RT<Real>::mk4(px, py, th, v);
    }
    const Real lmin = Real(1.2 * (2 * c.entity_size));  // is_landmark_collision july:916-925
#pragma unroll
    for (int k = 0; k < N; ++k) {
      Real px, py;
      bool bad;
      int tries = 0;
      do {
        px = Real(0.9 * ds.uniform(idx, -h, h));
        py = Real(0.9 * ds.uniform(idx + 1, -h, h));
        idx += 2;
        bad = false;
#pragma unroll
        for (int a = 0; a < k; ++a) bad = bad || (norm2(o.lm[a].x - px, o.lm[a].y - py) < lmin);
        if (bad && ++tries >= GMPE_PLACE_MAX_ATTEMPTS) { o.overflow = true; bad = false; }
      } while (bad);
      o.lm[k] = RT<Real>::mk2(px, py);
    }
  }
  return o;
}

}  // namespace gmpe
