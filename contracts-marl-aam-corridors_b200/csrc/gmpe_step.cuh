// gmpe_step.cuh -- the fused GraphMPE step kernel (sm_100a).
//
// One launch advances every world of the batch by one environment step and emits every
// tensor the reference's MultiAgentGraphEnv.step returns (multiagent/environment.py:
// 1021-1063), plus the auto-reset of graphworker (onpolicy/envs/env_wrappers.py:861-873).
//
// Mapping.  A CTA owns a tile of WPB consecutive worlds.  State is world-major
// ([B][N][4] etc.), so in phase A "thread = (world, agent)" reads and writes it with
// perfectly coalesced 128-bit accesses; entity state of the tile is staged in shared
// memory and every pairwise term (collision forces, collision penalties, neighbour
// selection, adjacency) is computed from there.  The reference evaluates agents in index
// order with in-place side effects; the kernel reproduces that exactly with
// a two-pass formulation: pass A computes what depends on an agent's own state only
// (phase machine, goal reach, heading re-draw), pass B evaluates agent j against
// post-reward values of agents < j and pre-reward values of agents > j.
// In phase B all threads of the CTA generate the output tensors element-parallel from
// shared memory, so that global stores are contiguous 128-bit vectors across the tile
// (obs is staged, node_obs / adj are generated in place, edges are stream-compacted per
// world with warp ballot + popc into the fixed-capacity padded buffers).
//
// The path is bandwidth / latency bound: no tensor cores by design.
#pragma once
#include <type_traits>

#include "gmpe_device.cuh"

namespace gmpe {

enum { MODE_STEP = 0, MODE_RESET = 1 };

// The first fifteen pointers of GmpeStepOut (everything but phase_reached, which the parameter block carries at
// its END as out_phase_reached).  The tile kernels proved sensitive to the offsets of the parameter block: with the
// sixteenth pointer in place here -- every later field 8 bytes further on -- the same source compiled to a
// different schedule and the N = 64 / N = 16 instantiations ran 6 % / 4 % slower (same-box A/B,
// profiles/r2_sweep.txt), so the fields those kernels read keep the offsets they had.
struct KOut {
  void *obs, *node_obs, *adj, *reward;
  uint8_t* done;
  void* info;
  int32_t* edge_index;
  void* edge_attr;
  int32_t *num_edges, *gnn_edge_index;
  void* gnn_edge_attr;
  int32_t* gnn_num_edges;
  void* indiv_reward;
  float *mask, *active_mask;
};
static_assert(sizeof(KOut) == 15 * sizeof(void*) && offsetof(GmpeStepOut, phase_reached) == sizeof(KOut) &&
              offsetof(KOut, active_mask) == offsetof(GmpeStepOut, active_mask) && offsetof(KOut, info) == offsetof(GmpeStepOut, info),
              "KOut mirrors the leading fields of GmpeStepOut");

struct KParams {
  GmpeConfig c;
  GmpeState st;
  KOut out;
  GmpeState rst;  // injected reset state (valid when has_rst)
  const int32_t* act;
  const void* theta;
  const uint8_t* env_mask;
  int has_rst;
  int mode;
  int wpb;        // worlds per tile
  int num_tiles;
  int E, F;       // entities per world, obs width
  int stage_graph; // small worlds: adjacency / edges staged in shared memory (see Smem)
  FastDiv dN, dE, dL, dO, dW, dF, d16, dE2, dNE2, dEE;
  int xmode;       // GMPE_EXPERIMENTS builds only (tools/ab.py): 1 = skip physics/reward, 2 = skip emission
  const void* act_tab;  // Real [32][2]: _set_action table (decode_action evaluated on the host), device memory owned by the handle
  const double* reset_draws;  // GmpeStepInject.reset_draws (already re-based on the launch's first world) or nullptr
  int reset_draws_n;          // ... doubles per world
  unsigned long long* overflow_cnt;  // worlds whose placement gave up a rejection loop (device counter owned by the handle)
  // lane-per-agent kernel (gmpe_coop.cuh): the round structure of the persistent grid, worked out
  // on the host (the 64-bit divisions cost ~150 instructions at the head of every warp), and the
  // _set_action table itself (Real [32][2]) so that the lookup is a constant-bank load instead of
  // a second dependent global load behind the action index
  int coop_groups;       // ceil(worlds of the launch / worlds per warp)
  int coop_full_rounds;  // coop_groups / (grid * warps per CTA)
  alignas(16) unsigned char act_tab_c[512];
  uint8_t* out_phase_reached;  // GmpeStepOut.phase_reached (rot-inv family, lane-per-agent kernel only)
};

// ---- shared memory carve-up --------------------------------------------------------
// The dynamic shared memory of the tile kernel.  The carve-up is kept as 32-bit OFFSETS into
// this array, not as pointers: the table does not fit in registers next to everything else,
// and a pointer that has been through local memory comes back as a generic address -- every
// access to the tile then went down the global load path (LD.E / ST.E, long-scoreboard
// stalls) instead of LDS / STS.  An offset added to the array keeps the address space.
extern __shared__ __align__(32) unsigned char gmpe_tile_smem[];

template <typename T>
struct SPtr {
  uint32_t off;
#ifdef __CUDA_ARCH__
  __device__ __forceinline__ operator T*() const { return reinterpret_cast<T*>(gmpe_tile_smem + off); }
#endif
};

template <typename Real>
struct Smem {
  // -- live for the whole tile
  SPtr<typename RT<Real>::V2> pos;      // [WPB*E]   post-step positions of every entity
  SPtr<typename RT<Real>::V2> velpre;   // [WPB*N]   p_vel before reward()
  SPtr<typename RT<Real>::V2> velpost;  // [WPB*N]   p_vel after reward() (heading re-draw on goal reach)
  SPtr<Real> wf;                        // [WPB*16]
  SPtr<Real> wl;                        // [WPB*W*4]
  SPtr<int32_t> wstep;                  // [WPB]
  SPtr<int32_t> wep;                    // [WPB]
  SPtr<uint8_t> stpre;                  // [WPB*N]
  SPtr<uint8_t> stpost;                 // [WPB*N]
  SPtr<uint8_t> mask;                   // [WPB*E]   entity disconnected from the graph
  SPtr<uint8_t> wreset;                 // [WPB]
  SPtr<uint8_t> wwrite;                 // [WPB]
  // -- phase-A scratch (dead once obs has left the SM) ...
  SPtr<typename RT<Real>::V4> nst;      // [WPB*N]   reset: freshly placed agent states
  SPtr<Real> obs;                       // [WPB*N*F] staging
  SPtr<Real> rew;                       // [WPB*N]
  SPtr<Real> sperr;                     // [WPB*N]
  SPtr<int32_t> cold;                   // [WPB*N*3] info counters before this step
  SPtr<int32_t> cnew;                   // [WPB*N*3] ... after
  SPtr<uint8_t> done;                   // [WPB*N]
  // -- ... aliased by the phase-B staging of small worlds (KParams::stage_graph)
  SPtr<Real> adjs;                      // [WPB*E*E]
  SPtr<int32_t> eidx;                   // [WPB*2*cap]
  SPtr<Real> eattr;                     // [WPB*cap]
  SPtr<int32_t> gidx;                   // [WPB*2*cap]
  SPtr<Real> gattr;                     // [WPB*cap]
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// stage_graph: 0 none, bit0 adjacency + env edges, bit1 gnn edges as well
template <typename Real>
__host__ __device__ inline size_t smem_layout(int wpb, int N, int E, int W, int F, int stage_graph, Smem<Real>* s) {
  size_t o = 0;
  auto take = [&](size_t bytes, size_t al) { o = align_up(o, al); size_t r = o; o += bytes; return r; };
  const size_t R = sizeof(Real);
  const size_t cap = (size_t)E * (E - 1);
  size_t o_pos = take(2 * R * wpb * E, 16);
  size_t o_vpre = take(2 * R * wpb * N, 16);
  size_t o_vpost = take(2 * R * wpb * N, 16);
  size_t o_wf = take(R * wpb * GMPE_WORLD_F, 16);
  size_t o_wl = take(R * wpb * (W > 0 ? W : 1) * GMPE_WALL_F, 16);
  size_t o_wstep = take(4 * wpb, 16);
  size_t o_wep = take(4 * wpb, 16);
  size_t o_stpre = take(wpb * N, 16);
  size_t o_stpost = take(wpb * N, 16);
  size_t o_mask = take(wpb * E, 16);
  size_t o_wreset = take(wpb, 16);
  size_t o_wwrite = take(wpb, 16);
  o = align_up(o, 32);
  const size_t scratch0 = o;
  size_t o_nst = take(4 * R * wpb * N, 32);
  size_t o_obs = take(R * wpb * N * F, 16);
  size_t o_rew = take(R * wpb * N, 16);
  size_t o_sp = take(R * wpb * N, 16);
  size_t o_cold = take(4 * wpb * N * 3, 16);
  size_t o_cnew = take(4 * wpb * N * 3, 16);
  size_t o_done = take(wpb * N, 16);
  const size_t scratch_end = o;
  o = scratch0;
  size_t o_adjs = 0, o_eidx = 0, o_eattr = 0, o_gidx = 0, o_gattr = 0;
  if (stage_graph) {
    o_adjs = take(R * wpb * E * E, 16);
    o_eidx = take(4 * wpb * 2 * cap, 16);
    o_eattr = take(R * wpb * cap, 16);
    if (stage_graph & 2) {
      o_gidx = take(4 * wpb * 2 * cap, 16);
      o_gattr = take(R * wpb * cap, 16);
    }
  }
  if (o < scratch_end) o = scratch_end;
  if (s) {
    s->pos.off = (uint32_t)o_pos;
    s->velpre.off = (uint32_t)o_vpre;
    s->velpost.off = (uint32_t)o_vpost;
    s->wf.off = (uint32_t)o_wf;
    s->wl.off = (uint32_t)o_wl;
    s->wstep.off = (uint32_t)o_wstep;
    s->wep.off = (uint32_t)o_wep;
    s->stpre.off = (uint32_t)o_stpre;
    s->stpost.off = (uint32_t)o_stpost;
    s->mask.off = (uint32_t)o_mask;
    s->wreset.off = (uint32_t)o_wreset;
    s->wwrite.off = (uint32_t)o_wwrite;
    s->nst.off = (uint32_t)o_nst;
    s->obs.off = (uint32_t)o_obs;
    s->rew.off = (uint32_t)o_rew;
    s->sperr.off = (uint32_t)o_sp;
    s->cold.off = (uint32_t)o_cold;
    s->cnew.off = (uint32_t)o_cnew;
    s->done.off = (uint32_t)o_done;
    s->adjs.off = (uint32_t)o_adjs;
    s->eidx.off = (uint32_t)o_eidx;
    s->eattr.off = (uint32_t)o_eattr;
    s->gidx.off = (uint32_t)o_gidx;
    s->gattr.off = (uint32_t)o_gattr;
  }
  return align_up(o, 16);
}

// contiguous shared -> global copy of n elements as 128-bit vectors (+ scalar tail) when both
// ends are 16-byte aligned (plan_tiles picks a tile size that makes every slab aligned)
template <typename T>
__device__ __forceinline__ void copy_out(T* __restrict__ dst, const T* __restrict__ src, int n, int tid, int nthr) {
  constexpr int VW = 16 / sizeof(T);
  const int nv = (((size_t)dst | (size_t)src) & 15) ? 0 : n / VW;  // unaligned slab: scalar stores
  const int4* s4 = (const int4*)src;
  int4* d4 = (int4*)dst;
  for (int q = tid; q < nv; q += nthr) d4[q] = s4[q];
  for (int q = nv * VW + tid; q < n; q += nthr) dst[q] = src[q];
}

// ---- per-agent pieces --------------------------------------------------------------

// MultiAgentBaseEnv._set_action (environment.py:336-475), index form of the one-hot
template <typename Real>
__host__ __device__ __forceinline__ void decode_action(const GmpeConfig& c, int a, Real& u0, Real& u1) {
  double x = 0.0, y = 0.0;
  if (c.dynamics == GMPE_DYN_DOUBLE_INTEGRATOR) {
    if (c.total_actions == 9) {  // environment.py:382-392
      const double d = 0.71;
      switch (a) {
        case 1: x = -1; break;
        case 2: x = -d; y = -d; break;
        case 3: y = -1; break;
        case 4: x = d; y = -d; break;
        case 5: x = 1; break;
        case 6: x = d; y = d; break;
        case 7: y = 1; break;
        case 8: x = -d; y = d; break;
        default: break;
      }
    } else {  // environment.py:404-406
      x = (double)((a == 1) - (a == 2));
      y = (double)((a == 3) - (a == 4));
    }
  } else {  // environment.py:424-449; np.linspace(start, stop, 5)[k] = k*step + start, last = stop
    int ar = a / 5, ac = a - ar * 5;
    double ws = (2.0 * c.ang_rate_max) / 4.0, as = (c.accel_max - c.accel_min) / 4.0;
    x = ar == 4 ? c.ang_rate_max : (double)ar * ws + (-c.ang_rate_max);
    y = ac == 4 ? c.accel_max : (double)ac * as + c.accel_min;
  }
  u0 = Real(x * c.sensitivity);  // environment.py:460-463
  u1 = Real(y * c.sensitivity);
}

// Closed form of solve_ivp(RK45) over [0, dt] (core.py:125-139, 207-218, 300-316),
// then the speed clamp and the travelled-distance counter.
template <typename Real>
__device__ __forceinline__ void integrate_kinematic(const GmpeConfig& c, typename RT<Real>::V4& s, Real u0, Real u1, Real& p_dist) {
  const Real dt = Real(c.dt);
  if (c.dynamics == GMPE_DYN_DOUBLE_INTEGRATOR) {
    s.x += s.z * dt + Real(0.5) * u0 * dt * dt;
    s.y += s.w * dt + Real(0.5) * u1 * dt * dt;
    s.z += u0 * dt;
    s.w += u1 * dt;
    Real speed = sqrt_(s.z * s.z + s.w * s.w);
    const Real vmax = Real(c.v_max);
    if (speed > vmax) {  // core.py:213-215
      s.z = vmax * s.z / speed;
      s.w = vmax * s.w / speed;
      speed = sqrt_(s.z * s.z + s.w * s.w);
    }
    p_dist += speed * dt;
    return;
  }
  Real th = s.z, v = s.w;
  Real th1 = th + u0 * dt, v1 = v + u1 * dt;
  Real s0, c0;
  sincos_(th, &s0, &c0);
  if (u0 != Real(0)) {
    Real s1, c1;
    sincos_(th1, &s1, &c1);
    Real iw = Real(1) / u0;
    s.x += (v1 * s1 - v * s0) * iw + u1 * (c1 - c0) * iw * iw;
    s.y += -(v1 * c1 - v * c0) * iw + u1 * (s1 - s0) * iw * iw;
  } else {
    Real d = v * dt + Real(0.5) * u1 * dt * dt;
    s.x += c0 * d;
    s.y += s0 * d;
  }
  v1 = clip(v1, Real(c.v_min), Real(c.v_max));  // core.py:132-135, 309-312
  s.z = wrap_heading(th1);
  s.w = v1;
  p_dist += v1 * dt;
}

// get_agent_phase (july:683-733) with _tube_coords / _in_tube_rect / _in_entrance_gate
// (july:621-637).  pp is agent.previous_phase and may be promoted to 2 (july:727).
template <typename Real>
__device__ __forceinline__ int tube_phase(const Real* wf, Real px, Real py, int& pp) {
  const Real eps = Real(0.05);
  Real rx = trunc32(px) - wf[GMPE_WF_ENT_X], ry = trunc32(py) - wf[GMPE_WF_ENT_Y];
  Real s = rx * wf[GMPE_WF_E_X] + ry * wf[GMPE_WF_E_Y];
  Real y = rx * wf[GMPE_WF_N_X] + ry * wf[GMPE_WF_N_Y];
  Real L = wf[GMPE_WF_L], hw = wf[GMPE_WF_HALF_W];
  bool lat = abs_(y) <= hw + eps;
  bool in_tube = (-eps <= s) && (s <= L + eps) && lat;
  Real tdx = wf[GMPE_WF_EXIT_X] - wf[GMPE_WF_ENT_X], tdy = wf[GMPE_WF_EXIT_Y] - wf[GMPE_WF_ENT_Y];
  Real tn = norm2(tdx, tdy);
  tdx = div_(tdx, tn); tdy = div_(tdy, tn);
  bool passed = ((px - wf[GMPE_WF_EXIT_X]) * tdx + (py - wf[GMPE_WF_EXIT_Y]) * tdy) > Real(0);
  Real gf = Real(0.08) * L, gb = Real(0.02) * L;
  bool gate = (-gb - eps <= s) && (s <= gf + eps) && lat;
  if (!in_tube && !passed) return 0;
  if (in_tube) return pp == 0 ? (gate ? 1 : 0) : 1;
  if (pp == 1) { pp = 2; return 2; }  // here passed is true
  if (pp == 2) return 2;
  return 0;
}

// ---- three_phase_graph.py ("tp"): the rotation-invariant scenario family -------------------
// _tube_coords (tp:592-598): float32 position, float32 normal
template <typename Real>
__device__ __forceinline__ void tp_tube_coords(const Real* wf, Real px, Real py, Real& s, Real& y) {
  const Real rx = trunc32(px) - wf[GMPE_WF_ENT_X], ry = trunc32(py) - wf[GMPE_WF_ENT_Y];
  s = rx * wf[GMPE_WF_E_X] + ry * wf[GMPE_WF_E_Y];
  y = rx * wf[GMPE_WF_N_X] + ry * wf[GMPE_WF_N_Y];
}
template <typename Real>
__device__ __forceinline__ bool tp_in_tube_rect(Real s, Real y, Real L, Real hw) {  // tp:600-601
  return (Real(-0.05) <= s) && (s <= L + Real(0.05)) && (abs_(y) <= hw + Real(0.05));
}
template <typename Real>
__device__ __forceinline__ bool tp_in_entrance_gate(Real s, Real y, Real L, Real hw) {  // tp:603-607, 584-585
  return (-(Real(0.02) * L) - Real(0.05) <= s) && (s <= Real(0.08) * L + Real(0.05)) && (abs_(y) <= hw + Real(0.05));
}
template <typename Real>
__device__ __forceinline__ bool tp_in_exit_gate(Real s, Real y, Real L, Real hw, Real back = Real(0.02)) {  // tp:618-621, 586-587
  return (L - back * L - Real(0.05) <= s) && (s <= L + Real(0.08) * L + Real(0.05)) && (abs_(y) <= hw + Real(0.05));
}
template <typename Real>
__device__ __forceinline__ Real tp_entrance_gate_distance(Real s, Real y, Real hw) {  // tp:610-615
  return hypot_(abs_(s), y - clip(y, -hw, hw));
}
template <typename Real>
__device__ __forceinline__ Real tp_exit_gate_distance(Real s, Real y, Real L, Real hw) {  // tp:624-633
  const Real ds = L - s > Real(0) ? L - s : Real(0);
  return hypot_(ds, y - clip(y, -hw, hw));
}
// get_agent_phase (tp:639-705); previous_phase / phase_reached are only read
// two_phase_graph.py ("tp2", TWO = true) lacks the "previous_phase == 2 and valid_exit" branch
// (tp:681-683 vs tp2:676-681); where tp2 falls off the end and returns None (an agent beyond the
// exit plane that never entered: float(None) raises in the reference) this build returns 0.
// VARIANT: 0 three_phase_graph.py, 1 two_phase_graph.py, 2 nav_graph_metered_single_corridor_rot_inv.py
// ("sc": inside the tube the phase stays 1, sc:693-695; exit gate reaches 0.05 L back, sc:621).
template <typename Real, int VARIANT>
__device__ __forceinline__ int tp_phase(Real s, Real y, Real L, Real hw, int pp, int pr) {
  constexpr bool TWO = VARIANT == 1, SC = VARIANT == 2;
  const bool in_tube = tp_in_tube_rect(s, y, L, hw), passed = s > L;
  const bool valid_exit = tp_in_exit_gate(s, y, L, hw, Real(SC ? 0.05 : 0.02));
  if (!in_tube && !passed) return 0;
  if (in_tube) {
    if (pp == 0) return tp_in_entrance_gate(s, y, L, hw) ? 1 : 0;
    if (SC) return 1;
    if (TWO) return (pp == 1 && valid_exit) ? 2 : 1;
    return valid_exit ? 2 : 1;  // pp is 1 or 2 here
  }
  if (pr >= 1) {  // passed the exit plane
    if (pp == 1 && valid_exit) return 2;
    if (pp == 2) return 2;
  }
  return 0;
}
// get_rotated_position_from_relative (tp:90-97) with the ego heading as a unit vector (hx, hy)
template <typename Real>
__device__ __forceinline__ void rot_into(Real hx, Real hy, Real vx, Real vy, Real& ox, Real& oy) {
  ox = hx * vx + hy * vy;
  oy = -hy * vx + hx * vy;
}

// is_obstacle_collision (july:864-890): obstacles within 2*(size_o + size), walls as a
// 1.5*size band around the segment.
template <typename Real>
__device__ __forceinline__ bool obstacle_collision(const GmpeConfig& c, const typename RT<Real>::V2* pos_w, const Real* wl_w,
                                                   int N, int L, int O, int W, Real px, Real py) {
  const Real size = Real(c.entity_size);
  const Real dmin = Real(2.0) * (size + size);
  for (int o = 0; o < O; ++o) {
    typename RT<Real>::V2 q = pos_w[N + L + o];
    if (norm2(q.x - px, q.y - py) < dmin) return true;
  }
  const Real band = Real(1.5) * size;
  for (int q = 0; q < W; ++q) {
    const Real* w = wl_w + q * GMPE_WALL_F;
    Real pa = w[0] == Real(0) ? py : px;  // coordinate along the wall's axis position
    Real pl = w[0] == Real(0) ? px : py;
    if (w[1] - band <= pa && pa <= w[1] + band && w[2] - band <= pl && pl <= w[3] + band) return true;
  }
  return false;
}

// get_wall_collision_force (core.py:909-964)
template <typename Real>
__device__ __forceinline__ void wall_force(const GmpeConfig& c, const Real* w, Real px, Real py, Real& fx, Real& fy) {
  const Real size = Real(c.entity_size);
  bool horiz = w[0] == Real(0);
  Real ppar = horiz ? px : py, pperp = horiz ? py : px;
  if (ppar < w[2] - size || ppar > w[3] + size) return;
  Real st = Real(0), ct = Real(1), dist_min;
  if (ppar < w[2] || ppar > w[3]) {
    Real past = ppar < w[2] ? ppar - w[2] : ppar - w[3];
    Real theta = asin_(past / size);
    sincos_(theta, &st, &ct);
    dist_min = ct * size + Real(0.5) * Real(c.wall_width);
  } else {
    dist_min = size + Real(0.5) * Real(c.wall_width);
  }
  Real delta = pperp - w[1];
  Real dist = abs_(delta);
  bool classic = (c.flags & GMPE_FLAG_CLASSIC_CONTACT) != 0;
  Real k = Real(classic ? c.contact_margin : c.wall_contact_margin);
  Real cf = Real(classic ? c.contact_force : c.wall_contact_force);
  Real pen = softplus(div_(-(dist - dist_min), k)) * k;
  Real fm = div_(cf * delta, dist) * pen;
  Real fperp = ct * fm, fpar = st * abs_(fm);
  if (horiz) { fx += fpar; fy += fperp; } else { fx += fperp; fy += fpar; }
}

// Scenario.observation (july:1337-1463) into the staging row o[19]
template <typename Real>
__device__ __forceinline__ void july_obs(const typename RT<Real>::V2* pos_w, const Real* wf, int N, int i,
                                         Real vx, Real vy, int phase, Real* o) {
  typename RT<Real>::V2 p = pos_w[i], g = pos_w[N + i];  // goal_match_index = arange (july:356)
  Real d1 = Real(0), d2 = Real(0);
  typename RT<Real>::V2 r1 = RT<Real>::mk2(0, 0), r2 = RT<Real>::mk2(0, 0);
  int cnt = 0;
  for (int k = 0; k < N; ++k) {  // two nearest others, stable on ties (july:1399-1415)
    if (k == i) continue;
    typename RT<Real>::V2 q = pos_w[k];
    Real rx = q.x - p.x, ry = q.y - p.y, d = norm2(rx, ry);
    if (cnt == 0 || d < d1) { d2 = d1; r2 = r1; d1 = d; r1 = RT<Real>::mk2(rx, ry); }
    else if (cnt == 1 || d < d2) { d2 = d; r2 = RT<Real>::mk2(rx, ry); }
    ++cnt;
  }
  Real gx = g.x - p.x, gy = g.y - p.y;
  o[0] = p.x; o[1] = p.y; o[2] = vx; o[3] = vy;
  o[4] = gx; o[5] = gy; o[6] = Real(0); o[7] = gx; o[8] = gy;
  o[9] = r1.x; o[10] = r1.y; o[11] = r2.x; o[12] = r2.y;
  o[13] = wf[GMPE_WF_ENT_X] - p.x; o[14] = wf[GMPE_WF_ENT_Y] - p.y;
  o[15] = wf[GMPE_WF_EXIT_X] - p.x; o[16] = wf[GMPE_WF_EXIT_Y] - p.y;
  o[17] = wf[GMPE_WF_WIDTH];
  o[18] = Real(phase);
}

// ---- reset (thread 0 of a world): random_scenario + setup_tube_params (july:440-613),
// set_landmarks_in_point (utils.py:165-193); nav placement is defined-by-build
template <typename Real>
__device__ __noinline__ void place_world(const KParams& p, const Smem<Real>& sm, int lw, int64_t b) {
  const GmpeConfig& c = p.c;
  const int N = c.num_agents, L = c.num_landmarks, O = c.num_obstacles, W = c.num_walls, E = p.E;
  typename RT<Real>::V2* pos_w = sm.pos + lw * E;
  typename RT<Real>::V4* nst = sm.nst + lw * N;
  Real* wf = sm.wf + lw * GMPE_WORLD_F;
  Real* wl = sm.wl + lw * W * GMPE_WALL_F;
  const DrawSrc ds{c.seed, c.world_offset + b, (uint32_t)sm.wep[lw],
                   p.reset_draws ? p.reset_draws + (size_t)b * p.reset_draws_n : nullptr, (uint32_t)p.reset_draws_n};
  uint32_t idx = 0;
  bool overflow = false;
  const double ws = c.world_size;
  const bool heading = c.dynamics != GMPE_DYN_DOUBLE_INTEGRATOR;
  const Real sep = Real(c.separation_dist);
  if (c.scenario == GMPE_SCEN_JULY_TUBE) {
    (void)ds.u01(idx++);  // wall_length draw (july:368), value unused
    double width = fmax(3 * c.entity_size * 2.5, ws * 0.15);
    double angle = ds.uniform(idx++, -1.5707963267948966, 1.5707963267948966);
    double tl = ws * 0.8, ca, sa;
    sincos(angle, &sa, &ca);
    double ent0 = sa * (tl / 4), ent1 = ca * (tl / 4), ext0 = sa * (-tl / 4), ext1 = ca * (-tl / 4);
    double Lh = sqrt((ext0 - ent0) * (ext0 - ent0) + (ext1 - ent1) * (ext1 - ent1)) + 1e-9;
    double e0 = (ext0 - ent0) / Lh, e1 = (ext1 - ent1) / Lh;
    wf[GMPE_WF_ENT_X] = Real(ent0); wf[GMPE_WF_ENT_Y] = Real(ent1);
    wf[GMPE_WF_EXIT_X] = Real(ext0); wf[GMPE_WF_EXIT_Y] = Real(ext1);
    wf[GMPE_WF_E_X] = Real(e0); wf[GMPE_WF_E_Y] = Real(e1);
    wf[GMPE_WF_N_X] = Real((float)(-e1)); wf[GMPE_WF_N_Y] = Real((float)e0);
    wf[GMPE_WF_L] = Real(Lh); wf[GMPE_WF_HALF_W] = Real(width * 0.5);
    wf[GMPE_WF_WIDTH] = Real(width); wf[GMPE_WF_ANGLE] = Real(angle);
    int k = 0, tries = 0;
    while (k < N) {  // july:452-486
      double jx = 0.2 * ds.uniform(idx, -ws, ws);
      double jy = 0.2 * ds.uniform(idx + 1, -ws, ws);
      idx += 2;
      double d = (ws + k) / 5;
      Real px = Real(ent0 + d * sa + jx), py = Real(ent1 + d * ca + jy);
      bool bad = obstacle_collision<Real>(c, pos_w, wl, N, L, O, W, px, py);
      for (int a = 0; a < k && !bad; ++a) bad = norm2(nst[a].x - px, nst[a].y - py) < sep;
      if (bad && ++tries < GMPE_PLACE_MAX_ATTEMPTS) continue;
      overflow = overflow || bad;  // bounded rejection: keep the last draw (gmpe_placement_overflows)
      tries = 0;
      Real th = Real(0), v = Real(0);
      if (heading) { th = wrap_heading(Real(ds.uniform(idx++, 0.0, 6.283185307179586))); v = Real(c.v_min); }
      nst[k] = RT<Real>::mk4(px, py, th, v);
      ++k;
    }
    for (int l = 0; l < L; ++l) {
      double lx, ly;
      formation_landmark(c, l, L, ext0, ext1, sa, ca, lx, ly);
      pos_w[N + l] = RT<Real>::mk2(Real(lx), Real(ly));
    }
  } else {
    const double h = ws / 2;
    for (int q = 0; q < W; ++q) {  // boundary box: H top, H bottom, V right, V left
      Real* w = wl + q * GMPE_WALL_F;
      w[0] = (q & 2) ? Real(1) : Real(0);
      w[1] = Real((q & 1) ? -h : h);
      w[2] = Real(-h); w[3] = Real(h);
      Real mid = Real(0.5) * (w[2] + w[3]);
      pos_w[N + L + O + q] = w[0] == Real(0) ? RT<Real>::mk2(mid, w[1]) : RT<Real>::mk2(w[1], mid);
    }
    for (int o = 0; o < O; ++o) {
      pos_w[N + L + o] = RT<Real>::mk2(Real(0.9 * ds.uniform(idx, -h, h)), Real(0.9 * ds.uniform(idx + 1, -h, h)));
      idx += 2;
    }
    int k = 0, tries = 0;
    while (k < N) {
      Real px = Real(0.9 * ds.uniform(idx, -h, h)), py = Real(0.9 * ds.uniform(idx + 1, -h, h));
      idx += 2;
      bool bad = obstacle_collision<Real>(c, pos_w, wl, N, L, O, W, px, py);
      for (int a = 0; a < k && !bad; ++a) bad = norm2(nst[a].x - px, nst[a].y - py) < sep;
      if (bad && ++tries < GMPE_PLACE_MAX_ATTEMPTS) continue;
      overflow = overflow || bad;
      tries = 0;
      Real th = Real(0), v = Real(0);
      if (heading) { th = wrap_heading(Real(ds.uniform(idx++, 0.0, 6.283185307179586))); v = Real(c.v_min); }
      nst[k] = RT<Real>::mk4(px, py, th, v);
      ++k;
    }
    const Real lmin = Real(1.2 * (2 * c.entity_size));  // is_landmark_collision july:916-925
    k = 0; tries = 0;
    while (k < L) {
      Real px = Real(0.9 * ds.uniform(idx, -h, h)), py = Real(0.9 * ds.uniform(idx + 1, -h, h));
      idx += 2;
      bool bad = obstacle_collision<Real>(c, pos_w, wl, N, L, O, W, px, py);
      for (int a = 0; a < k && !bad; ++a) bad = norm2(pos_w[N + a].x - px, pos_w[N + a].y - py) < lmin;
      if (bad && ++tries < GMPE_PLACE_MAX_ATTEMPTS) continue;
      overflow = overflow || bad;
      tries = 0;
      pos_w[N + k] = RT<Real>::mk2(px, py);
      ++k;
    }
  }
  if (overflow && p.overflow_cnt) atomicAdd(p.overflow_cnt, 1ull);
}

// -1 / 0 padding of one world's fixed-capacity edge buffers behind its last edge, by one warp:
// 128-bit stores over the aligned middle, scalars at the ragged ends.
template <typename T>
__device__ __forceinline__ void fill_span(T* p, T v, int lo, int hi, int lane) {
  constexpr int VW = 16 / (int)sizeof(T);
  int a = lo;
  while (a < hi && (reinterpret_cast<size_t>(p + a) & 15)) ++a;  // first 16-byte aligned element
  if (lo + lane < a) p[lo + lane] = v;                            // head: fewer than VW elements
  const int nv = (hi - a) / VW;
  if (nv > 0) {
    T vv[VW];
#pragma unroll
    for (int q = 0; q < VW; ++q) vv[q] = v;
    const int4 v4 = *reinterpret_cast<const int4*>(vv);
    int4* d4 = reinterpret_cast<int4*>(p + a);
    for (int q = lane; q < nv; q += 32) d4[q] = v4;
  }
  for (int q = a + (nv > 0 ? nv : 0) * VW + lane; q < hi; q += 32) p[q] = v;
}
template <typename Real>
__device__ __forceinline__ void pad_edges(int32_t* ei, Real* ea, int n, int cap, int lane) {
  fill_span<int32_t>(ei, -1, n, cap, lane);
  fill_span<int32_t>(ei + cap, -1, n, cap, lane);
  fill_span<Real>(ea, Real(0), n, cap, lane);
}

// ---- the kernel ----------------------------------------------------------------------
// Launch bounds picked on a B200: 128 threads with >= 5 resident CTAs per SM (<= 96 registers).
// Round 1 (profiles/r1_sweep.txt) ran 7 CTAs at 72 registers; with the kernel as it is now the spills
// at 72 cost more than the warps bring (profiles/r2_sweep.txt: N = 16: 247 -> 235 us, N = 64:
// 437 -> 420 us with 5; 4 and 3 are slower again, 6 and 8 as slow as 7).
#ifndef GMPE_TILE_NODE_ROWWISE_MIN_E  // node rows: warp per (world, ego) from this many entities on
#define GMPE_TILE_NODE_ROWWISE_MIN_E 64
#endif
#ifndef GMPE_XSKIP  // experiments only (tools/ab.py): bit0 edge stores, bit1 adj stores, bit2 node rows, bit3 padding
#define GMPE_XSKIP 0
#endif
#ifndef GMPE_EDGE_UNROLL
#define GMPE_EDGE_UNROLL 2
#endif
constexpr int kEdgeUnroll = GMPE_EDGE_UNROLL;
#ifndef GMPE_EDGE_UNROLL_MID  // compile-time shapes of up to 64 entities (the 16/16/8/4 instance): 236 -> 226 us with 8, N = 64 is best with 2
#define GMPE_EDGE_UNROLL_MID 8
#endif
constexpr int kEdgeUnrollMid = GMPE_EDGE_UNROLL_MID;
#ifndef GMPE_EDGE_LINEAR  // 1: linear walk of the E*E matrix for compile-time E that is not a multiple of 32 (see the edge emission)
#define GMPE_EDGE_LINEAR 1
#endif
#ifndef GMPE_EDGE_LINEAR_UNROLL
#define GMPE_EDGE_LINEAR_UNROLL 8
#endif
constexpr int kEdgeLinearUnroll = GMPE_EDGE_LINEAR_UNROLL;
#ifndef GMPE_FORCE_UNROLL  // pair loop of the force physics: independent MUFU chains overlap (N = 16: 226.3 -> 223.2 us with 8)
#define GMPE_FORCE_UNROLL 8
#endif
constexpr int kForceUnroll = GMPE_FORCE_UNROLL;
#ifndef GMPE_NODE_UNROLL
#define GMPE_NODE_UNROLL 1
#endif
constexpr int kNodeUnroll = GMPE_NODE_UNROLL;
#ifndef GMPE_LB_THREADS
#define GMPE_LB_THREADS 128
#endif
#ifndef GMPE_LB_MINBLOCKS
#define GMPE_LB_MINBLOCKS 5
#endif
// NC / LC / OC / WC > 0: entity counts known at compile time (instantiated for the BASELINE
// shapes 16/16/8/4 and 64/64/0/0 in fp32); NC == 0: every count is a launch-time value.
// SC: scenario known at compile time (GMPE_SCEN_*; the navigation instance also fixes the
// force-physics double-integrator state), -1 = launch-time value.
template <typename Real, int NC = 0, int LC = 0, int OC = 0, int WC = 0, int SC = -1>
__global__ void __launch_bounds__(GMPE_LB_THREADS, GMPE_LB_MINBLOCKS) gmpe_step_kernel(const __grid_constant__ KParams p) {
  using V2 = typename RT<Real>::V2;
  using V4 = typename RT<Real>::V4;
  const GmpeConfig& c = p.c;
  const int N = NC ? NC : c.num_agents, L = NC ? LC : c.num_landmarks, O = NC ? OC : c.num_obstacles, W = NC ? WC : c.num_walls;
  const int E = NC ? NC + LC + OC + WC : p.E, F = p.F;
  constexpr int EC = NC ? NC + LC + OC + WC : 0;
  const CDiv<NC> dN{p.dN};
  const CDiv<EC> dE{p.dE};
  const CDiv<EC * EC> dEE{p.dEE};
  const CDiv<LC> dL{p.dL};
  const CDiv<OC> dO{p.dO};
  const CDiv<WC> dW{p.dW};
  const int B = c.num_worlds, WPB = p.wpb;
  const bool july = SC < 0 ? c.scenario == GMPE_SCEN_JULY_TUBE : SC == GMPE_SCEN_JULY_TUBE;
  const bool heading = SC == GMPE_SCEN_NAV_GRAPH ? false : c.dynamics != GMPE_DYN_DOUBLE_INTEGRATOR;
  const bool do_info = (c.flags & GMPE_FLAG_INFO) != 0;
  const int tid = threadIdx.x, nthr = blockDim.x;
  Smem<Real> sm;
  smem_layout<Real>(WPB, N, E, W, F, p.stage_graph, &sm);

  Real* g_as = (Real*)p.st.agent_state;
  Real* g_aux = (Real*)p.st.agent_aux;
  uint32_t* g_fl = (uint32_t*)p.st.agent_flags;
  int32_t* g_cnt = p.st.agent_cnt;
  Real* g_wf = (Real*)p.st.world_f;
  int32_t* g_wi = p.st.world_i;
  Real* g_lm = (Real*)p.st.landmarks;
  Real* g_ob = (Real*)p.st.obstacles;
  Real* g_wl = (Real*)p.st.walls;

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int w0 = tile * WPB;
    const int nw = min(WPB, B - w0);
    // ---------------- cooperative, coalesced loads of world-level state -------------
    for (int q = tid; q < nw * GMPE_WORLD_F; q += nthr) sm.wf[q] = g_wf[(size_t)w0 * GMPE_WORLD_F + q];
    for (int q = tid; q < nw * L; q += nthr) {
      uint32_t lw, l; dL.divmod(q, lw, l);
      sm.pos[lw * E + N + l] = ((const V2*)g_lm)[(size_t)w0 * L + q];
    }
    for (int q = tid; q < nw * O; q += nthr) {
      uint32_t lw, o; dO.divmod(q, lw, o);
      sm.pos[lw * E + N + L + o] = ((const V2*)g_ob)[(size_t)w0 * O + q];
    }
    for (int q = tid; q < nw * W; q += nthr) {
      uint32_t lw, k; dW.divmod(q, lw, k);
      const Real* src = g_wl + ((size_t)w0 * W + q) * GMPE_WALL_F;
      Real* dst = sm.wl + q * GMPE_WALL_F;
      Real o0 = src[0], a = src[1], e0 = src[2], e1 = src[3];
      dst[0] = o0; dst[1] = a; dst[2] = e0; dst[3] = e1;
      Real mid = Real(0.5) * (e0 + e1);  // wall "position" = segment midpoint
      sm.pos[lw * E + N + L + O + k] = o0 == Real(0) ? RT<Real>::mk2(mid, a) : RT<Real>::mk2(a, mid);
    }
    for (int q = tid; q < nw; q += nthr) {
      int32_t stp = g_wi[(size_t)(w0 + q) * GMPE_WORLD_I + GMPE_WI_STEP];
      sm.wstep[q] = p.mode == MODE_STEP ? stp + 1 : stp;  // environment.py:1024
      sm.wep[q] = g_wi[(size_t)(w0 + q) * GMPE_WORLD_I + GMPE_WI_EPISODE];
      uint8_t m = p.mode == MODE_STEP ? 1 : (p.env_mask ? (p.env_mask[w0 + q] != 0) : 1);
      sm.wwrite[q] = m;
      sm.wreset[q] = p.mode == MODE_STEP ? 0 : m;
    }

    // ---------------- phase A: thread = (world, agent) -------------------------------
    const bool active = tid < nw * N;
    uint32_t lw = 0, i = 0;
    if (active) dN.divmod(tid, lw, i);
    const int64_t b = w0 + (int64_t)lw;
    const size_t ga = (size_t)w0 * N + tid;  // global agent slot
    V4 s = RT<Real>::mk4(0, 0, 0, 0);
    V2 aux = RT<Real>::mk2(0, 0);
    uint32_t flw = 0;
    Real u0 = Real(0), u1 = Real(0);
    int st_pre = 0, pp = 0, phase_reached = 0, goal_reached = 0;
    if (active) {
      s = ((const V4*)g_as)[ga];
      aux = ((const V2*)g_aux)[ga];
      flw = g_fl[ga];
      st_pre = flw & 0xff; pp = (flw >> 8) & 0xff; phase_reached = (flw >> 16) & 0xff; goal_reached = (flw >> 24) & 0xff;
    }
    V2 vpre = RT<Real>::mk2(0, 0), vpost = RT<Real>::mk2(0, 0);
    Real hx = Real(1), hy = Real(0);  // heading_vec of the PRE-reward state (july:1130-1131)
    int st_post = st_pre, ph = 0, pp_eff = pp;
    bool reach = false;
    Real d_goal = Real(0), rew = Real(0);
    int done = 0;
    const Real* wfw = sm.wf + lw * GMPE_WORLD_F;
    V2* pos_w = sm.pos + lw * E;
    const Real* wl_w = sm.wl + lw * W * GMPE_WALL_F;

    if (p.mode == MODE_STEP) {
      if (active) decode_action<Real>(c, p.act[ga], u0, u1);
      if (c.physics == GMPE_PHYS_FORCE) {
        // classic step order (onpolicy/envs/mpe/core.py:183-202) with the force functions of
        // multiagent/core.py:766-984: positions first, then pairwise penetration forces
        if (active) { pos_w[i] = RT<Real>::mk2(s.x, s.y); sm.stpre[tid] = (uint8_t)st_pre; }
        __syncthreads();
        if (active) {
          const bool classic = (c.flags & GMPE_FLAG_CLASSIC_CONTACT) != 0;
          const bool quirk = (c.flags & GMPE_FLAG_WALL_POINT_QUIRK) != 0;
          Real fx = u0, fy = u1;  // apply_action_force: mass 1, accel None, no noise
          const Real km = Real(c.contact_margin), cf = Real(c.contact_force);
          const int kend = quirk ? E : N + L + O;
#pragma unroll kForceUnroll
          for (int k = 0; k < kend; ++k) {
            if (k == (int)i || (k >= N && k < N + L)) continue;  // landmarks do not collide (july:298)
            if (k < N && !classic && st_pre) continue;           // core.py:899: no agent-agent force on a done entity
            V2 q = pos_w[k];
            Real dx = s.x - q.x, dy = s.y - q.y, dist = norm2(dx, dy);
            Real size_k = k >= N + L + O ? Real(c.wall_width) : Real(c.entity_size);
            Real dmin = classic ? Real(c.entity_size) + size_k : Real(c.separation_dist);
            Real pen = softplus(div_(-(dist - dmin), km)) * km;
            fx = div_(cf * dx, dist) * pen + fx;
            fy = div_(cf * dy, dist) * pen + fy;
          }
          for (int q = 0; q < W; ++q) wall_force<Real>(c, wl_w + q * GMPE_WALL_F, s.x, s.y, fx, fy);
          if (classic || !st_pre) {  // integrate_state core.py:827-845; done agents freeze
            const Real dt = Real(c.dt);
            s.z = s.z * Real(1 - c.damping) + fx * dt;
            s.w = s.w * Real(1 - c.damping) + fy * dt;
            if (c.max_speed > 0) {
              Real speed = sqrt_(s.z * s.z + s.w * s.w);
              if (speed > Real(c.max_speed)) { s.z = div_(s.z, speed) * Real(c.max_speed); s.w = div_(s.w, speed) * Real(c.max_speed); }
            }
            s.x += s.z * dt; s.y += s.w * dt;
            aux.x += norm2(s.z * dt, s.w * dt);
          }
        }
        __syncthreads();
      } else if (active && !st_pre) {  // core.py:819-826
        integrate_kinematic<Real>(c, s, u0, u1, aux.x);
      }
      if (active) {
        if (heading) { sincos_(s.z, &hy, &hx); vpre = RT<Real>::mk2(s.w * hx, s.w * hy); }
        else {
          vpre = RT<Real>::mk2(s.z, s.w);
          Real sp = norm2(s.z, s.w);  // DoubleIntegratorXYState.theta = arctan2(vy, vx), core.py:188-189
          if (sp > Real(0)) { hx = div_(s.z, sp); hy = div_(s.w, sp); }
        }
        pos_w[i] = RT<Real>::mk2(s.x, s.y);
        sm.velpre[tid] = vpre;
        sm.stpre[tid] = (uint8_t)st_pre;
      }
    }
    __syncthreads();  // world frame, static entities and post-step agent positions are in smem

    if (p.mode == MODE_STEP && active) {
      // ---- pass A: everything reward() derives from the agent's own state -----------
      V2 g = pos_w[N + i];
      d_goal = norm2(s.x - g.x, s.y - g.y);
      if (july) {
        ph = tube_phase<Real>(wfw, s.x, s.y, pp_eff);  // observation-time call may promote pp (july:727)
        reach = (ph == 2) && (phase_reached != 0) && (d_goal < Real(c.goal_dist_thresh)) && !st_pre;
      } else {
        reach = (d_goal < Real(c.goal_dist_thresh)) && (c.flags & GMPE_FLAG_USE_DONES) && !st_pre;
      }
      vpost = vpre;
      if (reach) {  // agent.status = True; state.reset_velocity() (july:1188-1190, core.py:324-333)
        st_post = 1;
        if (heading) {
          if (july) {
            Real th = p.theta ? ((const Real*)p.theta)[ga] : Real(NAN);
            if (isnan_(th))
              th = Real(philox_uniform(c.seed, c.world_offset + b, (uint32_t)sm.wep[lw],
                                       (1u << 30) + (uint32_t)sm.wstep[lw] * 64u + i, 0.0, 6.283185307179586));
            s.z = wrap_heading(th);
          }
          s.w = Real(c.v_min);
          Real sn, cs; sincos_(s.z, &sn, &cs);
          vpost = RT<Real>::mk2(s.w * cs, s.w * sn);
        } else { s.z = Real(0); s.w = Real(0); vpost = RT<Real>::mk2(0, 0); }
      }
      sm.stpost[tid] = (uint8_t)st_post;
      sm.velpost[tid] = vpost;
    }
    __syncthreads();

    Real sp_err = Real(0);
    int32_t cnt[GMPE_AGENT_CNT];
    if (p.mode == MODE_STEP && active) {
      // ---- pass B: terms that look at other agents (ordered semantics) ----
      const uint8_t* spre = sm.stpre + lw * N;
      const uint8_t* spost = sm.stpost + lw * N;
      const Real sep = Real(c.separation_dist);
      const Real coll = Real(c.collision_rew), goal = Real(c.goal_rew);
      int ncoll_rew = 0, ncoll_info = 0;
      int front = -1, back = -1;
      Real fproj = Real(0), bproj = Real(0), fdist = Real(0), bdist = Real(0);
      for (int k = 0; k < N; ++k) {
        if (k == (int)i) continue;
        V2 q = pos_w[k];
        Real rx = q.x - s.x, ry = q.y - s.y, d = norm2(rx, ry);
        int st_k = k < (int)i ? spost[k] : spre[k];
        bool close = d < sep;
        ncoll_rew += (close && !st_k && !st_pre);    // is_collision at reward(i) time (july:907-914)
        ncoll_info += (close && !st_k && !st_post);  // ... at info_callback(i) time (july:785)
        if (july) {  // front / back neighbours along the own heading (july:1136-1143)
          Real proj = rx * hx + ry * hy;
          if (proj > Real(0)) { if (front < 0 || proj < fproj) { front = k; fproj = proj; fdist = d; } }
          else { if (back < 0 || proj > bproj) { back = k; bproj = proj; bdist = d; } }
        }
      }
      bool obst = obstacle_collision<Real>(c, pos_w, wl_w, N, L, O, W, s.x, s.y);
      if (do_info) {
        const int32_t* gc = g_cnt + ga * GMPE_AGENT_CNT;
#pragma unroll
        for (int q = 0; q < GMPE_AGENT_CNT; ++q) cnt[q] = gc[q];
        sm.cold[tid * 3 + 0] = cnt[GMPE_AC_DIST_TRAV_I];
        sm.cold[tid * 3 + 1] = cnt[GMPE_AC_TIME_REQ_I];
        sm.cold[tid * 3 + 2] = cnt[GMPE_AC_SPACING_VIOL];
      } else {
#pragma unroll
        for (int q = 0; q < GMPE_AGENT_CNT; ++q) cnt[q] = 0;
      }
      int cur = ph;
      if (july) {
        rew -= Real(ncoll_rew) * (coll * Real(4));
        if (obst) rew -= coll * Real(3);
        Real tdx = wfw[GMPE_WF_EXIT_X] - wfw[GMPE_WF_ENT_X], tdy = wfw[GMPE_WF_EXIT_Y] - wfw[GMPE_WF_ENT_Y];
        Real tube_length = norm2(tdx, tdy);
        if (ph == 2 && ph > pp_eff + 1) rew -= goal * Real(3);  // july:1149
        Real ex = s.x - wfw[GMPE_WF_ENT_X], ey = s.y - wfw[GMPE_WF_ENT_Y];
        Real proj = ex * div_(tdx, tube_length) + ey * div_(tdy, tube_length);
        Real edist = norm2(ex - proj * tdx, ey - proj * tdy);  // un-normalised direction, july:1154
        if (ph == pp_eff + 1 && phase_reached == ph - 1) {     // july:1156-1160
          if (ph == 1 && Real(0) <= proj && proj < Real(0.1) * tube_length && edist < Real(0.2) * tube_length) rew += goal * Real(3);
          else if (ph == 2) rew += goal * Real(3);
        }
        if (ph == 0) {
          rew -= norm2(wfw[GMPE_WF_ENT_X] - s.x, wfw[GMPE_WF_ENT_Y] - s.y);
        } else if (ph == 1) {  // metering terms, july:1167-1181
          if (front >= 0) { Real diff = fdist - sep; sp_err += diff < Real(0) ? abs_(diff) : Real(0); }
          if (back >= 0) { Real diff = bdist - sep; sp_err += diff < Real(0) ? abs_(diff) : Real(0); }
          if (sp_err > Real(0)) cnt[GMPE_AC_SPACING_VIOL] += 1;
          rew -= sp_err * Real(c.formation_rew);
          rew -= norm2(wfw[GMPE_WF_EXIT_X] - s.x, wfw[GMPE_WF_EXIT_Y] - s.y);
          cnt[GMPE_AC_STEPS_IN_CORRIDOR] += 1;
        } else if (ph == 2 && phase_reached == 0) {
          cur = 0;
        } else {  // july:1185-1194
          if (d_goal < Real(c.goal_dist_thresh)) { if (reach) rew += goal * Real(5); }
          else rew -= d_goal;
        }
        if (phase_reached == 1 && cur == 0) cnt[GMPE_AC_CONFORMANCE] += 1;
        if (cur > phase_reached) phase_reached = cur;
        if (cur < pp_eff) rew -= coll * Real(3);
        if (cur < phase_reached) rew -= coll;
        pp = cur;
        rew = clip(rew, -Real(4) * coll, goal * Real(5));
        rew = clip(rew, Real(c.min_reward), Real(c.max_reward));
      } else {
        rew -= Real(ncoll_rew) * coll;
        if (obst) rew -= coll;
        if (d_goal < Real(c.goal_dist_thresh)) rew += goal; else rew -= d_goal;
      }
      done = (st_post || sm.wstep[lw] >= c.episode_length) ? 1 : 0;  // environment.py:264-271
      sm.done[tid] = (uint8_t)done;
      sm.rew[tid] = rew;
      sm.sperr[tid] = sp_err;
      // observation uses the PRE-reward velocity and the observation-time phase
      if (july) july_obs<Real>(pos_w, wfw, N, i, vpre.x, vpre.y, ph, sm.obs + (size_t)tid * F);
      else {
        Real* o = sm.obs + (size_t)tid * F;
        V2 g = pos_w[N + i];
        o[0] = vpre.x; o[1] = vpre.y; o[2] = s.x; o[3] = s.y; o[4] = g.x - s.x; o[5] = g.y - s.y;
      }
      if (do_info) {  // info_callback own-state part (july:741-787); int32 truncation as the int64 numpy arrays
        int nearest = 0;
        Real dmin = Real(0);
        for (int l = 0; l < L; ++l) {
          V2 q = pos_w[N + l];
          Real d = norm2(s.x - q.x, s.y - q.y);
          if (l == 0 || d < dmin) { dmin = d; nearest = l; }
        }
        int gr = goal_reached - 1;
        const Real thr = Real(c.goal_dist_thresh);
        const int32_t tnow = (int32_t)((double)sm.wstep[lw] * c.dt);
        const int32_t pd = (int32_t)aux.x, dl = (int32_t)dmin;
        if (dmin < thr && (nearest != gr && gr != -1)) { gr = nearest; cnt[GMPE_AC_DIST_LEFT_I] = dl; }
        if (dmin < thr && cnt[GMPE_AC_TIME_REQ_I] == -1) { cnt[GMPE_AC_TIME_REQ_I] = tnow; cnt[GMPE_AC_DIST_TRAV_I] = pd; cnt[GMPE_AC_DIST_LEFT_I] = dl; gr = nearest; }
        if (cnt[GMPE_AC_TIME_REQ_I] == -1) { cnt[GMPE_AC_DIST_TRAV_I] = pd; cnt[GMPE_AC_DIST_LEFT_I] = dl; }
        if (dmin > thr && cnt[GMPE_AC_TIME_REQ_I] != -1) { cnt[GMPE_AC_DIST_TRAV_I] = pd; cnt[GMPE_AC_TIME_REQ_I] = tnow; cnt[GMPE_AC_DIST_LEFT_I] = dl; }
        if (dmin < thr && nearest == gr) { cnt[GMPE_AC_DIST_LEFT_I] = dl; gr = nearest; }
        goal_reached = gr + 1;
        if (obst) cnt[GMPE_AC_OBST_COLL] += 1;
        cnt[GMPE_AC_AGENT_COLL] += ncoll_info;
        sm.cnew[tid * 3 + 0] = cnt[GMPE_AC_DIST_TRAV_I];
        sm.cnew[tid * 3 + 1] = cnt[GMPE_AC_TIME_REQ_I];
        sm.cnew[tid * 3 + 2] = cnt[GMPE_AC_SPACING_VIOL];
      }
    }
    __syncthreads();

    if (p.mode == MODE_STEP && active) {
      // ---- reward / done / info outputs, state write-back ------------------------------
      Real r_out = rew;
      if (c.flags & GMPE_FLAG_COLLABORATIVE) {  // environment.py:1056-1059
        Real sum = Real(0);
        for (int k = 0; k < N; ++k) sum += sm.rew[lw * N + k];
        r_out = sum;
      }
      ((Real*)p.out.reward)[ga] = r_out;
      if (p.out.indiv_reward) ((Real*)p.out.indiv_reward)[ga] = rew;
      p.out.done[ga] = (uint8_t)done;
      bool all_done = true;
      for (int k = 0; k < N; ++k) all_done = all_done && sm.done[lw * N + k];
      if (p.out.mask) p.out.mask[ga] = done ? 0.0f : 1.0f;  // GMPERunner.insert (graph_mpe_runner.py:397-405)
      if (p.out.active_mask) p.out.active_mask[ga] = (all_done || !done) ? 1.0f : 0.0f;
      if (do_info) {
        // world-level statistics are recomputed inside every agent's info_callback while
        // later agents still hold last step's values (july:789-793, 800-823)
        Real dm = Real(0), tm = Real(0), sv = Real(0), dsum = wfw[GMPE_WF_DELTA_SPACING];
        for (int k = 0; k < N; ++k) {
          const int32_t* cc = (k <= (int)i ? sm.cnew : sm.cold) + (lw * N + k) * 3;
          dm += Real(cc[0]); tm += Real(cc[1]); sv += Real(cc[2]);
          if (k <= (int)i) dsum += sm.sperr[lw * N + k];
        }
        dm = mean_div(dm, Real(N)); tm = mean_div(tm, Real(N));
        Real ds = Real(0), ts = Real(0);
        for (int k = 0; k < N; ++k) {
          const int32_t* cc = (k <= (int)i ? sm.cnew : sm.cold) + (lw * N + k) * 3;
          Real x = Real(cc[0]) - dm, y = Real(cc[1]) - tm;
          ds += x * x; ts += y * y;
        }
        ds = sqrt_(qdiv(ds, Real(N))); ts = sqrt_(qdiv(ts, Real(N)));
        if (p.out.info) {
          Real* info = (Real*)p.out.info + ga * GMPE_INFO_K;
          V4* i4 = (V4*)info;
          i4[0] = RT<Real>::mk4(Real(cnt[GMPE_AC_DIST_LEFT_I]), Real(cnt[GMPE_AC_TIME_REQ_I]), Real(cnt[GMPE_AC_AGENT_COLL]), Real(cnt[GMPE_AC_OBST_COLL]));
          i4[1] = RT<Real>::mk4(dm, ds, qdiv(dm, ds + Real(0.0001)), Real(cnt[GMPE_AC_DIST_TRAV_I]));
          i4[2] = RT<Real>::mk4(Real(cnt[GMPE_AC_TIME_REQ_I]), tm, ts, qdiv(tm, ts + Real(0.0001)));
          i4[3] = RT<Real>::mk4(qdiv(Real(cnt[GMPE_AC_CONFORMANCE]), Real(c.episode_length)),
                                qdiv(dsum, sv != Real(0) ? sv : Real(1)),
                                qdiv(Real(cnt[GMPE_AC_SPACING_VIOL]), cnt[GMPE_AC_STEPS_IN_CORRIDOR] != 0 ? Real(cnt[GMPE_AC_STEPS_IN_CORRIDOR]) : Real(1)),
                                aux.y);
        }
        if (i == (uint32_t)(N - 1)) g_wf[(size_t)b * GMPE_WORLD_F + GMPE_WF_DELTA_SPACING] = dsum;
        int4* gc = (int4*)(g_cnt + ga * GMPE_AGENT_CNT);
        gc[0] = make_int4(cnt[0], cnt[1], cnt[2], cnt[3]);
        gc[1] = make_int4(cnt[4], cnt[5], cnt[6], cnt[7]);
      }
      const bool will_reset = (c.flags & GMPE_FLAG_AUTO_RESET) && all_done;
      if (!will_reset) {
        ((V4*)g_as)[ga] = s;
        ((V2*)g_aux)[ga] = aux;
        g_fl[ga] = (uint32_t)st_post | ((uint32_t)pp << 8) | ((uint32_t)phase_reached << 16) | ((uint32_t)goal_reached << 24);
        if (i == 0) g_wi[(size_t)b * GMPE_WORLD_I + GMPE_WI_STEP] = sm.wstep[lw];
      } else if (i == 0) {
        sm.wreset[lw] = 1;
      }
    }
    __syncthreads();

    // ---------------- reset (auto-reset of finished worlds, or MODE_RESET) ------------------
    // reset_world (july:339-420): counters cleared, previous_phase survives
    const bool rs = active && sm.wreset[lw];
    if (rs && i == 0) {
      if (p.has_rst) {
        const Real* rwf = (const Real*)p.rst.world_f + (size_t)b * GMPE_WORLD_F;
        Real* wf = sm.wf + lw * GMPE_WORLD_F;
        for (int q = 0; q < 12; ++q) wf[q] = rwf[q];
        for (int l = 0; l < L; ++l) pos_w[N + l] = ((const V2*)p.rst.landmarks)[(size_t)b * L + l];
        for (int o = 0; o < O; ++o) pos_w[N + L + o] = ((const V2*)p.rst.obstacles)[(size_t)b * O + o];
        for (int q = 0; q < W; ++q) {
          const Real* src = (const Real*)p.rst.walls + ((size_t)b * W + q) * GMPE_WALL_F;
          Real* dst = sm.wl + (lw * W + q) * GMPE_WALL_F;
          for (int z = 0; z < GMPE_WALL_F; ++z) dst[z] = src[z];
          Real mid = Real(0.5) * (dst[2] + dst[3]);
          pos_w[N + L + O + q] = dst[0] == Real(0) ? RT<Real>::mk2(mid, dst[1]) : RT<Real>::mk2(dst[1], mid);
        }
      } else {
        // the out-of-line reset gets a COPY of the offset table: were `sm` itself to escape, it
        // would live in local memory and every hot loop would reload its offsets (LDL) per iteration
        const Smem<Real> sm_reset = sm;
        place_world<Real>(p, sm_reset, lw, b);
      }
      // persist the world-level arrays
      Real* wf = sm.wf + lw * GMPE_WORLD_F;
      wf[GMPE_WF_DELTA_SPACING] = Real(0);
      for (int q = 0; q < GMPE_WORLD_F; ++q) g_wf[(size_t)b * GMPE_WORLD_F + q] = wf[q];
      for (int l = 0; l < L; ++l) ((V2*)g_lm)[(size_t)b * L + l] = pos_w[N + l];
      for (int o = 0; o < O; ++o) ((V2*)g_ob)[(size_t)b * O + o] = pos_w[N + L + o];
      for (int q = 0; q < W * GMPE_WALL_F; ++q) g_wl[(size_t)b * W * GMPE_WALL_F + q] = sm.wl[lw * W * GMPE_WALL_F + q];
      g_wi[(size_t)b * GMPE_WORLD_I + GMPE_WI_STEP] = 0;
      g_wi[(size_t)b * GMPE_WORLD_I + GMPE_WI_EPISODE] = sm.wep[lw] + 1;
    }
    __syncthreads();
    if (rs) {
      if (p.has_rst) {
        s = ((const V4*)p.rst.agent_state)[ga];
        aux.y = ((const V2*)p.rst.agent_aux)[ga].y;
      } else {
        s = sm.nst[tid];
        V2 g = pos_w[N + i];
        aux.y = c.max_speed > 0 ? div_(norm2(s.x - g.x, s.y - g.y), Real(c.max_speed)) : Real(INFINITY);  // min_time july:941-951
      }
      aux.x = Real(0);
      st_post = 0; phase_reached = 0; goal_reached = 0;
      if (heading) { Real sn, cs; sincos_(s.z, &sn, &cs); vpre = RT<Real>::mk2(s.w * cs, s.w * sn); }
      else vpre = RT<Real>::mk2(s.z, s.w);
      pos_w[i] = RT<Real>::mk2(s.x, s.y);
      sm.velpre[tid] = vpre;
      sm.velpost[tid] = vpre;
      sm.stpost[tid] = 0;
    }
    __syncthreads();
    if (rs) {
      // observations of the fresh episode replace the terminal ones (env_wrappers.py:867-873)
      int ph0 = 0;
      if (july) {
        ph0 = tube_phase<Real>(wfw, s.x, s.y, pp);
        july_obs<Real>(pos_w, wfw, N, i, vpre.x, vpre.y, ph0, sm.obs + (size_t)tid * F);
      } else {
        Real* o = sm.obs + (size_t)tid * F;
        V2 g = pos_w[N + i];
        o[0] = vpre.x; o[1] = vpre.y; o[2] = s.x; o[3] = s.y; o[4] = g.x - s.x; o[5] = g.y - s.y;
      }
      ((V4*)g_as)[ga] = s;
      ((V2*)g_aux)[ga] = aux;
      g_fl[ga] = (uint32_t)0 | ((uint32_t)pp << 8);
      if (do_info || p.mode == MODE_RESET) {
        int4* gc = (int4*)(g_cnt + ga * GMPE_AGENT_CNT);
        gc[0] = make_int4(0, 0, 0, 0);
        gc[1] = make_int4(0, -1, -1, -1);
      }
    }
    __syncthreads();

    // ---------------- phase B: element-parallel output generation -------------------------
    // disconnected mask per entity (july:1627-1648): done agents, and landmark l once agent l
    // reached it (goal_tracker[l] == l; goal_match_index = arange)
    for (int q = tid; q < nw * E; q += nthr) {
      uint32_t w, k; dE.divmod(q, w, k);
      const uint8_t* stw = sm.stpost + w * N;
      sm.mask[q] = k < (uint32_t)N ? stw[k] : ((k < (uint32_t)(2 * N) && k < (uint32_t)(N + L)) ? stw[k - N] : (uint8_t)0);
    }
    // obs: staged rows -> contiguous 128-bit stores
    if (p.out.obs) {
      Real* g_obs = (Real*)p.out.obs + (size_t)w0 * N * F;
      const int total = nw * N * F;
      if (p.mode == MODE_STEP) {
        copy_out<Real>(g_obs, sm.obs, total, tid, nthr);
      } else {
        for (int q = tid; q < total; q += nthr) {
          uint32_t row = p.dF.div(q);
          uint32_t w = dN.div(row);
          if (sm.wwrite[w]) g_obs[q] = sm.obs[q];
        }
      }
    }
    // node_obs [lw][j][k][8]  (_get_entity_feat_relative, july:1694-1771)
    // Row-wise generation for large worlds (E >= 64): a warp owns one matrix row and its lanes
    // walk the entities, so the inner loops carry no index arithmetic.  (Measured on B200:
    // 951 -> 746 us at E = 128; for E = 44 sub-warp row groups of 16 lanes were slower than
    // the element-parallel form below, 461 vs 382 us, so smaller worlds keep that.)
    const bool rowwise = !p.stage_graph && E >= GMPE_TILE_NODE_ROWWISE_MIN_E;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    constexpr int LPR = 32, rpw = 1;
    const int sub = 0, sl = lane;
    if (p.out.node_obs && rowwise) {
      // consecutive lanes write consecutive 32-byte feature rows with one 256-bit store each
      Real* g_node = (Real*)p.out.node_obs + (size_t)w0 * N * E * 8;
      for (int row = warp * rpw + sub; row < nw * N; row += nwarps * rpw) {
        uint32_t w, j;
        dN.divmod(row, w, j);
        if (p.mode != MODE_STEP && !sm.wwrite[w]) continue;
        const V2* pw = sm.pos + w * E;
        const V2 pe = pw[j], ve = sm.velpost[w * N + j];
        Real* grow = g_node + (size_t)row * E * 8;
        for (int k = sl; k < E; k += LPR) {
          const V2 pk = pw[k];
          const Real rx = pk.x - pe.x, ry = pk.y - pe.y;
          if (k < N) {
            const V2 vk = (k <= (int)j ? sm.velpost : sm.velpre)[w * N + k];
            const V2 gk = pw[N + k];  // landmark_poses[entity.id] (july:1740)
            const Real f[8] = {vk.x - ve.x, vk.y - ve.y, rx, ry, gk.x - pe.x, gk.y - pe.y, Real(0), Real(0)};
            store_row8(grow + (size_t)k * 8, f);
          } else {
            const Real ty = k < N + L ? Real(1) : (k < N + L + O ? Real(2) : Real(3));
            const Real f[8] = {Real(0) - ve.x, Real(0) - ve.y, rx, ry, rx, ry, Real(1), ty};
            store_row8(grow + (size_t)k * 8, f);
          }
        }
      }
    } else if (p.out.node_obs) {
      // smaller worlds: element-parallel over (world, ego j, entity k), one 256-bit store each
      Real* g_node = (Real*)p.out.node_obs + (size_t)w0 * N * E * 8;
      const int total = nw * N * E;
#pragma unroll kNodeUnroll
      for (int q = tid; q < total; q += nthr) {
        uint32_t k, t2, j, w;
        dE.divmod(q, t2, k);
        dN.divmod(t2, w, j);
        if (p.mode != MODE_STEP && !sm.wwrite[w]) continue;
        const V2* pw = sm.pos + w * E;
        const V2 pe = pw[j], pk = pw[k], ve = sm.velpost[w * N + j];
        const Real rx = pk.x - pe.x, ry = pk.y - pe.y;
        if (k < (uint32_t)N) {
          const V2 vk = (k <= j ? sm.velpost : sm.velpre)[w * N + k];
          const V2 gk = pw[N + k];  // landmark_poses[entity.id] (july:1740)
          const Real f[8] = {vk.x - ve.x, vk.y - ve.y, rx, ry, gk.x - pe.x, gk.y - pe.y, Real(0), Real(0)};
          if (!(GMPE_XSKIP & 4)) store_row8(g_node + (size_t)q * 8, f);
        } else {
          const Real ty = k < (uint32_t)(N + L) ? Real(1) : (k < (uint32_t)(N + L + O) ? Real(2) : Real(3));
          const Real f[8] = {Real(0) - ve.x, Real(0) - ve.y, rx, ry, rx, ry, Real(1), ty};
          if (!(GMPE_XSKIP & 4)) store_row8(g_node + (size_t)q * 8, f);
        }
      }
    }
    __syncthreads();  // mask ready; phase-A scratch (obs staging, counters) is dead from here on

    const bool env_e = (c.flags & GMPE_FLAG_ENV_EDGES) && p.out.edge_index;
    const bool gnn_e = (c.flags & GMPE_FLAG_GNN_EDGES) && p.out.gnn_edge_index;
    const int cap = E * (E - 1), EE = E * E;
    const Real thr_env = Real(c.coordination_range);
    const float thr_gnn = (float)c.gnn_max_edge_dist;
    if (p.stage_graph && p.mode == MODE_STEP) {
      // ---- small worlds: the masked distance matrix of the tile is built ONCE in shared
      // memory; adj is a vector copy of it and the edge lists are compacted from it
      // (one thread per world walks its E*E entries in row-major order) into staged,
      // pre-padded buffers that leave the SM as contiguous 128-bit stores.
      for (int q = tid; q < nw * EE; q += nthr) {
        uint32_t w, rc, r, cc;
        dEE.divmod(q, w, rc);
        dE.divmod(rc, r, cc);
        const V2* pw = sm.pos + w * E;
        const uint8_t* mk = sm.mask + w * E;
        V2 a = pw[r], bb = pw[cc];
        sm.adjs[q] = (r == cc || mk[r] || mk[cc]) ? Real(0) : norm2(a.x - bb.x, a.y - bb.y);
      }
      if (env_e) {
        for (int q = tid; q < nw * 2 * cap; q += nthr) sm.eidx[q] = -1;
        for (int q = tid; q < nw * cap; q += nthr) sm.eattr[q] = Real(0);
      }
      if (gnn_e) {
        for (int q = tid; q < nw * 2 * cap; q += nthr) sm.gidx[q] = -1;
        for (int q = tid; q < nw * cap; q += nthr) sm.gattr[q] = Real(0);
      }
      __syncthreads();
      if (p.out.adj) copy_out<Real>((Real*)p.out.adj + (size_t)w0 * EE, sm.adjs, nw * EE, tid, nthr);
      if ((env_e || gnn_e) && tid < nw) {
        const int w = tid;
        const Real* aw = sm.adjs + w * EE;
        int32_t* ei = sm.eidx + w * 2 * cap; Real* ea = sm.eattr + w * cap;
        int32_t* gi = sm.gidx + w * 2 * cap; Real* ga_ = sm.gattr + w * cap;
        int n_env = 0, n_gnn = 0, r = 0, cc = 0;
        for (int q = 0; q < EE; ++q) {
          Real d = aw[q];
          if (env_e && d > Real(0) && d <= thr_env) {  // update_graph (july:1660-1665), row-major
            ei[n_env] = r; ei[cap + n_env] = cc; ea[n_env] = d; ++n_env;
          }
          if (gnn_e) {  // process_adj on the fp32 cast (gnn_new.py:339): strict <
            float df = (float)d;
            if (df > 0.0f && df < thr_gnn) { gi[n_gnn] = r; gi[cap + n_gnn] = cc; ga_[n_gnn] = Real(df); ++n_gnn; }
          }
          if (++cc == E) { cc = 0; ++r; }
        }
        if (env_e) p.out.num_edges[w0 + w] = n_env;
        if (gnn_e) p.out.gnn_num_edges[w0 + w] = n_gnn;
      }
      __syncthreads();
      if (env_e) {
        copy_out<int32_t>(p.out.edge_index + (size_t)w0 * 2 * cap, sm.eidx, nw * 2 * cap, tid, nthr);
        copy_out<Real>((Real*)p.out.edge_attr + (size_t)w0 * cap, sm.eattr, nw * cap, tid, nthr);
      }
      if (gnn_e) {
        copy_out<int32_t>(p.out.gnn_edge_index + (size_t)w0 * 2 * cap, sm.gidx, nw * 2 * cap, tid, nthr);
        copy_out<Real>((Real*)p.out.gnn_edge_attr + (size_t)w0 * cap, sm.gattr, nw * cap, tid, nthr);
      }
    } else {
      // ---- mid-size worlds (or reset-only launches of small ones): element-parallel, in place
      // adj: masked distance matrix, one copy per world (july:1625-1648)
      // (when an edge list is emitted as well, its pass below computes every distance anyway
      // and writes the adjacency row-major as it goes)
      if (p.out.adj && !(env_e || gnn_e)) {
        Real* g_adj = (Real*)p.out.adj + (size_t)w0 * EE;
        const int total = nw * EE;
        for (int q = tid; q < total; q += nthr) {
          uint32_t w, rc, r, cc;
          dEE.divmod(q, w, rc);
          dE.divmod(rc, r, cc);
          if (p.mode != MODE_STEP && !sm.wwrite[w]) continue;
          const V2* pw = sm.pos + w * E;
          const uint8_t* mk = sm.mask + w * E;
          V2 a = pw[r], bb = pw[cc];
          g_adj[q] = (r == cc || mk[r] || mk[cc]) ? Real(0) : norm2(a.x - bb.x, a.y - bb.y);
        }
      }
      // edges + adjacency: a warp walks the rows of one world.  Every lane keeps its (up to two)
      // columns -- position and mask of entities `lane` and `lane + 32` -- in registers for the
      // whole world, the row entity is a shared-memory broadcast, so a row costs two distance
      // evaluations per lane and no index arithmetic; each pair distance is evaluated once and
      // feeds the adjacency entry and the stream compaction (ballot + popc, row-major order:
      // columns < 32 first) into the padded buffers.
      if (env_e || gnn_e) {
        const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
        const unsigned lt = (1u << lane) - 1u;
        constexpr int MAXG = (GMPE_MAX_ENTITIES + 31) / 32;  // column groups of 32 entities
        const int ng = (E + 31) >> 5;
        for (int w = warp; w < nw; w += nwarps) {
          if (p.mode != MODE_STEP && !sm.wwrite[w]) continue;
          const V2* pw = sm.pos + w * E;
          const uint8_t* mk = sm.mask + w * E;
          const size_t gb = (size_t)(w0 + w);
          int32_t* ei = env_e ? p.out.edge_index + gb * 2 * cap : nullptr;
          Real* ea = env_e ? (Real*)p.out.edge_attr + gb * cap : nullptr;
          int32_t* gi = gnn_e ? p.out.gnn_edge_index + gb * 2 * cap : nullptr;
          Real* gea = gnn_e ? (Real*)p.out.gnn_edge_attr + gb * cap : nullptr;
          Real* g_adjw = p.out.adj ? (Real*)p.out.adj + gb * EE : nullptr;  // masked distance matrix, one copy per world (july:1625-1648)
          if constexpr (EC > 0 && (EC % 32) != 0 && GMPE_EDGE_LINEAR) {
            // Compile-time entity count that is not a multiple of the warp (E = 44): walk the E*E matrix
            // LINEARLY in chunks of 32 entries instead of row by row.  Row-major order is kept (the ballot
            // order inside a chunk is the matrix order), every lane is busy in every chunk but the last
            // (61 chunks against 44 x 2 half-empty row groups), and each adjacency store is one whole,
            // aligned 128-byte line (a world's matrix starts on a 32-byte boundary) instead of a 176-byte
            // row cut into a 128- and a 48-byte piece at odd 16-byte offsets.
            constexpr int EEC = EC * EC;
            int n_env = 0, n_gnn = 0;
#pragma unroll kEdgeLinearUnroll
            for (int q0 = 0; q0 < EEC; q0 += 32) {
              const int q = q0 + lane;
              const bool in = q < EEC;
              const int r = in ? q / EC : 0, cg = in ? q - r * EC : 0;
              const V2 a = pw[r], bq = pw[cg];
              const bool offd = !in || r == cg || mk[r] != 0 || mk[cg] != 0;
              const Real d = offd ? Real(0) : norm2(a.x - bq.x, a.y - bq.y);
              if (g_adjw && in) g_adjw[q] = d;
              if (env_e) {  // update_graph: (d <= max_edge_dist) & (d > 0), row-major (july:1660-1665)
                const bool on = d > Real(0) && d <= thr_env;
                const unsigned m = __ballot_sync(0xffffffffu, on);
                if (on) {
                  const int slot = n_env + __popc(m & lt);
                  ei[slot] = r; ei[cap + slot] = cg; ea[slot] = d;
                }
                n_env += __popc(m);
              }
              if (gnn_e) {  // process_adj on the fp32 cast: (adj < max_edge_dist) & (adj > 0) (gnn_new.py:339)
                const float df = (float)d;
                const bool on = df > 0.0f && df < thr_gnn;
                const unsigned m = __ballot_sync(0xffffffffu, on);
                if (on) {
                  const int slot = n_gnn + __popc(m & lt);
                  gi[slot] = r; gi[cap + slot] = cg; gea[slot] = Real(df);
                }
                n_gnn += __popc(m);
              }
            }
            if (env_e) {
              for (int q = n_env + lane; q < cap; q += 32) { ei[q] = -1; ei[cap + q] = -1; ea[q] = Real(0); }
              if (lane == 0) p.out.num_edges[gb] = n_env;
            }
            if (gnn_e) {
              for (int q = n_gnn + lane; q < cap; q += 32) { gi[q] = -1; gi[cap + q] = -1; gea[q] = Real(0); }
              if (lane == 0) p.out.gnn_num_edges[gb] = n_gnn;
            }
            continue;
          }
          V2 bcol[MAXG];
          uint32_t offm = 0, has = 0;  // bit g: the lane's column of group g is masked / exists
#pragma unroll
          for (int g = 0; g < MAXG; ++g) {
            const int cg = lane + 32 * g;
            const bool h = g < ng && cg < E;
            bcol[g] = pw[h ? cg : 0];
            has |= (uint32_t)h << g;
            offm |= (uint32_t)(!h || mk[h ? cg : 0] != 0) << g;
          }
          int n_env = 0, n_gnn = 0;
#pragma unroll (EC > 0 && EC <= 64 ? kEdgeUnrollMid : kEdgeUnroll)
          for (int r = 0; r < E; ++r) {
            const V2 a = pw[r];
            const bool mr = mk[r] != 0;
#pragma unroll
            for (int g = 0; g < MAXG; ++g) {
              if (g >= ng) break;  // uniform
              const int cg = lane + 32 * g;
              const Real d = (mr || ((offm >> g) & 1u) || r == cg) ? Real(0) : norm2(a.x - bcol[g].x, a.y - bcol[g].y);
              if (!(GMPE_XSKIP & 2) && g_adjw && ((has >> g) & 1u)) g_adjw[r * E + cg] = d;
              if (env_e) {  // update_graph: (d <= max_edge_dist) & (d > 0), row-major (july:1660-1665)
                const bool on = d > Real(0) && d <= thr_env;
                const unsigned m = __ballot_sync(0xffffffffu, on);
                if (on && !(GMPE_XSKIP & 1)) {
                  const int slot = n_env + __popc(m & lt);
                  ei[slot] = r; ei[cap + slot] = cg; ea[slot] = d;
                }
                n_env += __popc(m);
              }
              if (gnn_e) {  // process_adj on the fp32 cast: (adj < max_edge_dist) & (adj > 0) (gnn_new.py:339)
                const float df = (float)d;
                const bool on = df > 0.0f && df < thr_gnn;
                const unsigned m = __ballot_sync(0xffffffffu, on);
                if (on) {
                  const int slot = n_gnn + __popc(m & lt);
                  gi[slot] = r; gi[cap + slot] = cg; gea[slot] = Real(df);
                }
                n_gnn += __popc(m);
              }
            }
          }
          if (env_e) {
            if (!(GMPE_XSKIP & 8))
              for (int q = n_env + lane; q < cap; q += 32) { ei[q] = -1; ei[cap + q] = -1; ea[q] = Real(0); }
            if (lane == 0) p.out.num_edges[gb] = n_env;
          }
          if (gnn_e) {
            for (int q = n_gnn + lane; q < cap; q += 32) { gi[q] = -1; gi[cap + q] = -1; gea[q] = Real(0); }
            if (lane == 0) p.out.gnn_num_edges[gb] = n_gnn;
          }
        }
      }
    }
    __syncthreads();  // smem is reused by the next tile
  }
}

// argmax over the one-hot action rows the runner builds (graph_mpe_runner.py:375-377;
// environment.py:431, 444): first maximum wins, like np.argmax.
template <typename T>
__global__ void decode_onehot_kernel(const T* __restrict__ onehot, int32_t* __restrict__ idx, int64_t rows, int A) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const T* row = onehot + r * A;
  int best = 0;
  T bv = row[0];
  for (int a = 1; a < A; ++a) { T v = row[a]; if (v > bv) { bv = v; best = a; } }
  idx[r] = best;
}

}  // namespace gmpe
