// gmpe_small.cuh -- the GraphMPE step for SMALL worlds (N agents, L == N landmarks, no
// obstacles / walls, N known at compile time): one THREAD per world.
//
// Same contract as gmpe_step_kernel (gmpe_step.cuh): one launch = MultiAgentGraphEnv.step
// (multiagent/environment.py:1021-1063) for every world of the batch plus graphworker's
// auto-reset (onpolicy/envs/env_wrappers.py:861-873), or MultiAgentGraphEnv.reset
// (environment.py:1066-1081) for the masked worlds.
//
// Why a second mapping.  At N = 3 a world is ~0.3 KB of state in and ~1.4 KB of tensors out;
// the tile kernel spends most of its issue slots on index arithmetic and shared-memory
// round trips between its (world, agent) threads (profiles/r1c_*: 13.7 k thread
// instructions per world, issue-bound at 0.23 of the HBM roofline).  Here a lane owns a
// whole world: every entity lives in registers, all E(E-1)/2 distances are computed once
// and reused by the collision tests, the neighbour selection, the adjacency and both edge
// lists, and the reference's in-order side effects are reproduced by simply
// walking the agents in index order.  Loops over agents / entities are fully unrolled
// (N is a template parameter), so there is no index arithmetic left.
//
// Memory.  Inputs are read with direct vector loads (the 32 lanes of a warp cover one
// contiguous span of every array).  Outputs are bandwidth: a warp stages each output array
// of its 32 worlds in a private shared-memory slab and streams it out as contiguous
// 128-bit stores (warp-synchronous: only __syncwarp, no CTA barrier on the hot path).
#pragma once
#include "gmpe_step.cuh"

namespace gmpe {

template <typename Real, int SCEN, int N>
struct SmallTraits {
  static constexpr int E = 2 * N;
  static constexpr int EE = E * E;
  static constexpr int CAP = E * (E - 1);
  static constexpr int F = SCEN == GMPE_SCEN_JULY_TUBE ? 19 : 6;
  static constexpr int R = (int)sizeof(Real);
  static constexpr int VW = 16 / R;  // Reals per 128-bit vector
  // strided (vector-staged) rows: stride in Reals such that 8 consecutive lanes hit 8
  // distinct 16-byte bank groups (stride / VW odd)
  static constexpr int pad_stride(int rw) { return ((rw / VW) & 1) ? rw : rw + VW; }
  static constexpr int NODE_RW = 8 * E, NODE_STRIDE = pad_stride(NODE_RW);
  static constexpr int INFO_RW = GMPE_INFO_K * N, INFO_STRIDE = pad_stride(INFO_RW);
  static constexpr int ADJ_RW = EE, ADJ_STRIDE = pad_stride(ADJ_RW);
  static constexpr int OBS_RW = F * N;
  // bytes per lane of each staging round
  static constexpr int B_NODE = NODE_STRIDE * R;
  static constexpr int B_INFO = INFO_STRIDE * R;
  static constexpr int B_OBSADJ = ((OBS_RW * R + 15) / 16) * 16 + ADJ_STRIDE * R;
  static constexpr int B_EDGE = 2 * CAP * 4 + CAP * R;
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  static constexpr int LANE_BYTES = ((cmax(cmax(B_NODE, B_INFO), cmax(B_OBSADJ, B_EDGE)) + 15) / 16) * 16;
  static constexpr int SLAB_BYTES = 32 * LANE_BYTES;  // per warp
};

// launch shape of the small-world kernel
#ifndef GMPE_SMALL_WARPS
#define GMPE_SMALL_WARPS 2
#endif
#ifndef GMPE_SMALL_MINBLOCKS_F32
#define GMPE_SMALL_MINBLOCKS_F32 7
#endif
#ifndef GMPE_SMALL_MINBLOCKS_F64
#define GMPE_SMALL_MINBLOCKS_F64 3
#endif
// 1: node_obs / info rows leave the lane directly as 256-bit stores; 0: staged through the slab
#ifndef GMPE_SMALL_DIRECT_ROWS
#define GMPE_SMALL_DIRECT_ROWS 1
#endif

// ---- warp-cooperative slab -> global copies ------------------------------------------
// Rows of RW Reals per world, staged at `STRIDE` Reals per lane; world w's row goes to
// g + w * gstride.  Every 128-bit store of a row is contiguous with its neighbours.
template <typename Real, int RW, int STRIDE>
__device__ __forceinline__ void copy_rows(Real* __restrict__ g, size_t gstride, const Real* __restrict__ slab, int nw,
                                          uint32_t wmask, int lane) {
  constexpr int VW = 16 / (int)sizeof(Real);
  constexpr int VPW = RW / VW;
  static_assert(RW % VW == 0, "row must be a whole number of 128-bit vectors");
  if (nw == 32 && wmask == 0xffffffffu) {  // the hot case: a full warp of worlds, nothing masked
#pragma unroll
    for (int it = 0; it < VPW; ++it) {
      const int q = it * 32 + lane;
      const int w = q / VPW, r = q - w * VPW;
      const int4 v = *reinterpret_cast<const int4*>(slab + w * STRIDE + r * VW);
      *reinterpret_cast<int4*>(g + (size_t)w * gstride + r * VW) = v;
    }
    return;
  }
  const int total = nw * VPW;
#pragma unroll 1
  for (int q = lane; q < total; q += 32) {
    const int w = q / VPW, r = q - w * VPW;
    if ((wmask >> w) & 1u) {
      const int4 v = *reinterpret_cast<const int4*>(slab + w * STRIDE + r * VW);
      *reinterpret_cast<int4*>(g + (size_t)w * gstride + r * VW) = v;
    }
  }
}

// Flat array of RW elements per world staged contiguously (lane stride RW): the warp's
// span is one contiguous block of global memory.
template <typename T, int RW>
__device__ __forceinline__ void copy_flat(T* __restrict__ g, const T* __restrict__ slab, int nw, uint32_t wmask, uint32_t wfull,
                                          int lane) {
  constexpr int VW = 16 / (int)sizeof(T);
  const int n = nw * RW;
  if (wmask == wfull && ((reinterpret_cast<size_t>(g) & 15) == 0)) {
    constexpr int NV32 = 32 * RW / VW;  // 32 * RW * sizeof(T) is a multiple of 128
    if (nw == 32) {
#pragma unroll
      for (int it = 0; it < (NV32 + 31) / 32; ++it) {
        const int q = it * 32 + lane;
        if (NV32 % 32 == 0 || q < NV32) reinterpret_cast<int4*>(g)[q] = reinterpret_cast<const int4*>(slab)[q];
      }
    } else {
      const int nv = n / VW;
#pragma unroll 1
      for (int q = lane; q < nv; q += 32) reinterpret_cast<int4*>(g)[q] = reinterpret_cast<const int4*>(slab)[q];
      for (int q = nv * VW + lane; q < n; q += 32) g[q] = slab[q];
    }
  } else {  // masked reset launches / unaligned caller buffers
#pragma unroll 1
    for (int q = lane; q < n; q += 32) {
      const int w = q / RW;
      if ((wmask >> w) & 1u) g[q] = slab[q];
    }
  }
}

// ---- placement of one world in registers ---------------------------------------------
// random_scenario + setup_tube_params (july:440-613), set_landmarks_in_point
// (utils.py:165-193); navigation placement is defined-by-build.  Draw order and
// arithmetic are those of place_world (gmpe_step.cuh) with O = W = 0.
template <typename Real, int N>
struct Placed {
  typename RT<Real>::V4 s[N];
  typename RT<Real>::V2 lm[N];
  Real wf[12];
};

template <typename Real, int SCEN, int N>
__device__ __noinline__ Placed<Real, N> place_small(const GmpeConfig& c, int64_t gw, uint32_t ep) {
  Placed<Real, N> o;
  const uint64_t seed = c.seed;
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
    (void)philox_uniform01(seed, gw, ep, idx++);  // wall_length draw (july:368), value unused
    const double width = fmax(3 * c.entity_size * 2.5, ws * 0.15);
    const double angle = philox_uniform(seed, gw, ep, idx++, -1.5707963267948966, 1.5707963267948966);
    double tl = ws * 0.8;
    if (tp && !sc) tl += philox_uniform(seed, gw, ep, idx++, -ws * 0.3, ws * 0.1);  // tp:503
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
      do {
        const double jx = jit * philox_uniform(seed, gw, ep, idx, -ws, ws);
        const double jy = jit * philox_uniform(seed, gw, ep, idx + 1, -ws, ws);
        idx += 2;
        const double d = (ws + k) / spread;
        px = Real(ent0 + d * sa + jx); py = Real(ent1 + d * ca + jy);
        bad = false;
#pragma unroll
        for (int a = 0; a < k; ++a) bad = bad || (norm2(o.s[a].x - px, o.s[a].y - py) < sep);
      } while (bad);
      Real th = Real(0), v = Real(0);
      if (heading) { th = wrap_heading(Real(philox_uniform(seed, gw, ep, idx++, 0.0, 6.283185307179586))); v = Real(c.v_min); }
      o.s[k] = RT<Real>::mk4(px, py, th, v);
    }
#pragma unroll
    for (int l = 0; l < N; ++l) o.lm[l] = RT<Real>::mk2(Real(ext0 + sa * (-ws / 3)), Real(ext1 + ca * (-ws / 3)));
  } else {
    const double h = ws / 2;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      Real px, py;
      bool bad;
      do {
        px = Real(0.9 * philox_uniform(seed, gw, ep, idx, -h, h));
        py = Real(0.9 * philox_uniform(seed, gw, ep, idx + 1, -h, h));
        idx += 2;
        bad = false;
#pragma unroll
        for (int a = 0; a < k; ++a) bad = bad || (norm2(o.s[a].x - px, o.s[a].y - py) < sep);
      } while (bad);
      Real th = Real(0), v = Real(0);
      if (heading) { th = wrap_heading(Real(philox_uniform(seed, gw, ep, idx++, 0.0, 6.283185307179586))); v = Real(c.v_min); }
      o.s[k] = RT<Real>::mk4(px, py, th, v);
    }
    const Real lmin = Real(1.2 * (2 * c.entity_size));  // is_landmark_collision july:916-925
#pragma unroll
    for (int k = 0; k < N; ++k) {
      Real px, py;
      bool bad;
      do {
        px = Real(0.9 * philox_uniform(seed, gw, ep, idx, -h, h));
        py = Real(0.9 * philox_uniform(seed, gw, ep, idx + 1, -h, h));
        idx += 2;
        bad = false;
#pragma unroll
        for (int a = 0; a < k; ++a) bad = bad || (norm2(o.lm[a].x - px, o.lm[a].y - py) < lmin);
      } while (bad);
      o.lm[k] = RT<Real>::mk2(px, py);
    }
  }
  return o;
}

// ---- the kernel ---------------------------------------------------------------------
template <typename Real, int SCEN, int N>
__global__ void __launch_bounds__(32 * GMPE_SMALL_WARPS, sizeof(Real) == 4 ? GMPE_SMALL_MINBLOCKS_F32 : GMPE_SMALL_MINBLOCKS_F64)
gmpe_step_small_kernel(const __grid_constant__ KParams p) {
  using V2 = typename RT<Real>::V2;
  using V4 = typename RT<Real>::V4;
  using T = SmallTraits<Real, SCEN, N>;
  constexpr int E = T::E, CAP = T::CAP, F = T::F;
  constexpr bool july = SCEN == GMPE_SCEN_JULY_TUBE;
  extern __shared__ __align__(32) unsigned char smem_raw[];
  __shared__ Real acttab[2 * 32];

  const GmpeConfig& c = p.c;
  const int B = c.num_worlds;
  const bool heading = c.dynamics != GMPE_DYN_DOUBLE_INTEGRATOR;
  const bool do_info = (c.flags & GMPE_FLAG_INFO) != 0;
  const bool step_mode = p.mode == MODE_STEP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* slab = smem_raw + (size_t)warp * T::SLAB_BYTES;
  Real* slab_r = reinterpret_cast<Real*>(slab);

  // _set_action table (environment.py:336-475): identical arithmetic to decode_action
  if (threadIdx.x < 32) {
    Real u0 = Real(0), u1 = Real(0);
    if ((int)threadIdx.x < c.total_actions) decode_action<Real>(c, threadIdx.x, u0, u1);
    acttab[2 * threadIdx.x] = u0;
    acttab[2 * threadIdx.x + 1] = u1;
  }
  __syncthreads();

  Real* g_as = (Real*)p.st.agent_state;
  Real* g_aux = (Real*)p.st.agent_aux;
  uint32_t* g_fl = (uint32_t*)p.st.agent_flags;
  int32_t* g_cnt = p.st.agent_cnt;
  Real* g_wf = (Real*)p.st.world_f;
  int32_t* g_wi = p.st.world_i;
  Real* g_lm = (Real*)p.st.landmarks;

  const Real sep = Real(c.separation_dist), thr_goal = Real(c.goal_dist_thresh);
  const Real coll = Real(c.collision_rew), goal = Real(c.goal_rew);
  const Real dt = Real(c.dt);
  const bool env_e = (c.flags & GMPE_FLAG_ENV_EDGES) && p.out.edge_index;
  const bool gnn_e = (c.flags & GMPE_FLAG_GNN_EDGES) && p.out.gnn_edge_index;

  for (int64_t wbase = ((int64_t)blockIdx.x * GMPE_SMALL_WARPS + warp) * 32; wbase < B;
       wbase += (int64_t)gridDim.x * GMPE_SMALL_WARPS * 32) {
    const int nw = (int)min((int64_t)32, (int64_t)B - wbase);
    const bool valid = lane < nw;
    const int64_t b = wbase + (valid ? lane : 0);  // idle lanes shadow world wbase, never write
    const size_t ga0 = (size_t)b * N;
    const uint32_t wfull = nw == 32 ? 0xffffffffu : ((1u << nw) - 1u);

    // ---------------- loads: the whole world into registers ----------------------------
    V4 s[N];
    V2 aux[N], lm[N];
    int st[N], pp[N], pr[N], gr[N];
    Real wf[12];
    Real dsum = Real(0);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      s[i] = ((const V4*)g_as)[ga0 + i];
      aux[i] = ((const V2*)g_aux)[ga0 + i];
      const uint32_t w = g_fl[ga0 + i];
      st[i] = w & 0xff; pp[i] = (w >> 8) & 0xff; pr[i] = (w >> 16) & 0xff; gr[i] = (w >> 24) & 0xff;
      lm[i] = ((const V2*)g_lm)[ga0 + i];
    }
    if (july) {
      const Real* gw_ = g_wf + (size_t)b * GMPE_WORLD_F;
#pragma unroll
      for (int q = 0; q < 12; ++q) wf[q] = gw_[q];
      if (do_info) dsum = gw_[GMPE_WF_DELTA_SPACING];
    } else {
#pragma unroll
      for (int q = 0; q < 12; ++q) wf[q] = Real(0);
    }
    const int2 wi = *reinterpret_cast<const int2*>(g_wi + (size_t)b * GMPE_WORLD_I);
    const int wstep = step_mode ? wi.x + 1 : wi.x;  // environment.py:1024
    const int wep = wi.y;
    int32_t cnt[N][GMPE_AGENT_CNT];
    if (do_info && step_mode) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int4 a = ((const int4*)(g_cnt + (ga0 + i) * GMPE_AGENT_CNT))[0];
        const int4 bq = ((const int4*)(g_cnt + (ga0 + i) * GMPE_AGENT_CNT))[1];
        cnt[i][0] = a.x; cnt[i][1] = a.y; cnt[i][2] = a.z; cnt[i][3] = a.w;
        cnt[i][4] = bq.x; cnt[i][5] = bq.y; cnt[i][6] = bq.z; cnt[i][7] = bq.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int q = 0; q < GMPE_AGENT_CNT; ++q) cnt[i][q] = 0;
    }

    V2 vpre[N], vpost[N];
    int ph_obs[N];
    int stf[N];  // status after this step's reward() calls (graph mask, next state)
    bool rs = false;  // this world is (re)placed in this launch
    bool wr = valid;  // this world's tensors are emitted by this launch
#pragma unroll
    for (int i = 0; i < N; ++i) { vpre[i] = RT<Real>::mk2(0, 0); vpost[i] = vpre[i]; ph_obs[i] = 0; stf[i] = st[i]; }

    if (step_mode) {
      // ---------------- stage 2 + stage 1: action decode and physics --------------------
      Real u0[N], u1[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int a = p.act[ga0 + i];
        u0[i] = acttab[2 * (a & 31)];
        u1[i] = acttab[2 * (a & 31) + 1];
      }
      if (c.physics == GMPE_PHYS_FORCE) {
        // classic step order (onpolicy/envs/mpe/core.py:183-202) with the force functions of
        // multiagent/core.py:766-984; every pair term is evaluated once and applied to both
        // ends (f_b = -f_a exactly), accumulated per agent in entity order
        const bool classic = (c.flags & GMPE_FLAG_CLASSIC_CONTACT) != 0;
        const Real km = Real(c.contact_margin), cf = Real(c.contact_force);
        const Real dmin = classic ? Real(c.entity_size) + Real(c.entity_size) : sep;
        Real tx[N][N], ty[N][N], pen[N][N];
#pragma unroll
        for (int a = 0; a < N; ++a)
#pragma unroll
          for (int k = a + 1; k < N; ++k) {
            const Real dx = s[a].x - s[k].x, dy = s[a].y - s[k].y, dist = norm2(dx, dy);
            pen[a][k] = softplus(div_(-(dist - dmin), km)) * km;
            tx[a][k] = div_(cf * dx, dist);
            ty[a][k] = div_(cf * dy, dist);
          }
        Real fx[N], fy[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          fx[i] = u0[i]; fy[i] = u1[i];  // apply_action_force: mass 1, accel None, no noise
          if (classic || !st[i]) {       // core.py:899: no agent-agent force on a done entity
#pragma unroll
            for (int k = 0; k < N; ++k) {
              if (k == i) continue;
              if (i < k) { fx[i] = tx[i][k] * pen[i][k] + fx[i]; fy[i] = ty[i][k] * pen[i][k] + fy[i]; }
              else { fx[i] = (-tx[k][i]) * pen[k][i] + fx[i]; fy[i] = (-ty[k][i]) * pen[k][i] + fy[i]; }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
          if (classic || !st[i]) {  // integrate_state core.py:827-845; done agents freeze
            s[i].z = s[i].z * Real(1 - c.damping) + fx[i] * dt;
            s[i].w = s[i].w * Real(1 - c.damping) + fy[i] * dt;
            if (c.max_speed > 0) {
              const Real speed = sqrt_(s[i].z * s[i].z + s[i].w * s[i].w);
              if (speed > Real(c.max_speed)) { s[i].z = div_(s[i].z, speed) * Real(c.max_speed); s[i].w = div_(s[i].w, speed) * Real(c.max_speed); }
            }
            s[i].x += s[i].z * dt; s[i].y += s[i].w * dt;
            aux[i].x += norm2(s[i].z * dt, s[i].w * dt);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i)
          if (!st[i]) integrate_kinematic<Real>(c, s[i], u0[i], u1[i], aux[i].x);  // core.py:819-826
      }
    }

    // pre-reward velocities and heading vectors (july:1130-1131)
    Real hx[N], hy[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      hx[i] = Real(1); hy[i] = Real(0);
      if (heading) { sincos_(s[i].z, &hy[i], &hx[i]); vpre[i] = RT<Real>::mk2(s[i].w * hx[i], s[i].w * hy[i]); }
      else {
        vpre[i] = RT<Real>::mk2(s[i].z, s[i].w);
        const Real sp = norm2(s[i].z, s[i].w);  // DoubleIntegratorXYState.theta = arctan2(vy, vx), core.py:188-189
        if (sp > Real(0)) { hx[i] = div_(s[i].z, sp); hy[i] = div_(s[i].w, sp); }
      }
      vpost[i] = vpre[i];
    }

    // all pairwise distances of the world, once (calculate_distances, core.py:600-624)
    Real daa[N][N], dal[N][N], dll[N][N];
#pragma unroll
    for (int a = 0; a < N; ++a) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        dal[a][k] = norm2(s[a].x - lm[k].x, s[a].y - lm[k].y);
        if (k > a) {
          daa[a][k] = norm2(s[k].x - s[a].x, s[k].y - s[a].y);
          dll[a][k] = norm2(lm[a].x - lm[k].x, lm[a].y - lm[k].y);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < N; ++a) {
      daa[a][a] = Real(0); dll[a][a] = Real(0);
#pragma unroll
      for (int k = 0; k < a; ++k) { daa[a][k] = daa[k][a]; dll[a][k] = dll[k][a]; }
    }

    Real rew[N];
    int done[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { rew[i] = Real(0); done[i] = 0; }
    Real* info_s = slab_r + (size_t)lane * T::INFO_STRIDE;

    if (step_mode) {
      // ---------------- stages 3 + 4: agents in index order, side effects in place --------
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int st_pre = st[i];
        const Real px = s[i].x, py = s[i].y;
        const Real d_goal = dal[i][i];  // goal_match_index = arange (july:356)
        int pp_eff = pp[i], ph = 0;
        bool reach;
        if (july) {
          ph = tube_phase<Real>(wf, px, py, pp_eff);  // observation-time call may promote pp (july:727)
          reach = (ph == 2) && (pr[i] != 0) && (d_goal < thr_goal) && !st_pre;
        } else {
          reach = (d_goal < thr_goal) && (c.flags & GMPE_FLAG_USE_DONES) && !st_pre;
        }
        ph_obs[i] = ph;
        const int st_post = reach ? 1 : st_pre;
        // collisions and front / back neighbours; agents < i carry their post-reward status
        int ncoll_rew = 0, ncoll_info = 0;
        int front = -1, back = -1;
        Real fproj = Real(0), bproj = Real(0), fdist = Real(0), bdist = Real(0);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          if (k == i) continue;
          const Real d = daa[i][k];
          const int st_k = stf[k];
          const bool close = d < sep;
          ncoll_rew += (close && !st_k && !st_pre);    // is_collision at reward(i) time (july:907-914)
          ncoll_info += (close && !st_k && !st_post);  // ... at info_callback(i) time (july:785)
          if (july) {  // july:1136-1143
            const Real rx = s[k].x - px, ry = s[k].y - py;
            const Real proj = rx * hx[i] + ry * hy[i];
            if (proj > Real(0)) { if (front < 0 || proj < fproj) { front = k; fproj = proj; fdist = d; } }
            else { if (back < 0 || proj > bproj) { back = k; bproj = proj; bdist = d; } }
          }
        }
        Real r = Real(0), sp_err = Real(0);
        int cur = ph, prc = pr[i];
        if (july) {
          r -= Real(ncoll_rew) * (coll * Real(4));
          const Real tdx = wf[GMPE_WF_EXIT_X] - wf[GMPE_WF_ENT_X], tdy = wf[GMPE_WF_EXIT_Y] - wf[GMPE_WF_ENT_Y];
          const Real tube_length = norm2(tdx, tdy);
          if (ph == 2 && ph > pp_eff + 1) r -= goal * Real(3);  // july:1149
          const Real ex = px - wf[GMPE_WF_ENT_X], ey = py - wf[GMPE_WF_ENT_Y];
          const Real proj = ex * div_(tdx, tube_length) + ey * div_(tdy, tube_length);
          const Real edist = norm2(ex - proj * tdx, ey - proj * tdy);  // un-normalised direction, july:1154
          if (ph == pp_eff + 1 && prc == ph - 1) {                     // july:1156-1160
            if (ph == 1 && Real(0) <= proj && proj < Real(0.1) * tube_length && edist < Real(0.2) * tube_length) r += goal * Real(3);
            else if (ph == 2) r += goal * Real(3);
          }
          if (ph == 0) {
            r -= norm2(wf[GMPE_WF_ENT_X] - px, wf[GMPE_WF_ENT_Y] - py);
          } else if (ph == 1) {  // metering terms, july:1167-1181
            if (front >= 0) { const Real diff = fdist - sep; sp_err += diff < Real(0) ? abs_(diff) : Real(0); }
            if (back >= 0) { const Real diff = bdist - sep; sp_err += diff < Real(0) ? abs_(diff) : Real(0); }
            if (sp_err > Real(0)) cnt[i][GMPE_AC_SPACING_VIOL] += 1;
            r -= sp_err * Real(c.formation_rew);
            r -= norm2(wf[GMPE_WF_EXIT_X] - px, wf[GMPE_WF_EXIT_Y] - py);
            cnt[i][GMPE_AC_STEPS_IN_CORRIDOR] += 1;
          } else if (ph == 2 && prc == 0) {
            cur = 0;
          } else {  // july:1185-1194
            if (d_goal < thr_goal) { if (reach) r += goal * Real(5); }
            else r -= d_goal;
          }
          if (prc == 1 && cur == 0) cnt[i][GMPE_AC_CONFORMANCE] += 1;
          if (cur > prc) prc = cur;
          if (cur < pp_eff) r -= coll * Real(3);
          if (cur < prc) r -= coll;
          pp[i] = cur;
          pr[i] = prc;
          r = clip(r, -Real(4) * coll, goal * Real(5));
          r = clip(r, Real(c.min_reward), Real(c.max_reward));
        } else {
          r -= Real(ncoll_rew) * coll;
          if (d_goal < thr_goal) r += goal; else r -= d_goal;
        }
        if (reach) {  // agent.status = True; state.reset_velocity() (july:1188-1190, core.py:324-333)
          if (heading) {
            if (july) {
              Real th = p.theta ? ((const Real*)p.theta)[ga0 + i] : Real(NAN);
              if (isnan_(th))
                th = Real(philox_uniform(c.seed, c.world_offset + b, (uint32_t)wep,
                                         (1u << 30) + (uint32_t)wstep * 64u + (uint32_t)i, 0.0, 6.283185307179586));
              s[i].z = wrap_heading(th);
            }
            s[i].w = Real(c.v_min);
            Real sn, cs;
            sincos_(s[i].z, &sn, &cs);
            vpost[i] = RT<Real>::mk2(s[i].w * cs, s[i].w * sn);
          } else { s[i].z = Real(0); s[i].w = Real(0); vpost[i] = RT<Real>::mk2(0, 0); }
        }
        stf[i] = st_post;
        rew[i] = r;
        done[i] = (st_post || wstep >= c.episode_length) ? 1 : 0;  // environment.py:264-271

        if (do_info) {  // info_callback (july:741-829); int32 truncation as the int64 numpy arrays
          int nearest = 0;
          Real dmin = dal[i][0];
#pragma unroll
          for (int l = 1; l < N; ++l)
            if (dal[i][l] < dmin) { dmin = dal[i][l]; nearest = l; }
          int g = gr[i] - 1;
          const int32_t tnow = (int32_t)((double)wstep * c.dt);
          const int32_t pd = (int32_t)aux[i].x, dl = (int32_t)dmin;
          int32_t* cc = cnt[i];
          if (dmin < thr_goal && (nearest != g && g != -1)) { g = nearest; cc[GMPE_AC_DIST_LEFT_I] = dl; }
          if (dmin < thr_goal && cc[GMPE_AC_TIME_REQ_I] == -1) { cc[GMPE_AC_TIME_REQ_I] = tnow; cc[GMPE_AC_DIST_TRAV_I] = pd; cc[GMPE_AC_DIST_LEFT_I] = dl; g = nearest; }
          if (cc[GMPE_AC_TIME_REQ_I] == -1) { cc[GMPE_AC_DIST_TRAV_I] = pd; cc[GMPE_AC_DIST_LEFT_I] = dl; }
          if (dmin > thr_goal && cc[GMPE_AC_TIME_REQ_I] != -1) { cc[GMPE_AC_DIST_TRAV_I] = pd; cc[GMPE_AC_TIME_REQ_I] = tnow; cc[GMPE_AC_DIST_LEFT_I] = dl; }
          if (dmin < thr_goal && nearest == g) { cc[GMPE_AC_DIST_LEFT_I] = dl; g = nearest; }
          gr[i] = g + 1;
          cc[GMPE_AC_AGENT_COLL] += ncoll_info;
          // world-level statistics as of THIS agent's callback: agents > i still hold last
          // step's counters (july:789-793, 800-823)
          dsum += sp_err;
          Real dm = Real(0), tm = Real(0), sv = Real(0);
#pragma unroll
          for (int k = 0; k < N; ++k) {
            dm += Real(cnt[k][GMPE_AC_DIST_TRAV_I]); tm += Real(cnt[k][GMPE_AC_TIME_REQ_I]); sv += Real(cnt[k][GMPE_AC_SPACING_VIOL]);
          }
          dm = mean_div(dm, Real(N)); tm = mean_div(tm, Real(N));
          Real ds = Real(0), ts = Real(0);
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const Real x = Real(cnt[k][GMPE_AC_DIST_TRAV_I]) - dm, y = Real(cnt[k][GMPE_AC_TIME_REQ_I]) - tm;
            ds += x * x; ts += y * y;
          }
          ds = sqrt_(qdiv(ds, Real(N))); ts = sqrt_(qdiv(ts, Real(N)));
          const Real iv[GMPE_INFO_K] = {
              Real(cc[GMPE_AC_DIST_LEFT_I]), Real(cc[GMPE_AC_TIME_REQ_I]), Real(cc[GMPE_AC_AGENT_COLL]), Real(cc[GMPE_AC_OBST_COLL]),
              dm, ds, qdiv(dm, ds + Real(0.0001)), Real(cc[GMPE_AC_DIST_TRAV_I]),
              Real(cc[GMPE_AC_TIME_REQ_I]), tm, ts, qdiv(tm, ts + Real(0.0001)),
              qdiv(Real(cc[GMPE_AC_CONFORMANCE]), Real(c.episode_length)),
              qdiv(dsum, sv != Real(0) ? sv : Real(1)),
              qdiv(Real(cc[GMPE_AC_SPACING_VIOL]), cc[GMPE_AC_STEPS_IN_CORRIDOR] != 0 ? Real(cc[GMPE_AC_STEPS_IN_CORRIDOR]) : Real(1)),
              aux[i].y};
          if (GMPE_SMALL_DIRECT_ROWS) {
            if (valid && p.out.info) {
              Real* gi = (Real*)p.out.info + (ga0 + i) * GMPE_INFO_K;
              const Real lo[8] = {iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7]};
              const Real hi[8] = {iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15]};
              store_row8(gi, lo);
              store_row8(gi + 8, hi);
            }
          } else {
            V4* i4 = reinterpret_cast<V4*>(info_s + i * GMPE_INFO_K);
#pragma unroll
            for (int q = 0; q < 4; ++q) i4[q] = RT<Real>::mk4(iv[4 * q], iv[4 * q + 1], iv[4 * q + 2], iv[4 * q + 3]);
          }
        }
      }

      // ---------------- reward / done / info outputs ----------------------------------------
      Real rsum = Real(0);
      bool all_done = true;
#pragma unroll
      for (int k = 0; k < N; ++k) { rsum += rew[k]; all_done = all_done && done[k]; }
      if (valid) {
        const bool collab = (c.flags & GMPE_FLAG_COLLABORATIVE) != 0;  // environment.py:1056-1059
#pragma unroll
        for (int i = 0; i < N; ++i) {
          ((Real*)p.out.reward)[ga0 + i] = collab ? rsum : rew[i];
          if (p.out.indiv_reward) ((Real*)p.out.indiv_reward)[ga0 + i] = rew[i];
          p.out.done[ga0 + i] = (uint8_t)done[i];
          if (p.out.mask) p.out.mask[ga0 + i] = done[i] ? 0.0f : 1.0f;  // GMPERunner.insert (graph_mpe_runner.py:397-405)
          if (p.out.active_mask) p.out.active_mask[ga0 + i] = (all_done || !done[i]) ? 1.0f : 0.0f;
        }
      }
      if (do_info && !GMPE_SMALL_DIRECT_ROWS) {
        __syncwarp();
        if (p.out.info)
          copy_rows<Real, T::INFO_RW, T::INFO_STRIDE>((Real*)p.out.info + (size_t)wbase * T::INFO_RW, T::INFO_RW, slab_r, nw, wfull, lane);
        __syncwarp();
      }
      rs = valid && (c.flags & GMPE_FLAG_AUTO_RESET) && all_done;
      if (valid && !rs) {
        // ---------------- state write-back ---------------------------------------------------
#pragma unroll
        for (int i = 0; i < N; ++i) {
          ((V4*)g_as)[ga0 + i] = s[i];
          ((V2*)g_aux)[ga0 + i] = aux[i];
          g_fl[ga0 + i] = (uint32_t)stf[i] | ((uint32_t)pp[i] << 8) | ((uint32_t)pr[i] << 16) | ((uint32_t)gr[i] << 24);
        }
        g_wi[(size_t)b * GMPE_WORLD_I + GMPE_WI_STEP] = wstep;
        if (july && do_info) g_wf[(size_t)b * GMPE_WORLD_F + GMPE_WF_DELTA_SPACING] = dsum;
      }
      if (valid && do_info) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          int4* gc = (int4*)(g_cnt + (ga0 + i) * GMPE_AGENT_CNT);
          if (!rs) {
            gc[0] = make_int4(cnt[i][0], cnt[i][1], cnt[i][2], cnt[i][3]);
            gc[1] = make_int4(cnt[i][4], cnt[i][5], cnt[i][6], cnt[i][7]);
          }
        }
      }
    } else {
      const bool m = p.env_mask ? (p.env_mask[b] != 0) : true;
      rs = valid && m;
      wr = rs;
    }

    // ---------------- reset (auto-reset of finished worlds, or MODE_RESET) -------------------
    // reset_world (july:339-420): counters cleared, previous_phase survives
    if (rs) {
      if (p.has_rst) {
        const Real* rwf = (const Real*)p.rst.world_f + (size_t)b * GMPE_WORLD_F;
#pragma unroll
        for (int q = 0; q < 12; ++q) wf[q] = rwf[q];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          lm[i] = ((const V2*)p.rst.landmarks)[ga0 + i];
          s[i] = ((const V4*)p.rst.agent_state)[ga0 + i];
          aux[i].y = ((const V2*)p.rst.agent_aux)[ga0 + i].y;
        }
      } else {
        const Placed<Real, N> pl = place_small<Real, SCEN, N>(c, c.world_offset + b, (uint32_t)wep);
        if (july) {
#pragma unroll
          for (int q = 0; q < 12; ++q) wf[q] = pl.wf[q];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) { s[i] = pl.s[i]; lm[i] = pl.lm[i]; }
#pragma unroll
        for (int i = 0; i < N; ++i)  // min_time july:941-951
          aux[i].y = c.max_speed > 0 ? div_(norm2(s[i].x - lm[i].x, s[i].y - lm[i].y), Real(c.max_speed)) : Real(INFINITY);
      }
      // persist the world-level arrays
      Real* gw_ = g_wf + (size_t)b * GMPE_WORLD_F;
      if (july || p.has_rst) {
#pragma unroll
        for (int q = 0; q < 12; ++q) gw_[q] = wf[q];
      }
      gw_[GMPE_WF_DELTA_SPACING] = Real(0);
      g_wi[(size_t)b * GMPE_WORLD_I + GMPE_WI_STEP] = 0;
      g_wi[(size_t)b * GMPE_WORLD_I + GMPE_WI_EPISODE] = wep + 1;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        ((V2*)g_lm)[ga0 + i] = lm[i];
        aux[i].x = Real(0);
        stf[i] = 0;
        if (heading) { Real sn, cs; sincos_(s[i].z, &sn, &cs); vpre[i] = RT<Real>::mk2(s[i].w * cs, s[i].w * sn); }
        else vpre[i] = RT<Real>::mk2(s[i].z, s[i].w);
        vpost[i] = vpre[i];
        // observations of the fresh episode replace the terminal ones (env_wrappers.py:867-873)
        int pq = pp[i];
        ph_obs[i] = july ? tube_phase<Real>(wf, s[i].x, s[i].y, pq) : 0;
        pp[i] = pq;
        ((V4*)g_as)[ga0 + i] = s[i];
        ((V2*)g_aux)[ga0 + i] = aux[i];
        g_fl[ga0 + i] = (uint32_t)0 | ((uint32_t)pp[i] << 8);
        if (do_info || !step_mode) {
          int4* gc = (int4*)(g_cnt + (ga0 + i) * GMPE_AGENT_CNT);
          gc[0] = make_int4(0, 0, 0, 0);
          gc[1] = make_int4(0, -1, -1, -1);
        }
      }
      // distances of the fresh placement
#pragma unroll
      for (int a = 0; a < N; ++a) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
          dal[a][k] = norm2(s[a].x - lm[k].x, s[a].y - lm[k].y);
          if (k > a) {
            daa[a][k] = norm2(s[k].x - s[a].x, s[k].y - s[a].y);
            dll[a][k] = norm2(lm[a].x - lm[k].x, lm[a].y - lm[k].y);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < N; ++a)
#pragma unroll
        for (int k = 0; k < a; ++k) { daa[a][k] = daa[k][a]; dll[a][k] = dll[k][a]; }
    }
    const uint32_t wmask = __ballot_sync(0xffffffffu, wr);
    if (wmask == 0) continue;

    // ---------------- emission: obs + adj --------------------------------------------------
    // entity order: agents, landmarks (core.py:576-582); positions of the world
    {
      Real* obs_s = slab_r + (size_t)lane * T::OBS_RW;
      constexpr int ADJ_OFF = ((T::OBS_RW * T::R * 32 + 15) / 16) * 16;  // bytes, warp-level region after obs
      Real* adj_base = reinterpret_cast<Real*>(slab + ADJ_OFF);
      Real* adj_s = adj_base + (size_t)lane * T::ADJ_STRIDE;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        Real* o = obs_s + i * F;
        const Real px = s[i].x, py = s[i].y;
        if (july) {  // Scenario.observation (july:1337-1463)
          Real d1 = Real(0), d2 = Real(0);
          V2 r1 = RT<Real>::mk2(0, 0), r2 = RT<Real>::mk2(0, 0);
          int seen = 0;
#pragma unroll
          for (int k = 0; k < N; ++k) {  // two nearest others, stable on ties (july:1399-1415)
            if (k == i) continue;
            const Real rx = s[k].x - px, ry = s[k].y - py, d = daa[i][k];
            if (seen == 0 || d < d1) { d2 = d1; r2 = r1; d1 = d; r1 = RT<Real>::mk2(rx, ry); }
            else if (seen == 1 || d < d2) { d2 = d; r2 = RT<Real>::mk2(rx, ry); }
            ++seen;
          }
          const Real gx = lm[i].x - px, gy = lm[i].y - py;
          o[0] = px; o[1] = py; o[2] = vpre[i].x; o[3] = vpre[i].y;
          o[4] = gx; o[5] = gy; o[6] = Real(0); o[7] = gx; o[8] = gy;
          o[9] = r1.x; o[10] = r1.y; o[11] = r2.x; o[12] = r2.y;
          o[13] = wf[GMPE_WF_ENT_X] - px; o[14] = wf[GMPE_WF_ENT_Y] - py;
          o[15] = wf[GMPE_WF_EXIT_X] - px; o[16] = wf[GMPE_WF_EXIT_Y] - py;
          o[17] = wf[GMPE_WF_WIDTH];
          o[18] = Real(ph_obs[i]);
        } else {
          o[0] = vpre[i].x; o[1] = vpre[i].y; o[2] = px; o[3] = py; o[4] = lm[i].x - px; o[5] = lm[i].y - py;
        }
      }
      // masked distance matrix (july:1625-1648): done agents and the landmark a done agent
      // reached (goal_tracker[l] == l) are disconnected; one copy per world
      Real av[T::EE];
#pragma unroll
      for (int r = 0; r < E; ++r)
#pragma unroll
        for (int cc = 0; cc < E; ++cc) {
          const int mr = stf[r < N ? r : r - N], mc = stf[cc < N ? cc : cc - N];
          Real d;
          if (r < N && cc < N) d = daa[r][cc];
          else if (r < N) d = dal[r][cc - N];
          else if (cc < N) d = dal[cc][r - N];
          else d = dll[r - N][cc - N];
          av[r * E + cc] = (r == cc || mr || mc) ? Real(0) : d;
        }
#pragma unroll
      for (int q = 0; q < T::EE / T::VW; ++q) {
        if (T::VW == 4) reinterpret_cast<V4*>(adj_s)[q] = RT<Real>::mk4(av[4 * q], av[4 * q + 1], av[4 * q + 2], av[4 * q + 3]);
        else reinterpret_cast<V2*>(adj_s)[q] = RT<Real>::mk2(av[2 * q], av[2 * q + 1]);
      }
      __syncwarp();
      if (p.out.obs) copy_flat<Real, T::OBS_RW>((Real*)p.out.obs + (size_t)wbase * T::OBS_RW, slab_r, nw, wmask, wfull, lane);
      if (p.out.adj)
        copy_rows<Real, T::ADJ_RW, T::ADJ_STRIDE>((Real*)p.out.adj + (size_t)wbase * T::EE, T::EE, adj_base, nw, wmask, lane);
      __syncwarp();

      // ---------------- emission: edge lists ------------------------------------------------
      // update_graph (july:1651-1670): (d <= max_edge_dist) & (d > 0), row-major COO;
      // process_adj (gnn_new.py:328-358) on the fp32 cast: strict <
      const Real thr_env = Real(c.coordination_range);
      const float thr_gnn = (float)c.gnn_max_edge_dist;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const bool on_pass = pass == 0 ? env_e : gnn_e;
        if (!on_pass) continue;
        int32_t* ei = reinterpret_cast<int32_t*>(slab) + (size_t)lane * 2 * CAP;
        Real* ea_base = reinterpret_cast<Real*>(slab + (size_t)32 * 2 * CAP * 4);
        Real* ea = ea_base + (size_t)lane * CAP;
        // -1 / 0 padding of the fixed-capacity buffers (CAP is even: 8-byte vectors always fit)
#pragma unroll
        for (int q = 0; q < CAP; ++q) reinterpret_cast<int2*>(ei)[q] = make_int2(-1, -1);
#pragma unroll
        for (int q = 0; q < CAP / 2; ++q) reinterpret_cast<V2*>(ea)[q] = RT<Real>::mk2(0, 0);
        int n = 0;
#pragma unroll
        for (int r = 0; r < E; ++r)
#pragma unroll
          for (int cc = 0; cc < E; ++cc) {
            if (r == cc) continue;
            const Real d = av[r * E + cc];
            bool on;
            Real attr;
            if (pass == 0) { on = d > Real(0) && d <= thr_env; attr = d; }
            else { const float df = (float)d; on = df > 0.0f && df < thr_gnn; attr = Real(df); }
            if (on) { ei[n] = r; ei[CAP + n] = cc; ea[n] = attr; ++n; }
          }
        __syncwarp();
        int32_t* g_ei = pass == 0 ? p.out.edge_index : p.out.gnn_edge_index;
        Real* g_ea = (Real*)(pass == 0 ? p.out.edge_attr : p.out.gnn_edge_attr);
        int32_t* g_ne = pass == 0 ? p.out.num_edges : p.out.gnn_num_edges;
        copy_flat<int32_t, 2 * CAP>(g_ei + (size_t)wbase * 2 * CAP, reinterpret_cast<int32_t*>(slab), nw, wmask, wfull, lane);
        copy_flat<Real, CAP>(g_ea + (size_t)wbase * CAP, ea_base, nw, wmask, wfull, lane);
        if (wr) g_ne[b] = n;
        __syncwarp();
      }
    }

    // ---------------- emission: node features, one ego agent per round -------------------------
    // _get_entity_feat_relative (july:1694-1771): [rel_vel, rel_pos, rel_goal, occupied, type];
    // ego j sees agents <= j with their post-reward velocity, agents > j pre-reward
    if (p.out.node_obs) {
      Real* node_s = slab_r + (size_t)lane * T::NODE_STRIDE;
      Real* g_node = (Real*)p.out.node_obs + (size_t)wbase * N * T::NODE_RW;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const Real ex = s[j].x, ey = s[j].y;
        const V2 ve = vpost[j];
#pragma unroll
        for (int k = 0; k < E; ++k) {
          Real f[8];
          if (k < N) {
            const V2 vk = k <= j ? vpost[k] : vpre[k];
            f[0] = vk.x - ve.x; f[1] = vk.y - ve.y; f[2] = s[k].x - ex; f[3] = s[k].y - ey;
            f[4] = lm[k].x - ex; f[5] = lm[k].y - ey; f[6] = Real(0); f[7] = Real(0);  // landmark_poses[entity.id] (july:1740)
          } else {
            const V2 q = lm[k - N];
            f[0] = Real(0) - ve.x; f[1] = Real(0) - ve.y; f[2] = q.x - ex; f[3] = q.y - ey;
            f[4] = f[2]; f[5] = f[3]; f[6] = Real(1); f[7] = Real(1);
          }
          if (GMPE_SMALL_DIRECT_ROWS) {
            if (wr) store_row8(g_node + ((size_t)lane * N + j) * T::NODE_RW + k * 8, f);
          } else {
            V4* d4 = reinterpret_cast<V4*>(node_s + k * 8);
            d4[0] = RT<Real>::mk4(f[0], f[1], f[2], f[3]);
            d4[1] = RT<Real>::mk4(f[4], f[5], f[6], f[7]);
          }
        }
        if (!GMPE_SMALL_DIRECT_ROWS) {
          __syncwarp();
          copy_rows<Real, T::NODE_RW, T::NODE_STRIDE>(g_node + (size_t)j * T::NODE_RW, (size_t)N * T::NODE_RW, slab_r, nw, wmask, lane);
          __syncwarp();
        }
      }
    }
  }
}

}  // namespace gmpe
