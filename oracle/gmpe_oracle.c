/*
 * gmpe_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, float64, strictly sequential restatement of the reference's GraphMPE
 * step path (Jaroan/Contracts-MARL-AAM-Corridors).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (contracts-marl-aam-corridors_b200/) never does.
 *
 * Parity status: PINNED against the golden traces in tests/golden/ (.npz), which
 * tests/golden/make_golden.py produced by running the unmodified Python reference in
 * the build container (the reference ships no golden vectors of its own, SURVEY.md
 * section 4).  One third-party algorithm sits on the path and is absent from
 * /root/reference: scipy.integrate.solve_ivp(method='RK45') (scipy is unpinned in
 * requirements.txt; the traces were made with scipy 1.18.1).  rk45_solve() below
 * restates its published algorithm (Dormand-Prince 5(4) pair, Hairer/Norsett/Wanner
 * initial-step heuristic, rtol=1e-3, atol=1e-6, SAFETY 0.9, factors [0.2, 10]).
 *
 * Every function cites the reference file:line it follows.  "july" is
 * multiagent/custom_scenarios/nav_metered_one_goal_graph_rotate_tube_july.py,
 * "core" multiagent/core.py, "env" multiagent/environment.py, "classic"
 * onpolicy/envs/mpe/core.py, "wrap" onpolicy/envs/env_wrappers.py.
 *
 * The loop structure deliberately mirrors the reference (one world at a time, agents
 * in index order with their in-place side effects) and shares no code with the CUDA
 * implementation, which uses a different (two-pass, thread-per-agent) formulation.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gmpe.h"

#define MAXN GMPE_MAX_AGENTS
#define MAXE GMPE_MAX_ENTITIES
#define TWO_PI 6.283185307179586476925286766559
#define PI_D 3.14159265358979323846

/* ------------------------------------------------------------------ Philox4x32-10 */
/* Counter-based RNG shared (by specification, not by code) with the CUDA path so that
 * free-running batches can be compared: key = seed, counter = (idx, episode, world).
 * idx < 2^30: reset stream (draws in the order reset_world consumes np.random);
 * idx = 2^30 + step*64 + agent: the reset_velocity() heading draw of reward(). */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

double gmpe_oracle_uniform01(uint64_t seed, int64_t world, uint32_t episode, uint32_t idx) {
  uint32_t c[4] = {idx, episode, (uint32_t)world, (uint32_t)((uint64_t)world >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) / 9007199254740992.0;
}

static double uni(uint64_t seed, int64_t world, uint32_t ep, uint32_t idx, double lo, double hi) {
  return lo + (hi - lo) * gmpe_oracle_uniform01(seed, world, ep, idx); /* np.random.uniform */
}
/* bounded rejection sampling, as the kernels (gmpe_device.cuh): keep the last draw after this many */
#define PLACE_MAX_ATTEMPTS 1024

/* ------------------------------------------------------------------ small helpers */
static double norm2(double x, double y) { return sqrt(x * x + y * y); } /* np.linalg.norm */
static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* np.linspace(start, stop, 5)[k]  (env:429-430, 442-443) */
static double linspace5(double start, double stop, int k) {
  double step = (stop - start) / 4.0;
  return k == 4 ? stop : (double)k * step + start;
}

/* ------------------------------------------------------------------ scipy RK45 */
/* dstate: core:118-123 / 290-297 (unicycle, air_taxi), core:200-205 (double integrator) */
static void dstate(int dyn, const double y[4], const double u[2], double f[4]) {
  if (dyn == GMPE_DYN_DOUBLE_INTEGRATOR) {
    f[0] = y[2]; f[1] = y[3]; f[2] = u[0]; f[3] = u[1];
  } else {
    f[0] = y[3] * cos(y[2]); f[1] = y[3] * sin(y[2]); f[2] = u[0]; f[3] = u[1];
  }
}
static double rms4(const double x[4]) { /* scipy _ivp/common.py norm() */
  return sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]) / 2.0;
}
/* solve_ivp(ode, [0, dt], y0, method='RK45') final state (core:129, 211, 305) */
static void rk45_solve(int dyn, double y[4], const double u[2], double t_bound) {
  static const double C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1};
  static const double A[6][5] = {
      {0, 0, 0, 0, 0},
      {1.0 / 5, 0, 0, 0, 0},
      {3.0 / 40, 9.0 / 40, 0, 0, 0},
      {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
      {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
      {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
  static const double Bc[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
  static const double Ec[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
  const double rtol = 1e-3, atol = 1e-6;
  (void)C;
  double t = 0.0, f[4], K[7][4];
  dstate(dyn, y, u, f);
  /* select_initial_step (scipy _ivp/common.py) */
  double h_abs;
  {
    double sc[4], a0[4], a1[4];
    for (int i = 0; i < 4; ++i) { sc[i] = atol + fabs(y[i]) * rtol; a0[i] = y[i] / sc[i]; a1[i] = f[i] / sc[i]; }
    double d0 = rms4(a0), d1 = rms4(a1);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (h0 > t_bound) h0 = t_bound;
    double y1[4], f1[4], a2[4];
    for (int i = 0; i < 4; ++i) y1[i] = y[i] + h0 * f[i];
    dstate(dyn, y1, u, f1);
    for (int i = 0; i < 4; ++i) a2[i] = (f1[i] - f[i]) / sc[i];
    double d2 = rms4(a2) / h0;
    double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
    h_abs = fmin(fmin(100 * h0, h1), t_bound);
  }
  while (t < t_bound) {
    double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
    if (h_abs < min_step) h_abs = min_step;
    int rejected = 0;
    for (;;) {
      double t_new = t + h_abs;
      if (t_new - t_bound > 0) t_new = t_bound;
      double h = t_new - t;
      h_abs = fabs(h);
      for (int i = 0; i < 4; ++i) K[0][i] = f[i];
      for (int s = 1; s < 6; ++s) {
        double ys[4];
        for (int i = 0; i < 4; ++i) {
          double dy = 0;
          for (int j = 0; j < s; ++j) dy += K[j][i] * A[s][j];
          ys[i] = y[i] + dy * h;
        }
        dstate(dyn, ys, u, K[s]);
      }
      double yn[4], fn[4];
      for (int i = 0; i < 4; ++i) {
        double acc = 0;
        for (int j = 0; j < 6; ++j) acc += K[j][i] * Bc[j];
        yn[i] = y[i] + h * acc;
      }
      dstate(dyn, yn, u, fn);
      for (int i = 0; i < 4; ++i) K[6][i] = fn[i];
      double er[4];
      for (int i = 0; i < 4; ++i) {
        double acc = 0;
        for (int j = 0; j < 7; ++j) acc += K[j][i] * Ec[j];
        double scale = atol + fmax(fabs(y[i]), fabs(yn[i])) * rtol;
        er[i] = acc * h / scale;
      }
      double en = rms4(er);
      if (en < 1) {
        double factor = en == 0 ? 10.0 : fmin(10.0, 0.9 * pow(en, -0.2));
        if (rejected) factor = fmin(1.0, factor);
        h_abs *= factor;
        t = t_new;
        for (int i = 0; i < 4; ++i) { y[i] = yn[i]; f[i] = fn[i]; }
        break;
      }
      h_abs *= fmax(0.2, 0.9 * pow(en, -0.2));
      rejected = 1;
    }
  }
}

/* Closed form of the same ODE; what the CUDA kernels use.  Exposed so
 * tests can bound |closed form - RK45| independently of the GPU. */
static void closed_form(int dyn, double y[4], const double u[2], double dt) {
  if (dyn == GMPE_DYN_DOUBLE_INTEGRATOR) {
    y[0] += y[2] * dt + 0.5 * u[0] * dt * dt;
    y[1] += y[3] * dt + 0.5 * u[1] * dt * dt;
    y[2] += u[0] * dt; y[3] += u[1] * dt;
    return;
  }
  double th = y[2], v = y[3], w = u[0], a = u[1];
  double th1 = th + w * dt, v1 = v + a * dt;
  if (fabs(w) > 1e-9) {
    y[0] += (v1 * sin(th1) - v * sin(th)) / w + a * (cos(th1) - cos(th)) / (w * w);
    y[1] += -(v1 * cos(th1) - v * cos(th)) / w + a * (sin(th1) - sin(th)) / (w * w);
  } else {
    double d = v * dt + 0.5 * a * dt * dt;
    y[0] += cos(th) * d; y[1] += sin(th) * d;
  }
  y[2] = th1; y[3] = v1;
}

void gmpe_oracle_integrate(int dyn, int use_rk45, double y[4], const double u[2], double dt) {
  if (use_rk45) rk45_solve(dyn, y, u, dt); else closed_form(dyn, y, u, dt);
}

/* ------------------------------------------------------------------ world view */
typedef struct {
  const GmpeConfig* c;
  int N, L, O, W, E;
  int64_t b;
  double* as;    /* [N][4] */
  double* aux;   /* [N][2] */
  uint8_t* fl;   /* [N][4] */
  int32_t* cnt;  /* [N][8] */
  double* wf;    /* [16]   */
  int32_t* wi;   /* [4]    */
  double* lm;    /* [L][2] */
  double* ob;    /* [O][2] */
  double* wl;    /* [W][4] */
  int32_t* ext;  /* [N][2] (three_phase: entry_reward_cooldown, spare) or NULL */
  const double* draws; /* injected uniform [0,1) stream of this world's reset (GmpeStepInject.reset_draws) or NULL */
  int ndraws;
  int overflow;  /* a spawn rejection loop hit PLACE_MAX_ATTEMPTS */
  /* scratch for the current step */
  double pos[MAXE][2];
  double dist[MAXE][MAXE];
  double u[MAXN][2];
} World;

/* a uniform draw of world w's reset: the injected stream (the reference's own np.random draws,
 * reset-parity tests) where it has an entry, Philox otherwise */
static double wuni(const World* w, uint32_t idx, double lo, double hi) {
  if (w->draws && (int)idx < w->ndraws && w->draws[idx] == w->draws[idx]) return lo + (hi - lo) * w->draws[idx];
  return uni(w->c->seed, w->c->world_offset + w->b, (uint32_t)w->wi[GMPE_WI_EPISODE], idx, lo, hi);
}

/* GmpeConfig.formation: landmark l of L (july:492-497; utils.py:165-193 point, :77-131 line with
 * np.linspace's arithmetic for a vector with a zero step component, :231-262 circle) */
static void formation_landmark(const GmpeConfig* c, int l, int L, const double ext[2], double sa, double ca, double out[2]) {
  double ws = c->world_size;
  if (c->formation == GMPE_FORMATION_LINE) {
    double start = -ws / 2, stop = ws / 2;
    out[0] = (L > 1 && l == L - 1) ? stop : (L > 1 ? ((double)l / (double)(L - 1)) * (stop - start) + start : start);
    out[1] = -ws / 2;
  } else if (c->formation == GMPE_FORMATION_CIRCLE) {
    double ang = (double)l * (2 * PI_D / (double)L);
    out[0] = 0.0 + (ws / 3) * cos(ang);
    out[1] = (ext[1] + ws / 5) + (ws / 3) * sin(ang);
  } else {
    out[0] = ext[0] + sa * (-ws / 3);
    out[1] = ext[1] + ca * (-ws / 3);
  }
}

static void bind_world(World* w, const GmpeConfig* c, const GmpeState* st, int64_t b) {
  w->c = c; w->b = b;
  w->N = c->num_agents; w->L = c->num_landmarks; w->O = c->num_obstacles; w->W = c->num_walls;
  w->E = w->N + w->L + w->O + w->W;
  w->as = (double*)st->agent_state + b * w->N * 4;
  w->aux = (double*)st->agent_aux + b * w->N * 2;
  w->fl = st->agent_flags + b * w->N * 4;
  w->cnt = st->agent_cnt + b * w->N * GMPE_AGENT_CNT;
  w->wf = (double*)st->world_f + b * GMPE_WORLD_F;
  w->wi = st->world_i + b * GMPE_WORLD_I;
  w->lm = (double*)st->landmarks + b * w->L * 2;
  w->ob = st->obstacles ? (double*)st->obstacles + b * w->O * 2 : NULL;
  w->wl = st->walls ? (double*)st->walls + b * w->W * GMPE_WALL_F : NULL;
  w->ext = st->agent_ext ? st->agent_ext + b * w->N * 2 : NULL;
  w->draws = NULL; w->ndraws = 0; w->overflow = 0;
}

static int is_kin_heading(const GmpeConfig* c) { return c->dynamics != GMPE_DYN_DOUBLE_INTEGRATOR; }

/* state.p_vel: core:113-115, 192-193, 282-286 */
static void p_vel(const World* w, int i, double v[2]) {
  const double* s = w->as + i * 4;
  if (is_kin_heading(w->c)) { v[0] = s[3] * cos(s[2]); v[1] = s[3] * sin(s[2]); }
  else { v[0] = s[2]; v[1] = s[3]; }
}

/* Wall "position" for distance / node purposes.  The reference leaves Wall.state at
 * the origin (core:371); the build defines it as the segment midpoint. */
static void wall_pos(const double* wl, double p[2]) {
  double mid = 0.5 * (wl[2] + wl[3]);
  if (wl[0] == 0.0) { p[0] = mid; p[1] = wl[1]; } else { p[0] = wl[1]; p[1] = mid; }
}

static void gather_positions(World* w) { /* World.entities order, core:576-582 */
  int k = 0;
  for (int i = 0; i < w->N; ++i, ++k) { w->pos[k][0] = w->as[i * 4]; w->pos[k][1] = w->as[i * 4 + 1]; }
  for (int i = 0; i < w->L; ++i, ++k) { w->pos[k][0] = w->lm[i * 2]; w->pos[k][1] = w->lm[i * 2 + 1]; }
  for (int i = 0; i < w->O; ++i, ++k) { w->pos[k][0] = w->ob[i * 2]; w->pos[k][1] = w->ob[i * 2 + 1]; }
  for (int i = 0; i < w->W; ++i, ++k) wall_pos(w->wl + i * GMPE_WALL_F, w->pos[k]);
}

/* World.calculate_distances, core:600-624 */
static void calculate_distances(World* w) {
  gather_positions(w);
  for (int a = 0; a < w->E; ++a) {
    w->dist[a][a] = 0.0;
    for (int b = a + 1; b < w->E; ++b) {
      double d = norm2(w->pos[a][0] - w->pos[b][0], w->pos[a][1] - w->pos[b][1]);
      w->dist[a][b] = d; w->dist[b][a] = d;
    }
  }
}

/* MultiAgentBaseEnv._set_action, env:336-475 (one-hot branch; index = argmax) */
static void set_action(World* w, int i, int a) {
  const GmpeConfig* c = w->c;
  double u0 = 0.0, u1 = 0.0;
  if (c->dynamics == GMPE_DYN_DOUBLE_INTEGRATOR) {
    if (c->total_actions == 9) { /* env:382-392, 407-409 */
      static const double map[9][2] = {{0, 0}, {-1, 0}, {-0.71, -0.71}, {0, -1}, {0.71, -0.71},
                                       {1, 0}, {0.71, 0.71}, {0, 1}, {-0.71, 0.71}};
      u0 = map[a][0]; u1 = map[a][1];
    } else { /* env:404-406: u[0] += a[1]-a[2]; u[1] += a[3]-a[4] */
      u0 = (a == 1) - (a == 2); u1 = (a == 3) - (a == 4);
    }
  } else { /* env:424-449 */
    int ar = a / 5, ac = a - ar * 5;
    u0 = linspace5(-c->ang_rate_max, c->ang_rate_max, ar);
    u1 = linspace5(c->accel_min, c->accel_max, ac);
  }
  w->u[i][0] = u0 * c->sensitivity; /* env:460-463 */
  w->u[i][1] = u1 * c->sensitivity;
}

/* World.update_agent_state + *XYState.update_state, core:819-826, 125-139, 207-218, 300-316 */
static void kinematic_step(World* w, int use_rk45) {
  const GmpeConfig* c = w->c;
  for (int i = 0; i < w->N; ++i) {
    if (w->fl[i * 4 + GMPE_AF_STATUS]) continue; /* core:824 */
    double* s = w->as + i * 4;
    gmpe_oracle_integrate(c->dynamics, use_rk45, s, w->u[i], c->dt);
    double speed;
    if (c->dynamics == GMPE_DYN_DOUBLE_INTEGRATOR) {
      speed = sqrt(s[2] * s[2] + s[3] * s[3]);
      if (speed > c->v_max) { /* core:213-215, state.max_speed = VX_MAX */
        s[2] = c->v_max * s[2] / speed; s[3] = c->v_max * s[3] / speed;
        speed = sqrt(s[2] * s[2] + s[3] * s[3]);
      }
    } else {
      if (s[3] > c->v_max) s[3] = c->v_max;
      if (s[3] < c->v_min) s[3] = c->v_min;
      speed = s[3];
    }
    w->aux[i * 2] += speed * c->dt; /* p_dist */
  }
}

/* ------------------------------------------------------------------ force physics */
/* get_entity_collision_force core:872-906 (classic:264-295 with GMPE_FLAG_CLASSIC_CONTACT) */
static int entity_collides(const World* w, int k) { /* landmarks: collide=False (july:298) */
  return !(k >= w->N && k < w->N + w->L);
}
static int entity_movable(const World* w, int k) { return k < w->N; }

static void wall_force(const World* w, int i, const double* wl, double f[2], int* has) {
  /* get_wall_collision_force core:909-964 */
  const GmpeConfig* c = w->c;
  int prll = wl[0] == 0.0 ? 0 : 1, perp = 1 - prll;
  double size = c->entity_size;
  const double* p = w->pos[i];
  *has = 0;
  if (p[prll] < wl[2] - size || p[prll] > wl[3] + size) return;
  double theta, dist_min;
  if (p[prll] < wl[2] || p[prll] > wl[3]) {
    double past = p[prll] < wl[2] ? p[prll] - wl[2] : p[prll] - wl[3];
    theta = asin(past / size);
    dist_min = cos(theta) * size + 0.5 * c->wall_width;
  } else { theta = 0; dist_min = size + 0.5 * c->wall_width; }
  double delta = p[perp] - wl[1];
  double dist = fabs(delta);
  int classic = (c->flags & GMPE_FLAG_CLASSIC_CONTACT) != 0;
  double k = classic ? c->contact_margin : c->wall_contact_margin;
  double cf = classic ? c->contact_force : c->wall_contact_force;
  double x = -(dist - dist_min) / k;
  double pen = (x > 0 ? x + log1p(exp(-x)) : log1p(exp(x))) * k; /* np.logaddexp(0, x)*k */
  double fm = cf * delta / dist * pen;
  f[perp] = cos(theta) * fm;
  f[prll] = sin(theta) * fabs(fm);
  *has = 1;
}

static double logaddexp0(double x) { return x > 0 ? x + log1p(exp(-x)) : log1p(exp(x)); }

/* classic World.step order (classic:183-202): action force, environment force, integrate */
static void force_step(World* w) {
  const GmpeConfig* c = w->c;
  int E = w->E, N = w->N;
  int quirk = (c->flags & GMPE_FLAG_WALL_POINT_QUIRK) != 0;
  int classic = (c->flags & GMPE_FLAG_CLASSIC_CONTACT) != 0;
  static _Thread_local double F[MAXE][2];
  static _Thread_local int has[MAXE];
  gather_positions(w); /* == cached_dist_vect of the previous step's calculate_distances */
  for (int k = 0; k < E; ++k) { has[k] = 0; F[k][0] = F[k][1] = 0; }
  /* apply_action_force core:766-778: mass=1, accel=None, u_noise=None -> F = u */
  for (int i = 0; i < N; ++i) { F[i][0] = w->u[i][0]; F[i][1] = w->u[i][1]; has[i] = 1; }
  /* apply_environment_force core:781-814 */
  for (int a = 0; a < E; ++a) {
    for (int b = a + 1; b < E; ++b) {
      int is_wall_a = a >= N + w->L + w->O, is_wall_b = b >= N + w->L + w->O;
      if ((is_wall_a || is_wall_b) && !quirk) continue; /* walls are in world.entities (core:580) */
      if (!entity_collides(w, a) || !entity_collides(w, b)) continue;
      if (!entity_movable(w, a) && !entity_movable(w, b)) continue;
      double dx = w->pos[a][0] - w->pos[b][0], dy = w->pos[a][1] - w->pos[b][1];
      double dist = norm2(dx, dy);
      double size_b = is_wall_b ? c->wall_width : c->entity_size; /* Wall.size = width (core:366) */
      double dist_min = classic ? c->entity_size + size_b : c->separation_dist;
      double k = c->contact_margin;
      double pen = logaddexp0(-(dist - dist_min) / k) * k;
      double fx = c->contact_force * dx / dist * pen, fy = c->contact_force * dy / dist * pen;
      int mov_a = entity_movable(w, a), mov_b = entity_movable(w, b);
      if (mov_a && mov_b) {
        int sa = w->fl[a * 4 + GMPE_AF_STATUS], sb = w->fl[b * 4 + GMPE_AF_STATUS];
        if (classic) sa = sb = 0;
        if (!sa) { F[a][0] = fx + F[a][0]; F[a][1] = fy + F[a][1]; has[a] = 1; } /* core:899 */
        if (!sb) { F[b][0] = -fx + F[b][0]; F[b][1] = -fy + F[b][1]; has[b] = 1; }
      } else {
        if (mov_a) { F[a][0] = fx + F[a][0]; F[a][1] = fy + F[a][1]; has[a] = 1; }
        if (mov_b) { F[b][0] = -fx + F[b][0]; F[b][1] = -fy + F[b][1]; has[b] = 1; }
      }
    }
    if (entity_movable(w, a)) {
      for (int q = 0; q < w->W; ++q) {
        double wf[2] = {0, 0}; int h;
        wall_force(w, a, w->wl + q * GMPE_WALL_F, wf, &h);
        if (h) { F[a][0] = F[a][0] + wf[0]; F[a][1] = F[a][1] + wf[1]; has[a] = 1; }
      }
    }
  }
  /* integrate_state core:827-845 (classic:239-252) */
  for (int i = 0; i < N; ++i) {
    double* s = w->as + i * 4;
    if (!classic && w->fl[i * 4 + GMPE_AF_STATUS]) continue; /* build-defined: done agents freeze */
    s[2] = s[2] * (1 - c->damping); s[3] = s[3] * (1 - c->damping);
    if (has[i]) { s[2] += (F[i][0] / 1.0) * c->dt; s[3] += (F[i][1] / 1.0) * c->dt; }
    if (c->max_speed > 0) {
      double speed = sqrt(s[2] * s[2] + s[3] * s[3]);
      if (speed > c->max_speed) {
        double q = sqrt(s[2] * s[2] + s[3] * s[3]);
        s[2] = s[2] / q * c->max_speed; s[3] = s[3] / q * c->max_speed;
      }
    }
    s[0] += s[2] * c->dt; s[1] += s[3] * c->dt;
    w->aux[i * 2] += norm2(s[2] * c->dt, s[3] * c->dt); /* core:844 p_dist */
  }
}

/* ------------------------------------------------------------------ july scenario */
/* _tube_coords july:621-627: pos is cast to float32 first, n is a float32 vector */
static void tube_coords(const World* w, const double p[2], double* s, double* y) {
  double rx = (double)(float)p[0] - w->wf[GMPE_WF_ENT_X];
  double ry = (double)(float)p[1] - w->wf[GMPE_WF_ENT_Y];
  *s = rx * w->wf[GMPE_WF_E_X] + ry * w->wf[GMPE_WF_E_Y];
  *y = rx * w->wf[GMPE_WF_N_X] + ry * w->wf[GMPE_WF_N_Y];
}

/* get_agent_phase july:683-733 (mutates previous_phase at :727) */
static int get_agent_phase(World* w, int i) {
  const double eps = 0.05;
  const double* p = w->as + i * 4;
  double s, y, L = w->wf[GMPE_WF_L], hw = w->wf[GMPE_WF_HALF_W];
  tube_coords(w, p, &s, &y);
  int in_tube = (-eps <= s && s <= L + eps) && (fabs(y) <= hw + eps); /* july:630-631 */
  double tdx = w->wf[GMPE_WF_EXIT_X] - w->wf[GMPE_WF_ENT_X], tdy = w->wf[GMPE_WF_EXIT_Y] - w->wf[GMPE_WF_ENT_Y];
  double tn = norm2(tdx, tdy);
  tdx /= tn; tdy /= tn;
  int passed = ((p[0] - w->wf[GMPE_WF_EXIT_X]) * tdx + (p[1] - w->wf[GMPE_WF_EXIT_Y]) * tdy) > 0;
  double gf = 0.08 * L, gb = 0.02 * L; /* july:612-613, 633-637 */
  int valid_entrance = (-gb - eps <= s && s <= gf + eps) && (fabs(y) <= hw + eps);
  uint8_t* pp = &w->fl[i * 4 + GMPE_AF_PREV_PHASE];
  if (!in_tube && !passed) return 0;
  if (in_tube) {
    if (*pp == 0) return valid_entrance ? 1 : 0;
    return 1;
  }
  if (*pp == 1) { if (passed) { *pp = 2; return 2; } }
  else if (*pp == 2 && passed) return 2;
  return 0;
}

/* is_collision july:907-914 */
static int is_collision(const World* w, int a, int b) {
  if (w->fl[a * 4 + GMPE_AF_STATUS] || w->fl[b * 4 + GMPE_AF_STATUS]) return 0;
  const double *pa = w->as + a * 4, *pb = w->as + b * 4;
  return norm2(pa[0] - pb[0], pa[1] - pb[1]) < w->c->separation_dist;
}

/* is_obstacle_collision july:864-890 */
static int is_obstacle_collision(const World* w, const double p[2], double size) {
  for (int o = 0; o < w->O; ++o) {
    double d = norm2(w->ob[o * 2] - p[0], w->ob[o * 2 + 1] - p[1]);
    if (d < 2.0 * (w->c->entity_size + size)) return 1;
  }
  for (int q = 0; q < w->W; ++q) {
    const double* wl = w->wl + q * GMPE_WALL_F;
    int ax = wl[0] == 0.0 ? 1 : 0, al = 1 - ax; /* H wall: axis is y */
    if (wl[1] - 1.5 * size <= p[ax] && p[ax] <= wl[1] + 1.5 * size)
      if (wl[2] - 1.5 * size <= p[al] && p[al] <= wl[3] + 1.5 * size) return 1;
  }
  return 0;
}

/* Scenario.observation july:1337-1463 */
static void july_observation(World* w, int i, double* o) {
  const double* p = w->as + i * 4;
  double v[2];
  p_vel(w, i, v);
  const double* g = w->lm + i * 2; /* goal_match_index = arange (july:356) */
  double nd[MAXN], nr[MAXN][2];
  int order[MAXN], m = 0;
  for (int k = 0; k < w->N; ++k) {
    if (k == i) continue;
    nr[m][0] = w->as[k * 4] - p[0]; nr[m][1] = w->as[k * 4 + 1] - p[1];
    nd[m] = norm2(nr[m][0], nr[m][1]);
    order[m] = m; ++m;
  }
  for (int a = 1; a < m; ++a) { /* stable sort by distance, july:1409 */
    int key = order[a], b = a - 1;
    while (b >= 0 && nd[order[b]] > nd[key]) { order[b + 1] = order[b]; --b; }
    order[b + 1] = key;
  }
  int phase = get_agent_phase(w, i);
  o[0] = p[0]; o[1] = p[1]; o[2] = v[0]; o[3] = v[1];
  o[4] = g[0] - p[0]; o[5] = g[1] - p[1];
  o[6] = 0.0; /* landmark_poses_occupied is all zeros (july:506) */
  o[7] = o[4]; o[8] = o[5];
  for (int q = 0; q < 2; ++q) {
    o[9 + 2 * q] = q < m ? nr[order[q]][0] : 0.0;
    o[10 + 2 * q] = q < m ? nr[order[q]][1] : 0.0;
  }
  o[13] = w->wf[GMPE_WF_ENT_X] - p[0]; o[14] = w->wf[GMPE_WF_ENT_Y] - p[1];
  o[15] = w->wf[GMPE_WF_EXIT_X] - p[0]; o[16] = w->wf[GMPE_WF_EXIT_Y] - p[1];
  o[17] = w->wf[GMPE_WF_WIDTH];
  o[18] = (double)phase;
}

/* Scenario.reward july:1105-1221; theta_draw replaces np.random in reset_velocity (core:324-333) */
static double july_reward(World* w, int i, double theta_draw) {
  const GmpeConfig* c = w->c;
  double rew = 0;
  int phase = get_agent_phase(w, i);
  double* s = w->as + i * 4;
  uint8_t* fl = w->fl + i * 4;
  int32_t* cnt = w->cnt + i * GMPE_AGENT_CNT;
  for (int a = 0; a < w->N; ++a) { /* july:1117-1122 */
    if (a == i) continue;
    if (is_collision(w, a, i)) rew -= c->collision_rew * 4;
  }
  if (is_obstacle_collision(w, s, c->entity_size)) rew -= c->collision_rew * 3;
  double tdx = w->wf[GMPE_WF_EXIT_X] - w->wf[GMPE_WF_ENT_X], tdy = w->wf[GMPE_WF_EXIT_Y] - w->wf[GMPE_WF_ENT_Y];
  double tube_length = norm2(tdx, tdy);
  /* agent.state.theta: the heading state, or arctan2(vy, vx) for the double integrator (core:188-189) */
  double theta = is_kin_heading(c) ? s[2] : atan2(s[3], s[2]);
  double hx = cos(theta), hy = sin(theta);
  int front = -1, back = -1;
  double fproj = 0, bproj = 0;
  for (int k = 0; k < w->N; ++k) { /* july:1136-1143 */
    if (k == i) continue;
    double proj = (w->as[k * 4] - s[0]) * hx + (w->as[k * 4 + 1] - s[1]) * hy;
    if (proj > 0) { if (front < 0 || proj < fproj) { front = k; fproj = proj; } }
    else { if (back < 0 || proj > bproj) { back = k; bproj = proj; } }
  }
  int pp = fl[GMPE_AF_PREV_PHASE];
  if (phase == 2 && phase > pp + 1) rew -= c->goal_rew * 3; /* july:1149 */
  double ux = tdx / tube_length, uy = tdy / tube_length;
  double ex = s[0] - w->wf[GMPE_WF_ENT_X], ey = s[1] - w->wf[GMPE_WF_ENT_Y];
  double proj = ex * ux + ey * uy;
  double entrance_dist = norm2(ex - proj * tdx, ey - proj * tdy); /* un-normalised, july:1154 */
  if (phase == pp + 1 && fl[GMPE_AF_PHASE_REACHED] == phase - 1) { /* july:1156-1160 */
    if (phase == 1 && 0 <= proj && proj < 0.1 * tube_length && entrance_dist < 0.2 * tube_length) rew += c->goal_rew * 3;
    else if (phase == 2) rew += c->goal_rew * 3;
  }
  if (phase == 0) {
    rew -= norm2(w->wf[GMPE_WF_ENT_X] - s[0], w->wf[GMPE_WF_ENT_Y] - s[1]);
  } else if (phase == 1) { /* july:1167-1181 */
    double spacing_error = 0;
    if (front >= 0) {
      double diff = norm2(w->as[front * 4] - s[0], w->as[front * 4 + 1] - s[1]) - c->separation_dist;
      spacing_error += diff < 0 ? fabs(diff) : 0;
    }
    if (back >= 0) {
      double diff = norm2(w->as[back * 4] - s[0], w->as[back * 4 + 1] - s[1]) - c->separation_dist;
      spacing_error += diff < 0 ? fabs(diff) : 0;
    }
    if (spacing_error > 0) cnt[GMPE_AC_SPACING_VIOL] += 1;
    rew -= spacing_error * c->formation_rew;
    rew -= norm2(w->wf[GMPE_WF_EXIT_X] - s[0], w->wf[GMPE_WF_EXIT_Y] - s[1]);
    w->wf[GMPE_WF_DELTA_SPACING] += spacing_error;
    cnt[GMPE_AC_STEPS_IN_CORRIDOR] += 1;
  } else if (phase == 2 && fl[GMPE_AF_PHASE_REACHED] == 0) {
    phase = 0;
  } else { /* july:1185-1194 */
    double d = norm2(s[0] - w->lm[i * 2], s[1] - w->lm[i * 2 + 1]);
    if (d < c->goal_dist_thresh) {
      if (!fl[GMPE_AF_STATUS]) {
        fl[GMPE_AF_STATUS] = 1;
        if (is_kin_heading(c)) { s[2] = theta_draw; s[3] = c->v_min; } /* core:145-153, 324-333 */
        else { s[2] = 0; s[3] = 0; }                                  /* core:223-225 */
        rew += c->goal_rew * 5;
      }
    } else rew -= d;
  }
  if (fl[GMPE_AF_PHASE_REACHED] == 1 && phase == 0) cnt[GMPE_AC_CONFORMANCE] += 1;
  if (phase > fl[GMPE_AF_PHASE_REACHED]) fl[GMPE_AF_PHASE_REACHED] = (uint8_t)phase;
  if (phase < fl[GMPE_AF_PREV_PHASE]) rew -= c->collision_rew * 3;
  if (phase < fl[GMPE_AF_PHASE_REACHED]) rew -= c->collision_rew;
  fl[GMPE_AF_PREV_PHASE] = (uint8_t)phase;
  rew = clipd(rew, -4 * c->collision_rew, c->goal_rew * 5);
  return clipd(rew, c->min_reward, c->max_reward);
}

/* _get_entity_feat_relative july:1694-1771 (8-wide; wall rows padded to 8, type 3) */
static void node_features(World* w, int ego, double* out /* [E][8] */) {
  double ve[2];
  p_vel(w, ego, ve);
  const double* pe = w->as + ego * 4;
  int k = 0;
  for (int a = 0; a < w->N; ++a, ++k) {
    double va[2];
    p_vel(w, a, va);
    double* f = out + k * 8;
    f[0] = va[0] - ve[0]; f[1] = va[1] - ve[1];
    f[2] = w->as[a * 4] - pe[0]; f[3] = w->as[a * 4 + 1] - pe[1];
    f[4] = w->lm[a * 2] - pe[0]; f[5] = w->lm[a * 2 + 1] - pe[1]; /* landmark_poses[entity.id] */
    f[6] = 0.0; f[7] = 0.0;
  }
  for (; k < w->E; ++k) {
    double* f = out + k * 8;
    f[0] = 0.0 - ve[0]; f[1] = 0.0 - ve[1];
    f[2] = w->pos[k][0] - pe[0]; f[3] = w->pos[k][1] - pe[1];
    f[4] = f[2]; f[5] = f[3];
    f[6] = 1.0;
    f[7] = k < w->N + w->L ? 1.0 : (k < w->N + w->L + w->O ? 2.0 : 3.0);
  }
}

/* graph_observation mask july:1627-1648: done agents and reached landmarks */
static void mask_adj(World* w) {
  int dis[MAXE];
  for (int k = 0; k < w->E; ++k) dis[k] = 0;
  for (int a = 0; a < w->N; ++a) dis[a] = w->fl[a * 4 + GMPE_AF_STATUS];
  for (int l = 0; l < w->L; ++l) {
    /* np.any(goal_tracker == landmark.id); goal_tracker[a] = a once agent a reached (july:1192) */
    dis[w->N + l] = (l < w->N) ? w->fl[l * 4 + GMPE_AF_STATUS] : 0;
  }
  for (int a = 0; a < w->E; ++a)
    for (int b = 0; b < w->E; ++b)
      if (dis[a] || dis[b]) w->dist[a][b] = 0.0;
}

/* info_callback july:741-829.  times_required / dists_to_goal / dist_left_to_goal are
 * int64 numpy arrays in the reference (np.full(N, -1)) so assignments truncate. */
static void july_info(World* w, int i, double* info) {
  const GmpeConfig* c = w->c;
  const double* p = w->as + i * 4;
  int32_t* cnt = w->cnt + i * GMPE_AGENT_CNT;
  uint8_t* fl = w->fl + i * 4;
  int nearest = 0;
  double dmin = 0;
  for (int l = 0; l < w->L; ++l) {
    double d = norm2(p[0] - w->lm[l * 2], p[1] - w->lm[l * 2 + 1]);
    if (l == 0 || d < dmin) { dmin = d; nearest = l; }
  }
  int gr = (int)fl[GMPE_AF_GOAL_REACHED] - 1;
  double thr = c->goal_dist_thresh, pd = w->aux[i * 2];
  int32_t tnow = (int32_t)((double)w->wi[GMPE_WI_STEP] * c->dt);
  if (dmin < thr && (nearest != gr && gr != -1)) { gr = nearest; cnt[GMPE_AC_DIST_LEFT_I] = (int32_t)dmin; }
  if (dmin < thr && cnt[GMPE_AC_TIME_REQ_I] == -1) {
    cnt[GMPE_AC_TIME_REQ_I] = tnow; cnt[GMPE_AC_DIST_TRAV_I] = (int32_t)pd;
    cnt[GMPE_AC_DIST_LEFT_I] = (int32_t)dmin; gr = nearest;
  }
  if (cnt[GMPE_AC_TIME_REQ_I] == -1) { cnt[GMPE_AC_DIST_TRAV_I] = (int32_t)pd; cnt[GMPE_AC_DIST_LEFT_I] = (int32_t)dmin; }
  if (dmin > thr && cnt[GMPE_AC_TIME_REQ_I] != -1) {
    cnt[GMPE_AC_DIST_TRAV_I] = (int32_t)pd; cnt[GMPE_AC_TIME_REQ_I] = tnow; cnt[GMPE_AC_DIST_LEFT_I] = (int32_t)dmin;
  }
  if (dmin < thr && nearest == gr) { cnt[GMPE_AC_DIST_LEFT_I] = (int32_t)dmin; gr = nearest; }
  fl[GMPE_AF_GOAL_REACHED] = (uint8_t)(gr + 1);
  if (is_obstacle_collision(w, p, c->entity_size)) cnt[GMPE_AC_OBST_COLL] += 1;
  for (int a = 0; a < w->N; ++a) {
    if (a == i) continue;
    if (is_collision(w, i, a)) cnt[GMPE_AC_AGENT_COLL] += 1;
  }
  double dm = 0, tm = 0, ds = 0, ts = 0, sv = 0;
  for (int a = 0; a < w->N; ++a) {
    dm += w->cnt[a * GMPE_AGENT_CNT + GMPE_AC_DIST_TRAV_I];
    tm += w->cnt[a * GMPE_AGENT_CNT + GMPE_AC_TIME_REQ_I];
    sv += w->cnt[a * GMPE_AGENT_CNT + GMPE_AC_SPACING_VIOL];
  }
  dm /= w->N; tm /= w->N;
  for (int a = 0; a < w->N; ++a) {
    double x = w->cnt[a * GMPE_AGENT_CNT + GMPE_AC_DIST_TRAV_I] - dm, y = w->cnt[a * GMPE_AGENT_CNT + GMPE_AC_TIME_REQ_I] - tm;
    ds += x * x; ts += y * y;
  }
  ds = sqrt(ds / w->N); ts = sqrt(ts / w->N);
  info[GMPE_INFO_DIST_TO_GOAL] = cnt[GMPE_AC_DIST_LEFT_I];
  info[GMPE_INFO_TIME_REQ] = cnt[GMPE_AC_TIME_REQ_I];
  info[GMPE_INFO_AGENT_COLL] = cnt[GMPE_AC_AGENT_COLL];
  info[GMPE_INFO_OBST_COLL] = cnt[GMPE_AC_OBST_COLL];
  info[GMPE_INFO_DIST_MEAN] = dm;
  info[GMPE_INFO_DIST_STD] = ds;
  info[GMPE_INFO_MEAN_BY_STD] = dm / (ds + 0.0001);
  info[GMPE_INFO_DISTS_TRAVELED] = cnt[GMPE_AC_DIST_TRAV_I];
  info[GMPE_INFO_TIME_TAKEN] = cnt[GMPE_AC_TIME_REQ_I];
  info[GMPE_INFO_TIME_MEAN] = tm;
  info[GMPE_INFO_TIME_STD] = ts;
  info[GMPE_INFO_TIME_MEAN_BY_STD] = tm / (ts + 0.0001);
  info[GMPE_INFO_CONFORMANCE] = (double)cnt[GMPE_AC_CONFORMANCE] / (double)c->episode_length;
  info[GMPE_INFO_DELTA_SPACING] = w->wf[GMPE_WF_DELTA_SPACING] / (sv != 0 ? sv : 1.0);
  info[GMPE_INFO_SPACING_VIOL] = (double)cnt[GMPE_AC_SPACING_VIOL] /
                                 (cnt[GMPE_AC_STEPS_IN_CORRIDOR] != 0 ? (double)cnt[GMPE_AC_STEPS_IN_CORRIDOR] : 1.0);
  info[GMPE_INFO_MIN_TIME] = w->aux[i * 2 + 1];
}


/* ------------------------------------------------------------------ three_phase scenario */
/* "tp" = multiagent/custom_scenarios/three_phase_graph.py: the rotation-invariant family
 * (SURVEY 8f rank 3).  reset_world / info_callback / is_collision / min_time are identical to
 * july (diff of tp:315-399, 723-815, 847-956 against july:339-420, 741-829, 864-973 is empty up
 * to a float32 cast and one extra info key), so those are shared. */
static int is_two_phase(const World* w) { return w->c->scenario == GMPE_SCEN_TWO_PHASE; } /* two_phase_graph.py ("tp2") */
/* nav_graph_metered_single_corridor_rot_inv.py ("sc"): same family again */
static int is_corridor(const World* w) { return w->c->scenario == GMPE_SCEN_CORRIDOR_ROT_INV; }
static double agent_theta(const World* w, int i) { /* state.theta; arctan2(vy, vx) for the DI (core:188-189) */
  const double* s = w->as + i * 4;
  return is_kin_heading(w->c) ? s[2] : atan2(s[3], s[2]);
}
/* get_rotated_position_from_relative tp:90-97: float64 rotation matrix times the vector */
static void rot_rel(double h, double vx, double vy, double out[2]) {
  double c = cos(h), s = sin(h);
  out[0] = c * vx + s * vy;
  out[1] = -s * vx + c * vy;
}
static double pymod2pi(double x) { /* Python float %: result takes the sign of the (positive) divisor */
  double m = fmod(x, TWO_PI);
  if (m < 0) m += TWO_PI;
  return m;
}
static double tp_entrance_gate_distance(double s, double y, double hw) { /* tp:610-615 */
  double cy = clipd(y, -hw, hw);
  return hypot(fabs(s), y - cy);
}
static double tp_exit_gate_distance(double s, double y, double L, double hw) { /* tp:624-633, penalize_backward=False */
  double cy = clipd(y, -hw, hw);
  double ds = L - s > 0.0 ? L - s : 0.0;
  return hypot(ds, y - cy);
}
static int tp_in_tube_rect(double s, double y, double L, double hw) { return (-0.05 <= s && s <= L + 0.05) && (fabs(y) <= hw + 0.05); }
static int tp_in_entrance_gate(double s, double y, double L, double hw) { /* tp:603-607, ratios tp:584-585 */
  double gf = 0.08 * L, gb = 0.02 * L;
  return (-gb - 0.05 <= s && s <= gf + 0.05) && (fabs(y) <= hw + 0.05);
}
static int tp_in_exit_gate_r(double s, double y, double L, double hw, double back_ratio) { /* tp:618-621, ratios tp:586-587 (sc:621: 0.05) */
  double eb = back_ratio * L, ef = 0.08 * L;
  return (L - eb - 0.05 <= s && s <= L + ef + 0.05) && (fabs(y) <= hw + 0.05);
}
#define tp_in_exit_gate(s, y, L, hw) tp_in_exit_gate_r(s, y, L, hw, is_corridor(w) ? 0.05 : 0.02)

/* get_agent_phase tp:639-705: decrements the entry cooldown on EVERY call (observation and
 * reward both call it); does not touch previous_phase */
static int tp_get_agent_phase(World* w, int i) {
  const double* p = w->as + i * 4;
  double s, y, L = w->wf[GMPE_WF_L], hw = w->wf[GMPE_WF_HALF_W];
  tube_coords(w, p, &s, &y);
  int in_tube = tp_in_tube_rect(s, y, L, hw);
  int passed = s > L;
  int valid_entrance = tp_in_entrance_gate(s, y, L, hw), valid_exit = tp_in_exit_gate(s, y, L, hw);
  if (w->ext[i * 2] > 0) w->ext[i * 2] -= 1;
  int pp = w->fl[i * 4 + GMPE_AF_PREV_PHASE];
  if (!in_tube && !passed) return 0;
  if (in_tube) {
    if (pp == 0) return valid_entrance ? 1 : 0;
    if (is_corridor(w)) return 1; /* sc:693-695 */
    if (pp == 1 && valid_exit) return 2;
    if (pp == 2 && valid_exit && !is_two_phase(w)) return 2; /* tp:681-683; absent from tp2:676-681 */
    return 1;
  }
  /* tp2 falls off the end (returns None -> float(None) raises) when an agent passes the exit
   * plane without ever having entered; this build returns 0 there like tp:704 */
  if (passed && w->fl[i * 4 + GMPE_AF_PHASE_REACHED] >= 1) {
    if (pp == 1 && valid_exit) return 2;
    if (pp == 2) return 2;
    return 0;
  }
  return 0;
}

/* Scenario.observation tp:1161-1250 -> 15 float32 values */
static void tp_observation(World* w, int i, double* o) {
  const double* p = w->as + i * 4;
  double h = agent_theta(w, i);
  double speed = is_kin_heading(w->c) ? p[3] : norm2(p[2], p[3]);
  double g[2];
  if (is_two_phase(w)) /* tp2:1196-1198: the goal slot holds the float32 tube exit, rotated */
    rot_rel(h, (double)(float)w->wf[GMPE_WF_EXIT_X] - p[0], (double)(float)w->wf[GMPE_WF_EXIT_Y] - p[1], g);
  else rot_rel(h, w->lm[i * 2] - p[0], w->lm[i * 2 + 1] - p[1], g); /* goal_match_index = arange (tp:325) */
  double nd[MAXN], nr[MAXN][2];
  int order[MAXN], m = 0;
  for (int k = 0; k < w->N; ++k) { /* tp:1182-1191: done agents are not neighbours */
    if (k == i || (w->fl[k * 4 + GMPE_AF_STATUS] && !is_corridor(w))) continue; /* sc:1470-1476 keeps done agents */
    nr[m][0] = w->as[k * 4] - p[0]; nr[m][1] = w->as[k * 4 + 1] - p[1];
    nd[m] = norm2(nr[m][0], nr[m][1]);
    order[m] = m; ++m;
  }
  for (int a = 1; a < m; ++a) { /* list.sort is stable */
    int key = order[a], b = a - 1;
    while (b >= 0 && nd[order[b]] > nd[key]) { order[b + 1] = order[b]; --b; }
    order[b + 1] = key;
  }
  double nn[4] = {0, 0, 0, 0};
  for (int q = 0; q < 2 && q < m; ++q) { /* np.asarray(vec, float32) BEFORE the rotation (tp:1199) */
    double r[2];
    rot_rel(h, (double)(float)nr[order[q]][0], (double)(float)nr[order[q]][1], r);
    nn[2 * q] = r[0]; nn[2 * q + 1] = r[1];
  }
  int phase = tp_get_agent_phase(w, i);
  double s, y, L = w->wf[GMPE_WF_L], hw = w->wf[GMPE_WF_HALF_W];
  tube_coords(w, p, &s, &y);
  double he = pymod2pi(h - atan2(w->wf[GMPE_WF_E_Y], w->wf[GMPE_WF_E_X]) + PI_D) - PI_D;
  double raw[15] = {cos(h), sin(h), speed, g[0], g[1], nn[0], nn[1], nn[2], nn[3],
                    clipd(s / L, -2.0, 2.0), clipd(y / (hw + 1e-9), -2.0, 2.0),
                    tp_exit_gate_distance(s, y, L, hw) / (L + 1e-9), cos(he), sin(he), (double)phase};
  if (is_corridor(w)) { /* sc:1536-1548: no heading-alignment pair -> 13 values */
    raw[12] = raw[14];
    for (int q = 0; q < 13; ++q) o[q] = (double)(float)raw[q];
    return;
  }
  for (int q = 0; q < 15; ++q) o[q] = (double)(float)raw[q]; /* .astype(np.float32), tp:1250 */
}

/* Scenario.reward tp:957-1159 */
static double tp_reward(World* w, int i, double theta_draw) {
  const GmpeConfig* c = w->c;
  double rew = 0;
  int phase = tp_get_agent_phase(w, i);
  double* st = w->as + i * 4;
  uint8_t* fl = w->fl + i * 4;
  int32_t* cnt = w->cnt + i * GMPE_AGENT_CNT;
  for (int a = 0; a < w->N && !is_two_phase(w); ++a) { /* tp:965-970; commented out in tp2:962-970; x4 in sc:1131-1137 */
    if (a == i) continue;
    if (is_collision(w, a, i)) rew -= c->collision_rew * (is_corridor(w) ? 4 : 1);
  }
  double theta = agent_theta(w, i);
  double hx = cos(theta), hy = sin(theta);
  double s, y, L = w->wf[GMPE_WF_L], hw = w->wf[GMPE_WF_HALF_W];
  tube_coords(w, st, &s, &y);
  int valid_exit = tp_in_exit_gate(s, y, L, hw);
  int front = -1, back = -1;
  double fproj = 0, bproj = 0;
  for (int k = 0; k < w->N; ++k) { /* tp:991-1003 */
    if (k == i) continue;
    double proj = (w->as[k * 4] - st[0]) * hx + (w->as[k * 4 + 1] - st[1]) * hy;
    if (proj > 0) { if (front < 0 || proj < fproj) { front = k; fproj = proj; } }
    else { if (back < 0 || proj > bproj) { back = k; bproj = proj; } }
  }
  int pp = fl[GMPE_AF_PREV_PHASE];
  if (phase == 2 && phase > pp + 1) rew -= c->goal_rew; /* tp:1018-1019 */
  if (phase == pp + 1 && fl[GMPE_AF_PHASE_REACHED] == phase - 1) { /* tp:1028-1042 */
    if (phase == 1 && tp_in_entrance_gate(s, y, L, hw) && w->ext[i * 2] == 0) {
      rew += c->goal_rew;
      /* phase_reward_cooldown_steps = episode_length (tp:255), episode_length / 10 in sc:268 (a float,
       * truncated by the int32 array it is stored in) */
      w->ext[i * 2] = is_corridor(w) ? (int32_t)((double)c->episode_length / 10) : c->episode_length;
      fl[GMPE_AF_PHASE_REACHED] = 1;
    } else if (phase == 2) {
      rew += c->goal_rew;
      fl[GMPE_AF_PHASE_REACHED] = 2;
      if (is_two_phase(w) && !fl[GMPE_AF_STATUS]) { /* tp2:1041-1044: leaving through the exit gate IS the goal */
        fl[GMPE_AF_STATUS] = 1;
        if (is_kin_heading(c)) { st[2] = theta_draw; st[3] = c->v_min; }
        else { st[2] = 0; st[3] = 0; }
        rew += c->goal_rew * 5;
      }
    }
  }
  double corridor_heading = atan2(w->wf[GMPE_WF_E_Y], w->wf[GMPE_WF_E_X]);
  if (phase == 0) { /* tp:1044-1063 */
    double d = tp_entrance_gate_distance(s, y, hw);
    rew -= d;
    double he = fabs(pymod2pi(theta - corridor_heading + PI_D) - PI_D);
    if (d < c->world_size * 0.1 && !is_corridor(w)) rew -= he * c->formation_rew * 0.5; /* no heading term in sc:1246-1249 */
  } else if (phase == 1) { /* tp:1065-1108 */
    double spacing_error = 0;
    if (front >= 0) {
      double diff = norm2(w->as[front * 4] - st[0], w->as[front * 4 + 1] - st[1]) - c->separation_dist;
      spacing_error += diff < 0 ? fabs(diff) : 0;
    }
    if (back >= 0) {
      double diff = norm2(w->as[back * 4] - st[0], w->as[back * 4 + 1] - st[1]) - c->separation_dist;
      spacing_error += diff < 0 ? fabs(diff) : 0;
    }
    if (spacing_error > 0) cnt[GMPE_AC_SPACING_VIOL] += 1;
    rew -= spacing_error * c->formation_rew;
    rew -= tp_exit_gate_distance(s, y, L, hw);
    cnt[GMPE_AC_STEPS_IN_CORRIDOR] += 1;
    if (!is_two_phase(w)) w->wf[GMPE_WF_DELTA_SPACING] += spacing_error; /* tp:1101; not in tp2 */
    double he = fabs(pymod2pi(theta - corridor_heading + PI_D) - PI_D);
    if (!is_corridor(w)) rew -= he * c->formation_rew * 0.1;
    else { /* sc:1265-1276: progress along the tube axis since the last in-tube step (prev_proj is float32) */
      double tdx = w->wf[GMPE_WF_EXIT_X] - w->wf[GMPE_WF_ENT_X], tdy = w->wf[GMPE_WF_EXIT_Y] - w->wf[GMPE_WF_ENT_Y];
      double tl = norm2(tdx, tdy);
      double proj = (st[0] - w->wf[GMPE_WF_ENT_X]) * (tdx / tl) + (st[1] - w->wf[GMPE_WF_ENT_Y]) * (tdy / tl);
      float prev;
      memcpy(&prev, &w->ext[i * 2 + 1], 4);
      double dp = proj - (double)prev;
      double gain = c->goal_rew / (c->world_size * 0.8 * 10); /* progress_gain, sc:538 */
      rew += gain * (dp > -0.05 ? dp : -0.05);
      float pf = (float)proj;
      memcpy(&w->ext[i * 2 + 1], &pf, 4);
    }
  } else if (phase == 2 && is_corridor(w) && fl[GMPE_AF_PHASE_REACHED] == 0) {
    phase = 0; /* sc:1278-1279 */
  } else if (phase == 2 && !is_two_phase(w)) { /* tp:1110-1123; goal_tracker is NOT updated here (line commented out) */
    double d = norm2(st[0] - w->lm[i * 2], st[1] - w->lm[i * 2 + 1]);
    if (d < c->goal_dist_thresh) {
      if (!fl[GMPE_AF_STATUS]) {
        fl[GMPE_AF_STATUS] = 1;
        if (is_kin_heading(c)) { st[2] = theta_draw; st[3] = c->v_min; }
        else { st[2] = 0; st[3] = 0; }
        rew += c->goal_rew * 5; /* sc:1294 also records goal_tracker: see tp_mask_adj */
      }
    } else rew -= d;
  }
  if (fl[GMPE_AF_PHASE_REACHED] == 1 && phase == 0) cnt[GMPE_AC_CONFORMANCE] += 1;
  if (phase > fl[GMPE_AF_PHASE_REACHED]) fl[GMPE_AF_PHASE_REACHED] = (uint8_t)phase;
  if (phase < fl[GMPE_AF_PREV_PHASE]) rew -= c->collision_rew;
  if (phase < fl[GMPE_AF_PHASE_REACHED]) rew -= c->collision_rew;
  fl[GMPE_AF_PREV_PHASE] = (uint8_t)phase;
  if (tp_in_tube_rect(s, y, L, hw) && phase != 1 && (is_two_phase(w) || is_corridor(w) || !valid_exit)) rew -= c->collision_rew; /* tp:1145-1147, tp2:1130-1132, sc:1308-1310 */
  if (s > L && fl[GMPE_AF_PHASE_REACHED] < 1) rew -= c->goal_rew;                             /* tp:1149-1151 */
  return clipd(rew, -4 * c->collision_rew, c->goal_rew * 5);
}

/* _get_entity_feat_relative tp:1391-1430: 7 float32 features per node, float32 differences
 * rotated into the ego heading frame with a float64 matrix */
static void tp_node_features(World* w, int ego, double* out /* [E][7] */) {
  double ve64[2];
  p_vel(w, ego, ve64);
  float pe[2] = {(float)w->as[ego * 4], (float)w->as[ego * 4 + 1]}, ve[2] = {(float)ve64[0], (float)ve64[1]};
  double h = agent_theta(w, ego);
  for (int k = 0; k < w->E; ++k) {
    float pk[2], vk[2] = {0.0f, 0.0f};
    if (k < w->N) {
      double v64[2];
      p_vel(w, k, v64);
      pk[0] = (float)w->as[k * 4]; pk[1] = (float)w->as[k * 4 + 1];
      vk[0] = (float)v64[0]; vk[1] = (float)v64[1];
    } else { pk[0] = (float)w->pos[k][0]; pk[1] = (float)w->pos[k][1]; }
    double rp[2], rv[2], rg[2];
    rot_rel(h, (double)(float)(pk[0] - pe[0]), (double)(float)(pk[1] - pe[1]), rp);
    rot_rel(h, (double)(float)(vk[0] - ve[0]), (double)(float)(vk[1] - ve[1]), rv);
    if (k < w->N) {
      float gx = (float)w->lm[k * 2], gy = (float)w->lm[k * 2 + 1]; /* landmark_poses[goal_match_index[entity.id]] */
      if (is_two_phase(w)) { gx = (float)w->wf[GMPE_WF_EXIT_X]; gy = (float)w->wf[GMPE_WF_EXIT_Y]; } /* tp2:1405 */
      rot_rel(h, (double)(float)(gx - pe[0]), (double)(float)(gy - pe[1]), rg);
    } else { rg[0] = rp[0]; rg[1] = rp[1]; }
    double* f = out + k * 7;
    f[0] = (double)(float)rv[0]; f[1] = (double)(float)rv[1];
    f[2] = (double)(float)rp[0]; f[3] = (double)(float)rp[1];
    f[4] = (double)(float)rg[0]; f[5] = (double)(float)rg[1];
    f[6] = k < w->N ? 0.0 : 1.0; /* entity_mapping: agent 0, landmark 1 */
  }
}

/* graph_observation mask tp:1331-1343: goal_tracker is never set in this scenario, so only
 * done agents are disconnected */
static void tp_mask_adj(World* w) {
  if (is_corridor(w)) { mask_adj(w); return; } /* sc:1294 records goal_tracker like july */
  for (int a = 0; a < w->N; ++a)
    if (w->fl[a * 4 + GMPE_AF_STATUS])
      for (int b = 0; b < w->E; ++b) { w->dist[a][b] = 0.0; w->dist[b][a] = 0.0; }
}

/* random_scenario + setup_tube_params tp:416-587 */
static void tp_random_scenario(World* w) {
  const GmpeConfig* c = w->c;
  uint32_t idx = 0;
  double ws = c->world_size;
  (void)wuni(w, idx++, 0.2, 0.8); /* wall_length draw, tp:343 (value unused) */
  double width = fmax(3 * c->entity_size * 2.5, ws * 0.15);
  double angle = wuni(w, idx++, -PI_D / 2, PI_D / 2);
  double tl = ws * 0.8;
  if (!is_corridor(w)) tl += wuni(w, idx++, -ws * 0.3, ws * 0.1); /* tp:503; fixed length in sc:537 */
  double ca = cos(angle), sa = sin(angle);
  double ent[2] = {sa * (tl / 4), ca * (tl / 4)};
  double ext[2] = {sa * (-tl / 4), ca * (-tl / 4)};
  double L = norm2(ext[0] - ent[0], ext[1] - ent[1]) + 1e-9;
  double e[2] = {(ext[0] - ent[0]) / L, (ext[1] - ent[1]) / L};
  w->wf[GMPE_WF_ENT_X] = ent[0]; w->wf[GMPE_WF_ENT_Y] = ent[1];
  w->wf[GMPE_WF_EXIT_X] = ext[0]; w->wf[GMPE_WF_EXIT_Y] = ext[1];
  w->wf[GMPE_WF_E_X] = e[0]; w->wf[GMPE_WF_E_Y] = e[1];
  w->wf[GMPE_WF_N_X] = (double)(float)(-e[1]); w->wf[GMPE_WF_N_Y] = (double)(float)e[0];
  w->wf[GMPE_WF_L] = L; w->wf[GMPE_WF_HALF_W] = width * 0.5;
  w->wf[GMPE_WF_WIDTH] = width; w->wf[GMPE_WF_ANGLE] = angle;
  int k = 0, tries = 0;
  while (k < w->N) { /* tp:424-443 */
    double jx = 0.3 * wuni(w, idx, -ws, ws), jy = 0.3 * wuni(w, idx + 1, -ws, ws);
    idx += 2;
    double d = (ws + k) / 3;
    double p[2] = {ent[0] + d * sa + jx, ent[1] + d * ca + jy};
    int bad = 0;
    for (int a = 0; a < k && !bad; ++a)
      if (norm2(w->as[a * 4] - p[0], w->as[a * 4 + 1] - p[1]) < c->separation_dist) bad = 1;
    if (bad && ++tries < PLACE_MAX_ATTEMPTS) continue;
    if (bad) w->overflow = 1;
    tries = 0;
    double* s = w->as + k * 4;
    s[0] = p[0]; s[1] = p[1];
    if (is_kin_heading(c)) { s[2] = wuni(w, idx++, 0, TWO_PI); s[3] = c->v_min; }
    else { s[2] = 0; s[3] = 0; }
    ++k;
  }
  for (int l = 0; l < w->L; ++l) formation_landmark(c, l, w->L, ext, sa, ca, w->lm + l * 2); /* tp:459-464 */
}

/* ------------------------------------------------------------------ nav scenario (D1) */
/* navigation_graph-class scenario: absent from the reference, defined by the
 * build as a composition of reference pieces.  6-d observation
 * [vel, pos, goal-pos]; reward -dist or +goal_rew, -collision_rew per agent within
 * COLLISION_DISTANCE and per obstacle hit. */
static void nav_observation(World* w, int i, double* o) {
  const double* s = w->as + i * 4;
  double v[2];
  p_vel(w, i, v);
  o[0] = v[0]; o[1] = v[1]; o[2] = s[0]; o[3] = s[1];
  o[4] = w->lm[i * 2] - s[0]; o[5] = w->lm[i * 2 + 1] - s[1];
}

static double nav_reward(World* w, int i) {
  const GmpeConfig* c = w->c;
  double* s = w->as + i * 4;
  uint8_t* fl = w->fl + i * 4;
  double rew = 0;
  for (int a = 0; a < w->N; ++a) { /* collisions first, as july:1117-1124 */
    if (a == i) continue;
    if (is_collision(w, a, i)) rew -= c->collision_rew;
  }
  if (is_obstacle_collision(w, s, c->entity_size)) rew -= c->collision_rew;
  double d = norm2(s[0] - w->lm[i * 2], s[1] - w->lm[i * 2 + 1]);
  if (d < c->goal_dist_thresh) {
    rew += c->goal_rew;
    if ((c->flags & GMPE_FLAG_USE_DONES) && !fl[GMPE_AF_STATUS]) {
      fl[GMPE_AF_STATUS] = 1;
      if (is_kin_heading(c)) { s[3] = c->v_min; } else { s[2] = 0; s[3] = 0; }
    }
  } else rew -= d;
  return rew;
}

/* ------------------------------------------------------------------ edges */
/* update_graph july:1651-1670: (d <= max_edge_dist) & (d > 0), csr->coo = row-major */
static void emit_edges(const World* w, double thr, int strict, int as_f32, int32_t* ei, double* ea, int32_t* ne) {
  int E = w->E, cap = E * (E - 1), n = 0;
  for (int q = 0; q < cap; ++q) { ei[q] = -1; ei[cap + q] = -1; ea[q] = 0.0; }
  for (int a = 0; a < E; ++a)
    for (int b = 0; b < E; ++b) {
      double d = w->dist[a][b];
      int on;
      if (as_f32) { /* process_adj on the fp32 cast the buffer holds (gnn_new.py:339; graph_buffer.py:84-113) */
        float df = (float)d;
        on = (df < (float)thr) && (df > 0.0f);
        d = (double)df;
      } else on = strict ? (d < thr && d > 0) : (d <= thr && d > 0);
      if (on) { ei[n] = a; ei[cap + n] = b; ea[n] = d; ++n; }
    }
  *ne = n;
}

/* ------------------------------------------------------------------ reset */
static void reset_counters(World* w) { /* reset_world july:339-374 */
  for (int i = 0; i < w->N; ++i) {
    uint8_t* fl = w->fl + i * 4;
    int32_t* cnt = w->cnt + i * GMPE_AGENT_CNT;
    fl[GMPE_AF_STATUS] = 0; fl[GMPE_AF_PHASE_REACHED] = 0; fl[GMPE_AF_GOAL_REACHED] = 0;
    /* previous_phase deliberately survives the reset */
    cnt[GMPE_AC_SPACING_VIOL] = 0; cnt[GMPE_AC_STEPS_IN_CORRIDOR] = 0; cnt[GMPE_AC_CONFORMANCE] = 0;
    cnt[GMPE_AC_AGENT_COLL] = 0; cnt[GMPE_AC_OBST_COLL] = 0;
    cnt[GMPE_AC_TIME_REQ_I] = -1; cnt[GMPE_AC_DIST_TRAV_I] = -1; cnt[GMPE_AC_DIST_LEFT_I] = -1;
    w->aux[i * 2] = 0.0;
    if (w->ext) { w->ext[i * 2] = 0; w->ext[i * 2 + 1] = 0; } /* entry_reward_cooldown, tp:348 */
  }
  w->wf[GMPE_WF_DELTA_SPACING] = 0.0;
  w->wi[GMPE_WI_STEP] = 0;
}

/* random_scenario + setup_tube_params july:440-613, set_landmarks_in_point utils.py:165-193 */
static void july_random_scenario(World* w) {
  const GmpeConfig* c = w->c;
  uint32_t idx = 0;
  double ws = c->world_size;
  (void)wuni(w, idx++, 0.2, 0.8); /* wall_length draw, july:368 (value unused) */
  double width = fmax(3 * c->entity_size * 2.5, ws * 0.15);
  double angle = wuni(w, idx++, -PI_D / 2, PI_D / 2);
  double tl = ws * 0.8, ca = cos(angle), sa = sin(angle);
  double ent[2] = {sa * (tl / 4), ca * (tl / 4)};
  double ext[2] = {sa * (-tl / 4), ca * (-tl / 4)};
  double L = norm2(ext[0] - ent[0], ext[1] - ent[1]) + 1e-9;
  double e[2] = {(ext[0] - ent[0]) / L, (ext[1] - ent[1]) / L};
  w->wf[GMPE_WF_ENT_X] = ent[0]; w->wf[GMPE_WF_ENT_Y] = ent[1];
  w->wf[GMPE_WF_EXIT_X] = ext[0]; w->wf[GMPE_WF_EXIT_Y] = ext[1];
  w->wf[GMPE_WF_E_X] = e[0]; w->wf[GMPE_WF_E_Y] = e[1];
  w->wf[GMPE_WF_N_X] = (double)(float)(-e[1]); w->wf[GMPE_WF_N_Y] = (double)(float)e[0];
  w->wf[GMPE_WF_L] = L; w->wf[GMPE_WF_HALF_W] = width * 0.5;
  w->wf[GMPE_WF_WIDTH] = width; w->wf[GMPE_WF_ANGLE] = angle;
  int k = 0, tries = 0;
  while (k < w->N) { /* july:452-486 */
    double jx = 0.2 * wuni(w, idx, -ws, ws), jy = 0.2 * wuni(w, idx + 1, -ws, ws);
    idx += 2;
    double d = (ws + k) / 5;
    double p[2] = {ent[0] + d * sa + jx, ent[1] + d * ca + jy};
    int bad = is_obstacle_collision(w, p, c->entity_size);
    for (int a = 0; a < k && !bad; ++a)
      if (norm2(w->as[a * 4] - p[0], w->as[a * 4 + 1] - p[1]) < c->separation_dist) bad = 1;
    if (bad && ++tries < PLACE_MAX_ATTEMPTS) continue;
    if (bad) w->overflow = 1;
    tries = 0;
    double* s = w->as + k * 4;
    s[0] = p[0]; s[1] = p[1];
    if (is_kin_heading(c)) { s[2] = wuni(w, idx++, 0, TWO_PI); s[3] = c->v_min; }
    else { s[2] = 0; s[3] = 0; }
    ++k;
  }
  for (int l = 0; l < w->L; ++l) formation_landmark(c, l, w->L, ext, sa, ca, w->lm + l * 2); /* july:492-497 */
}

/* nav reset (defined-by-build): obstacles, agents, goals at
 * 0.9*U(-ws/2, ws/2)^2 with the reference's rejection tests; 4 boundary walls. */
static void nav_random_scenario(World* w) {
  const GmpeConfig* c = w->c;
  uint32_t idx = 0;
  double ws = c->world_size, h = ws / 2;
  for (int q = 0; q < w->W; ++q) { /* q: 0 H top, 1 H bottom, 2 V right, 3 V left, then repeats */
    double* wl = w->wl + q * GMPE_WALL_F;
    wl[0] = (q & 2) ? 1.0 : 0.0; wl[1] = (q & 1) ? -h : h; wl[2] = -h; wl[3] = h;
  }
  for (int o = 0; o < w->O; ++o) {
    w->ob[o * 2] = 0.9 * wuni(w, idx, -h, h); w->ob[o * 2 + 1] = 0.9 * wuni(w, idx + 1, -h, h);
    idx += 2;
  }
  int saveW = w->W; /* spawn rejection ignores the boundary walls' 1.5*size band except via the 0.9 factor */
  int k = 0, tries = 0;
  while (k < w->N) {
    double p[2] = {0.9 * wuni(w, idx, -h, h), 0.9 * wuni(w, idx + 1, -h, h)};
    idx += 2;
    int bad = is_obstacle_collision(w, p, c->entity_size);
    for (int a = 0; a < k && !bad; ++a)
      if (norm2(w->as[a * 4] - p[0], w->as[a * 4 + 1] - p[1]) < c->separation_dist) bad = 1;
    if (bad && ++tries < PLACE_MAX_ATTEMPTS) continue;
    if (bad) w->overflow = 1;
    tries = 0;
    double* s = w->as + k * 4;
    s[0] = p[0]; s[1] = p[1];
    if (is_kin_heading(c)) { s[2] = wuni(w, idx++, 0, TWO_PI); s[3] = c->v_min; }
    else { s[2] = 0; s[3] = 0; }
    ++k;
  }
  k = 0; tries = 0;
  while (k < w->L) {
    double p[2] = {0.9 * wuni(w, idx, -h, h), 0.9 * wuni(w, idx + 1, -h, h)};
    idx += 2;
    int bad = is_obstacle_collision(w, p, c->entity_size);
    for (int a = 0; a < k && !bad; ++a) /* is_landmark_collision july:916-925 */
      if (norm2(w->lm[a * 2] - p[0], w->lm[a * 2 + 1] - p[1]) < 1.2 * (2 * c->entity_size)) bad = 1;
    if (bad && ++tries < PLACE_MAX_ATTEMPTS) continue;
    if (bad) w->overflow = 1;
    tries = 0;
    w->lm[k * 2] = p[0]; w->lm[k * 2 + 1] = p[1];
    ++k;
  }
  (void)saveW;
}

static void adopt_injected(World* w, const GmpeState* src) {
  World s;
  bind_world(&s, w->c, src, w->b);
  memcpy(w->as, s.as, sizeof(double) * w->N * 4);
  for (int i = 0; i < w->N; ++i) w->aux[i * 2 + 1] = s.aux[i * 2 + 1];
  memcpy(w->wf, s.wf, sizeof(double) * 12);
  memcpy(w->lm, s.lm, sizeof(double) * w->L * 2);
  if (w->O) memcpy(w->ob, s.ob, sizeof(double) * w->O * 2);
  if (w->W) memcpy(w->wl, s.wl, sizeof(double) * w->W * GMPE_WALL_F);
}

static void write_graph_outputs(World* w, const GmpeStepOut* out) {
  const GmpeConfig* c = w->c;
  int E = w->E, cap = E * (E - 1);
  if (out->adj) {
    double* adj = (double*)out->adj + w->b * E * E;
    for (int a = 0; a < E; ++a) for (int b = 0; b < E; ++b) adj[a * E + b] = w->dist[a][b];
  }
  if (out->edge_index && (c->flags & GMPE_FLAG_ENV_EDGES))
    emit_edges(w, c->coordination_range, 0, 0, out->edge_index + w->b * 2 * cap,
               (double*)out->edge_attr + w->b * cap, out->num_edges + w->b);
  if (out->gnn_edge_index && (c->flags & GMPE_FLAG_GNN_EDGES))
    emit_edges(w, c->gnn_max_edge_dist, 1, 1, out->gnn_edge_index + w->b * 2 * cap,
               (double*)out->gnn_edge_attr + w->b * cap, out->gnn_num_edges + w->b);
}

/* MultiAgentGraphEnv.reset env:1066-1081 -> reset_world july:339-420 */
static int rot_inv(const GmpeConfig* c) {
  return c->scenario == GMPE_SCEN_THREE_PHASE || c->scenario == GMPE_SCEN_TWO_PHASE || c->scenario == GMPE_SCEN_CORRIDOR_ROT_INV;
}
static int obs_dim_of(const GmpeConfig* c) {
  return c->scenario == GMPE_SCEN_JULY_TUBE ? 19 : (c->scenario == GMPE_SCEN_CORRIDOR_ROT_INV ? 13 : (rot_inv(c) ? 15 : 6));
}
static int node_dim_of(const GmpeConfig* c) { return rot_inv(c) ? 7 : 8; }
static void scen_observation(World* w, int i, double* o) {
  switch (w->c->scenario) {
    case GMPE_SCEN_JULY_TUBE: july_observation(w, i, o); break;
    case GMPE_SCEN_THREE_PHASE: case GMPE_SCEN_TWO_PHASE: case GMPE_SCEN_CORRIDOR_ROT_INV: tp_observation(w, i, o); break;
    default: nav_observation(w, i, o);
  }
}
static void scen_node_features(World* w, int i, double* out) {
  if (rot_inv(w->c)) tp_node_features(w, i, out); else node_features(w, i, out);
}
static void scen_mask_adj(World* w) {
  if (rot_inv(w->c)) tp_mask_adj(w); else mask_adj(w);
}

static void reset_world(World* w, const GmpeStepInject* inj, const GmpeStepOut* out) {
  const GmpeConfig* c = w->c;
  int odim = obs_dim_of(c), ndim = node_dim_of(c);
  reset_counters(w);
  w->draws = NULL; w->ndraws = 0;
  if (inj && inj->reset_draws && inj->reset_draws_per_world > 0) {
    w->draws = inj->reset_draws + w->b * inj->reset_draws_per_world;
    w->ndraws = inj->reset_draws_per_world;
  }
  if (inj && inj->reset_state) adopt_injected(w, inj->reset_state);
  else {
    if (c->scenario == GMPE_SCEN_JULY_TUBE) july_random_scenario(w);
    else if (rot_inv(c)) tp_random_scenario(w);
    else nav_random_scenario(w);
    for (int i = 0; i < w->N; ++i) /* min_time july:941-951 */
      w->aux[i * 2 + 1] = norm2(w->as[i * 4] - w->lm[i * 2], w->as[i * 4 + 1] - w->lm[i * 2 + 1]) / c->max_speed;
  }
  w->wi[GMPE_WI_EPISODE] += 1;
  calculate_distances(w); /* initialize_min_time_distance_graph july:735-739 */
  for (int i = 0; i < w->N; ++i) {
    if (out->obs) scen_observation(w, i, (double*)out->obs + (w->b * w->N + i) * odim);
    else if (c->scenario == GMPE_SCEN_JULY_TUBE) (void)get_agent_phase(w, i);
    else if (rot_inv(c)) (void)tp_get_agent_phase(w, i);
    if (out->node_obs) scen_node_features(w, i, (double*)out->node_obs + (w->b * w->N + i) * w->E * ndim);
  }
  scen_mask_adj(w);
  write_graph_outputs(w, out);
}

/* ------------------------------------------------------------------ step */
/* MultiAgentGraphEnv.step env:1021-1063 (+ graphworker auto-reset wrap:861-873) */
static void step_world(World* w, const int32_t* act, const GmpeStepInject* inj, const GmpeStepOut* out, int use_rk45) {
  const GmpeConfig* c = w->c;
  int N = w->N, E = w->E;
  int odim = obs_dim_of(c), ndim = node_dim_of(c);
  w->wi[GMPE_WI_STEP] += 1;
  for (int i = 0; i < N; ++i) set_action(w, i, act[w->b * N + i]);
  if (c->physics == GMPE_PHYS_FORCE) force_step(w); else kinematic_step(w, use_rk45);
  calculate_distances(w);
  double rew[MAXN];
  int all_done = 1;
  for (int i = 0; i < N; ++i) { /* env:1036-1053, agents in order */
    double otmp[19];
    double* o = out->obs ? (double*)out->obs + (w->b * N + i) * odim : otmp;
    scen_observation(w, i, o);
    double th = NAN;
    if (inj && inj->theta_draw) th = ((const double*)inj->theta_draw)[w->b * N + i];
    if (isnan(th))
      th = uni(c->seed, c->world_offset + w->b, (uint32_t)w->wi[GMPE_WI_EPISODE],
               (1u << 30) + (uint32_t)w->wi[GMPE_WI_STEP] * 64u + (uint32_t)i, 0, TWO_PI);
    rew[i] = c->scenario == GMPE_SCEN_JULY_TUBE ? july_reward(w, i, th)
             : (rot_inv(c) ? tp_reward(w, i, th) : nav_reward(w, i));
    if (out->node_obs) scen_node_features(w, i, (double*)out->node_obs + (w->b * N + i) * E * ndim);
    scen_mask_adj(w);
    int done = w->fl[i * 4 + GMPE_AF_STATUS] || w->wi[GMPE_WI_STEP] >= c->episode_length; /* env:264-271 */
    out->done[w->b * N + i] = (uint8_t)done;
    all_done &= done;
    if (c->flags & GMPE_FLAG_INFO) {
      double itmp[GMPE_INFO_K];
      double* info = out->info ? (double*)out->info + (w->b * N + i) * GMPE_INFO_K : itmp;
      july_info(w, i, info);
      /* info['Phase_reached'] of the rot-inv family (tp:806, tp2:804, sc:840): phase_reached[agent.id] as reward(i) left it */
      if (rot_inv(c) && out->phase_reached) out->phase_reached[w->b * N + i] = w->fl[i * 4 + GMPE_AF_PHASE_REACHED];
    }
  }
  double* r = (double*)out->reward + w->b * N;
  if (c->flags & GMPE_FLAG_COLLABORATIVE) { /* env:1056-1059 */
    double sum = 0;
    for (int i = 0; i < N; ++i) sum += rew[i];
    for (int i = 0; i < N; ++i) r[i] = sum;
  } else for (int i = 0; i < N; ++i) r[i] = rew[i];
  if (out->indiv_reward) for (int i = 0; i < N; ++i) ((double*)out->indiv_reward)[w->b * N + i] = rew[i]; /* env:1050 */
  /* GMPERunner.insert masks / active_masks (onpolicy/runner/shared/graph_mpe_runner.py:397-405) */
  for (int i = 0; i < N; ++i) {
    int d = out->done[w->b * N + i] != 0;
    if (out->mask) out->mask[w->b * N + i] = d ? 0.0f : 1.0f;
    if (out->active_mask) out->active_mask[w->b * N + i] = (all_done || !d) ? 1.0f : 0.0f;
  }
  write_graph_outputs(w, out);
  if ((c->flags & GMPE_FLAG_AUTO_RESET) && all_done) reset_world(w, inj, out); /* wrap:869-870 */
}

/* ------------------------------------------------------------------ batch entry points */
/* Worlds are independent: a plain pthread parallel-for over contiguous world ranges
 * (the image's gcc has no usable libgomp spec, so no OpenMP). */
#include <pthread.h>
typedef struct {
  const GmpeConfig* cfg; const GmpeState* st; const int32_t* act; const uint8_t* mask;
  const GmpeStepInject* inj; const GmpeStepOut* out; int use_rk45, is_reset;
  int64_t b0, b1;
} Job;

static long long g_overflows = 0; /* worlds whose placement gave up a rejection loop (gmpe_placement_overflows) */
long long gmpe_oracle_overflows(int clear) {
  long long v = __atomic_load_n(&g_overflows, __ATOMIC_RELAXED);
  if (clear) __atomic_store_n(&g_overflows, 0, __ATOMIC_RELAXED);
  return v;
}

static void* run_job(void* arg) {
  Job* j = (Job*)arg;
  World* w = (World*)malloc(sizeof(World));
  for (int64_t b = j->b0; b < j->b1; ++b) {
    if (j->is_reset && j->mask && !j->mask[b]) continue;
    bind_world(w, j->cfg, j->st, b);
    if (j->is_reset) reset_world(w, j->inj, j->out);
    else step_world(w, j->act, j->inj, j->out, j->use_rk45);
    if (w->overflow) __atomic_fetch_add(&g_overflows, 1, __ATOMIC_RELAXED);
  }
  free(w);
  return NULL;
}

static int run_parallel(Job proto, int nthreads) {
  int64_t B = proto.cfg->num_worlds;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if ((int64_t)nthreads > B) nthreads = (int)(B > 0 ? B : 1);
  Job jobs[256];
  pthread_t tid[256];
  int64_t chunk = (B + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = proto;
    jobs[t].b0 = t * chunk;
    jobs[t].b1 = (t + 1) * chunk < B ? (t + 1) * chunk : B;
  }
  for (int t = 1; t < nthreads; ++t) pthread_create(&tid[t], NULL, run_job, &jobs[t]);
  run_job(&jobs[0]);
  for (int t = 1; t < nthreads; ++t) pthread_join(tid[t], NULL);
  return GMPE_OK;
}

int gmpe_oracle_step(const GmpeConfig* cfg, const GmpeState* st, const int32_t* action_idx,
                     const GmpeStepInject* inj, const GmpeStepOut* out, int use_rk45, int nthreads) {
  if (!cfg || !st || !action_idx || !out || !out->reward || !out->done) return GMPE_E_NULL;
  if (cfg->num_agents > MAXN || cfg->num_agents + cfg->num_landmarks + cfg->num_obstacles + cfg->num_walls > MAXE)
    return GMPE_E_INVALID;
  Job j = {cfg, st, action_idx, NULL, inj, out, use_rk45, 0, 0, 0};
  return run_parallel(j, nthreads);
}

int gmpe_oracle_reset(const GmpeConfig* cfg, const GmpeState* st, const uint8_t* env_mask,
                      const GmpeStepInject* inj, const GmpeStepOut* out, int nthreads) {
  if (!cfg || !st || !out) return GMPE_E_NULL;
  Job j = {cfg, st, NULL, env_mask, inj, out, 0, 1, 0, 0};
  return run_parallel(j, nthreads);
}

/* argmax of a one-hot row (env:431, 444), first maximum wins like np.argmax */
int gmpe_oracle_decode_onehot(const GmpeConfig* cfg, const double* onehot, int32_t* action_idx) {
  int64_t rows = (int64_t)cfg->num_worlds * cfg->num_agents;
  int A = cfg->total_actions;
  for (int64_t r = 0; r < rows; ++r) {
    int best = 0;
    for (int a = 1; a < A; ++a) if (onehot[r * A + a] > onehot[r * A + best]) best = a;
    action_idx[r] = best;
  }
  return GMPE_OK;
}
