// gmpe_api.cu -- C ABI (include/gmpe.h) over the fused step kernel.
// Built by csrc/build.py:  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <sched.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <utility>
#include <algorithm>

#include "gmpe_coop.cuh"
#include "gmpe_coop_table.h"
#include "gmpe_step.cuh"
#include "gmpe_flat.cuh"

using namespace gmpe;

// Staging of a pageable action block into page-locked memory (gmpe_step_host) by a few helper
// threads plus the caller: a single-threaded memcpy moves ~10 GB/s, which would put 0.8 ms of copy
// in front of a 1.3 ms step for the runner's 7.9 MB float64 one-hot block.  The copy is cut into
// 256 KB pieces handed out IN ORDER through an atomic counter, and every piece raises a flag when it
// is done: the caller starts the job, then for each upload block of the step waits (copying pieces
// itself meanwhile) only until the pieces of THAT block are in place -- the upload of block k and
// its kernels run while the helpers are still staging block k+1.  The helpers are created on the
// first host step that needs staging; between steps they poll for a few milliseconds, then sleep.
class StagePool {
 public:
  enum { PIECE = 256 * 1024, MAX_PIECES = 4096 };
  explicit StagePool(int helpers) : done_(new std::atomic<uint8_t>[MAX_PIECES]) {
    for (int t = 0; t < helpers; ++t) th_.emplace_back([this] { run(); });
  }
  ~StagePool() {
    {
      std::lock_guard<std::mutex> g(m_);
      quit_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
    delete[] done_;
  }
  // dst[0, n) = src[0, n), asynchronously; wait_until(bytes) / finish() below
  void start(char* dst, const char* src, size_t n) {
    {  // a helper that woke late for the previous job has left it before the job variables change
      std::unique_lock<std::mutex> g(m_);
      idle_.wait(g, [this] { return active_ == 0; });
    }
    npieces_ = (n + PIECE - 1) / PIECE;
    if (npieces_ > MAX_PIECES) npieces_ = 0;  // (never at these sizes) -> the caller copies everything in wait_until
    dst_ = dst; src_ = src; n_ = n;
    for (size_t q = 0; q < npieces_; ++q) done_[q].store(0, std::memory_order_relaxed);
    next_.store(0, std::memory_order_relaxed);
    if (npieces_ > 1 && !th_.empty()) {
      {
        std::lock_guard<std::mutex> g(m_);
        active_ = (int)th_.size();
        ++gen_;
        gen_atomic_.store(gen_, std::memory_order_release);
      }
      cv_.notify_all();
    }
    waited_ = 0;
  }
  // returns once dst[0, upto) is complete
  void wait_until(size_t upto) {
    if (npieces_ == 0) {  // job too large for the flag table: plain copy of the missing part
      if (upto > waited_) { memcpy(dst_ + waited_, src_ + waited_, upto - waited_); waited_ = upto; }
      return;
    }
    const size_t last = (upto + PIECE - 1) / PIECE;  // pieces [0, last) are needed
    for (size_t q = waited_; q < last; ++q) {
      while (!done_[q].load(std::memory_order_acquire)) {
        if (!copy_one()) std::this_thread::yield();
      }
    }
    if (last > waited_) waited_ = last;
  }
  void finish() { wait_until(n_); }  // every piece is in place: no helper touches the buffers any more

 private:
  bool copy_one() {
    const size_t q = next_.fetch_add(1, std::memory_order_relaxed);
    if (q >= npieces_) return false;
    const size_t lo = q * (size_t)PIECE;
    memcpy(dst_ + lo, src_ + lo, std::min((size_t)PIECE, n_ - lo));
    done_[q].store(1, std::memory_order_release);
    return true;
  }
  void run() {
    uint64_t seen = 0;
    for (;;) {
      // A rollout steps every ~1.3 ms: between steps a helper polls the generation counter for a few
      // milliseconds before it goes to sleep -- waking a sleeping thread costs 0.1 - 1 ms on these boxes,
      // more than the whole staging copy (OpenMP runtimes spin between parallel regions for the same reason).
      bool got = false;
      const auto t0 = std::chrono::steady_clock::now();
      for (int spin = 0;; ++spin) {
        if (gen_atomic_.load(std::memory_order_acquire) != seen) { got = true; break; }
        if ((spin & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(4)) break;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      if (!got) {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return quit_ || gen_ != seen; });
        if (quit_) return;
      }
      {
        std::lock_guard<std::mutex> g(m_);
        if (quit_) return;
        seen = gen_;
      }
      while (copy_one()) {}
      {
        std::lock_guard<std::mutex> g(m_);
        if (--active_ == 0) idle_.notify_one();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, idle_;
  std::atomic<uint8_t>* done_;
  std::atomic<size_t> next_{0};
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t n_ = 0, npieces_ = 0, waited_ = 0;
  uint64_t gen_ = 0;
  std::atomic<uint64_t> gen_atomic_{0};
  int active_ = 0;
  bool quit_ = false;
};

struct gmpe_ctx {
  GmpeConfig cfg;
  int device;
  int sm_count;
  int threads;
  int wpb;
  int coop_warps;
  int num_tiles;
  int grid;
  size_t smem_bytes;
  int stage_graph;
  int small;  // KERNEL_TILE (gmpe_step.cuh) / KERNEL_COOP (gmpe_coop.cuh)
  const void* fn;
  int64_t resident;  // CTAs of the step kernel that fit on the device at once
  unsigned char act_tab_host[512];  // the same table for the kernel parameter block (KParams::act_tab_c)
  void* act_tab;  // device copy of the _set_action table (Real [32][2]) followed by the placement-overflow counter
                  // (one 64-bit word at byte ACT_TAB_BYTES): the only device memory the library owns
  int64_t launches;
  std::string err;
  // gmpe_step_host: copy streams and per-slice events, created on first use
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev;
  // tail of the last host step's device->host copies: a later launch on ANY stream waits for it
  // before it overwrites the device outputs those copies are reading
  cudaEvent_t ev_tail = nullptr;
  bool tail_pending = false;
  StagePool* pool = nullptr;
#ifdef GMPE_EXPERIMENTS
  // GMPE_TRACE=1 (tools/e2e_sweep.py --trace): timeline of one host step, printed by gmpe_host_wait
  std::vector<cudaEvent_t> tr_ev;   // [0] call start, then per slice: H2D done, kernels done, D2H done
  std::vector<double> tr_host;      // per slice: staged, enqueued (us since the call); last: returned
  std::vector<std::pair<int64_t, int64_t>> tr_sl;
  bool tr_on = false;
#endif
};

#ifdef GMPE_EXPERIMENTS
#include <chrono>
static double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#endif

static thread_local std::string g_create_err;

static int fail(gmpe_ctx* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}

static bool rot_inv(const GmpeConfig& c) {
  return c.scenario == GMPE_SCEN_THREE_PHASE || c.scenario == GMPE_SCEN_TWO_PHASE || c.scenario == GMPE_SCEN_CORRIDOR_ROT_INV;
}
static int obs_dim_of(const GmpeConfig& c) {
  return c.scenario == GMPE_SCEN_JULY_TUBE ? 19 : (c.scenario == GMPE_SCEN_CORRIDOR_ROT_INV ? 13 : (rot_inv(c) ? 15 : 6));
}
static int node_dim_of(const GmpeConfig& c) { return rot_inv(c) ? 7 : 8; }

enum { ACT_TAB_BYTES = 512 };  // Real [32][2] at fp64; the overflow counter sits right behind it

// Tuning knobs read from the environment exist only in GMPE_EXPERIMENTS builds (tools/sweep.py,
// tools/ab.py): the product library's behaviour depends on its arguments alone.
static const char* tuning_env(const char* name) {
#ifdef GMPE_EXPERIMENTS
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

static const char* validate(const GmpeConfig& c) {
  if (c.scenario != GMPE_SCEN_JULY_TUBE && c.scenario != GMPE_SCEN_NAV_GRAPH && !rot_inv(c)) return "unknown scenario";
  if (c.dynamics < 0 || c.dynamics > 2) return "unknown dynamics";
  if (c.physics != GMPE_PHYS_KINEMATIC && c.physics != GMPE_PHYS_FORCE) return "unknown physics";
  if (c.precision != GMPE_F32 && c.precision != GMPE_F64) return "unknown precision";
  if (c.num_worlds < 1) return "num_worlds must be >= 1";
  if (c.num_agents < 1 || c.num_agents > GMPE_MAX_AGENTS) return "num_agents out of range [1, 64]";
  if (c.num_landmarks < c.num_agents) return "num_landmarks must be >= num_agents";
  if (c.num_obstacles < 0 || c.num_walls < 0) return "negative entity count";
  if (c.num_agents + c.num_landmarks + c.num_obstacles + c.num_walls > GMPE_MAX_ENTITIES) return "too many entities";
  if (rot_inv(c) && !(c.num_agents >= COOP_MIN_AGENTS && c.num_agents <= COOP_MAX_AGENTS_ROT_INV &&
                      c.num_landmarks == c.num_agents && c.num_obstacles == 0 && c.num_walls == 0 && c.physics == GMPE_PHYS_KINEMATIC))
    return "the rot-inv scenarios are served for 2 <= num_agents == num_landmarks <= 16 without obstacles / walls, kinematic step";
  if (c.scenario == GMPE_SCEN_JULY_TUBE && (c.num_obstacles || c.num_walls))
    return "july tube scenario supports num_obstacles == num_walls == 0 only (the reference graph builder raises, SURVEY G7)";
  if (c.physics == GMPE_PHYS_FORCE && c.dynamics != GMPE_DYN_DOUBLE_INTEGRATOR) return "force physics needs double_integrator state";
  if (c.dynamics == GMPE_DYN_DOUBLE_INTEGRATOR && c.total_actions != 5 && c.total_actions != 9) return "double_integrator: total_actions must be 5 or 9";
  if (c.dynamics != GMPE_DYN_DOUBLE_INTEGRATOR && c.total_actions != 25) return "unicycle / air_taxi use the 25 motion primitives";
  if (c.episode_length < 1) return "episode_length must be >= 1";
  if (!(c.world_size > 0)) return "world_size must be positive";
  if (c.formation < GMPE_FORMATION_POINT || c.formation > GMPE_FORMATION_CIRCLE) return "unknown formation";
  if (c.reserved0 != 0) return "GmpeConfig.reserved0 must be 0";
  {
    // Spawn placement is rejection sampling on the device (bounded, see gmpe_placement_overflows):
    // refuse what cannot be packed.  Agents keep separation_dist between each other inside the
    // spawn box -- 0.9 * world_size square (navigation), jitter box of side 2 * 0.2 (july) or
    // 2 * 0.3 (rot-inv) * world_size around spawn points staggered by 1/5 (july:470) or 1/3 (tp:430)
    // per agent along one axis.
    const double side = c.scenario == GMPE_SCEN_NAV_GRAPH ? 0.9 * c.world_size : (rot_inv(c) ? 0.6 : 0.4) * c.world_size;
    const double length = c.scenario == GMPE_SCEN_NAV_GRAPH ? side : side + (c.num_agents - 1) / (rot_inv(c) ? 3.0 : 5.0);
    const double need = c.num_agents * 3.141592653589793 * 0.25 * c.separation_dist * c.separation_dist +
                        c.num_obstacles * 3.141592653589793 * 16.0 * c.entity_size * c.entity_size;
    if (need > 0.5 * side * length) return "agents cannot be packed: N * pi * (separation_dist / 2)^2 exceeds half of the spawn area (raise world_size)";
  }
  return nullptr;
}

static int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

// Worlds per tile: as many as the CTA has (world, agent) threads for, rounded down to the
// multiple that keeps every tile's output slab 16-byte aligned (obs N*F, adj E*E, edges
// E*(E-1) elements per world) when that is possible; copy_out falls back to scalar stores
// otherwise.
static void plan_tiles(const GmpeConfig& c, int threads, int* wpb, int* tiles) {
  const int E = c.num_agents + c.num_landmarks + c.num_obstacles + c.num_walls;
  const int F = obs_dim_of(c);
  const int vw = c.precision == GMPE_F64 ? 2 : 4;  // elements per 16 bytes
  int m = 1;
  const int counts[3] = {c.num_agents * F, E * E, E * (E - 1)};
  for (int k = 0; k < 3; ++k) {
    int need = vw / gcd_i(vw, counts[k]);
    m = m * need / gcd_i(m, need);
  }
  int w = threads / c.num_agents;
  if (w < 1) w = 1;
  if (w >= m) w -= w % m;
  *wpb = w;
  *tiles = (c.num_worlds + w - 1) / w;
}

template <typename Real>
static size_t smem_for(const GmpeConfig& c, int wpb, int stage_graph) {
  int E = c.num_agents + c.num_landmarks + c.num_obstacles + c.num_walls;
  int F = obs_dim_of(c);
  return smem_layout<Real>(wpb, c.num_agents, E, c.num_walls, F, stage_graph, nullptr);
}

// Small worlds stage the adjacency matrix and the padded edge buffers of a tile in shared
// memory (thread-per-world compaction, 128-bit stores); large worlds generate in place.
static int pick_stage_graph(const GmpeConfig& c, int wpb) {
  int E = c.num_agents + c.num_landmarks + c.num_obstacles + c.num_walls;
  if (E > 12) return 0;
  int mode = (c.flags & GMPE_FLAG_GNN_EDGES) ? 3 : 1;
  size_t need = c.precision == GMPE_F64 ? smem_for<double>(c, wpb, mode) : smem_for<float>(c, wpb, mode);
  return need <= 72 * 1024 ? mode : 0;
}

// Worlds with 2 <= N = L <= 10 and no obstacles / walls (BASELINE configs 1-3 among them) run on the
// lane-per-agent kernel (gmpe_coop.cuh).  GMPE_FLAG_FORCE_TILE in the configuration overrides
// the choice (A/B timing, and the tests that keep both mappings pinned to the oracle).
enum { KERNEL_TILE = 0, KERNEL_COOP = 2 };
static int pick_kernel(const GmpeConfig& c) {
  if (rot_inv(c)) return KERNEL_COOP;  // only the lane-per-agent kernel serves that family (2 <= N = L <= 16)
  const bool small = c.num_agents >= COOP_MIN_AGENTS && c.num_agents <= COOP_MAX_AGENTS && c.num_landmarks == c.num_agents &&
                     c.num_obstacles == 0 && c.num_walls == 0;
  if (!small) return KERNEL_TILE;
  if (c.flags & GMPE_FLAG_FORCE_TILE) return KERNEL_TILE;
  return KERNEL_COOP;
}

// Persistent grid: at most one wave of resident CTAs, units strided over it.  When the batch
// needs r > 1 rounds the grid is shrunk to ceil(units / r) CTAs so that every CTA runs exactly
// r rounds (a last round that only a few CTAs take part in leaves the SMs idle).
static int grid_for(const gmpe_ctx* h, int64_t units) {
  const int64_t g = h->resident;
  bool persistent = true;
  if (const char* t = tuning_env("GMPE_PERSISTENT")) persistent = atoi(t) != 0;  // tuning knob (tools/ab.py)
  if (units <= g || !persistent) return (int)units;
  if (const char* t = tuning_env("GMPE_GRID")) { const int v = atoi(t); if (v > 0 && v <= units) return v; }  // tuning knob
  // lane-per-agent kernel: every SM gets the same number of CTAs, the partial last round is
  // spread over all of them at warp granularity (gmpe_coop.cuh); 30.8 -> 30.1 us at 65,536
  // navigation worlds, 60.9 -> 59.8 us at 131,072 july worlds against the rounded-down grid below
  if (h->small == KERNEL_COOP && GMPE_COOP_TRANSPOSE_TAIL) return (int)g;
  const int64_t rounds = (units + g - 1) / g;
  if (const char* t = tuning_env("GMPE_BALANCED")) if (!atoi(t)) return (int)g;
  return (int)((units + rounds - 1) / rounds);
}

template <typename T>
static void shift(T*& ptr, int64_t elems) { if (ptr) ptr += elems; }
static void shift_bytes(void*& ptr, int64_t bytes) { if (ptr) ptr = (char*)ptr + bytes; }
static void shift_bytes(const void*& ptr, int64_t bytes) { if (ptr) ptr = (const char*)ptr + bytes; }

// Re-base every world-major array on world `b0` (gmpe_step_range / gmpe_reset_range).
static void rebase_state(GmpeState& s, const GmpeConfig& c, int64_t b0, int64_t R) {
  const int64_t N = c.num_agents;
  shift_bytes(s.agent_state, b0 * N * 4 * R);
  shift_bytes(s.agent_aux, b0 * N * 2 * R);
  shift(s.agent_flags, b0 * N * 4);
  shift(s.agent_cnt, b0 * N * GMPE_AGENT_CNT);
  shift_bytes(s.world_f, b0 * GMPE_WORLD_F * R);
  shift(s.world_i, b0 * GMPE_WORLD_I);
  shift_bytes(s.landmarks, b0 * c.num_landmarks * 2 * R);
  shift_bytes(s.obstacles, b0 * c.num_obstacles * 2 * R);
  shift_bytes(s.walls, b0 * c.num_walls * GMPE_WALL_F * R);
  shift(s.agent_ext, b0 * N * 2);
}

static void rebase_out(GmpeStepOut& o, const GmpeConfig& c, int64_t b0, int64_t R) {
  const int64_t N = c.num_agents, E = N + c.num_landmarks + c.num_obstacles + c.num_walls, cap = E * (E - 1);
  const int64_t F = obs_dim_of(c);
  shift_bytes(o.obs, b0 * N * F * R);
  shift_bytes(o.node_obs, b0 * N * E * node_dim_of(c) * R);
  shift_bytes(o.adj, b0 * E * E * R);
  shift_bytes(o.reward, b0 * N * R);
  shift(o.done, b0 * N);
  shift_bytes(o.info, b0 * N * GMPE_INFO_K * R);
  shift(o.edge_index, b0 * 2 * cap);
  shift_bytes(o.edge_attr, b0 * cap * R);
  shift(o.num_edges, b0);
  shift(o.gnn_edge_index, b0 * 2 * cap);
  shift_bytes(o.gnn_edge_attr, b0 * cap * R);
  shift(o.gnn_num_edges, b0);
  shift_bytes(o.indiv_reward, b0 * N * R);
  shift(o.mask, b0 * N);
  shift(o.active_mask, b0 * N);
  shift(o.phase_reached, b0 * N);
}

extern "C" {

int gmpe_abi_version(void) { return GMPE_ABI_VERSION; }

int gmpe_layout(const GmpeConfig* cfg, GmpeLayout* out) {
  if (!cfg || !out) return GMPE_E_NULL;
  if (validate(*cfg)) return GMPE_E_INVALID;
  const int64_t B = cfg->num_worlds, N = cfg->num_agents, L = cfg->num_landmarks, O = cfg->num_obstacles, W = cfg->num_walls;
  const int64_t E = N + L + O + W;
  out->real_bytes = cfg->precision == GMPE_F64 ? 8 : 4;
  out->num_entities = (int32_t)E;
  out->obs_dim = obs_dim_of(*cfg);
  out->node_dim = node_dim_of(*cfg);
  out->edge_cap = (int32_t)(E * (E - 1));
  out->info_dim = GMPE_INFO_K;
  out->agent_state = B * N * 4;
  out->agent_aux = B * N * 2;
  out->agent_flags = B * N * 4;
  out->agent_cnt = B * N * GMPE_AGENT_CNT;
  out->world_f = B * GMPE_WORLD_F;
  out->world_i = B * GMPE_WORLD_I;
  out->landmarks = B * L * 2;
  out->obstacles = B * O * 2;
  out->walls = B * W * GMPE_WALL_F;
  out->agent_ext = B * N * 2;
  out->obs = B * N * out->obs_dim;
  out->node_obs = B * N * E * out->node_dim;
  out->adj = B * E * E;
  out->reward = B * N;
  out->done = B * N;
  out->info = B * N * GMPE_INFO_K;
  out->edge_index = B * 2 * out->edge_cap;
  out->edge_attr = B * out->edge_cap;
  out->num_edges = B;
  return GMPE_OK;
}

int gmpe_create(const GmpeConfig* cfg, int device, gmpe_handle* out) {
  if (!cfg || !out) return fail(nullptr, GMPE_E_NULL, "gmpe_create: NULL argument");
  if (const char* why = validate(*cfg)) return fail(nullptr, GMPE_E_INVALID, std::string("gmpe_create: ") + why);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev)
    return fail(nullptr, GMPE_E_CUDA, std::string("gmpe_create: no such CUDA device (") + cudaGetErrorString(e) + ")");
  gmpe_ctx* h = new (std::nothrow) gmpe_ctx();
  if (!h) return fail(nullptr, GMPE_E_INVALID, "gmpe_create: out of host memory");
  h->cfg = *cfg;
  h->device = device;
  h->launches = 0;
  h->act_tab = nullptr;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  const bool f64 = cfg->precision == GMPE_F64;
  h->small = pick_kernel(h->cfg);
  int units = 0;  // CTAs needed to cover the batch once
  if (h->small == KERNEL_COOP) {
    // one lane per agent, floor(32 / N) worlds per warp, warps independent: gmpe_coop.cuh
    h->stage_graph = 0;
    h->coop_warps = 0;  // warps per CTA of the instantiation (4; fewer for the widest worlds, whose slabs are large)
    const bool found = f64 ? coop_lookup_f64(cfg->scenario, cfg->num_agents, &h->fn, &h->smem_bytes, &h->coop_warps)
                           : coop_lookup_f32(cfg->scenario, cfg->num_agents, &h->fn, &h->smem_bytes, &h->coop_warps);
    h->threads = 32 * h->coop_warps;
    h->wpb = h->coop_warps * (32 / cfg->num_agents);
    h->num_tiles = found ? (cfg->num_worlds + h->wpb - 1) / h->wpb : 0;
    if (!found) {
      cudaSetDevice(prev);
      const int code = fail(nullptr, GMPE_E_INVALID, "no lane-per-agent kernel for this scenario / num_agents");
      delete h;
      return code;
    }
  } else {
    h->threads = GMPE_LB_THREADS;
    if (const char* t = tuning_env("GMPE_THREADS")) {  // tuning knob (tools/sweep.py); must not exceed the launch bound
      int v = atoi(t);
      if (v >= 32 && v <= GMPE_LB_THREADS && v % 32 == 0 && v >= cfg->num_agents) h->threads = v;  // phase A needs a thread per agent of a world
    }
    plan_tiles(h->cfg, h->threads, &h->wpb, &h->num_tiles);
    h->stage_graph = pick_stage_graph(h->cfg, h->wpb);
    h->smem_bytes = f64 ? smem_for<double>(h->cfg, h->wpb, h->stage_graph) : smem_for<float>(h->cfg, h->wpb, h->stage_graph);
    h->fn = f64 ? (const void*)gmpe_step_kernel<double> : (const void*)gmpe_step_kernel<float>;
    // fp32 instantiations with compile-time entity counts for the BASELINE shapes
    const GmpeConfig& q = h->cfg;
    const char* gs = tuning_env("GMPE_GENERIC_SHAPE");
    const bool generic = gs && atoi(gs);
    if (!f64 && !generic) {
      if (q.num_agents == 16 && q.num_landmarks == 16 && q.num_obstacles == 8 && q.num_walls == 4 &&
          q.scenario == GMPE_SCEN_NAV_GRAPH && q.dynamics == GMPE_DYN_DOUBLE_INTEGRATOR)
        h->fn = (const void*)gmpe_step_kernel<float, 16, 16, 8, 4, GMPE_SCEN_NAV_GRAPH>;
      else if (q.num_agents == 64 && q.num_landmarks == 64 && q.num_obstacles == 0 && q.num_walls == 0 &&
               q.scenario == GMPE_SCEN_JULY_TUBE)
        h->fn = (const void*)gmpe_step_kernel<float, 64, 64, 0, 0, GMPE_SCEN_JULY_TUBE>;
    }
  }
  units = h->num_tiles;
  const void* fn = h->fn;
  int max_optin = 0;
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (h->smem_bytes > (size_t)max_optin) {
    std::string m = "gmpe_create: tile needs " + std::to_string(h->smem_bytes) + " B of shared memory, device allows " + std::to_string(max_optin);
    delete h;
    cudaSetDevice(prev);
    return fail(nullptr, GMPE_E_INVALID, m);
  }
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
  int per_sm = 1;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, h->threads, h->smem_bytes);
  if (e != cudaSuccess) {
    std::string m = std::string("gmpe_create: ") + cudaGetErrorString(e);
    delete h;
    cudaSetDevice(prev);
    return fail(nullptr, GMPE_E_CUDA, m);
  }
  if (per_sm < 1) per_sm = 1;
  {  // _set_action (environment.py:336-475) evaluated once per action index
    double td[64];
    float tf[64];
    for (int a = 0; a < 32; ++a) {
      double u0 = 0.0, u1 = 0.0;
      if (a < cfg->total_actions) decode_action<double>(h->cfg, a, u0, u1);
      td[2 * a] = u0; td[2 * a + 1] = u1;
      float f0 = 0.f, f1 = 0.f;
      if (a < cfg->total_actions) decode_action<float>(h->cfg, a, f0, f1);
      tf[2 * a] = f0; tf[2 * a + 1] = f1;
    }
    const size_t nbytes = f64 ? sizeof(td) : sizeof(tf);
    memset(h->act_tab_host, 0, sizeof(h->act_tab_host));
    memcpy(h->act_tab_host, f64 ? (const void*)td : (const void*)tf, nbytes);
    static_assert(sizeof(td) <= ACT_TAB_BYTES, "action table");
    e = cudaMalloc(&h->act_tab, ACT_TAB_BYTES + 8);
    if (e == cudaSuccess) e = cudaMemset(h->act_tab, 0, ACT_TAB_BYTES + 8);
    if (e == cudaSuccess) e = cudaMemcpy(h->act_tab, f64 ? (const void*)td : (const void*)tf, nbytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      std::string m = std::string("gmpe_create: ") + cudaGetErrorString(e);
      if (h->act_tab) cudaFree(h->act_tab);
      delete h;
      cudaSetDevice(prev);
      return fail(nullptr, GMPE_E_CUDA, m);
    }
  }
  // persistent grid: one wave of resident CTAs, tiles strided over it (148 SMs on B200)
  int64_t g = (int64_t)h->sm_count * per_sm;
  h->resident = g;
  h->grid = grid_for(h, units);
  cudaSetDevice(prev);
  *out = h;
  return GMPE_OK;
}

int gmpe_destroy(gmpe_handle h) {
  if (!h) return GMPE_E_NULL;
  {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    if (h->act_tab) cudaFree(h->act_tab);
    delete h->pool;
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->ev_tail) cudaEventDestroy(h->ev_tail);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    cudaSetDevice(prev);
  }
  delete h;
  return GMPE_OK;
}

const char* gmpe_last_error(gmpe_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int64_t gmpe_launch_count(gmpe_handle h) { return h ? h->launches : 0; }

const char* gmpe_kernel_name(gmpe_handle h) {
  if (!h) return "";
  return h->small == KERNEL_COOP ? "gmpe_step_coop_kernel" : "gmpe_step_kernel";
}

int gmpe_placement_overflows(gmpe_handle h, int64_t* count, int clear) {
  if (!h || !count) return GMPE_E_NULL;
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  unsigned long long v = 0;
  void* cnt = (char*)h->act_tab + ACT_TAB_BYTES;
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(&v, cnt, 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && clear) e = cudaMemset(cnt, 0, 8);
  if (prev != h->device) cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(h, GMPE_E_CUDA, std::string("gmpe_placement_overflows: ") + cudaGetErrorString(e));
  *count = (int64_t)v;
  return GMPE_OK;
}

int gmpe_set_seed(gmpe_handle h, uint64_t seed, int64_t world_offset) {
  if (!h) return GMPE_E_NULL;
  h->cfg.seed = seed;
  h->cfg.world_offset = world_offset;
  return GMPE_OK;
}

static int launch(gmpe_handle h, int mode, const GmpeState* st, const int32_t* act, const uint8_t* mask,
                  const GmpeStepInject* inj, const GmpeStepOut* out, void* stream, int64_t b0 = 0, int64_t count = -1,
                  bool in_host_step = false) {
  const GmpeConfig& c = h->cfg;
  if (count < 0) count = c.num_worlds;
  if (b0 < 0 || count < 1 || b0 + count > c.num_worlds) return fail(h, GMPE_E_INVALID, "world range outside [0, num_worlds)");
  if (!st || !out) return fail(h, GMPE_E_NULL, "state / out struct is NULL");
  if (!st->agent_state || !st->agent_aux || !st->agent_flags || !st->agent_cnt || !st->world_f || !st->world_i || !st->landmarks)
    return fail(h, GMPE_E_NULL, "a required state array is NULL");
  if ((c.num_obstacles && !st->obstacles) || (c.num_walls && !st->walls)) return fail(h, GMPE_E_NULL, "obstacles / walls array is NULL");
  if (rot_inv(c) && !st->agent_ext) return fail(h, GMPE_E_NULL, "the rot-inv scenarios need state.agent_ext");
  if (mode == MODE_STEP && (!act || !out->reward || !out->done)) return fail(h, GMPE_E_NULL, "gmpe_step needs action_idx, out.reward and out.done");
  if ((c.flags & GMPE_FLAG_ENV_EDGES) && out->edge_index && (!out->edge_attr || !out->num_edges)) return fail(h, GMPE_E_NULL, "edge_attr / num_edges missing");
  if ((c.flags & GMPE_FLAG_GNN_EDGES) && out->gnn_edge_index && (!out->gnn_edge_attr || !out->gnn_num_edges)) return fail(h, GMPE_E_NULL, "gnn_edge_attr / gnn_num_edges missing");
  // feature rows leave the SM as 256-bit stores (store_row8): torch / cudaMalloc buffers are
  // 256-byte aligned, an arbitrary caller pointer must at least be 32-byte aligned
  if ((node_dim_of(c) == 8 && (reinterpret_cast<size_t>(out->node_obs) & 31)) || (reinterpret_cast<size_t>(out->info) & 31))
    return fail(h, GMPE_E_INVALID, "out.node_obs and out.info must be 32-byte aligned");
  KParams p;
  memset(&p, 0, sizeof(p));
  p.c = c;
  p.st = *st;
  GmpeStepOut o = *out;
  p.act = act;
  p.env_mask = mask;
  p.theta = inj ? inj->theta_draw : nullptr;
  p.has_rst = (inj && inj->reset_state) ? 1 : 0;
  if (p.has_rst) p.rst = *inj->reset_state;
  if (inj && inj->reset_draws && inj->reset_draws_per_world > 0) {
    p.reset_draws = inj->reset_draws;
    p.reset_draws_n = inj->reset_draws_per_world;
  }
  p.overflow_cnt = reinterpret_cast<unsigned long long*>((char*)h->act_tab + ACT_TAB_BYTES);
  p.mode = mode;
  p.wpb = h->wpb;
  p.num_tiles = h->num_tiles;
  int grid = h->grid;
  if (b0 != 0 || count != c.num_worlds) {  // a slice of the batch: same arrays, re-based on world b0
    const int64_t R = c.precision == GMPE_F64 ? 8 : 4;
    rebase_state(p.st, c, b0, R);
    rebase_out(o, c, b0, R);
    if (p.has_rst) rebase_state(p.rst, c, b0, R);
    shift(p.act, b0 * c.num_agents);
    shift(p.env_mask, b0);
    shift_bytes(p.theta, b0 * c.num_agents * R);
    shift(p.reset_draws, b0 * p.reset_draws_n);
    p.c.num_worlds = (int32_t)count;
    p.c.world_offset += b0;  // Philox streams are keyed by the global world id
    p.num_tiles = (int)((count + h->wpb - 1) / h->wpb);
    grid = grid_for(h, p.num_tiles);
  }
  memcpy(&p.out, &o, sizeof(p.out));  // the leading fifteen pointers (KOut) ...
  p.out_phase_reached = o.phase_reached;  // ... and the sixteenth, kept at the end of the parameter block
  p.stage_graph = h->stage_graph;
  memcpy(p.act_tab_c, h->act_tab_host, sizeof(p.act_tab_c));
  if (h->small == KERNEL_COOP) {  // round structure of the persistent grid (num_tiles counts CTAs of coop_warps warps)
    const int wpw = 32 / c.num_agents;
    p.coop_groups = (int)((p.c.num_worlds + wpw - 1) / wpw);
    p.coop_full_rounds = p.coop_groups / (grid * h->coop_warps);
  }
  const int N = c.num_agents, L = c.num_landmarks, O = c.num_obstacles, W = c.num_walls;
  p.E = N + L + O + W;
  p.F = obs_dim_of(c);
  p.dN = FastDiv::make(N); p.dE = FastDiv::make(p.E); p.dL = FastDiv::make(L); p.dO = FastDiv::make(O);
  p.dW = FastDiv::make(W); p.dF = FastDiv::make(p.F); p.d16 = FastDiv::make(16);
  if (const char* t = tuning_env("GMPE_XMODE")) p.xmode = atoi(t);
  p.act_tab = h->act_tab;
  p.dE2 = FastDiv::make(p.E * 2); p.dNE2 = FastDiv::make(N * p.E * 2); p.dEE = FastDiv::make(p.E * p.E);
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  void* kargs[1] = {(void*)&p};
  cudaError_t e = cudaSuccess;
  // the copies of an uncollected host step still read the device outputs this launch rewrites
  if (h->tail_pending && !in_host_step) e = cudaStreamWaitEvent(s, h->ev_tail, 0);
  if (e == cudaSuccess) e = cudaLaunchKernel(h->fn, dim3(grid), dim3(h->threads), kargs, h->smem_bytes, s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (prev != h->device) cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(h, GMPE_E_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString(e));
  h->launches += 1;
  return GMPE_OK;
}

int gmpe_reset(gmpe_handle h, const GmpeState* st, const uint8_t* env_mask, const GmpeStepInject* inject,
               const GmpeStepOut* out, void* cuda_stream) {
  if (!h) return GMPE_E_NULL;
  return launch(h, MODE_RESET, st, nullptr, env_mask, inject, out, cuda_stream);
}

int gmpe_step(gmpe_handle h, const GmpeState* st, const int32_t* action_idx, const GmpeStepInject* inject,
              const GmpeStepOut* out, void* cuda_stream) {
  if (!h) return GMPE_E_NULL;
  return launch(h, MODE_STEP, st, action_idx, nullptr, inject, out, cuda_stream);
}

int gmpe_step_range(gmpe_handle h, const GmpeState* st, const int32_t* action_idx, const GmpeStepInject* inject,
                    const GmpeStepOut* out, int64_t world_begin, int64_t world_count, void* cuda_stream) {
  if (!h) return GMPE_E_NULL;
  return launch(h, MODE_STEP, st, action_idx, nullptr, inject, out, cuda_stream, world_begin, world_count);
}

int gmpe_decode_onehot(gmpe_handle h, const void* onehot, int elem_bytes, int32_t* action_idx, void* cuda_stream) {
  if (!h) return GMPE_E_NULL;
  return gmpe_decode_onehot_range(h, onehot, elem_bytes, action_idx, 0, h->cfg.num_worlds, cuda_stream);
}

int gmpe_decode_onehot_range(gmpe_handle h, const void* onehot, int elem_bytes, int32_t* action_idx, int64_t world_begin,
                             int64_t world_count, void* cuda_stream) {
  if (!h) return GMPE_E_NULL;
  if (!onehot || !action_idx) return fail(h, GMPE_E_NULL, "gmpe_decode_onehot: NULL buffer");
  if (elem_bytes != 4 && elem_bytes != 8) return fail(h, GMPE_E_INVALID, "gmpe_decode_onehot: elem_bytes must be 4 or 8");
  if (world_begin < 0 || world_count < 1 || world_begin + world_count > h->cfg.num_worlds)
    return fail(h, GMPE_E_INVALID, "world range outside [0, num_worlds)");
  const int64_t rows = world_count * h->cfg.num_agents;
  const int A = h->cfg.total_actions;
  onehot = (const char*)onehot + world_begin * h->cfg.num_agents * A * elem_bytes;
  action_idx += world_begin * h->cfg.num_agents;
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const int threads = 256;
  const int blocks = (int)((rows + threads - 1) / threads);
  if (elem_bytes == 4) decode_onehot_kernel<float><<<blocks, threads, 0, s>>>((const float*)onehot, action_idx, rows, A);
  else decode_onehot_kernel<double><<<blocks, threads, 0, s>>>((const double*)onehot, action_idx, rows, A);
  cudaError_t e = cudaGetLastError();
  if (prev != h->device) cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(h, GMPE_E_CUDA, std::string("decode_onehot launch failed: ") + cudaGetErrorString(e));
  h->launches += 1;
  return GMPE_OK;
}

// Upload blocks of the batch for gmpe_step_host: `k` near-equal blocks of multiples of 64 worlds
// (every slab of a block stays 16-byte aligned).
static std::vector<std::pair<int64_t, int64_t>> host_slices(int64_t B, int k) {
  std::vector<std::pair<int64_t, int64_t>> out;
  const int64_t al = 64;
  if (k < 1 || B < 64 * al) k = 1;
  int64_t step = ((B + k - 1) / k + al - 1) / al * al;
  for (int64_t b0 = 0; b0 < B; b0 += step) out.push_back({b0, std::min(B, b0 + step) - b0});
  return out;
}

int gmpe_step_host(gmpe_handle h, const GmpeState* st, const void* actions_host, int action_kind, void* actions_pinned,
                   void* actions_dev, int32_t* idx_dev, const GmpeStepOut* dev_out, const GmpeStepOut* host_out, int slices,
                   void* cuda_stream) {
  if (!h) return GMPE_E_NULL;
  if (!st || !actions_host || !actions_pinned || !actions_dev || !dev_out || !host_out)
    return fail(h, GMPE_E_NULL, "gmpe_step_host: NULL argument");
  if (action_kind != 0 && action_kind != 4 && action_kind != 8) return fail(h, GMPE_E_INVALID, "gmpe_step_host: action_kind must be 0, 4 or 8");
  if (action_kind != 0 && !idx_dev) return fail(h, GMPE_E_NULL, "gmpe_step_host: idx_dev is NULL");
  const GmpeConfig& c = h->cfg;
  const int64_t B = c.num_worlds, N = c.num_agents, E = N + c.num_landmarks + c.num_obstacles + c.num_walls, cap = E * (E - 1);
  const int64_t R = c.precision == GMPE_F64 ? 8 : 4, F = obs_dim_of(c), ND = node_dim_of(c);
  const int64_t act_bpw = action_kind == 0 ? N * 4 : N * (int64_t)c.total_actions * action_kind;  // bytes per world
  // GmpeStepOut is sixteen pointers in declaration order: bytes per world of each
  constexpr int NOUT = 16;
  const int64_t bpw[NOUT] = {N * F * R, N * E * ND * R, E * E * R, N * R, N, N * GMPE_INFO_K * R, 2 * cap * 4, cap * R, 4,
                             2 * cap * 4, cap * R, 4, N * R, N * 4, N * 4, N};
  static_assert(sizeof(GmpeStepOut) == NOUT * sizeof(void*), "GmpeStepOut layout");
  void* const* dptr = reinterpret_cast<void* const*>(dev_out);
  void* const* hptr = reinterpret_cast<void* const*>(host_out);
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  cudaStream_t main_s = (cudaStream_t)cuda_stream;
  const auto sl = host_slices(B, slices);
  cudaError_t e = cudaSuccess;
  if (!h->s_in) e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess && !h->s_out) e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
  while (e == cudaSuccess && h->ev.size() < 2 * sl.size() + 1) {
    cudaEvent_t v;
    e = cudaEventCreateWithFlags(&v, cudaEventDisableTiming);
    if (e == cudaSuccess) h->ev.push_back(v);
  }
  if (e == cudaSuccess && !h->ev_tail) e = cudaEventCreateWithFlags(&h->ev_tail, cudaEventDisableTiming);
  if (actions_pinned != actions_host && !h->pool) {
    unsigned hw = std::thread::hardware_concurrency();
    cpu_set_t cs;  // the cores this process may run on (a rank pinned to its share of the box), not the box
    if (sched_getaffinity(0, sizeof(cs), &cs) == 0 && CPU_COUNT(&cs) > 0) hw = (unsigned)CPU_COUNT(&cs);
    h->pool = new (std::nothrow) StagePool(hw >= 16 ? 7 : (hw >= 8 ? 3 : (hw >= 4 ? 2 : (hw >= 2 ? 1 : 0))));
  }
  const bool staging = actions_pinned != actions_host;
  if (staging && h->pool) h->pool->start((char*)actions_pinned, (const char*)actions_host, (size_t)(B * act_bpw));
  int rc = GMPE_OK;
#ifdef GMPE_EXPERIMENTS
  double tr_t0 = 0;
  h->tr_on = getenv("GMPE_TRACE") && atoi(getenv("GMPE_TRACE"));
  if (h->tr_on) {
    while (h->tr_ev.size() < 3 * sl.size() + 1) { cudaEvent_t v; cudaEventCreate(&v); h->tr_ev.push_back(v); }
    h->tr_host.assign(2 * sl.size() + 1, 0.0);
    h->tr_sl = sl;
    tr_t0 = now_us();
    cudaEventRecord(h->tr_ev[0], main_s);
  }
#endif
  // a previous host step whose results were never collected: its copies read the device outputs
  if (e == cudaSuccess && h->tail_pending) e = cudaStreamWaitEvent(main_s, h->ev_tail, 0);
  if (e == cudaSuccess) {  // the copy-in stream starts behind whatever the caller has queued on its stream
    e = cudaEventRecord(h->ev[2 * sl.size()], main_s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->s_in, h->ev[2 * sl.size()], 0);
  }
  for (size_t k = 0; k < sl.size() && e == cudaSuccess && rc == GMPE_OK; ++k) {
    const int64_t b0 = sl[k].first, cnt = sl[k].second;
    const int64_t off = b0 * act_bpw, nb = cnt * act_bpw;
    if (staging) {
      if (h->pool) h->pool->wait_until((size_t)(off + nb));
      else memcpy((char*)actions_pinned + off, (const char*)actions_host + off, (size_t)nb);
    }
#ifdef GMPE_EXPERIMENTS
    if (h->tr_on) h->tr_host[2 * k] = now_us() - tr_t0;
#endif
    e = cudaMemcpyAsync((char*)actions_dev + off, (const char*)actions_pinned + off, (size_t)nb, cudaMemcpyHostToDevice, h->s_in);
    if (e == cudaSuccess) e = cudaEventRecord(h->ev[2 * k], h->s_in);
#ifdef GMPE_EXPERIMENTS
    if (h->tr_on) cudaEventRecord(h->tr_ev[1 + 3 * k], h->s_in);
#endif
    if (e == cudaSuccess) e = cudaStreamWaitEvent(main_s, h->ev[2 * k], 0);
    if (e != cudaSuccess) break;
    const int32_t* idx = action_kind == 0 ? (const int32_t*)actions_dev : idx_dev;
    if (action_kind != 0) rc = gmpe_decode_onehot_range(h, actions_dev, action_kind, idx_dev, b0, cnt, cuda_stream);
    if (rc == GMPE_OK) rc = launch(h, MODE_STEP, st, idx, nullptr, nullptr, dev_out, cuda_stream, b0, cnt, true);
    if (rc != GMPE_OK) break;
    cudaSetDevice(h->device);  // (the calls above restore the caller's device)
#ifdef GMPE_EXPERIMENTS
    if (h->tr_on) { cudaEventRecord(h->tr_ev[2 + 3 * k], main_s); h->tr_host[2 * k + 1] = now_us() - tr_t0; }
#endif
  }
  if (staging && h->pool) h->pool->finish();  // also on the error paths: no helper may outlive the caller's buffers
  // Results go back once every block has been uploaded and stepped, one copy per field over the
  // whole batch: on the B200 boxes measured (profiles/r2_e2e_timeline.txt) a device->host copy that
  // overlaps a host->device copy drops from ~50 GB/s to ~19 GB/s while the upload itself crawls at
  // ~8 GB/s, and short per-slice copies reach 33-40 GB/s -- serial directions and long copies win.
  if (e == cudaSuccess && rc == GMPE_OK) {
    e = cudaEventRecord(h->ev[1], main_s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->s_out, h->ev[1], 0);
    for (int f = 0; f < NOUT && e == cudaSuccess; ++f)
      if (dptr[f] && hptr[f])
        e = cudaMemcpyAsync(hptr[f], dptr[f], (size_t)(B * bpw[f]), cudaMemcpyDeviceToHost, h->s_out);
#ifdef GMPE_EXPERIMENTS
    if (h->tr_on) for (size_t k = 0; k < sl.size(); ++k) cudaEventRecord(h->tr_ev[3 + 3 * k], h->s_out);
#endif
  }
#ifdef GMPE_EXPERIMENTS
  if (h->tr_on) h->tr_host[2 * sl.size()] = now_us() - tr_t0;
#endif
  if (e == cudaSuccess && rc == GMPE_OK) {
    e = cudaEventRecord(h->ev_tail, h->s_out);
    h->tail_pending = e == cudaSuccess;
  }
  if (prev != h->device) cudaSetDevice(prev);
  if (rc != GMPE_OK) return rc;
  if (e != cudaSuccess) return fail(h, GMPE_E_CUDA, std::string("gmpe_step_host: ") + cudaGetErrorString(e));
  return GMPE_OK;
}

int64_t gmpe_gnn_flat_scratch_elems(gmpe_handle h) {
  if (!h) return 0;
  const int64_t B = h->cfg.num_worlds;
  return B + (B + kFlatBlock - 1) / kFlatBlock + 1;
}

int gmpe_gnn_flatten(gmpe_handle h, const GmpeStepOut* out, int64_t* flat_edge_index, void* flat_edge_attr, int64_t capacity,
                     int64_t* total, int64_t* scratch, void* cuda_stream) {
  if (!h) return GMPE_E_NULL;
  if (!out || !flat_edge_index || !flat_edge_attr || !total || !scratch) return fail(h, GMPE_E_NULL, "gmpe_gnn_flatten: NULL argument");
  if (!out->gnn_edge_index || !out->gnn_edge_attr || !out->gnn_num_edges)
    return fail(h, GMPE_E_NULL, "gmpe_gnn_flatten: out.gnn_edge_index / gnn_edge_attr / gnn_num_edges missing (GMPE_FLAG_GNN_EDGES)");
  if (capacity < 1) return fail(h, GMPE_E_INVALID, "gmpe_gnn_flatten: capacity must be positive");
  const GmpeConfig& c = h->cfg;
  const int64_t B = c.num_worlds;
  const int N = c.num_agents, E = N + c.num_landmarks + c.num_obstacles + c.num_walls, cap = E * (E - 1);
  const int nblk = (int)((B + kFlatBlock - 1) / kFlatBlock);
  int64_t* local = scratch;
  int64_t* block_pre = scratch + B;
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  cudaStream_t s = (cudaStream_t)cuda_stream;
  flat_scan_blocks<<<nblk, kFlatBlock, 0, s>>>(out->gnn_num_edges, B, local, block_pre);
  flat_scan_tops<<<1, kFlatBlock, 0, s>>>(block_pre, nblk, N, total);
  const int grid = (int)std::min<int64_t>((B + 7) / 8, (int64_t)h->sm_count * 8);
  if (c.precision == GMPE_F64)
    flat_expand<double><<<grid, 256, 0, s>>>(out->gnn_edge_index, (const double*)out->gnn_edge_attr, out->gnn_num_edges, local, block_pre,
                                            B, N, E, cap, flat_edge_index, flat_edge_index + capacity, (double*)flat_edge_attr, capacity);
  else
    flat_expand<float><<<grid, 256, 0, s>>>(out->gnn_edge_index, (const float*)out->gnn_edge_attr, out->gnn_num_edges, local, block_pre,
                                           B, N, E, cap, flat_edge_index, flat_edge_index + capacity, (float*)flat_edge_attr, capacity);
  cudaError_t e = cudaGetLastError();
  if (prev != h->device) cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(h, GMPE_E_CUDA, std::string("gmpe_gnn_flatten: ") + cudaGetErrorString(e));
  h->launches += 3;
  return GMPE_OK;
}

int gmpe_host_wait(gmpe_handle h) {
  if (!h) return GMPE_E_NULL;
  if (!h->s_out) return GMPE_OK;
  cudaError_t e = cudaStreamSynchronize(h->s_out);
  if (e != cudaSuccess) return fail(h, GMPE_E_CUDA, std::string("gmpe_host_wait: ") + cudaGetErrorString(e));
  h->tail_pending = false;
#ifdef GMPE_EXPERIMENTS
  if (h->tr_on && !h->tr_sl.empty()) {
    fprintf(stderr, "[trace] host step, %zu slices; device times in us since the call reached the stream\n", h->tr_sl.size());
    for (size_t k = 0; k < h->tr_sl.size(); ++k) {
      float a = 0, b = 0, c = 0;
      cudaEventElapsedTime(&a, h->tr_ev[0], h->tr_ev[1 + 3 * k]);
      cudaEventElapsedTime(&b, h->tr_ev[0], h->tr_ev[2 + 3 * k]);
      cudaEventElapsedTime(&c, h->tr_ev[0], h->tr_ev[3 + 3 * k]);
      fprintf(stderr, "[trace]  slice %zu worlds %6lld: host staged %7.1f enqueued %7.1f | dev H2D done %7.1f kernels done %7.1f D2H done %7.1f\n",
              k, (long long)h->tr_sl[k].second, h->tr_host[2 * k], h->tr_host[2 * k + 1], a * 1e3, b * 1e3, c * 1e3);
    }
    fprintf(stderr, "[trace]  call returned at %7.1f us\n", h->tr_host[2 * h->tr_sl.size()]);
    h->tr_sl.clear();
  }
#endif
  return GMPE_OK;
}

}  // extern "C"
