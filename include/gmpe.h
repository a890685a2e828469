/*
 * gmpe.h -- C ABI of the B200-native GraphMPE step path.
 *
 * The reference (Jaroan/Contracts-MARL-AAM-Corridors) has no FFI: its boundary for
 * this path is three duck-typed Python surfaces.  This header is the
 * thin layer *under* the Python mirror of those surfaces; every entry point cites
 * the reference interface whose work it replaces.  Plain pointers and sizes only --
 * no torch types -- so the same library binds from ctypes, cgo, JNI or a C++ host.
 *
 * Conventions
 *   - every buffer is CALLER-OWNED device memory (the Python host passes
 *     torch.Tensor.data_ptr()); the library never allocates outputs (its only own device
 *     memory is a 256/512-byte action table and one counter per handle).  Arrays of Real must be aligned to
 *     16 bytes, out.node_obs and out.info to 32 bytes (256-bit row stores);
 *   - 0 = ok, negative = error (GMPE_E_*), text via gmpe_last_error();
 *   - asynchronous on the caller's stream, no host synchronisation inside
 *     reset/step; one handle per device, not re-entrant; the only internal threads are the
 *     up-to-three helpers gmpe_step_host uses to stage a pageable action block;
 *   - `Real` below means float when GmpeConfig.precision == GMPE_F32 and double
 *     when GMPE_F64 (strict-parity instantiation);
 *   - all arrays are world-major, C-contiguous: [B][N][...] .
 */
#ifndef GMPE_H_
#define GMPE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMPE_ABI_VERSION 13

/* error codes */
#define GMPE_OK 0
#define GMPE_E_INVALID (-1)   /* bad argument / unsupported configuration */
#define GMPE_E_CUDA (-2)      /* CUDA runtime error (launch, device)        */
#define GMPE_E_NULL (-3)      /* required pointer is NULL                   */

/* GmpeConfig.scenario */
#define GMPE_SCEN_JULY_TUBE 0 /* nav_metered_one_goal_graph_rotate_tube_july.py */
#define GMPE_SCEN_NAV_GRAPH 1 /* navigation_graph-class, defined-by-build */
#define GMPE_SCEN_THREE_PHASE 2 /* three_phase_graph.py: rotation-invariant fp32 features, 15-d obs,
                                   7-d nodes, exit gate, random tube length;
                                   served for 2 <= N = L <= 16 (the reference runs it with 3, 5 and 10
                                   agents: README.md:88, eval_scripts/commands_test_run.txt:7,30), O = W = 0 */
#define GMPE_SCEN_TWO_PHASE 3 /* two_phase_graph.py (eval_scripts/commands_test_run.txt:14): the same family;
                                 the goal is the tube exit -- an agent is done when it leaves through the
                                 exit gate; no agent-collision penalty; same layouts and restrictions */
#define GMPE_SCEN_CORRIDOR_ROT_INV 4 /* nav_graph_metered_single_corridor_rot_inv.py (the scenario of
                                 model_weights/tube/rot_inv/airtaxi): 13-d obs, progress reward along the
                                 tube (prev_proj), fixed tube length; same restrictions */
/* GmpeConfig.dynamics -- multiagent/core.py:23-26 EntityDynamicsType */
#define GMPE_DYN_DOUBLE_INTEGRATOR 0
#define GMPE_DYN_UNICYCLE 1
#define GMPE_DYN_AIR_TAXI 2
/* GmpeConfig.physics */
#define GMPE_PHYS_KINEMATIC 0 /* multiagent/core.py:687-756 (live step, closed form of RK45) */
#define GMPE_PHYS_FORCE 1     /* onpolicy/envs/mpe/core.py:183-202 order, multiagent/core.py:766-984 forces */
/* GmpeConfig.precision */
#define GMPE_F32 0
#define GMPE_F64 1

/* GmpeConfig.flags */
#define GMPE_FLAG_USE_DONES (1u << 0)       /* nav: goal reach sets status (done)            */
#define GMPE_FLAG_COLLABORATIVE (1u << 1)   /* environment.py:1056-1061 shared reward = sum   */
#define GMPE_FLAG_AUTO_RESET (1u << 2)      /* env_wrappers.py:861-873 reset when all(done)   */
#define GMPE_FLAG_ENV_EDGES (1u << 3)       /* emit update_graph edges (july:1651-1670)       */
#define GMPE_FLAG_GNN_EDGES (1u << 4)       /* emit process_adj edges (gnn_new.py:328-358)    */
#define GMPE_FLAG_INFO (1u << 5)            /* maintain + emit info_callback stats (july:741) */
#define GMPE_FLAG_WALL_POINT_QUIRK (1u << 6)/* walls also act as point colliders (core.py:580) */
#define GMPE_FLAG_CLASSIC_CONTACT (1u << 7) /* d_min = size_a+size_b (onpolicy/envs/mpe/core.py:282) */
#define GMPE_FLAG_FORCE_TILE (1u << 8)      /* testing / A-B timing: run the generic tile kernel even where a
                                               specialised lane-per-agent instantiation exists (ignored by the
                                               rot-inv scenarios, which only that kernel serves) */

/* GmpeConfig.formation -- args.formation_type (july:192, 492-497; multiagent/custom_scenarios/utils.py) */
#define GMPE_FORMATION_POINT 0  /* set_landmarks_in_point  utils.py:165-193 */
#define GMPE_FORMATION_LINE 1   /* set_landmarks_in_line   utils.py:77-131  */
#define GMPE_FORMATION_CIRCLE 2 /* set_landmarks_in_circle utils.py:231-262 */

#define GMPE_MAX_AGENTS 64
#define GMPE_MAX_ENTITIES 160

/* Layout constants */
#define GMPE_WORLD_F 16  /* Reals per world in GmpeState.world_f  */
#define GMPE_WORLD_I 4   /* int32 per world in GmpeState.world_i  */
#define GMPE_AGENT_CNT 8 /* int32 per agent in GmpeState.agent_cnt */
#define GMPE_INFO_K 16   /* Reals per agent in GmpeStepOut.info    */
#define GMPE_WALL_F 4    /* Reals per wall: orient(0=H,1=V), axis_pos, end0, end1 */

/* world_f slots (july tube frame: july:518-613) */
#define GMPE_WF_ENT_X 0
#define GMPE_WF_ENT_Y 1
#define GMPE_WF_EXIT_X 2
#define GMPE_WF_EXIT_Y 3
#define GMPE_WF_E_X 4
#define GMPE_WF_E_Y 5
#define GMPE_WF_N_X 6 /* fp32-rounded, july:602 */
#define GMPE_WF_N_Y 7
#define GMPE_WF_L 8
#define GMPE_WF_HALF_W 9
#define GMPE_WF_WIDTH 10
#define GMPE_WF_ANGLE 11
#define GMPE_WF_DELTA_SPACING 12 /* running sum of Scenario.delta_spacing (july:1180) */
/* world_i slots */
#define GMPE_WI_STEP 0     /* env.current_step (environment.py:1024)      */
#define GMPE_WI_EPISODE 1  /* number of resets so far (RNG counter)       */
/* agent_flags bytes */
#define GMPE_AF_STATUS 0
#define GMPE_AF_PREV_PHASE 1   /* agent.previous_phase; survives resets */
#define GMPE_AF_PHASE_REACHED 2
#define GMPE_AF_GOAL_REACHED 3 /* Scenario.goal_reached[id] + 1 (0 = -1)      */
/* agent_cnt slots (all integer-valued in the reference; the three *_I slots are
 * int64 numpy arrays there, np.full(N, -1), so float assignments truncate) */
#define GMPE_AC_SPACING_VIOL 0
#define GMPE_AC_STEPS_IN_CORRIDOR 1
#define GMPE_AC_CONFORMANCE 2
#define GMPE_AC_AGENT_COLL 3
#define GMPE_AC_OBST_COLL 4
#define GMPE_AC_TIME_REQ_I 5
#define GMPE_AC_DIST_TRAV_I 6
#define GMPE_AC_DIST_LEFT_I 7
/* info columns (july:806-829, same order as the reference dict after 'individual_reward') */
#define GMPE_INFO_DIST_TO_GOAL 0
#define GMPE_INFO_TIME_REQ 1
#define GMPE_INFO_AGENT_COLL 2
#define GMPE_INFO_OBST_COLL 3
#define GMPE_INFO_DIST_MEAN 4
#define GMPE_INFO_DIST_STD 5
#define GMPE_INFO_MEAN_BY_STD 6
#define GMPE_INFO_DISTS_TRAVELED 7
#define GMPE_INFO_TIME_TAKEN 8
#define GMPE_INFO_TIME_MEAN 9
#define GMPE_INFO_TIME_STD 10
#define GMPE_INFO_TIME_MEAN_BY_STD 11
#define GMPE_INFO_CONFORMANCE 12
#define GMPE_INFO_DELTA_SPACING 13
#define GMPE_INFO_SPACING_VIOL 14
#define GMPE_INFO_MIN_TIME 15

/*
 * Frozen POD configuration.  Replaces the argparse Namespace handed to
 * Scenario.make_world(args) (july:109-335) plus the constants of
 * multiagent/config.py and multiagent/core.py:537-566.  Passed by pointer, copied.
 */
typedef struct GmpeConfig {
  int32_t scenario;
  int32_t dynamics;
  int32_t physics;
  int32_t precision;
  int32_t num_worlds;     /* B on this device                         */
  int32_t num_agents;     /* N                                        */
  int32_t num_landmarks;  /* L (july requires L >= N)                 */
  int32_t num_obstacles;  /* O                                        */
  int32_t num_walls;      /* W                                        */
  int32_t episode_length; /* world.world_length                       */
  int32_t total_actions;  /* A: 25 (unicycle/air_taxi), 5 or 9 (DI)   */
  uint32_t flags;
  int32_t formation;      /* GMPE_FORMATION_* (tube scenarios; navigation_graph places goals at random) */
  int32_t reserved0;      /* must be 0 */
  uint64_t seed;          /* Philox key                               */
  int64_t world_offset;   /* global index of local world 0 (multi-GPU sharding: RNG is invariant to G) */
  double world_size;
  double goal_rew, collision_rew, formation_rew;
  double max_speed;       /* args.max_speed (entity.max_speed, july:291) */
  double dt;              /* config.DT                                */
  double v_min, v_max;    /* unicycle/air_taxi speed clamp            */
  double accel_min, accel_max, ang_rate_max;
  double goal_dist_thresh;    /* DISTANCE_TO_GOAL_THRESHOLD           */
  double separation_dist;     /* COLLISION_DISTANCE                   */
  double coordination_range;  /* env-side max_edge_dist (july:242)    */
  double gnn_max_edge_dist;   /* args.max_edge_dist (config.py:256)   */
  double entity_size;         /* Entity.size (core.py:385)            */
  double sensitivity;         /* environment.py:460                   */
  double damping, contact_force, contact_margin;
  double wall_contact_force, wall_contact_margin, wall_width;
  double min_reward, max_reward;
} GmpeConfig;

/* Element counts of every state / output array for a config, so that hosts in
 * any language can size their allocations.  Counts are in ELEMENTS of the
 * array's own type (Real / int32 / uint8). */
typedef struct GmpeLayout {
  int32_t real_bytes;   /* 4 or 8 */
  int32_t num_entities; /* E = N+L+O+W */
  int32_t obs_dim;      /* F_obs: 19 july, 6 nav, 15 two/three_phase, 13 corridor_rot_inv */
  int32_t node_dim;     /* F_node: 8 (7 for the rot-inv family) */
  int32_t edge_cap;     /* E*(E-1) */
  int32_t info_dim;     /* GMPE_INFO_K */
  int64_t agent_state, agent_aux, agent_flags, agent_cnt;
  int64_t world_f, world_i, landmarks, obstacles, walls, agent_ext;
  int64_t obs, node_obs, adj, reward, done, info;
  int64_t edge_index, edge_attr, num_edges;
} GmpeLayout;

/* Persistent simulator state (device pointers, caller-owned). */
typedef struct GmpeState {
  void* agent_state;    /* Real [B][N][4]: (x,y,theta,v) or DI (x,y,vx,vy)   core.py:80-341 state.values */
  void* agent_aux;      /* Real [B][N][2]: p_dist, goal_min_time              core.py:36, july:941      */
  uint8_t* agent_flags; /* u8   [B][N][4]: GMPE_AF_*                                                   */
  int32_t* agent_cnt;   /* i32  [B][N][GMPE_AGENT_CNT]: GMPE_AC_* (only touched with GMPE_FLAG_INFO,
                           except SPACING_VIOL/STEPS/CONFORMANCE which reward() owns)                  */
  void* world_f;        /* Real [B][GMPE_WORLD_F]                                                      */
  int32_t* world_i;     /* i32  [B][GMPE_WORLD_I]                                                      */
  void* landmarks;      /* Real [B][L][2]                                                              */
  void* obstacles;      /* Real [B][O][2]  (may be NULL when O == 0)                                   */
  void* walls;          /* Real [B][W][GMPE_WALL_F] (may be NULL when W == 0)                          */
  int32_t* agent_ext;   /* i32  [B][N][2]: entry_reward_cooldown (three_phase_graph.py:348,652-654,1037);
                           the bits of the float32 prev_proj (nav_graph_metered_single_corridor_rot_inv.py:
                           374,1268-1276).  Rot-inv family only; may be NULL for the other scenarios           */
} GmpeState;

/* Per-step outputs (device pointers, caller-owned).  Any pointer may be NULL to
 * skip that output, except reward/done for gmpe_step. */
typedef struct GmpeStepOut {
  void* obs;           /* Real [B][N][F_obs]      Scenario.observation (july:1337-1463)                */
  void* node_obs;      /* Real [B][N][E][F_node]  Scenario.graph_observation (july:1584-1649,
                                                  three_phase_graph.py:1280-1445)                      */
  void* adj;           /* Real [B][E][E]          masked cached_dist_mag; the reference returns the SAME
                                                  ndarray for all N agents -> one copy     */
  void* reward;        /* Real [B][N]             Scenario.reward (july:1105-1221)                     */
  uint8_t* done;       /* u8   [B][N]             MultiAgentBaseEnv._get_done (environment.py:264)     */
  void* info;          /* Real [B][N][GMPE_INFO_K] Scenario.info_callback (july:741-829)               */
  int32_t* edge_index; /* i32  [B][2][edge_cap]   update_graph row-major COO (july:1651-1670), -1 pad  */
  void* edge_attr;     /* Real [B][edge_cap]      dists[row,col], 0 pad                                */
  int32_t* num_edges;  /* i32  [B]                                                                     */
  int32_t* gnn_edge_index; /* i32 [B][2][edge_cap] process_adj (gnn_new.py:328-358), per world, strict < */
  void* gnn_edge_attr;     /* Real [B][edge_cap]                                                        */
  int32_t* gnn_num_edges;  /* i32 [B]                                                                   */
  void* indiv_reward;      /* Real [B][N] reward before the collaborative sum: info['individual_reward']
                                          (environment.py:1050)                                          */
  float* mask;             /* f32 [B][N] GMPERunner.insert masks: 0 where done, else 1
                                          (onpolicy/runner/shared/graph_mpe_runner.py:397-402)           */
  float* active_mask;      /* f32 [B][N] ... active_masks: like mask, but 1 for every agent of a world
                                          whose agents are all done (graph_mpe_runner.py:403-405); both
                                          are float32 like the reference buffer, written by gmpe_step only */
  uint8_t* phase_reached;  /* u8 [B][N] rot-inv family only: info['Phase_reached'] = Scenario.phase_reached[agent.id]
                                          at info_callback time (three_phase_graph.py:806, two_phase_graph.py:804,
                                          nav_graph_metered_single_corridor_rot_inv.py:840); like the info rows it keeps
                                          the terminal step's value when the world auto-resets; NULL = not wanted;
                                          written by gmpe_step only (ABI v13) */
} GmpeStepOut;

/* Teacher forcing for bit-parity runs.  NULL pointers = on-device Philox. */
typedef struct GmpeStepInject {
  const void* theta_draw;   /* Real [B][N]: heading used by reset_velocity() when agent i reaches its
                               goal this step (core.py:324-333); NaN entries fall back to Philox       */
  const GmpeState* reset_state; /* state to adopt instead of a Philox reset when a world auto-resets
                               (agent_state, agent_aux, world_f, landmarks, obstacles, walls are read;
                               flags/counters follow reset_world semantics)                            */
  const double* reset_draws;    /* double [B][reset_draws_per_world] (device): uniform [0,1) numbers a
                               world's reset consumes INSTEAD of its Philox stream, in the order
                               reset_world draws from np.random (july:368 wall_length, :530 tube angle,
                               then per spawn attempt jitter x, y (:467) and, when accepted, the heading
                               of reset_velocity (core.py:328)); NaN / past the end = Philox.  Lets a
                               reset be replayed on the reference's own random stream (reset parity);
                               ignored when reset_state is given                                        */
  int32_t reset_draws_per_world;
  int32_t reserved0;
} GmpeStepInject;

typedef struct gmpe_ctx* gmpe_handle;

/* -- lifecycle ------------------------------------------------------------ */
int gmpe_abi_version(void);
/* Validates + copies the config, binds to `device`.  Replaces Scenario.make_world
 * (july:109-335) + MultiAgentGraphEnv.__init__ (environment.py:945-984) for a batch. */
int gmpe_create(const GmpeConfig* cfg, int device, gmpe_handle* out);
int gmpe_destroy(gmpe_handle h);
/* Last error text for `h` (or for the calling thread's last failed create when h == NULL). */
const char* gmpe_last_error(gmpe_handle h);
/* Pure host function: array sizes for a config (no GPU needed). */
int gmpe_layout(const GmpeConfig* cfg, GmpeLayout* out);

/* -- hot path ---------------------------------------------------------------- */
/* Replaces MultiAgentGraphEnv.reset (environment.py:1066-1081) ->
 * Scenario.reset_world/random_scenario/setup_tube_params (july:339-613) for every
 * world whose env_mask byte is non-zero (NULL = all).  With `inject` the initial
 * condition is adopted from inject->reset_state instead of Philox draws.
 * Fills out->obs/node_obs/adj(/edges) for the reset worlds. */
int gmpe_reset(gmpe_handle h, const GmpeState* st, const uint8_t* env_mask,
               const GmpeStepInject* inject, const GmpeStepOut* out, void* cuda_stream);

/* Replaces GraphSubprocVecEnv.step for the whole batch: graphworker 'step'
 * (env_wrappers.py:851-873) -> MultiAgentGraphEnv.step (environment.py:1021-1063) ->
 * _set_action (environment.py:336-475) -> World.step (core.py:687-756 or the force
 * variant) -> observation / reward / graph_observation / _get_done / info_callback
 * (+ auto-reset when GMPE_FLAG_AUTO_RESET).  `action_idx` is int32 [B][N] (the argmax
 * of the reference's one-hot; see gmpe_decode_onehot). */
int gmpe_step(gmpe_handle h, const GmpeState* st, const int32_t* action_idx,
              const GmpeStepInject* inject, const GmpeStepOut* out, void* cuda_stream);

/* gmpe_step for the worlds [world_begin, world_begin + world_count) only.  All pointers are
 * the SAME whole-batch arrays gmpe_step takes; the library re-bases them.  Worlds are
 * independent (graphworker processes in the reference, env_wrappers.py:851-901), so a host can
 * issue slices on different streams and overlap one slice's device->host copies with the
 * next slice's host->device copy and kernel (env_wrappers.py here does). */
int gmpe_step_range(gmpe_handle h, const GmpeState* st, const int32_t* action_idx,
                    const GmpeStepInject* inject, const GmpeStepOut* out, int64_t world_begin,
                    int64_t world_count, void* cuda_stream);

/* Front half of _set_action (environment.py:431,444: np.argmax of the one-hot the
 * runner builds with np.eye(A)[a], graph_mpe_runner.py:375-377).
 * onehot: [B][N][A] of float (elem_bytes 4) or double (8); first maximum wins. */
int gmpe_decode_onehot(gmpe_handle h, const void* onehot, int elem_bytes,
                       int32_t* action_idx, void* cuda_stream);
/* ... for a slice of worlds (whole-batch pointers, as gmpe_step_range). */
int gmpe_decode_onehot_range(gmpe_handle h, const void* onehot, int elem_bytes, int32_t* action_idx,
                             int64_t world_begin, int64_t world_count, void* cuda_stream);

/* GraphSubprocVecEnv.step (onpolicy/envs/env_wrappers.py:959-1004) for a caller whose actions
 * and results live in HOST memory.  The way up is pipelined: the batch is cut into `slices`
 * blocks and for every block the library stages the caller's action block into page-locked
 * memory (helper threads), sends it up and decodes + steps the block (gmpe_decode_onehot_range +
 * gmpe_step_range), so that staging block k+1 overlaps the upload and the kernels of block k.
 * The way down starts when the last block has been stepped: one copy per field that is non-NULL
 * in BOTH dev_out and host_out over the whole batch (measured: copies in opposite directions at
 * the same time, and short per-slice copies, both cost PCIe throughput on the B200 boxes --
 * profiles/r2_e2e_timeline.txt).  Streams are owned by the handle.  The call returns when
 * everything is ENQUEUED; gmpe_host_wait blocks until the results are in host_out.  One host
 * step in flight per handle.
 *   actions_host   [B][N] int32 indices (action_kind 0) or [B][N][A] one-hot float / double
 *                  (action_kind 4 / 8), pageable or page-locked;
 *   actions_pinned page-locked staging block of the same size (may equal actions_host when
 *                  that is page-locked itself: no staging copy then);
 *   actions_dev    device block of the same size; idx_dev int32 [B][N] device (ignored for
 *                  action_kind 0, where actions_dev already holds the indices);
 *   dev_out        device outputs exactly as for gmpe_step; host_out page-locked mirrors. */
int gmpe_step_host(gmpe_handle h, const GmpeState* st, const void* actions_host, int action_kind,
                   void* actions_pinned, void* actions_dev, int32_t* idx_dev, const GmpeStepOut* dev_out,
                   const GmpeStepOut* host_out, int slices, void* cuda_stream);
int gmpe_host_wait(gmpe_handle h);

/* The consumer-side edge list in the layout the consumer takes: what GNNBase.forward
 * (onpolicy/algorithms/utils/gnn_new.py:492-500) gets from process_adj (:328-358) on the
 * (B*N, E, E) stack of adjacency matrices -- one graph per (world, agent), the N graphs of a world
 * sharing one matrix.  From out->gnn_edge_index / gnn_edge_attr / gnn_num_edges of the last step:
 *   flat_edge_index  int64 [2][capacity] (device): row 0 sources, row 1 targets, node id =
 *                    (world * N + agent) * E + entity, edges in (graph, row, col) order;
 *   flat_edge_attr   Real  [capacity]: the distance (edge_dim = 1, environment.py:997);
 *   total            int64 [1] (device): nnz over all B*N graphs.  Entries [total, capacity) are not
 *                    written; when total > capacity nothing may be read (raise capacity: at most
 *                    B * N * E * (E - 1));
 *   scratch          int64 [gmpe_gnn_flat_scratch_elems(h)] (device).
 * Three launches (scan of the per-world counts, gather), asynchronous on the caller's stream. */
int64_t gmpe_gnn_flat_scratch_elems(gmpe_handle h);
int gmpe_gnn_flatten(gmpe_handle h, const GmpeStepOut* out, int64_t* flat_edge_index, void* flat_edge_attr,
                     int64_t capacity, int64_t* total, int64_t* scratch, void* cuda_stream);

/* Re-key the Philox streams: replaces env.seed(seed) (environment.py:219-223, called by
 * train_mpe.py:31 with seed + rank*1000); world_offset as in GmpeConfig. */
int gmpe_set_seed(gmpe_handle h, uint64_t seed, int64_t world_offset);

/* Spawn placement is rejection sampling (july:452-486).  The reference loops until a draw fits;
 * the kernels give up after 1024 rejected draws per entity, keep the last one and count the
 * world here.  Reads (and optionally clears) the handle's device counter: synchronises the
 * device, so not for the hot loop.  gmpe_create already refuses configurations whose agents
 * cannot be packed (N * pi * (separation / 2)^2 > half of the spawn area). */
int gmpe_placement_overflows(gmpe_handle h, int64_t* count, int clear);

/* Number of kernel launches issued through this handle so far (bench gpu_launches). */
int64_t gmpe_launch_count(gmpe_handle h);

/* Name of the step kernel this handle launches (which mapping gmpe_create picked for the
 * configuration: "gmpe_step_coop_kernel" or "gmpe_step_kernel");
 * bench / profile bookkeeping only. */
const char* gmpe_kernel_name(gmpe_handle h);

#ifdef __cplusplus
}
#endif
#endif /* GMPE_H_ */
