#!/usr/bin/env python
"""Generate the committed golden traces from the LIVE Python reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Every trace is produced by stepping the *unmodified* reference
(`multiagent.environment.MultiAgentGraphEnv` + the july scenario, or the classic
`onpolicy.envs.mpe` force step) through the import shims in
`oracle/ref_harness.py`.  A trace holds the initial condition, the action stream,
the random draws the step consumed (teacher-forcing inputs) and every
per-step output of the path.  The reference publishes no golden vectors of its own
(SURVEY.md section 4), so these files ARE the golden set; regenerate only when the
reference changes.
"""
from __future__ import annotations

import ast
import json
import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import (  # noqa: E402
    REFERENCE_ROOT,
    build_july_env,
    install_shims,
    july_args,
)

INFO_KEYS = [
    "Dist_to_goal", "Time_req_to_goal", "Num_agent_collisions", "Num_obst_collisions",
    "Distance_mean", "Distance_variance", "Mean_by_variance", "Dists_traveled",
    "Time_taken", "Time_mean", "Time_stddev", "Time_mean_by_stddev", "Conformance",
    "Delta_spacing", "Spacing_violations", "Min_time_to_goal",
]


def load_process_adj():
    """Execute the reference's own `process_adj` source text (gnn_new.py:328-358)
    without importing its module (torch_geometric is absent here)."""
    import torch
    from torch import Tensor
    from typing import Tuple

    path = os.path.join(REFERENCE_ROOT, "onpolicy/algorithms/utils/gnn_new.py")
    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "process_adj":
            node.decorator_list = []
            code = ast.unparse(node)
            ns = {"torch": torch, "Tensor": Tensor, "Tuple": Tuple}
            exec(textwrap.dedent(code), ns)
            return ns["process_adj"]
    raise RuntimeError("process_adj not found")


def wrap_pi(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


class PursuitPolicy:
    """Goal-seeking action stream so that all three phases, the gate bonus and
    goal-reach events occur inside a 200-step episode (random actions alone never
    leave phase 0).  `eps` mixes in uniform random actions."""

    def __init__(self, env, scenario, rs, eps):
        self.env, self.sc, self.rs, self.eps = env, scenario, rs, eps
        self.stage = [0] * env.n

    def reset(self):
        self.stage = [0] * self.env.n

    def act(self):
        w = self.env.world
        tp = w.tube_params
        e = np.asarray(tp["e"], dtype=np.float64)
        out = []
        for i, ag in enumerate(w.agents):
            p = ag.state.p_pos
            if self.rs.rand() < self.eps:
                out.append(int(self.rs.randint(25)))
                continue
            if self.stage[i] == 0:
                tgt = tp["entrance"] - 0.3 * e
                if np.linalg.norm(tgt - p) < 0.12:
                    self.stage[i] = 1
            if self.stage[i] == 1:
                tgt = tp["exit"] + 0.3 * e
                if np.dot(p - tp["exit"], e) > 0.12:
                    self.stage[i] = 2
            if self.stage[i] == 2:
                tgt = self.sc.landmark_poses[i]
            err = wrap_pi(np.arctan2(tgt[1] - p[1], tgt[0] - p[0]) - ag.state.theta)
            omegas = np.linspace(-0.5, 0.5, 5)
            ar = int(np.argmin(np.abs(omegas - np.clip(err, -0.5, 0.5))))
            ac = 4 if abs(err) < 0.6 else 0
            out.append(ar * 5 + ac)
        return np.array(out, dtype=np.int64)


def tube_vector(w):
    tp = w.tube_params
    return np.array(
        [tp["entrance"][0], tp["entrance"][1], tp["exit"][0], tp["exit"][1],
         tp["e"][0], tp["e"][1], float(tp["n"][0]), float(tp["n"][1]),
         tp["L"], tp["half_width"], tp["width"], tp["angle"]], dtype=np.float64)


def snapshot_init(env, sc):
    w = env.world
    return dict(
        agent_state=np.array([a.state.values.copy() for a in w.agents]),
        prev_phase=np.array([getattr(a, "previous_phase", 0) for a in w.agents], dtype=np.int64),
        tube=tube_vector(w),
        landmarks=np.array([l.state.p_pos.copy() for l in w.landmarks]),
        obstacles=np.array([o.state.p_pos.copy() for o in w.obstacles]).reshape(-1, 2),
        goal_min_time=np.array([a.goal_min_time for a in w.agents]),
    )


def counters(env, sc):
    w = env.world
    n = env.n
    c = np.zeros((n, 8), dtype=np.int64)
    c[:, 0] = sc.spacing_violation
    c[:, 1] = sc.steps_in_corridor
    c[:, 2] = sc.conformance_percent
    c[:, 3] = w.num_agent_collisions
    c[:, 4] = w.num_obstacle_collisions
    c[:, 5] = w.times_required
    c[:, 6] = w.dists_to_goal
    c[:, 7] = w.dist_left_to_goal
    return c


def pad_edges(edge_list, weight, cap):
    ei = -np.ones((2, cap), dtype=np.int64)
    ea = np.zeros(cap)
    k = edge_list.shape[1]
    ei[:, :k] = edge_list
    ea[:k] = weight
    return ei, ea, k


def run_july_trace(path, seed, T, policy, eps=0.15, episodes=1, **over):
    """One env, `episodes` episodes of T steps each (re-reset by the harness with the
    reference's own reset()), outputs concatenated; `ep_start` marks reset rows."""
    import torch

    process_adj = load_process_adj()
    args = july_args(episode_length=T, **over)
    env, sc = build_july_env(args, seed)
    env.seed(seed)
    rs = np.random.RandomState(seed + 12345)
    n = env.n
    E = len(env.world.entities)
    cap = E * (E - 1)
    pol = PursuitPolicy(env, sc, rs, eps)

    rec = {k: [] for k in (
        "actions theta_draw state p_dist status prev_phase phase_reached goal_reached obs node_obs adj "
        "reward done info cnt delta_spacing_sum edge_index edge_attr num_edges "
        "gnn_edge_index gnn_edge_attr gnn_num_edges cooldown prev_proj").split()}
    inits = {k: [] for k in "agent_state prev_phase tube landmarks obstacles goal_min_time reset_obs reset_node_obs reset_adj".split()}

    for ep in range(episodes):
        obs_n, id_n, node_n, adj_n = env.reset(ep)
        pol.reset()
        ini = snapshot_init(env, sc)
        for k, v in ini.items():
            inits[k].append(v)
        inits["reset_obs"].append(np.array(obs_n))
        inits["reset_node_obs"].append(np.array(node_n))
        inits["reset_adj"].append(np.array(adj_n[0]).copy())
        for t in range(T):
            a = pol.act() if policy == "pursuit" else rs.randint(25, size=n)
            status_before = np.array([bool(ag.status) for ag in env.world.agents])
            obs_n, id_n, node_n, adj_n, rew_n, done_n, info_n = env.step([np.eye(25)[i] for i in a])
            w = env.world
            status_after = np.array([bool(ag.status) for ag in w.agents])
            th = np.full(n, np.nan)
            for i in range(n):
                if status_after[i] and not status_before[i]:
                    th[i] = w.agents[i].state.theta  # the reset_velocity() draw (core.py:328)
            assert all(x is adj_n[0] for x in adj_n)
            adj = np.array(adj_n[0]).copy()
            # env-side edges on the CURRENT masked matrix (what the next step's
            # update_graph call would produce, environment.py:1022-1023)
            sc.update_graph(w)
            ei, ea, ne = pad_edges(w.edge_list, w.edge_weight, cap)
            # consumer-side edges: reference process_adj on the fp32 cast the buffer makes
            gi, ga = process_adj(torch.from_numpy(adj.astype(np.float32)), float(args.max_edge_dist))
            gei, gea, gne = pad_edges(gi.numpy(), ga.numpy()[:, 0].astype(np.float64), cap)
            rec["actions"].append(a)
            rec["theta_draw"].append(th)
            rec["state"].append(np.array([ag.state.values.copy() for ag in w.agents]))
            rec["p_dist"].append(np.array([ag.state.p_dist for ag in w.agents]))
            rec["status"].append(status_after)
            rec["prev_phase"].append(np.array([ag.previous_phase for ag in w.agents]))
            rec["phase_reached"].append(sc.phase_reached.copy())
            if "Phase_reached" in info_n[0]:  # rot-inv family (three_phase_graph.py:806): the info entry IS that array's value
                assert [int(d["Phase_reached"]) for d in info_n] == [int(x) for x in sc.phase_reached], "info Phase_reached"
                assert list(info_n[0]).index("Phase_reached") == list(info_n[0]).index("Min_time_to_goal") - 1
            rec["goal_reached"].append(sc.goal_reached.copy())
            rec["obs"].append(np.array(obs_n))
            rec["node_obs"].append(np.array(node_n))
            rec["adj"].append(adj)
            rec["reward"].append(np.array(rew_n, dtype=np.float64).reshape(n))
            rec["done"].append(np.array(done_n, dtype=bool))
            rec["info"].append(np.array([[float(info_n[i][k]) for k in INFO_KEYS] for i in range(n)]))
            rec["cnt"].append(counters(env, sc))
            rec["delta_spacing_sum"].append(float(np.sum(sc.delta_spacing)))
            # three_phase family: entry_reward_cooldown (three_phase_graph.py:348, 652-654, 1037)
            rec["cooldown"].append(np.array(getattr(sc, "entry_reward_cooldown", np.zeros(n)), dtype=np.int64).copy())
            rec["prev_proj"].append(np.array(getattr(sc, "prev_proj", np.zeros(n)), dtype=np.float32).copy())
            rec["edge_index"].append(ei); rec["edge_attr"].append(ea); rec["num_edges"].append(ne)
            rec["gnn_edge_index"].append(gei); rec["gnn_edge_attr"].append(gea); rec["gnn_num_edges"].append(gne)

    out = {k: np.array(v) for k, v in rec.items()}
    out.update({"init_" + k if not k.startswith("reset_") else k: np.array(v) for k, v in inits.items()})
    out["cfg_json"] = np.array(json.dumps({k: v for k, v in vars(args).items()}))
    out["episodes"] = np.array(episodes)
    out["T"] = np.array(T)
    out["individual_reward_check"] = np.array(1)
    np.savez_compressed(path, **out)
    st = out["status"]
    print(f"{os.path.basename(path)}: N={n} E={E} T={T}x{episodes} reached={int(st[-1].sum())} "
          f"phases={sorted(set(out['obs'][..., -1].ravel().tolist()))} "
          f"rew[{out['reward'].min():.2f},{out['reward'].max():.2f}] size={os.path.getsize(path)//1024}KB")


def run_classic_trace(path, seed, T, num_agents=3, num_landmarks=3, pull=0.8):
    """Stage-1 FORCE physics oracle: the classic MPE copy whose World.step runs
    apply_action_force / apply_environment_force / integrate_state live
    (onpolicy/envs/mpe/core.py:183-335) through onpolicy/envs/mpe/environment.py +
    scenarios/simple_spread.py.  The trace holds the agent (x, y, vx, vy) trajectory under a
    seeded action stream in which agents chase their nearest neighbour, so that they actually collide."""
    import argparse
    import contextlib
    import io

    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        from onpolicy.envs.mpe.environment import MultiAgentEnv
        from onpolicy.envs.mpe.scenarios import load
    finally:
        os.chdir(cwd)
    args = argparse.Namespace(num_agents=num_agents, num_landmarks=num_landmarks, episode_length=T, scenario_name="simple_spread")
    np.random.seed(seed)
    scenario = load("simple_spread.py").Scenario()
    world = scenario.make_world(args)
    env = MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation, scenario.info)
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        env.reset()
    ag = world.agents
    init_state = np.array([[*a.state.p_pos, *a.state.p_vel] for a in ag])
    init_lm = np.array([l.state.p_pos.copy() for l in world.landmarks])
    rs = np.random.RandomState(seed + 100)
    acts, states, hits = [], [], 0
    for t in range(T):
        a_idx = []
        for a in ag:
            if rs.rand() < pull:  # chase the nearest other agent; one-hot decode u = [a1 - a2, a3 - a4] (environment.py:238-239)
                others = [o for o in ag if o is not a]
                tgt = min(others, key=lambda o: np.linalg.norm(o.state.p_pos - a.state.p_pos))
                d = tgt.state.p_pos - a.state.p_pos
                k = (1 if d[0] > 0 else 2) if abs(d[0]) > abs(d[1]) else (3 if d[1] > 0 else 4)
            else:
                k = int(rs.randint(5))
            a_idx.append(k)
        onehot = [np.eye(5)[k] for k in a_idx]
        with contextlib.redirect_stdout(io.StringIO()):
            env.step(onehot)
        acts.append(a_idx)
        states.append([[*a.state.p_pos, *a.state.p_vel] for a in ag])
        for i in range(len(ag)):
            for j in range(i + 1, len(ag)):
                hits += int(np.linalg.norm(ag[i].state.p_pos - ag[j].state.p_pos) < ag[i].size + ag[j].size)
    out = dict(init_agent_state=init_state, init_landmarks=init_lm, actions=np.array(acts, dtype=np.int32),
               state=np.array(states), T=T, size=float(ag[0].size), dt=float(world.dt), damping=float(world.damping),
               contact_force=float(world.contact_force), contact_margin=float(world.contact_margin),
               num_agents=num_agents, num_landmarks=num_landmarks)
    np.savez_compressed(path, **out)
    print(f"{os.path.basename(path)}: classic MPE force step N={num_agents} T={T} contact pair-steps={hits} "
          f"size={os.path.getsize(path)//1024}KB")


def run_multiagent_force_trace(path, seed, T, num_agents=4, num_obstacles=3, wall_points=True, done_at=None):
    """Stage-1 FORCE physics with the `multiagent` parameterisation: the reference's OWN
    `World.apply_action_force / apply_environment_force / get_entity_collision_force /
    get_wall_collision_force / integrate_state` (multiagent/core.py:766-964) -- dead code on the
    live kinematic step, but callable -- driven unmodified in the classic step order
    (onpolicy/envs/mpe/core.py:183-202) on a double-integrator world the harness assembles from the
    reference's entity classes: agents, non-colliding landmarks, immovable obstacles, four wall
    segments shorter than the box they frame (agents reach the end caps and walk around them).
    `wall_points`: walls are members of world.entities (core.py:580) and therefore ALSO act as point
    colliders at Wall.state.p_pos; the harness puts that point at the segment midpoint (the build's
    GMPE_FLAG_WALL_POINT_QUIRK convention).  With wall_points=False the walls get collide=False --
    only the segment force remains (the build's default).
    `done_at`: from that step on agent 0 has status=True, zero velocity and the no-op action: the
    reference suppresses the agent-agent force ON it (core.py:899-900) while the others still feel it."""
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        from multiagent.config import DoubleIntegratorConfig
        from multiagent.core import Agent, Entity, EntityDynamicsType, Landmark, Wall, World
    finally:
        os.chdir(cwd)
    rs = np.random.RandomState(seed)
    N, O, h = num_agents, num_obstacles, 1.0
    # P3: World.__init__ reads config_class.COORDINATION_RANGE, which the DI config lacks
    DoubleIntegratorConfig.COORDINATION_RANGE = DoubleIntegratorConfig.COMMUNICATION_RANGE
    world = World(EntityDynamicsType.DoubleIntegratorXY)
    world.agents = [Agent(EntityDynamicsType.DoubleIntegratorXY) for _ in range(N)]
    for i, a in enumerate(world.agents):
        a.id, a.name, a.collide, a.status, a.max_speed = i, f"agent {i}", True, False, 2  # july:283-291 (args.max_speed)
        a.action.u = np.zeros(2)
    world.landmarks = [Landmark() for _ in range(N)]
    for i, l in enumerate(world.landmarks):
        l.id, l.name, l.collide, l.movable = i, f"landmark {i}", False, False  # july:295-300
    world.obstacles = [Entity() for _ in range(O)]
    for i, o in enumerate(world.obstacles):
        o.id, o.name, o.collide, o.movable = i, f"obstacle {i}", True, False   # july:302-309
    segs = [("H", h, (-0.6 * h, 0.6 * h)), ("H", -h, (-0.6 * h, 0.6 * h)), ("V", h, (-0.6 * h, 0.6 * h)), ("V", -h, (-0.6 * h, 0.6 * h))]
    world.walls = [Wall(orient=o, axis_pos=a, endpoints=e) for o, a, e in segs]
    for i, w in enumerate(world.walls):
        w.id, w.name, w.collide, w.movable = i, f"wall {i}", bool(wall_points), False  # july:317-321
        mid = 0.5 * (w.endpoints[0] + w.endpoints[1])
        w.state.p_pos = np.array([mid, w.axis_pos]) if w.orient == "H" else np.array([w.axis_pos, mid])
    for o in world.obstacles:
        o.state.p_pos = rs.uniform(-0.55 * h, 0.55 * h, 2)
    for l in world.landmarks:
        l.state.p_pos = rs.uniform(-0.8 * h, 0.8 * h, 2)
    placed = []
    for a in world.agents:  # spawn clear of the obstacles and of each other (bounded rejection)
        for attempt in range(10000):
            p = rs.uniform(-0.85 * h, 0.85 * h, 2)
            if all(np.linalg.norm(p - q) > 0.52 for q in [o.state.p_pos for o in world.obstacles] + placed):
                break
        else:
            raise RuntimeError("could not place the agents")
        placed.append(p.copy())
        a.state.p_pos = p
        a.state.p_vel = np.zeros(2)
    ag = world.agents
    init_state = np.array([[*a.state.p_pos, *a.state.p_vel] for a in ag])
    u_tab = np.array([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1]], dtype=np.float64) * 5.0  # environment.py:399-409, 460-463
    dmin = DoubleIntegratorConfig.COLLISION_DISTANCE
    acts, states, status = [], [], []
    n_agent = n_done = n_obst = n_wall = n_cap = 0

    def toward(d):
        return (1 if d[0] > 0 else 2) if abs(d[0]) > abs(d[1]) else (3 if d[1] > 0 else 4)

    for t in range(T):
        if done_at is not None and t == done_at:
            ag[0].status = True
            ag[0].state.p_vel = np.zeros(2)  # DoubleIntegratorXYState.reset_velocity
        a_idx = []
        for a in ag:
            if a.status:
                k = 0
            else:
                r = rs.rand()
                if r < 0.35:    # chase the nearest other agent
                    tgt = min((o for o in ag if o is not a), key=lambda o: np.linalg.norm(o.state.p_pos - a.state.p_pos))
                    k = toward(tgt.state.p_pos - a.state.p_pos)
                elif r < 0.55:  # run into the nearest obstacle
                    tgt = min(world.obstacles, key=lambda o: np.linalg.norm(o.state.p_pos - a.state.p_pos))
                    k = toward(tgt.state.p_pos - a.state.p_pos)
                elif r < 0.85:  # head for the end of the nearest wall segment
                    p = a.state.p_pos
                    k = toward(np.array([np.sign(p[0]) * (0.6 * h if abs(p[1]) > abs(p[0]) else h) - p[0],
                                         np.sign(p[1]) * (h if abs(p[1]) > abs(p[0]) else 0.6 * h) - p[1]]))
                else:
                    k = int(rs.randint(5))
            a_idx.append(k)
            a.action.u = u_tab[k].copy()
        # statistics of the configuration the forces are evaluated on
        for i in range(N):
            for j in range(i + 1, N):
                if np.linalg.norm(ag[i].state.p_pos - ag[j].state.p_pos) < dmin:
                    n_agent += 1
                    n_done += int(bool(ag[i].status) or bool(ag[j].status))
            n_obst += sum(np.linalg.norm(ag[i].state.p_pos - o.state.p_pos) < dmin for o in world.obstacles)
            for w in world.walls:
                prll, perp = (0, 1) if w.orient == "H" else (1, 0)
                p = ag[i].state.p_pos
                if w.endpoints[0] - ag[i].size <= p[prll] <= w.endpoints[1] + ag[i].size and abs(p[perp] - w.axis_pos) < ag[i].size + 0.5 * w.width:
                    n_wall += 1
                    n_cap += int(p[prll] < w.endpoints[0] or p[prll] > w.endpoints[1])
        # classic World.step order (onpolicy/envs/mpe/core.py:183-202) on the multiagent force functions
        p_force = [None] * len(world.entities)
        p_force = world.apply_action_force(p_force)
        p_force = world.apply_environment_force(p_force)
        world.integrate_state(p_force)
        acts.append(a_idx)
        states.append([[*a.state.p_pos, *a.state.p_vel] for a in ag])
        status.append([int(bool(a.status)) for a in ag])
    out = dict(init_agent_state=init_state, init_landmarks=np.array([l.state.p_pos for l in world.landmarks]),
               obstacles=np.array([o.state.p_pos for o in world.obstacles]),
               walls=np.array([[0.0 if w.orient == "H" else 1.0, w.axis_pos, w.endpoints[0], w.endpoints[1]] for w in world.walls]),
               actions=np.array(acts, dtype=np.int32), state=np.array(states), status=np.array(status, dtype=np.uint8),
               T=T, size=float(ag[0].size), dt=float(world.dt), damping=float(world.damping),
               contact_force=float(world.contact_force), contact_margin=float(world.contact_margin),
               wall_contact_force=float(world.wall_contact_force), wall_contact_margin=float(world.wall_contact_margin),
               wall_width=float(world.walls[0].width), collision_distance=float(dmin), max_speed=float(ag[0].max_speed),
               wall_points=int(bool(wall_points)), done_at=-1 if done_at is None else int(done_at),
               num_agents=N, num_landmarks=N, num_obstacles=O, num_walls=len(world.walls))
    np.savez_compressed(path, **out)
    print(f"{os.path.basename(path)}: multiagent force step N={N} O={O} W=4 T={T}: pair-steps inside COLLISION_DISTANCE {n_agent} "
          f"(with the done agent {n_done}), obstacle {n_obst}, wall contacts {n_wall} (end caps {n_cap}) "
          f"size={os.path.getsize(path)//1024}KB")


MULTIAGENT_FORCE_TRACES = (
    ("multiagent_force_n4_o3_w4_s51.npz", dict(seed=51, T=200, done_at=100)),
    ("multiagent_force_n4_o3_w4_nopoint_s52.npz", dict(seed=52, T=200, wall_points=False, done_at=120)),
    ("multiagent_force_n6_o4_w4_s53.npz", dict(seed=53, T=150, num_agents=6, num_obstacles=4)),
)

# the rot-inv family beyond the shapes of the reference's README: 11 <= N = L <= 16 (E = 2N <= 32), served by
# the wide instantiations of the lane-per-agent kernel (csrc/gmpe_coop_wide_f32.cu / _f64.cu)
ROT_INV_WIDE_TRACES = (
    ("three_phase_n12_pursuit_s26.npz", dict(seed=26, T=120, policy="pursuit", scenario_name="three_phase_graph",
                                             num_agents=12, num_landmarks=12, world_size=5)),
    ("two_phase_n13_pursuit_s35.npz", dict(seed=35, T=100, policy="pursuit", scenario_name="two_phase_graph",
                                           num_agents=13, num_landmarks=13, world_size=5)),
    ("corridor_rot_inv_n16_pursuit_s44.npz", dict(seed=44, T=80, policy="pursuit", episodes=2,
                                                  scenario_name="nav_graph_metered_single_corridor_rot_inv",
                                                  num_agents=16, num_landmarks=16, world_size=6)),
)


def main():
    install_shims()
    g = lambda name: os.path.join(HERE, name)
    if "--rot-inv-wide" in sys.argv:  # only the wide rot-inv traces (11 <= N <= 16) added late in round 2
        for name, kw in ROT_INV_WIDE_TRACES:
            run_july_trace(g(name), **kw)
        return
    if "--multiagent-force" in sys.argv:  # only the traces added in round 2
        for name, kw in MULTIAGENT_FORCE_TRACES:
            run_multiagent_force_trace(g(name), **kw)
        return
    # C3: july / air_taxi / N=3 -- README.md:84-105 flags
    run_july_trace(g("july_n3_pursuit_s1.npz"), seed=1, T=200, policy="pursuit")
    run_july_trace(g("july_n3_pursuit_s2.npz"), seed=2, T=200, policy="pursuit", eps=0.3)
    run_july_trace(g("july_n3_pursuit_s3_ep3.npz"), seed=3, T=120, policy="pursuit", episodes=3)
    run_july_trace(g("july_n3_random_s4.npz"), seed=4, T=200, policy="random")
    run_july_trace(g("july_n3_collab_s5.npz"), seed=5, T=100, policy="pursuit", collaborative=True)
    # wider worlds (C5 family): N=8 and N=16
    run_july_trace(g("july_n8_pursuit_s6.npz"), seed=6, T=150, policy="pursuit",
                   num_agents=8, num_landmarks=8, world_size=5)
    run_july_trace(g("july_n16_pursuit_s7.npz"), seed=7, T=60, policy="pursuit",
                   num_agents=16, num_landmarks=16, world_size=8)
    # C5: N=64, world_size=12 (node_obs is 512 KB/step in fp64 -> keep T tiny)
    run_july_trace(g("july_n64_pursuit_s8.npz"), seed=8, T=4, policy="pursuit",
                   num_agents=64, num_landmarks=64, world_size=12)
    # unicycle dynamics through the same scenario (P3 patch)
    from multiagent.config import UnicycleVehicleConfig
    UnicycleVehicleConfig.COORDINATION_RANGE = UnicycleVehicleConfig.COMMUNICATION_RANGE
    run_july_trace(g("july_n3_unicycle_s9.npz"), seed=9, T=120, policy="pursuit", dynamics_type="unicycle_vehicle")
    # rotation-invariant family: three_phase_graph.py, air_taxi, N = 3
    run_july_trace(g("three_phase_n3_pursuit_s21.npz"), seed=21, T=150, policy="pursuit", scenario_name="three_phase_graph")
    run_july_trace(g("three_phase_n3_pursuit_s22_ep2.npz"), seed=22, T=110, policy="pursuit", eps=0.3, episodes=2,
                   scenario_name="three_phase_graph")
    run_july_trace(g("three_phase_n3_random_s23.npz"), seed=23, T=60, policy="random", scenario_name="three_phase_graph")
    # ... and two_phase_graph.py (eval_scripts/commands_test_run.txt:14)
    run_july_trace(g("two_phase_n3_pursuit_s31.npz"), seed=31, T=150, policy="pursuit", scenario_name="two_phase_graph")
    run_july_trace(g("two_phase_n3_pursuit_s32_ep2.npz"), seed=32, T=110, policy="pursuit", eps=0.3, episodes=2,
                   scenario_name="two_phase_graph")
    # the shapes of the reference's eval commands (eval_scripts/commands_test_run.txt:7,30)
    run_july_trace(g("two_phase_n5_pursuit_s33.npz"), seed=33, T=100, policy="pursuit", scenario_name="two_phase_graph",
                   num_agents=5, num_landmarks=5, world_size=3)
    run_july_trace(g("three_phase_n10_pursuit_s24.npz"), seed=24, T=150, policy="pursuit", scenario_name="three_phase_graph",
                   num_agents=10, num_landmarks=10, world_size=4)
    # other agent counts of the family (the lane-per-agent kernel is instantiated for N = 2 .. 10)
    run_july_trace(g("two_phase_n2_pursuit_s34.npz"), seed=34, T=60, policy="pursuit", scenario_name="two_phase_graph",
                   num_agents=2, num_landmarks=2, world_size=3)
    run_july_trace(g("three_phase_n4_pursuit_s25.npz"), seed=25, T=80, policy="pursuit", scenario_name="three_phase_graph",
                   num_agents=4, num_landmarks=4, world_size=3)
    run_july_trace(g("corridor_rot_inv_n7_pursuit_s43.npz"), seed=43, T=80, policy="pursuit",
                   scenario_name="nav_graph_metered_single_corridor_rot_inv", num_agents=7, num_landmarks=7, world_size=4)
    # ... and nav_graph_metered_single_corridor_rot_inv.py (model_weights/tube/rot_inv/airtaxi)
    run_july_trace(g("corridor_rot_inv_n3_pursuit_s41.npz"), seed=41, T=150, policy="pursuit",
                   scenario_name="nav_graph_metered_single_corridor_rot_inv")
    run_july_trace(g("corridor_rot_inv_n3_pursuit_s42_ep2.npz"), seed=42, T=110, policy="pursuit", eps=0.3, episodes=2,
                   scenario_name="nav_graph_metered_single_corridor_rot_inv")
    for name, kw in ROT_INV_WIDE_TRACES:
        run_july_trace(g(name), **kw)
    # stage-1 force physics: the live classic-MPE World.step
    run_classic_trace(g("classic_spread_n3_s11.npz"), seed=11, T=120)
    run_classic_trace(g("classic_spread_n6_s12.npz"), seed=12, T=80, num_agents=6, num_landmarks=6)
    # ... and the multiagent/core.py force functions themselves: walls, obstacles, a done agent
    for name, kw in MULTIAGENT_FORCE_TRACES:
        run_multiagent_force_trace(g(name), **kw)


if __name__ == "__main__":
    main()
