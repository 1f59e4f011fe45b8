"""JSRL (Jump-Start RL, arXiv 2204.02372) control logic on top of the B200 IQL engine:
guide/learner arbitration, the horizon curriculum and the online-buffer switch.

Host-side mirror of ``algorithms/finetune/jsrl_utils.py`` (and the JSRL pieces of
``jsrl_w_iql.py:46-60, 182-263``) of the reference: same function names, arguments,
mutable-config protocol and quirks.  It is
pure control flow -- no device work happens here except through ``ImplicitQLearning``
and ``ReplayBuffer``.  Not mirrored: ``get_var_predictor`` / ``VarianceLearner``
(experimental, env-bound, hard-coded paths) and Stable-Baselines3 ``.pth`` guides.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from pathlib import Path, PosixPath
from typing import Optional

import numpy as np
import torch

from .iql import (DeterministicPolicy, GaussianPolicy, ImplicitQLearning, ReplayBuffer, TrainConfig, TwinQ,
                  ValueFunction)

horizon_str = ""  # set by the driver, like the reference's module global (jsrl_w_iql.py sets jsrl.horizon_str)


@dataclass(kw_only=True)
class JsrlTrainConfig(TrainConfig):
    """Fields of the reference's JsrlTrainConfig (jsrl_w_iql.py:46-60).  As in the reference the
    object doubles as a mutable bag of runtime curriculum state."""
    n_curriculum_stages: int = 10
    tolerance: float = 0.05
    rolling_mean_n: int = 5
    horizon_fn: str = "time_step"
    new_online_buffer: bool = True
    online_buffer_size: int = 10000
    max_init_horizon: bool = False
    guide_heuristic_fn: str = None
    no_agent_types: bool = True
    variance_learn_frac: float = 0.9
    env_config: dict = field(default_factory=lambda: {})
    downloaded_dataset: str = None
    pretrained_policy_path: str = None


# ---------------------------------------------------------------------------
# metrics
# ---------------------------------------------------------------------------
def add_jsrl_metrics(eval_log, config):
    eval_log.update({
        "eval/jsrl/curriculum_stage_idx": config.curriculum_stage_idx,
        "eval/jsrl/curriculum_stage": config.curriculum_stage,
        "eval/jsrl/best_eval_score": config.best_eval_score,
        "eval/jsrl/mean_horizon_reached": config.mean_horizon_reached,
        "eval/jsrl/mean_agent_type": config.eval_mean_agent_type,
    })
    return eval_log


# ---------------------------------------------------------------------------
# curriculum (jsrl_utils.py:50-95, 137-174)
# ---------------------------------------------------------------------------
def horizon_update_callback(config, eval_reward):
    """Advance the curriculum when the rolling mean of eval returns, over a FULL window, reaches
    ``best - tolerance*best``.  Reference quirks kept: the bar uses ``best - tol*best`` (which raises the
    bar for negative returns); ``best`` is overwritten by the rolling mean on every advance (it may
    decrease); the window is not cleared on advance; nothing happens once the last stage is reached; a
    ``-inf`` best passes the first full window."""
    config.rolling_mean_rews.append(eval_reward)
    rolling = np.mean(config.rolling_mean_rews)
    if config.curriculum_stage == config.all_curriculum_stages[-1]:
        return config
    bar = config.best_eval_score
    if not np.isinf(bar):
        bar = bar - config.tolerance * bar
    print(f"prev best: {bar}, rolling: {rolling}, best eval: {config.best_eval_score}")
    if len(config.rolling_mean_rews) == config.rolling_mean_n and rolling >= bar:
        config.curriculum_stage_idx += 1
        config.curriculum_stage = config.all_curriculum_stages[config.curriculum_stage_idx]
        config.agent_type_stage = config.all_agent_types[config.curriculum_stage_idx]
        config.best_eval_score = rolling
    return config


def max_to_min_curriculum(init_horizon, n_curriculum_stages):
    return np.linspace(init_horizon, 0, n_curriculum_stages)


def min_to_max_curriculum(init_horizon, n_curriculum_stages):
    return np.linspace(0, init_horizon, n_curriculum_stages)


def prepare_finetuning(init_horizon, config):
    stages = HORIZON_FNS[config.horizon_fn]["generate_curriculum_fn"](init_horizon, config.n_curriculum_stages)
    lo = 1 if config.no_agent_types else 0
    config.all_agent_types = np.linspace(lo, 1, config.n_curriculum_stages)
    config.all_curriculum_stages = stages
    config.curriculum_stage_idx = 0
    config.curriculum_stage = stages[0]
    config.agent_type_stage = config.all_agent_types[0]
    if config.n_curriculum_stages == 1:
        config.agent_type_stage = 1
    config.best_eval_score = -np.inf
    config.rolling_mean_rews = deque(maxlen=config.rolling_mean_n)
    return config


# ---------------------------------------------------------------------------
# horizon functions: (use_learner, horizon) = f(step, state, env, config)
# ---------------------------------------------------------------------------
def _learner_turn(reached: bool, config) -> bool:
    last = config.curriculum_stage_idx == (config.n_curriculum_stages - 1)
    return bool((reached or last) and config.ep_agent_type <= config.agent_type_stage)


def timestep_horizon(step, _s, _e, config):
    if np.isnan(config.curriculum_stage):  # offline-phase / guide-only evaluation: the learner always acts
        return True, step
    return _learner_turn(step >= config.curriculum_stage, config), step


def _antmaze_goal_dist(_, env):
    return np.linalg.norm(np.array(env.target_goal) - np.array(env.get_xy()))


def _lunar_goal_dist(state, _):
    return np.linalg.norm(np.zeros(2) - np.asarray(state)[:2])


GOAL_MAP = {name: _antmaze_goal_dist for name in (
    "antmaze-umaze-v2", "antmaze-umaze-diverse-v2", "antmaze-medium-play-v2", "antmaze-medium-diverse-v2",
    "antmaze-large-play-v2", "antmaze-large-diverse-v2")}
GOAL_MAP["LunarLander-v2"] = _lunar_goal_dist


def goal_dist_calc(state, env):
    return GOAL_MAP[env.spec.id](state, env)


def goal_distance_horizon(_t, s, env, config):
    dist = goal_dist_calc(s, env)
    if np.isnan(config.curriculum_stage):
        return True, dist
    return _learner_turn(dist <= config.curriculum_stage, config), dist


def variance_horizon(_, s, _e, config):
    var = config.vf(torch.Tensor(s))
    if np.isnan(config.curriculum_stage):
        return True, var
    return _learner_turn(var <= config.curriculum_stage, config), var


def agent_type_horizon(_st, _s, _e, config):
    if np.isnan(config.curriculum_stage):
        return True, config.ep_agent_type
    use_learner = False
    last = config.curriculum_stage_idx == (config.n_curriculum_stages - 1)
    if last or config.ep_agent_type <= config.curriculum_stage:
        use_learner = bool(np.random.sample() < config.curriculum_stage)  # consumes the global numpy stream
    return use_learner, config.ep_agent_type


def max_accumulator(v):
    return np.max(v)


def mean_accumulator(v):
    return np.mean(v)


def static_accumulator(v):
    return 1


HORIZON_FNS = {
    "time_step": {"horizon_fn": timestep_horizon, "accumulator_fn": mean_accumulator,
                  "generate_curriculum_fn": max_to_min_curriculum},
    "agent_type": {"horizon_fn": agent_type_horizon, "accumulator_fn": static_accumulator,
                   "generate_curriculum_fn": min_to_max_curriculum},
    "goal_dist": {"horizon_fn": goal_distance_horizon, "accumulator_fn": max_accumulator,
                  "generate_curriculum_fn": min_to_max_curriculum},
    "variance": {"horizon_fn": variance_horizon, "accumulator_fn": mean_accumulator,
                 "generate_curriculum_fn": min_to_max_curriculum},
}


def accumulate(vals):
    return HORIZON_FNS[horizon_str]["accumulator_fn"](vals)


# ---------------------------------------------------------------------------
# agents
# ---------------------------------------------------------------------------
def _is_policy_module(agent) -> bool:
    return isinstance(agent, (GaussianPolicy, DeterministicPolicy))


def _agent_action(agent, env, state, device):
    if not _is_policy_module(agent):
        return agent(env, state)  # heuristic guide: f(env, state)
    try:
        return agent.act(state, device)
    except AttributeError:
        return agent(torch.tensor(state.reshape(1, -1), device=device, dtype=torch.float32))


def learner_or_guide_action(state, step, env, learner, guide, config, device, eval=False, as_numpy=False):
    """Pick the acting agent for this env step (jsrl_utils.py:547-622): with no guide the learner always
    acts (the horizon is still evaluated for logging).  ``as_numpy`` (not in the reference): hand a numpy action back
    as it is instead of wrapping it in a CPU tensor the caller converts back (16 us per env step)."""
    use_learner, horizon = HORIZON_FNS[horizon_str]["horizon_fn"](step, state, env, config)
    if guide is None:
        use_learner = True
    if use_learner:
        action = _agent_action(learner, env, state, device)
    else:
        action = _agent_action(guide, env, state, device)
    if as_numpy and not eval and isinstance(action, np.ndarray):
        return action.reshape(-1), use_learner, horizon
    if not use_learner:
        if not eval and not isinstance(action, torch.Tensor):
            action = torch.tensor(action)
    if eval and isinstance(action, torch.Tensor):
        action = action.cpu().numpy().flatten()
    elif not eval:
        if isinstance(action, np.ndarray):
            action = torch.tensor(action)
        elif len(action.size()) > 1:
            action = action.flatten()
    return action, use_learner, horizon


_actors_built = 0


def make_actor(config, state_dim, action_dim, max_action, device=None, max_steps=None, **engine_kw):
    """Build Q / V / policy networks, their three Adam optimizers and the trainer exactly as the
    reference does (jsrl_utils.py:219-282: class-default 2x256 networks, TwinQ -> V -> policy
    construction order, which fixes the init RNG stream)."""
    if device is None:
        device = config.device
    q_network = TwinQ(state_dim, action_dim).to(device)
    v_network = ValueFunction(state_dim).to(device)
    policy_cls = DeterministicPolicy if config.iql_deterministic else GaussianPolicy
    actor = policy_cls(state_dim, action_dim, max_action, dropout=config.actor_dropout).to(device)
    v_optimizer = torch.optim.Adam(v_network.parameters(), lr=config.vf_lr)
    q_optimizer = torch.optim.Adam(q_network.parameters(), lr=config.qf_lr)
    actor_optimizer = torch.optim.Adam(actor.parameters(), lr=config.actor_lr)
    # the Philox dropout key follows the run's seed (the reference's masks follow torch.manual_seed(config.seed));
    # every trainer built for a run (guide, learner) gets a distinct stream via the construction counter
    global _actors_built
    engine_kw.setdefault("seed", (int(getattr(config, "seed", 0)) << 8) + (_actors_built & 0xFF))
    _actors_built += 1
    return ImplicitQLearning(max_action=max_action, actor=actor, actor_optimizer=actor_optimizer,
                             q_network=q_network, q_optimizer=q_optimizer, v_network=v_network,
                             v_optimizer=v_optimizer, discount=config.discount, tau=config.tau, device=device,
                             beta=config.beta, iql_tau=config.iql_tau, max_steps=max_steps, **engine_kw)


def load_guide(trainer, pretrained):
    """Load a pretrained IQL checkpoint into ``trainer`` and return its actor in eval mode; falls back to a
    CPU ``map_location`` like the reference (jsrl_utils.py:98-134).  SB3 ``.pth`` guides are not supported."""
    if not isinstance(pretrained, PosixPath):
        return pretrained
    if pretrained.suffix == ".pth":
        raise NotImplementedError("Stable-Baselines3 guides (.pth) need stable_baselines3, which is out of scope")
    try:
        trainer.load_state_dict(torch.load(pretrained))
    except RuntimeError:
        trainer.load_state_dict(torch.load(pretrained, map_location=torch.device("cpu")))
    guide = trainer.actor
    guide.eval()
    return guide


def get_guide_agent(config, trainer, state_dim, action_dim, max_action, heuristics=None):
    if config.guide_heuristic_fn is not None:
        if heuristics is None:
            raise ValueError("guide_heuristic_fn needs a module/namespace of heuristic guides (heuristics=...)")
        return getattr(heuristics, config.guide_heuristic_fn), None
    if trainer is None:
        guide_trainer = make_actor(config, state_dim, action_dim, max_action)
        guide = load_guide(guide_trainer, Path(config.pretrained_policy_path))
        guide.eval()
        return guide, guide_trainer
    guide = trainer.actor
    guide.eval()
    return guide, trainer


def get_learning_agent(config, guide_trainer, init_horizon, state_dim, action_dim, max_action):
    """A FRESH random-init learner (no LR schedule: max_steps=None) unless there is a single curriculum
    stage, in which case the guide's networks are copied (optimizers stay fresh); ``total_it`` continues
    from ``offline_iterations`` (jsrl_utils.py:326-357)."""
    trainer = make_actor(config, state_dim, action_dim, max_action)
    if config.n_curriculum_stages == 1 and config.guide_heuristic_fn is None:
        trainer.partial_load_state_dict(guide_trainer.state_dict())
    trainer.total_it = config.offline_iterations
    config = prepare_finetuning(init_horizon, config)
    return trainer, config


def get_online_buffer(config, replay_buffer, state_dim, action_dim):
    """A new ``online_buffer_size`` ring (default 10,000 rows) replaces the offline buffer unless
    ``new_online_buffer`` is False (jsrl_w_iql.py:232-263)."""
    if config.new_online_buffer:
        del replay_buffer
        return ReplayBuffer(state_dim, action_dim, config.online_buffer_size, config.device)
    return replay_buffer
