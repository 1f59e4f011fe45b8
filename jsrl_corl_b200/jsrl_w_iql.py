"""Offline -> online JSRL driver loop on the B200 IQL engine: the host-side mirror of
``algorithms/finetune/jsrl_w_iql.py:62-263, 432-606`` of the reference for continuous-action envs.

Kept from the reference: guide-only evaluation to find the initial horizon, the fresh online learner
and the 10k-row online ring, ``ep_agent_type`` bookkeeping, exploration noise / sampling of learner
actions, ``real_done`` excluding time-outs, the update gate on the GLOBAL iteration counter
(``t >= batch_size`` with a new online buffer), evaluation every ``eval_freq`` iterations, the
curriculum callback and checkpoints named ``checkpoint_{t}.pt``.  Replaced: wandb / ray.tune by a
``log`` callable; env construction, D4RL download and dataset normalisation stay with the caller.
Discrete-action heuristics guides (LunarLander / CartPole) are out of scope.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch

from . import jsrl_utils as jsrl
from .iql import (DeterministicPolicy, GaussianPolicy, ReplayBuffer, is_goal_reached, modify_reward_online)


def _is_gymnasium(env) -> bool:
    return "gymnasium" in str(type(env))


def _reset(env, seed=None):
    if _is_gymnasium(env):
        state, _ = env.reset(seed=seed) if seed is not None else env.reset()
        return state
    if seed is not None and hasattr(env, "seed"):
        env.seed(seed)
    state = env.reset()
    return state[0] if isinstance(state, tuple) else state


def _step(env, action):
    out = env.step(action)
    if len(out) == 5:
        state, reward, term, trunc, info = out
        return state, reward, bool(term or trunc), info
    return out


@torch.no_grad()
def eval_actor(env, learner, guide, config) -> Tuple[np.ndarray, float, float, float]:
    """n_episodes mixed guide/learner rollouts; returns (returns, success rate, horizon, mean agent type).
    With ``guide is None`` the "learner" is the guide being probed for the initial horizon."""
    is_module = isinstance(learner, (GaussianPolicy, DeterministicPolicy))
    if is_module:
        learner.eval()
    returns, successes, horizons, agent_types = [], [], [], []
    for ep in range(config.n_episodes):
        state = _reset(env, config.seed if ep == 0 else None)
        done, ts, ep_ret, goal = False, 0, 0.0, False
        ep_horizons, ep_types = [], []
        while not done:
            config.ep_agent_type = 0 if ts == 0 else np.mean(ep_types)
            action, use_learner, horizon = jsrl.learner_or_guide_action(state, ts, env, learner, guide, config,
                                                                        config.device, eval=True)
            ep_horizons.append(horizon)
            ep_types.append(1 if use_learner else 0)
            state, reward, done, info = _step(env, action)
            ep_ret += reward
            ts += 1
            goal = goal or is_goal_reached(reward, info)
        successes.append(float(goal))
        returns.append(ep_ret)
        probing = guide is None and config.max_init_horizon
        horizons.append(np.max(ep_horizons) if probing else jsrl.accumulate(ep_horizons))
        agent_types.append(np.mean(ep_types))
    horizon = np.max(horizons) if (guide is None and config.max_init_horizon) else np.mean(horizons)
    if is_module:
        learner.train()
    return np.asarray(returns), float(np.mean(successes)), horizon, float(np.mean(agent_types))


def jsrl_online_actor(config, env, trainer, state_dim, action_dim, max_action, heuristics=None):
    """Switch to the online phase: evaluate the guide alone for the initial horizon, then build the learner."""
    config.curriculum_stage = np.nan
    guide, guide_trainer = jsrl.get_guide_agent(config, trainer, state_dim, action_dim, max_action, heuristics)
    _, _, init_horizon, _ = eval_actor(env, guide, None, config)
    trainer, config = jsrl.get_learning_agent(config, guide_trainer, init_horizon, state_dim, action_dim, max_action)
    return trainer, guide, config


def make_offline_buffer(config, dataset: Dict[str, np.ndarray], state_dim: int, action_dim: int, max_episode_steps: int = 1000):
    """Dataset -> device replay buffer the way the reference driver does it (jsrl_w_iql.py:344-368), with the
    arithmetic on the GPU: ``return_reward_range`` (host scan, as in the reference) gives the locomotion reward range,
    ``ReplayBuffer.ingest_d4rl_dataset`` computes the state statistics, normalises, rescales rewards and packs.
    Returns (replay_buffer, state_mean, state_std, reward_mod_dict); wrap the envs with the statistics as before."""
    from .iql import return_reward_range

    reward_mod, shift = {}, 0.0
    if config.normalize_reward:
        if any(s in config.env for s in ("halfcheetah", "hopper", "walker2d")):
            lo, hi = return_reward_range(dataset, max_episode_steps)
            reward_mod = {"max_ret": hi, "min_ret": lo, "max_episode_steps": max_episode_steps}
        elif "antmaze" in config.env:
            shift = 1.0
    rb = ReplayBuffer(state_dim, action_dim, config.buffer_size, config.device)
    mean, std = rb.ingest_d4rl_dataset(dataset, normalize=bool(config.normalize), eps=1e-3, reward_mod=reward_mod or None,
                                       reward_shift=shift)
    return rb, mean, std, reward_mod


def train_loop(config, env, eval_env, replay_buffer: Optional[ReplayBuffer], state_dim: int, action_dim: int,
               max_action: float, max_steps: int, log: Optional[Callable[[Dict, int], None]] = None,
               reward_mod_dict: Optional[Dict] = None, heuristics=None, is_env_with_goal: bool = False):
    """Run ``offline_iterations`` updates on ``replay_buffer`` then ``online_iterations`` env steps with one
    update each.  Returns (trainer, config, history of eval logs)."""
    log = log or (lambda d, step: None)
    config.discrete = False
    jsrl.horizon_str = config.horizon_fn
    if config.pretrained_policy_path is not None:
        config.offline_iterations = 0  # reference quirk: a pretrained guide skips the offline phase
    trainer = actor = guide = None
    if config.offline_iterations > 0 or config.pretrained_policy_path is None:
        trainer = jsrl.make_actor(config, state_dim, action_dim, max_action, max_steps=config.offline_iterations)
        actor = trainer.actor
    if config.checkpoints_path is not None:
        os.makedirs(config.checkpoints_path, exist_ok=True)
    state = _reset(env, config.seed)
    ep_ret, ep_step, goal = 0.0, 0, False
    eval_successes, train_successes, history = [], [], []
    online_buffer, ep_types = None, []
    for t in range(int(config.offline_iterations) + int(config.online_iterations)):
        if t == config.offline_iterations:
            trainer, guide, config = jsrl_online_actor(config, env, trainer, state_dim, action_dim, max_action, heuristics)
            actor = trainer.actor
            online_buffer = jsrl.get_online_buffer(config, replay_buffer, state_dim, action_dim)
            state = _reset(env)
        online_log = {}
        if t >= config.offline_iterations:
            if ep_step == 0:
                ep_types = []
                config.ep_agent_type = 0
            else:
                config.ep_agent_type = np.mean(ep_types)
            action, use_learner, _ = jsrl.learner_or_guide_action(state, ep_step, env, actor, guide, config, config.device,
                                                                  as_numpy=True)
            ep_types.append(1 if use_learner else 0)
            if isinstance(action, np.ndarray):
                # engine-backed policies hand back numpy: the exploration noise still comes from torch's CPU generator
                # (what `torch.randn_like` of the reference's CPU action tensor draws from), everything else stays numpy
                if use_learner and config.iql_deterministic:
                    noise = (torch.randn(action.shape[0]) * config.expl_noise).clamp(-config.noise_clip, config.noise_clip)
                    action = action + noise.numpy()
                action = np.clip(max_action * action, -max_action, max_action).astype(np.float32).flatten()
            else:
                if use_learner and config.iql_deterministic:
                    noise = (torch.randn_like(action) * config.expl_noise).clamp(-config.noise_clip, config.noise_clip)
                    action = action + noise
                action = torch.clamp(max_action * action, -max_action, max_action).cpu().numpy().flatten()
            next_state, reward, done, info = _step(env, action)
            ep_step += 1
            goal = goal or is_goal_reached(reward, info)
            ep_ret += reward
            real_done = bool(done and ep_step < max_steps)  # time-outs are not terminals
            if config.normalize_reward and reward_mod_dict is not None:
                reward = modify_reward_online(reward, config.env, **reward_mod_dict)
            online_buffer.add_transition(state, action, reward, next_state, real_done)
            state = next_state
            if done:
                state = _reset(env)
                if is_env_with_goal:
                    train_successes.append(goal)
                    online_log["train/regret"] = float(np.mean(1 - np.array(train_successes)))
                    online_log["train/is_success"] = float(goal)
                online_log.update({"train/episode_return": ep_ret, "train/mean_ep_agent_type": float(np.mean(ep_types)),
                                   "train/episode_length": ep_step})
                ep_ret, ep_step, goal = 0.0, 0, False
        if t >= config.batch_size or not config.new_online_buffer:
            buf = online_buffer if t >= config.offline_iterations else replay_buffer
            log_dict = trainer.train(buf.sample(config.batch_size))
            if t < config.offline_iterations:
                log_dict["offline_iter"] = t
            else:
                log_dict["online_iter"] = t - config.offline_iterations
            log_dict.update(online_log)
            log(log_dict, trainer.total_it)
            if (t + 1) % config.eval_freq == 0:
                if t < config.offline_iterations:
                    guide = None
                if guide is None:
                    config.curriculum_stage = np.nan
                scores, success, config.mean_horizon_reached, config.eval_mean_agent_type = eval_actor(eval_env, actor, guide, config)
                score = float(scores.mean())
                # reference jsrl_w_iql.py:573-592: with normalize_reward the curriculum callback and the log see the
                # D4RL-normalised score (an affine rescale changes when `best - tol * best` is reached)
                norm_fn = getattr(eval_env, "get_normalized_score", None) if config.normalize_reward else None
                if norm_fn is not None:
                    score = float(norm_fn(score))  # the callback sees this value; the log 100x it (reference :587)
                eval_log = {}
                if is_env_with_goal:
                    eval_successes.append(success)
                    eval_log["eval/regret"] = float(np.mean(1 - np.array(eval_successes)))
                    eval_log["eval/success_rate"] = success
                if t >= config.offline_iterations:
                    config = jsrl.horizon_update_callback(config, score)
                    eval_log = jsrl.add_jsrl_metrics(eval_log, config)
                if norm_fn is not None:
                    eval_log["eval/d4rl_normalized_score"] = score * 100.0
                else:
                    eval_log["eval/score"] = score
                if config.checkpoints_path is not None:
                    torch.save(trainer.state_dict(), os.path.join(config.checkpoints_path, f"checkpoint_{t}.pt"))
                log(eval_log, trainer.total_it)
                history.append(dict(eval_log, t=t))
    return trainer, config, history
