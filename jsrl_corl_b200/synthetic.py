"""Synthetic D4RL-shaped datasets for benchmarks:
obs/next_obs ~ N(0,1), actions ~ U(-1,1), rewards ~ N(0,1) (antmaze: in {-1,0}),
terminals ~ Bernoulli(1e-3), all from numpy RandomState(seed)."""
import numpy as np


def synthetic_dataset(n: int, state_dim: int, action_dim: int, seed: int = 0, antmaze_rewards: bool = False):
    rng = np.random.RandomState(seed)
    obs = rng.standard_normal((n, state_dim)).astype(np.float32)
    nobs = rng.standard_normal((n, state_dim)).astype(np.float32)
    act = rng.uniform(-1.0, 1.0, (n, action_dim)).astype(np.float32)
    if antmaze_rewards:
        rew = -(rng.uniform(size=n) < 0.98).astype(np.float32)
    else:
        rew = rng.standard_normal(n).astype(np.float32)
    term = rng.uniform(size=n) < 1e-3
    return {"observations": obs, "actions": act, "rewards": rew, "next_observations": nobs, "terminals": term}
