"""jsrl_corl_b200 -- B200-native IQL(+JSRL) update engine behind the Python API
of LaurenYTaylor/jsrl-CORL (see DESIGN.md / INTEGRATION.md)."""
from .iql import (  # noqa: F401
    ENVS_WITH_GOAL, EXP_ADV_MAX, LOG_STD_MAX, LOG_STD_MIN, MLP, DeterministicPolicy, GaussianPolicy,
    ImplicitQLearning, OfflineTrainConfig, ReplayBuffer, Squeeze, TrainConfig, TwinQ, ValueFunction,
    asymmetric_l2_loss, compute_mean_std, modify_reward, modify_reward_online, normalize_states, set_seed,
    soft_update,
)
from .engine import EnsembleEngine, query_layout  # noqa: F401
from .ensemble import IQLEnsemble, reference_init, shard_members  # noqa: F401
