"""B200-native PPO-update hot path with the API of ProximalPolicyOptimization.jl.

Host-side mirror of the reference's Julia interface (`!` becomes a trailing underscore) over
the C ABI of libppo_b200.so.  Julia itself is not available in the build image, so this Python
layer is the executable stand-in for `julia/PPOB200.jl`; both bind exactly the same symbols.
"""
from ._lib import PPOError, LIB_PATH, load as load_library
from .context import Context, default_context
from .rollout_buffer import (DeviceRollouts, DeviceDataset, StateData, batch_state, update_, length,
                             compute_state_value_, permute_, shuffle_, construct_dataset, get_sample, get_batch,
                             pad_vertex_scores, pad_action_mask, prepare_state_data_for_batching_, pack_action_mask)
from .collect_rollouts import (collect_step_data_, collect_episode_data_, collect_rollouts_, compute_returns)
from .policy import Policy, Adam, Optimiser, action_probabilities, batch_action_probabilities, \
    number_of_actions_per_state, batch_sample_actions
from .train import (simplified_ppo_clip, get_linear_action_index, batch_advantage, step_batch_, step_epoch_,
                    ppo_train_, ppo_iterate_, get_optimizer_learning_rate, ppo_loss_with_entropy_from_logits,
                    format_epoch_line)
from .evaluate import single_trajectory_return, average_returns
from .rollouts_to_disk import DiskRollouts, write_returns_to_disk, collect_rollouts_disk_
from .dataset import DiskDataset, load_sample, load_batch, construct_disk_dataset
from . import distributed, bson_io

GEMM_FP32_SIMT, GEMM_TF32X3_TC, GEMM_BF16_TC, GEMM_F16X3_TC = 0, 1, 2, 3
GEMM_AUTO = -1
