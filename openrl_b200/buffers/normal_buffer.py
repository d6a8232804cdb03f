"""NormalReplayBuffer(cfg, num_agents, obs_space, act_space, data_client, episode_length)
(reference: openrl/buffers/normal_buffer.py:22) wrapping the device ReplayData as `.data`."""
from .replay_data import ReplayData


class NormalReplayBuffer:
    def __init__(self, cfg, num_agents, obs_space, act_space, data_client=None, episode_length=None, device="cuda:0"):
        self.data = ReplayData(cfg, num_agents, obs_space, act_space, data_client, episode_length, device=device)

    def init_buffer(self, raw_obs, action_masks=None):
        self.data.init_buffer(raw_obs, action_masks)

    def after_update(self):
        self.data.after_update()

    def compute_returns(self, next_value, value_normalizer=None):
        self.data.compute_returns(next_value, value_normalizer)

    def get_buffer_size(self):
        d = self.data
        return d.episode_length * d.n_rollout_threads
