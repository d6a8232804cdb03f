"""Oracle: vectorised env semantics of the reference for the simple gym-class envs (numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  SyncVectorEnv._reset / _step (auto-reset)   openrl/envs/vec_env/sync_venv.py:129-247
  Single2MultiAgentWrapper / RemoveTruncated  openrl/envs/wrappers/multiagent_wrapper.py:33-79,
                                              openrl/envs/wrappers/extra_wrappers.py:52-134
  gymnasium TimeLimit (500 steps) + CartPole  oracle/cartpole_ref.py (third-party restated)
  GridWorldEnv                                openrl/envs/gridworld/gridworld_env.py:21-86
Returns the reference's 4-tuple: obs (N,A,d) , rewards (N,A,1) float64, dones (N,A) bool, infos.
"""
import numpy as np

from .cartpole_ref import MAX_EPISODE_STEPS, cartpole_step_f64


def pcg64_np_random(seed):
    """gymnasium.utils.seeding.np_random: PCG64(SeedSequence(seed))."""
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


class CartPoleVec:
    obs_dim = 4
    n_actions = 2
    agent_num = 1

    def __init__(self, env_num):
        self.N = env_num
        self.rng = [pcg64_np_random(None) for _ in range(env_num)]
        self.state = np.zeros((env_num, 4), np.float64)
        self.elapsed = np.zeros(env_num, np.int64)

    def _reset_one(self, i):
        self.state[i] = self.rng[i].uniform(low=-0.05, high=0.05, size=(4,))
        self.elapsed[i] = 0

    def reset(self, seed=None):
        for i in range(self.N):
            if seed is not None:
                self.rng[i] = pcg64_np_random(seed + i * 10086)  # sync_venv.py:137
            self._reset_one(i)
        return self.state.astype(np.float32)[:, None, :].copy()

    def step(self, actions):
        """actions (N,1,1) -> obs (N,1,4) f32, rewards (N,1,1) f64, dones (N,1) bool, final_obs (N,4)."""
        obs = np.zeros((self.N, 1, 4), np.float32)
        rewards = np.ones((self.N, 1, 1), np.float64)
        dones = np.zeros((self.N, 1), bool)
        final_obs = np.full((self.N, 4), np.nan, np.float32)
        for i in range(self.N):
            s, terminated = cartpole_step_f64(tuple(self.state[i]), int(actions[i, 0, 0]))
            self.state[i] = s
            self.elapsed[i] += 1
            truncated = self.elapsed[i] >= MAX_EPISODE_STEPS
            done = terminated or truncated
            dones[i, 0] = done
            if done:  # sync_venv.py:213-218 auto-reset; the terminal obs goes to info["final_observation"]
                final_obs[i] = np.array(s, dtype=np.float32)
                self._reset_one(i)
            obs[i, 0] = self.state[i].astype(np.float32)
        return obs, rewards, dones, final_obs


class GridWorldVec:
    """GridWorldEnv (10x10 via make(): nrow=ncol=10, goal (1,1)); resets consume the GLOBAL
    np.random (MT19937) in env order (gridworld_env.py:76-81) — or a caller-supplied table
    `reset_positions[k]` (k-th reset overall) for device-parity runs."""
    obs_dim = 4
    n_actions = 5
    agent_num = 1

    def __init__(self, env_num, nrow=10, ncol=10, reset_table=None):
        self.N, self.nrow, self.ncol = env_num, nrow, ncol
        self.pos = np.zeros((env_num, 2), np.int64)
        self.steps = np.zeros(env_num, np.int64)
        self.goal = np.array([1, 1])
        self.reset_table = reset_table
        self.reset_count = 0

    def _reset_one(self, i):
        self.steps[i] = 0
        if self.reset_table is not None:
            self.pos[i] = self.reset_table[self.reset_count]
            self.reset_count += 1
            return
        while True:
            p = np.random.randint(low=[0, 0], high=[self.nrow, self.ncol])
            if not (p == self.goal).all():
                self.pos[i] = p
                return

    def _obs(self):
        return np.concatenate([self.pos, np.tile(self.goal, (self.N, 1))], axis=1)[:, None, :]

    def reset(self, seed=None):
        for i in range(self.N):
            self._reset_one(i)
        return self._obs()

    def step(self, actions):
        rewards = np.zeros((self.N, 1, 1), np.float64)
        dones = np.zeros((self.N, 1), bool)
        delta = {0: (0, 0), 1: (-1, 0), 2: (1, 0), 3: (0, -1), 4: (0, 1)}
        final_obs = np.full((self.N, 4), -1, np.int64)
        for i in range(self.N):
            d = delta[int(actions[i, 0, 0])]
            self.pos[i] = np.clip(self.pos[i] + np.array(d), [0, 0], [self.nrow - 1, self.ncol - 1])
            reward, done = 0, False
            if (self.pos[i] == self.goal).all():
                reward += 10
                done = True
            else:
                reward -= 1
            if self.steps[i] == 100:
                done = True
                reward -= 10
            else:
                self.steps[i] += 1
            rewards[i, 0, 0] = reward
            dones[i, 0] = done
            if done:
                final_obs[i] = np.concatenate([self.pos[i], self.goal])
                self._reset_one(i)
        return self._obs(), rewards, dones, final_obs


ENVS = {"CartPole-v1": CartPoleVec, "GridWorldEnv": GridWorldVec}


class SimpleSpreadVec:
    """MPE simple_spread (3 agents, 3 landmarks, world_length 25), float64, as the reference:
      World.step / apply_action_force / apply_environment_force / integrate_state /
      get_entity_collision_force      openrl/envs/mpe/core.py:216-344
      MultiAgentEnv.step/_set_action/reset/construct_obs   openrl/envs/mpe/multiagent_env.py:167-243,274-339
      Scenario.reset_world / reward / observation           openrl/envs/mpe/scenarios/simple_spread.py:46-125
    Vec semantics (auto-reset when all agents are done, seeds seed + i*10086): sync_venv.py:129-247.
    Returns obs dict {"policy": (N,3,18), "critic": (N,3,54)} float64 like the reference."""
    obs_dim = 18
    critic_obs_dim = 54
    n_actions = 5
    agent_num = 3
    DT, DAMPING, CONTACT_FORCE, CONTACT_MARGIN = 0.1, 0.25, 1e2, 1e-3
    AGENT_SIZE, SENSITIVITY, WORLD_LENGTH = 0.15, 5.0, 25

    def __init__(self, env_num):
        self.N = env_num
        self.rng = [pcg64_np_random(None) for _ in range(env_num)]
        self.pos = np.zeros((env_num, 3, 2))
        self.vel = np.zeros((env_num, 3, 2))
        self.lm = np.zeros((env_num, 3, 2))
        self.step_count = np.zeros(env_num, np.int64)

    def _reset_one(self, i):
        for a in range(3):
            self.pos[i, a] = self.rng[i].uniform(-1, +1, 2)
            self.vel[i, a] = 0.0
        for l in range(3):
            self.lm[i, l] = 0.8 * self.rng[i].uniform(-1, +1, 2)
        self.step_count[i] = 0

    def _obs_one(self, i):
        obs = []
        for a in range(3):
            parts = [self.vel[i, a], self.pos[i, a]]
            parts += [self.lm[i, l] - self.pos[i, a] for l in range(3)]
            parts += [self.pos[i, o] - self.pos[i, a] for o in range(3) if o != a]
            parts += [np.zeros(2), np.zeros(2)]
            obs.append(np.concatenate(parts))
        return obs

    def _obs(self):
        pol = np.zeros((self.N, 3, 18))
        cri = np.zeros((self.N, 3, 54))
        for i in range(self.N):
            o = self._obs_one(i)
            pol[i] = np.stack(o)
            cri[i] = np.concatenate(o)[None].repeat(3, axis=0)
        return {"policy": pol, "critic": cri}

    def reset(self, seed=None):
        for i in range(self.N):
            if seed is not None:
                self.rng[i] = pcg64_np_random(seed + i * 10086)
            self._reset_one(i)
        return self._obs()

    def _world_step(self, i, actions):
        p_force = [None, None, None]
        for a in range(3):
            onehot = np.zeros(5)
            onehot[int(actions[a])] = 1
            u = np.zeros(2)
            u[0] += onehot[1] - onehot[2]
            u[1] += onehot[3] - onehot[4]
            u *= self.SENSITIVITY
            p_force[a] = 1.0 * u + 0.0
        for a in range(3):
            for b in range(a + 1, 3):
                delta = self.pos[i, a] - self.pos[i, b]
                dist = np.sqrt(np.sum(np.square(delta)))
                dist_min = self.AGENT_SIZE + self.AGENT_SIZE
                k = self.CONTACT_MARGIN
                penetration = np.logaddexp(0, -(dist - dist_min) / k) * k
                force = self.CONTACT_FORCE * delta / dist * penetration
                force_ratio = 1.0 / 1.0
                p_force[a] = force_ratio * force + p_force[a]
                p_force[b] = -(1 / force_ratio) * force + p_force[b]
        for a in range(3):
            self.vel[i, a] = self.vel[i, a] * (1 - self.DAMPING)
            self.vel[i, a] += (p_force[a] / 1.0) * self.DT
            self.pos[i, a] += self.vel[i, a] * self.DT

    def _reward(self, i, agent):
        rew = 0
        for l in range(3):
            dists = [np.sqrt(np.sum(np.square(self.pos[i, a] - self.lm[i, l]))) for a in range(3)]
            rew -= min(dists)
        for a in range(3):
            d = np.sqrt(np.sum(np.square(self.pos[i, a] - self.pos[i, agent])))
            if d < 2 * self.AGENT_SIZE:
                rew -= 1
        return rew

    def step(self, actions):
        """actions (N,3,1) -> obs dict, rewards (N,3,1) f64, dones (N,3) bool."""
        rewards = np.zeros((self.N, 3, 1))
        dones = np.zeros((self.N, 3), bool)
        for i in range(self.N):
            self.step_count[i] += 1
            self._world_step(i, actions[i, :, 0])
            r = np.sum([[self._reward(i, a)] for a in range(3)])
            rewards[i, :, 0] = r
            done = self.step_count[i] >= self.WORLD_LENGTH
            dones[i] = done
            if done:
                self._reset_one(i)
        return self._obs(), rewards, dones, None


ENVS["simple_spread"] = SimpleSpreadVec


class IdentityContinuousVec:
    """IdentityEnvcontinuous (openrl/envs/toy_envs/identity_env.py:87-152; dim=2, ep_length=4) under
    SyncVectorEnv + Single2MultiAgentWrapper: obs Box(1) = the hidden integer state, action Box(1),
    reward = 1 - |state - clip(action, 0, 1)| computed BEFORE the next state is drawn, per-env PCG64."""
    obs_dim = 1
    act_dim = 1
    agent_num = 1

    def __init__(self, env_num):
        self.N = env_num
        self.rng = [pcg64_np_random(None) for _ in range(env_num)]
        self.state = np.zeros(env_num, np.int64)
        self.cur = np.zeros(env_num, np.int64)

    def _next(self, i):
        self.state[i] = self.rng[i].integers(0, 2)

    def reset(self, seed=None):
        for i in range(self.N):
            if seed is not None:
                self.rng[i] = pcg64_np_random(seed + i * 10086)
            self.cur[i] = 0
            self._next(i)
        return self.state.astype(np.float32)[:, None, None].copy()

    def step(self, actions):
        rewards = np.zeros((self.N, 1, 1), np.float64)
        dones = np.zeros((self.N, 1), bool)
        for i in range(self.N):
            a = np.float32(actions[i, 0, 0])
            rewards[i, 0, 0] = 1 - np.abs(self.state[i] - np.clip(a, 0, 1))
            self._next(i)
            self.cur[i] += 1
            done = self.cur[i] >= 4
            dones[i, 0] = done
            if done:  # auto-reset
                self.cur[i] = 0
                self._next(i)
        return self.state.astype(np.float32)[:, None, None].copy(), rewards, dones, None


ENVS["IdentityEnvcontinuous"] = IdentityContinuousVec
