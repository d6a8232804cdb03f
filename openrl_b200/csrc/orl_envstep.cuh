// Device-side env.step wrappers (auto-reset + episode statistics) and the categorical sampler shared by the
// feed-forward rollout (orl_rollout.cu) and the recurrent rollout (orl_rnn.cu).
#pragma once
#include "orl_envs.cuh"
#include "orl_mlp.cuh"

namespace orl {

struct EnvPtrs {
    double* f64; uint64_t* u64; int32_t* i32; const int32_t* table; int table_len; uint64_t seed;
    float* ep_return; int32_t* ep_length; double* episode_stats;
    int env_offset = 0;   // first GLOBAL env index of this shard: Philox-keyed env randomness is drawn per global env
};

// One env.step of a single-agent, 4-wide-observation env (CartPole-v1 / GridWorldEnv) with the
// reference's auto-reset (sync_venv.py:213-218): on done the returned obs is the reset obs and the
// terminal obs goes to `fin` (info["final_observation"]).
__device__ __forceinline__ void env_step_single(const EnvPtrs& E, int kind, int e, int N, int act, float (&ob)[4],
                                                float& reward, bool& done, float (&fin)[4]) {
    if (kind == ORL_ENV_CARTPOLE) {
        double s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = E.f64[(size_t)k * N + e];
        int elapsed = E.i32[e];
        const bool terminated = cartpole_dynamics(s, act);
        elapsed += 1;
        done = terminated || (elapsed >= 500);
        reward = 1.0f;
        float ret = E.ep_return[e] + 1.0f;
        int len = E.ep_length[e] + 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) fin[k] = (float)s[k];
        if (done) {
            Pcg64 g = pcg_load(E.u64, e, N);
            cartpole_reset(s, g);
            pcg_store(E.u64, e, N, g);
            elapsed = 0;
            atomicAdd(E.episode_stats + 0, (double)ret);
            atomicAdd(E.episode_stats + 1, (double)len);
            atomicAdd(E.episode_stats + 2, 1.0);
            ret = 0.f; len = 0;
        }
        E.ep_return[e] = ret; E.ep_length[e] = len;
        E.i32[e] = elapsed;
#pragma unroll
        for (int k = 0; k < 4; ++k) { E.f64[(size_t)k * N + e] = s[k]; ob[k] = (float)s[k]; }
    } else {  // ORL_ENV_GRIDWORLD
        int x = E.i32[0 * N + e], y = E.i32[1 * N + e], steps = E.i32[2 * N + e];
        int nreset = E.i32[3 * N + e];
        const int nrow = 10, ncol = 10;
        if (act == 1) x -= 1; else if (act == 2) x += 1; else if (act == 3) y -= 1; else if (act == 4) y += 1;
        x = min(max(x, 0), nrow - 1); y = min(max(y, 0), ncol - 1);
        done = false;
        if (x == 1 && y == 1) { reward = 10.f; done = true; } else reward = -1.f;
        if (steps == 100) { done = true; reward -= 10.f; } else steps += 1;  // gridworld_env.py:68-72
        float ret = E.ep_return[e] + reward;
        int len = E.ep_length[e] + 1;
        fin[0] = (float)x; fin[1] = (float)y; fin[2] = 1.f; fin[3] = 1.f;
        if (done) {
            gridworld_reset(x, y, e, nreset, E.seed, E.table, E.table_len, nrow, ncol, e + E.env_offset);
            nreset += 1; steps = 0;
            atomicAdd(E.episode_stats + 0, (double)ret);
            atomicAdd(E.episode_stats + 1, (double)len);
            atomicAdd(E.episode_stats + 2, 1.0);
            ret = 0.f; len = 0;
        }
        E.ep_return[e] = ret; E.ep_length[e] = len;
        E.i32[0 * N + e] = x; E.i32[1 * N + e] = y; E.i32[2 * N + e] = steps; E.i32[3 * N + e] = nreset;
        ob[0] = (float)x; ob[1] = (float)y; ob[2] = 1.f; ob[3] = 1.f;
    }
}

// One simple_spread env.step (3 agents) with auto-reset at world_length = 25.
__device__ __forceinline__ void env_step_mpe(const EnvPtrs& E, int e, int N, const int (&acts)[3], float (&ob)[3][18],
                                             float& reward, bool& done) {
    MpeState s;
    mpe_load(E.f64, e, N, s);
    int step = E.i32[e] + 1;
    mpe_world_step(s, acts);
    const double r = mpe_shared_reward(s);
    reward = (float)r;
    done = step >= 25;
    float ret = E.ep_return[e] + reward;
    int len = E.ep_length[e] + 1;
    if (done) {
        Pcg64 g = pcg_load(E.u64, e, N);
        mpe_reset(s, g);
        pcg_store(E.u64, e, N, g);
        step = 0;
        atomicAdd(E.episode_stats + 0, (double)ret);
        atomicAdd(E.episode_stats + 1, (double)len);
        atomicAdd(E.episode_stats + 2, 1.0);
        ret = 0.f; len = 0;
    }
    E.ep_return[e] = ret; E.ep_length[e] = len;
    E.i32[e] = step;
    mpe_store(E.f64, e, N, s);
#pragma unroll
    for (int ag = 0; ag < 3; ++ag) mpe_obs(s, ag, ob[ag]);
}

__device__ __forceinline__ int sample_categorical(const float (&p)[MAX_OUT], int n, const float (&q)[MAX_OUT]) {
    // torch.multinomial(probs, 1) == argmax(probs / q), first index wins ties
    int best = 0;
    float bv = p[0] / q[0];
#pragma unroll
    for (int j = 1; j < MAX_OUT; ++j)
        if (j < n) { const float v = p[j] / q[j]; if (v > bv) { bv = v; best = j; } }
    return best;
}

}  // namespace orl
