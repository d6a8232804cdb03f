// Fused rollout kernel: policy MLP forward + categorical sampling + device env.step + in-place
// buffer insert, for steps [t_begin, t_end) of one rollout, plus env reset and the batched critic
// forward.  See include/openrl_b200.h for the reference functions this replaces.
//
// Mapping: a CTA of 128 threads owns ROWS = 32 consecutive rows (env, agent) for the whole step
// range — envs never interact, so there is no grid-wide dependency and ONE launch covers all T
// steps; the policy weights (folded, 37 KB) stay resident in shared memory.  Per step the trunk is
// two register-tiled tile GEMMs (32x64xd, 32x64x64; see orl_mlp.cuh), then 4 lanes per row compute
// the logits, one lane samples and steps the env, writing slot t / t+1 of the buffers directly.
// The step chain is latency-bound (T sequential steps); the grid is N*A/32 CTAs.
#include "orl_envstep.cuh"

namespace {
using namespace orl;

constexpr int R_NT = 128;  // threads per CTA; rows per CTA R_M is a template parameter (8 / 16 / 32)


__global__ void env_step_mpe_kernel(int N, EnvPtrs E, const float* __restrict__ actions, float* __restrict__ obs_out,
                                    float* __restrict__ critic_obs_out, float* __restrict__ rewards_out,
                                    float* __restrict__ dones_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const int acts[3] = {(int)actions[e * 3], (int)actions[e * 3 + 1], (int)actions[e * 3 + 2]};
    float ob[3][18], reward; bool done;
    env_step_mpe(E, e, N, acts, ob, reward, done);
    for (int ag = 0; ag < 3; ++ag) {
        for (int k = 0; k < 18; ++k) {
            obs_out[((size_t)e * 3 + ag) * 18 + k] = ob[ag][k];
            if (critic_obs_out) for (int dst = 0; dst < 3; ++dst) critic_obs_out[((size_t)e * 3 + dst) * 54 + ag * 18 + k] = ob[ag][k];
        }
        rewards_out[e * 3 + ag] = reward;
        dones_out[e * 3 + ag] = done ? 1.f : 0.f;
    }
}

__global__ void env_step_kernel(int kind, int N, EnvPtrs E, const float* __restrict__ actions, float* __restrict__ obs_out,
                                float* __restrict__ rewards_out, float* __restrict__ dones_out, float* __restrict__ final_obs_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    float ob[4], fin[4], reward; bool done;
    env_step_single(E, kind, e, N, (int)actions[e], ob, reward, done, fin);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        obs_out[(size_t)e * 4 + k] = ob[k];
        if (final_obs_out) final_obs_out[(size_t)e * 4 + k] = fin[k];
    }
    rewards_out[e] = reward;
    dones_out[e] = done ? 1.f : 0.f;
}


template <int R_M, int ENV>
__global__ void __launch_bounds__(R_NT) rollout_kernel(const OrlRolloutArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int N = a.n_envs, A = a.n_agents, B = N * A, d = a.obs_dim, n = a.n_actions;
    const int ldx = pad4(d) + 4;
    float* p = smem;
    SmemWeights w = carve_weights(p, d, false);
    float* Xs = p;  p += R_M * ldx;
    float* N1s = p; p += R_M * LDA;
    float* N3s = p; p += R_M * LDA;
    int* act_s = reinterpret_cast<int*>(p); p += R_M;

    // rows of this CTA: whole envs only
    const int envs_per_cta = R_M / A;
    const int env0 = blockIdx.x * envs_per_cta;
    const int n_env_here = min(envs_per_cta, N - env0);
    const int row0 = env0 * A;
    const int rows_here = n_env_here * A;
    const int tid = threadIdx.x;

    load_weights_folded<R_NT>(w, a.policy_params, d, n, false);   // (the logstd tail, if any, is read directly)

    // stage obs of slot t_begin (zero padding for the k tail and for idle rows)
    for (int i = tid; i < R_M * ldx; i += R_NT) {
        const int r = i / ldx, k = i % ldx;
        Xs[i] = (r < rows_here && k < d) ? a.policy_obs[((size_t)a.t_begin * B + row0 + r) * d + k] : 0.f;
    }
    __syncthreads();

    constexpr int PPR = R_NT / R_M;
    const int hrow = tid / PPR, hpart = tid % PPR;
    const uint64_t rng_base = a.rng_step_base + (a.rng_counter ? *a.rng_counter : 0ull);

    for (int t = a.t_begin; t < a.t_end; ++t) {
        float mu1[R_M / (R_NT / 16)], rstd1[R_M / (R_NT / 16)], rstd3[R_M / (R_NT / 16)];
        unsigned pm;
        trunk_forward<R_M, R_NT, false>(w, Xs, ldx, d, a.activation_id, N1s, N3s, mu1, rstd1, rstd3, pm);
        __syncthreads();
        float logit[MAX_OUT];
        head_dots<R_M, R_NT>(w, N3s, n, logit);
        if (hpart == 0 && hrow < rows_here && a.head_kind == ORL_HEAD_GAUSSIAN) {
            // DiagGaussian (distributions.py:75-98): action = noise*std + mean, per-dimension log-probs
            const size_t grow = (size_t)t * B + row0 + hrow;
            const float* logstd = a.policy_params + net_offsets(d, n, 1).ls;
            uint32_t rr[8];
            if (!a.exp_noise && !a.deterministic) {
                const uint64_t step = rng_base + (uint64_t)t;
                const uint2 key = make_uint2((uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
                const uint4 r0 = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(row0 + hrow + a.rng_row_offset), 2u), key);
                const uint4 r1 = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(row0 + hrow + a.rng_row_offset), 3u), key);
                const uint4 r2 = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(row0 + hrow + a.rng_row_offset), 4u), key);
                const uint4 r3 = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(row0 + hrow + a.rng_row_offset), 5u), key);
                // Box-Muller: 8 normals from 16 uniforms (pairs (r0,r1) and (r2,r3))
                const uint32_t u1[8] = {r0.x, r0.y, r0.z, r0.w, r2.x, r2.y, r2.z, r2.w};
                const uint32_t u2[8] = {r1.x, r1.y, r1.z, r1.w, r3.x, r3.y, r3.z, r3.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float rad = sqrtf(-2.0f * logf(u32_to_unit_open(u1[j])));
                    rr[j] = __float_as_uint(rad * cospif(2.0f * u32_to_unit_open(u2[j])));
                }
            }
#pragma unroll
            for (int j = 0; j < MAX_OUT; ++j) {
                if (j < n) {
                    const float mean = logit[j], ls = logstd[j], std = expf(ls);
                    float act = mean;
                    if (!a.deterministic) {
                        const float eps = a.exp_noise ? a.exp_noise[grow * n + j] : __uint_as_float(rr[j]);
                        act = __fadd_rn(__fmul_rn(eps, std), mean);
                    }
                    const float diff = act - mean;
                    // Normal.log_prob: -((x-mu)^2)/(2 var) - log(std) - log(sqrt(2 pi))
                    const float lp = -(diff * diff) / (2.0f * (std * std)) - ls - 0.9189385332046727f;
                    a.actions[grow * n + j] = act;
                    a.action_log_probs[grow * n + j] = lp;
                }
            }
        }
        if (hpart == 0 && hrow < rows_here && a.head_kind != ORL_HEAD_GAUSSIAN) {
            const size_t grow = (size_t)t * B + row0 + hrow;
            if (a.action_masks) {
#pragma unroll
                for (int j = 0; j < MAX_OUT; ++j)
                    if (j < n && a.action_masks[grow * n + j] == 0.f) logit[j] = -6e4f;
            }
            float nl[MAX_OUT], pr[MAX_OUT];
            log_softmax_n(logit, n, nl, pr);
            int act;
            if (a.deterministic) {
                act = 0;
#pragma unroll
                for (int j = 1; j < MAX_OUT; ++j) if (j < n && pr[j] > pr[act]) act = j;
            } else {
                float q[MAX_OUT];
                if (a.exp_noise) {
#pragma unroll
                    for (int j = 0; j < MAX_OUT; ++j) q[j] = (j < n) ? a.exp_noise[grow * n + j] : 1.f;
                } else {
                    const uint64_t step = rng_base + (uint64_t)t;
                    const uint2 key = make_uint2((uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
                    const uint4 r0 = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(row0 + hrow + a.rng_row_offset), 0u), key);
                    const uint4 r1 = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(row0 + hrow + a.rng_row_offset), 1u), key);
                    const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int j = 0; j < MAX_OUT; ++j) q[j] = -logf(u32_to_unit_open(rr[j]));
                }
                act = sample_categorical(pr, n, q);
            }
            float lp = nl[0];
#pragma unroll
            for (int j = 1; j < MAX_OUT; ++j) if (j == act) lp = nl[j];
            a.actions[grow] = (float)act;
            a.action_log_probs[grow] = lp;
            act_s[hrow] = act;
        }
        __syncthreads();

        // ---- env.step for the envs of this CTA (one thread per env) ----
        if (ENV == ORL_ENV_MPE_SPREAD && tid < n_env_here) {
          if constexpr (ENV == ORL_ENV_MPE_SPREAD) {
            const int e = env0 + tid;
            EnvPtrs E{a.env_f64, a.env_u64, a.env_i32, a.env_table, a.env_table_len, a.rng_seed,
                      a.ep_return, a.ep_length, a.episode_stats};
            const int acts[3] = {act_s[tid * 3], act_s[tid * 3 + 1], act_s[tid * 3 + 2]};
            float ob[3][18], reward; bool done;
            env_step_mpe(E, e, N, acts, ob, reward, done);
            const size_t r1 = (size_t)(t + 1) * B + (size_t)e * 3;
#pragma unroll
            for (int ag = 0; ag < 3; ++ag) {
#pragma unroll
                for (int k = 0; k < 18; ++k) {
                    Xs[(tid * 3 + ag) * ldx + k] = ob[ag][k];
                    a.policy_obs[(r1 + ag) * 18 + k] = ob[ag][k];
#pragma unroll
                    for (int dst = 0; dst < 3; ++dst) a.critic_obs[(r1 + dst) * 54 + ag * 18 + k] = ob[ag][k];
                }
                a.rewards[(size_t)t * B + (size_t)e * 3 + ag] = reward;
                a.masks[r1 + ag] = done ? 0.f : 1.f;
                a.active_masks[r1 + ag] = 1.f;  // all agents finish together (onpolicy_driver.py:118-124)
            }
          }
        } else if (ENV != ORL_ENV_NONE && ENV != ORL_ENV_MPE_SPREAD && tid < n_env_here) {
          if constexpr (ENV == ORL_ENV_CARTPOLE || ENV == ORL_ENV_GRIDWORLD) {
            const int e = env0 + tid;
            EnvPtrs E{a.env_f64, a.env_u64, a.env_i32, a.env_table, a.env_table_len, a.rng_seed,
                      a.ep_return, a.ep_length, a.episode_stats, a.rng_row_offset / max(a.n_agents, 1)};
            float ob[4], fin[4], reward; bool done;
            env_step_single(E, ENV, e, N, act_s[tid], ob, reward, done, fin);
            const size_t o1 = ((size_t)(t + 1) * B + e);
#pragma unroll
            for (int k = 0; k < 4; ++k) { Xs[tid * ldx + k] = ob[k]; a.policy_obs[o1 * 4 + k] = ob[k]; }
            a.rewards[(size_t)t * B + e] = reward;
            a.masks[o1] = done ? 0.f : 1.f;
            a.active_masks[o1] = 1.f;  // onpolicy_driver.py:118-124 with A == 1
          }
        }
        __syncthreads();
    }
}

__global__ void bump_counter_kernel(uint64_t* c, uint64_t by) { *c += by; }

__global__ void env_reset_kernel(int env_kind, int N, double* env_f64, uint64_t* env_u64, int32_t* env_i32,
                                 const int32_t* env_table, int env_table_len, uint64_t seed, float* obs_out,
                                 float* critic_obs_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    if (env_kind == ORL_ENV_MPE_SPREAD) {
        MpeState s;
        Pcg64 g = pcg_load(env_u64, e, N);
        mpe_reset(s, g);
        pcg_store(env_u64, e, N, g);
        mpe_store(env_f64, e, N, s);
        env_i32[e] = 0;
        for (int ag = 0; ag < 3; ++ag) {
            float o[18];
            mpe_obs(s, ag, o);
            for (int k = 0; k < 18; ++k) {
                obs_out[((size_t)e * 3 + ag) * 18 + k] = o[k];
                if (critic_obs_out) for (int dst = 0; dst < 3; ++dst) critic_obs_out[((size_t)e * 3 + dst) * 54 + ag * 18 + k] = o[k];
            }
        }
        return;
    }
    if (env_kind == ORL_ENV_CARTPOLE) {
        double s[4];
        Pcg64 g = pcg_load(env_u64, e, N);
        cartpole_reset(s, g);
        pcg_store(env_u64, e, N, g);
#pragma unroll
        for (int k = 0; k < 4; ++k) { env_f64[(size_t)k * N + e] = s[k]; obs_out[(size_t)e * 4 + k] = (float)s[k]; }
        env_i32[e] = 0;
    } else if (env_kind == ORL_ENV_GRIDWORLD) {
        int x, y, nreset = env_i32[3 * N + e];
        gridworld_reset(x, y, e, nreset, seed, env_table, env_table_len, 10, 10);
        env_i32[0 * N + e] = x; env_i32[1 * N + e] = y; env_i32[2 * N + e] = 0; env_i32[3 * N + e] = nreset + 1;
        obs_out[(size_t)e * 4 + 0] = (float)x; obs_out[(size_t)e * 4 + 1] = (float)y;
        obs_out[(size_t)e * 4 + 2] = 1.f; obs_out[(size_t)e * 4 + 3] = 1.f;
    }
}

// ---- batched critic forward --------------------------------------------------------------------
constexpr int C_M = 128, C_NT = 256;

__global__ void __launch_bounds__(C_NT) critic_values_kernel(const float* __restrict__ params, int d, int activation_id,
                                                             const float* __restrict__ obs, float* __restrict__ values,
                                                             long long rows) {
    extern __shared__ __align__(16) float smem[];
    const int ldx = pad4(d) + 4;
    float* p = smem;
    SmemWeights w = carve_weights(p, d, false);
    float* Xs = p;  p += C_M * ldx;
    float* N1s = p; p += C_M * LDA;
    float* N3s = p; p += C_M * LDA;
    const int tid = threadIdx.x;
    load_weights_folded<C_NT>(w, params, d, 1, false);
    const long long n_tiles = (rows + C_M - 1) / C_M;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long r0 = tile * C_M;
        const int rows_here = (int)min((long long)C_M, rows - r0);
        for (int i = tid; i < C_M * ldx; i += C_NT) {
            const int r = i / ldx, k = i % ldx;
            Xs[i] = (r < rows_here && k < d) ? obs[(r0 + r) * d + k] : 0.f;
        }
        __syncthreads();
        float mu1[C_M / (C_NT / 16)], rstd1[C_M / (C_NT / 16)], rstd3[C_M / (C_NT / 16)];
        unsigned pm;
        trunk_forward<C_M, C_NT, false>(w, Xs, ldx, d, activation_id, N1s, N3s, mu1, rstd1, rstd3, pm);
        __syncthreads();
        float out[MAX_OUT];
        head_dots<C_M, C_NT>(w, N3s, 1, out);
        constexpr int PPR = C_NT / C_M;
        if (tid % PPR == 0 && tid / PPR < rows_here) values[r0 + tid / PPR] = out[0];
        __syncthreads();
    }
}

// ---- insert of one host env.step into the rollout buffer (OnPolicyDriver.add2buffer, onpolicy_driver.py:80-152) -----------
// staged = [obs (B*d) | rewards (B) | dones (B)] as uploaded from the host in ONE copy; one thread per row.
__global__ void host_insert_kernel(const float* __restrict__ staged, int n_envs, int n_agents, int d, float* __restrict__ obs_next,
                                   float* __restrict__ rewards, float* __restrict__ masks_next, float* __restrict__ active_next) {
    const int B = n_envs * n_agents;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const float* so = staged;
    const float* sr = staged + (size_t)B * d;
    const float* sd = sr + B;
    for (int k = 0; k < d; ++k) obs_next[(size_t)r * d + k] = so[(size_t)r * d + k];
    rewards[r] = sr[r];
    const int e = r / n_agents;
    bool all_done = true;
    for (int a = 0; a < n_agents; ++a) all_done = all_done && (sd[e * n_agents + a] != 0.f);
    const bool done = sd[r] != 0.f;
    masks_next[r] = all_done ? 0.f : 1.f;                    // masks[dones_env] = 0
    active_next[r] = (done && !all_done) ? 0.f : 1.f;        // active[dones] = 0, active[dones_env] = 1
}

// ---- PolicyNetwork.eval_actions over a flat batch (policy_network.py:164-203, act.py:160-168 / 150-158) ----------------
// log-prob of the given action and the entropy of the action distribution per row (the caller takes the masked mean).
__global__ void __launch_bounds__(C_NT) policy_eval_kernel(const float* __restrict__ params, int d, int n, int activation_id, int head_kind,
                                                           const float* __restrict__ obs, const float* __restrict__ actions,
                                                           const float* __restrict__ action_masks, float* __restrict__ logp_out,
                                                           float* __restrict__ entropy_out, long long rows) {
    extern __shared__ __align__(16) float smem[];
    const int ldx = pad4(d) + 4;
    float* p = smem;
    SmemWeights w = carve_weights(p, d, false);
    float* Xs = p;  p += C_M * ldx;
    float* N1s = p; p += C_M * LDA;
    float* N3s = p; p += C_M * LDA;
    const int tid = threadIdx.x;
    load_weights_folded<C_NT>(w, params, d, n, false);
    const bool gaussian = head_kind == ORL_HEAD_GAUSSIAN;
    const float* logstd = params + net_offsets(d, n, 1).ls;
    const long long n_tiles = (rows + C_M - 1) / C_M;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long r0 = tile * C_M;
        const int rows_here = (int)min((long long)C_M, rows - r0);
        for (int i = tid; i < C_M * ldx; i += C_NT) {
            const int r = i / ldx, k = i % ldx;
            Xs[i] = (r < rows_here && k < d) ? obs[(r0 + r) * d + k] : 0.f;
        }
        __syncthreads();
        float mu1[C_M / (C_NT / 16)], rstd1[C_M / (C_NT / 16)], rstd3[C_M / (C_NT / 16)];
        unsigned pm;
        trunk_forward<C_M, C_NT, false>(w, Xs, ldx, d, activation_id, N1s, N3s, mu1, rstd1, rstd3, pm);
        __syncthreads();
        float out[MAX_OUT];
        head_dots<C_M, C_NT>(w, N3s, n, out);
        constexpr int PPR = C_NT / C_M;
        if (tid % PPR == 0 && tid / PPR < rows_here) {
            const long long g = r0 + tid / PPR;
            if (gaussian) {   // per-dimension log-probs and entropies (distributions.py:34-47)
                for (int j = 0; j < n; ++j) {
                    const float ls = logstd[j], std = expf(ls), diff = actions[g * n + j] - out[j];
                    logp_out[g * n + j] = -(diff * diff) / (2.0f * (std * std)) - ls - 0.9189385332046727f;
                    entropy_out[g * n + j] = 1.4189385332046727f + ls;
                }
            } else {
                if (action_masks) {
#pragma unroll
                    for (int j = 0; j < MAX_OUT; ++j) if (j < n && action_masks[g * n + j] == 0.f) out[j] = -6e4f;
                }
                float nl[MAX_OUT], pr[MAX_OUT];
                log_softmax_n(out, n, nl, pr);
                const int act = (int)actions[g];
                float lp = nl[0], ent = 0.f;
#pragma unroll
                for (int j = 0; j < MAX_OUT; ++j) if (j < n) { if (j == act) lp = nl[j]; ent -= pr[j] * nl[j]; }
                logp_out[g] = lp;
                entropy_out[g] = ent;
            }
        }
        __syncthreads();
    }
}

}  // namespace

namespace orl {
bool fwd_tc_enabled();
bool rollout_tc_eligible(const OrlRolloutArgs& a);
int launch_rollout_tc(const OrlRolloutArgs& a, cudaStream_t st);
int launch_critic_values_tc(const float* params, int d, int activation_id, const float* obs, float* values, long long rows, cudaStream_t st);
}  // namespace orl

extern "C" int orl_env_reset(int env_kind, int n_envs, int n_agents, double* env_f64, uint64_t* env_u64,
                             int32_t* env_i32, const int32_t* env_table, int env_table_len, uint64_t rng_seed,
                             float* policy_obs_out, float* critic_obs_out, void* stream) {
    ORL_CHECK_ARG(n_envs > 0 && n_agents > 0, "n_envs/n_agents");
    ORL_CHECK_ARG(policy_obs_out, "policy_obs_out");
    if (env_kind == ORL_ENV_CARTPOLE) ORL_CHECK_ARG(env_f64 && env_u64 && env_i32 && n_agents == 1, "cartpole state");
    else if (env_kind == ORL_ENV_GRIDWORLD) ORL_CHECK_ARG(env_i32 && n_agents == 1, "gridworld state");
    else if (env_kind == ORL_ENV_MPE_SPREAD) ORL_CHECK_ARG(env_f64 && env_u64 && env_i32 && n_agents == 3, "simple_spread state");
    else { orl::set_last_error("orl_env_reset: unsupported env_kind %d", env_kind); return ORL_ERR_UNSUPPORTED; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    env_reset_kernel<<<(n_envs + 127) / 128, 128, 0, st>>>(env_kind, n_envs, env_f64, env_u64, env_i32, env_table,
                                                          env_table_len, rng_seed, policy_obs_out, critic_obs_out);
    ORL_LAUNCH_CHECK("env_reset_kernel");
    return 0;
}


extern "C" int orl_env_step(int env_kind, int n_envs, int n_agents, double* env_f64, uint64_t* env_u64, int32_t* env_i32,
                            const int32_t* env_table, int env_table_len, uint64_t rng_seed, float* ep_return,
                            int32_t* ep_length, double* episode_stats, const float* actions, float* obs_out,
                            float* rewards_out, float* dones_out, float* final_obs_out, void* stream) {
    ORL_CHECK_ARG(n_envs > 0, "n_envs");
    ORL_CHECK_ARG(actions && obs_out && rewards_out && dones_out && ep_return && ep_length && episode_stats, "null buffer");
    if (env_kind == ORL_ENV_MPE_SPREAD) {
        ORL_CHECK_ARG(n_agents == 3 && env_f64 && env_u64 && env_i32, "simple_spread state");
        EnvPtrs Em{env_f64, env_u64, env_i32, env_table, env_table_len, rng_seed, ep_return, ep_length, episode_stats};
        // final_obs_out doubles as the critic observation output (B, 54) for this env kind
        env_step_mpe_kernel<<<(n_envs + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            n_envs, Em, actions, obs_out, final_obs_out, rewards_out, dones_out);
        ORL_LAUNCH_CHECK("env_step_mpe_kernel");
        return 0;
    }
    ORL_CHECK_ARG(n_agents == 1, "n_agents");
    if (env_kind == ORL_ENV_CARTPOLE) ORL_CHECK_ARG(env_f64 && env_u64 && env_i32, "cartpole state");
    else if (env_kind == ORL_ENV_GRIDWORLD) ORL_CHECK_ARG(env_i32, "gridworld state");
    else { orl::set_last_error("orl_env_step: unsupported env_kind %d", env_kind); return ORL_ERR_UNSUPPORTED; }
    EnvPtrs E{env_f64, env_u64, env_i32, env_table, env_table_len, rng_seed, ep_return, ep_length, episode_stats};
    env_step_kernel<<<(n_envs + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        env_kind, n_envs, E, actions, obs_out, rewards_out, dones_out, final_obs_out);
    ORL_LAUNCH_CHECK("env_step_kernel");
    return 0;
}

extern "C" int orl_rollout(const OrlRolloutArgs* args, void* stream) {
    ORL_CHECK_ARG(args, "args");
    const OrlRolloutArgs& a = *args;
    ORL_CHECK_ARG(a.n_envs > 0 && a.n_agents > 0 && a.n_agents <= 32, "n_envs / n_agents");
    ORL_CHECK_ARG(a.obs_dim > 0 && a.obs_dim <= 64, "obs_dim must be in 1..64");
    ORL_CHECK_ARG(a.n_actions > 0 && a.n_actions <= orl::MAX_OUT, "n_actions must be in 1..8");
    ORL_CHECK_ARG(a.t_begin >= 0 && a.t_begin < a.t_end && a.t_end <= a.episode_length, "step range");
    ORL_CHECK_ARG(a.activation_id >= 0 && a.activation_id <= 3, "activation_id");
    ORL_CHECK_ARG(a.policy_params && a.policy_obs && a.actions && a.action_log_probs, "null buffer");
    ORL_CHECK_ARG(a.head_kind == ORL_HEAD_CATEGORICAL || (a.head_kind == ORL_HEAD_GAUSSIAN && a.env_kind == ORL_ENV_NONE),
                  "Gaussian heads act on host-stepped envs (ORL_ENV_NONE)");
    if (a.env_kind == ORL_ENV_NONE) {
        ORL_CHECK_ARG(a.t_end == a.t_begin + 1, "ORL_ENV_NONE acts for one step per call");
    } else if (a.env_kind == ORL_ENV_CARTPOLE) {
        ORL_CHECK_ARG(a.n_agents == 1 && a.obs_dim == 4 && a.n_actions == 2, "CartPole shapes");
        ORL_CHECK_ARG(a.env_f64 && a.env_u64 && a.env_i32 && a.rewards && a.masks && a.active_masks && a.ep_return &&
                          a.ep_length && a.episode_stats, "CartPole state buffers");
    } else if (a.env_kind == ORL_ENV_GRIDWORLD) {
        ORL_CHECK_ARG(a.n_agents == 1 && a.obs_dim == 4 && a.n_actions == 5, "GridWorld shapes");
        ORL_CHECK_ARG(a.env_i32 && a.rewards && a.masks && a.active_masks && a.ep_return && a.ep_length &&
                          a.episode_stats, "GridWorld state buffers");
    } else if (a.env_kind == ORL_ENV_MPE_SPREAD) {
        ORL_CHECK_ARG(a.n_agents == 3 && a.obs_dim == 18 && a.critic_obs_dim == 54 && a.n_actions == 5, "simple_spread shapes");
        ORL_CHECK_ARG(a.env_f64 && a.env_u64 && a.env_i32 && a.rewards && a.masks && a.active_masks && a.ep_return &&
                          a.ep_length && a.episode_stats && a.critic_obs, "simple_spread state buffers");
    } else {
        orl::set_last_error("orl_rollout: unsupported env_kind %d", a.env_kind);
        return ORL_ERR_UNSUPPORTED;
    }
    if (orl::rollout_tc_eligible(a)) {   // single-agent device envs: the tensor-core rollout (orl_fwd_tc.cu), one CTA per 128 envs
        if (int e = orl::launch_rollout_tc(a, reinterpret_cast<cudaStream_t>(stream))) return e;
        if (a.rng_counter) {
            bump_counter_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a.rng_counter, (uint64_t)(a.t_end - a.t_begin));
            ORL_LAUNCH_CHECK("bump_counter_kernel");
        }
        return 0;
    }
    // rows per CTA: the step chain is latency-bound, so prefer many small CTAs (>= ~4 per SM) and
    // only grow the tile when there are enough rows to keep that many CTAs anyway
    const long long B = (long long)a.n_envs * a.n_agents;
    const long long want = 4LL * orl::sm_count();
    int rm = 32;
    if (B / 32 < want) rm = 16;
    if (B / 16 < want) rm = 8;
    while (rm < a.n_agents) rm *= 2;
    const int envs_per_cta = rm / a.n_agents;
    const int grid = (a.n_envs + envs_per_cta - 1) / envs_per_cta;
    const int ldx = orl::pad4(a.obs_dim) + 4;
    const size_t smem = sizeof(float) * (orl::smem_weights_floats(a.obs_dim, false) + rm * ldx + 2 * rm * orl::LDA + rm);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ORL_LAUNCH_ROLLOUT(RM, EK)                                                                                   \
    do {                                                                                                             \
        static bool attr_done = false;                                                                               \
        if (!attr_done) {                                                                                            \
            int e_ = orl::check_cuda(cudaFuncSetAttribute(rollout_kernel<RM, EK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "attr"); \
            if (e_) return e_;                                                                                       \
            attr_done = true;                                                                                        \
        }                                                                                                            \
        rollout_kernel<RM, EK><<<grid, R_NT, smem, st>>>(a);                                                         \
    } while (0)
#define ORL_LAUNCH_ROLLOUT_RM(EK)                                                        \
    do {                                                                                 \
        if (rm == 8) ORL_LAUNCH_ROLLOUT(8, EK);                                          \
        else if (rm == 16) ORL_LAUNCH_ROLLOUT(16, EK);                                   \
        else ORL_LAUNCH_ROLLOUT(32, EK);                                                 \
    } while (0)
    switch (a.env_kind) {
        case ORL_ENV_NONE: ORL_LAUNCH_ROLLOUT_RM(ORL_ENV_NONE); break;
        case ORL_ENV_CARTPOLE: ORL_LAUNCH_ROLLOUT_RM(ORL_ENV_CARTPOLE); break;
        case ORL_ENV_GRIDWORLD: ORL_LAUNCH_ROLLOUT_RM(ORL_ENV_GRIDWORLD); break;
        default: ORL_LAUNCH_ROLLOUT_RM(ORL_ENV_MPE_SPREAD); break;
    }
    ORL_LAUNCH_CHECK("rollout_kernel");
    if (a.rng_counter) {
        bump_counter_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a.rng_counter, (uint64_t)(a.t_end - a.t_begin));
        ORL_LAUNCH_CHECK("bump_counter_kernel");
    }
    return 0;
}

extern "C" int orl_critic_values(const float* critic_params, int obs_dim, int activation_id, const float* obs,
                                 float* values, long long rows, void* stream) {
    ORL_CHECK_ARG(critic_params && obs && values, "null buffer");
    ORL_CHECK_ARG(obs_dim > 0 && obs_dim <= 64, "obs_dim must be in 1..64");
    ORL_CHECK_ARG(rows > 0, "rows");
    ORL_CHECK_ARG(activation_id >= 0 && activation_id <= 3, "activation_id");
    if (orl::fwd_tc_enabled() && obs_dim <= 8)   // tensor-core forward (orl_fwd_tc.cu)
        return orl::launch_critic_values_tc(critic_params, obs_dim, activation_id, obs, values, rows, reinterpret_cast<cudaStream_t>(stream));
    const int ldx = orl::pad4(obs_dim) + 4;
    const size_t smem = sizeof(float) * (orl::smem_weights_floats(obs_dim, false) + C_M * ldx + 2 * C_M * orl::LDA);
    static bool attr_set = false;
    if (!attr_set) {
        int e = orl::check_cuda(cudaFuncSetAttribute(critic_values_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                                "cudaFuncSetAttribute(critic_values)");
        if (e) return e;
        attr_set = true;
    }
    const long long n_tiles = (rows + C_M - 1) / C_M;
    const int grid = (int)std::min<long long>(n_tiles, 2LL * orl::sm_count());
    critic_values_kernel<<<grid, C_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(critic_params, obs_dim, activation_id,
                                                                                      obs, values, rows);
    ORL_LAUNCH_CHECK("critic_values_kernel");
    return 0;
}

extern "C" int orl_policy_eval(const float* policy_params, int obs_dim, int n_actions, int activation_id, int head_kind,
                               const float* obs, const float* actions, const float* action_masks, float* log_probs,
                               float* entropy, long long rows, void* stream) {
    ORL_CHECK_ARG(policy_params && obs && actions && log_probs && entropy, "null buffer");
    ORL_CHECK_ARG(obs_dim > 0 && obs_dim <= 64 && n_actions > 0 && n_actions <= orl::MAX_OUT, "shapes");
    ORL_CHECK_ARG(rows > 0 && activation_id >= 0 && activation_id <= 3, "rows / activation_id");
    ORL_CHECK_ARG(head_kind == ORL_HEAD_CATEGORICAL || head_kind == ORL_HEAD_GAUSSIAN, "head_kind");
    const int ldx = orl::pad4(obs_dim) + 4;
    const size_t smem = sizeof(float) * (orl::smem_weights_floats(obs_dim, false) + C_M * ldx + 2 * C_M * orl::LDA);
    static bool attr_set = false;
    if (!attr_set) {
        int e = orl::check_cuda(cudaFuncSetAttribute(policy_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                                "cudaFuncSetAttribute(policy_eval)");
        if (e) return e;
        attr_set = true;
    }
    const long long n_tiles = (rows + C_M - 1) / C_M;
    const int grid = (int)std::min<long long>(n_tiles, 2LL * orl::sm_count());
    policy_eval_kernel<<<grid, C_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(policy_params, obs_dim, n_actions, activation_id, head_kind,
                                                                                    obs, actions, action_masks, log_probs, entropy, rows);
    ORL_LAUNCH_CHECK("policy_eval_kernel");
    return 0;
}

extern "C" int orl_host_insert(const float* staged, int n_envs, int n_agents, int obs_dim, float* policy_obs_next, float* rewards,
                               float* masks_next, float* active_masks_next, void* stream) {
    ORL_CHECK_ARG(staged && policy_obs_next && rewards && masks_next && active_masks_next, "null buffer");
    ORL_CHECK_ARG(n_envs > 0 && n_agents > 0 && obs_dim > 0, "shapes");
    const int B = n_envs * n_agents;
    host_insert_kernel<<<(B + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(staged, n_envs, n_agents, obs_dim, policy_obs_next,
                                                                                         rewards, masks_next, active_masks_next);
    ORL_LAUNCH_CHECK("host_insert_kernel");
    return 0;
}
