// Device-resident single-env step functions for the simple gym-class envs, with the reference's
// vectorised-env conventions (auto-reset, done = terminated|truncated, per-env RNG streams).
//
//   CartPole-v1   gymnasium 0.29 classic_control/cartpole.py (third-party; restated in
//                 oracle/cartpole_ref.py), TimeLimit 500, reached through
//                 openrl/envs/gymnasium/__init__.py:27-54 + sync_venv.py:178-247.
//   GridWorldEnv  openrl/envs/gridworld/gridworld_env.py:21-86 (10x10 through make()).
//
// Reset randomness: every env owns a PCG64 stream bit-compatible with numpy's
// Generator(PCG64(SeedSequence(seed + i*10086))) (sync_venv.py:137,
// gymnasium/utils/seeding.py), so CartPole reset states equal the reference's draw for draw.
#pragma once
#include "orl_common.cuh"

namespace orl {

// ---- numpy-compatible PCG64 (setseq 128, XSL-RR 64 output) --------------------------------
struct Pcg64 {
    unsigned __int128 state, inc;
};
__device__ __forceinline__ Pcg64 pcg_load(const uint64_t* __restrict__ u, int i, int N) {
    Pcg64 g;
    g.state = ((unsigned __int128)u[0 * N + i] << 64) | u[1 * N + i];
    g.inc = ((unsigned __int128)u[2 * N + i] << 64) | u[3 * N + i];
    return g;
}
__device__ __forceinline__ void pcg_store(uint64_t* __restrict__ u, int i, int N, const Pcg64& g) {
    u[0 * N + i] = (uint64_t)(g.state >> 64);
    u[1 * N + i] = (uint64_t)g.state;
}
__device__ __forceinline__ uint64_t pcg_next64(Pcg64& g) {
    const unsigned __int128 mult = ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    g.state = g.state * mult + g.inc;
    const uint64_t hi = (uint64_t)(g.state >> 64), lo = (uint64_t)g.state;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}
__device__ __forceinline__ double pcg_next_double(Pcg64& g) {
    return (double)(pcg_next64(g) >> 11) * (1.0 / 9007199254740992.0);
}
// Generator.uniform(low, high): low + (high - low) * next_double, unfused
__device__ __forceinline__ double pcg_uniform(Pcg64& g, double low, double range) {
    return __dadd_rn(low, __dmul_rn(range, pcg_next_double(g)));
}

// ---- Philox4x32-10 (fast-mode sampling noise and GridWorld resets) --------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u32_to_unit_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// ---- CartPole-v1 -------------------------------------------------------------------------
// env_f64: [4][N] state (x, x_dot, theta, theta_dot);  env_i32: [0][N] elapsed steps
struct CartPoleOut { float obs[4]; float reward; bool done; };

__device__ __forceinline__ void cartpole_reset(double (&s)[4], Pcg64& g) {
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = pcg_uniform(g, -0.05, 0.05 - (-0.05));
}

__device__ __forceinline__ bool cartpole_dynamics(double (&s)[4], int action) {
    const double gravity = 9.8, masscart = 1.0, masspole = 0.1, total_mass = masspole + masscart, length = 0.5,
                 polemass_length = masspole * length, force_mag = 10.0, tau = 0.02;
    const double theta_thr = 12 * 2 * 3.141592653589793 / 360, x_thr = 2.4;
    double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    const double force = action == 1 ? force_mag : -force_mag;
    const double costheta = cos(theta), sintheta = sin(theta);
    const double temp = __ddiv_rn(
        __dadd_rn(force, __dmul_rn(__dmul_rn(polemass_length, __dmul_rn(theta_dot, theta_dot)), sintheta)),
        total_mass);
    const double thetaacc = __ddiv_rn(
        __dsub_rn(__dmul_rn(gravity, sintheta), __dmul_rn(costheta, temp)),
        __dmul_rn(length, __dsub_rn(4.0 / 3.0, __ddiv_rn(__dmul_rn(masspole, __dmul_rn(costheta, costheta)), total_mass))));
    const double xacc = __dsub_rn(temp, __ddiv_rn(__dmul_rn(__dmul_rn(polemass_length, thetaacc), costheta), total_mass));
    x = __dadd_rn(x, __dmul_rn(tau, x_dot));
    x_dot = __dadd_rn(x_dot, __dmul_rn(tau, xacc));
    theta = __dadd_rn(theta, __dmul_rn(tau, theta_dot));
    theta_dot = __dadd_rn(theta_dot, __dmul_rn(tau, thetaacc));
    s[0] = x; s[1] = x_dot; s[2] = theta; s[3] = theta_dot;
    return x < -x_thr || x > x_thr || theta < -theta_thr || theta > theta_thr;
}

// ---- GridWorld -----------------------------------------------------------------------------
// env_i32: [0][N] x, [1][N] y, [2][N] steps, [3][N] resets done so far.
// Resets: the reference draws from the process-global MT19937 in env order
// (gridworld_env.py:76-81), which cannot be reproduced by independent device streams; the start
// cell is drawn uniformly over the 99 non-goal cells from Philox keyed by (seed, env, #reset), or
// read from a host table reset_table[(env*max_resets + k)*2] when one is supplied (parity runs).
__device__ __forceinline__ void gridworld_reset(int& x, int& y, int env, int nreset, uint64_t seed,
                                                const int* __restrict__ table, int max_resets, int nrow, int ncol,
                                                int env_key = -1) {   // env_key: GLOBAL env index for the Philox counter (default: env)
    if (table) {
        const int k = min(nreset, max_resets - 1);
        x = table[(env * max_resets + k) * 2 + 0];
        y = table[(env * max_resets + k) * 2 + 1];
        return;
    }
    // rejection sampling like the reference, bounded
    uint4 c = make_uint4((uint32_t)(env_key >= 0 ? env_key : env), (uint32_t)nreset, 0x47726964u, 0u);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int it = 0; it < 16; ++it) {
        c.w = it;
        const uint4 r = philox4x32_10(c, key);
        x = (int)(((uint64_t)r.x * (uint64_t)nrow) >> 32);
        y = (int)(((uint64_t)r.y * (uint64_t)ncol) >> 32);
        if (!(x == 1 && y == 1)) return;
    }
    x = 0; y = 0;
}

}  // namespace orl

// ---- MPE simple_spread (3 agents, 3 landmarks) ------------------------------------------------
// Reference: World.step + forces + integrate  openrl/envs/mpe/core.py:216-344,
// MultiAgentEnv.step/_set_action/reset        openrl/envs/mpe/multiagent_env.py:167-243,274-339,
// Scenario.reset_world/reward/observation     openrl/envs/mpe/scenarios/simple_spread.py:46-125.
// float64 state like the reference (numpy): env_f64 [18][N] = agent pos (3x2), agent vel (3x2),
// landmark pos (3x2); env_i32 [N] = current_step; env_u64 = the env's PCG64 np_random.
// One thread steps one env (its 3 agents).
namespace orl {

struct MpeState { double pos[3][2], vel[3][2], lm[3][2]; };

__device__ __forceinline__ void mpe_load(const double* __restrict__ f, int e, int N, MpeState& s) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            s.pos[a][c] = f[(size_t)(a * 2 + c) * N + e];
            s.vel[a][c] = f[(size_t)(6 + a * 2 + c) * N + e];
            s.lm[a][c] = f[(size_t)(12 + a * 2 + c) * N + e];
        }
}
__device__ __forceinline__ void mpe_store(double* __restrict__ f, int e, int N, const MpeState& s) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            f[(size_t)(a * 2 + c) * N + e] = s.pos[a][c];
            f[(size_t)(6 + a * 2 + c) * N + e] = s.vel[a][c];
            f[(size_t)(12 + a * 2 + c) * N + e] = s.lm[a][c];
        }
}
__device__ __forceinline__ void mpe_reset(MpeState& s, Pcg64& g) {  // simple_spread.py:46-62
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        s.pos[a][0] = pcg_uniform(g, -1.0, 2.0); s.pos[a][1] = pcg_uniform(g, -1.0, 2.0);
        s.vel[a][0] = 0.0; s.vel[a][1] = 0.0;
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        s.lm[l][0] = __dmul_rn(0.8, pcg_uniform(g, -1.0, 2.0)); s.lm[l][1] = __dmul_rn(0.8, pcg_uniform(g, -1.0, 2.0));
    }
}
__device__ __forceinline__ double mpe_dist(const double (&p)[2], const double (&q)[2]) {
    const double dx = __dsub_rn(p[0], q[0]), dy = __dsub_rn(p[1], q[1]);
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
// numpy.logaddexp(0, v)
__device__ __forceinline__ double logaddexp0(double v) {
    if (v == 0.0) return 0.6931471805599453;
    return v < 0.0 ? log1p(exp(v)) : __dadd_rn(v, log1p(exp(-v)));
}
__device__ __forceinline__ void mpe_world_step(MpeState& s, const int (&act)[3]) {
    const double dt = 0.1, damping = 0.25, contact_force = 1e2, k = 1e-3, dist_min = 0.15 + 0.15;
    double f[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {  // _set_action + apply_action_force (mass 1, sensitivity 5)
        const double u0 = (act[a] == 1 ? 1.0 : 0.0) - (act[a] == 2 ? 1.0 : 0.0);
        const double u1 = (act[a] == 3 ? 1.0 : 0.0) - (act[a] == 4 ? 1.0 : 0.0);
        f[a][0] = __dmul_rn(u0, 5.0); f[a][1] = __dmul_rn(u1, 5.0);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a + 1; b < 3; ++b) {  // apply_environment_force: agent-agent contacts only
            const double dx = __dsub_rn(s.pos[a][0], s.pos[b][0]), dy = __dsub_rn(s.pos[a][1], s.pos[b][1]);
            const double dist = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            const double pen = __dmul_rn(logaddexp0(__ddiv_rn(-__dsub_rn(dist, dist_min), k)), k);
            const double fx = __dmul_rn(__ddiv_rn(__dmul_rn(contact_force, dx), dist), pen);
            const double fy = __dmul_rn(__ddiv_rn(__dmul_rn(contact_force, dy), dist), pen);
            f[a][0] = __dadd_rn(fx, f[a][0]); f[a][1] = __dadd_rn(fy, f[a][1]);
            f[b][0] = __dadd_rn(-fx, f[b][0]); f[b][1] = __dadd_rn(-fy, f[b][1]);
        }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // integrate_state
            double v = __dmul_rn(s.vel[a][c], 1.0 - damping);
            v = __dadd_rn(v, __dmul_rn(f[a][c], dt));
            s.vel[a][c] = v;
            s.pos[a][c] = __dadd_rn(s.pos[a][c], __dmul_rn(v, dt));
        }
}
// shared reward: sum over agents of (-sum_l min_a dist(a,l) - #collisions incl. self)
__device__ __forceinline__ double mpe_shared_reward(const MpeState& s) {
    double lm_term = 0.0;  // rew -= min(dists) per landmark, starting from integer 0
    double rew_base;
    {
        double r = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const double d0 = mpe_dist(s.pos[0], s.lm[l]), d1 = mpe_dist(s.pos[1], s.lm[l]), d2 = mpe_dist(s.pos[2], s.lm[l]);
            r = __dsub_rn(r, fmin(fmin(d0, d1), d2));
        }
        rew_base = r;
    }
    (void)lm_term;
    double total = 0.0;
#pragma unroll
    for (int ag = 0; ag < 3; ++ag) {
        double r = rew_base;
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (mpe_dist(s.pos[a], s.pos[ag]) < 0.3) r = __dsub_rn(r, 1.0);
        total = (ag == 0) ? r : __dadd_rn(total, r);
    }
    return total;
}
// observation of agent a (18 values) as float32
__device__ __forceinline__ void mpe_obs(const MpeState& s, int a, float (&o)[18]) {
    o[0] = (float)s.vel[a][0]; o[1] = (float)s.vel[a][1]; o[2] = (float)s.pos[a][0]; o[3] = (float)s.pos[a][1];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        o[4 + 2 * l] = (float)__dsub_rn(s.lm[l][0], s.pos[a][0]);
        o[5 + 2 * l] = (float)__dsub_rn(s.lm[l][1], s.pos[a][1]);
    }
    int w = 10;
#pragma unroll
    for (int ot = 0; ot < 3; ++ot) {
        if (ot != a) {
            o[w] = (float)__dsub_rn(s.pos[ot][0], s.pos[a][0]);
            o[w + 1] = (float)__dsub_rn(s.pos[ot][1], s.pos[a][1]);
            w += 2;
        }
    }
    o[14] = 0.f; o[15] = 0.f; o[16] = 0.f; o[17] = 0.f;
}

}  // namespace orl
