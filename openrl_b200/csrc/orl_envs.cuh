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
                                                const int* __restrict__ table, int max_resets, int nrow, int ncol) {
    if (table) {
        const int k = min(nreset, max_resets - 1);
        x = table[(env * max_resets + k) * 2 + 0];
        y = table[(env * max_resets + k) * 2 + 1];
        return;
    }
    // rejection sampling like the reference, bounded
    uint4 c = make_uint4((uint32_t)env, (uint32_t)nreset, 0x47726964u, 0u);
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
