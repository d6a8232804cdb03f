// Loss-side scalar helpers shared by the feed-forward (orl_ppo.cu) and recurrent (orl_rnn.cu) PPO updates.
#pragma once
#include "orl_mlp.cuh"

namespace orl {

struct AdvNorm { float m0, s0, m1, s1; bool two_stage; };
__device__ __forceinline__ AdvNorm make_adv_norm(const double* __restrict__ gs, bool use_adv_normalize) {
    // ppo.py:402-409
    AdvNorm r;
    const double n_all = gs[ORL_GS_COUNT], n_act = gs[ORL_GS_ACT_COUNT];
    const double mean_all = gs[ORL_GS_ADV_SUM] / n_all;
    const double var_all = fmax(gs[ORL_GS_ADV_SQSUM] / n_all - mean_all * mean_all, 0.0);
    double mean_act = gs[ORL_GS_ADV_ACT_SUM] / n_act;
    const double var_act = fmax(gs[ORL_GS_ADV_ACT_SQSUM] / n_act - mean_act * mean_act, 0.0);
    double std_act = sqrt(var_act);
    r.two_stage = use_adv_normalize;
    r.m0 = 0.f; r.s0 = 1.f;
    if (use_adv_normalize) {
        const double s0 = (double)((float)sqrt(var_all)) + 1e-5;
        r.m0 = (float)mean_all;
        r.s0 = (float)s0;
        mean_act = (mean_act - mean_all) / s0;
        std_act = std_act / s0;
    }
    r.m1 = (float)mean_act;
    r.s1 = (float)((double)((float)std_act) + 1e-5);
    return r;
}
__device__ __forceinline__ float apply_adv_norm(const AdvNorm& r, float a) {
    if (r.two_stage) a = (a - r.m0) / r.s0;
    return (a - r.m1) / r.s1;
}

// ValueNorm.update (valuenorm.py:59-76) applied to the old state with this minibatch's moments.
__device__ __forceinline__ void vn_updated(const float* __restrict__ vn_state, const double* __restrict__ mb_stats,
                                           double batch_rows, double beta_d, float (&out)[3]) {
    const float bm = (float)(mb_stats[0] / batch_rows);
    const float bsq = (float)(mb_stats[1] / batch_rows);
    const float beta = (float)beta_d;
    const float omw = (float)(1.0 - beta_d);
    out[0] = __fadd_rn(__fmul_rn(vn_state[0], beta), __fmul_rn(bm, omw));
    out[1] = __fadd_rn(__fmul_rn(vn_state[1], beta), __fmul_rn(bsq, omw));
    out[2] = __fadd_rn(__fmul_rn(vn_state[2], beta), __fmul_rn(1.0f, omw));
}

__device__ __forceinline__ float huber(float e, float d) { return fabsf(e) <= d ? 0.5f * e * e : d * (fabsf(e) - 0.5f * d); }
__device__ __forceinline__ float huber_grad(float e, float d) { return fabsf(e) <= d ? e : (e > 0.f ? d : -d); }

// Clipped value loss of one row (ppo.py:344-386 cal_value_loss): returns the loss and d(loss)/d(value).
struct ValueTerm { float loss, dv; };
__device__ __forceinline__ ValueTerm value_term(float v, float vp, float target, float clip, float delta, int flags) {
    const float diff = v - vp;
    const float clipped = vp + fminf(fmaxf(diff, -clip), clip);
    const float e_c = target - clipped, e_o = target - v;
    const bool hub = flags & ORL_PPO_HUBER;
    const float l_c = hub ? huber(e_c, delta) : 0.5f * e_c * e_c;
    const float l_o = hub ? huber(e_o, delta) : 0.5f * e_o * e_o;
    const float gc = hub ? huber_grad(e_c, delta) : e_c;
    const float go = hub ? huber_grad(e_o, delta) : e_o;
    ValueTerm o; o.loss = l_o; o.dv = -go;
    if (flags & ORL_PPO_CLIP_VALUE) {
        const bool inrange = diff >= -clip && diff <= clip;
        const float dc = inrange ? -gc : 0.f;
        if (l_o > l_c) { o.loss = l_o; o.dv = -go; }
        else if (l_c > l_o) { o.loss = l_c; o.dv = dc; }
        else { o.loss = l_o; o.dv = 0.5f * (-go) + 0.5f * dc; }
    }
    return o;
}

}  // namespace orl
