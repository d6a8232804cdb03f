// Warp-cooperative recurrent network step: ONE WARP per row, every 64-vector held as two registers per lane
// (elements `lane` and `lane + 32`), weights of one net staged once per CTA in shared memory as padded
// row-major matrices (leading dimension 65 floats, so that both W v — lane owns an output row — and
// W^T v — lane owns an input column — read conflict-free banks), mat-vecs by shuffle broadcast, LayerNorm
// by xor-shuffle reductions.  Nothing lives in local memory.
//
// Same math, same tape fields as the sequential restatement in orl_rnn_core.h (which the CPU test pins to
// the torch oracle); tests/debug_gru.py compares the tape of this implementation element-wise with the
// CPU core on the device's own rollout.  Summation orders differ (warp tree vs sequential): ~1e-7 relative.
//
// Tape row (floats): [0,1224) the P / Q / S fields of orl_rnn_core.h (inputs of the dW = sum P^T Q
// reductions), then the forward activations the backward step re-reads: A1 N1 N3 R Z NN GHN NO (64 each)
// and 8 scalars (rstd1, rstd3, rstd_rnn, mask, -).
#pragma once
#include "orl_mlp.cuh"
#include "orl_rnn_core.h"

namespace orl_rnnw {
using orl::warp_sum;
namespace rc = orl_rnn;

constexpr int H = 64, G3 = 192, LD = 65, MAXN = 8;
constexpr int TW_A1 = rc::TAPE, TW_N1 = TW_A1 + 64, TW_N3 = TW_N1 + 64, TW_R = TW_N3 + 64, TW_Z = TW_R + 64, TW_NN = TW_Z + 64,
              TW_GHN = TW_NN + 64, TW_NO = TW_GHN + 64, TW_SC = TW_NO + 64, TAPE_W = TW_SC + 8;   // 1744

struct V2 { float a, b; };

struct SmemNet {
    const float *w1, *w3, *wih, *whh, *wh;                                  // [rows][LD]
    const float *b1, *g1, *be1, *b3, *g3, *be3, *bih, *bhh, *gr, *ber, *bh;  // vectors
};
__host__ __device__ constexpr int smem_net_floats() { return (H + H + G3 + G3 + MAXN) * LD + 8 * H + 2 * G3 + MAXN; }

// Stage one net (flat layout of rc::rnn_offsets) into shared memory; W1 columns k >= d and head rows m >= n are zero.
__device__ inline SmemNet load_net(float* s, const float* __restrict__ P, const rc::Offsets& o, int tid, int nthreads) {
    float* w1 = s; s += H * LD;
    float* w3 = s; s += H * LD;
    float* wih = s; s += G3 * LD;
    float* whh = s; s += G3 * LD;
    float* wh = s; s += MAXN * LD;
    float* vec = s;   // b1 g1 be1 b3 g3 be3 gr ber (64 each) | bih bhh (192 each) | bh (8)
    for (int i = tid; i < H * LD; i += nthreads) {
        const int j = i / LD, k = i % LD;
        w1[i] = k < o.d ? P[o.w1 + j * o.d + k] : 0.f;
        w3[i] = k < H ? P[o.w3 + j * H + k] : 0.f;
    }
    for (int i = tid; i < G3 * LD; i += nthreads) {
        const int j = i / LD, k = i % LD;
        wih[i] = k < H ? P[o.wih + j * H + k] : 0.f;
        whh[i] = k < H ? P[o.whh + j * H + k] : 0.f;
    }
    for (int i = tid; i < MAXN * LD; i += nthreads) {
        const int j = i / LD, k = i % LD;
        wh[i] = (j < o.n && k < H) ? P[o.wh + j * H + k] : 0.f;
    }
    for (int i = tid; i < H; i += nthreads) {
        vec[i] = P[o.b1 + i]; vec[H + i] = P[o.g1 + i]; vec[2 * H + i] = P[o.be1 + i];
        vec[3 * H + i] = P[o.b3 + i]; vec[4 * H + i] = P[o.g3 + i]; vec[5 * H + i] = P[o.be3 + i];
        vec[6 * H + i] = P[o.gr + i]; vec[7 * H + i] = P[o.ber + i];
    }
    for (int i = tid; i < G3; i += nthreads) { vec[8 * H + i] = P[o.bih + i]; vec[8 * H + G3 + i] = P[o.bhh + i]; }
    for (int i = tid; i < MAXN; i += nthreads) vec[8 * H + 2 * G3 + i] = i < o.n ? P[o.bh + i] : 0.f;
    SmemNet n;
    n.w1 = w1; n.w3 = w3; n.wih = wih; n.whh = whh; n.wh = wh;
    n.b1 = vec; n.g1 = vec + H; n.be1 = vec + 2 * H; n.b3 = vec + 3 * H; n.g3 = vec + 4 * H; n.be3 = vec + 5 * H;
    n.gr = vec + 6 * H; n.ber = vec + 7 * H; n.bih = vec + 8 * H; n.bhh = vec + 8 * H + G3; n.bh = vec + 8 * H + 2 * G3;
    return n;
}

__device__ __forceinline__ V2 ldv(const float* p, int lane) { return V2{p[lane], p[lane + 32]}; }
__device__ __forceinline__ void stv(float* p, int lane, V2 v) { p[lane] = v.a; p[lane + 32] = v.b; }
__device__ __forceinline__ float bcast(V2 v, int k) {   // element k of a lane-distributed 64-vector (k uniform)
    return __shfl_sync(0xffffffffu, k < 32 ? v.a : v.b, k & 31);
}

// normalised = (v - mean) * rstd over the 64 elements
__device__ __forceinline__ V2 ln_fwd(V2 v, float& rstd) {
    const float m = warp_sum(v.a + v.b) * (1.f / H);
    const float da = v.a - m, db = v.b - m;
    const float q = warp_sum(da * da + db * db) * (1.f / H);
    rstd = 1.f / sqrtf(q + rc::LN_EPS);
    return V2{da * rstd, db * rstd};
}
__device__ __forceinline__ V2 ln_bwd(V2 dn, V2 n, float rstd) {
    const float s1 = warp_sum(dn.a + dn.b) * (1.f / H);
    const float s2 = warp_sum(dn.a * n.a + dn.b * n.b) * (1.f / H);
    return V2{rstd * (dn.a - s1 - n.a * s2), rstd * (dn.b - s1 - n.b * s2)};
}

// out_j = bias_j + sum_{k<K} W[j][k] v_k for the 64 rows of W starting at W (lane owns rows lane, lane+32)
__device__ __forceinline__ V2 matvec64(const float* W, const float* bias, V2 v, int K, int lane) {
    float s0 = bias[lane], s1 = bias[lane + 32];
    const float* r0 = W + lane * LD;
    const float* r1 = W + (lane + 32) * LD;
    const int k1 = K < 32 ? K : 32;
    for (int k = 0; k < k1; ++k) {
        const float vk = __shfl_sync(0xffffffffu, v.a, k);
        s0 = fmaf(r0[k], vk, s0); s1 = fmaf(r1[k], vk, s1);
    }
    for (int k = 32; k < K; ++k) {
        const float vk = __shfl_sync(0xffffffffu, v.b, k - 32);
        s0 = fmaf(r0[k], vk, s0); s1 = fmaf(r1[k], vk, s1);
    }
    return V2{s0, s1};
}
// the three gate blocks of a [192][64] matrix at once: out[g] (g = r, z, n), one shuffle per k for six FMAs
__device__ __forceinline__ void matvec192(const float* W, const float* bias, V2 v, int lane, V2 (&out)[3]) {
    float s[6];
#pragma unroll
    for (int g = 0; g < 3; ++g) { s[2 * g] = bias[g * H + lane]; s[2 * g + 1] = bias[g * H + lane + 32]; }
    const float* r = W + lane * LD;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
        const float vk = bcast(v, k);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            s[2 * g] = fmaf(r[(g * H) * LD + k], vk, s[2 * g]);
            s[2 * g + 1] = fmaf(r[(g * H + 32) * LD + k], vk, s[2 * g + 1]);
        }
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) out[g] = V2{s[2 * g], s[2 * g + 1]};
}
// dv_k = sum_{j<64} W[j][k] dz_j (lane owns columns lane, lane+32)
__device__ __forceinline__ V2 matvecT64(const float* W, V2 dz, int lane) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int j = 0; j < H; ++j) {
        const float dj = bcast(dz, j);
        s0 = fmaf(W[j * LD + lane], dj, s0); s1 = fmaf(W[j * LD + 32 + lane], dj, s1);
    }
    return V2{s0, s1};
}
__device__ __forceinline__ V2 matvecT192(const float* W, const V2 (&dz)[3], int lane) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll 4
        for (int j = 0; j < H; ++j) {
            const float dj = bcast(dz[g], j);
            s0 = fmaf(W[(g * H + j) * LD + lane], dj, s0); s1 = fmaf(W[(g * H + j) * LD + 32 + lane], dj, s1);
        }
    }
    return V2{s0, s1};
}

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// One forward step of a row.  x: observation (zero beyond d), h_in: hidden state, out[m] valid for m < n on
// every lane.  With `tape` the Q operands and the activations for the backward step go to the row's tape.
__device__ __forceinline__ void step_forward(const SmemNet& W, int d, int n, int act_id, V2 x, V2 h_in, float mask, V2& h_out,
                                             float (&out)[MAXN], float* tape, int lane) {
    V2 z1 = matvec64(W.w1, W.b1, x, d, lane);
    const V2 a1{rc::act_fwd(z1.a, act_id), rc::act_fwd(z1.b, act_id)};
    float rstd1, rstd3, rstdr;
    const V2 n1 = ln_fwd(a1, rstd1);
    const V2 g1 = ldv(W.g1, lane), be1 = ldv(W.be1, lane);
    const V2 y1{fmaf(n1.a, g1.a, be1.a), fmaf(n1.b, g1.b, be1.b)};
    const V2 z3 = matvec64(W.w3, W.b3, y1, H, lane);
    const V2 n3 = ln_fwd(z3, rstd3);
    const V2 g3 = ldv(W.g3, lane), be3 = ldv(W.be3, lane);
    const V2 y3{fmaf(n3.a, g3.a, be3.a), fmaf(n3.b, g3.b, be3.b)};
    const V2 hm{h_in.a * mask, h_in.b * mask};
    V2 gi[3], gh[3];
    matvec192(W.wih, W.bih, y3, lane, gi);
    matvec192(W.whh, W.bhh, hm, lane, gh);
    const V2 r{sigm(gi[0].a + gh[0].a), sigm(gi[0].b + gh[0].b)};
    const V2 z{sigm(gi[1].a + gh[1].a), sigm(gi[1].b + gh[1].b)};
    const V2 nn{tanhf(fmaf(r.a, gh[2].a, gi[2].a)), tanhf(fmaf(r.b, gh[2].b, gi[2].b))};
    h_out = V2{fmaf(z.a, hm.a, (1.f - z.a) * nn.a), fmaf(z.b, hm.b, (1.f - z.b) * nn.b)};
    const V2 no = ln_fwd(h_out, rstdr);
    const V2 gr = ldv(W.gr, lane), ber = ldv(W.ber, lane);
    const V2 o{fmaf(no.a, gr.a, ber.a), fmaf(no.b, gr.b, ber.b)};
#pragma unroll
    for (int m = 0; m < MAXN; ++m) {
        out[m] = 0.f;
        if (m < n) out[m] = W.bh[m] + warp_sum(fmaf(W.wh[m * LD + lane], o.a, W.wh[m * LD + 32 + lane] * o.b));
    }
    if (tape) {
        stv(tape + rc::TQ_X, lane, x); stv(tape + rc::TQ_Y1, lane, y1); stv(tape + rc::TQ_Y3, lane, y3);
        stv(tape + rc::TQ_HM, lane, hm); stv(tape + rc::TQ_O, lane, o);
        stv(tape + TW_A1, lane, a1); stv(tape + TW_N1, lane, n1); stv(tape + TW_N3, lane, n3);
        stv(tape + TW_R, lane, r); stv(tape + TW_Z, lane, z); stv(tape + TW_NN, lane, nn);
        stv(tape + TW_GHN, lane, gh[2]); stv(tape + TW_NO, lane, no);
        if (lane == 0) { tape[TW_SC] = rstd1; tape[TW_SC + 1] = rstd3; tape[TW_SC + 2] = rstdr; tape[TW_SC + 3] = mask; }
    }
}

// Backward of one row-step from its tape row (written by step_forward; dL/dout in TP_DLOG).  Returns
// dL/dh_in (already multiplied by the mask) and writes the P / S tape fields.
__device__ __forceinline__ V2 step_backward(const SmemNet& W, int n, int act_id, float* tape, V2 dh_next, int lane) {
    float dl[MAXN];
#pragma unroll
    for (int m = 0; m < MAXN; ++m) dl[m] = tape[rc::TP_DLOG + m];
    const float rstd1 = tape[TW_SC], rstd3 = tape[TW_SC + 1], rstdr = tape[TW_SC + 2], mask = tape[TW_SC + 3];
    V2 dov{0.f, 0.f};
#pragma unroll
    for (int m = 0; m < MAXN; ++m)
        if (m < n) { dov.a = fmaf(W.wh[m * LD + lane], dl[m], dov.a); dov.b = fmaf(W.wh[m * LD + 32 + lane], dl[m], dov.b); }
    const V2 no = ldv(tape + TW_NO, lane), gr = ldv(W.gr, lane);
    stv(tape + rc::TS_DONO, lane, V2{dov.a * no.a, dov.b * no.b});
    stv(tape + rc::TS_DO, lane, dov);
    V2 dh = ln_bwd(V2{dov.a * gr.a, dov.b * gr.b}, no, rstdr);
    dh.a += dh_next.a; dh.b += dh_next.b;
    const V2 r = ldv(tape + TW_R, lane), z = ldv(tape + TW_Z, lane), nn = ldv(tape + TW_NN, lane);
    const V2 ghn = ldv(tape + TW_GHN, lane), hm = ldv(tape + rc::TQ_HM, lane);
    V2 dgi[3], dgh[3];
    float dhm[2];
    {
        const float dhv[2] = {dh.a, dh.b}, rv[2] = {r.a, r.b}, zv[2] = {z.a, z.b}, nv[2] = {nn.a, nn.b};
        const float gv[2] = {ghn.a, ghn.b}, hv[2] = {hm.a, hm.b};
        float o_r[2], o_z[2], o_n[2], o_hn[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float dnn = dhv[e] * (1.f - zv[e]), dz = dhv[e] * (hv[e] - nv[e]);
            dhm[e] = dhv[e] * zv[e];
            const float dnpre = dnn * (1.f - nv[e] * nv[e]);
            const float dr = dnpre * gv[e];
            o_z[e] = dz * zv[e] * (1.f - zv[e]);
            o_r[e] = dr * rv[e] * (1.f - rv[e]);
            o_n[e] = dnpre;
            o_hn[e] = dnpre * rv[e];
        }
        dgi[0] = V2{o_r[0], o_r[1]}; dgi[1] = V2{o_z[0], o_z[1]}; dgi[2] = V2{o_n[0], o_n[1]};
        dgh[0] = dgi[0]; dgh[1] = dgi[1]; dgh[2] = V2{o_hn[0], o_hn[1]};
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) { stv(tape + rc::TP_DGI + g * H, lane, dgi[g]); stv(tape + rc::TP_DGH + g * H, lane, dgh[g]); }
    const V2 dy3 = matvecT192(W.wih, dgi, lane);
    const V2 t = matvecT192(W.whh, dgh, lane);
    const V2 dh_prev{(dhm[0] + t.a) * mask, (dhm[1] + t.b) * mask};
    const V2 n3 = ldv(tape + TW_N3, lane), g3 = ldv(W.g3, lane);
    stv(tape + rc::TS_DY3N3, lane, V2{dy3.a * n3.a, dy3.b * n3.b});
    stv(tape + rc::TS_DY3, lane, dy3);
    const V2 dz3 = ln_bwd(V2{dy3.a * g3.a, dy3.b * g3.b}, n3, rstd3);
    stv(tape + rc::TP_DZ3, lane, dz3);
    const V2 dy1 = matvecT64(W.w3, dz3, lane);
    const V2 n1 = ldv(tape + TW_N1, lane), g1 = ldv(W.g1, lane), a1 = ldv(tape + TW_A1, lane);
    stv(tape + rc::TS_DY1N1, lane, V2{dy1.a * n1.a, dy1.b * n1.b});
    stv(tape + rc::TS_DY1, lane, dy1);
    const V2 da1 = ln_bwd(V2{dy1.a * g1.a, dy1.b * g1.b}, n1, rstd1);
    stv(tape + rc::TP_DZ1, lane, V2{da1.a * rc::act_bwd_from_out(a1.a, act_id), da1.b * rc::act_bwd_from_out(a1.b, act_id)});
    return dh_prev;
}

}  // namespace orl_rnnw
