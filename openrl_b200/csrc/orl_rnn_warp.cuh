// Warp-cooperative recurrent network step: ONE WARP per row, every 64-vector held as two registers per lane
// (elements `lane` and `lane + 32`), weights of one net staged once per CTA in shared memory as padded
// row-major matrices (leading dimension 65 floats, so that both W v — lane owns an output row — and
// W^T v — lane owns an input column — read conflict-free banks).  A warp advances R rows together: each weight
// read from shared memory feeds R FMAs (the mat-vecs are shared-memory-bandwidth bound, 4 B of weight per
// lane-FMA at R = 1), and the input vector of a mat-vec is parked in a per-warp shared scratch and read back as
// broadcast float4 (one wavefront per four inputs) instead of one shuffle per input.  LayerNorm by xor-shuffle
// reductions.  Nothing lives in local memory.
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

constexpr int SCR = 256;   // per-row scratch floats in shared memory: slot A [0,64), slot B [64,256)
__host__ __device__ constexpr int smem_scratch_floats(int warps, int R) { return warps * R * SCR; }

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

// Park the R lane-distributed 64-vectors in scratch slot `off` (row r at scr + r*SCR + off).
template <int R>
__device__ __forceinline__ void park(float* scr, int off, const V2 (&v)[R], int lane) {
    __syncwarp();   // every lane is done reading what the slot held before
#pragma unroll
    for (int r = 0; r < R; ++r) stv(scr + r * SCR + off, lane, v[r]);
    __syncwarp();
}

// out[r]_j = bias_j + sum_{k<K4} W[j][k] * in[r][k]; lane owns rows j = lane, lane+32 of the 64-row block at W;
// in = scratch slot (K4 = K rounded up to 4: the inputs / weight columns beyond K are zero).
template <int R>
__device__ __forceinline__ void matvec64(const float* W, const float* bias, const float* scr, int off, int K, int lane, V2 (&out)[R]) {
    float s0[R], s1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { s0[r] = bias[lane]; s1[r] = bias[lane + 32]; }
    const float* r0 = W + lane * LD;
    const float* r1 = W + (lane + 32) * LD;
    const int K4 = (K + 3) & ~3;
    for (int k = 0; k < K4; k += 4) {
        const float w00 = r0[k], w01 = r0[k + 1], w02 = r0[k + 2], w03 = r0[k + 3];
        const float w10 = r1[k], w11 = r1[k + 1], w12 = r1[k + 2], w13 = r1[k + 3];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(scr + r * SCR + off + k);
            s0[r] = fmaf(w00, v.x, s0[r]); s0[r] = fmaf(w01, v.y, s0[r]); s0[r] = fmaf(w02, v.z, s0[r]); s0[r] = fmaf(w03, v.w, s0[r]);
            s1[r] = fmaf(w10, v.x, s1[r]); s1[r] = fmaf(w11, v.y, s1[r]); s1[r] = fmaf(w12, v.z, s1[r]); s1[r] = fmaf(w13, v.w, s1[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = V2{s0[r], s1[r]};
}
// the three gate blocks (r, z, n) of a [192][64] matrix: three 64-row passes over the same parked input
template <int R>
__device__ __forceinline__ void matvec192(const float* W, const float* bias, const float* scr, int off, int lane, V2 (&out)[3][R]) {
#pragma unroll
    for (int g = 0; g < 3; ++g) matvec64<R>(W + g * H * LD, bias + g * H, scr, off, H, lane, out[g]);
}
// out[r]_k = sum_{j<J} W[j][k] * in[r][j]; lane owns columns k = lane, lane+32; in = scratch slot of J floats
template <int R>
__device__ __forceinline__ void matvecT(const float* W, const float* scr, int off, int J, int lane, V2 (&out)[R]) {
    float s0[R], s1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { s0[r] = 0.f; s1[r] = 0.f; }
    for (int j = 0; j < J; j += 4) {
        const float* w = W + j * LD + lane;
        const float w00 = w[0], w01 = w[LD], w02 = w[2 * LD], w03 = w[3 * LD];
        const float w10 = w[32], w11 = w[LD + 32], w12 = w[2 * LD + 32], w13 = w[3 * LD + 32];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(scr + r * SCR + off + j);
            s0[r] = fmaf(w00, v.x, s0[r]); s0[r] = fmaf(w01, v.y, s0[r]); s0[r] = fmaf(w02, v.z, s0[r]); s0[r] = fmaf(w03, v.w, s0[r]);
            s1[r] = fmaf(w10, v.x, s1[r]); s1[r] = fmaf(w11, v.y, s1[r]); s1[r] = fmaf(w12, v.z, s1[r]); s1[r] = fmaf(w13, v.w, s1[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = V2{s0[r], s1[r]};
}

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// One forward step of R rows.  x: observation (zero beyond d), h_in: hidden state, out[r][m] valid for m < n on
// every lane.  tape[r] != nullptr: the Q operands and the activations for the backward step go to the row's tape.
// scr: this warp's scratch (R * SCR floats of shared memory).
template <int R>
__device__ __forceinline__ void step_forward(const SmemNet& W, float* scr, int d, int n, int act_id, const V2 (&x)[R], const V2 (&h_in)[R],
                                             const float (&mask)[R], V2 (&h_out)[R], float (&out)[R][MAXN], float* const (&tape)[R],
                                             int lane) {
    const V2 g1 = ldv(W.g1, lane), be1 = ldv(W.be1, lane), g3 = ldv(W.g3, lane), be3 = ldv(W.be3, lane);
    V2 t0[R], y1[R], y3[R], hm[R];
    park<R>(scr, 0, x, lane);
    matvec64<R>(W.w1, W.b1, scr, 0, d, lane, t0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const V2 a1{rc::act_fwd(t0[r].a, act_id), rc::act_fwd(t0[r].b, act_id)};
        float rstd1;
        const V2 n1 = ln_fwd(a1, rstd1);
        y1[r] = V2{fmaf(n1.a, g1.a, be1.a), fmaf(n1.b, g1.b, be1.b)};
        if (tape[r]) {
            stv(tape[r] + rc::TQ_X, lane, x[r]); stv(tape[r] + TW_A1, lane, a1); stv(tape[r] + TW_N1, lane, n1);
            stv(tape[r] + rc::TQ_Y1, lane, y1[r]);
            if (lane == 0) { tape[r][TW_SC] = rstd1; tape[r][TW_SC + 3] = mask[r]; }
        }
    }
    park<R>(scr, 0, y1, lane);
    matvec64<R>(W.w3, W.b3, scr, 0, H, lane, t0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float rstd3;
        const V2 n3 = ln_fwd(t0[r], rstd3);
        y3[r] = V2{fmaf(n3.a, g3.a, be3.a), fmaf(n3.b, g3.b, be3.b)};
        hm[r] = V2{h_in[r].a * mask[r], h_in[r].b * mask[r]};
        if (tape[r]) {
            stv(tape[r] + TW_N3, lane, n3); stv(tape[r] + rc::TQ_Y3, lane, y3[r]); stv(tape[r] + rc::TQ_HM, lane, hm[r]);
            if (lane == 0) tape[r][TW_SC + 1] = rstd3;
        }
    }
    V2 gi[3][R], gh[3][R];
    park<R>(scr, 0, y3, lane);
    matvec192<R>(W.wih, W.bih, scr, 0, lane, gi);
    park<R>(scr, 0, hm, lane);
    matvec192<R>(W.whh, W.bhh, scr, 0, lane, gh);
    const V2 gr = ldv(W.gr, lane), ber = ldv(W.ber, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const V2 rg{sigm(gi[0][r].a + gh[0][r].a), sigm(gi[0][r].b + gh[0][r].b)};
        const V2 z{sigm(gi[1][r].a + gh[1][r].a), sigm(gi[1][r].b + gh[1][r].b)};
        const V2 nn{tanhf(fmaf(rg.a, gh[2][r].a, gi[2][r].a)), tanhf(fmaf(rg.b, gh[2][r].b, gi[2][r].b))};
        h_out[r] = V2{fmaf(z.a, hm[r].a, (1.f - z.a) * nn.a), fmaf(z.b, hm[r].b, (1.f - z.b) * nn.b)};
        float rstdr;
        const V2 no = ln_fwd(h_out[r], rstdr);
        const V2 o{fmaf(no.a, gr.a, ber.a), fmaf(no.b, gr.b, ber.b)};
#pragma unroll
        for (int m = 0; m < MAXN; ++m) {
            out[r][m] = 0.f;
            if (m < n) out[r][m] = W.bh[m] + warp_sum(fmaf(W.wh[m * LD + lane], o.a, W.wh[m * LD + 32 + lane] * o.b));
        }
        if (tape[r]) {
            stv(tape[r] + rc::TQ_O, lane, o); stv(tape[r] + TW_R, lane, rg); stv(tape[r] + TW_Z, lane, z);
            stv(tape[r] + TW_NN, lane, nn); stv(tape[r] + TW_GHN, lane, gh[2][r]); stv(tape[r] + TW_NO, lane, no);
            if (lane == 0) tape[r][TW_SC + 2] = rstdr;
        }
    }
}

// Backward of one step of R rows from their tape rows (written by step_forward; dL/dout in TP_DLOG).  dh holds
// dL/dh_out arriving from the following step on entry and dL/dh_in (already multiplied by the mask) on exit;
// the P / S tape fields are written.
template <int R>
__device__ __forceinline__ void step_backward(const SmemNet& W, float* scr, int n, int act_id, float* const (&tape)[R], V2 (&dh)[R], int lane) {
    const V2 gr = ldv(W.gr, lane), g3 = ldv(W.g3, lane), g1 = ldv(W.g1, lane);
    V2 dgi[3][R], dgh[3][R], dhm[R];
    float mask[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float dl[MAXN];
#pragma unroll
        for (int m = 0; m < MAXN; ++m) dl[m] = tape[r][rc::TP_DLOG + m];
        const float rstdr = tape[r][TW_SC + 2];
        mask[r] = tape[r][TW_SC + 3];
        V2 dov{0.f, 0.f};
#pragma unroll
        for (int m = 0; m < MAXN; ++m)
            if (m < n) { dov.a = fmaf(W.wh[m * LD + lane], dl[m], dov.a); dov.b = fmaf(W.wh[m * LD + 32 + lane], dl[m], dov.b); }
        const V2 no = ldv(tape[r] + TW_NO, lane);
        stv(tape[r] + rc::TS_DONO, lane, V2{dov.a * no.a, dov.b * no.b});
        stv(tape[r] + rc::TS_DO, lane, dov);
        V2 d = ln_bwd(V2{dov.a * gr.a, dov.b * gr.b}, no, rstdr);
        d.a += dh[r].a; d.b += dh[r].b;
        const V2 rg = ldv(tape[r] + TW_R, lane), z = ldv(tape[r] + TW_Z, lane), nn = ldv(tape[r] + TW_NN, lane);
        const V2 ghn = ldv(tape[r] + TW_GHN, lane), hm = ldv(tape[r] + rc::TQ_HM, lane);
        const float dhv[2] = {d.a, d.b}, rv[2] = {rg.a, rg.b}, zv[2] = {z.a, z.b}, nv[2] = {nn.a, nn.b};
        const float gv[2] = {ghn.a, ghn.b}, hv[2] = {hm.a, hm.b};
        float o_r[2], o_z[2], o_n[2], o_hn[2], o_hm[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float dnn = dhv[e] * (1.f - zv[e]), dz = dhv[e] * (hv[e] - nv[e]);
            o_hm[e] = dhv[e] * zv[e];
            const float dnpre = dnn * (1.f - nv[e] * nv[e]);
            const float dr = dnpre * gv[e];
            o_z[e] = dz * zv[e] * (1.f - zv[e]);
            o_r[e] = dr * rv[e] * (1.f - rv[e]);
            o_n[e] = dnpre;
            o_hn[e] = dnpre * rv[e];
        }
        dgi[0][r] = V2{o_r[0], o_r[1]}; dgi[1][r] = V2{o_z[0], o_z[1]}; dgi[2][r] = V2{o_n[0], o_n[1]};
        dgh[0][r] = dgi[0][r]; dgh[1][r] = dgi[1][r]; dgh[2][r] = V2{o_hn[0], o_hn[1]};
        dhm[r] = V2{o_hm[0], o_hm[1]};
#pragma unroll
        for (int g = 0; g < 3; ++g) { stv(tape[r] + rc::TP_DGI + g * H, lane, dgi[g][r]); stv(tape[r] + rc::TP_DGH + g * H, lane, dgh[g][r]); }
    }
    V2 dy3[R], t[R], dz3[R], dy1[R];
#pragma unroll
    for (int g = 0; g < 3; ++g) park<R>(scr, 64 + g * H, dgi[g], lane);
    matvecT<R>(W.wih, scr, 64, G3, lane, dy3);
#pragma unroll
    for (int g = 0; g < 3; ++g) park<R>(scr, 64 + g * H, dgh[g], lane);
    matvecT<R>(W.whh, scr, 64, G3, lane, t);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        dh[r] = V2{(dhm[r].a + t[r].a) * mask[r], (dhm[r].b + t[r].b) * mask[r]};
        const float rstd3 = tape[r][TW_SC + 1];
        const V2 n3 = ldv(tape[r] + TW_N3, lane);
        stv(tape[r] + rc::TS_DY3N3, lane, V2{dy3[r].a * n3.a, dy3[r].b * n3.b});
        stv(tape[r] + rc::TS_DY3, lane, dy3[r]);
        dz3[r] = ln_bwd(V2{dy3[r].a * g3.a, dy3[r].b * g3.b}, n3, rstd3);
        stv(tape[r] + rc::TP_DZ3, lane, dz3[r]);
    }
    park<R>(scr, 0, dz3, lane);
    matvecT<R>(W.w3, scr, 0, H, lane, dy1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float rstd1 = tape[r][TW_SC];
        const V2 n1 = ldv(tape[r] + TW_N1, lane), a1 = ldv(tape[r] + TW_A1, lane);
        stv(tape[r] + rc::TS_DY1N1, lane, V2{dy1[r].a * n1.a, dy1[r].b * n1.b});
        stv(tape[r] + rc::TS_DY1, lane, dy1[r]);
        const V2 da1 = ln_bwd(V2{dy1[r].a * g1.a, dy1[r].b * g1.b}, n1, rstd1);
        stv(tape[r] + rc::TP_DZ1, lane, V2{da1.a * rc::act_bwd_from_out(a1.a, act_id), da1.b * rc::act_bwd_from_out(a1.b, act_id)});
    }
}

}  // namespace orl_rnnw
