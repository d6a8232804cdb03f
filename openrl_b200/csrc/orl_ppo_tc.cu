// PPO minibatch forward+backward on the 5th-gen tensor cores (tcgen05, TF32 in / FP32 accumulate in
// TMEM) — the "fast mode" sibling of ppo_fwdbwd_kernel (orl_ppo.cu), producing the same folded
// partial gradients and loss sums.  Selected with ORL_PPO_TF32; obs widths <= 8 (CartPole, GridWorld).
//
// CTA = 512 threads = 128 tile rows x 4 column quarters: warps w, w+4, w+8, w+12 own the same 32 rows
// (the TMEM lane quadrant w%4) and 16 columns each, so every row-wise operation (fc1 with K = d <= 8,
// LayerNorm forward/backward, head, loss) is thread-local apart from a handful of two-float
// exchanges through shared memory (named barriers per row group), and the three 64-wide GEMMs of the row tile go to the tensor
// core, issued by one elected thread:
//     GEMM1  Z3 [128x64]  = n1 [128x64]  . W3f^T            (M=128, N=64, K=64)    fwd fc3
//     GEMM2  dN1[128x64]  = dZ3[128x64]  . W3f              (M=128, N=64, K=64)    bwd-data fc3
//     GEMM3  G  [128x80] += [dZ3^T;dZ1^T][128x128] . [n1^T;X^T;1^T]^T  (M=128, N=80, K=128 rows)
// GEMM3's accumulator stays in TMEM for the whole kernel: rows 0..63 are G3 = dZ3^T n1 (cols 0..63)
// and db3 (col 72, the ones row); rows 64..127 are G1 = dZ1^T X (cols 64..64+d) and db1 (col 72).
// Operands are staged by the row-owning threads in the canonical no-swizzle K-major layout
// (orl_tc.cuh) — transposed operands are written with conflict-free scalar stores thanks to a
// 16-byte pad on the panel stride.  GH = dL^T n3 (n x 64) is a small FFMA reduction.
// TF32 operands are rounded to nearest on store (the MMA truncates); accumulation is fp32.
#include <algorithm>

#include "orl_mlp.cuh"
#include "orl_tc.cuh"

namespace orl {
int ppo_stride_host(int obs_dim, int critic_obs_dim, int n_actions);
}

namespace {
using namespace orl;
using namespace orl::tc;

constexpr int T_M = 128, T_NT = 512;   // four threads per row: thread = (row, column quarter)
constexpr int T_Q = T_NT / T_M;        // threads per row
constexpr int T_CH = 8 / T_Q;          // 8-column chunks per thread
constexpr int NB3 = 80;                       // rows of B3: 64 (n1^T) + 8 (X^T) + 8 (ones / zero)
constexpr uint32_t LBO_A = 128 * 16 + 16;     // panel stride of 128-row tiles (padded)
constexpr uint32_t LBO_B3 = NB3 * 16 + 16;
constexpr uint32_t LBO_W = 64 * 16 + 16;
constexpr int N_LOSS_TC = 8;

__device__ __forceinline__ uint32_t poff(uint32_t lbo, int row, int col) { return (uint32_t)(col >> 2) * lbo + row * 16 + (col & 3) * 4; }
__device__ __forceinline__ void pst(uint8_t* t, uint32_t lbo, int row, int col, float v) { *reinterpret_cast<float*>(t + poff(lbo, row, col)) = v; }
__device__ __forceinline__ float pld(const uint8_t* t, uint32_t lbo, int row, int col) { return *reinterpret_cast<const float*>(t + poff(lbo, row, col)); }
__device__ __forceinline__ uint64_t kdesc(const uint8_t* tile, uint32_t lbo, int k0) {
    return desc_common(smem_u32(tile) + (uint32_t)(k0 >> 2) * lbo, lbo, 128);
}

struct AdvNormTc { float m0, s0, m1, s1; bool two; };

// NOUT: head width known at compile time (1 critic, 2, 5) or 8 = generic (runtime n <= 8)
#define FOR_OUT(j) _Pragma("unroll") for (int j = 0; j < NOUT; ++j) if (NOUT != 8 || j < n)

// row-group barrier: the T_Q warps that share rows [32g, 32g+32) (g = warp % 4)
#define ROWGROUP_SYNC() asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp & 3)), "r"(32 * T_Q) : "memory")
// exchange: the T_Q column slices of a row publish two partial sums and read the others'
#define PAIR_SUM2(v0, v1)                                                          \
    do {                                                                           \
        xch[(half * T_M + row) * 8 + 0] = (v0); xch[(half * T_M + row) * 8 + 1] = (v1); \
        ROWGROUP_SYNC();                                                           \
        float t0_ = 0.f, t1_ = 0.f;                                                \
        _Pragma("unroll") for (int q_ = 0; q_ < T_Q; ++q_) { t0_ += xch[(q_ * T_M + row) * 8 + 0]; t1_ += xch[(q_ * T_M + row) * 8 + 1]; } \
        (v0) = t0_; (v1) = t1_;                                                    \
        ROWGROUP_SYNC();                                                           \
    } while (0)

template <bool POLICY, int NOUT>
__device__ __forceinline__ void tc_net_pass(const OrlPpoArgs& a, uint8_t* smem, int cta, int G, int stride) {
    const int d = POLICY ? a.obs_dim : a.critic_obs_dim;
    const int n = POLICY ? (NOUT == 8 ? a.n_actions : NOUT) : 1;
    const float* params = POLICY ? a.policy_params : a.critic_params;
    const float* obs = POLICY ? a.policy_obs : a.critic_obs;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;   // this thread: tile row, column slice `half`
    constexpr int CW = 8 * T_CH;                  // columns per thread
    const int cb = CW * half;                     // first column of the slice
    const NetOffsets po = net_offsets(d, n);

    // ---- shared memory carve-up (all tensor-core tiles 128-byte aligned) ----
    uint8_t* A12 = smem;                         // n1 -> dZ3 : rows m, cols 64
    uint8_t* A3 = A12 + 16 * LBO_A;              // rows 0..63 dZ3^T, 64..127 dZ1^T ; cols m
    uint8_t* B3 = A3 + 32 * LBO_A;               // rows 0..63 n1^T, 64..71 X^T, 72 ones ; cols m
    uint8_t* B1 = B3 + 32 * LBO_B3;              // W3f   rows j, cols k
    uint8_t* B2 = B1 + 16 * LBO_W;               // W3f^T rows k, cols j
    uint8_t* N3s = B2 + 16 * LBO_W;              // n3 rows (fp32) for the GH reduction: rows m, cols 64
    float* w1t = reinterpret_cast<float*>(N3s + 16 * LBO_A);  // [8][64] k-major, zero padded
    float* b1s = w1t + 8 * H;
    float* b3f = b1s + H;
    float* whf = b3f + H;                        // [8][64] folded head
    float* bhf = whf + MAX_OUT * H;
    float* swh = bhf + MAX_OUT;                  // [8] row sums of whf
    float* DLs = swh + MAX_OUT;                  // [128][8]
    float* xch = DLs + T_M * 8;                  // [T_Q][128][8] row exchange
    float* red = xch + T_Q * T_M * 8;            // [64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 32);  // 3 mbarriers
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 3);

    // ---- stage weights (folded; tensor-core copies rounded to TF32) ----
    for (int i = tid; i < 8 * H; i += T_NT) { const int k = i / H, j = i % H; w1t[i] = (k < d) ? params[po.w1 + j * d + k] : 0.f; }
    for (int i = tid; i < H; i += T_NT) b1s[i] = params[po.b1 + i];
    for (int i = tid; i < H * H; i += T_NT) {
        const int j = i / H, k = i % H;
        const float v = to_tf32(params[po.w3 + i] * params[po.g1 + k]);
        pst(B1, LBO_W, j, k, v);
        pst(B2, LBO_W, k, j, v);
    }
    for (int i = tid; i < MAX_OUT * H; i += T_NT) { const int j = i / H, k = i % H; whf[i] = (j < n) ? params[po.wh + j * H + k] * params[po.g3 + k] : 0.f; }
    for (int j = tid; j < H; j += T_NT) {
        float s = params[po.b3 + j];
        for (int k = 0; k < H; ++k) s = fmaf(params[po.w3 + j * H + k], params[po.be1 + k], s);
        b3f[j] = s;
    }
    for (int j = tid; j < MAX_OUT; j += T_NT) {
        float s = 0.f;
        if (j < n) { s = params[po.bh + j]; for (int k = 0; k < H; ++k) s = fmaf(params[po.wh + j * H + k], params[po.be3 + k], s); }
        bhf[j] = s;
        float rs = 0.f;
        if (j < n) for (int k = 0; k < H; ++k) rs += params[po.wh + j * H + k] * params[po.g3 + k];
        swh[j] = rs;
    }
    // constant rows of B3: ones row 72, zero rows 64+d..71 and 73..79 (this thread's column m = tid)
    if (half == 0) for (int r = 64; r < NB3; ++r) pst(B3, LBO_B3, r, row, r == 72 ? 1.0f : 0.f);
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); }
    if (warp == 0) tmem_alloc(tmem_holder, 256);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_holder;
    const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t idesc64 = make_idesc_tf32(128, 64, false, false);
    const uint32_t idesc80 = make_idesc_tf32(128, NB3, false, false);

    // ---- minibatch constants ----
    const bool pol_masks = a.flags & ORL_PPO_POLICY_ACTIVE_MASKS, val_masks = a.flags & ORL_PPO_VALUE_ACTIVE_MASKS;
    const double rows_d = (double)(a.norm_rows > 0 ? a.norm_rows : a.batch_rows);
    const float inv_rows = (float)(1.0 / rows_d), inv_act = (float)(1.0 / a.mb_stats[2]);
    AdvNormTc an = {0.f, 1.f, 0.f, 1.f, false};
    float vn_mean = 0.f, vn_std = 1.f;
    if (POLICY) {  // ppo.py:402-409
        const double* gs = a.gae_stats;
        const double n_all = gs[ORL_GS_COUNT], n_act = gs[ORL_GS_ACT_COUNT];
        const double mean_all = gs[ORL_GS_ADV_SUM] / n_all;
        const double var_all = fmax(gs[ORL_GS_ADV_SQSUM] / n_all - mean_all * mean_all, 0.0);
        double mean_act = gs[ORL_GS_ADV_ACT_SUM] / n_act;
        double std_act = sqrt(fmax(gs[ORL_GS_ADV_ACT_SQSUM] / n_act - mean_act * mean_act, 0.0));
        if (a.flags & ORL_PPO_ADV_NORMALIZE) {
            const double s0 = (double)((float)sqrt(var_all)) + 1e-5;
            an.two = true; an.m0 = (float)mean_all; an.s0 = (float)s0;
            mean_act = (mean_act - mean_all) / s0; std_act = std_act / s0;
        }
        an.m1 = (float)mean_act; an.s1 = (float)((double)((float)std_act) + 1e-5);
    } else if (a.flags & ORL_PPO_VALUENORM) {
        const float bm = (float)(a.mb_stats[0] / rows_d), bsq = (float)(a.mb_stats[1] / rows_d);
        const float beta = (float)a.vn_beta, omw = (float)(1.0 - a.vn_beta);
        float st[3];
        st[0] = __fadd_rn(__fmul_rn(a.vn_state[0], beta), __fmul_rn(bm, omw));
        st[1] = __fadd_rn(__fmul_rn(a.vn_state[1], beta), __fmul_rn(bsq, omw));
        st[2] = __fadd_rn(__fmul_rn(a.vn_state[2], beta), __fmul_rn(1.0f, omw));
        const VnScalars s = vn_mean_std(st);
        vn_mean = s.mean; vn_std = s.std;
    }

    float gh_acc[1] = {0.f};                  // GH output o = tid  (j = o/64, k = o%64), o < n*64 <= 512
    float dbh_acc = 0.f;                        // threads with k == 0: sum_m dL[m][j]
    float loss0 = 0.f, loss1 = 0.f, loss2 = 0.f;
    uint32_t it = 0;
    // cycle attribution of the tile pipeline (thread 0 only; written to the spare loss slots 3..7 of the
    // partial row and read by tools/tc_phase_profile.py): [0] gather+fc1+LN1 until GEMM1 is issued,
    // [1] wait GEMM1, [2] LN3+head+loss+dZ3 until GEMM2 is issued, [3] wait GEMM2, [4] LN1-bwd, GEMM3 issue, GH
    uint32_t prof[5] = {0u, 0u, 0u, 0u, 0u};
    uint32_t tprev = (uint32_t)clock();
#define PROF_MARK(i) do { if (tid == 0) { const uint32_t now_ = (uint32_t)clock(); prof[i] += now_ - tprev; tprev = now_; } } while (0)

    const long long n_tiles = (a.batch_rows + T_M - 1) / T_M;
    bool nx_valid; long long nx_gi; float nx_x[8], nx_a = 0.f, nx_b = 0.f, nx_c = 0.f, nx_act = 0.f;
    {   // row data of the first tile
        const long long r0 = (long long)cta * T_M + row;
        nx_valid = (cta < n_tiles) && r0 < a.batch_rows;
        nx_gi = nx_valid ? (a.indices ? a.indices[r0] : a.row_begin + r0) : -1;
#pragma unroll
        for (int k = 0; k < 8; ++k) nx_x[k] = (nx_valid && k < d) ? obs[nx_gi * d + k] : 0.f;
        if (nx_valid) {
            if (POLICY) { nx_a = a.actions[nx_gi]; nx_b = a.old_log_probs[nx_gi]; nx_c = a.advantages[nx_gi]; }
            else { nx_a = a.value_preds[nx_gi]; nx_b = a.returns[nx_gi]; }
            nx_act = a.active_masks[nx_gi];
        }
    }
    for (long long tile = cta; tile < n_tiles; tile += G, ++it) {
        const uint32_t par = it & 1u;
        const bool valid = nx_valid;
        const long long gi = nx_gi;
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = nx_x[k];
        float row_a = nx_a, row_b = nx_b, row_c = nx_c, active = nx_act;
        {   // prefetch the NEXT tile's row data (index -> observation / scalars: two dependent global
            // round trips) so that it overlaps this tile's work
            const long long rn = (tile + G) * T_M + row;
            nx_valid = (tile + G < n_tiles) && rn < a.batch_rows;
            nx_gi = nx_valid ? (a.indices ? a.indices[rn] : a.row_begin + rn) : -1;
#pragma unroll
            for (int k = 0; k < 8; ++k) nx_x[k] = (nx_valid && k < d) ? obs[nx_gi * d + k] : 0.f;
            nx_a = nx_b = nx_c = nx_act = 0.f;
            if (nx_valid) {
                if (POLICY) { nx_a = a.actions[nx_gi]; nx_b = a.old_log_probs[nx_gi]; nx_c = a.advantages[nx_gi]; }
                else { nx_a = a.value_preds[nx_gi]; nx_b = a.returns[nx_gi]; }
                nx_act = a.active_masks[nx_gi];
            }
        }

        // Row-wise phases: this thread owns CW = 16 columns [cb, cb+16) of its row and keeps them in
        // registers across the passes (n1 until the LayerNorm-1 backward, n3 until dZ3); the row's other
        // slices live in the 3 sibling threads, reached through the two-float exchanges.
        // ---- fc1 + activation + LayerNorm-1 ----
        float n1[CW];
#pragma unroll
        for (int q4 = 0; q4 < CW; q4 += 4) {
            const float4 b = *reinterpret_cast<const float4*>(b1s + cb + q4);
            n1[q4] = b.x; n1[q4 + 1] = b.y; n1[q4 + 2] = b.z; n1[q4 + 3] = b.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < d) {
#pragma unroll
                for (int q4 = 0; q4 < CW; q4 += 4) {
                    const float4 wv = *reinterpret_cast<const float4*>(w1t + k * H + cb + q4);
                    n1[q4] = fmaf(x[k], wv.x, n1[q4]); n1[q4 + 1] = fmaf(x[k], wv.y, n1[q4 + 1]);
                    n1[q4 + 2] = fmaf(x[k], wv.z, n1[q4 + 2]); n1[q4 + 3] = fmaf(x[k], wv.w, n1[q4 + 3]);
                }
            }
        }
        unsigned posmask = 0u;   // sign bits of this thread's pre-activations
        float s = 0.f, sq = 0.f;
#pragma unroll
        for (int i = 0; i < CW; ++i) {
            if (n1[i] > 0.f) posmask |= 1u << i;
            n1[i] = act_fwd(n1[i], a.activation_id);
            s += n1[i]; sq = fmaf(n1[i], n1[i], sq);
        }
        __syncthreads();    // orders this tile after every thread's GH reads (N3s, DLs) of the previous tile
        PAIR_SUM2(s, sq);
        const float mu1 = s * (1.f / H);
        const float rstd1 = 1.0f / sqrtf(fmaxf(sq * (1.f / H) - mu1 * mu1, 0.f) + LN_EPS);

        // previous tile's GEMM3 must have finished reading A3 / B3
        if (it > 0) mbar_wait(&bars[2], (it - 1) & 1u);
        // n1 row slice (GEMM1 A operand) and n1^T (GEMM3 B operand), TF32-rounded
#pragma unroll
        for (int q4 = 0; q4 < CW; q4 += 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) n1[q4 + i] = to_tf32((n1[q4 + i] - mu1) * rstd1);
            *reinterpret_cast<float4*>(A12 + poff(LBO_A, row, cb + q4)) = make_float4(n1[q4], n1[q4 + 1], n1[q4 + 2], n1[q4 + 3]);
            uint8_t* bt = B3 + poff(LBO_B3, cb + q4, row);
            *reinterpret_cast<float*>(bt) = n1[q4]; *reinterpret_cast<float*>(bt + 16) = n1[q4 + 1];
            *reinterpret_cast<float*>(bt + 32) = n1[q4 + 2]; *reinterpret_cast<float*>(bt + 48) = n1[q4 + 3];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) if (half == 0 && k < d) pst(B3, LBO_B3, 64 + k, row, to_tf32(x[k]));
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        if (warp == 0 && elect_one()) {   // GEMM1: Z3 = n1 . W3f^T
            tcgen05_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) mma_tf32(tmem, kdesc(A12, LBO_A, kk * 8), kdesc(B1, LBO_W, kk * 8), idesc64, kk > 0);
            mma_commit(&bars[0]);
        }

        PROF_MARK(0);
        mbar_wait(&bars[0], par);
        tcgen05_fence_after();
        PROF_MARK(1);

        // ---- Z3 (TMEM) + b3f -> LayerNorm-3 -> n3 (registers; a copy goes to N3s for the GH reduction) ----
        float n3[CW];
        tmem_ld_row16(tmem_row + cb, n3);
        float s3 = 0.f, q3 = 0.f;
#pragma unroll
        for (int i = 0; i < CW; ++i) { n3[i] += b3f[cb + i]; s3 += n3[i]; q3 = fmaf(n3[i], n3[i], q3); }
        PAIR_SUM2(s3, q3);
        const float mu3 = s3 * (1.f / H);
        const float rstd3 = 1.0f / sqrtf(fmaxf(q3 * (1.f / H) - mu3 * mu3, 0.f) + LN_EPS);
        float out[MAX_OUT];
#pragma unroll
        for (int j = 0; j < MAX_OUT; ++j) out[j] = 0.f;
#pragma unroll
        for (int q4 = 0; q4 < CW; q4 += 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) n3[q4 + i] = (n3[q4 + i] - mu3) * rstd3;
            *reinterpret_cast<float4*>(N3s + poff(LBO_A, row, cb + q4)) = make_float4(n3[q4], n3[q4 + 1], n3[q4 + 2], n3[q4 + 3]);
            FOR_OUT(j) {
                const float4 wv = *reinterpret_cast<const float4*>(whf + j * H + cb + q4);
                out[j] = fmaf(n3[q4], wv.x, fmaf(n3[q4 + 1], wv.y, fmaf(n3[q4 + 2], wv.z, fmaf(n3[q4 + 3], wv.w, out[j]))));
            }
        }
        {   // exchange of the partial head dots
            FOR_OUT(j) xch[(half * T_M + row) * 8 + j] = out[j];
            ROWGROUP_SYNC();
            FOR_OUT(j) {
                float t_ = 0.f;
#pragma unroll
                for (int q_ = 0; q_ < T_Q; ++q_) t_ += xch[(q_ * T_M + row) * 8 + j];
                out[j] = t_;
            }
            ROWGROUP_SYNC();
        }
        // dot[j] = sum_k Whf[j][k] n3[k] (needed by the LayerNorm-3 backward); logits add the folded bias
        float dot[MAX_OUT];
#pragma unroll
        for (int j = 0; j < MAX_OUT; ++j) { dot[j] = out[j]; out[j] += (j < NOUT && j < n) ? bhf[j] : 0.f; }

        // ---- head loss + dL/dhead ----
        float dl[MAX_OUT];
#pragma unroll
        for (int j = 0; j < MAX_OUT; ++j) dl[j] = 0.f;
        if (valid) {
            if (POLICY) {
                unsigned masked = 0;
                if (a.action_masks) {
#pragma unroll
                    for (int j = 0; j < MAX_OUT; ++j)
                        if (j < n && a.action_masks[gi * n + j] == 0.f) { out[j] = -6e4f; masked |= 1u << j; }
                }
                float nl[MAX_OUT], pr[MAX_OUT];
                log_softmax_n(out, n, nl, pr);
                const int act = (int)row_a;
                float lp = nl[0];
#pragma unroll
                for (int j = 1; j < MAX_OUT; ++j) if (j == act) lp = nl[j];
                float adv = row_c;
                if (an.two) adv = (adv - an.m0) / an.s0;
                adv = (adv - an.m1) / an.s1;
                const PgTerm pg = pg_term(lp, row_b, adv, a.clip_param, a.flags, a.dual_clip_coeff);
                const float wrow = pol_masks ? active * inv_act : inv_rows;
                float ent = 0.f;
#pragma unroll
                for (int j = 0; j < MAX_OUT; ++j) if (j < n) ent -= pr[j] * nl[j];
                if (half == 0) { loss0 += pg.loss * wrow; loss1 += ent * wrow; loss2 += pg.ratio; }
                const float dlp = pg.dlogp * wrow, went = a.entropy_coef * wrow;
#pragma unroll
                for (int j = 0; j < MAX_OUT; ++j)
                    if (j < n && !((masked >> j) & 1u)) dl[j] = dlp * ((j == act ? 1.f : 0.f) - pr[j]) + went * pr[j] * (nl[j] + ent);
            } else {
                const float v = out[0], vp = row_a, ret = row_b;
                const float target = (a.flags & ORL_PPO_VALUENORM) ? (ret - vn_mean) / vn_std : ret;
                const float diff = v - vp;
                const float clipped = vp + fminf(fmaxf(diff, -a.clip_param), a.clip_param);
                const float e_c = target - clipped, e_o = target - v, dlt = a.huber_delta;
                const bool hub = a.flags & ORL_PPO_HUBER;
                const float l_c = hub ? (fabsf(e_c) <= dlt ? 0.5f * e_c * e_c : dlt * (fabsf(e_c) - 0.5f * dlt)) : 0.5f * e_c * e_c;
                const float l_o = hub ? (fabsf(e_o) <= dlt ? 0.5f * e_o * e_o : dlt * (fabsf(e_o) - 0.5f * dlt)) : 0.5f * e_o * e_o;
                const float gc = hub ? (fabsf(e_c) <= dlt ? e_c : (e_c > 0.f ? dlt : -dlt)) : e_c;
                const float go = hub ? (fabsf(e_o) <= dlt ? e_o : (e_o > 0.f ? dlt : -dlt)) : e_o;
                float l = l_o, dv = -go;
                if (a.flags & ORL_PPO_CLIP_VALUE) {
                    const bool inrange = diff >= -a.clip_param && diff <= a.clip_param;
                    const float dc = inrange ? -gc : 0.f;
                    if (l_o > l_c) { l = l_o; dv = -go; } else if (l_c > l_o) { l = l_c; dv = dc; } else { l = l_o; dv = 0.5f * (-go) + 0.5f * dc; }
                }
                const float wrow = val_masks ? active * inv_act : inv_rows;
                if (half == 0) loss0 += l * wrow;
                dl[0] = a.value_loss_coef * wrow * dv;
            }
        }
        // ---- dn3 = dL . Whf ; LayerNorm-3 backward -> dZ3 (single pass: the two row means are
        //      mean(dn3) = sum_j dL[j] rowsum(Whf[j]) / 64 and mean(dn3 n3) = sum_j dL[j] dot[j] / 64) ----
        float m1 = 0.f, m2 = 0.f;
        FOR_OUT(j) { m1 = fmaf(dl[j], swh[j], m1); m2 = fmaf(dl[j], dot[j], m2); }
        m1 *= (1.f / H); m2 *= (1.f / H);
#pragma unroll
        for (int q4 = 0; q4 < CW; q4 += 4) {
            float g4[4] = {0.f, 0.f, 0.f, 0.f};
            FOR_OUT(j) {
                const float4 wv = *reinterpret_cast<const float4*>(whf + j * H + cb + q4);
                g4[0] = fmaf(dl[j], wv.x, g4[0]); g4[1] = fmaf(dl[j], wv.y, g4[1]); g4[2] = fmaf(dl[j], wv.z, g4[2]); g4[3] = fmaf(dl[j], wv.w, g4[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) g4[i] = to_tf32(rstd3 * (g4[i] - m1 - n3[q4 + i] * m2));
            *reinterpret_cast<float4*>(A12 + poff(LBO_A, row, cb + q4)) = make_float4(g4[0], g4[1], g4[2], g4[3]);   // dZ3 row
            uint8_t* at = A3 + poff(LBO_A, cb + q4, row);                                                              // dZ3^T
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<float*>(at + 16 * i) = g4[i];
        }
        if (half == 0) {
            *reinterpret_cast<float4*>(DLs + row * 8) = make_float4(dl[0], dl[1], dl[2], dl[3]);
            *reinterpret_cast<float4*>(DLs + row * 8 + 4) = make_float4(dl[4], dl[5], dl[6], dl[7]);
        }
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        if (warp == 0 && elect_one()) {   // GEMM2: dN1 = dZ3 . W3f
            tcgen05_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) mma_tf32(tmem, kdesc(A12, LBO_A, kk * 8), kdesc(B2, LBO_W, kk * 8), idesc64, kk > 0);
            mma_commit(&bars[1]);
        }
        PROF_MARK(2);
        mbar_wait(&bars[1], par);
        tcgen05_fence_after();
        PROF_MARK(3);
        // ---- dN1 (TMEM) -> LayerNorm-1 backward -> activation backward -> dZ1^T ----
        {
            float g[CW];
            tmem_ld_row16(tmem_row + cb, g);
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int i = 0; i < CW; ++i) { t1 += g[i]; t2 = fmaf(g[i], n1[i], t2); }
            PAIR_SUM2(t1, t2);
            t1 *= (1.f / H); t2 *= (1.f / H);
            const float std1 = 1.0f / rstd1;
            uint8_t* at = A3 + poff(LBO_A, 64 + cb, row);
#pragma unroll
            for (int i = 0; i < CW; ++i) {
                const float da = rstd1 * (g[i] - t1 - n1[i] * t2);
                const float aval = fmaf(n1[i], std1, mu1);                          // activation output
                *reinterpret_cast<float*>(at + 16 * i) = to_tf32(da * act_bwd(aval, (posmask >> i) & 1u, a.activation_id));
            }
        }
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        if (warp == 0 && elect_one()) {   // GEMM3: G += [dZ3^T; dZ1^T] . [n1^T; X^T; 1^T]^T   (K = 128 tile rows)
            tcgen05_fence_after();
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) mma_tf32(tmem + 64, kdesc(A3, LBO_A, kk * 8), kdesc(B3, LBO_B3, kk * 8), idesc80, (it | kk) > 0);
            mma_commit(&bars[2]);
        }
        // ---- GH += dL^T n3 (FFMA reduction over the tile rows, n3 read back from N3s) ----
        {
            const int o = tid;
            if (o < n * H) {
                const int j = o >> 6, k = o & 63;
                float acc = gh_acc[0], accb = dbh_acc;
                const uint8_t* np = N3s + poff(LBO_A, 0, k);
#pragma unroll 8
                for (int m = 0; m < T_M; ++m) {
                    const float dlv = DLs[m * 8 + j];
                    acc = fmaf(dlv, *reinterpret_cast<const float*>(np + m * 16), acc);
                    if (k == 0) accb += dlv;     // dbh[j] = sum_m dL[m][j], carried by the k == 0 thread of row j
                }
                gh_acc[0] = acc; dbh_acc = accb;
            }
        }
        PROF_MARK(4);
    }

    // ---- flush: G (TMEM) -> partial folded gradients ----
    float* part = a.partials + (size_t)((POLICY ? 0 : G) + cta) * stride;
    const FoldOffsets fo = fold_offsets(d, n);
    if (it > 0) {
        mbar_wait(&bars[2], (it - 1) & 1u);
        tcgen05_fence_after();
        {   // G columns [16*half, 16*half+16) of this row
            uint32_t r16[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r16[0]), "=r"(r16[1]), "=r"(r16[2]), "=r"(r16[3]), "=r"(r16[4]), "=r"(r16[5]), "=r"(r16[6]), "=r"(r16[7]),
                           "=r"(r16[8]), "=r"(r16[9]), "=r"(r16[10]), "=r"(r16[11]), "=r"(r16[12]), "=r"(r16[13]), "=r"(r16[14]), "=r"(r16[15])
                         : "r"(tmem_row + 64 + 16 * half));
            tmem_ld_wait();
            if (row < 64) {
#pragma unroll
                for (int c = 0; c < 16; ++c) part[fo.g3 + row * H + 16 * half + c] = __uint_as_float(r16[c]);
            }
        }
        if (half == 0) {
            uint32_t r16[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r16[0]), "=r"(r16[1]), "=r"(r16[2]), "=r"(r16[3]), "=r"(r16[4]), "=r"(r16[5]), "=r"(r16[6]), "=r"(r16[7]),
                           "=r"(r16[8]), "=r"(r16[9]), "=r"(r16[10]), "=r"(r16[11]), "=r"(r16[12]), "=r"(r16[13]), "=r"(r16[14]), "=r"(r16[15])
                         : "r"(tmem_row + 128));
            tmem_ld_wait();
            if (row < 64) {
                part[fo.db3 + row] = __uint_as_float(r16[8]);
            } else {
                const int j = row - 64;
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c < d) part[fo.g1 + j * d + c] = __uint_as_float(r16[c]);
                part[fo.db1 + j] = __uint_as_float(r16[8]);
            }
        }
    } else {
        for (int i = tid; i < fo.gh; i += T_NT) part[i] = 0.f;   // G1, db1, G3, db3 of an idle CTA
    }
#pragma unroll
    { const int o = tid; if (o < n * H) part[fo.gh + o] = gh_acc[0]; }
    if (tid < n * H && (tid & 63) == 0) { part[fo.dbh + (tid >> 6)] = dbh_acc; part[fo.dls + (tid >> 6)] = 0.f; }
    {
        float v[3] = {loss0, loss1, loss2};
        const int lane = tid & 31;
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float sv = warp_sum(v[k]); if (lane == 0) red[k * 16 + warp] = sv; }
        __syncthreads();
        if (tid < N_LOSS_TC) {
            float sv = 0.f;
            if (tid < 3) for (int wv = 0; wv < T_NT / 32; ++wv) sv += red[tid * 16 + wv];
            part[stride - N_LOSS_TC + tid] = sv;
        }
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) part[stride - N_LOSS_TC + 3 + i] = (float)prof[i];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

__global__ void __launch_bounds__(T_NT, 1) ppo_fwdbwd_tc_kernel(const OrlPpoArgs a, int stride) {
    extern __shared__ __align__(1024) uint8_t smem_tc[];
    const int G = a.grid_per_net;
    if ((int)blockIdx.x < G) {
        if (a.n_actions == 2) tc_net_pass<true, 2>(a, smem_tc, blockIdx.x, G, stride);
        else if (a.n_actions == 5) tc_net_pass<true, 5>(a, smem_tc, blockIdx.x, G, stride);
        else tc_net_pass<true, 8>(a, smem_tc, blockIdx.x, G, stride);
    } else {
        tc_net_pass<false, 1>(a, smem_tc, blockIdx.x - G, G, stride);
    }
}

}  // namespace

namespace orl {
int launch_ppo_fwdbwd_tc(const OrlPpoArgs& a, cudaStream_t st) {
    if (a.obs_dim > 8 || a.critic_obs_dim > 8) {
        set_last_error("orl_ppo_fwdbwd: ORL_PPO_TF32 supports observation widths <= 8 (got %d / %d)", a.obs_dim, a.critic_obs_dim);
        return ORL_ERR_UNSUPPORTED;
    }
    const size_t smem = 16 * LBO_A + 32 * LBO_A + 32 * LBO_B3 + 2 * 16 * LBO_W + 16 * LBO_A +
                        sizeof(float) * (8 * H + H + H + MAX_OUT * H + 2 * MAX_OUT + T_M * 8 + T_Q * T_M * 8 + 64) + 3 * 8 + 16 + 128;
    static bool attr_set = false;
    if (!attr_set) {
        int e = check_cuda(cudaFuncSetAttribute(ppo_fwdbwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                           "cudaFuncSetAttribute(ppo_fwdbwd_tc)");
        if (e) return e;
        attr_set = true;
    }
    const int stride = ppo_stride_host(a.obs_dim, a.critic_obs_dim, a.n_actions);
    ppo_fwdbwd_tc_kernel<<<2 * a.grid_per_net, T_NT, smem, st>>>(a, stride);
    return check_cuda(cudaGetLastError(), "ppo_fwdbwd_tc_kernel");
}
}  // namespace orl
