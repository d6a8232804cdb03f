// Standalone validation of the fp16 hi/lo split ("2xFP16", fp32-class accuracy) tcgen05 building blocks
// of the round-2 PPO update kernel.  Every operand lives in ONE row-major panel buffer
//     element (row r, feature f) at (f/8)*PANEL + r*16 + (f%8)*2      (PANEL = rows*16 bytes)
// and is read K-major (rows = M/N index) or MN-major (rows = K index) by descriptor only — no
// transposed copies.  x = hi + lo with hi = fp16(x), lo = fp16(x - hi); D = Al*Bh + Ah*Bl + Ah*Bh.
//   test A: D[128x64]  = P[128x64] . W[64x64]^T          (A K-major, B K-major)
//   test B: D[128x64]  = P[128x64] . W[64x64]            (A K-major, B MN-major: same W buffer)
//   test C: G[128x80] += P2[m][128 feats]^T . Q[m][80 feats]   (both MN-major, K = 128 tile rows, 2 tiles)
// nvcc -gencode arch=compute_100a,code=sm_100a tools/tc_test2.cu -o tools/tc_test2
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../openrl_b200/csrc/orl_tc.cuh"

using namespace orl::tc;

__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N, bool a_mn, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                  // c_format = F32 ; a_format = b_format = 0 (F16)
    d |= (a_mn ? 1u : 0u) << 15;
    d |= (b_mn ? 1u : 0u) << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool acc) {
    const uint32_t a = acc ? 1u : 0u;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(a)
                 : "memory");
}
// K-major: k-slice of 16 features starting at feature f0 (panels f0/8, f0/8+1): LBO = panel stride, SBO = 128
__device__ __forceinline__ uint64_t dk(const uint8_t* buf, uint32_t panel, int f0) { return desc_common(smem_u32(buf) + (uint32_t)(f0 >> 3) * panel, panel, 128); }
// MN-major: k-slice = 16 rows starting at row r0; MN blocks of 8 features at SBO = panel stride; 8-row groups at LBO = 128
__device__ __forceinline__ uint64_t dmn(const uint8_t* buf, uint32_t panel, int r0) { return desc_common(smem_u32(buf) + (uint32_t)r0 * 16, 128, panel); }

__device__ __forceinline__ void split_store(uint8_t* hi, uint8_t* lo, uint32_t panel, int row, int f, float x) {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    const uint32_t off = (uint32_t)(f >> 3) * panel + row * 16 + (f & 7) * 2;
    *reinterpret_cast<__half*>(hi + off) = h;
    *reinterpret_cast<__half*>(lo + off) = l;
}

__global__ void __launch_bounds__(128) tc2_kernel(const float* P, const float* W, float* DA, float* DB, const float* P2,
                                                  const float* Q, float* G) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t PA = 128 * 16, PW = 64 * 16;
    uint8_t* Ph = smem;              // 8 panels
    uint8_t* Pl = Ph + 8 * PA;
    uint8_t* Wh = Pl + 8 * PA;
    uint8_t* Wl = Wh + 8 * PW;
    uint8_t* P2h = Wl + 8 * PW;      // 16 panels
    uint8_t* P2l = P2h + 16 * PA;
    uint8_t* Qh = P2l + 16 * PA;     // 10 panels
    uint8_t* Ql = Qh + 10 * PA;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_holder;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc(&tmem_holder, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_holder;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;

    for (int f = 0; f < 64; ++f) split_store(Ph, Pl, PA, tid, f, P[tid * 64 + f]);
    if (tid < 64) for (int f = 0; f < 64; ++f) split_store(Wh, Wl, PW, tid, f, W[tid * 64 + f]);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0 && elect_one()) {
        const uint32_t id = make_idesc_f16(128, 64, false, false);
        int n = 0;
        for (int pass = 0; pass < 3; ++pass) {
            const uint8_t* a = pass == 0 ? Pl : Ph;
            const uint8_t* b = pass == 1 ? Wl : Wh;
            for (int k = 0; k < 64; k += 16) mma_f16(tmem, dk(a, PA, k), dk(b, PW, k), id, n++ > 0);
        }
        // test B: D = P . W : N = k (feature of W rows?)  D[m][k] = sum_j P[m][j] W[j][k]; B is W read MN-major (K index = row j)
        const uint32_t idb = make_idesc_f16(128, 64, false, true);
        n = 0;
        for (int pass = 0; pass < 3; ++pass) {
            const uint8_t* a = pass == 0 ? Pl : Ph;
            const uint8_t* b = pass == 1 ? Wl : Wh;
            for (int k = 0; k < 64; k += 16) mma_f16(tmem + 64, dk(a, PA, k), dmn(b, PW, k), idb, n++ > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, phase); phase ^= 1;
    tcgen05_fence_after();
    {
        float v[64];
        tmem_ld_row64(trow, v);
        for (int c = 0; c < 64; ++c) DA[tid * 64 + c] = v[c];
        tmem_ld_row64(trow + 64, v);
        for (int c = 0; c < 64; ++c) DB[tid * 64 + c] = v[c];
    }
    tcgen05_fence_before();
    __syncthreads();

    for (int tile = 0; tile < 2; ++tile) {
        for (int f = 0; f < 128; ++f) split_store(P2h, P2l, PA, tid, f, P2[((size_t)tile * 128 + tid) * 128 + f]);
        for (int f = 0; f < 80; ++f) split_store(Qh, Ql, PA, tid, f, Q[((size_t)tile * 128 + tid) * 80 + f]);
        fence_proxy_async();
        __syncthreads();
        tcgen05_fence_after();
        if (warp == 0 && elect_one()) {
            const uint32_t id = make_idesc_f16(128, 80, true, true);
            int n = tile;
            for (int pass = 0; pass < 3; ++pass) {
                const uint8_t* a = pass == 0 ? P2l : P2h;
                const uint8_t* b = pass == 1 ? Ql : Qh;
                for (int r = 0; r < 128; r += 16) mma_f16(tmem + 128, dmn(a, PA, r), dmn(b, PA, r), id, n++ > 0);
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tcgen05_fence_after();
        __syncthreads();
    }
    {
        float v[64];
        tmem_ld_row64(trow + 128, v);
        for (int c = 0; c < 64; ++c) G[tid * 80 + c] = v[c];
        float w[16];
        tmem_ld_row16(trow + 192, w);
        for (int c = 0; c < 16; ++c) G[tid * 80 + 64 + c] = w[c];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    std::vector<float> P(128 * 64), W(64 * 64), P2(2 * 128 * 128), Q(2 * 128 * 80);
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto& x : P) x = 3.f * rnd();
    for (auto& x : W) x = 0.3f * rnd();
    for (auto& x : P2) x = 5.f * rnd() * ((rand() & 7) == 0 ? 1e-3f : 1.f);
    for (auto& x : Q) x = 3.f * rnd();
    for (int t = 0; t < 2; ++t) for (int m = 0; m < 128; ++m) Q[((size_t)t * 128 + m) * 80 + 64] = 1.0f;   // ones column
    float *dP, *dW, *dDA, *dDB, *dP2, *dQ, *dG;
    cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dDA, 128 * 64 * 4); cudaMalloc(&dDB, 128 * 64 * 4);
    cudaMalloc(&dP2, P2.size() * 4); cudaMalloc(&dQ, Q.size() * 4); cudaMalloc(&dG, 128 * 80 * 4);
    cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dP2, P2.data(), P2.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 2 * 8 * 2048 + 2 * 8 * 1024 + 2 * 16 * 2048 + 2 * 10 * 2048 + 1024;
    cudaFuncSetAttribute(tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc2_kernel<<<1, 128, smem>>>(dP, dW, dDA, dDB, dP2, dQ, dG);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> DA(128 * 64), DB(128 * 64), G(128 * 80);
    cudaMemcpy(DA.data(), dDA, DA.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(DB.data(), dDB, DB.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(G.data(), dG, G.size() * 4, cudaMemcpyDeviceToHost);
    double ea = 0, eb = 0, ec = 0, fa = 0, fb = 0, fc = 0, ma = 0, mb = 0, mc = 0;   // e*: device err, f*: fp32 FFMA chain err, m*: scale
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        double s = 0, sb = 0; float f = 0.f, fb32 = 0.f;
        for (int k = 0; k < 64; ++k) {
            s += (double)P[m * 64 + k] * W[n * 64 + k]; f = fmaf(P[m * 64 + k], W[n * 64 + k], f);
            sb += (double)P[m * 64 + k] * W[k * 64 + n]; fb32 = fmaf(P[m * 64 + k], W[k * 64 + n], fb32);
        }
        ea = fmax(ea, fabs(DA[m * 64 + n] - s)); fa = fmax(fa, fabs(f - s)); ma = fmax(ma, fabs(s));
        eb = fmax(eb, fabs(DB[m * 64 + n] - sb)); fb = fmax(fb, fabs(fb32 - sb)); mb = fmax(mb, fabs(sb));
    }
    for (int i = 0; i < 128; ++i) for (int j = 0; j < 80; ++j) {
        double s = 0; float f = 0.f;
        for (int t = 0; t < 2; ++t) for (int m = 0; m < 128; ++m) {
            const float p = P2[((size_t)t * 128 + m) * 128 + i], q = Q[((size_t)t * 128 + m) * 80 + j];
            s += (double)p * q; f = fmaf(p, q, f);
        }
        ec = fmax(ec, fabs(G[i * 80 + j] - s)); fc = fmax(fc, fabs(f - s)); mc = fmax(mc, fabs(s));
    }
    printf("test A K/K   : max|D|=%.3f  err=%.3e  (fp32 FFMA chain err=%.3e)\n", ma, ea, fa);
    printf("test B K/MN  : max|D|=%.3f  err=%.3e  (fp32 FFMA chain err=%.3e)\n", mb, eb, fb);
    printf("test C MN/MN : max|G|=%.3f  err=%.3e  (fp32 FFMA chain err=%.3e)\n", mc, ec, fc);
    const bool ok = ea < 2e-5 * ma && eb < 2e-5 * mb && ec < 2e-5 * mc;
    printf("%s\n", ok ? "TC_TEST2 PASS" : "TC_TEST2 FAIL");
    return 0;
}
