// Standalone validation of the hand-written tcgen05 (TF32) building blocks used by the PPO update
// kernel: smem SW128 descriptors (K-major and MN-major), instruction descriptor, TMEM alloc/ld,
// commit -> mbarrier.   nvcc -gencode arch=compute_100a,code=sm_100a tools/tc_test.cu -o tc_test
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../openrl_b200/csrc/orl_tc.cuh"

using namespace orl::tc;

// test 1: D[128x64] = A[128x64] . B[64x64]^T   (A, B K-major)
// test 2: G[128x96] = P^T . Q  with P[128(m) x 128(cols)] and Q[128(m) x 96(cols)] both read MN-major
//         (K = m), accumulated over 2 "tiles".
__global__ void __launch_bounds__(128) tc_test_kernel(const float* A, const float* B, float* D, const float* P,
                                                      const float* Q, float* G) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* sA = (float*)smem;                 // 2 col-blocks x 16 KB
    float* sB = (float*)(smem + 32768);       // 2 col-blocks x 8 KB
    float* sP = (float*)(smem + 49152);       // 4 col-blocks x 16 KB
    float* sQ = (float*)(smem + 49152 + 65536);  // 3 col-blocks x 16 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_holder;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc(&tmem_holder, 256);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_holder;

    // ---- test 1 ----
    for (int k = 0; k < 64; ++k) panel_store(sA, 128, tid, k, to_tf32(A[tid * 64 + k]));
    if (tid < 64) for (int k = 0; k < 64; ++k) panel_store(sB, 64, tid, k, to_tf32(B[tid * 64 + k]));
    fence_proxy_async();
    __syncthreads();
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = make_idesc_tf32(128, 64, false, false);
        for (int kk = 0; kk < 8; ++kk) {   // K = 64 = 8 MMAs of K=8
            const uint64_t da = make_desc_kmajor(sA, 128, kk * 8);
            const uint64_t db = make_desc_kmajor(sB, 64, kk * 8);
            mma_tf32(tmem + 0, da, db, idesc, kk > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tcgen05_fence_after();
    {
        float v[64];
        tmem_ld_row64(tmem + ((uint32_t)(warp * 32) << 16) + 0, v);
        for (int c = 0; c < 64; ++c) D[(warp * 32 + lane) * 64 + c] = v[c];
    }
    tcgen05_fence_before();
    __syncthreads();

    // ---- test 2: two tiles accumulated ----
    for (int tile = 0; tile < 2; ++tile) {
        const float* Pt = P + (size_t)tile * 128 * 128;
        const float* Qt = Q + (size_t)tile * 128 * 96;
        for (int c = 0; c < 128; ++c) panel_store(sP, 128, tid, c, to_tf32(Pt[tid * 128 + c]));
        for (int c = 0; c < 96; ++c) panel_store(sQ, 128, tid, c, to_tf32(Qt[tid * 96 + c]));
        fence_proxy_async();
        __syncthreads();
        tcgen05_fence_after();
        if (warp == 0 && elect_one()) {
            const uint32_t idesc = make_idesc_tf32(128, 96, true, true);
            for (int kk = 0; kk < 16; ++kk) {  // K = m = 128 = 16 MMAs
                const uint64_t da = make_desc_mnmajor(sP, 128, kk * 8);
                const uint64_t db = make_desc_mnmajor(sQ, 128, kk * 8);
                mma_tf32(tmem + 64, da, db, idesc, (tile | kk) > 0);
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, (tile + 1) & 1);
        tcgen05_fence_after();
        __syncthreads();
    }
    {
        float v[64];
        tmem_ld_row64(tmem + ((uint32_t)(warp * 32) << 16) + 64, v);
        for (int c = 0; c < 64; ++c) G[(warp * 32 + lane) * 96 + c] = v[c];
        float w[32];
        tmem_ld_row32(tmem + ((uint32_t)(warp * 32) << 16) + 128, w);
        for (int c = 0; c < 32; ++c) G[(warp * 32 + lane) * 96 + 64 + c] = w[c];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static float tf32r(float x) {  // round-to-nearest-even to 10 mantissa bits (what the tensor core sees, approx.)
    uint32_t u; memcpy(&u, &x, 4);
    u += 0xFFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u;
    float y; memcpy(&y, &u, 4); return y;
}

int main() {
    std::vector<float> A(128 * 64), B(64 * 64), P(2 * 128 * 128), Q(2 * 128 * 96);
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto& x : A) x = rnd(); for (auto& x : B) x = rnd(); for (auto& x : P) x = rnd(); for (auto& x : Q) x = rnd();
    float *dA, *dB, *dD, *dP, *dQ, *dG;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * 64 * 4);
    cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dQ, Q.size() * 4); cudaMalloc(&dG, 128 * 96 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 49152 + 65536 + 49152 + 1024;
    cudaFuncSetAttribute(tc_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_test_kernel<<<1, 128, smem>>>(dA, dB, dD, dP, dQ, dG);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> D(128 * 64), G(128 * 96);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(G.data(), dG, G.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e1t = 0, e2 = 0, e2t = 0, m1 = 0, m2 = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        double s = 0, st = 0;
        for (int k = 0; k < 64; ++k) { s += (double)A[m * 64 + k] * B[n * 64 + k]; st += (double)tf32r(A[m * 64 + k]) * tf32r(B[n * 64 + k]); }
        e1 = fmax(e1, fabs(D[m * 64 + n] - s)); e1t = fmax(e1t, fabs(D[m * 64 + n] - st)); m1 = fmax(m1, fabs(s));
    }
    for (int i = 0; i < 128; ++i) for (int j = 0; j < 96; ++j) {
        double s = 0, st = 0;
        for (int t = 0; t < 2; ++t) for (int m = 0; m < 128; ++m) {
            const float p = P[(t * 128 + m) * 128 + i], q = Q[(t * 128 + m) * 96 + j];
            s += (double)p * q; st += (double)tf32r(p) * tf32r(q);
        }
        e2 = fmax(e2, fabs(G[i * 96 + j] - s)); e2t = fmax(e2t, fabs(G[i * 96 + j] - st)); m2 = fmax(m2, fabs(s));
    }
    printf("test1 K-major:  max|D|=%.3f  err vs fp64=%.3e  err vs tf32-rounded inputs=%.3e\n", m1, e1, e1t);
    printf("test2 MN-major: max|G|=%.3f  err vs fp64=%.3e  err vs tf32-rounded inputs=%.3e\n", m2, e2, e2t);
    printf("%s\n", (e1t < 1e-3 && e2t < 1e-3) ? "TC_TEST PASS" : "TC_TEST FAIL");
    return 0;
}
