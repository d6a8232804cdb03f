// Hand-written tcgen05 (5th-gen tensor core) primitives for sm_100a, TF32 operands / FP32
// accumulate in TMEM: shared-memory matrix descriptors (K-major and MN-major), instruction
// descriptor, TMEM allocation, MMA issue, commit -> mbarrier, TMEM -> register loads.
// Bit layouts follow the PTX ISA tcgen05 matrix / instruction descriptors (cross-checked against the
// CUTLASS headers cute/arch/mma_sm100_desc.hpp vendored in this image; no CUTLASS code is used).
//
// Shared-memory operand tile ("panel tile", the canonical no-swizzle / INTERLEAVE layout): a matrix
// of R rows x C float columns (R % 8 == 0, C % 4 == 0) is stored as C/4 column panels; panel p holds
// columns [4p, 4p+4) of every row as R consecutive 16-byte units.  Element (row, col) lives at byte
//        (col/4) * R*16 + row*16 + (col%4)*4.
// An 8-row x 16-byte block (128 contiguous bytes) is one "core matrix".  The SAME buffer can be
// read by the tensor core in both majors, which the update kernel uses to avoid transposed copies:
//   * K-major  (rows = M/N index, columns = K): LBO = R*16 (next core matrix along K),
//     SBO = 128 (next 8-row group); the k-slice of one MMA (K = 8 floats) starts at panel k0/4.
//   * MN-major (rows = K index, columns = M/N): SBO = R*16 (next 4 columns along M/N),
//     LBO = 128 (next 8 rows along K); the k-slice of one MMA starts at row k0.
// A thread that owns a row stores float4s; the 32 rows of a warp are 512 contiguous bytes per
// panel (conflict-free).  (TF32 MN-major operands admit no 128B swizzle other than the
// 32-byte-atomic one, so a swizzled buffer could not be shared between the two majors.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace orl {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t panel_offset(int rows, int row, int col) {
    return (uint32_t)((col >> 2) * rows * 16 + row * 16 + (col & 3) * 4);
}
// round-to-nearest TF32 (the tensor core truncates the low 13 mantissa bits of fp32 operands)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void panel_store(float* tile, int rows, int row, int col, float v) {
    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tile) + panel_offset(rows, row, col)) = v;
}
// 16-byte store of columns [col, col+4), col % 4 == 0
__device__ __forceinline__ void panel_store4(float* tile, int rows, int row, int col, float4 v) {
    *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(tile) + panel_offset(rows, row, col)) = v;
}
__device__ __forceinline__ float4 panel_load4(const float* tile, int rows, int row, int col) {
    return *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(tile) + panel_offset(rows, row, col));
}

__device__ __forceinline__ uint64_t desc_common(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);             // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    return d;                                          // layout type bits [61,64) = 0: SWIZZLE_NONE
}
// K-major operand: MMA k-slice starting at float column k0 (k0 % 8 == 0); first M/N row r0 (r0 % 8 == 0)
__device__ __forceinline__ uint64_t make_desc_kmajor(const float* tile, int rows, int k0, int r0 = 0) {
    return desc_common(smem_u32(tile) + (uint32_t)((k0 >> 2) * rows * 16 + r0 * 16), (uint32_t)rows * 16, 128);
}
// MN-major operand: MMA k-slice = rows [k0, k0+8) (k0 % 8 == 0); first M/N index = column c0 (c0 % 4 == 0)
__device__ __forceinline__ uint64_t make_desc_mnmajor(const float* tile, int rows, int k0, int c0 = 0) {
    return desc_common(smem_u32(tile) + (uint32_t)(k0 * 16 + (c0 >> 2) * rows * 16), 128, (uint32_t)rows * 16);
}
// instruction descriptor, kind::tf32, FP32 accumulate
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                        // c_format = F32
    d |= 2u << 7;                        // a_format = TF32
    d |= 2u << 10;                       // b_format = TF32
    d |= (a_mn_major ? 1u : 0u) << 15;   // a_major
    d |= (b_mn_major ? 1u : 0u) << 16;   // b_major
    d |= (uint32_t)(N >> 3) << 17;       // n_dim
    d |= (uint32_t)(M >> 4) << 24;       // m_dim
    return d;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* holder, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(a), "r"(parity) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
// make the mbarrier track completion of all MMAs issued so far by this thread
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this thread's lane (row), 32 consecutive fp32 columns starting at taddr's column
__device__ __forceinline__ void tmem_ld_row32_nowait(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 consecutive fp32 columns of this thread's lane (row)
__device__ __forceinline__ void tmem_ld_row8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 consecutive fp32 columns of this thread's lane (row)
__device__ __forceinline__ void tmem_ld_row16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_row32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_row64(uint32_t taddr, float (&v)[64]) {
    float a[32], b[32];
    tmem_ld_row32(taddr, a);
    tmem_ld_row32(taddr + 32, b);
#pragma unroll
    for (int i = 0; i < 32; ++i) { v[i] = a[i]; v[32 + i] = b[i]; }
}

}  // namespace tc
}  // namespace orl
