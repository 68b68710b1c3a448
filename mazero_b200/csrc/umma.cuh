// umma.cuh -- thin inline-PTX layer over the Blackwell (sm_100a) tensor-core path used by the fused
// inference kernel: tcgen05.mma with shared-memory operand descriptors and TMEM accumulators, 1-D bulk
// async copies (cp.async.bulk) signalled through mbarriers, tcgen05.ld/st for the epilogues.
//
// Operand layout in shared memory (both A = activations [M rows x K] and B = weights [N rows x K]):
// K-major, no swizzle, built from 8x16-byte "core matrices" (8 rows x 8 bf16):
//      byte_offset(r, k) = (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2
// with LBO = 128 (core matrices adjacent along K are contiguous) and SBO = (K / 8) * 128.
// One tcgen05.mma consumes K = 16 (two core matrices); stepping K by 16 advances the start address by 256 B.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace maz {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- bulk async copy global -> shared (1-D TMA), completion counted in bytes on an mbarrier ------------
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// make generic-proxy shared-memory writes (st.shared) visible to the async proxy (tcgen05.mma / TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM allocation ----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result, uint32_t ncols)  // one full warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)       // the same warp
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (layout_type 0), sm_100 descriptor version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major, M x N tile
__host__ __device__ constexpr uint32_t instr_desc_bf16(uint32_t M, uint32_t N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K=16 step; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM <-> registers: 32 lanes x 32-bit, 16 consecutive columns per call ------------------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- operand tiles ----------------------------------------------------------------------------------------
constexpr uint32_t kLBO = 128;
__host__ __device__ constexpr uint32_t sbo_bytes(uint32_t K) { return (K / 8) * 128; }
__host__ __device__ constexpr uint32_t operand_bytes(uint32_t rows, uint32_t K) { return rows * K * 2; }
// byte offset of the 16-byte chunk holding elements (r, 8*k8 .. 8*k8+7)
__device__ __forceinline__ uint32_t chunk_off(uint32_t r, uint32_t k8, uint32_t K)
{
    return (r >> 3) * sbo_bytes(K) + k8 * kLBO + (r & 7) * 16;
}

// Issue the K/16 MMAs of one GEMM stage: D[128 x N] (+)= A[128 x K] * W[N x K]^T.  ONE thread.
// a_addr / w_addr: shared-memory byte addresses of operand tiles laid out for THEIR full K (a_K, w_K);
// a_k0 / w_k0: first K index used from each (multiples of 16); kk: number of K elements (multiple of 16).
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_addr, uint32_t a_K, uint32_t a_k0, uint32_t w_addr,
                                           uint32_t w_K, uint32_t w_k0, uint32_t kk, uint32_t N, bool accumulate)
{
    const uint32_t idesc = instr_desc_bf16(128, N);
    for (uint32_t k = 0; k < kk; k += 16) {
        const uint64_t da = smem_desc(a_addr + ((a_k0 + k) >> 3) * kLBO, kLBO, sbo_bytes(a_K));
        const uint64_t db = smem_desc(w_addr + ((w_k0 + k) >> 3) * kLBO, kLBO, sbo_bytes(w_K));
        mma_bf16(d_tmem, da, db, idesc, (accumulate || k > 0) ? 1u : 0u);
    }
}

}  // namespace umma
}  // namespace maz
