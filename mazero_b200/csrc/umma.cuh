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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// non-blocking test (try_wait may suspend the thread for a system-dependent time when the phase is not complete)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
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

// ---- TMEM <-> registers: 32 lanes x 32-bit, 16 / 32 consecutive columns per call -------------------------
// The loaded registers are passed THROUGH the wait statement ("+r") so that neither nvcc nor ptxas can
// schedule a consumer between the asynchronous load and tcgen05.wait::ld.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// named barrier over `nthreads` threads (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts2f(uint32_t saddr, float a, float b)
{
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(saddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float2 lds2f(uint32_t saddr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts1f(uint32_t saddr, float a)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(a) : "memory");
}
__device__ __forceinline__ float lds1v(uint32_t saddr)   // volatile-ordered scalar load (exchange buffers)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}
// shared-space vector load / store with 32-bit shared addresses
__device__ __forceinline__ float4 lds4(uint32_t saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void lds4u(uint32_t saddr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d)
{
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(saddr) : "memory");
}
__device__ __forceinline__ float lds1(uint32_t saddr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t saddr, uint4 v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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

// ---- 128-byte-swizzled operand tiles (K a multiple of 64) ----------------------------------------------------
// The no-swizzle layout above makes the tensor core fetch operands with shared-memory bank conflicts (measured:
// ~1150 cycles per 128x128x128 GEMM instead of ~520).  SWIZZLE_128B, K-major: an "atom" is 8 rows x 64 bf16
// (8 x 128 B = 1 KB, tile base 1 KB aligned); inside an atom the 16-byte chunk j of row r sits at chunk (j ^ r);
// atoms are stored [K/64][rows/8]:
//      byte_offset(r, k) = (k / 64) * rows * 128 + (r / 8) * 1024 + (r % 8) * 128 + (((k % 64) / 8) ^ (r % 8)) * 16 + (k % 8) * 2
// Descriptor: layout_type 2 (SWIZZLE_128B), SBO = 1024, LBO unused (1).  A K = 16 step inside an atom advances the
// start address by 32 B (the hardware applies the XOR to the generated addresses); the next 64-wide K block is
// `rows * 128` bytes further.
__device__ __forceinline__ uint32_t chunk_off_sw(uint32_t r, uint32_t k8, uint32_t rows)
{
    return (k8 >> 3) * rows * 128u + (r >> 3) * 1024u + (r & 7u) * 128u + (((k8 & 7u) ^ (r & 7u)) << 4);
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;      // SBO: 8-row groups are 1 KB apart
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// Measured on B200 (3m, 26 tiles): the swizzled layout is functionally correct (tests/test_infer_gpu.py) but did NOT
// speed up the GEMM stages (105 us vs 93 us per launch: the per-chunk time is not operand-fetch bound and the
// epilogue's address arithmetic grows), so it is compiled out.  Set MAZ_SWIZZLE to 1 to re-enable (and
// mazero_b200.fused.SWIZZLE = True on the host side).
#ifndef MAZ_SWIZZLE
#define MAZ_SWIZZLE 0
#endif
__host__ __device__ constexpr bool use_swizzle(uint32_t K) { return MAZ_SWIZZLE && (K % 64u) == 0; }
// operand-layout-aware chunk address: swizzled for K % 64 == 0, core-matrix layout otherwise
__device__ __forceinline__ uint32_t operand_chunk_off(uint32_t r, uint32_t k8, uint32_t K, uint32_t rows)
{
    return use_swizzle(K) ? chunk_off_sw(r, k8, rows) : ((r >> 3) * ((K / 8) * 128u) + k8 * 128u + (r & 7u) * 16u);
}

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
// a_rows / w_rows: number of rows of the A tile (always 128) and of the W chunk (needed by the swizzled layout)
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_addr, uint32_t a_K, uint32_t a_k0, uint32_t w_addr,
                                           uint32_t w_K, uint32_t w_k0, uint32_t kk, uint32_t N, bool accumulate,
                                           uint32_t w_rows = 0)
{
    const uint32_t idesc = instr_desc_bf16(128, N);
    uint32_t acc = accumulate ? 1u : 0u;
    if (w_rows == 0) w_rows = N;
    const bool sa = use_swizzle(a_K), sw = use_swizzle(w_K);
    if (!sa && !sw) {
        // descriptors of consecutive K=16 steps differ only in the start-address field: +256 B = +16 (address >> 4)
        uint64_t da = smem_desc(a_addr + (a_k0 >> 3) * kLBO, kLBO, sbo_bytes(a_K));
        uint64_t db = smem_desc(w_addr + (w_k0 >> 3) * kLBO, kLBO, sbo_bytes(w_K));
#pragma unroll 1
        for (uint32_t k = 0; k < kk; k += 16) {
            mma_bf16(d_tmem, da, db, idesc, acc);
            da += 16;
            db += 16;
            acc = 1u;
        }
        return;
    }
#pragma unroll 1
    for (uint32_t k = 0; k < kk; k += 16) {
        const uint32_t ka = a_k0 + k, kw = w_k0 + k;
        const uint64_t da = sa ? smem_desc_sw128(a_addr + (ka >> 6) * 128u * 128u + ((ka & 63u) >> 3) * 16u)
                               : smem_desc(a_addr + (ka >> 3) * kLBO, kLBO, sbo_bytes(a_K));
        const uint64_t db = sw ? smem_desc_sw128(w_addr + (kw >> 6) * w_rows * 128u + ((kw & 63u) >> 3) * 16u)
                               : smem_desc(w_addr + (kw >> 3) * kLBO, kLBO, sbo_bytes(w_K));
        mma_bf16(d_tmem, da, db, idesc, acc);
        acc = 1u;
    }
}
// polling wait with back-off, for the single-thread producer / MMA warps (keeps them out of the epilogue
// warps' issue slots and shared-memory pipe while they wait)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}

}  // namespace umma
}  // namespace maz
