// Raw sm_100a PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc/mma/commit/ld), fences.
// No CUTLASS/CuTe — everything here is inline PTX assembled by nvcc 12.9 for compute_100a.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

#ifndef B200_SPIN_LIMIT
#define B200_SPIN_LIMIT (1u << 28)  // bounded mbarrier spins: a protocol bug traps instead of hanging the box
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the warp is parked by the hardware until the phase completes (or the hint, in ns, expires)
// instead of re-issuing the test — a plain spin loop spent ~12 % of the attention kernel's issue slots on TRYWAIT + BRA (ncu r1h).
#ifndef B200_WAIT_HINT_NS
#define B200_WAIT_HINT_NS 0x989680u
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#ifdef B200_WAIT_NO_HINT   // developer A/B: the plain form (hardware default suspend time)
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(B200_WAIT_HINT_NS)
        : "memory");
#endif
    return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a protocol bug traps after B200_WAIT_TIMEOUT_NS instead of hanging the box (each failed try_wait may already
// have been parked for up to the suspend hint, so the bound is on wall time, not on the iteration count).
#ifndef B200_WAIT_TIMEOUT_NS
#define B200_WAIT_TIMEOUT_NS 4000000000ull
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (global_ns() - t0 > B200_WAIT_TIMEOUT_NS) __trap();
    }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2-D tiled load, coordinates (c0 = innermost element index, c1 = row index); completes `bytes` on `bar`.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// TMA tile STORE (shared -> global, bulk async group): rows / columns of the box that fall outside the tensor are clipped by the hardware.
// The issuing thread must have ordered the generic-proxy smem writes of the whole warp before it (fence.proxy.async by every writer,
// then a warp barrier); the staging buffer may be rewritten once cp.async.bulk.wait_group.read has returned.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
// Programmatic dependent launch (PDL). Every kernel of the library executes pdl_wait() before its first global-memory access: when
// the kernel was launched with the programmatic-stream-serialisation attribute (B200_PDL=1, common.cuh) its CTAs may become resident
// and run their prologue while the previous kernel in the stream is still draining, and pdl_wait() blocks until that kernel has
// completed and flushed; launched normally, the instruction returns immediately. pdl_launch_dependents() lets the NEXT kernel start
// that early once every CTA of this grid has issued it (or exited) — placed after the main loop of the long persistent kernels.
// Compiled in only with `make PDL=1` (-DB200_PDL_BUILD=1): the default build is instruction-for-instruction the one validated on
// hardware in round 1; the PDL build is the first experiment of round 2 (enable at run time with B200_PDL=1).
#ifndef B200_PDL_BUILD
#define B200_PDL_BUILD 0
#endif
#if B200_PDL_BUILD
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_launch_dependents() {}
#endif

// 1-D bulk copy global -> shared (TMA engine, no tensor map): `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = lane i of the warp's quadrant)
// ---- cta_group::2 (CTA pair) forms. The pair shares one MMA of M = 256: each CTA supplies 128 rows of A and half of B's N rows,
// each CTA's TMEM receives the accumulators of its own 128 rows. All pair-level mbarriers live in the even ("leader") CTA.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    // non-.aligned forms: the single-lane producer / MMA roles leave their warps diverged when they get here
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// TMA load issued by either CTA of the pair into ITS OWN smem; the transaction bytes are credited to the leader CTA's mbarrier
// (bit 24 of a shared::cluster address selects the odd CTA of a pair — clearing it addresses the same offset in the even CTA).
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, both CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit of the pair's MMAs: one arrival on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 "version 1"): start addr>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), base_offset [49,52)=0, layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for kind::f16: c_format F32 (1) [4,6), a/b format BF16 (1) [7,10)/[10,13),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Coalesced fp32 accumulation of a row-per-lane 32 x 32 block: lane l holds v[0..31] = 32 consecutive columns of row (row0 + l).
// The block is transposed through a padded smem tile so that every RED instruction adds 32 CONSECUTIVE floats of one row
// (one 128-byte L2 transaction instead of 32 scattered 4-byte ones).  stg: 32 x 33 floats, private to the warp.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// stg: 32 x 32 fp32 staging tile (4 KB), 16-byte groups XOR-swizzled by the row so that both the row-per-lane writes and the
// 8-lanes-per-row reads are bank-conflict free without padding (the 4 KB tile doubles as the TMA-store staging buffer of the bf16 path)
__device__ __forceinline__ void warp_red_rows_f32(float* stg, const float (&v)[32], float* base, long long ld, int row0, int nrows_total,
                                                  int ncols_valid, int lane) {
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 8; ++g)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((g ^ (lane & 7)) << 2)) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    __syncwarp();
    if (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) {
        // 16-byte vector reductions (REDG.F32x4): 8 lanes cover the 128 bytes of a row, 4 rows per instruction
        const int g = lane & 7, cg = g * 4, rsub = lane >> 3;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            if (row0 + r < nrows_total) {
                const float4 s4 = *reinterpret_cast<const float4*>(stg + r * 32 + ((g ^ (r & 7)) << 2));
                float* dst = base + (long long)(row0 + r) * ld + cg;
                if (cg + 4 <= ncols_valid) red_add_v4(dst, s4.x, s4.y, s4.z, s4.w);
                else {
                    const float sp[4] = {s4.x, s4.y, s4.z, s4.w};
                    for (int j = 0; j < 4; ++j)
                        if (cg + j < ncols_valid) atomicAdd(dst + j, sp[j]);
                }
            }
        }
    } else if (lane < ncols_valid) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r)
            if (row0 + r < nrows_total) atomicAdd(base + (long long)(row0 + r) * ld + lane, stg[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))]);
    }
    __syncwarp();
}

// ---------------------------------------------------------------- misc
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// counter-based RNG for dropout masks (recomputed in backward from the same (seed, index))
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, float p) {
    return (hash_u32(seed, idx) >> 8) * (1.0f / 16777216.0f) >= p;
}

// Dropout masks: counter-based, recomputed by the backward kernels from the same (seed, index). One mixing step decides TWO
// neighbouring elements: x = pair * golden + seed; x ^= x >> 15; then the HIGH bits of two different odd multiples of x are compared
// against the threshold as full 32-bit words — keep iff word >= thresh16 << 16, i.e. P(drop) = thresh16 / 65536. idx = row_id *
// stride + column with an even stride, so (idx >> 1) pairs columns (2k, 2k+1) of one row; only the low 32 bits of the pair index
// are hashed (the pattern repeats after 2^33 elements).
// Cost per pair: 3 IMAD (FMA pipe) + SHF + LOP3 + 2 ISETP (+ the two selects): the round-1 hash (two xorshift-multiply rounds, then
// 16-bit field extraction) put 10 instructions per pair on the half-rate ALU pipe, which the softmax threads of the attention kernels
// are bound by (ncu r2c: ALU pipe 42 %, dropout = 160 of the 509 instructions per warp and key tile).
struct DropWords { uint32_t a, b; };
__device__ __forceinline__ DropWords drop_words(uint32_t seedmix, uint32_t pair) {
    uint32_t x = pair * 0x9E3779B1u + seedmix;
    x ^= x >> 15;
    DropWords w;
    w.a = x * 0x85EBCA6Bu;
    w.b = x * 0xC2B2AE35u;
    return w;
}
__device__ __forceinline__ uint32_t drop_thresh32(uint32_t thresh16) { return thresh16 << 16; }
__device__ __forceinline__ uint32_t seed_mix32(uint64_t seed) { return (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA77u); }
__device__ __forceinline__ bool dropout_keep16(uint64_t seed, uint64_t idx, uint32_t thresh) {
    const DropWords w = drop_words(seed_mix32(seed), (uint32_t)(idx >> 1));
    return ((idx & 1) ? w.b : w.a) >= drop_thresh32(thresh);
}

}  // namespace b200
