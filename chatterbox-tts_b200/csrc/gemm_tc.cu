// tcgen05 / TMEM / TMA GEMM for sm_100a: C[128 x BN] tile per CTA, operands staged by TMA into a
// SWIZZLE_128B smem ring, one elected thread issues tcgen05.mma (M=128, N=BN, K=16 per instruction) into a
// TMEM accumulator, then all eight warps read it back with tcgen05.ld and apply the fused epilogue.  Two CTAs per SM.
// The A operand is described by a 4-D tensor map (channel, row, tap, batch): row stride lda, tap stride
// tap_stride -- conv1d over time-major channels-last activations becomes plain TMA boxes (implicit GEMM).
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace {

constexpr int TM = 128, TK = 64, A_BYTES = TM * TK * 2;
__device__ unsigned long long g_tc_trace[16];
template <int BN> struct TcCfg {
    // shallow rings (K is 256..1024 here) so that TWO CTAs fit one SM: the prologue / epilogue of one tile overlaps the
    // main loop of the other, which is what multi-wave (batched) launches need; the ring doubles as the epilogue tile
    static constexpr int STAGES = BN <= 64 ? 4 : 3;
    static constexpr int B_BYTES = BN * TK * 2;
    static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TC_TRACE(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_trace[i] = gtime(); } while (0)

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 24); spin++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();   // a lost arrival must fail loudly, never hang the device
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t lo = ((smem_addr >> 4) & 0x3FFF) | (1u << 16);                     // start address, LBO = 1 (ignored for swizzled K-major)
    uint64_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);                       // SBO = 1024 B, descriptor version 1 (sm_100), SWIZZLE_128B
    return lo | (hi << 32);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// LayerNorm of the finished row fused into the epilogue (LN = true; N == cluster size x BN, the CTAs of a row's tiles form
// one thread-block cluster along N): every CTA writes its part of the fp32 row (bias + residual), computes {mean, M2}
// of its BN columns, the cluster exchanges them through distributed shared memory, and every CTA writes its part of
// LayerNorm(row) * gamma + beta as bf16 -- the separate norm kernel between a residual GEMM and the next projection
// (two per CFM transformer block, a fifth of all S3Gen launches) disappears.
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float2 ld_cluster_f2(const float* local_smem_ptr, uint32_t cta_rank) {
    uint32_t la = smem_u32(local_smem_ptr), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(cta_rank));
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(ra) : "memory");
    return v;
}

template <int BN, int ACT, int ACT2, bool LN = false>
__global__ void __launch_bounds__(256, 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                         const GemmParams p, int kc_blocks, int w_batched) {
    using C = TcCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + C::STAGES * A_BYTES;
    const uint32_t bars = sB + C::STAGES * C::B_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * C::STAGES, tmem_full = bars + 16 * C::STAGES, tmem_slot = tmem_full + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * BN, b = blockIdx.z;
    const int KB = p.K / TK;
    if (!LN && p.row_div > 0 && m0 >= p.row_len[(b / p.row_div) & 15]) { pdl_launch_dependents(); return; }     // tile in a slab's padding

    if (threadIdx.x == 0) TC_TRACE(0);
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pdl_wait();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (threadIdx.x == 0) TC_TRACE(1);
    // warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer; afterwards all 8 warps run the epilogue
    // (TMEM lane quarter = warp % 4, column half = warp / 4)
    if (warp == 0) {
        if (elect_one()) {   // TMA producer
            for (int kb = 0; kb < KB; kb++) {
                const int s = kb % C::STAGES, ph = (kb / C::STAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                mbar_expect_tx(full0 + 8 * s, A_BYTES + C::B_BYTES);
                const int tap = kb / kc_blocks, ci = (kb % kc_blocks) * TK;
                tma_load_4d(sA + s * A_BYTES, &tmA, full0 + 8 * s, ci, m0, tap, b);
                tma_load_3d(sB + s * C::B_BYTES, &tmW, full0 + 8 * s, kb * TK, n0, w_batched ? b : 0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {   // MMA issuer
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            for (int kb = 0; kb < KB; kb++) {
                const int s = kb % C::STAGES, ph = (kb / C::STAGES) & 1;
                mbar_wait(full0 + 8 * s, ph);
                if (kb == 0) TC_TRACE(2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t ad = umma_desc(sA + s * A_BYTES), bd = umma_desc(sB + s * C::B_BYTES);
#pragma unroll
                for (int k = 0; k < TK / 16; k++) umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);   // +32 B per K=16 step
                umma_commit(empty0 + 8 * s);     // smem slot is free once these MMAs retire
            }
            umma_commit(tmem_full);
            TC_TRACE(3);
        }
        __syncwarp();
    }
    {
        const int q = warp & 3, half = warp >> 2, et = threadIdx.x;
        constexpr int LDT = BN + 4, CPR = BN / 8, RSTEP = 256 / CPR;
        const int cc = (et % CPR) * 8;
        ColOps co;
        load_colops(p, b, n0 + cc, co);      // overlaps the main loop
        mbar_wait(tmem_full, 0);
        if (threadIdx.x == 64) TC_TRACE(4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // TMEM -> registers (thread = row) -> shared staging tile; the operand ring is idle by now and is reused for it
        float* tile = reinterpret_cast<float*>(smem_raw + (sA - smem_u32(smem_raw)));
        {
            float* trow = tile + (q * 32 + lane) * LDT + half * (BN / 2);
#pragma unroll
            for (int g = 0; g < BN / 64; g++) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + half * (BN / 2) + g * 32;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                               "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                               "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 8; c++)
                    *reinterpret_cast<uint4*>(trow + g * 32 + 4 * c) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            }
        }
        if (threadIdx.x == 64) TC_TRACE(7);
        __syncthreads();
        if (threadIdx.x == 64) TC_TRACE(8);
        if constexpr (!LN) {
            epilogue_rows<ACT, ACT2>(p, b, m0, et / CPR, RSTEP, TM, tile, LDT, cc, co);
        } else {
            float* stats = tile + TM * LDT;                       // [TM][2]: {mean, M2} of this CTA's BN columns per row
            const int n = n0 + cc, sub = et % CPR;
            uint32_t cs, rank;
            asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
            asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
            (void)rank;
            float gam[8], bet[8];
            ld8(p.ln_gamma + n, 8, gam);
            ld8(p.ln_beta + n, 8, bet);
            // pass 1: v = acc + bias + residual -> fp32 row (global) and back into the staged tile; per-row {mean, M2}.
            // All residual rows of this thread are requested first: one L2 round trip instead of one per row.
            constexpr int NR = TM / RSTEP;
            float rr[NR][8];
#pragma unroll
            for (int k = 0; k < NR; k++) {
                const int m = m0 + et / CPR + k * RSTEP;
                if (m < p.M && p.res) ld8(p.res + (long)b * p.r_bs + (long)m * p.ldr + n, 8, rr[k]);
                else {
#pragma unroll
                    for (int i = 0; i < 8; i++) rr[k][i] = 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < NR; k++) {
                const int r = et / CPR + k * RSTEP, m = m0 + r;
                const bool valid = m < p.M;
                float v[8];
                const float4 x0 = *reinterpret_cast<const float4*>(tile + r * LDT + cc), x1 = *reinterpret_cast<const float4*>(tile + r * LDT + cc + 4);
                v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) { v[i] = v[i] + co.t0[i] + rr[k][i]; s += v[i]; }
                if (valid && p.outF) st8(p.outF + (long)b * p.c_bs + (long)m * p.ldc + n, 8, v);
                *reinterpret_cast<float4*>(tile + r * LDT + cc) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(tile + r * LDT + cc + 4) = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
                for (int o = CPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mean_c = s * (1.f / BN);
                float d2 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) { const float d = v[i] - mean_c; d2 += d * d; }
#pragma unroll
                for (int o = CPR / 2; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
                if (sub == 0) *reinterpret_cast<float2*>(stats + r * 2) = make_float2(mean_c, d2);
            }
            cluster_sync_all();
            // pass 2: merge the cluster's partial statistics (equal counts: Chan's formula), normalise this CTA's columns
            for (int r = et / CPR; r < TM; r += RSTEP) {
                const int m = m0 + r;
                float mc[4], dc[4], mean = 0.f;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (c < (int)cs) { const float2 t = ld_cluster_f2(stats + r * 2, (uint32_t)c); mc[c] = t.x; dc[c] = t.y; mean += t.x; }
                }
                mean /= (float)cs;
                float m2 = 0.f;
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c < (int)cs) { const float d = mc[c] - mean; m2 += dc[c] + (float)BN * d * d; }
                const float rstd = rsqrtf(m2 / (float)(BN * cs) + p.ln_eps);
                if (m < p.M) {
                    const float4 x0 = *reinterpret_cast<const float4*>(tile + r * LDT + cc), x1 = *reinterpret_cast<const float4*>(tile + r * LDT + cc + 4);
                    float y[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) y[i] = (y[i] - mean) * rstd * gam[i] + bet[i];
                    st8b(p.outB2 + (long)b * p.c2_bs + (long)m * p.ldc2 + n, 8, y);
                }
            }
            cluster_sync_all();   // nobody leaves while a peer may still read its statistics
        }
    }
    if (threadIdx.x == 64) TC_TRACE(5);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) TC_TRACE(6);
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent variant for multi-wave launches (batched S3Gen: M = calls x 2 x T rows).  One CTA per SM walks a static
// list of output tiles; the TMA warp keeps the operand ring full ACROSS tiles, the MMA warp alternates between two TMEM
// accumulators, and eight dedicated epilogue warps drain accumulator i while the tensor core fills accumulator i^1 --
// the per-tile prologue / epilogue latency that bounds the one-tile kernel above disappears from the critical path.
template <int BN> struct TcPCfg {
    static constexpr int STAGES = 4;
    static constexpr int B_BYTES = BN * TK * 2;
    static constexpr int RING = STAGES * (A_BYTES + B_BYTES);
    static constexpr int LDT = BN + 4;
    static constexpr int TILE_BYTES = TM * LDT * 4;
    static constexpr int SMEM = RING + TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
};

template <int BN, int ACT, int ACT2>
__global__ void __launch_bounds__(320, 1) gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                                    const GemmParams p, int kc_blocks, int w_batched, int tiles_m, int tiles_n, int n_tiles) {
    using C = TcPCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + C::STAGES * A_BYTES;
    const uint32_t stile = base + C::RING;
    const uint32_t bars = stile + C::TILE_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * C::STAGES, tfull0 = bars + 16 * C::STAGES, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = p.K / TK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pdl_wait();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // tile t -> (n tile fastest: neighbouring CTAs share the A rows in L2, m tile, batch)
    auto tile_coords = [&](int t, int& m0, int& n0, int& b) { n0 = (t % tiles_n) * BN; m0 = ((t / tiles_n) % tiles_m) * TM; b = t / (tiles_n * tiles_m); };

    if (warp == 0) {
        if (elect_one()) {   // TMA producer: the ring runs ahead into the next tile
            int g = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int m0, n0, b;
                tile_coords(t, m0, n0, b);
                for (int kb = 0; kb < KB; kb++, g++) {
                    const int s = g % C::STAGES, ph = (g / C::STAGES) & 1;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    mbar_expect_tx(full0 + 8 * s, A_BYTES + C::B_BYTES);
                    const int tap = kb / kc_blocks, ci = (kb % kc_blocks) * TK;
                    tma_load_4d(sA + s * A_BYTES, &tmA, full0 + 8 * s, ci, m0, tap, b);
                    tma_load_3d(sB + s * C::B_BYTES, &tmW, full0 + 8 * s, kb * TK, n0, w_batched ? b : 0);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {   // MMA issuer: accumulator it & 1
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            int g = 0, it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
                const int acc = it & 1;
                mbar_wait(tempty0 + 8 * acc, ((it >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < KB; kb++, g++) {
                    const int s = g % C::STAGES, ph = (g / C::STAGES) & 1;
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t ad = umma_desc(sA + s * A_BYTES), bd = umma_desc(sB + s * C::B_BYTES);
#pragma unroll
                    for (int k = 0; k < TK / 16; k++) umma_bf16(tmem_base + acc * BN, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(empty0 + 8 * s);
                }
                umma_commit(tfull0 + 8 * acc);
            }
        }
    } else {               // epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4
        const int q = warp & 3, half = (warp - 2) >> 2, et = threadIdx.x - 64;
        constexpr int LDT = C::LDT, CPR = BN / 8, RSTEP = 256 / CPR;
        const int cc = (et % CPR) * 8;
        float* tile = reinterpret_cast<float*>(smem_raw + (stile - smem_u32(smem_raw)));
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
            const int acc = it & 1;
            int m0, n0, b;
            tile_coords(t, m0, n0, b);
            ColOps co;
            load_colops(p, b, n0 + cc, co);
            mbar_wait(tfull0 + 8 * acc, (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                float* trow = tile + (q * 32 + lane) * LDT + half * (BN / 2);
#pragma unroll
                for (int gq = 0; gq < BN / 64; gq++) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * (BN / 2) + gq * 32;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        *reinterpret_cast<uint4*>(trow + gq * 32 + 4 * c) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                }
            }
            // this thread's part of the accumulator is in shared memory: hand it back to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8 * acc) : "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");     // the staged tile is complete
            epilogue_rows<ACT, ACT2>(p, b, m0, et / CPR, RSTEP, TM, tile, LDT, cc, co);
            asm volatile("bar.sync 1, 256;" ::: "memory");     // everyone is done reading it: the next tile may overwrite
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
int g_sms = 148, g_persistent_min = 0;   // tiles from which the persistent kernel is used (0 = never)
bool g_tc_ok = false;
long g_tc_launches = 0;

template <int BN>
bool launch_tc(const GemmParams& p, cudaStream_t st) {
    using C = TcCfg<BN>;
    const int ntaps = p.K / p.kc;
    alignas(64) CUtensorMap tmA, tmW;
    {
        cuuint64_t dim[4] = {(cuuint64_t)p.kc, (cuuint64_t)p.M, (cuuint64_t)ntaps, (cuuint64_t)p.batch};
        cuuint64_t str[3] = {(cuuint64_t)p.lda * 2, (cuuint64_t)(ntaps > 1 ? p.tap_stride : p.lda) * 2, (cuuint64_t)(p.batch > 1 ? p.a_bs : p.lda) * 2};
        cuuint32_t box[4] = {TK, TM, 1, 1}, es[4] = {1, 1, 1, 1};
        if (g_encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)p.A, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
    const int w_batched = (p.batch > 1 && p.w_bs != 0) ? 1 : 0;
    {
        cuuint64_t dim[3] = {(cuuint64_t)p.K, (cuuint64_t)p.N, (cuuint64_t)(w_batched ? p.batch : 1)};
        cuuint64_t str[2] = {(cuuint64_t)p.ldw * 2, (cuuint64_t)(w_batched ? p.w_bs : p.ldw) * 2};
        cuuint32_t box[3] = {TK, BN, 1}, es[3] = {1, 1, 1};
        if (g_encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)p.W, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
    dim3 grid(cdiv(p.M, TM), cdiv(p.N, BN), p.batch);
    bool done = false;
    if (p.ln_gamma) {   // LayerNorm fused into the epilogue: the N tiles of a row form one cluster
        if (p.act != ACT_NONE || p.act2 != ACT_NONE || p.N % BN != 0 || grid.y > 4 || !p.outB2 || !p.ln_beta || p.glu || p.ct_u || p.accumulate ||
            p.out_scale != 1.f || p.bias2 || p.outB) return false;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = grid.y; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        CBX_CHECK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, ACT_NONE, ACT_NONE, true>, tmA, tmW, p, p.kc / TK, w_batched));
        CBX_CHECK(cudaGetLastError());
        g_tc_launches++;
        return true;
    }
    const int n_tiles = (int)grid.x * (int)grid.y * (int)grid.z;
    if (g_persistent_min > 0 && n_tiles >= g_persistent_min) {
        const int ctas = n_tiles < g_sms ? n_tiles : g_sms;
#define CBX_LAUNCHP(A1, A2) if (!done && p.act == A1 && p.act2 == A2) { launch_pdl(gemm_tc_persistent_kernel<BN, A1, A2>, dim3(ctas), dim3(320), TcPCfg<BN>::SMEM, st, tmA, tmW, p, p.kc / TK, w_batched, (int)grid.x, (int)grid.y, n_tiles); done = true; }
        CBX_FOR_ACT_PAIRS(CBX_LAUNCHP)
#undef CBX_LAUNCHP
    }
#define CBX_LAUNCH(A1, A2) if (!done && p.act == A1 && p.act2 == A2) { launch_pdl(gemm_tc_kernel<BN, A1, A2>, grid, dim3(256), C::SMEM, st, tmA, tmW, p, p.kc / TK, w_batched); done = true; }
    CBX_FOR_ACT_PAIRS(CBX_LAUNCH)
#undef CBX_LAUNCH
    if (!done) return false;
    CBX_CHECK(cudaGetLastError());
    g_tc_launches++;
    return true;
}

}  // namespace

void gemm_tc_init() {
    g_tc_ok = false;
    if (const char* d = getenv("CBX_DISABLE_TC")) { if (d[0] == '1') return; }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return;
    g_encode = (EncodeFn)fn;
#define CBX_ATTR(A1, A2) \
    CBX_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<64, A1, A2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::SMEM)); \
    CBX_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<128, A1, A2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM));
    CBX_FOR_ACT_PAIRS(CBX_ATTR)
#undef CBX_ATTR
    CBX_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<64, ACT_NONE, ACT_NONE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::SMEM));
    CBX_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<128, ACT_NONE, ACT_NONE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM));
#define CBX_ATTRP(A1, A2) \
    CBX_CHECK(cudaFuncSetAttribute(gemm_tc_persistent_kernel<64, A1, A2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcPCfg<64>::SMEM)); \
    CBX_CHECK(cudaFuncSetAttribute(gemm_tc_persistent_kernel<128, A1, A2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcPCfg<128>::SMEM));
    CBX_FOR_ACT_PAIRS(CBX_ATTRP)
#undef CBX_ATTRP
    {
        int dev = 0;
        CBX_CHECK(cudaGetDevice(&dev));
        CBX_CHECK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
        // more than two waves of the two-CTAs-per-SM kernel: the persistent one wins (CBX_GEMM_PERSISTENT_MIN overrides, 0 = off)
        const char* e = getenv("CBX_GEMM_PERSISTENT_MIN");
        g_persistent_min = e ? atoi(e) : 2 * 2 * g_sms;
    }
    g_tc_ok = true;
}

// returns false when the problem does not fit the TMA/tcgen05 path (caller falls back to the mma.sync kernel)
bool launch_gemm_tc(const GemmParams& p, cudaStream_t st) {
    if (!g_tc_ok) return false;
    if (p.kc % TK != 0 || p.K % p.kc != 0 || p.K % TK != 0) return false;
    if (p.lda % 8 || p.ldw % 8 || p.tap_stride % 8 || p.a_bs % 8 || p.w_bs % 8) return false;
    if (((uintptr_t)p.A & 15) || ((uintptr_t)p.W & 15)) return false;
    // wide tiles only when they still leave enough CTAs in flight
    const long ctas128 = (long)cdiv(p.M, TM) * cdiv(p.N, 128) * p.batch;
    static const long wide_min = [] { const char* e = getenv("CBX_GEMM_WIDE_MIN"); return e ? atol(e) : 149L; }();
    if (p.N >= 128 && ctas128 >= wide_min) return launch_tc<128>(p, st);
    return launch_tc<64>(p, st);
}

extern "C" long long cbx_gemm_tc_launches(void) { return g_tc_launches; }

extern "C" int cbx_gemm_tc_trace(unsigned long long* out_h) { return cudaMemcpyFromSymbol(out_h, g_tc_trace, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : 1; }
