// Fused tail of a CFM transformer block (matcha BasicTransformerBlock) for sm_100a: everything between one
// attention and the next as ONE kernel per 128-row tile, with the row tile stationary in shared memory / TMEM and only the
// weights streamed (128 FLOP per byte fetched from L2 instead of 64 for 128 x 128 x K=256 GEMM tiles, and no activation
// round trips through HBM / L2 between the five GEMMs and two LayerNorms):
//
//   OUT : h   = h + attn_o . Wout^T + b_out                      (M128 N256 K512, accumulator ACC0)
//   FF  : x3  = LayerNorm3(h)                      -> bf16 A operand X in TENSOR MEMORY (two values per column along K)
//         for 16 chunks of 64 hidden units:  S_j = X . W0_j^T    (M128 N64 K256, A from TMEM, two S buffers)
//                                            G_j = GELU(S_j + b0) -> bf16, written over S_j in TMEM
//                                            Y  += G_j . W2_j^T  (M128 N256 K64, A from TMEM, ACC)
//         h   = h + Y + b2
//   QKV : xn  = LayerNorm1_next(h)                 -> X (TMEM)
//         qkv = xn . Wqkv_next^T                   (12 chunks of M128 N128 K256, A from TMEM, two accumulators) -> bf16
//
// warp 0 = TMA producer (5 x 32 KB weight ring), warp 1 = tcgen05.mma issuer, warps 2..9 = 256 element-wise threads (thread =
// one row x one column half; TMEM lane quarter = warp % 4).  The activations never touch shared memory: with both operands in
// shared memory the loop was bound by shared-memory bandwidth (the tensor core re-reads the 4 KB A slice for every K = 16 step;
// measured 1.1 us per 64-unit chunk against 0.54 us of tensor time), so X and G live in TMEM and only the weights stream
// through the ring.  S_{j+1} is issued before GELU(S_j) starts and Y += G_{j-1} W2 runs meanwhile; the QKV accumulators
// ping-pong.  Residual rows and QKV tiles move through per-warp TMA pipelines ([32 x 32] fp32 / [32 x 64] bf16 atoms, two
// staging slots per warp): thread-per-row global accesses cost 32 L1 wavefronts per instruction.
// `mode` selects the phases: QKV alone opens a stage (LayerNorm1 + QKV of its first block), OUT|FF|QKV follows every
// attention but the stage's last, OUT|FF closes it.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "cfm_tail.cuh"

namespace {

constexpr int TM = 128, C = 256, CI = 512, CF = 1024, NQKV = 1536;
constexpr int ATOM = 16384;                    // [128 rows x 64 bf16] SWIZZLE_128B tile
constexpr int STAGE = 32768, NST = 5;
constexpr int RING_OFF = 0, IO_OFF = NST * STAGE, BAR_OFF = IO_OFF + 8 * 8192;      // weight ring, 8 warps x 2 x 4 KB I/O staging, barriers
constexpr int SMEM = BAR_OFF + 512 + 1024;     // + barriers / TMEM slot, + alignment slack
// tensor memory: ACC = out-proj / FF accumulator, row cache of the LayerNorms, QKV accumulators (2 x 128); SB = two score buffers
// (GELU output overwrites them in place); XT = LayerNorm output as the bf16 A operand (256 values = 128 columns per row)
constexpr int ACC = 0, SB = 256, XT = 384, TMEM_COLS = 512;
constexpr int THREADS = 320;

__device__ unsigned long long g_tail_trace[16];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TAIL_TRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 64) g_tail_trace[i] = gtime(); } while (0)

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 26); spin++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();   // a lost arrival must fail loudly, never hang the device
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {   // K-major SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart
    uint64_t lo = ((smem_addr >> 4) & 0x3FFF) | (1u << 16);
    uint64_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return lo | (hi << 32);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&f)[32]) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) f[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&f)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])), "r"(__float_as_uint(f[3])),
                   "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])), "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7])),
                   "r"(__float_as_uint(f[8])), "r"(__float_as_uint(f[9])), "r"(__float_as_uint(f[10])), "r"(__float_as_uint(f[11])),
                   "r"(__float_as_uint(f[12])), "r"(__float_as_uint(f[13])), "r"(__float_as_uint(f[14])), "r"(__float_as_uint(f[15])),
                   "r"(__float_as_uint(f[16])), "r"(__float_as_uint(f[17])), "r"(__float_as_uint(f[18])), "r"(__float_as_uint(f[19])),
                   "r"(__float_as_uint(f[20])), "r"(__float_as_uint(f[21])), "r"(__float_as_uint(f[22])), "r"(__float_as_uint(f[23])),
                   "r"(__float_as_uint(f[24])), "r"(__float_as_uint(f[25])), "r"(__float_as_uint(f[26])), "r"(__float_as_uint(f[27])),
                   "r"(__float_as_uint(f[28])), "r"(__float_as_uint(f[29])), "r"(__float_as_uint(f[30])), "r"(__float_as_uint(f[31]))
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&f)[16]) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) f[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&f)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])), "r"(__float_as_uint(f[3])),
                   "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])), "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7])),
                   "r"(__float_as_uint(f[8])), "r"(__float_as_uint(f[9])), "r"(__float_as_uint(f[10])), "r"(__float_as_uint(f[11])),
                   "r"(__float_as_uint(f[12])), "r"(__float_as_uint(f[13])), "r"(__float_as_uint(f[14])), "r"(__float_as_uint(f[15]))
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
    uint32_t x, y;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    a = __uint_as_float(x); b = __uint_as_float(y);
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be overwritten
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }          // writes are complete
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, float& a, float& b, float& c, float& d) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
}
__device__ __forceinline__ void ew_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// GELU of the feed-forward's hidden units, written to tensor memory as bf16.  The reference applies the erf form; here
// 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) with ONE MUFU (tanh.approx) and four FMA-pipe instructions per element.  Its
// distance to the erf form is at most 4.7e-4 ABSOLUTE (around |x| = 2.3), i.e. under a tenth of the bf16 rounding the value gets
// right after (half an ulp at |x| = 2 is 7.8e-3), and the GELU of a 128 x 1024 tile is what the element-wise warps spend their
// time on: the Abramowitz-Stegun erf used before (|error| 1.5e-7, a reciprocal and an exp2 on the XU pipe plus ten FMAs) made the
// feed-forward loop XU-bound at 1.22 us per 64-unit chunk against 0.74 us of weight streaming.  CBX_CFM_TAIL_GELU_ERF=1 (read at
// launch, mode bit 32) selects the erf form for A/B parity runs.
__device__ __forceinline__ float gelu_erf_as(float x) {
    const float a = fabsf(x) * 0.70710678118654752f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.f)));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * a * a));
    const float e = 1.f - p * ex;                                       // erf(|x| / sqrt 2)
    return 0.5f * x * (1.f + copysignf(e, x));
}
__device__ __forceinline__ float gelu_fast(float x) {
    const float u = x * fmaf(x * x, 0.0356774081f, 0.7978845608f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

// A operand from tensor memory (rows = lanes, two bf16 per 32-bit column along K), B from shared memory
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_c, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(tmem_c), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) cfm_tail_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWout,
                                                              const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW2,
                                                              const __grid_constant__ CUtensorMap tmWqkv, const __grid_constant__ CUtensorMap tmH,
                                                              const __grid_constant__ CUtensorMap tmQ, const CfmTailArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sR = base + RING_OFF, sIO = base + IO_OFF, bars = base + BAR_OFF;
    const uint32_t full0 = bars, empty0 = bars + 8 * NST, y_full = bars + 96, x_ready = bars + 104, s_full0 = bars + 112, g_ready0 = bars + 128,
                   q_full0 = bars + 144, q_empty0 = bars + 160, tmem_slot = bars + 176, hbar0 = bars + 192;     // hbar: [8 warps][2 slots]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TM;
    const bool do_out = p.mode & CFM_TAIL_OUT, do_ff = p.mode & CFM_TAIL_FF, do_qkv = p.mode & CFM_TAIL_QKV;
    if (p.seq_T > 0) {      // a tile wholly inside one slab's padding: nothing to do (before any barrier / TMEM state exists)
        const int s0 = m0 / p.seq_T, s1 = min(m0 + TM - 1, p.M - 1) / p.seq_T;
        if (s0 == s1 && m0 - s0 * p.seq_T >= p.seq_len[(s0 >> 1) & 15]) { pdl_launch_dependents(); return; }
    }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < NST; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(y_full, 1); mbar_init(x_ready, 256);
        for (int s = 0; s < 2; s++) {
            mbar_init(s_full0 + 8 * s, 1); mbar_init(g_ready0 + 8 * s, 256);
            mbar_init(q_full0 + 8 * s, 1); mbar_init(q_empty0 + 8 * s, 256);
        }
        for (int s = 0; s < 16; s++) mbar_init(hbar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    TAIL_TRACE(0);
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    TAIL_TRACE(1);
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (elect_one()) {   // ------------------------------------------------------------------ TMA producer (weights + attn_o)
            int g = 0;
            auto acquire = [&](uint32_t bytes) -> uint32_t {     // next ring stage, armed for `bytes`
                const int s = g % NST;
                mbar_wait(empty0 + 8 * s, ((g / NST) & 1) ^ 1);
                mbar_expect_tx(full0 + 8 * s, bytes);
                return (uint32_t)s;
            };
            if (do_out) {
                for (int kb = 0; kb < CI / 64; kb++) {
                    uint32_t s = acquire(ATOM);
                    tma_load_2d(sR + s * STAGE, &tmO, full0 + 8 * s, kb * 64, m0);
                    g++;
                    s = acquire(STAGE);
                    tma_load_2d(sR + s * STAGE, &tmWout, full0 + 8 * s, kb * 64, 0);
                    g++;
                }
            }
            if (do_ff) {
                auto load_w0 = [&](int j) {
                    const uint32_t s = acquire(STAGE);
                    for (int kb = 0; kb < 4; kb++) tma_load_2d(sR + s * STAGE + kb * 8192, &tmW0, full0 + 8 * s, kb * 64, j * 64);
                    g++;
                };
                load_w0(0);
                for (int j = 0; j < CF / 64; j++) {
                    if (j + 1 < CF / 64) load_w0(j + 1);
                    const uint32_t s = acquire(STAGE);
                    tma_load_2d(sR + s * STAGE, &tmW2, full0 + 8 * s, j * 64, 0);
                    g++;
                }
            }
            if (do_qkv) {
                for (int c = 0; c < NQKV / 128; c++)
                    for (int kp = 0; kp < 2; kp++) {      // two k-blocks of [128 rows x 64] per stage
                        const uint32_t s = acquire(STAGE);
                        tma_load_2d(sR + s * STAGE, &tmWqkv, full0 + 8 * s, (2 * kp) * 64, c * 128);
                        tma_load_2d(sR + s * STAGE + ATOM, &tmWqkv, full0 + 8 * s, (2 * kp + 1) * 64, c * 128);
                        g++;
                    }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {   // ------------------------------------------------------------------ MMA issuer
            constexpr uint32_t ID = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TM >> 4) << 24);   // D f32, A = B = bf16, K-major, M = 128
            constexpr uint32_t ID256 = ID | ((uint32_t)(256 >> 3) << 17), ID128 = ID | ((uint32_t)(128 >> 3) << 17), ID64 = ID | ((uint32_t)(64 >> 3) << 17);
            int g = 0, xr = 0;
            auto stage_wait = [&]() -> uint32_t {
                const int s = g % NST;
                mbar_wait(full0 + 8 * s, (g / NST) & 1);
                tc_fence_after();
                return sR + s * STAGE;
            };
            auto stage_free = [&]() { umma_commit(empty0 + 8 * (g % NST)); g++; };
            if (do_out) {      // ACC = attn_o . Wout^T: both operands from shared memory
                for (int kb = 0; kb < CI / 64; kb++) {
                    const uint32_t a = stage_wait();
                    const int ga = g;
                    g++;
                    const uint32_t b = stage_wait();
                    const uint64_t ad = umma_desc(a), bd = umma_desc(b);
#pragma unroll
                    for (int k = 0; k < 4; k++) umma_bf16(tmem_base + ACC, ad + 2 * k, bd + 2 * k, ID256, (kb | k) != 0);
                    umma_commit(empty0 + 8 * (ga % NST));
                    stage_free();
                }
                umma_commit(y_full);
            }
            if (do_ff) {
                mbar_wait(x_ready, xr & 1); xr++;
                tc_fence_after();
                auto issue_s = [&](int j) {      // S_j = X . W0_j^T into S buffer j & 1; X (the LayerNorm output) is read from TMEM
                    const uint32_t w = stage_wait();
#pragma unroll
                    for (int kk = 0; kk < 16; kk++) {
                        if ((p.mode & 16) && (kk & 1)) continue;     // debug: half the S instructions (timing experiment, wrong results)
                        const uint64_t bd = umma_desc(w + (kk >> 2) * 8192) + 2 * (kk & 3);
                        umma_ts_bf16(tmem_base + SB + 64 * (j & 1), tmem_base + XT + kk * 8, bd, ID64, kk != 0);
                    }
                    umma_commit(s_full0 + 8 * (j & 1));
                    stage_free();
                };
                issue_s(0);
                for (int j = 0; j < CF / 64; j++) {
                    // S buffer (j+1)&1 holds G_{j-1}: Y += G_{j-1} W2 was issued one iteration ago and the tensor pipe runs in order
                    if (j + 1 < CF / 64) issue_s(j + 1);
                    mbar_wait(g_ready0 + 8 * (j & 1), (j >> 1) & 1);
                    tc_fence_after();
                    const uint32_t w = stage_wait();
                    const uint64_t bd = umma_desc(w);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++)       // G_j sits in its S buffer: hidden units 0..31 in columns 0..15, 32..63 in columns 32..47
                        umma_ts_bf16(tmem_base + ACC, tmem_base + SB + 64 * (j & 1) + (kk < 2 ? kk * 8 : 32 + (kk - 2) * 8), bd + 2 * kk, ID256, (j | kk) != 0);
                    stage_free();
                }
                umma_commit(y_full);
            }
            if (do_qkv) {
                mbar_wait(x_ready, xr & 1); xr++;
                tc_fence_after();
                for (int c = 0; c < NQKV / 128; c++) {
                    const int a = c & 1;
                    mbar_wait(q_empty0 + 8 * a, ((c >> 1) & 1) ^ 1);
                    tc_fence_after();
                    for (int kp = 0; kp < 2; kp++) {
                        const uint32_t w = stage_wait();
#pragma unroll
                        for (int kk = 0; kk < 8; kk++) {
                            const uint64_t bd = umma_desc(w + (kk >> 2) * ATOM) + 2 * (kk & 3);
                            umma_ts_bf16(tmem_base + ACC + 128 * a, tmem_base + XT + (kp * 8 + kk) * 8, bd, ID128, (kp | kk) != 0);
                        }
                        stage_free();
                    }
                    umma_commit(q_full0 + 8 * a);
                }
            }
        }
    } else {   // ---------------------------------------------------------------------------------- element-wise warps 2..9
        const int wid = warp - 2, hh = wid >> 2, q = warp & 3;  // column half, TMEM lane quarter
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t swz = (uint32_t)(lane & 7);
        const uint32_t io = sIO + wid * 8192;                   // this warp's two 4 KB staging slots
        const uint32_t hb = hbar0 + wid * 16;
        const int mrow = m0 + q * 32;                           // first global row of this warp's 32-row slab
        int yf = 0, qslot = 0;

        // One pass over the warp's slab of the residual rows: v = acc (+ bias) + h -> h (if `store`), cached in ACC; then
        // LayerNorm(v) -> X (bf16, TMEM).  Every warp runs its own TMA pipeline over [32 rows x 32 cols] fp32 atoms (two staging
        // slots, results written back in place and stored from there): no block barriers, the eight warps overlap each other's
        // latencies.  The warp pair (q, hh = 0 / 1) of a row quarter takes the even / odd atoms.
        auto row_pass = [&](bool has_acc, const float* bias, bool store, const float* gamma, const float* beta, bool make_x) {
            if (lane == 0) {
                bulk_wait_all();               // this warp's earlier stores of h are complete before the rows are read again
                for (int k = 0; k < 2; k++) { mbar_expect_tx(hb + 8 * k, 4096); tma_load_2d(io + k * 4096, &tmH, hb + 8 * k, (hh + 2 * k) * 32, mrow); }
            }
            float sum = 0.f, sq = 0.f;
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                const int c0 = (hh + 2 * k) * 32, slot = k & 1;
                float v[32];
                if (has_acc) tmem_ld32(trow + ACC + c0, v);
                else {
#pragma unroll
                    for (int i = 0; i < 32; i++) v[i] = 0.f;
                }
                if (bias) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) { const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + i); v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w; }
                }
                mbar_wait(hb + 8 * slot, (k >> 1) & 1);
                const uint32_t row = io + slot * 4096 + lane * 128;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    float x0, x1, x2, x3;
                    ld_shared_v4(row + (((uint32_t)c ^ swz) << 4), x0, x1, x2, x3);
                    v[4 * c] += x0; v[4 * c + 1] += x1; v[4 * c + 2] += x2; v[4 * c + 3] += x3;
                }
                if (make_x) {
#pragma unroll
                    for (int i = 0; i < 32; i++) { sum += v[i]; sq = fmaf(v[i], v[i], sq); }
                    tmem_st32(trow + ACC + c0, v);
                }
                if (store) {
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)c ^ swz) << 4)),
                                     "f"(v[4 * c]), "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3]) : "memory");
                    fence_async_smem();
                }
                __syncwarp();
                if (lane == 0) {
                    if (store) tma_store_2d(&tmH, io + slot * 4096, c0, mrow);
                    if (k + 2 < 4) {
                        if (store) bulk_wait_read0();      // the store has drained the slot
                        mbar_expect_tx(hb + 8 * slot, 4096);
                        tma_load_2d(io + slot * 4096, &tmH, hb + 8 * slot, (hh + 2 * (k + 2)) * 32, mrow);
                    }
                }
            }
            if (!make_x) return;
            // the two warps of a row quarter exchange their partial sums through two spare TMEM columns of the row's lane
            tmem_st2(trow + SB + hh * 2, sum, sq);
            tc_fence_before();
            asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
            tc_fence_after();
            {
                float s2, q2;
                tmem_ld2(trow + SB + (hh ^ 1) * 2, s2, q2);
                sum += s2; sq += q2;
            }
            const float mean = sum * (1.f / C);
            const float rstd = rsqrtf(fmaxf(sq * (1.f / C) - mean * mean, 0.f) + 1e-5f);
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                const int c0 = (hh + 2 * k) * 32;
                float v[32];
                tmem_ld32(trow + ACC + c0, v);
                float pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 gg = *reinterpret_cast<const float4*>(gamma + c0 + i), bb = *reinterpret_cast<const float4*>(beta + c0 + i);
                    pk[i >> 1] = __uint_as_float(pack_bf16((v[i] - mean) * rstd * gg.x + bb.x, (v[i + 1] - mean) * rstd * gg.y + bb.y));
                    pk[(i >> 1) + 1] = __uint_as_float(pack_bf16((v[i + 2] - mean) * rstd * gg.z + bb.z, (v[i + 3] - mean) * rstd * gg.w + bb.w));
                }
                tmem_st16(trow + XT + (c0 >> 1), pk);      // K-major A operand: two bf16 per column
            }
            tc_fence_before();
            mbar_arrive(x_ready);
        };

        if (do_out) {
            mbar_wait(y_full, yf & 1); yf++;
            tc_fence_after();
        }
        TAIL_TRACE(2);
        if (do_ff) {
            row_pass(do_out, do_out ? p.b_out : nullptr, do_out, p.ln3_g, p.ln3_b, true);
            TAIL_TRACE(3);
            for (int j = 0; j < CF / 64; j++) {
                mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);      // S_j is complete (and, the pipe being in order, G_{j-2} has been consumed)
                tc_fence_after();
                float v[32];
                tmem_ld32(trow + SB + 64 * (j & 1) + hh * 32, v);
                const float* b0 = p.b0 + j * 64 + hh * 32;
                float pk[16];
                if (p.mode & 8) {      // debug: identity instead of GELU (locates the bottleneck of the loop)
#pragma unroll
                    for (int i = 0; i < 32; i += 2) pk[i >> 1] = __uint_as_float(pack_bf16(v[i], v[i + 1]));
                } else {
                    if (p.mode & 32) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 bb = *reinterpret_cast<const float4*>(b0 + i);
                            pk[i >> 1] = __uint_as_float(pack_bf16(gelu_erf_as(v[i] + bb.x), gelu_erf_as(v[i + 1] + bb.y)));
                            pk[(i >> 1) + 1] = __uint_as_float(pack_bf16(gelu_erf_as(v[i + 2] + bb.z), gelu_erf_as(v[i + 3] + bb.w)));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 bb = *reinterpret_cast<const float4*>(b0 + i);
                            pk[i >> 1] = __uint_as_float(pack_bf16(gelu_fast(v[i] + bb.x), gelu_fast(v[i + 1] + bb.y)));
                            pk[(i >> 1) + 1] = __uint_as_float(pack_bf16(gelu_fast(v[i + 2] + bb.z), gelu_fast(v[i + 3] + bb.w)));
                        }
                    }
                }
                tmem_st16(trow + SB + 64 * (j & 1) + hh * 32, pk);   // G_j overwrites the first half of this thread's own score columns
                tc_fence_before();
                mbar_arrive(g_ready0 + 8 * (j & 1));
            }
            TAIL_TRACE(4);
            mbar_wait(y_full, yf & 1); yf++;
            tc_fence_after();
            TAIL_TRACE(5);
            // h = h + Y + b2 (and, for the QKV phase, LayerNorm1 of the next block)
            row_pass(true, p.b2, true, p.ln1_g, p.ln1_b, do_qkv);
            TAIL_TRACE(6);
        } else if (do_qkv) {
            row_pass(do_out, do_out ? p.b_out : nullptr, do_out, p.ln1_g, p.ln1_b, true);
        } else if (do_out) {
            row_pass(true, p.b_out, true, nullptr, nullptr, false);
        }
        if (do_qkv) {
            // accumulator -> bf16 atoms [32 rows x 64 cols] in the warp's staging slots -> TMA store (coalesced by the copy engine)
            for (int c = 0; c < NQKV / 128; c++) {
                const int a = c & 1;
                mbar_wait(q_full0 + 8 * a, (c >> 1) & 1);
                tc_fence_after();
                float v0[32], v1[32];
                tmem_ld32(trow + ACC + 128 * a + hh * 64, v0);
                tmem_ld32(trow + ACC + 128 * a + hh * 64 + 32, v1);
                tc_fence_before();
                mbar_arrive(q_empty0 + 8 * a);             // the accumulator is in registers: hand it back to the tensor core
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the store that last used this slot has drained it
                __syncwarp();
                const uint32_t row = io + qslot * 4096 + lane * 128;
#pragma unroll
                for (int c8 = 0; c8 < 4; c8++)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)c8 ^ swz) << 4)),
                                 "r"(pack_bf16(v0[8 * c8], v0[8 * c8 + 1])), "r"(pack_bf16(v0[8 * c8 + 2], v0[8 * c8 + 3])),
                                 "r"(pack_bf16(v0[8 * c8 + 4], v0[8 * c8 + 5])), "r"(pack_bf16(v0[8 * c8 + 6], v0[8 * c8 + 7])) : "memory");
#pragma unroll
                for (int c8 = 0; c8 < 4; c8++)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)(c8 + 4) ^ swz) << 4)),
                                 "r"(pack_bf16(v1[8 * c8], v1[8 * c8 + 1])), "r"(pack_bf16(v1[8 * c8 + 2], v1[8 * c8 + 3])),
                                 "r"(pack_bf16(v1[8 * c8 + 4], v1[8 * c8 + 5])), "r"(pack_bf16(v1[8 * c8 + 6], v1[8 * c8 + 7])) : "memory");
                fence_async_smem();
                __syncwarp();
                if (lane == 0) tma_store_2d(&tmQ, io + qslot * 4096, c * 128 + hh * 64, mrow);
                qslot ^= 1;
            }
        }
        TAIL_TRACE(7);
        if (lane == 0) bulk_wait_all();       // shared memory must outlive the bulk stores
        TAIL_TRACE(8);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
bool g_ok = false;
long g_launches = 0;

bool map2d(CUtensorMap* tm, const void* ptr, long cols, long rows, long ld, int box_rows, bool f32 = false) {
    cuuint64_t dim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)box_rows}, es[2] = {1, 1};       // 128-byte rows either way
    return g_encode(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

void cfm_tail_init() {
    g_ok = false;
    if (const char* d = getenv("CBX_DISABLE_CFM_TAIL")) { if (d[0] == '1') return; }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return;
    g_encode = (EncodeFn)fn;
    CBX_CHECK(cudaFuncSetAttribute(cfm_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    g_ok = true;
}

bool cfm_tail_available() { return g_ok; }

void cfm_tail_weights(CfmTailWeights& w, const bf16* wout, const bf16* w0, const bf16* w2, const bf16* wqkv) {
    CBX_REQUIRE(g_ok, "cfm_tail: not initialised");
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    bool ok = true;
    if (wout) ok = ok && map2d(reinterpret_cast<CUtensorMap*>(w.out), wout, CI, C, CI, 256);
    if (w0) ok = ok && map2d(reinterpret_cast<CUtensorMap*>(w.w0), w0, C, CF, C, 64);
    if (w2) ok = ok && map2d(reinterpret_cast<CUtensorMap*>(w.w2), w2, CF, C, CF, 256);
    if (wqkv) ok = ok && map2d(reinterpret_cast<CUtensorMap*>(w.qkv), wqkv, C, NQKV, C, 128);
    CBX_REQUIRE(ok, "cfm_tail: cuTensorMapEncodeTiled failed for a weight");
    w.has_out = wout != nullptr; w.has_ff = w0 != nullptr && w2 != nullptr; w.has_qkv = wqkv != nullptr;
}

// blk: the block whose attention just ran (OUT / FF weights); nxt: the block whose LayerNorm1 + QKV follow (QKV weights)
void launch_cfm_tail(const CfmTailArgs& a, const bf16* attn_o, const CfmTailWeights* blk, const CfmTailWeights* nxt, cudaStream_t st) {
    CBX_REQUIRE(g_ok, "cfm_tail: not initialised");
    CBX_REQUIRE(a.M > 0 && a.h && a.mode != 0, "cfm_tail: bad arguments");
    CBX_REQUIRE(!(a.mode & CFM_TAIL_OUT) || (attn_o && blk && blk->has_out && a.b_out), "cfm_tail: OUT phase needs attn_o and the out projection");
    CBX_REQUIRE(!(a.mode & CFM_TAIL_FF) || (blk && blk->has_ff && a.b0 && a.b2 && a.ln3_g && a.ln3_b), "cfm_tail: FF phase needs the feed-forward weights");
    CBX_REQUIRE(!(a.mode & CFM_TAIL_QKV) || (nxt && nxt->has_qkv && a.ln1_g && a.ln1_b && a.qkv), "cfm_tail: QKV phase needs the next block's weights");
    alignas(64) CUtensorMap tmO, tmH, tmQ;
    CBX_REQUIRE(map2d(&tmH, a.h, C, a.M, C, 32, true), "cfm_tail: tensor map for the residual rows");
    if (a.mode & CFM_TAIL_QKV) CBX_REQUIRE(map2d(&tmQ, a.qkv, NQKV, a.M, NQKV, 32), "cfm_tail: tensor map for the qkv output");
    else tmQ = tmH;
    const CfmTailWeights* any = blk ? blk : nxt;
    if (a.mode & CFM_TAIL_OUT) CBX_REQUIRE(map2d(&tmO, attn_o, CI, a.M, CI, 128), "cfm_tail: tensor map for attn_o");
    else tmO = *reinterpret_cast<const CUtensorMap*>(any->has_qkv ? any->qkv : any->out);      // never dereferenced by the kernel
    const CUtensorMap* dummy = &tmO;
    const CUtensorMap* mo = (blk && blk->has_out) ? reinterpret_cast<const CUtensorMap*>(blk->out) : dummy;
    const CUtensorMap* m0 = (blk && blk->has_ff) ? reinterpret_cast<const CUtensorMap*>(blk->w0) : dummy;
    const CUtensorMap* m2 = (blk && blk->has_ff) ? reinterpret_cast<const CUtensorMap*>(blk->w2) : dummy;
    const CUtensorMap* mq = (nxt && nxt->has_qkv) ? reinterpret_cast<const CUtensorMap*>(nxt->qkv) : dummy;
    double flops = 0;
    if (a.mode & CFM_TAIL_OUT) flops += 2.0 * a.M * C * CI;
    if (a.mode & CFM_TAIL_FF) flops += 4.0 * a.M * C * CF;
    if (a.mode & CFM_TAIL_QKV) flops += 2.0 * a.M * C * NQKV;
    ProfScope ps(PC_GEMM, flops, st);
    static const bool gelu_erf = [] { const char* e = getenv("CBX_CFM_TAIL_GELU_ERF"); return e && e[0] == '1'; }();
    CfmTailArgs aa = a;
    if (gelu_erf) aa.mode |= 32;
    launch_pdl(cfm_tail_kernel, dim3(cdiv(a.M, TM)), dim3(THREADS), SMEM, st, tmO, *mo, *m0, *m2, *mq, tmH, tmQ, aa);
    CBX_CHECK(cudaGetLastError());
    g_launches++;
}

extern "C" long long cbx_cfm_tail_launches(void) { return g_launches; }

// debug: %globaltimer stamps (ns) of CTA 0's element-wise leader in the last launch: start, dependencies resolved, out-proj accumulator
// ready, LayerNorm3 done, GELU loop done, FF accumulator ready, second row pass done, QKV tiles stored, bulk stores complete
extern "C" int cbx_cfm_tail_trace(unsigned long long* out_h) { return cudaMemcpyFromSymbol(out_h, g_tail_trace, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : 1; }
