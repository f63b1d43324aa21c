// tcgen05 flash attention, second generation (sm_100a, head_dim 64): full or causal, optionally ragged (keys >= kv_len masked).
//
// The first kernel (attention_tc.cu) ran the chain S -> softmax -> P (shared memory) -> PV -> O (registers) serially per CTA and
// hid it with two CTAs per SM: 2.2 us per 128 x 128 score block and SM against ~0.55 us of MUFU time.  This one follows the
// Blackwell recipe for that chain:
//
//   * a CTA owns TWO 128-query tiles of one (batch, head) -- groups A and B, four softmax warps each, ONE THREAD PER QUERY ROW
//     (no cross-thread max / sum exchange, no block barrier in the loop); K/V tiles are loaded once for both groups.  Each
//     group has its own MMA-issuing warp, so neither group's chain queues behind the other's barrier waits.
//   * S (fp32, 128 columns) is read out of TMEM once, into registers, and released at once: the tensor core computes
//     S(j+1) while the softmax threads are still exponentiating block j.  P (bf16) goes back INTO TMEM (its own 64 columns)
//     and is the A operand of P.V (tcgen05.mma with A in tensor memory): no shared-memory round trip, no proxy fence.
//   * O lives in TMEM and is accumulated there by the tensor core across all key blocks.  A row's running maximum is only
//     raised when it grew by more than 2^8 (warp-uniform vote); until then P is computed against the stale maximum, which the
//     final division by the row sum (same stale maximum) cancels exactly.  The O rescale (tcgen05.ld / st) is therefore rare.
//   * three-stage TMA rings for K and V.
//   * the producer and issuer threads are chosen with elect.sync, not `lane == 0`: under a lane test the compiler wraps EVERY
//     tcgen05.mma in an election loop with four R2UR broadcasts (~80 cycles per MMA against a 32-64 cycle MMA); after
//     elect.sync it knows the region is single-threaded and issues the MMAs back to back.
//
// TMEM (512 columns): S_A 0..127, S_B 128..255, O_A 256..319, O_B 320..383, P_A 384..447, P_B 448..511.
// warp 0 = TMA producer, warps 1 / 2 = MMA issuers of groups A / B, warps 3..6 = softmax A, warps 7..10 = softmax B
// (TMEM lane quarter = warp % 4).
#include <cuda.h>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"

namespace {

constexpr int D = 64, BQ = 128, BK = 128, THREADS = 352, NS = 3;
constexpr int TILE = 128 * 128;                        // bytes of a [128 rows x 64 bf16] SWIZZLE_128B tile
constexpr int SMEM = (2 + 2 * NS) * TILE + 1024 + 256; // Q_A, Q_B, K ring, V ring, alignment slack, barriers
constexpr int TMEM_COLS = 512, S_COL = 0, O_COL = 256, P_COL = 384;
constexpr float RESCALE_THRESHOLD = 8.f;               // log2 domain: P <= 2^8 against a stale maximum

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row groups 1024 B apart (K-major Q / K tiles and the MN-major V tile)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t lo = ((smem_addr >> 4) & 0x3FFF) | (1u << 16);
    uint64_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return lo | (hi << 32);
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {   // A operand in tensor memory
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(tmem_c), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 consecutive fp32 columns of this thread's TMEM lane; the caller waits (tmem_ld_wait) before touching the registers
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* f) {
    uint32_t* v = reinterpret_cast<uint32_t*>(f);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32_nowait(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                   "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                   "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ unsigned long long g_fa_trace[256];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define FA_TRACE(cond, i) do { if ((p.dbg & 4) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (cond)) g_fa_trace[i] = gtime(); } while (0)

// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial on [0, 1), max relative error 8.8e-5 -- P is
// rounded to bf16 (4e-3) right after).  OPT-IN (CBX_ATTN_FA_DBG bit 4 = 16), measured NEGATIVE on B200: the XU pipe takes 8.0
// cycles per warp-wide ex2 and 3.8 per bf16x2 pack (tools/ubench/mufu.cu), i.e. 1 267 cycles per 128 x 128 score block, and
// moving every second exponential here should halve the ex2 share -- but with the 128 scores of a row held in registers
// there is no room to interleave enough of the 8-deep polynomial chains: the exponential phase went from 860 to 1 150 ns
// per block (T = 668, 16 sequences: 56.7 vs 49.1 us per launch).
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.f);
    float t;
    asm("add.rm.ftz.f32 %0, %1, 0f4B400000;" : "=f"(t) : "f"(x));      // 1.5 * 2^23 + floor(x): the integer sits in the low mantissa bits
    const float f = x - (t - 12582912.f);                                // fractional part in [0, 1)
    float q = fmaf(f, 0.077119089663028717041015625f, 0.227564394474029541015625f);
    q = fmaf(q, f, 0.695146143436431884765625f);
    q = fmaf(q, f, 1.f);
    return __uint_as_float(__float_as_uint(q) + (__float_as_uint(t) << 23));   // q * 2^floor(x)
}

struct AttnFaArgs {
    bf16* o; long ldo, o_bs;
    int T, H, causal; float sl2;   // sl2 = scale * log2(e): scores are kept in the log2 domain
    int kv_len[16]; int kv_div; int dbg;
    const float* relbias; long rb_ld, rb_hs, rb_bs;      // additive score bias (ESPnet relative position): bias[b][h][i][T - 1 - i + j]
};

__global__ void __launch_bounds__(THREADS, 1) attn_fa_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                             const __grid_constant__ CUtensorMap tmV, const AttnFaArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base, sK = base + 2 * TILE, sV = sK + NS * TILE, bars = sV + NS * TILE;
    const uint32_t q_full = bars, k_full = bars + 8, k_empty = k_full + 8 * NS, v_full = k_empty + 8 * NS, v_empty = v_full + 8 * NS,
                   s_full = v_empty + 8 * NS, p_ready = s_full + 16, o_done = p_ready + 16, s_free = o_done + 16, tmem_slot = s_free + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * 2 * BQ;
    const int klen_raw = p.kv_len[(b / p.kv_div) & 15];
    const int Tk = klen_raw > 0 ? klen_raw : p.T;
    pdl_launch_dependents();
    if (q0 >= Tk) return;                              // padded query tiles of a ragged batch: never consumed
    const int ngroups = (q0 + BQ < Tk) ? 2 : 1;        // group B's tile may lie wholly beyond the sequence
    // key blocks this CTA visits: all of them, or (causal) up to the diagonal of its last query row
    const int nkv = p.causal ? min((Tk + BK - 1) / BK, (min(q0 + ngroups * BQ, Tk) + BK - 1) / BK) : (Tk + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < NS; s++) { mbar_init(k_full + 8 * s, 1); mbar_init(k_empty + 8 * s, ngroups); mbar_init(v_full + 8 * s, 1); mbar_init(v_empty + 8 * s, ngroups); }
        for (int g = 0; g < 2; g++) { mbar_init(s_full + 8 * g, 1); mbar_init(p_ready + 8 * g, 128); mbar_init(o_done + 8 * g, 1); mbar_init(s_free + 8 * g, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmV) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (elect_one()) {   // ----------------------------------------------------------------------------- TMA producer
            mbar_expect_tx(q_full, ngroups * TILE);
            for (int g = 0; g < ngroups; g++) tma_load_3d(sQ + g * TILE, &tmQ, q_full, h * D, q0 + g * BQ, b);
            for (int j = 0; j < nkv; j++) {
                const int s = j % NS, ph = (j / NS) & 1;
                mbar_wait(k_empty + 8 * s, ph ^ 1);
                mbar_expect_tx(k_full + 8 * s, TILE);
                tma_load_3d(sK + s * TILE, &tmK, k_full + 8 * s, h * D, j * BK, b);
                mbar_wait(v_empty + 8 * s, ph ^ 1);
                mbar_expect_tx(v_full + 8 * s, TILE);
                tma_load_3d(sV + s * TILE, &tmV, v_full + 8 * s, h * D, j * BK, b);
            }
        }
    } else if (warp == 1 || warp == 2) {
        const int g = warp - 1;                 // one issuing warp per query group: neither group's chain waits behind the other's
        if (g < ngroups && elect_one()) {   // ------------------------------------------------------------------ MMA issuers
            // instruction descriptors: D = f32, A = B = bf16, M = 128; S: N = 128, both K-major; PV: N = 64, B (= V) MN-major
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BK >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
            const uint64_t qd = umma_desc(sQ + g * TILE);
            auto issue_s = [&](int j) {     // S_g(j) = Q_g K_j^T; a K / V stage is free again when BOTH groups' products on it have completed
                const int s = j % NS;
                mbar_wait(k_full + 8 * s, (j / NS) & 1);
                tc_fence_after();
                const uint64_t kd = umma_desc(sK + s * TILE);
#pragma unroll
                for (int k = 0; k < D / 16; k++) umma_ss(tmem_base + S_COL + g * 128, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);   // +32 B per K = 16
                umma_commit(s_full + 8 * g);
                umma_commit(k_empty + 8 * s);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nkv; j++) {
                const int s = j % NS, ph = (j / NS) & 1;
                if (j + 1 < nkv) {
                    mbar_wait(s_free + 8 * g, j & 1);           // S_g(j) sits in the softmax threads' registers: S_g(j+1) runs under their exponentials
                    issue_s(j + 1);
                }
                FA_TRACE(j < 8, 64 + (j * 2 + g) * 3);
                mbar_wait(p_ready + 8 * g, j & 1);              // P_g(j) is in tensor memory, O_g is rescaled
                FA_TRACE(j < 8, 64 + (j * 2 + g) * 3 + 1);
                mbar_wait(v_full + 8 * s, ph);
                tc_fence_after();
                const uint64_t vd = umma_desc(sV + s * TILE);
#pragma unroll
                for (int k = 0; k < BK / 16; k++)                // A: 16 keys = 8 TMEM columns; B: V rows 16k.. = +2048 B
                    umma_ts(tmem_base + O_COL + g * 64, tmem_base + P_COL + g * 64 + 8 * k, vd + (uint64_t)(k * (2048 >> 4)), idesc_pv, (j | k) != 0);
                umma_commit(o_done + 8 * g);
                umma_commit(v_empty + 8 * s);
                FA_TRACE(j < 8, 64 + (j * 2 + g) * 3 + 2);
            }
        }
    } else {               // ------------------------------------------------------------------------------- softmax groups
        const int g = (warp - 3) >> 2;
        if (g < ngroups) {
            const int qr = (warp & 3) * 32 + lane;                                   // row inside the group's tile = TMEM lane
            const int qi = q0 + g * BQ + qr;
            const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
            const uint32_t scol = trow + S_COL + g * 128, ocol = trow + O_COL + g * 64, pcol = trow + P_COL + g * 64;
            float m = -INFINITY, l = 0.f;
            const bool pingpong = ngroups == 2 && !(p.dbg & 8);
            if (pingpong && g == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");      // group A goes first
            // a warp whose 32 query rows all lie beyond the sequence (the tail of the last query tile: 668 rows = 5.2 tiles) only
            // keeps the barrier protocol going: no TMEM traffic, no exponentials.  Its P / O rows hold garbage that is never stored.
            const bool dead = q0 + g * BQ + (warp & 3) * 32 >= Tk;
            for (int j = 0; j < nkv; j++) {
                FA_TRACE(j < 8 && qr == 0 , 128 + (j * 2 + g) * 4);
                mbar_wait(s_full + 8 * g, j & 1);
                tc_fence_after();
                FA_TRACE(j < 8 && qr == 0, 128 + (j * 2 + g) * 4 + 1);
                if (dead) {
                    tc_fence_before();
                    mbar_arrive(s_free + 8 * g);
                    if (j > 0) mbar_wait(o_done + 8 * g, (j - 1) & 1);
                    if (pingpong) {
                        if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
                        if (!(g == 1 && j == nkv - 1)) { if (g == 0) asm volatile("bar.arrive 2, 256;" ::: "memory"); else asm volatile("bar.arrive 1, 256;" ::: "memory"); }
                    }
                    mbar_arrive(p_ready + 8 * g);
                    continue;
                }
                float s[128];
#pragma unroll
                for (int c = 0; c < 4; c++) tmem_ld32_nowait(scol + c * 32, s + c * 32);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(s_free + 8 * g);                                              // the tensor core may overwrite S_g
                FA_TRACE(j < 8 && qr == 0, 128 + (j * 2 + g) * 4 + 2);
                const int kbase = j * BK;
                if (p.relbias && qi < p.T) {          // rel-pos term (q + v) P^T, rel-shifted by the index: consecutive keys are consecutive floats
                    const float* br = p.relbias + (long)b * p.rb_bs + (long)h * p.rb_hs + (long)qi * p.rb_ld + (p.T - 1 - qi + kbase);
                    const int nk = min(BK, Tk - kbase);
#pragma unroll
                    for (int i = 0; i < 128; i++) if (i < nk) s[i] += __ldg(br + i);
                }
                const bool masked = kbase + BK > Tk || (p.causal && kbase + BK - 1 > q0 + g * BQ);      // block with masked keys (CTA-group-uniform)
                if (masked) {
                    const int lim = p.causal ? min(Tk, qi + 1) : Tk;                      // keys < lim are visible to this row
#pragma unroll
                    for (int i = 0; i < 128; i++) if (kbase + i >= lim) s[i] = -INFINITY;
                }
                float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
                for (int i = 4; i < 128; i += 4) { mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1])); mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3])); }
                const float mnew = fmaxf(m, fmaxf(mx0, mx1) * p.sl2);                     // sl2 > 0: scaling commutes with the max
                // raise the running maximum only when some row of the warp needs it (tcgen05.ld / st are warp-collective)
                FA_TRACE(j < 8 && qr == 0, 192 + (j * 2 + g) * 4);
                if (j > 0) {                                                              // P_g(j-1) V has landed in O_g: the P buffer is free
                    mbar_wait(o_done + 8 * g, (j - 1) & 1);
                    tc_fence_after();
                }
                FA_TRACE(j < 8 && qr == 0, 192 + (j * 2 + g) * 4 + 1);
                // Every row decides for ITSELF whether to raise its maximum (a row's result must not depend on which other rows
                // share its warp: batched and single calls agree bit for bit); the warp runs the collective tcgen05.ld / st
                // when any of its rows does, the others multiply by one.
                const bool raise = mnew - m > RESCALE_THRESHOLD;                          // m = -inf on the first block: taken (false for NaN rows)
                if (__any_sync(0xffffffffu, raise)) {
                    if (j > 0) {
                        const float corr = raise ? ex2(m - mnew) : 1.f;
#pragma unroll 1
                        for (int hh = 0; hh < 2; hh++) {      // 32 columns at a time: the 128 scores stay in registers
                            float o[32];
                            tmem_ld32_nowait(ocol + hh * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; i++) o[i] *= corr;
                            tmem_st32_nowait(ocol + hh * 32, reinterpret_cast<const uint32_t*>(o));
                        }
                        l *= corr;
                    }
                    if (raise) m = mnew;
                }
                const float neg_m = (m == -INFINITY) ? 0.f : -m;                          // a row with every key so far masked (causal padding): P = 0
                // the exponentials of the two groups take turns on the MUFU pipe (named barriers 1 / 2, 256 = one group syncing +
                // the other arriving): left alone the groups fall into lockstep -- both in their MUFU-free phases (TMEM load, max,
                // hand-over) at once, then both halving each other's exponential rate
                if (pingpong) { if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory"); }
                FA_TRACE(j < 8 && qr == 0, 192 + (j * 2 + g) * 4 + 2);
                float l0 = 0.f, l1 = 0.f;
                // 32-key chunks of the last key block that lie wholly beyond the sequence get P = 0 without an exponential
                const int live_chunks = (masked && !p.causal) ? (min(BK, Tk - kbase) + 31) >> 5 : 4;
                auto exp_pass = [&](auto use_poly) {         // masked blocks (exact zeros wanted) take the MUFU for every element
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t pk[16];
                        if (c >= live_chunks) {
#pragma unroll
                            for (int i = 0; i < 16; i++) pk[i] = 0u;
                        } else
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float x0 = fmaf(s[c * 32 + i], p.sl2, neg_m), x1 = fmaf(s[c * 32 + i + 1], p.sl2, neg_m);
                            const float e0 = ex2(x0), e1 = decltype(use_poly)::value ? ex2_poly(x1) : ex2(x1);
                            l0 += e0; l1 += e1;
                            pk[i >> 1] = pack_bf16(e0, e1);
                        }
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                                     ::"r"(pcol + c * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]),
                                       "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15]) : "memory");
                    }
                };
                if (masked || !(p.dbg & 16)) exp_pass(std::false_type{}); else exp_pass(std::true_type{});
                if (pingpong && !(g == 1 && j == nkv - 1)) { if (g == 0) asm volatile("bar.arrive 2, 256;" ::: "memory"); else asm volatile("bar.arrive 1, 256;" ::: "memory"); }
                l += l0 + l1;
                FA_TRACE(j < 8 && qr == 0, 192 + (j * 2 + g) * 4 + 3);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(p_ready + 8 * g);
                FA_TRACE(j < 8 && qr == 0, 128 + (j * 2 + g) * 4 + 3);
            }
            mbar_wait(o_done + 8 * g, (nkv - 1) & 1);
            tc_fence_after();
            if (!dead) {                  // (dead warps: nothing to store, rows beyond the sequence are never consumed)
            float o[64];
            tmem_ld32_nowait(ocol, o);
            tmem_ld32_nowait(ocol + 32, o + 32);
            tmem_ld_wait();
            if (qi < p.T) {
                const float inv = 1.f / l;
                bf16* orow = p.o + (long)b * p.o_bs + (long)qi * p.ldo + h * D;
#pragma unroll
                for (int i = 0; i < 64; i += 8)
                    *reinterpret_cast<uint4*>(orow + i) = make_uint4(pack_bf16(o[i] * inv, o[i + 1] * inv), pack_bf16(o[i + 2] * inv, o[i + 3] * inv),
                                                                     pack_bf16(o[i + 4] * inv, o[i + 5] * inv), pack_bf16(o[i + 6] * inv, o[i + 7] * inv));
            }
            }
        }
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

bool make_map(CUtensorMap* tm, const bf16* ptr, long ld, long bs, int T, int H, int batch) {
    cuuint64_t dim[3] = {(cuuint64_t)H * D, (cuuint64_t)T, (cuuint64_t)batch};
    cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? bs : ld * (long)T) * 2};
    cuuint32_t box[3] = {D, 128, 1}, es[3] = {1, 1, 1};
    return g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)ptr, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

void attention_fa_init() {
    g_ok = false;
    if (const char* d = getenv("CBX_DISABLE_ATTN_FA")) { if (d[0] == '1') return; }
    if (const char* d = getenv("CBX_DISABLE_ATTN_TC")) { if (d[0] == '1') return; }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return;
    g_encode = (EncodeFn)fn;
    CBX_CHECK(cudaFuncSetAttribute(attn_fa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    g_ok = true;
}

// returns false when the problem does not fit this kernel (alignment): the caller falls back to the older kernels
bool launch_attention_fa(const AttnParams& p, cudaStream_t st) {
    if (!g_ok) return false;
    if (p.ldq % 8 || p.ldk % 8 || p.ldv % 8 || p.q_bs % 8 || p.k_bs % 8 || p.v_bs % 8 || p.ldo % 8 || p.o_bs % 8) return false;
    if (((uintptr_t)p.q & 15) || ((uintptr_t)p.k & 15) || ((uintptr_t)p.v & 15) || ((uintptr_t)p.o & 15)) return false;
    alignas(64) CUtensorMap tq, tk, tv;
    if (!make_map(&tq, p.q, p.ldq, p.q_bs, p.T, p.H, p.batch) || !make_map(&tk, p.k, p.ldk, p.k_bs, p.T, p.H, p.batch) ||
        !make_map(&tv, p.v, p.ldv, p.v_bs, p.T, p.H, p.batch)) return false;
    AttnFaArgs a;
    a.o = p.o; a.ldo = p.ldo; a.o_bs = p.o_bs; a.T = p.T; a.H = p.H; a.causal = p.causal; a.sl2 = p.scale * 1.4426950408889634f;
    for (int i = 0; i < 16; i++) a.kv_len[i] = p.kv_len[i];
    a.kv_div = p.kv_div;
    a.relbias = p.relbias; a.rb_ld = p.rb_ld; a.rb_hs = p.rb_hs; a.rb_bs = p.rb_bs;
    { static const int dbg = [] { const char* e = getenv("CBX_ATTN_FA_DBG"); return e ? atoi(e) : 0; }(); a.dbg = dbg; }
    ProfScope ps(PC_ATTN, 4.0 * p.T * p.T * D * p.H * p.batch * (p.causal ? 0.5 : 1.0), st);
    launch_pdl(attn_fa_kernel, dim3(cdiv(p.T, 2 * BQ), p.H, p.batch), dim3(THREADS), SMEM, st, tq, tk, tv, a);
    CBX_CHECK(cudaGetLastError());
    g_launches++;
    return true;
}

extern "C" long long cbx_attn_fa_launches(void) { return g_launches; }
// debug (CBX_ATTN_FA_DBG bit 2): %globaltimer stamps of CTA (0,0,0) of the last launch: [64 + (2j+g)*3 + {0,1,2}] MMA issuer before / after the wait for
// P_g(j) and after issuing PV_g(j), S_g(j+1); [128 + (2j+g)*4 + {0..3}] softmax row 0 of group g: before / after the wait for S_g(j), scores in registers, P handed over
extern "C" int cbx_attn_fa_trace(unsigned long long* out_h) { return cudaMemcpyFromSymbol(out_h, g_fa_trace, sizeof(unsigned long long) * 256) == cudaSuccess ? 0 : 1; }
