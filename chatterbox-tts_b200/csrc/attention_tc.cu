// tcgen05 / TMEM / TMA flash attention for head_dim 64 (sm_100a): the CFM transformer blocks' full attention
// (no bias, not causal), optionally ragged (keys >= kv_len masked).  The mma.sync kernel in attention.cu is capped by
// the legacy tensor path (~140 TFLOP/s); this one puts both products on the 5th-generation tensor cores:
//
//   CTA = 128 queries of one (batch, head); per 128-key block j
//     S_j  = Q K_j^T   tcgen05.mma M128 N128 K64 (Q, K: K-major SWIZZLE_128B tiles loaded by TMA)  -> TMEM S
//     softmax          8 warps, two threads per query row (64 score columns / 32 output columns each): tcgen05.ld S twice
//                      (max pass, exp pass), online rescale of the register-resident O row, P (bf16) written to shared
//                      memory in the K-major SWIZZLE_128B layout
//     PV_j = P_j V_j   tcgen05.mma M128 N64 K128 (V tile as loaded by TMA = MN-major B operand)    -> TMEM PV
//     O += PV_j        read back by the softmax threads during block j+1
//
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = softmax.  The per-block chain S -> softmax -> PV -> O is serial
// inside a CTA, so TWO CTAs share an SM (256 TMEM columns and ~100 KB of shared memory each: one S buffer, two K stages,
// one V stage): while one CTA's softmax warps wait for the tensor core, the other CTA's run.  Round 1 ran one CTA per SM
// (512 TMEM columns, double-buffered S) at 2.6 us per 128 x 128 score block against ~0.9 us of softmax issue time.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int D = 64, BQ = 128, BK = 128, THREADS = 320;
constexpr int TILE = 128 * 128;                 // bytes of a [128 rows x 64 bf16] SWIZZLE_128B tile
constexpr int SMEM = 6 * TILE + 1024 + 2048 + 256;   // Q, 2 K, V, P (two 64-key halves), alignment slack, row exchange, barriers
constexpr int TMEM_COLS = 256, S_COL = 0, PV_COL = 128;

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
// SWIZZLE_128B shared-memory matrix descriptor (sm_100 version 1): rows of 128 B, 8-row groups 1024 B apart.  The same
// encoding serves K-major tiles (rows = M/N index, 128 B = 64 K elements) and the MN-major V tile (rows = K index,
// 128 B = 64 N elements): in both the 8-row group stride is SBO = 1024 B and LBO is unused (one swizzle atom wide).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct AttnTcArgs {
    bf16* o; long ldo, o_bs;
    int T, H; float sl2;           // sl2 = scale * log2(e): scores are kept in the log2 domain
    int kv_len[16]; int kv_div;
};

__global__ void __launch_bounds__(THREADS, 2) attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                             const __grid_constant__ CUtensorMap tmV, const AttnTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base, sK = base + TILE, sV = base + 3 * TILE, sP = base + 4 * TILE, xch = base + 6 * TILE, bars = xch + 2048;
    const uint32_t q_full = bars, k_full = bars + 8, k_empty = bars + 24, v_full = bars + 40, v_empty = bars + 48, s_full = bars + 56,
                   p_ready = bars + 64, pv_done = bars + 72, tmem_slot = bars + 80;   // the K barriers come in pairs (stage 0, 1)
    float* xchf = reinterpret_cast<float*>(smem_raw + (xch - smem_u32(smem_raw)));     // [2 parities][2 halves][128 rows] row max / row sum exchange
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * BQ;
    const int klen_raw = p.kv_len[(b / p.kv_div) & 15];
    const int Tk = klen_raw > 0 ? klen_raw : p.T;
    pdl_launch_dependents();
    if (q0 >= Tk) return;                           // padded query tile of a ragged batch: never consumed
    const int nkv = (Tk + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; s++) { mbar_init(k_full + 8 * s, 1); mbar_init(k_empty + 8 * s, 1); }
        mbar_init(v_full, 1); mbar_init(v_empty, 1); mbar_init(s_full, 1);
        mbar_init(p_ready, 256); mbar_init(pv_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pdl_wait();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (elect_one()) {   // TMA producer: two K stages (S_{j+1} follows PV_j at once), one V stage
            mbar_expect_tx(q_full, TILE);
            tma_load_3d(sQ, &tmQ, q_full, h * D, q0, b);
            for (int j = 0; j < nkv; j++) {
                const int s = j & 1, ph = (j >> 1) & 1;
                mbar_wait(k_empty + 8 * s, ph ^ 1);
                mbar_expect_tx(k_full + 8 * s, TILE);
                tma_load_3d(sK + s * TILE, &tmK, k_full + 8 * s, h * D, j * BK, b);
                mbar_wait(v_empty, (j & 1) ^ 1);
                mbar_expect_tx(v_full, TILE);
                tma_load_3d(sV, &tmV, v_full, h * D, j * BK, b);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {   // MMA issuer
            // instruction descriptors: D=f32, A=B=bf16; S: N=128, both K-major; PV: N=64, B (= V) MN-major
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BK >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
            const uint64_t qd = umma_desc(sQ), pd = umma_desc(sP);
            auto issue_s = [&](int j) {   // S_j = Q K_j^T (one S buffer: the softmax of block j-1 has finished reading it)
                const int s = j & 1;
                mbar_wait(k_full + 8 * s, (j >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t kd = umma_desc(sK + s * TILE);
#pragma unroll
                for (int k = 0; k < D / 16; k++) umma_bf16(tmem_base + S_COL, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);   // +32 B per K=16
                umma_commit(s_full);
                umma_commit(k_empty + 8 * s);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nkv; j++) {
                mbar_wait(p_ready, j & 1);      // P_j is in shared memory and S_j has been read out of TMEM
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (j + 1 < nkv) issue_s(j + 1);   // first: the next block's softmax starts with its scores
                mbar_wait(v_full, j & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t vd = umma_desc(sV);
#pragma unroll
                for (int k = 0; k < BK / 16; k++)   // A: P half k/4, +32 B per 16 keys; B: V rows 16k.. = +2048 B
                    umma_bf16(tmem_base + PV_COL, pd + (uint64_t)((k >> 2) * (TILE >> 4) + 2 * (k & 3)), vd + (uint64_t)(k * (2048 >> 4)), idesc_pv, k != 0);
                umma_commit(pv_done);
                umma_commit(v_empty);
            }
        }
    } else {               // softmax warps 2..9: TMEM lane quarter = warp % 4; score-column / output-column half = (warp - 2) / 4
        const int half = (warp - 2) >> 2;
        const int qr = (warp & 3) * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        float m = -INFINITY, l = 0.f;
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; i++) o[i] = 0.f;
        for (int j = 0; j < nkv; j++) {
            mbar_wait(s_full, j & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t scol = trow + S_COL + half * 64;
            const int kbase = j * BK + half * 64;
            const bool tail = kbase + 64 > Tk;
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                uint32_t v[32];
                tmem_ld32(scol + c * 32, v);
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const float s = __uint_as_float(v[i]);
                    if (!tail || kbase + c * 32 + i < Tk) mx = fmaxf(mx, s);
                }
            }
            float* xj = xchf + (j & 1) * 256;                 // parity-buffered: the next block's write cannot overtake a partner's read
            xj[half * 128 + qr] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mx = fmaxf(mx, xj[(half ^ 1) * 128 + qr]);
            const float mnew = fmaxf(m, mx * p.sl2);          // sl2 > 0: scaling commutes with the max
            const float corr = exp2f(m - mnew);               // m = -inf on the first block -> 0
            if (j > 0) {                                      // O += P_{j-1} V_{j-1} (this thread's 32 output columns)
                mbar_wait(pv_done, (j - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t v[32];
                tmem_ld32(trow + PV_COL + half * 32, v);
#pragma unroll
                for (int i = 0; i < 32; i++) o[i] += __uint_as_float(v[i]);
            }
#pragma unroll
            for (int i = 0; i < 32; i++) o[i] *= corr;
            l *= corr;
            m = mnew;
            // exp pass: this thread's 64 keys of P_j (bf16) -> shared memory half-tile `half`, K-major SWIZZLE_128B
            const uint32_t prow = sP + half * TILE + qr * 128;
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                uint32_t v[32];
                tmem_ld32(scol + c * 32, v);
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float e0 = exp2f(__uint_as_float(v[i]) * p.sl2 - mnew), e1 = exp2f(__uint_as_float(v[i + 1]) * p.sl2 - mnew);
                    if (tail) {
                        if (kbase + c * 32 + i >= Tk) e0 = 0.f;
                        if (kbase + c * 32 + i + 1 >= Tk) e1 = 0.f;
                    }
                    l += e0 + e1;
                    pk[i >> 1] = pack_bf16(e0, e1);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int chunk = c * 4 + q;              // 16-byte chunk (8 keys) inside the 64-key half
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((chunk ^ (qr & 7)) << 4)),
                                 "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(p_ready);
        }
        mbar_wait(pv_done, (nkv - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            uint32_t v[32];
            tmem_ld32(trow + PV_COL + half * 32, v);
#pragma unroll
            for (int i = 0; i < 32; i++) o[i] += __uint_as_float(v[i]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");       // everyone is past the last read of the max exchange
        xchf[half * 128 + qr] = l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        l += xchf[(half ^ 1) * 128 + qr];
        const int qi = q0 + qr;
        if (qi < p.T) {
            const float inv = 1.f / l;
            bf16* orow = p.o + (long)b * p.o_bs + (long)qi * p.ldo + h * D + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8)
                *reinterpret_cast<uint4*>(orow + i) = make_uint4(pack_bf16(o[i] * inv, o[i + 1] * inv), pack_bf16(o[i + 2] * inv, o[i + 3] * inv),
                                                                 pack_bf16(o[i + 4] * inv, o[i + 5] * inv), pack_bf16(o[i + 6] * inv, o[i + 7] * inv));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
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

void attention_tc_init() {
    g_ok = false;
    if (const char* d = getenv("CBX_DISABLE_ATTN_TC")) { if (d[0] == '1') return; }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return;
    g_encode = (EncodeFn)fn;
    CBX_CHECK(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    g_ok = true;
}

// returns false when the problem does not fit this kernel (bias / causal / alignment): the caller uses the mma.sync kernel
bool launch_attention_tc(const AttnParams& p, cudaStream_t st) {
    if (!g_ok || p.causal || p.relbias) return false;
    if (p.ldq % 8 || p.ldk % 8 || p.ldv % 8 || p.q_bs % 8 || p.k_bs % 8 || p.v_bs % 8 || p.ldo % 8 || p.o_bs % 8) return false;
    if (((uintptr_t)p.q & 15) || ((uintptr_t)p.k & 15) || ((uintptr_t)p.v & 15) || ((uintptr_t)p.o & 15)) return false;
    alignas(64) CUtensorMap tq, tk, tv;
    if (!make_map(&tq, p.q, p.ldq, p.q_bs, p.T, p.H, p.batch) || !make_map(&tk, p.k, p.ldk, p.k_bs, p.T, p.H, p.batch) ||
        !make_map(&tv, p.v, p.ldv, p.v_bs, p.T, p.H, p.batch)) return false;
    AttnTcArgs a;
    a.o = p.o; a.ldo = p.ldo; a.o_bs = p.o_bs; a.T = p.T; a.H = p.H; a.sl2 = p.scale * 1.4426950408889634f;
    for (int i = 0; i < 16; i++) a.kv_len[i] = p.kv_len[i];
    a.kv_div = p.kv_div;
    ProfScope ps(PC_ATTN, 4.0 * p.T * p.T * D * p.H * p.batch, st);
    launch_pdl(attn_tc_kernel, dim3(cdiv(p.T, BQ), p.H, p.batch), dim3(THREADS), SMEM, st, tq, tk, tv, a);
    CBX_CHECK(cudaGetLastError());
    g_launches++;
    return true;
}

extern "C" long long cbx_attn_fa_launches(void);
// both tcgen05 attention kernels (this one and attention_fa.cu)
extern "C" long long cbx_attn_tc_launches(void) { return g_launches + cbx_attn_fa_launches(); }
