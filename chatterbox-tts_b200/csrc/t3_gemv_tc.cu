// T3 decode projections for BATCHED rows on the 5th-generation tensor cores (sm_100a): y[r][f] = sum_k W[f][k] x[r][k] for the
// 9..32 rows (2 x streams) of a decode step, "swap-AB": the WEIGHT tile is the M operand.
//
//   CTA (mt, ks) of a cluster of `ksplit` CTAs: features [128 mt, +128), K slice [ks Kc, +Kc)
//     A = W[128 x Kc]   row-major bf16 weights, TMA -> SWIZZLE_128B shared memory; ALL k-blocks of the slice are requested
//                       BEFORE the programmatic-dependent-launch wait (weights do not depend on the previous kernel), so the
//                       weight stream of kernel i+1 runs under the tail of kernel i
//     B = x[32 x Kc]    the step's bf16 activation rows (row-mapped, <= 32), staged by all threads into the K-major
//                       SWIZZLE_128B layout after the wait
//     D = TMEM [128 lanes x 32 columns] fp32, tcgen05.mma M128 N32 K16
//   split-K partial sums meet in the cluster leader (ks = 0) through distributed shared memory; the leader runs the same three
//   epilogues as the mma.sync GEMV (t3_kernels.cu): STORE (x RMS scale), RESID (+ residual, bf16 hand-over x next gain, per-strip
//   sums of squares), GLU (SiLU(gate) * up from adjacent weight rows).  One thread = one feature: stores of a row are
//   coalesced across the warp.
//
// The GEMV streams fragment-ordered weights through mma.sync with the rows as the 8-wide N operand: at 16-32 rows every CTA
// re-stages all activation rows and the instruction stream, not HBM, bounds it (7-10 us per projection).  Here a projection is
// one wave of <= 128 CTAs with 32-128 KB of weights each.
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include "t3_kernels.cuh"

namespace {

constexpr int TM = 128, TK = 64, RP = 32, THREADS = 192, MAX_KB = 8;
constexpr int A_BYTES = TM * TK * 2, B_BYTES = RP * TK * 2;
constexpr int SMEM_FIXED = 1024 + 512;      // alignment slack + barriers / TMEM slot / row scales
// Every launch asks for the full MAX_KB stages, i.e. ONE CTA per SM.  Sizing the request to the K slice (40-176 KB) lets the next
// kernel's CTAs move in beside the running ones under PDL -- measured slower (16 rows 1.27 vs 1.17 ms, 32 rows 1.64 vs 1.53 ms per step).

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
    __trap();
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
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_dsmem(uint32_t local_addr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    return remote;
}
__device__ __forceinline__ float4 ld_dsmem4(uint32_t remote) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
    return v;
}

struct GemvTcArgs {
    int N, K, kb_per_cta, ksplit, rows, epi, n_strips;
    const int* row_map;
    const bf16* xb; long ldxb;
    const float* ss_in; int n_ss; float eps;
    float* out; long ld_out;
    bf16* out_b; long ld_out_b;
    const float* next_gain; float* ss_out;
};

__global__ void __launch_bounds__(THREADS, 1) gemv_tc_kernel(const __grid_constant__ CUtensorMap tmW, const GemvTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int nkb = p.kb_per_cta;
    const uint32_t sA = base, sB = base + nkb * A_BYTES, bars = sB + nkb * B_BYTES;
    const uint32_t full0 = bars, acc_full = bars + 8 * MAX_KB, tmem_slot = acc_full + 8;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    float* scl = reinterpret_cast<float*>(gen + nkb * (A_BYTES + B_BYTES) + 8 * MAX_KB + 16);        // [32] RMS scale per row
    int* rmap = reinterpret_cast<int*>(scl + RP);                                                      // [32]
    float* part = reinterpret_cast<float*>(gen);                 // [128][32] split-K partials (chunk-swizzled): over the weight stages (free once the MMAs have completed)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x, ks = blockIdx.y;

    pdl_launch_dependents();
    if (warp == 0 && elect_one()) {
        for (int i = 0; i < nkb; i++) mbar_init(full0 + 8 * i, 1);
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW) : "memory");
        // the whole weight slice of this CTA, before the dependency wait
        for (int i = 0; i < nkb; i++) {
            mbar_expect_tx(full0 + 8 * i, A_BYTES);
            tma_load_2d(sA + i * A_BYTES, &tmW, full0 + 8 * i, (ks * nkb + i) * TK, mt * TM);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x < RP) rmap[threadIdx.x] = threadIdx.x < p.rows ? p.row_map[threadIdx.x] : 0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pdl_wait();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // ---- activation rows of this K slice -> K-major SWIZZLE_128B tiles [32 rows x 64] per k-block (16-byte chunks)
    {
        const int cpr = nkb * 8;                                   // chunks per row in the slice
        const bf16* xb = p.xb + (long)ks * nkb * TK;
        for (int i0 = 0; i0 < RP * cpr; i0 += THREADS * 4) {        // four 16-byte loads in flight per thread
            uint4 v4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * THREADS + threadIdx.x, r = i / cpr, ch = i - r * cpr;
                v4[u] = (i < RP * cpr && r < p.rows) ? __ldg(reinterpret_cast<const uint4*>(xb + (long)rmap[r] * p.ldxb + ch * 8)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * THREADS + threadIdx.x, r = i / cpr, ch = i - r * cpr, kb = ch >> 3, c8 = ch & 7;
                if (i < RP * cpr) {
                    const uint32_t dst = sB + kb * B_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v4[u].x), "r"(v4[u].y), "r"(v4[u].z), "r"(v4[u].w) : "memory");
                }
            }
        }
        if (threadIdx.x < RP) {
            float sc = 1.f;
            if (p.ss_in && threadIdx.x < p.rows) {
                const float4* sp = reinterpret_cast<const float4*>(p.ss_in + (long)rmap[threadIdx.x] * p.n_ss);
                float ss = 0.f;
                for (int i = 0; i < p.n_ss / 4; i++) { const float4 t = sp[i]; ss += (t.x + t.y) + (t.z + t.w); }
                sc = rsqrtf(ss / p.K + p.eps);
            }
            scl[threadIdx.x] = sc;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }

    if (warp == 1 && elect_one()) {       // ---------------------------------------------------------------- MMA issuer
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        for (int i = 0; i < nkb; i++) {
            mbar_wait(full0 + 8 * i, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t ad = umma_desc(sA + i * A_BYTES), bd = umma_desc(sB + i * B_BYTES);
#pragma unroll
            for (int k = 0; k < TK / 16; k++) umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (i | k) != 0);
        }
        umma_commit(acc_full);
    }
    // ---- accumulator -> registers (warps 2..5: TMEM lane quarter = warp % 4; one thread = one feature)
    float v[RP];
    const int q = warp & 3, f = q * 32 + lane;                 // feature inside the tile
    const bool epi_warp = warp >= 2;
    if (epi_warp) {
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t* u = reinterpret_cast<uint32_t*>(v);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                       "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                       "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                       "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                     : "r"(tmem_base + ((uint32_t)(q * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    // ---- split-K reduce-scatter over the cluster: every CTA parks its partial tile in its own shared memory, then finalises
    // 128 / ksplit of the features: one thread = one feature x four rows, its ksplit partial float4s fetched with ALL loads in
    // flight (distributed shared memory), summed in rank order (deterministic), then the epilogue.  Every CTA of the cluster
    // stores outputs: no idle peers, one DSMEM round trip.
    if (epi_warp) {
#pragma unroll
        for (int j = 0; j < 8; j++)       // row of feature f: eight 16-byte chunks, chunk index xor (f & 7) against bank conflicts
            *reinterpret_cast<float4*>(part + f * RP + ((j ^ (f & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncwarp();
    if (p.ksplit > 1) { cluster_arrive(); cluster_wait(); } else __syncthreads();
    if (epi_warp) {
        const int nf = TM / p.ksplit, te = threadIdx.x - 64;
        for (int item = te; item < nf * 8; item += 128) {
            const int j = item / nf, ff = ks * nf + item % nf;            // row chunk, feature inside the tile
            const uint32_t la = smem_u32(part) + (uint32_t)(ff * RP + ((j ^ (ff & 7)) << 2)) * 4;
            float4 t[8];
#pragma unroll
            for (int c = 0; c < 8; c++) if (c < p.ksplit) t[c] = ld_dsmem4(map_dsmem(la, (uint32_t)c));
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 8; c++) if (c < p.ksplit) { s4[0] += t[c].x; s4[1] += t[c].y; s4[2] += t[c].z; s4[3] += t[c].w; }
            const int col = mt * TM + ff;
            const float ng = (p.epi == GEMV_RESID && p.next_gain && col < p.N) ? p.next_gain[col] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = 4 * j + i;
                const bool live = r < p.rows && col < p.N;
                const float val = s4[i] * scl[r];
                if (p.epi == GEMV_GLU) {          // weight rows (2j, 2j+1) = (gate_j, up_j): even lanes combine with their neighbour
                    const float up = __shfl_down_sync(0xffffffffu, val, 1);
                    if (live && !(ff & 1)) p.out_b[(long)rmap[r] * p.ld_out_b + (col >> 1)] = __float2bfloat16(val / (1.f + expf(-val)) * up);
                } else if (p.epi == GEMV_RESID) {
                    float nv = 0.f;
                    if (live) {
                        float* o = p.out + (long)rmap[r] * p.ld_out + col;
                        nv = *o + val;
                        *o = nv;
                        if (p.out_b) p.out_b[(long)rmap[r] * p.ld_out_b + col] = __float2bfloat16(nv * ng);
                    }
                    if (p.ss_out) {               // sum of squares of the strip's 16 new values, fixed order (xor tree inside the 16-lane group)
                        float ss = nv * nv;
                        ss += __shfl_xor_sync(0xffffffffu, ss, 8); ss += __shfl_xor_sync(0xffffffffu, ss, 4);
                        ss += __shfl_xor_sync(0xffffffffu, ss, 2); ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                        if (live && !(ff & 15)) p.ss_out[(long)rmap[r] * p.n_strips + (col >> 4)] = ss;
                    }
                } else if (live) p.out[(long)rmap[r] * p.ld_out + col] = val;
            }
        }
    }
    if (p.ksplit > 1) {                   // nobody leaves while a peer may still read its partials
        __syncwarp();
        cluster_arrive();
        cluster_wait();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
bool g_ok = false;
long g_launches = 0;

}  // namespace

void gemv_tc_init() {
    g_ok = false;
    if (const char* d = getenv("CBX_DISABLE_T3_TC")) { if (d[0] == '1') return; }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return;
    g_encode = (EncodeFn)fn;
    CBX_CHECK(cudaFuncSetAttribute(gemv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_KB * (A_BYTES + B_BYTES) + SMEM_FIXED));
    g_ok = true;
}
bool gemv_tc_available() { return g_ok; }

// TMA descriptor of a row-major bf16 weight [N][K] (static for the life of the engine): boxes of [128 rows x 64 columns]
void gemv_tc_weight_map(unsigned char (&map)[128], const bf16* w, int N, int K) {
    CBX_REQUIRE(g_ok, "gemv_tc: not initialised");
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    cuuint64_t dim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t str[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {TK, TM}, es[2] = {1, 1};
    CBX_REQUIRE(g_encode(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS,
                "gemv_tc: cuTensorMapEncodeTiled failed");
}

// p: the GEMV parameter block (bf16 input rows required); map: the weight's TMA descriptor.  Returns false when the shape
// does not fit (the caller launches the mma.sync GEMV instead).
bool launch_gemv_tc(const GemvParams& p, const unsigned char (&map)[128], cudaStream_t st) {
    if (!g_ok || !p.xb || p.rows < 1 || p.rows > RP || p.K % TK) return false;
    if (p.epi == GEMV_GLU && !p.out_b) return false;
    const int kb = p.K / TK, mtiles = cdiv(p.N, TM);
    int ksplit = 1;
    while (ksplit < 8 && kb % (ksplit * 2) == 0 && mtiles * ksplit * 2 <= 160) ksplit *= 2;      // fill the SMs, keep slices whole
    while (kb / ksplit > MAX_KB) { if (ksplit >= 8) return false; ksplit *= 2; }
    if (kb % ksplit) return false;
    GemvTcArgs a;
    a.N = p.N; a.K = p.K; a.kb_per_cta = kb / ksplit; a.ksplit = ksplit; a.rows = p.rows; a.epi = p.epi; a.n_strips = p.n_strips;
    a.row_map = p.row_map; a.xb = p.xb; a.ldxb = p.ldxb; a.ss_in = p.ss_in; a.n_ss = p.n_ss; a.eps = p.eps;
    a.out = p.out; a.ld_out = p.ld_out; a.out_b = p.out_b; a.ld_out_b = p.ld_out_b; a.next_gain = p.next_gain; a.ss_out = p.ss_out;
    alignas(64) CUtensorMap tm;
    memcpy(&tm, map, 128);
    ProfScope ps(PC_GEMV, (double)p.N * p.K * 2 + (double)p.rows * (p.K + p.N) * 4, st);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(mtiles, ksplit); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = (size_t)MAX_KB * (A_BYTES + B_BYTES) + SMEM_FIXED; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = ksplit; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    CBX_CHECK(cudaLaunchKernelEx(&cfg, gemv_tc_kernel, tm, a));
    CBX_CHECK(cudaGetLastError());
    g_launches++;
    return true;
}

extern "C" long long cbx_t3_tc_launches(void) { return g_launches; }
