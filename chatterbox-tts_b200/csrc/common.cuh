// Shared device helpers for the Chatterbox B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

typedef __nv_bfloat16 bf16;

#define CBX_CHECK(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            throw std::runtime_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + \
                                     " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

#define CBX_REQUIRE(cond, msg)                                                             \
    do {                                                                                  \
        if (!(cond)) throw std::runtime_error(std::string("cbx: ") + (msg) + " [" #cond "] at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- activations
enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2, ACT_MISH = 3, ACT_LRELU = 4, ACT_ELU = 5, ACT_SNAKE = 6 };

// GELU of a bf16-bound activation: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) with tanh.approx (one MUFU).  The reference
// applies the erf form; the two differ by at most 4.7e-4 absolute, a tenth of the bf16 rounding the value gets right after, and
// libdevice's erff costs ~30 issue slots per element in the epilogue of the widest GEMM of a CFM block.  (The fp32 conditioning
// encoders keep the exact erf form, cond.cu.)
__device__ __forceinline__ float gelu_tanh_form(float x) {
    const float u = x * fmaf(x * x, 0.0356774081f, 0.7978845608f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

// mish(x) = x tanh(softplus(x)) without log1p / tanh: with e = exp(x), tanh(log(1 + e)) = ((1 + e)^2 - 1) / ((1 + e)^2 + 1) = n / (n + 2),
// n = e (e + 2).  One exp and one division on the fast paths; fp32-exact to rounding (x > 20: tanh(softplus) = 1 in fp32).
__device__ __forceinline__ float mish_fast(float v) {
    const float e = __expf(fminf(v, 20.f)), n = e * (e + 2.f);
    return v > 20.f ? v : v * __fdividef(n, n + 2.f);
}

__device__ __forceinline__ float act_apply(int act, float v, float param) {
    switch (act) {
        case ACT_GELU: return gelu_tanh_form(v);
        case ACT_SILU: return __fdividef(v, 1.f + __expf(-v));
        case ACT_MISH: return mish_fast(v);
        case ACT_LRELU: return v > 0.f ? v : v * param;
        case ACT_ELU: return v > 0.f ? v : expm1f(v);
        case ACT_SNAKE: {
            float s = __sinf(v * param);
            return v + s * s / (param + 1e-9f);
        }
        default: return v;
    }
}

template <int ACT>
__device__ __forceinline__ void act_apply8(float (&v)[8], const float (&alpha)[8] /*per-column parameter (snake alpha / lrelu slope)*/) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if constexpr (ACT == ACT_GELU) v[i] = gelu_tanh_form(v[i]);
        else if constexpr (ACT == ACT_SILU) v[i] = __fdividef(v[i], 1.f + __expf(-v[i]));
        else if constexpr (ACT == ACT_MISH) v[i] = mish_fast(v[i]);
        else if constexpr (ACT == ACT_LRELU) v[i] = v[i] > 0.f ? v[i] : v[i] * alpha[i];
        else if constexpr (ACT == ACT_ELU) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
        else if constexpr (ACT == ACT_SNAKE) { float a = alpha[i]; float s = __sinf(v[i] * a); v[i] = v[i] + s * s / (a + 1e-9f); }
    }
}

// activation pairs (main, second output) that the model actually uses; every GEMM kernel is instantiated per pair so that
// only one activation's code is resident (a runtime switch over all of them overflowed the instruction cache)
#define CBX_FOR_ACT_PAIRS(X) \
    X(ACT_NONE, ACT_NONE) X(ACT_GELU, ACT_NONE) X(ACT_SILU, ACT_NONE) X(ACT_MISH, ACT_NONE) X(ACT_LRELU, ACT_NONE) \
    X(ACT_ELU, ACT_NONE) X(ACT_SNAKE, ACT_NONE) X(ACT_NONE, ACT_SNAKE) X(ACT_NONE, ACT_LRELU)

// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- async copy / mma primitives
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// One lane of a converged warp (elect.sync).  Use this, not `lane == 0`, to pick the thread that issues tcgen05.mma / TMA
// instructions: their operands live in uniform registers, and under a lane test the compiler cannot prove the region is
// single-threaded, so it wraps EVERY such instruction in an election loop with one R2UR broadcast per operand (~80 cycles per
// tcgen05.mma, measured in attention_fa.cu, against 32-64 cycles of tensor-pipe time).  After elect.sync the instructions are
// issued back to back.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- Philox4x32-10 (counter-based RNG)
struct Philox {
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // counter (c0..c3), key (seed lo/hi) -> 4 x u32
    __host__ __device__ static inline void gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) {
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
        for (int i = 0; i < 10; i++) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};
// u32 -> float in (0,1]
__host__ __device__ static inline float u32_to_unit(uint32_t x) { return ((x >> 8) + 1) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel calls pdl_prologue() (or its two halves) and is launched through launch_pdl(): the next kernel's CTAs are
// scheduled as soon as all CTAs of the current one have started, run their prologue (barrier init, TMEM allocation,
// weight prefetch) and then block in griddepcontrol.wait until the current kernel has completed and flushed its writes.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CBX_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// ---------------------------------------------------------------- per-launch profiler (bench.py roofline pass)
enum ProfClass { PC_GEMM = 0, PC_ATTN = 1, PC_GEMV = 2, PC_DECODE_ATTN = 3, PC_SAMPLER = 4, PC_NORM = 5, PC_ELEMWISE = 6, PC_HIFT_MISC = 7, PC_COUNT = 8 };
bool prof_enabled();   // per-launch bracketing is on for this thread
bool prof_active();    // a profiling pass is running (even if this thread has suspended per-launch bracketing)
void prof_suspend(int delta);   // +1 / -1: units that are timed as a whole (a graph-replayed T3 decode step)
void prof_begin_launch(int cls, double work, cudaStream_t st);   // work: FLOPs (GEMM/ATTN) or bytes (everything else)
void prof_end_launch(cudaStream_t st);
struct ProfScope {
    cudaStream_t st; bool on;
    ProfScope(int cls, double work, cudaStream_t s) : st(s), on(prof_enabled()) { if (on) prof_begin_launch(cls, work, st); }
    ~ProfScope() { if (on) prof_end_launch(st); }
};

// ---------------------------------------------------------------- GEMM descriptor (implemented in gemm.cu)
// C[b][m][n] = epilogue( sum_kk A(b,m,kk) * W(b,n,kk) )
//   A(b,m,kk) = A[b*a_bs + m*lda + (kk/kc)*tap_stride + kk%kc]   (bf16)  -> conv1d over time-major channels-last activations
//   W(b,n,kk) = W[b*w_bs + n*ldw + kk]                            (bf16)
struct GemmParams {
    const bf16* A = nullptr; long lda = 0; int kc = 0; long tap_stride = 0; long a_bs = 0;
    const bf16* W = nullptr; long ldw = 0; long w_bs = 0;
    int M = 0, N = 0, K = 0, batch = 1;
    const float* bias = nullptr;            // [N]
    const float* bias2 = nullptr; long bias2_bs = 0;   // per-batch [N] added before act (e.g. time embedding)
    int act = ACT_NONE; float act_param = 0.f; const float* act_alpha = nullptr;  // per-column alpha for snake
    int glu = 0;                            // columns (2j,2j+1) = (gate,up): out[j] = silu(gate)*up ; output width N/2
    const float* res = nullptr; long ldr = 0; long r_bs = 0;   // fp32 residual added after act
    float out_scale = 1.f; int accumulate = 0;                   // outF = (accumulate? outF : 0) + scale*v
    float* outF = nullptr; bf16* outB = nullptr; long ldc = 0; long c_bs = 0;
    bf16* outB2 = nullptr; int act2 = ACT_NONE; float act2_param = 0.f; const float* act2_alpha = nullptr; long ldc2 = 0; long c2_bs = 0;
    // LayerNorm of the finished row fused into the epilogue (tcgen05 kernel only; N must be 1..4 tiles wide): outF = acc + bias +
    // res (fp32), outB2 = LayerNorm(outF row) * ln_gamma + ln_beta (bf16)
    const float* ln_gamma = nullptr; const float* ln_beta = nullptr; float ln_eps = 1e-5f;
    // transposed-conv scatter mask: column n belongs to phase n/ct_cout; output time = m*ct_u + phase - ct_pad must be in [0, ct_len)
    int ct_u = 0, ct_cout = 0, ct_pad = 0, ct_len = 0;
    // ragged batches of slabs (batch = slab index): row tiles that start at or beyond row_len[b / row_div] lie in the slab's padding
    // and are skipped by the tcgen05 kernel (their rows are never consumed).  row_div == 0: off
    int row_len[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; int row_div = 0;
};
void launch_gemm(const GemmParams& p, cudaStream_t st);
void gemm_init();
void gemm_tc_init();
bool launch_gemm_tc(const GemmParams& p, cudaStream_t st);   // per-device kernel attributes (call once after cudaSetDevice)

// flash attention (attention.cu): o[b][t][h*64+d] = softmax_j(scale*(q.k + bias)) v
struct AttnParams {
    const bf16 *q = nullptr, *k = nullptr, *v = nullptr; long ldq = 0, ldk = 0, ldv = 0; long q_bs = 0, k_bs = 0, v_bs = 0;
    bf16* o = nullptr; long ldo = 0; long o_bs = 0;
    int T = 0, H = 0, batch = 1; int causal = 0; float scale = 0.125f;
    const float* relbias = nullptr; long rb_ld = 0; long rb_hs = 0; long rb_bs = 0;  // bias[b*rb_bs + h*rb_hs + i*rb_ld + (T-1-i+j)]
    // ragged batches (sequences right-padded to T): keys >= kv_len[b / kv_div] are masked and query tiles beyond it are
    // skipped; 0 = the full T
    int kv_len[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; int kv_div = 1;
};
void launch_attention(const AttnParams& p, cudaStream_t st);
void attention_init();
void attention_tc_init();
void attention_fa_init();
bool launch_attention_fa(const AttnParams& p, cudaStream_t st);   // false: not applicable (additive bias / alignment)
bool launch_attention_tc(const AttnParams& p, cudaStream_t st);   // false: not applicable (bias / causal / alignment)
