// Fused GEMM epilogue shared by the mma.sync (gemm.cu) and tcgen05 (gemm_tc.cu) kernels.
// One call handles 8 consecutive output columns of one row: all global loads (bias, per-batch bias,
// activation alphas, residual, accumulate target) are issued first as independent 16-byte loads, then the
// math, then 16-byte stores -- the earlier per-pair load/store chain serialised on memory latency.
#pragma once
#include "common.cuh"

__device__ __forceinline__ bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

__device__ __forceinline__ void ld8(const float* __restrict__ ptr, int nv, float (&o)[8]) {
    if (nv == 8 && al16(ptr)) {
        float4 a = *reinterpret_cast<const float4*>(ptr), b = *reinterpret_cast<const float4*>(ptr + 4);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = i < nv ? ptr[i] : 0.f;
    }
}
__device__ __forceinline__ void st8(float* ptr, int nv, const float (&v)[8]) {
    if (nv == 8 && al16(ptr)) {
        *reinterpret_cast<float4*>(ptr) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(ptr + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) if (i < nv) ptr[i] = v[i];
    }
}
__device__ __forceinline__ void st8b(bf16* ptr, int nv, const float (&v)[8]) {
    if (nv == 8 && al16(ptr)) {
        uint4 u = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        *reinterpret_cast<uint4*>(ptr) = u;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) if (i < nv) ptr[i] = __float2bfloat16(v[i]);
    }
}

// math + stores for 8 columns whose per-column / per-element operands are already in registers
template <int ACT, int ACT2>
__device__ __forceinline__ void epilogue_math_store8(const GemmParams& p, int b, int m, int n, int nv, float (&v)[8], const float (&t0)[8],
                                                     const float (&t1)[8], const float (&a1)[8], const float (&a2)[8], const float (&r)[8],
                                                     const float (&old)[8]) {
    const long o = (long)b * p.c_bs + (long)m * p.ldc + n;
    // ---- math
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[i] = v[i] + t0[i] + t1[i];   // zero-filled when the bias vectors are absent
    }
    if (p.glu) {   // columns (2j, 2j+1) = (gate, up)
        const long og = (long)b * p.c_bs + (long)m * p.ldc + (n >> 1);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (2 * j + 1 < nv) {
                float gte = v[2 * j], g = gte / (1.f + expf(-gte)) * v[2 * j + 1];   // exact division: T3 logits parity
                if (p.outB) p.outB[og + j] = __float2bfloat16(g);
                if (p.outF) p.outF[og + j] = g;
            }
        }
        return;
    }
    act_apply8<ACT>(v, a1);
    if (p.res) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] += r[i];
    }
    if (p.out_scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] *= p.out_scale;
    }
    if (p.outF && p.accumulate) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] += old[i];
    }
    // ---- scatter
    if (p.outF) st8(p.outF + o, nv, v);
    if (p.outB) st8b(p.outB + o, nv, v);
    if (p.outB2) {
        float u[8];
#pragma unroll
        for (int i = 0; i < 8; i++) u[i] = v[i];
        act_apply8<ACT2>(u, a2);
        st8b(p.outB2 + (long)b * p.c2_bs + (long)m * p.ldc2 + n, nv, u);
    }
}

// v: 8 accumulators of row m, columns [n, n+8); n % 8 == 0
template <int ACT, int ACT2>
__device__ __forceinline__ void epilogue_chunk8(const GemmParams& p, int b, int m, int n, float (&v)[8]) {
    if (m >= p.M || n >= p.N) return;
    const int nv = min(8, p.N - n);
    if (p.ct_u) {   // transposed-conv scatter: the whole chunk lies in one output phase (ct_cout % 8 == 0)
        const int t = m * p.ct_u + n / p.ct_cout - p.ct_pad;
        if (t < 0 || t >= p.ct_len) return;
    }
    // ---- gather
    float t0[8], t1[8], r[8], old[8], a1[8], a2[8];
    if (p.bias) ld8(p.bias + n, nv, t0);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) t0[i] = 0.f;
    }
    if (p.bias2) ld8(p.bias2 + (long)b * p.bias2_bs + n, nv, t1);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) t1[i] = 0.f;
    }
    if (p.act && p.act_alpha) ld8(p.act_alpha + n, nv, a1);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) a1[i] = p.act_param;
    }
    if (p.outB2 && p.act2_alpha) ld8(p.act2_alpha + n, nv, a2);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) a2[i] = p.act2_param;
    }
    const long o = (long)b * p.c_bs + (long)m * p.ldc + n;
    if (p.res) ld8(p.res + (long)b * p.r_bs + (long)m * p.ldr + n, nv, r);
    if (p.outF && p.accumulate) ld8(p.outF + o, nv, old);
    epilogue_math_store8<ACT, ACT2>(p, b, m, n, nv, v, t0, t1, a1, a2, r, old);
}

// Per-column operands of one thread's fixed 8-column strip (loaded once, reused for every row it handles).
struct ColOps { float t0[8], t1[8], a1[8], a2[8]; int n, nv; };

__device__ __forceinline__ void load_colops(const GemmParams& p, int b, int n, ColOps& c) {
    c.n = n; c.nv = min(8, p.N - n);
    const int nv = c.nv > 0 ? c.nv : 0;
    if (p.bias) ld8(p.bias + n, nv, c.t0); else { for (int i = 0; i < 8; i++) c.t0[i] = 0.f; }
    if (p.bias2) ld8(p.bias2 + (long)b * p.bias2_bs + n, nv, c.t1); else { for (int i = 0; i < 8; i++) c.t1[i] = 0.f; }
    if (p.act_alpha) ld8(p.act_alpha + n, nv, c.a1); else { for (int i = 0; i < 8; i++) c.a1[i] = p.act_param; }
    if (p.act2_alpha) ld8(p.act2_alpha + n, nv, c.a2); else { for (int i = 0; i < 8; i++) c.a2[i] = p.act2_param; }
}

// rows r0, r0+rstep, ... < rows_in_tile of a staged fp32 tile (row stride ldt) -> fused epilogue, two rows in flight
template <int ACT, int ACT2>
__device__ __forceinline__ void epilogue_rows(const GemmParams& p, int b, int m0, int r0, int rstep, int rows_in_tile, const float* tile, int ldt, int cc,
                                              const ColOps& c) {
    if (c.nv <= 0) return;
    const int ph = p.ct_u ? c.n / p.ct_cout - p.ct_pad : 0;
    for (int r = r0; r < rows_in_tile; r += 2 * rstep) {
        float v[2][8], rr[2][8], oo[2][8];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int ru = r + u * rstep, m = m0 + ru;
            ok[u] = ru < rows_in_tile && m < p.M;
            if (ok[u] && p.ct_u) { const int tt = m * p.ct_u + ph; ok[u] = tt >= 0 && tt < p.ct_len; }
            if (ok[u]) {
                if (p.res && !p.glu) ld8(p.res + (long)b * p.r_bs + (long)m * p.ldr + c.n, c.nv, rr[u]);
                if (p.outF && p.accumulate) ld8(p.outF + (long)b * p.c_bs + (long)m * p.ldc + c.n, c.nv, oo[u]);
                const float4 x0 = *reinterpret_cast<const float4*>(tile + ru * ldt + cc), x1 = *reinterpret_cast<const float4*>(tile + ru * ldt + cc + 4);
                v[u][0] = x0.x; v[u][1] = x0.y; v[u][2] = x0.z; v[u][3] = x0.w; v[u][4] = x1.x; v[u][5] = x1.y; v[u][6] = x1.z; v[u][7] = x1.w;
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++)
            if (ok[u]) epilogue_math_store8<ACT, ACT2>(p, b, m0 + r + u * rstep, c.n, c.nv, v[u], c.t0, c.t1, c.a1, c.a2, rr[u], oo[u]);
    }
}
