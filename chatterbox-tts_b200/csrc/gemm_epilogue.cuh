// Fused GEMM epilogue shared by the mma.sync (gemm.cu) and tcgen05 (gemm_tc.cu) kernels.
#pragma once
#include "common.cuh"

__device__ __forceinline__ void epilogue_pair(const GemmParams& p, int b, int m, int n, float v0, float v1) {
    // n is even; handles columns n, n+1 of row m
    if (m >= p.M || n >= p.N) return;
    const bool has1 = (n + 1) < p.N;
    if (p.ct_u) {
        int ph = n / p.ct_cout;
        int t = m * p.ct_u + ph - p.ct_pad;
        if (t < 0 || t >= p.ct_len) return;
    }
    if (p.bias) { v0 += p.bias[n]; if (has1) v1 += p.bias[n + 1]; }
    if (p.bias2) { const float* b2 = p.bias2 + (long)b * p.bias2_bs; v0 += b2[n]; if (has1) v1 += b2[n + 1]; }
    if (p.glu) {
        float g = v0 / (1.f + expf(-v0)) * v1;
        long o = (long)b * p.c_bs + (long)m * p.ldc + (n >> 1);
        if (p.outB) p.outB[o] = __float2bfloat16(g);
        if (p.outF) p.outF[o] = g;
        return;
    }
    if (p.act) {
        float a0 = p.act_alpha ? p.act_alpha[n] : p.act_param;
        float a1 = (p.act_alpha && has1) ? p.act_alpha[n + 1] : p.act_param;
        v0 = act_apply(p.act, v0, a0);
        v1 = act_apply(p.act, v1, a1);
    }
    if (p.res) {
        const float* r = p.res + (long)b * p.r_bs + (long)m * p.ldr + n;
        v0 += r[0]; if (has1) v1 += r[1];
    }
    long o = (long)b * p.c_bs + (long)m * p.ldc + n;
    if (p.outF) {
        float w0 = p.out_scale * v0, w1 = p.out_scale * v1;
        if (p.accumulate) { w0 += p.outF[o]; if (has1) w1 += p.outF[o + 1]; }
        p.outF[o] = w0; if (has1) p.outF[o + 1] = w1;
        v0 = w0; v1 = w1;
    } else {
        v0 *= p.out_scale; v1 *= p.out_scale;
    }
    if (p.outB) {
        if (has1 && ((o & 1) == 0)) *reinterpret_cast<uint32_t*>(p.outB + o) = pack_bf16(v0, v1);
        else { p.outB[o] = __float2bfloat16(v0); if (has1) p.outB[o + 1] = __float2bfloat16(v1); }
    }
    if (p.outB2) {
        float a0 = p.act2_alpha ? p.act2_alpha[n] : p.act2_param;
        float a1 = (p.act2_alpha && has1) ? p.act2_alpha[n + 1] : p.act2_param;
        long o2 = (long)b * p.c2_bs + (long)m * p.ldc2 + n;
        float u0 = act_apply(p.act2, v0, a0), u1 = act_apply(p.act2, v1, a1);
        if (has1 && ((o2 & 1) == 0)) *reinterpret_cast<uint32_t*>(p.outB2 + o2) = pack_bf16(u0, u1);
        else { p.outB2[o2] = __float2bfloat16(u0); if (has1) p.outB2[o2 + 1] = __float2bfloat16(u1); }
    }
}

