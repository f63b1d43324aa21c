// Launch declarations for the non-GEMM kernels (norm_act.cu, t3_kernels.cu, hift_kernels.cu).
#pragma once
#include "common.cuh"

struct NormParams {
    const float* in = nullptr; long ld_in = 0; long in_bs = 0;
    int rows = 0, C = 0, batch = 1;
    const float* gain = nullptr; const float* bias = nullptr;  // bias == nullptr and rms == 1 -> RMSNorm
    int rms = 0; float eps = 1e-5f;
    int act = ACT_NONE;
    const float* add = nullptr; long add_bs = 0;   // per-batch vector added after the activation (time embedding)
    float out_scale = 1.f;
    float* outF = nullptr; long ld_outF = 0; long outF_bs = 0;
    bf16* outB = nullptr; long ld_outB = 0; long outB_bs = 0;
};
void launch_norm(const NormParams& p, cudaStream_t st);
void launch_gather_rows_bf16(const float* table, const int* idx, int n, int C, bf16* out, long ld, cudaStream_t st);
void launch_f32_to_bf16_rows(const float* in, long ld_in, bf16* out, long ld_out, int rows, int C, int act, float act_param, cudaStream_t st);
void launch_f32_to_bf16_slabs(const float* in, long ld_in, long in_bs, bf16* out, long ld_out, long out_bs, int rows, int C, int batch, cudaStream_t st);
void launch_add_rows(float* a, long lda, const float* b, long ldb, int rows, int C, cudaStream_t st);
void launch_upsample2(const float* in, bf16* out, long ld_out, int T, int C, cudaStream_t st);
void launch_pack_cfm_input(const float* x, const float* mu, const float* spks, const float* cond, bf16* out, long out_bs, int T, int mel, cudaStream_t st);
void launch_euler_update(float* x, const float* v, long v_bs, long n, float dt, float r, cudaStream_t st);
void launch_small_attention(const bf16* q, long ldq, const bf16* k, const bf16* v, long ldk, bf16* out, long ldo, int Tq, int Tk, int H, int hd, float scale, cudaStream_t st);
