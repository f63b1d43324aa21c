#pragma once
#include "common.cuh"

struct SourceDyn { long long cache_len; unsigned long long seed; };   // per-call values read from device memory (CUDA-graph replay)

struct SourceParams {
    const float* f0 = nullptr; double* cum = nullptr;
    float* s = nullptr; long L = 0; int up = 480; float sr = 24000.f;
    int n_harm = 9; float sine_amp = 0.1f, noise_std = 0.003f, voiced_thr = 10.f;
    const float* lw = nullptr; const float* lb = nullptr;   // l_linear [n_harm], [1]
    const float* phase = nullptr;    // optional explicit initial phases [n_harm]
    const float* noise = nullptr;    // optional explicit N(0,1) noise [n_harm][L]
    const float* cache = nullptr; long cache_len = 0;   // cache_source overwrite of the first samples
    unsigned long long seed = 0;
    const SourceDyn* dyn = nullptr;   // non-null: cache_len / seed come from here instead of the fields above
};
void hift_init_constants();
void launch_f0_classifier(const bf16* x, long ld, const float* w, const float* b, float* f0, int T, int C, cudaStream_t st);
void launch_source(const SourceParams& p, int T, cudaStream_t st);
void launch_stft16(const float* s, long L, bf16* out, long ld, int F, cudaStream_t st);
void launch_istft16(const float* y, long ldy, int F, float* wav, long L, float limit, const float* fade, int fade_len, long n0, cudaStream_t st);
void launch_snake_rows(const float* in, long ld_in, bf16* out, long ld_out, int rows, int C, const float* alpha, cudaStream_t st);
void launch_relpos_table(bf16* out, int T, int D, cudaStream_t st);
void launch_copy_row(float* dst, const float* src, int C, cudaStream_t st);
void launch_spk_affine(const float* emb, int D, const float* w, const float* b, float* out, int N, cudaStream_t st);
