// HiFT vocoder kernels that are not GEMM-shaped (sm_100a): F0 classifier, harmonic+noise source,
// 16-point STFT of the source, fused exp/sin -> iSTFT -> clamp -> trim-fade, Snake, ESPnet rel-pos table.
#include "common.cuh"
#include "hift_kernels.cuh"

namespace {

__constant__ float c_hann16[16];
__constant__ float c_cos16[16];
__constant__ float c_sin16[16];

// f0[t] = | w . x[t] + b |  (x bf16 [T][C], one warp per frame)
__global__ void f0_classifier_kernel(const bf16* __restrict__ x, long ld, const float* __restrict__ w, const float* __restrict__ b, float* f0, int T, int C) {
    pdl_prologue();
    int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= T) return;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __bfloat162float(x[(long)t * ld + c]) * w[c];
    s = warp_sum(s);
    if (lane == 0) f0[t] = fabsf(s + b[0]);
}

// per-frame, per-harmonic phase prefix (cycles, double): cum[t][h] = sum_{u<t} up * fp32(f0[u]*(h+1)/sr)
__global__ void f0_prefix_kernel(const float* __restrict__ f0, double* cum, int T, int up, float sr, int H) {
    pdl_prologue();
    int h = threadIdx.x;
    if (h < H && blockIdx.x == 0) {
        double c = 0.0;
        for (int t = 0; t < T; t++) {
            cum[(long)t * H + h] = c;
            float inc = (f0[t] * (float)(h + 1)) / sr;
            c += (double)inc * up;
        }
    }
}

// source sample n (frame t = n / up, j = n % up): harmonics h=1..H+1
//   theta_h = 2*pi*frac(cum[t][h] + (j+1) * fp32(f0[t]*h/sr)),  sine = amp*sin(theta_h + phase_h)
//   s = tanh( sum_h lw[h] * (sine*uv + namp*noise_h) + lb )
__global__ void source_kernel(const SourceParams p) {
    pdl_prologue();
    long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.L) return;
    const long cache_len = p.dyn ? (long)p.dyn->cache_len : p.cache_len;
    const unsigned long long seed = p.dyn ? p.dyn->seed : p.seed;
    if (n < cache_len) { p.s[n] = p.cache[n]; return; }
    int t = n / p.up, j = n % p.up;
    float f = p.f0[t];
    float uv = f > p.voiced_thr ? 1.f : 0.f;
    float namp = uv * p.noise_std + (1.f - uv) * p.sine_amp / 3.f;
    float acc = p.lb[0];
    const int H = p.n_harm;
    for (int h = 0; h < H; h++) {
        float inc = (f * (float)(h + 1)) / p.sr;
        double ph = p.cum[(long)t * H + h] + (double)(j + 1) * (double)inc;
        float theta = 6.283185307179586f * (float)(ph - floor(ph));
        float phase0 = (h == 0) ? 0.f : (p.phase ? p.phase[h] : 0.f);
        if (!p.phase && h > 0) {
            uint32_t r4[4];
            Philox::gen(seed, (uint32_t)h, 0u, 0x5048u, 0u, r4);
            phase0 = (u32_to_unit(r4[0]) * 2.f - 1.f) * 3.14159265358979f;
        }
        float sine = p.sine_amp * sinf(theta + phase0);
        float nz;
        if (p.noise) nz = p.noise[(long)h * p.L + n];
        else {
            uint32_t r4[4];
            Philox::gen(seed, (uint32_t)n, (uint32_t)(n >> 32), 0x4E5Au + h, 0u, r4);
            float u1 = u32_to_unit(r4[0]), u2 = u32_to_unit(r4[1]);
            nz = sqrtf(-2.f * logf(u1)) * cosf(6.283185307179586f * u2);
        }
        acc += p.lw[h] * (sine * uv + namp * nz);
    }
    p.s[n] = tanhf(acc);
}

// STFT (n_fft 16, hop 4, hann, center/reflect): out[f][c] c<9 real, 9..17 imag, bf16, row stride ld (>=18; extra channels zero)
__global__ void stft16_kernel(const float* __restrict__ s, long L, bf16* out, long ld, int F) {
    pdl_prologue();
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float x[16];
#pragma unroll
    for (int n = 0; n < 16; n++) {
        long i = (long)f * 4 + n - 8;
        if (i < 0) i = -i;
        if (i >= L) i = 2 * (L - 1) - i;
        x[n] = s[i] * c_hann16[n];
    }
    bf16* o = out + (long)f * ld;
#pragma unroll
    for (int c = 0; c < 9; c++) {
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int n = 0; n < 16; n++) { int k = (c * n) & 15; re += x[n] * c_cos16[k]; im -= x[n] * c_sin16[k]; }
        o[c] = __float2bfloat16(re);
        o[9 + c] = __float2bfloat16(im);
    }
    for (int c = 18; c < ld; c++) o[c] = __float2bfloat16(0.f);
}

// y[F][18] (fp32: 9 log-magnitudes, 9 phase pre-activations) -> wav[4*(F-1)], with clamp and leading trim-fade
__global__ void istft16_kernel(const float* __restrict__ y, long ldy, int F, float* wav, long L, float limit, const float* __restrict__ fade, int fade_len, long n0) {
    pdl_prologue();
    long n = n0 + (long)blockIdx.x * blockDim.x + threadIdx.x;      // samples [n0, L): a decode window leaves the earlier ones alone
    if (n >= L) return;
    // frames f with 0 <= n + 8 - 4f < 16
    int fhi = (int)((n + 8) >> 2);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        int f = fhi - d;
        if (f < 0 || f >= F) continue;
        int m = (int)(n + 8 - 4 * (long)f);
        const float* yr = y + (long)f * ldy;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 9; c++) {
            float mag = fminf(expf(yr[c]), 100.f);
            float ph = sinf(yr[9 + c]);
            float re = mag * cosf(ph), im = mag * sinf(ph);
            int k = (c * m) & 15;
            float wgt = (c == 0 || c == 8) ? 1.f : 2.f;
            float term = re * c_cos16[k] - ((c == 0 || c == 8) ? 0.f : im * c_sin16[k]);
            acc += wgt * term;
        }
        float w = c_hann16[m];
        num += w * acc * (1.f / 16.f);
        den += w * w;
    }
    float v = num / den;
    v = fminf(fmaxf(v, -limit), limit);
    if (n < fade_len) v *= fade[n];
    wav[n] = v;
}

__global__ void snake_rows_kernel(const float* __restrict__ in, long ld_in, bf16* out, long ld_out, int rows, int C, const float* __restrict__ alpha) {
    pdl_prologue();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * C) return;
    int r = i / C, c = i % C;
    out[(long)r * ld_out + c] = __float2bfloat16(act_apply(ACT_SNAKE, in[(long)r * ld_in + c], alpha[c]));
}

// ESPnet relative positional table: row j <-> relative position T-1-j; bf16 [2T-1][D]
__global__ void relpos_table_kernel(bf16* out, int T, int D) {
    pdl_prologue();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)(2 * T - 1) * D) return;
    int j = i / D, c = i % D;
    float pos = (float)(T - 1 - j);
    float div = expf((float)(c & ~1) * (-9.210340371976184f / D));
    float a = pos * div;
    out[i] = __float2bfloat16((c & 1) ? cosf(a) : sinf(a));
}

__global__ void copy_row_kernel(float* dst, const float* src, int C) {
    pdl_prologue();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) dst[c] = src[c];
}

__global__ void spk_affine_kernel(const float* __restrict__ emb, int D, const float* __restrict__ w, const float* __restrict__ b, float* out, int N) {
    pdl_prologue();
    // out = W * (emb / max(||emb||, 1e-12)) + b ; one CTA
    __shared__ float red[32];
    __shared__ float nrm;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) s += emb[i] * emb[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < (blockDim.x >> 5); i++) t += red[i]; nrm = fmaxf(sqrtf(t), 1e-12f); }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float a = b[n];
        for (int i = 0; i < D; i++) a += w[(long)n * D + i] * (emb[i] / nrm);
        out[n] = a;
    }
}

}  // namespace

static inline dim3 g1(long n, int t = 256) { return dim3((unsigned)((n + t - 1) / t)); }

void hift_init_constants() {
    float hann[16], cs[16], sn[16];
    for (int n = 0; n < 16; n++) {
        hann[n] = 0.5f - 0.5f * cosf(6.283185307179586f * n / 16.f);
        cs[n] = cosf(6.283185307179586f * n / 16.f);
        sn[n] = sinf(6.283185307179586f * n / 16.f);
    }
    CBX_CHECK(cudaMemcpyToSymbol(c_hann16, hann, sizeof(hann)));
    CBX_CHECK(cudaMemcpyToSymbol(c_cos16, cs, sizeof(cs)));
    CBX_CHECK(cudaMemcpyToSymbol(c_sin16, sn, sizeof(sn)));
}
void launch_f0_classifier(const bf16* x, long ld, const float* w, const float* b, float* f0, int T, int C, cudaStream_t st) {
    launch_pdl(f0_classifier_kernel, dim3(g1((long)T * 32)), dim3(256), 0, st, x, ld, w, b, f0, T, C);
    CBX_CHECK(cudaGetLastError());
}
void launch_source(const SourceParams& p, int T, cudaStream_t st) {
    ProfScope ps(PC_HIFT_MISC, (double)p.L * 4, st);
    launch_pdl(f0_prefix_kernel, dim3(1), dim3(32), 0, st, p.f0, p.cum, T, p.up, p.sr, p.n_harm);
    launch_pdl(source_kernel, dim3(g1(p.L)), dim3(256), 0, st, p);
    CBX_CHECK(cudaGetLastError());
}
void launch_stft16(const float* s, long L, bf16* out, long ld, int F, cudaStream_t st) {
    ProfScope ps(PC_HIFT_MISC, (double)L * 4 + (double)F * ld * 2, st);
    launch_pdl(stft16_kernel, g1(F, 128), dim3(128), 0, st, s, L, out, ld, F);
    CBX_CHECK(cudaGetLastError());
}
void launch_istft16(const float* y, long ldy, int F, float* wav, long L, float limit, const float* fade, int fade_len, long n0, cudaStream_t st) {
    ProfScope ps(PC_HIFT_MISC, (double)F * ldy * 4 + (double)L * 4, st);
    if (n0 >= L) return;
    launch_pdl(istft16_kernel, dim3(g1(L - n0)), dim3(256), 0, st, y, ldy, F, wav, L, limit, fade, fade_len, n0);
    CBX_CHECK(cudaGetLastError());
}
void launch_snake_rows(const float* in, long ld_in, bf16* out, long ld_out, int rows, int C, const float* alpha, cudaStream_t st) {
    ProfScope ps(PC_HIFT_MISC, (double)rows * C * 6, st);
    launch_pdl(snake_rows_kernel, dim3(g1((long)rows * C)), dim3(256), 0, st, in, ld_in, out, ld_out, rows, C, alpha);
    CBX_CHECK(cudaGetLastError());
}
void launch_relpos_table(bf16* out, int T, int D, cudaStream_t st) {
    launch_pdl(relpos_table_kernel, dim3(g1((long)(2 * T - 1) * D)), dim3(256), 0, st, out, T, D);
    CBX_CHECK(cudaGetLastError());
}
void launch_copy_row(float* dst, const float* src, int C, cudaStream_t st) {
    launch_pdl(copy_row_kernel, dim3(g1(C)), dim3(256), 0, st, dst, src, C);
    CBX_CHECK(cudaGetLastError());
}
void launch_spk_affine(const float* emb, int D, const float* w, const float* b, float* out, int N, cudaStream_t st) {
    launch_pdl(spk_affine_kernel, dim3(1), dim3(128), 0, st, emb, D, w, b, out, N);
    CBX_CHECK(cudaGetLastError());
}
