// Voice-conditioning encoders (reference src/tts_streaming.py:357-384: s3gen.embed_ref, s3gen.tokenizer.forward,
// ve.embeds_from_wavs): fp32 building blocks on the CUDA cores.
//
// This path runs once per VOICE (its result is cached, reference voice_cache :178), not per token, and its results feed
// hard decisions -- the S3Tokenizer's FSQ codebook rounds eight tanh outputs to {-1, 0, 1}, so a bf16 tensor-core GEMM in
// front of it flips token ids against an fp32 reference.  Everything here is therefore plain fp32 (a 10 s clip is ~35 GFLOP:
// a few milliseconds at SIMT rates); the tcgen05 kernels stay on the per-token path.  The host side
// (cbx_b200/conditioning.py) strings these ops together the way the upstream modules do.
//
// Layout: time-major, channels-last ([T][C]); a conv1d is a GEMM whose A row t gathers the taps t*stride + j*dil - pad
// (zero outside the sequence), weights repacked by the host to [C_out][k][C_in].
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string>
#include <stdexcept>
#include "common.cuh"
#include "../../include/cbx_b200.h"

void cbx_set_error(const char* msg);     // engine.cu: thread-local message behind cbx_last_error()

#define COND_API_BEGIN try {
#define COND_API_END                                               \
    CBX_CHECK(cudaGetLastError());                                 \
    return 0;                                                      \
    } catch (const std::exception& ex) { cbx_set_error(ex.what()); return 1; } \
    catch (...) { cbx_set_error("unknown error"); return 1; }

namespace {

__device__ __forceinline__ float act_f(int act, float v) {
    switch (act) {
        case 1: return fmaxf(v, 0.f);
        case 2: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
        case 3: return 1.f / (1.f + expf(-v));
        case 4: return tanhf(v);
        default: return v;
    }
}

// ------------------------------------------------------------------------------------------------------------ sgemm
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256) sgemm_kernel(const cbx_sgemm_args p) {
    __shared__ __align__(16) float As[GK][GT + 4], Ws[GK][GT + 4];
    const int b = blockIdx.z, m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* A = p.A + (long)b * p.a_bs;
    const float* W = p.W + (long)b * p.w_bs;
    float acc[4][4] = {};
    // fast path: every 16-wide K tile lies inside one conv tap and all operand rows are 16-byte aligned -> one float4 per thread
    // and tile for A and for W (no per-element index arithmetic), float4 reads of the tiles in the inner loop
    const bool vec = (p.kc % GK == 0) && (p.K % GK == 0) && (p.lda % 4 == 0) && (p.ldw % 4 == 0) && (p.a_bs % 4 == 0) && (p.w_bs % 4 == 0) && !p.w_trans &&
                     ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) == 0;
    for (int k0 = 0; k0 < p.K; k0 += GK) {
        if (vec) {
            const int r = threadIdx.x >> 2, q4 = (threadIdx.x & 3) * 4;       // 64 rows x 4 float4 columns
            const int tap = k0 / p.kc, c0 = k0 - tap * p.kc + q4;
            {
                const int m = m0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const long t_in = (long)m * p.a_stride + (long)tap * p.a_dil - p.a_pad;
                if (m < p.M && t_in >= 0 && t_in < p.a_rows) {
                    v = *reinterpret_cast<const float4*>(A + t_in * p.lda + c0);
                    if (p.a_scale) {
                        const float4 sc = *reinterpret_cast<const float4*>(p.a_scale + c0), sh = *reinterpret_cast<const float4*>(p.a_shift + c0);
                        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                    }
                    if (p.a_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                }
                As[q4][r] = v.x; As[q4 + 1][r] = v.y; As[q4 + 2][r] = v.z; As[q4 + 3][r] = v.w;
            }
            {
                const int n = n0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < p.N) v = *reinterpret_cast<const float4*>(W + (long)n * p.ldw + k0 + q4);
                Ws[q4][r] = v.x; Ws[q4 + 1][r] = v.y; Ws[q4 + 2][r] = v.z; Ws[q4 + 3][r] = v.w;
            }
        } else {
        // A tile: 64 rows x 16 k (gathered conv taps, optional per-channel affine + ReLU on the way in)
        for (int i = threadIdx.x; i < GT * GK; i += 256) {
            const int r = i / GK, kk = k0 + i % GK, m = m0 + r;
            float v = 0.f;
            if (m < p.M && kk < p.K) {
                const int tap = kk / p.kc, c = kk - tap * p.kc;
                const long t_in = (long)m * p.a_stride + (long)tap * p.a_dil - p.a_pad;
                if (t_in >= 0 && t_in < p.a_rows) {
                    v = A[t_in * p.lda + c];
                    if (p.a_scale) v = fmaf(v, p.a_scale[c], p.a_shift[c]);
                    if (p.a_relu) v = fmaxf(v, 0.f);
                }
            }
            As[i % GK][r] = v;
        }
        for (int i = threadIdx.x; i < GT * GK; i += 256) {
            float v = 0.f;
            if (p.w_trans) {         // W given as [K][N]
                const int kk = k0 + i / GT, n = n0 + i % GT;
                if (kk < p.K && n < p.N) v = W[(long)kk * p.ldw + n];
                Ws[i / GT][i % GT] = v;
            } else {                 // W given as [N][K]
                const int n = n0 + i / GK, kk = k0 + i % GK;
                if (kk < p.K && n < p.N) v = W[(long)n * p.ldw + kk];
                Ws[i % GK][i / GK] = v;
            }
        }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; k++) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]), w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* C = p.C + (long)b * p.c_bs;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j] * p.alpha;
            if (p.bias) v += p.bias[n];
            if (p.o_scale) v = fmaf(v, p.o_scale[n], p.o_shift[n]);
            v = act_f(p.act, v);
            if (p.mul) v *= p.mul[(long)b * p.mul_bs + (long)(m / p.mul_div) * p.ldm + n];
            if (p.res) v += p.res[(long)b * p.res_bs + (long)m * p.ldr + n];
            if (p.res2) v += p.res2[(long)b * p.res_bs + (long)m * p.ldr + n];
            C[(long)m * p.ldc + n] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------ framing + DFT
// One block per frame: |DFT|^2-style spectra by direct summation against a shared twiddle table (n_fft <= 2048; a clip has a
// few hundred frames, so this is microseconds).  mode 0: magnitude, 1: power, 2: sqrt(power + 1e-9).
__global__ void __launch_bounds__(256) frames_dft_kernel(const float* __restrict__ wav, long n, int n_fft, int hop, int pad, int frame_len,
                                                         const float* __restrict__ window, int remove_dc, float preemph, int mode,
                                                         float* __restrict__ out, int ld_out) {
    extern __shared__ float sm[];
    float* fr = sm;                 // [n_fft] windowed frame (zero beyond frame_len)
    float* cs = sm + n_fft;         // [n_fft] cos(2 pi m / N)
    float* sn = cs + n_fft;         // [n_fft] sin
    __shared__ float red[8];
    const int f = blockIdx.x;
    for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
        float s, c;
        sincospif(2.f * (float)i / (float)n_fft, &s, &c);
        cs[i] = c; sn[i] = s;
        float v = 0.f;
        if (i < frame_len) {
            long t = (long)f * hop + i - pad;
            if (t < 0) t = -t;
            if (t >= n) t = 2 * (n - 1) - t;
            v = (t >= 0 && t < n) ? wav[t] : 0.f;
        }
        fr[i] = v;
    }
    __syncthreads();
    if (remove_dc) {                // Kaldi: subtract the frame mean, then pre-emphasis with the first sample replicated
        float s = 0.f;
        for (int i = threadIdx.x; i < frame_len; i += blockDim.x) s += fr[i];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        float mean = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); w++) mean += red[w];
        mean /= (float)frame_len;
        __syncthreads();
        for (int i = threadIdx.x; i < frame_len; i += blockDim.x) fr[i] -= mean;
        __syncthreads();
    }
    float keep[8];
    if (preemph != 0.f) {
        for (int i = threadIdx.x, q = 0; i < frame_len; i += blockDim.x, q++) keep[q] = fr[i] - preemph * fr[i > 0 ? i - 1 : 0];
        __syncthreads();
        for (int i = threadIdx.x, q = 0; i < frame_len; i += blockDim.x, q++) fr[i] = keep[q];
        __syncthreads();
    }
    for (int i = threadIdx.x; i < frame_len; i += blockDim.x) fr[i] *= window[i];
    __syncthreads();
    const int bins = n_fft / 2 + 1;
    for (int k = threadIdx.x; k < bins; k += blockDim.x) {
        float re0 = 0.f, im0 = 0.f, re1 = 0.f, im1 = 0.f;
        int idx = 0;
        int i = 0;
        for (; i + 1 < frame_len; i += 2) {
            re0 = fmaf(fr[i], cs[idx], re0); im0 = fmaf(fr[i], sn[idx], im0);
            idx += k; if (idx >= n_fft) idx -= n_fft;
            re1 = fmaf(fr[i + 1], cs[idx], re1); im1 = fmaf(fr[i + 1], sn[idx], im1);
            idx += k; if (idx >= n_fft) idx -= n_fft;
        }
        if (i < frame_len) { re0 = fmaf(fr[i], cs[idx], re0); im0 = fmaf(fr[i], sn[idx], im0); }
        const float re = re0 + re1, im = im0 + im1, pw = re * re + im * im;
        out[(long)f * ld_out + k] = mode == 0 ? sqrtf(pw) : mode == 1 ? pw : sqrtf(pw + 1e-9f);
    }
}

// windowed-sinc polyphase resampler (torchaudio.functional.resample): out[j] = sum_k kern[j % up][k] * x[(j / up) * down + k - width]
__global__ void resample_kernel(const float* __restrict__ x, long n_in, float* __restrict__ y, long n_out, int down, int up,
                                const float* __restrict__ kern, int klen, int width) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const int ph = (int)(j % up);
    const long base = (j / up) * down - width;
    float acc = 0.f;
    for (int k = 0; k < klen; k++) {
        const long t = base + k;
        if (t >= 0 && t < n_in) acc = fmaf(kern[ph * klen + k], x[t], acc);
    }
    y[j] = acc;
}

// ------------------------------------------------------------------------------------------------------------ row-wise ops
__global__ void layernorm_rows_kernel(const float* __restrict__ x, long ld_in, float* __restrict__ y, long ld_out, int rows, int C,
                                      const float* __restrict__ g, const float* __restrict__ b, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long)row * ld_in;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    for (int c = lane; c < C; c += 32) y[(long)row * ld_out + c] = (xr[c] - mean) * rstd * g[c] + b[c];
}

__global__ void softmax_rows_kernel(float* __restrict__ x, long ld, long bs, int rows, int cols) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* xr = x + (long)blockIdx.y * bs + (long)row * ld;
    float m = -INFINITY;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, xr[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) { const float e = expf(xr[c] - m); xr[c] = e; s += e; }
    const float inv = 1.f / warp_sum(s);
    for (int c = lane; c < cols; c += 32) xr[c] *= inv;
}

// rotate-half rotary embedding in place on [T][H][hd] (angles repeated over both halves, theta 10000), then a scale
__global__ void rotary_kernel(float* __restrict__ x, long ld, int T, int H, int hd, float scale) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int half = hd / 2;
    if (i >= (long)T * H * half) return;
    const int d = (int)(i % half), h = (int)((i / half) % H), t = (int)(i / ((long)half * H));
    const float inv = 1.f / powf(10000.f, (float)(2 * d) / (float)hd);
    float s, c;
    sincosf((float)t * inv, &s, &c);
    float* p = x + (long)t * ld + h * hd;
    const float a = p[d], b = p[d + half];
    p[d] = (a * c - b * s) * scale;
    p[d + half] = (b * c + a * s) * scale;
}

// depthwise conv along time with zero padding (k - 1) / 2 plus the input itself (FSMN memory block): [T][C]
__global__ void dwconv_add_kernel(const float* __restrict__ x, long ld, const float* __restrict__ w, int k, float* __restrict__ y, long ld_out, int T, int C) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)T * C) return;
    const int c = (int)(i % C), t = (int)(i / C);
    float acc = x[(long)t * ld + c];
    for (int j = 0; j < k; j++) {
        const int tt = t + j - (k - 1) / 2;
        if (tt >= 0 && tt < T) acc = fmaf(w[c * k + j], x[(long)tt * ld + c], acc);
    }
    y[(long)t * ld_out + c] = acc;
}

// FSQ codebook: h [T][8] (projected down) -> tanh * 0.999 -> round -> + 1 -> base-3 digits
__global__ void fsq_kernel(const float* __restrict__ h, int* __restrict__ out, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int code = 0, pw = 1;
    for (int d = 0; d < 8; d++) {
        const float v = rintf(tanhf(h[t * 8 + d]) * 0.9990000128746033f) + 1.f;
        code += (int)v * pw;
        pw *= 3;
    }
    out[t] = code;
}

// mel post-processing.  mode 0: log(max(x, floor)); 1: whisper log-mel (log10(max(x, 1e-10)), clipped to global max - 8,
// (x + 4) / 4; needs gmax = max over the tensor of log10 values, computed by colmax_kernel first)
__global__ void mel_log_kernel(float* __restrict__ x, long n, int mode, float floor_v, const float* __restrict__ gmax) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mode == 0) x[i] = logf(fmaxf(x[i], floor_v));
    else {
        const float v = log10f(fmaxf(x[i], 1e-10f));
        x[i] = (fmaxf(v, gmax[0] - 8.f) + 4.f) * 0.25f;
    }
}
__global__ void max_log10_kernel(const float* __restrict__ x, long n, float* __restrict__ out) {      // single block
    __shared__ float red[32];
    float m = -INFINITY;
    for (long i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, log10f(fmaxf(x[i], 1e-10f)));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
        m = warp_max(m);
        if (threadIdx.x == 0) out[0] = m;
    }
}

// per-column statistics over time of [T][C]: mode 0 subtracts the column mean in place (fbank mean normalisation);
// mode 1 writes [mean | unbiased std] (2C) (statistics pooling) after an optional per-channel affine + ReLU;
// mode 2 writes mean (C) (+ per-segment means into seg_out [nseg][C] with the global mean added: CAM context)
__global__ void col_stats_kernel(float* __restrict__ x, long ld, int T, int C, int mode, const float* __restrict__ a_scale, const float* __restrict__ a_shift,
                                 float* __restrict__ out, int seg, float* __restrict__ seg_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    auto at = [&](int t) { float v = x[(long)t * ld + c]; if (a_scale) v = fmaxf(fmaf(v, a_scale[c], a_shift[c]), 0.f); return v; };
    double s = 0.0;
    for (int t = 0; t < T; t++) s += at(t);
    const float mean = (float)(s / T);
    if (mode == 0) { for (int t = 0; t < T; t++) x[(long)t * ld + c] -= mean; return; }
    if (mode == 1) {
        double q = 0.0;
        for (int t = 0; t < T; t++) { const double d = at(t) - mean; q += d * d; }
        out[c] = mean; out[C + c] = (float)sqrt(q / (T - 1));
        return;
    }
    out[c] = mean;
    for (int s0 = 0, i = 0; s0 < T; s0 += seg, i++) {
        const int e = min(T, s0 + seg);
        float ss = 0.f;
        for (int t = s0; t < e; t++) ss += at(t);
        seg_out[(long)i * C + c] = ss / (float)(e - s0) + mean;
    }
}

__global__ void l2norm_rows_kernel(float* __restrict__ x, int rows, int C, int relu) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* xr = x + (long)row * C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { float v = xr[c]; if (relu) v = fmaxf(v, 0.f); q += v * v; }
    const float inv = rsqrtf(warp_sum(q));
    for (int c = lane; c < C; c += 32) { float v = xr[c]; if (relu) v = fmaxf(v, 0.f); xr[c] = v * inv; }
}
__global__ void mean_rows_kernel(const float* __restrict__ x, int rows, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int r = 0; r < rows; r++) s += x[(long)r * C + c];
    out[c] = s / rows;
}

// ------------------------------------------------------------------------------------------------------------ conv2d (FCM head)
// x [Cin][F][T] -> y = relu?(bn(conv(x)) (+ res)); 3x3 (pad 1) or 1x1 (pad 0), stride (sf, 1); y [Cout][F'][T] or, t_major,
// [T][Cout * F'] (the reshape that feeds the TDNN)
__global__ void conv2d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale, const float* __restrict__ shift,
                              const float* __restrict__ res, float* __restrict__ y, int Cin, int Cout, int F, int T, int ks, int sf, int relu, int t_major) {
    const int Fo = (F + 2 * (ks / 2) - ks) / sf + 1;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)Cout * Fo * T) return;
    const int t = (int)(i % T), fo = (int)((i / T) % Fo), co = (int)(i / ((long)T * Fo));
    const int pad = ks / 2;
    float acc = 0.f;
    for (int ci = 0; ci < Cin; ci++)
        for (int a = 0; a < ks; a++) {
            const int f = fo * sf + a - pad;
            if (f < 0 || f >= F) continue;
            for (int bb = 0; bb < ks; bb++) {
                const int tt = t + bb - pad;
                if (tt < 0 || tt >= T) continue;
                acc = fmaf(w[((co * Cin + ci) * ks + a) * ks + bb], x[((long)ci * F + f) * T + tt], acc);
            }
        }
    float v = fmaf(acc, scale[co], shift[co]);
    if (res) v += res[((long)co * Fo + fo) * T + t];
    if (relu) v = fmaxf(v, 0.f);
    if (t_major) y[(long)t * Cout * Fo + co * Fo + fo] = v;
    else y[((long)co * Fo + fo) * T + t] = v;
}

// ------------------------------------------------------------------------------------------------------------ LSTM layer
// One block per sequence, H threads: thread j accumulates the four consecutive gate rows 4j .. 4j+3 over the hidden state
// (w_hh_t [H][4H]: one 16-byte load per k, coalesced over j, eight in flight per thread), then combines the four gates of unit j.
// xp [B][T][4H] = x W_ih^T + b_ih + b_hh (one GEMM for all steps).  The accumulation order per gate row (input projection, then k
// ascending) is the oracle's.
__global__ void __launch_bounds__(256) lstm_layer_kernel(const float* __restrict__ xp, const float* __restrict__ w_hh_t, float* __restrict__ h_seq, float* __restrict__ h_last, int T, int H) {
    extern __shared__ __align__(16) float sm[];
    float* hs = sm;            // [H]
    float* gates = sm + H;     // [4H]
    const int b = blockIdx.x, j = threadIdx.x, G = 4 * H;
    float c = 0.f;
    hs[j] = 0.f;
    __syncthreads();
    for (int t = 0; t < T; t++) {
        float4 acc = *reinterpret_cast<const float4*>(xp + ((long)b * T + t) * G + 4 * j);
        const float4* w = reinterpret_cast<const float4*>(w_hh_t) + j;
#pragma unroll 8
        for (int k = 0; k < H; k++) {
            const float4 wv = __ldg(w + (long)k * H);
            const float hk = hs[k];
            acc.x = fmaf(wv.x, hk, acc.x); acc.y = fmaf(wv.y, hk, acc.y); acc.z = fmaf(wv.z, hk, acc.z); acc.w = fmaf(wv.w, hk, acc.w);
        }
        *reinterpret_cast<float4*>(gates + 4 * j) = acc;
        __syncthreads();       // every thread is past its reads of hs; the gates are complete
        const float gi = gates[j], gf = gates[H + j], gg = gates[2 * H + j], go = gates[3 * H + j];
        c = c / (1.f + expf(-gf)) + tanhf(gg) / (1.f + expf(-gi));
        const float h = tanhf(c) / (1.f + expf(-go));
        hs[j] = h;
        if (h_seq) h_seq[((long)b * T + t) * H + j] = h;
        __syncthreads();
    }
    if (h_last) h_last[(long)b * H + j] = hs[j];
}

}  // namespace

extern "C" {

int cbx_cond_sgemm(const cbx_sgemm_args* a, void* stream) {
    COND_API_BEGIN
    CBX_REQUIRE(a && a->A && a->W && a->C && a->M > 0 && a->N > 0 && a->K > 0 && a->batch >= 1, "cond sgemm: bad arguments");
    cbx_sgemm_args p = *a;
    if (p.kc <= 0) p.kc = p.K;
    if (p.a_stride <= 0) p.a_stride = 1;
    if (p.a_rows <= 0) p.a_rows = (long)p.M * p.a_stride;
    if (p.mul_div <= 0) p.mul_div = 1;
    if (p.alpha == 0.f) p.alpha = 1.f;
    sgemm_kernel<<<dim3(cdiv(p.N, GT), cdiv(p.M, GT), p.batch), 256, 0, (cudaStream_t)stream>>>(p);
    COND_API_END
}

int cbx_cond_frames_dft(const float* wav, int64_t n, int n_fft, int hop, int pad, int frame_len, const float* window, int remove_dc, float preemph,
                        int mode, int n_frames, float* out, int ld_out, void* stream) {
    COND_API_BEGIN
    CBX_REQUIRE(wav && window && out && n > 0 && n_fft >= 16 && n_fft <= 2048 && frame_len <= n_fft && frame_len <= 2048 && n_frames > 0, "cond frames_dft: bad arguments");
    frames_dft_kernel<<<n_frames, 256, 3 * n_fft * sizeof(float), (cudaStream_t)stream>>>(wav, n, n_fft, hop, pad, frame_len, window, remove_dc, preemph, mode, out, ld_out);
    COND_API_END
}

int cbx_cond_resample(const float* x, int64_t n_in, float* y, int64_t n_out, int down, int up, const float* kern, int klen, int width, void* stream) {
    COND_API_BEGIN
    CBX_REQUIRE(x && y && kern && n_in > 0 && n_out > 0 && down > 0 && up > 0, "cond resample: bad arguments");
    resample_kernel<<<cdiv(n_out, 256), 256, 0, (cudaStream_t)stream>>>(x, n_in, y, n_out, down, up, kern, klen, width);
    COND_API_END
}

int cbx_cond_layernorm(const float* x, int64_t ld_in, float* y, int64_t ld_out, int rows, int C, const float* g, const float* b, float eps, void* stream) {
    COND_API_BEGIN
    layernorm_rows_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, ld_in, y, ld_out, rows, C, g, b, eps);
    COND_API_END
}

int cbx_cond_softmax(float* x, int64_t ld, int64_t bs, int rows, int cols, int batch, void* stream) {
    COND_API_BEGIN
    softmax_rows_kernel<<<dim3(cdiv(rows, 8), batch), 256, 0, (cudaStream_t)stream>>>(x, ld, bs, rows, cols);
    COND_API_END
}

int cbx_cond_rotary(float* x, int64_t ld, int T, int H, int hd, float scale, void* stream) {
    COND_API_BEGIN
    rotary_kernel<<<cdiv((long)T * H * (hd / 2), 256), 256, 0, (cudaStream_t)stream>>>(x, ld, T, H, hd, scale);
    COND_API_END
}

int cbx_cond_dwconv_add(const float* x, int64_t ld, const float* w, int k, float* y, int64_t ld_out, int T, int C, void* stream) {
    COND_API_BEGIN
    dwconv_add_kernel<<<cdiv((long)T * C, 256), 256, 0, (cudaStream_t)stream>>>(x, ld, w, k, y, ld_out, T, C);
    COND_API_END
}

int cbx_cond_fsq(const float* h, int* out, int T, void* stream) {
    COND_API_BEGIN
    fsq_kernel<<<cdiv(T, 128), 128, 0, (cudaStream_t)stream>>>(h, out, T);
    COND_API_END
}

int cbx_cond_mel_log(float* x, int64_t n, int mode, float floor_v, float* scratch1, void* stream) {
    COND_API_BEGIN
    if (mode == 1) {
        CBX_REQUIRE(scratch1, "cond mel_log: the whisper form needs one float of scratch");
        max_log10_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, scratch1);
    }
    mel_log_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, mode, floor_v, scratch1);
    COND_API_END
}

int cbx_cond_col_stats(float* x, int64_t ld, int T, int C, int mode, const float* a_scale, const float* a_shift, float* out, int seg, float* seg_out, void* stream) {
    COND_API_BEGIN
    CBX_REQUIRE(x && T > 0 && C > 0 && (mode == 0 || out) && (mode != 2 || (seg > 0 && seg_out)), "cond col_stats: bad arguments");
    col_stats_kernel<<<cdiv(C, 64), 64, 0, (cudaStream_t)stream>>>(x, ld, T, C, mode, a_scale, a_shift, out, seg, seg_out);
    COND_API_END
}

int cbx_cond_l2norm_rows(float* x, int rows, int C, int relu, void* stream) {
    COND_API_BEGIN
    l2norm_rows_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, C, relu);
    COND_API_END
}

int cbx_cond_mean_rows(const float* x, int rows, int C, float* out, void* stream) {
    COND_API_BEGIN
    mean_rows_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(x, rows, C, out);
    COND_API_END
}

int cbx_cond_conv2d(const float* x, const float* w, const float* scale, const float* shift, const float* res, float* y, int Cin, int Cout, int F, int T,
                    int ks, int stride_f, int relu, int t_major, void* stream) {
    COND_API_BEGIN
    CBX_REQUIRE(x && w && scale && shift && y && (ks == 1 || ks == 3) && stride_f >= 1, "cond conv2d: bad arguments");
    const int Fo = (F + 2 * (ks / 2) - ks) / stride_f + 1;
    conv2d_kernel<<<cdiv((long)Cout * Fo * T, 256), 256, 0, (cudaStream_t)stream>>>(x, w, scale, shift, res, y, Cin, Cout, F, T, ks, stride_f, relu, t_major);
    COND_API_END
}

int cbx_cond_lstm_layer(const float* xp, const float* w_hh_t, float* h_seq, float* h_last, int B, int T, int H, void* stream) {
    COND_API_BEGIN
    CBX_REQUIRE(xp && w_hh_t && B > 0 && T > 0 && H > 0 && H <= 256 && H % 4 == 0, "cond lstm: bad arguments (hidden size must fit one block)");
    lstm_layer_kernel<<<B, H, 5 * H * sizeof(float), (cudaStream_t)stream>>>(xp, w_hh_t, h_seq, h_last, T, H);
    COND_API_END
}

}  // extern "C"
