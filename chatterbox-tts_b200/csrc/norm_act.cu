// Row-wise normalisation and small elementwise kernels shared by T3 prefill, the flow encoder,
// the CFM estimator and HiFT (sm_100a).  All activations are time-major, channels-last.
#include "common.cuh"
#include "kernels.cuh"

namespace {

// One warp per row, the row cached in registers as float4 (C = 128 * V4).  Templated on the row width and the fused
// activation: the generic version (runtime width, per-element activation switch) was 35% of an S3Gen call.
template <int V4, int ACT>
__global__ void __launch_bounds__(256) norm_kernel(NormParams p) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const long total = (long)p.rows * p.batch;
    if (warp >= total) return;
    const int b = warp / p.rows, r = warp % p.rows;
    const float* x = p.in + (long)b * p.in_bs + (long)r * p.ld_in;
    constexpr int C = 128 * V4;
    float4 v[V4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; i++) {
        v[i] = *reinterpret_cast<const float4*>(x + (i * 32 + lane) * 4);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = p.rms ? 0.f : warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V4; i++) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + p.eps);
    const float* add = p.add ? p.add + (long)b * p.add_bs : nullptr;
#pragma unroll
    for (int i = 0; i < V4; i++) {
        const int c = (i * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(p.gain + c);
        float y[4] = {v[i].x * rstd * g.x, v[i].y * rstd * g.y, v[i].z * rstd * g.z, v[i].w * rstd * g.w};
        if (p.bias) { const float4 bb = *reinterpret_cast<const float4*>(p.bias + c); y[0] += bb.x; y[1] += bb.y; y[2] += bb.z; y[3] += bb.w; }
        if (ACT == ACT_MISH) {
#pragma unroll
            for (int j = 0; j < 4; j++) y[j] = mish_fast(y[j]);      // exp + division (common.cuh), not log1p + tanh: the Mish, not the 6 B/element, bounded this kernel
        }
        if (add) { const float4 aa = *reinterpret_cast<const float4*>(add + c); y[0] += aa.x; y[1] += aa.y; y[2] += aa.z; y[3] += aa.w; }
#pragma unroll
        for (int j = 0; j < 4; j++) y[j] *= p.out_scale;
        if (p.outF) *reinterpret_cast<float4*>(p.outF + (long)b * p.outF_bs + (long)r * p.ld_outF + c) = make_float4(y[0], y[1], y[2], y[3]);
        if (p.outB) *reinterpret_cast<uint2*>(p.outB + (long)b * p.outB_bs + (long)r * p.ld_outB + c) = make_uint2(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]));
    }
}

__global__ void gather_rows_bf16_kernel(const float* __restrict__ table, const int* __restrict__ idx, int n, int C, bf16* out, long ld) {
    pdl_prologue();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n * C) return;
    int r = i / C, c = i % C;
    int t = idx[r];
    out[(long)r * ld + c] = __float2bfloat16(table[(long)t * C + c]);
}

__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ in, long ld_in, bf16* out, long ld_out, int rows, int C, int act, float act_param) {
    pdl_prologue();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * C) return;
    int r = i / C, c = i % C;
    out[(long)r * ld_out + c] = __float2bfloat16(act_apply(act, in[(long)r * ld_in + c], act_param));
}

// batched form: slab b of `rows` rows at in + b*in_bs -> out + b*out_bs (4 columns per thread; C % 4 == 0)
__global__ void f32_to_bf16_slabs_kernel(const float* __restrict__ in, long ld_in, long in_bs, bf16* out, long ld_out, long out_bs, int rows, int C) {
    pdl_prologue();
    const int b = blockIdx.y;
    const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= (long)rows * C) return;
    const int r = i / C, c = i % C;
    const float4 v = *reinterpret_cast<const float4*>(in + b * in_bs + (long)r * ld_in + c);
    *reinterpret_cast<uint2*>(out + b * out_bs + (long)r * ld_out + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}

__global__ void add_rows_kernel(float* a, long lda, const float* b, long ldb, int rows, int C) {
    pdl_prologue();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * C) return;
    int r = i / C, c = i % C;
    a[(long)r * lda + c] += b[(long)r * ldb + c];
}

// nearest x2 upsample along time, fp32 [T][C] -> bf16 [2T][C]
__global__ void upsample2_kernel(const float* __restrict__ in, bf16* out, long ld_out, int T, int C) {
    pdl_prologue();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)2 * T * C) return;
    int r = i / C, c = i % C;
    out[(long)r * ld_out + c] = __float2bfloat16(in[(long)(r >> 1) * C + c]);
}

// CFM estimator input: rows [x | mu | spks | cond] (cond row) and [x | 0 | 0 | 0] (uncond row), bf16, 4*mel channels
__global__ void pack_cfm_input_kernel(const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ spks,
                                      const float* __restrict__ cond, bf16* out, long out_bs, int T, int mel) {
    pdl_prologue();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int C = 4 * mel;
    if (i >= (long)2 * T * C) return;
    int b = i / ((long)T * C);
    long rem = i % ((long)T * C);
    int t = rem / C, c = rem % C;
    int part = c / mel, cc = c % mel;
    float v;
    if (part == 0) v = x[(long)t * mel + cc];
    else if (b == 1) v = 0.f;
    else if (part == 1) v = mu[(long)t * mel + cc];
    else if (part == 2) v = spks[cc];
    else v = cond[(long)t * mel + cc];
    out[(long)b * out_bs + (long)t * C + c] = __float2bfloat16(v);
}

// x += dt * ((1+r) v_c - r v_u)
__global__ void euler_update_kernel(float* x, const float* __restrict__ v, long v_bs, long n, float dt, float r) {
    pdl_prologue();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] += dt * ((1.f + r) * v[i] - r * v[v_bs + i]);
}

// one CTA per (query row, head); generic head_dim (perceiver: 256).  q [Tq][ldq], k/v [Tk][ldk]
__global__ void small_attention_kernel(const bf16* __restrict__ q, long ldq, const bf16* __restrict__ k, const bf16* __restrict__ v, long ldk,
                                       bf16* out, long ldo, int Tk, int hd, float scale) {
    pdl_prologue();
    extern __shared__ float sm[];  // scores [Tk]
    const int qi = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const bf16* qr = q + (long)qi * ldq + h * hd;
    for (int j = warp; j < Tk; j += nw) {
        const bf16* kr = k + (long)j * ldk + h * hd;
        float s = 0.f;
        for (int d = lane; d < hd; d += 32) s += __bfloat162float(qr[d]) * __bfloat162float(kr[d]);
        s = warp_sum(s);
        if (lane == 0) sm[j] = s * scale;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = 0; j < Tk; j++) mx = fmaxf(mx, sm[j]);
    float den = 0.f;
    for (int j = 0; j < Tk; j++) den += expf(sm[j] - mx);
    for (int d = tid; d < hd; d += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < Tk; j++) acc += expf(sm[j] - mx) * __bfloat162float(v[(long)j * ldk + h * hd + d]);
        out[(long)qi * ldo + h * hd + d] = __float2bfloat16(acc / den);
    }
}

}  // namespace

static inline dim3 g1(long n, int t = 256) { return dim3((unsigned)((n + t - 1) / t)); }

void launch_norm(const NormParams& p, cudaStream_t st) {
    long warps = (long)p.rows * p.batch;
    if (warps == 0) return;
    CBX_REQUIRE(p.C == 256 || p.C == 512 || p.C == 1024, "norm: row width must be 256, 512 or 1024");
    CBX_REQUIRE(p.act == ACT_NONE || p.act == ACT_MISH, "norm: only Mish is fused");
    CBX_REQUIRE(p.ld_in % 4 == 0 && p.in_bs % 4 == 0 && p.ld_outF % 4 == 0 && p.outF_bs % 4 == 0 && p.ld_outB % 4 == 0 && p.outB_bs % 4 == 0 && p.add_bs % 4 == 0,
                "norm: rows must be 16-byte aligned");
    ProfScope ps(PC_NORM, (double)warps * p.C * 6, st);
    const dim3 grid = g1(warps * 32, 256), block(256);
    const bool mish = p.act == ACT_MISH;
#define CBX_NORM(V4) (mish ? launch_pdl(norm_kernel<V4, ACT_MISH>, grid, block, 0, st, p) : launch_pdl(norm_kernel<V4, ACT_NONE>, grid, block, 0, st, p))
    if (p.C == 256) CBX_NORM(2); else if (p.C == 512) CBX_NORM(4); else CBX_NORM(8);
#undef CBX_NORM
}
void launch_gather_rows_bf16(const float* table, const int* idx, int n, int C, bf16* out, long ld, cudaStream_t st) {
    launch_pdl(gather_rows_bf16_kernel, dim3(g1((long)n * C)), dim3(256), 0, st, table, idx, n, C, out, ld);
    CBX_CHECK(cudaGetLastError());
}
void launch_f32_to_bf16_rows(const float* in, long ld_in, bf16* out, long ld_out, int rows, int C, int act, float act_param, cudaStream_t st) {
    ProfScope ps(PC_ELEMWISE, (double)rows * C * 6, st);
    if (rows == 0) return;
    launch_pdl(f32_to_bf16_rows_kernel, dim3(g1((long)rows * C)), dim3(256), 0, st, in, ld_in, out, ld_out, rows, C, act, act_param);
    CBX_CHECK(cudaGetLastError());
}
void launch_f32_to_bf16_slabs(const float* in, long ld_in, long in_bs, bf16* out, long ld_out, long out_bs, int rows, int C, int batch, cudaStream_t st) {
    ProfScope ps(PC_ELEMWISE, (double)rows * C * 6 * batch, st);
    if (rows == 0 || batch == 0) return;
    CBX_REQUIRE(C % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && in_bs % 4 == 0 && out_bs % 4 == 0, "f32_to_bf16_slabs: 4-column alignment");
    launch_pdl(f32_to_bf16_slabs_kernel, dim3(g1((long)rows * C / 4).x, batch), dim3(256), 0, st, in, ld_in, in_bs, out, ld_out, out_bs, rows, C);
    CBX_CHECK(cudaGetLastError());
}
void launch_add_rows(float* a, long lda, const float* b, long ldb, int rows, int C, cudaStream_t st) {
    launch_pdl(add_rows_kernel, dim3(g1((long)rows * C)), dim3(256), 0, st, a, lda, b, ldb, rows, C);
    CBX_CHECK(cudaGetLastError());
}
void launch_upsample2(const float* in, bf16* out, long ld_out, int T, int C, cudaStream_t st) {
    launch_pdl(upsample2_kernel, dim3(g1((long)2 * T * C)), dim3(256), 0, st, in, out, ld_out, T, C);
    CBX_CHECK(cudaGetLastError());
}
void launch_pack_cfm_input(const float* x, const float* mu, const float* spks, const float* cond, bf16* out, long out_bs, int T, int mel, cudaStream_t st) {
    ProfScope ps(PC_ELEMWISE, (double)T * mel * 4 * 2 * 6, st);
    launch_pdl(pack_cfm_input_kernel, dim3(g1((long)2 * T * 4 * mel)), dim3(256), 0, st, x, mu, spks, cond, out, out_bs, T, mel);
    CBX_CHECK(cudaGetLastError());
}
void launch_euler_update(float* x, const float* v, long v_bs, long n, float dt, float r, cudaStream_t st) {
    ProfScope ps(PC_ELEMWISE, (double)n * 16, st);
    launch_pdl(euler_update_kernel, dim3(g1(n)), dim3(256), 0, st, x, v, v_bs, n, dt, r);
    CBX_CHECK(cudaGetLastError());
}
void launch_small_attention(const bf16* q, long ldq, const bf16* k, const bf16* v, long ldk, bf16* out, long ldo, int Tq, int Tk, int H, int hd, float scale, cudaStream_t st) {
    launch_pdl(small_attention_kernel, dim3(Tq, H), dim3(128), Tk * sizeof(float), st, q, ldq, k, v, ldk, out, ldo, Tk, hd, scale);
    CBX_CHECK(cudaGetLastError());
}
