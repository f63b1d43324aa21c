// T3 decode-step kernels (sm_100a): HBM-streaming GEMV over fragment-ordered bf16 weights with
// fused RMSNorm / residual / SwiGLU epilogues, RoPE + paged-KV decode attention, and the one-kernel
// CFG-mix -> repetition-penalty -> temperature -> min-p -> top-p -> sample step.
// Also the prefill helpers (embedding assembly, RoPE + KV page write).
#include <cstdlib>
#include "common.cuh"
#include "t3_kernels.cuh"

namespace {

// This CTA's share of an L2 prefetch of [ptr, ptr + bytes): one 128-byte line per instruction, fire and forget.
__device__ __forceinline__ void l2_prefetch_slice(const void* ptr, size_t bytes, int cta, int n_cta, int tid, int n_thr) {
    const size_t lines = bytes >> 7, per = (lines + n_cta - 1) / n_cta, l0 = (size_t)cta * per, l1 = l0 + per < lines ? l0 + per : lines;
    const char* base = static_cast<const char*>(ptr);
    for (size_t i = l0 + tid; i < l1; i += n_thr) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (i << 7)));
}

// ------------------------------------------------------------------------------------------------
// GEMV: y[r][f] = sum_k W[f][k] * xin[r][k] for r < rows (<= 8*NT).  W is stored in mma.m16n8k16
// A-fragment order: tile (strip s, ktile kt) = 32 lanes x 8 bf16 (512 B, contiguous), tiles ordered
// [strip][ktile], so every warp streams 512-byte coalesced lines straight into tensor-core fragments.
// Each CTA owns `strips_per_cta` strips; its warps split K.
// ------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(512) gemv_kernel(const GemvParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int ldx = p.K + 8;  // bf16 elements per staged row (+8 pad: conflict-free fragment reads)
    bf16* xs = reinterpret_cast<bf16*>(smem);                                   // [8*NT][ldx]
    float* part = reinterpret_cast<float*>(smem + (size_t)8 * NT * ldx * 2);    // [nwarps][S][16][8*NT]
    float* sq = part + (size_t)nwarps * p.strips_per_cta * 16 * 8 * NT;         // [S][16][8*NT] squares of the new residual values
    float* scl = sq + (size_t)p.strips_per_cta * 16 * 8 * NT;                   // [8*NT] RMS scale per row (ss_in)
    int* rmap = reinterpret_cast<int*>(scl + 8 * NT);                           // [8*NT] compact row -> global row
    constexpr int R = 8 * NT;

    // weights and the row map do not depend on the previous kernel: fetch this warp's first batch of fragments and the
    // map before waiting for it
    pdl_launch_dependents();
    uint4 wpre[8];
    {
        const int KTp = p.K >> 4, perp = KTp / nwarps, strip0 = blockIdx.x * p.strips_per_cta;
        const uint4* wp0 = reinterpret_cast<const uint4*>(p.Wf) + ((size_t)strip0 * KTp + warp * perp) * 32 + lane;
#pragma unroll
        for (int u = 0; u < 8; u++) wpre[u] = (strip0 < p.n_strips && u < perp) ? __ldg(wp0 + (size_t)u * 32) : make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < R) rmap[tid] = tid < p.rows ? p.row_map[tid] : 0;
    if (p.pf_ptr) l2_prefetch_slice(p.pf_ptr, p.pf_bytes, blockIdx.x, gridDim.x, tid, blockDim.x);
    __syncthreads();
    pdl_wait();
    // residual epilogue: the old value of this thread's output element is requested now, not after the main loop
    float x_old = 0.f;
    if (p.epi == GEMV_RESID && p.strips_per_cta == 1 && tid < 16 * R) {
        const int f = tid / R, r = tid % R, col = blockIdx.x * 16 + f;
        if (r < p.rows && col < p.N) x_old = p.out[(long)rmap[r] * p.ld_out + col];
    }
    if (p.xb) {
        // ---- bf16 rows written by the producing kernel: 16-byte chunks, eight in flight per thread
        const int nch = p.K >> 3, sh = 31 - __clz(nch), total = R * nch;
        for (int c0 = 0; c0 < total; c0 += blockDim.x * 8) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int c = c0 + u * blockDim.x + tid, r = c >> sh, ch = c & (nch - 1);
                v[u] = (c < total && r < p.rows) ? *reinterpret_cast<const uint4*>(p.xb + (long)rmap[r] * p.ldxb + ch * 8) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int c = c0 + u * blockDim.x + tid, r = c >> sh, ch = c & (nch - 1);
                if (c < total) *reinterpret_cast<uint4*>(xs + (size_t)r * ldx + ch * 8) = v[u];
            }
        }
        if (tid < R) {
            float sc = 1.f;
            if (p.ss_in && tid < p.rows) {
                const float4* sp = reinterpret_cast<const float4*>(p.ss_in + (long)rmap[tid] * p.n_ss);
                float ss = 0.f;
                for (int i = 0; i < p.n_ss / 4; i++) { const float4 t = sp[i]; ss += (t.x + t.y) + (t.z + t.w); }
                sc = rsqrtf(ss / p.K + p.eps);
            }
            scl[tid] = sc;
        }
    } else {
    // ---- stage fp32 input rows (optionally RMS-normalised) as bf16: every thread owns float4 columns of all rows, so the
    // whole staging is one global round trip; per-row sums of squares are reduced through shared memory
        if (tid < R) scl[tid] = 1.f;
        float* red = part;   // [8*NT][nwarps] scratch inside the partial-sum area (free until the main loop ends)
        const int nvec = p.K >> 2;
        for (int c0 = 0; c0 < nvec; c0 += blockDim.x) {
            const int c = c0 + tid;
            float4 v[R];
#pragma unroll
            for (int r = 0; r < R; r++)
                v[r] = (r < p.rows && c < nvec) ? *reinterpret_cast<const float4*>(p.x + (long)rmap[r] * p.ldx_in + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.gain) {   // K == blockDim.x * 4 is not required: accumulate partial sums over the column passes
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float ss = warp_sum(v[r].x * v[r].x + v[r].y * v[r].y + v[r].z * v[r].z + v[r].w * v[r].w);
                    if (lane == 0) red[r * nwarps + warp] = (c0 == 0 ? 0.f : red[r * nwarps + warp]) + ss;
                }
            }
            if (!p.gain) {
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (c < nvec) *reinterpret_cast<uint2*>(xs + (size_t)r * ldx + c * 4) = make_uint2(pack_bf16(v[r].x, v[r].y), pack_bf16(v[r].z, v[r].w));
            }
        }
        if (p.gain) {
            __syncthreads();
            for (int c = tid; c < nvec; c += blockDim.x) {
                const float4 gg = *reinterpret_cast<const float4*>(p.gain + c * 4);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    uint2 o = make_uint2(0u, 0u);
                    if (r < p.rows) {
                        float ss = 0.f;
                        for (int w = 0; w < nwarps; w++) ss += red[r * nwarps + w];
                        const float scale = rsqrtf(ss / p.K + p.eps);
                        const float4 a = *reinterpret_cast<const float4*>(p.x + (long)rmap[r] * p.ldx_in + c * 4);   // L1 hit
                        o = make_uint2(pack_bf16(a.x * scale * gg.x, a.y * scale * gg.y), pack_bf16(a.z * scale * gg.z, a.w * scale * gg.w));
                    }
                    *reinterpret_cast<uint2*>(xs + (size_t)r * ldx + c * 4) = o;
                }
            }
        }
    }
    __syncthreads();

    const int KT = p.K >> 4;
    const int kt_per = KT / nwarps;            // host guarantees divisibility
    const int kt0 = warp * kt_per;
    const int S = p.strips_per_cta;
    for (int sl = 0; sl < S; sl++) {
        const int strip = blockIdx.x * S + sl;
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[j][r] = 0.f;
        if (strip < p.n_strips) {
            const uint4* wp = reinterpret_cast<const uint4*>(p.Wf) + ((size_t)strip * KT + kt0) * 32 + lane;
            for (int kt = 0; kt < kt_per; kt += 8) {
                uint4 w[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (sl == 0 && kt == 0) w[u] = wpre[u];
                    else if (kt + u < kt_per) w[u] = __ldg(wp + (size_t)(kt + u) * 32);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (kt + u < kt_per) {
                        const int k0 = (kt0 + kt + u) << 4;
                        uint32_t a[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
                        for (int j = 0; j < NT; j++) {
                            const bf16* xr = xs + (size_t)(j * 8 + g) * ldx + k0 + tg * 2;
                            uint32_t b0 = *reinterpret_cast<const uint32_t*>(xr);
                            uint32_t b1 = *reinterpret_cast<const uint32_t*>(xr + 8);
                            mma_bf16(acc[j], a, b0, b1);
                        }
                    }
                }
            }
        }
        float* pw = part + ((size_t)(warp * S + sl) * 16) * (8 * NT);
#pragma unroll
        for (int j = 0; j < NT; j++) {
            pw[(g) * (8 * NT) + j * 8 + tg * 2] = acc[j][0];
            pw[(g) * (8 * NT) + j * 8 + tg * 2 + 1] = acc[j][1];
            pw[(g + 8) * (8 * NT) + j * 8 + tg * 2] = acc[j][2];
            pw[(g + 8) * (8 * NT) + j * 8 + tg * 2 + 1] = acc[j][3];
        }
    }
    __syncthreads();
    // ---- cross-warp reduction + epilogue: one thread per (strip_local, f, r)
    const int per_strip = 16 * 8 * NT;
    if (p.epi == GEMV_GLU) {
        // strips come in (gate, up) pairs
        for (int i = tid; i < (S / 2) * per_strip; i += blockDim.x) {
            int pr = i / per_strip, e = i % per_strip, f = e / (8 * NT), r = e % (8 * NT);
            if (r >= p.rows) continue;
            float gsum = 0.f, usum = 0.f;
            for (int w = 0; w < nwarps; w++) {
                gsum += part[((size_t)(w * S + 2 * pr) * 16 + f) * (8 * NT) + r];
                usum += part[((size_t)(w * S + 2 * pr + 1) * 16 + f) * (8 * NT) + r];
            }
            gsum *= scl[r]; usum *= scl[r];
            int pair = (blockIdx.x * S) / 2 + pr;
            if (pair * 2 + 1 < p.n_strips) {
                const float a = gsum / (1.f + expf(-gsum)) * usum;
                if (p.out_b) p.out_b[(long)rmap[r] * p.ld_out_b + pair * 16 + f] = __float2bfloat16(a);
                else p.out[(long)rmap[r] * p.ld_out + pair * 16 + f] = a;
            }
        }
    } else {
        for (int i = tid; i < S * per_strip; i += blockDim.x) {
            int sl = i / per_strip, e = i % per_strip, f = e / (8 * NT), r = e % (8 * NT);
            int strip = blockIdx.x * S + sl;
            if (r >= p.rows || strip >= p.n_strips) continue;
            float s = 0.f;
            for (int w = 0; w < nwarps; w++) s += part[((size_t)(w * S + sl) * 16 + f) * (8 * NT) + r];
            s *= scl[r];
            int col = strip * 16 + f;
            if (col >= p.N) continue;
            float* o = p.out + (long)rmap[r] * p.ld_out + col;
            if (p.epi == GEMV_RESID) {
                const float v = ((S == 1 && i == tid) ? x_old : *o) + s;
                *o = v;
                if (p.out_b) p.out_b[(long)rmap[r] * p.ld_out_b + col] = __float2bfloat16(v * p.next_gain[col]);
                if (p.ss_out) sq[(sl * 16 + f) * (8 * NT) + r] = v * v;
            } else *o = s;
        }
        if (p.epi == GEMV_RESID && p.ss_out) {   // sum of squares of this strip's 16 new values per row, fixed order
            __syncthreads();
            for (int i = tid; i < S * 8 * NT; i += blockDim.x) {
                const int sl = i / (8 * NT), r = i % (8 * NT), strip = blockIdx.x * S + sl;
                if (r >= p.rows || strip >= p.n_strips) continue;
                float ss = 0.f;
#pragma unroll
                for (int f = 0; f < 16; f++) ss += sq[(sl * 16 + f) * (8 * NT) + r];
                p.ss_out[(long)rmap[r] * p.n_strips + strip] = ss;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Decode attention for one (row, head): RoPE on the new q/k, append k/v to the row's KV pages, then
// softmax(q K^T / 8) V over positions [0, pos].  KV pool layout per layer: [2][page][H][16][64] bf16.
// ------------------------------------------------------------------------------------------------
constexpr int PAGE = 16, HD = 64;

// Flash-decoding form: each of the 8 warps streams whole KV pages (K and V of a page are fetched together) and keeps an
// online-softmax partial (m, l, o[64]); the partials are merged in shared memory -- two block barriers and two global
// round trips instead of five and three.
template <bool PIPE, int NW>
__global__ void __launch_bounds__(32 * NW) decode_attn_kernel(const DecodeAttnParams p) {
    __shared__ float qs[HD];
    __shared__ float wm[NW], wl[NW];
    __shared__ float wo[NW * HD];
    const int h = blockIdx.x, r = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the row map and the page table are written by the host between steps only: read them before waiting for the QKV kernel
    pdl_launch_dependents();
    const int row = p.row_map[r];
    const int slot = row >> 1;
    const int* pt = p.page_table + (long)row * p.max_pages;
    int pte0 = warp < p.max_pages ? pt[warp] : 0, pte1 = warp + NW < p.max_pages ? pt[warp + NW] : 0;
    bf16* kpool = p.kv;
    bf16* vpool = p.kv + p.kv_half;
    const int pp = lane >> 1, half = lane & 1;
    // K and V of a page are fetched together; the page after it is in flight while this one is reduced.  The first page of
    // every warp is fetched BEFORE waiting for the QKV kernel: pages below the one that receives the new position were
    // completed a whole step ago (unused table entries are 0, a valid page), so the read is safe and overlaps the
    // previous kernel's tail; it is only used if it turns out not to be the last page.
    uint4 ku[4], vu[4], kn[4], vn[4];
    auto fetch = [&](int pte, uint4* k, uint4* v) {
        const long off = (((long)pte * p.H + h) * PAGE + pp) * HD + half * 32;
        const uint4* kp = reinterpret_cast<const uint4*>(kpool + off);
        const uint4* vp = reinterpret_cast<const uint4*>(vpool + off);
#pragma unroll
        for (int c = 0; c < 4; c++) { k[c] = kp[c]; v[c] = vp[c]; }
    };
    if (PIPE) fetch(pte0, ku, vu);
    if (p.pf_ptr) l2_prefetch_slice(p.pf_ptr, p.pf_bytes, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, tid, blockDim.x);
    pdl_wait();
    const int pos = p.slot_pos[slot];
    const float* qkv = p.qkv + (long)row * (3 * p.H * HD);
    if (tid < 32) {
        // rotate_half RoPE: pairs (d, d+32)
        float sn, cs;
        sincosf((float)pos * p.inv_freq[tid], &sn, &cs);
        const float q0 = qkv[h * HD + tid], q1 = qkv[h * HD + tid + 32];
        const float k0 = qkv[p.H * HD + h * HD + tid], k1 = qkv[p.H * HD + h * HD + tid + 32];
        qs[tid] = (q0 * cs - q1 * sn) * 0.125f;
        qs[tid + 32] = (q1 * cs + q0 * sn) * 0.125f;
        const long base = (((long)pt[pos / PAGE] * p.H + h) * PAGE + (pos % PAGE)) * HD;
        kpool[base + tid] = __float2bfloat16(k0 * cs - k1 * sn);
        kpool[base + tid + 32] = __float2bfloat16(k1 * cs + k0 * sn);
        vpool[base + tid] = __float2bfloat16(qkv[2 * p.H * HD + h * HD + tid]);
        vpool[base + tid + 32] = __float2bfloat16(qkv[2 * p.H * HD + h * HD + tid + 32]);
    }
    const int n = pos + 1, npages = (n + PAGE - 1) / PAGE;
    const bool early = PIPE && warp < npages - 1;
    __syncthreads();
    if (p.q_save && tid < HD) p.q_save[(long)row * (p.H * HD) + h * HD + tid] = qs[tid];
    if (!early && warp < npages) fetch(pte0, ku, vu);
    float m = -INFINITY, lsum = 0.f;
    float acc[32];
#pragma unroll
    for (int d = 0; d < 32; d++) acc[d] = 0.f;
    for (int pg = warp; pg < npages; pg += NW) {
        const int nx = pg + NW;
        int pte2 = 0;
        if (PIPE) {
            if (nx < npages) fetch(pte1, kn, vn);
            pte2 = nx + NW < npages ? pt[nx + NW] : 0;
        } else if (pg != warp) fetch(pt[pg], ku, vu);
        float sc = 0.f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&ku[c]);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __bfloat1622float2(b2[e]);
                sc += f.x * qs[half * 32 + c * 8 + e * 2] + f.y * qs[half * 32 + c * 8 + e * 2 + 1];
            }
        }
        sc += __shfl_xor_sync(0xffffffffu, sc, 1);
        if (pg * PAGE + pp >= n) sc = -INFINITY;           // stale tail of the last page
        const float mnew = fmaxf(m, warp_max(sc));         // finite: every page holds at least one valid position
        const float corr = expf(m - mnew), pj = expf(sc - mnew);
        m = mnew;
        lsum = lsum * corr + pj;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&vu[c]);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __bfloat1622float2(b2[e]);
                acc[c * 8 + e * 2] = acc[c * 8 + e * 2] * corr + (pj > 0.f ? pj * f.x : 0.f);
                acc[c * 8 + e * 2 + 1] = acc[c * 8 + e * 2 + 1] * corr + (pj > 0.f ? pj * f.y : 0.f);
            }
        }
        if (PIPE) {
#pragma unroll
            for (int c = 0; c < 4; c++) { ku[c] = kn[c]; vu[c] = vn[c]; }
            pte1 = pte2;
        }
    }
#pragma unroll
    for (int d = 0; d < 32; d++) {
        float v = acc[d];
        v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[d] = v;
    }
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 2); lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 8); lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
    if (lane < 2) {
#pragma unroll
        for (int d = 0; d < 32; d++) wo[warp * HD + lane * 32 + d] = acc[d];
        if (lane == 0) { wm[warp] = m; wl[warp] = lsum; }
    }
    __syncthreads();
    if (tid < HD) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; w++) M = fmaxf(M, wm[w]);
        float Lt = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const float e = wm[w] == -INFINITY ? 0.f : expf(wm[w] - M);
            Lt += wl[w] * e; O += wo[w * HD + tid] * e;
        }
        if (p.out_b) p.out_b[(long)row * (p.H * HD) + h * HD + tid] = __float2bfloat16(O / Lt);
        else p.out[(long)row * (p.H * HD) + h * HD + tid] = O / Lt;
    }
}

// ------------------------------------------------------------------------------------------------
// Sampler: one CTA (1024 threads) per stream.  Order fixed by BASELINE.json north_star:
// CFG mix -> repetition penalty -> temperature -> min-p -> top-p -> sample (argmax p/q, q~Exp(1)).
// Also advances the stream state and writes the next input embedding for both CFG rows.
// ------------------------------------------------------------------------------------------------
constexpr int SAMP_T = 1024, SAMP_E = 9;  // 9 * 1024 >= 8194

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[threadIdx.x & 31];
    r = warp_max(r);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[threadIdx.x & 31];
    r = warp_sum(r);
    __syncthreads();
    return r;
}

// The bf16 hand-over form of a new input row (both CFG rows of `slot`): bf16(v * gain0[d]) and the sum of squares of every 16
// columns (16 consecutive lanes).  Called by whole warps.
__device__ __forceinline__ void write_handover(bf16* xb, float* ss, const float* gain0, int slot, int d, int dim, float v) {
    const bf16 b = __float2bfloat16(v * gain0[d]);
    xb[(long)(slot * 2) * dim + d] = b;
    xb[(long)(slot * 2 + 1) * dim + d] = b;
    float s = v * v;
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4); s += __shfl_xor_sync(0xffffffffu, s, 8);
    if ((d & 15) == 0) {
        ss[(long)(slot * 2) * (dim / 16) + (d >> 4)] = s;
        ss[(long)(slot * 2 + 1) * (dim / 16) + (d >> 4)] = s;
    }
}

__global__ void __launch_bounds__(SAMP_T) sampler_kernel(const SamplerParams p) {
    pdl_prologue();
    __shared__ float red[32];
    __shared__ float wsum2[2 * 32];
    __shared__ float bestv[32];
    __shared__ int besti[32];
    const int slot = p.slots[blockIdx.x], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T3SlotState* st = p.state + slot;
    if (st->done) return;
    const int V = p.V;
    const float* lc = p.logits + (long)(slot * 2) * p.ld_logits;
    const float* lu = lc + p.ld_logits;
    const float w = st->cfg_w, inv_temp = 1.f / st->temp, rp = st->rep_pen;
    const uint8_t* seen = p.seen + (long)slot * p.seen_stride;
    const int ctl = p.eos_ctl ? p.eos_ctl[slot] : 0;
    float l[SAMP_E];
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < SAMP_E; e++) {
        int i = tid + e * SAMP_T;
        float v = -INFINITY;
        if (i < V) {
            float c = lc[i];
            v = w > 0.f ? c + w * (c - lu[i]) : c;
            if (ctl & 2) v = i == p.eos ? 32768.f : -32768.f;      // forced end of speech: every logit -2^15, EOS +2^15
            if ((ctl & 1) && i == p.eos) v = -32768.f;             // EOS suppressed while the text is not finished
            if (seen[i]) v = v < 0.f ? v * rp : v / rp;
            v *= inv_temp;
        }
        l[e] = v;
        mx = fmaxf(mx, v);
    }
    mx = block_reduce_max(mx, red);
    float ex[SAMP_E];
    float z = 0.f;
#pragma unroll
    for (int e = 0; e < SAMP_E; e++) { ex[e] = expf(l[e] - mx); z += ex[e]; }
    z = block_reduce_sum(z, red);
    // min-p: drop p_i < min_p * p_max, p_max = 1/z
    if (st->min_p > 0.f) {
        const float thr = st->min_p * (1.f / z);
#pragma unroll
        for (int e = 0; e < SAMP_E; e++)
            if (ex[e] / z < thr && ex[e] != 1.0f) ex[e] = 0.f;
    }
    // top-p: remove the ascending-probability prefix whose cumulative mass <= 1 - top_p
    if (st->top_p < 1.f) {
        float z2 = 0.f;
#pragma unroll
        for (int e = 0; e < SAMP_E; e++) z2 += ex[e];
        z2 = block_reduce_sum(z2, red);
        const float target = (1.f - st->top_p) * z2;
        // bisection on the float bit pattern: invariant mass(e <= lo) <= target < mass(e <= hi).  The whole block sits on one
        // SM, so the search is bound by instruction issue, not latency: one threshold per round over the nine register-resident
        // values of every thread (~30 rounds x ~60 instructions per warp) costs a fifth of evaluating 32 thresholds per round.
        // Reduction order is fixed (thread: e ascending; warp: xor 16..1; block: the 32 warp totals by xor 16..1), every thread
        // takes the same decision, one block barrier per round (double-buffered warp totals).
        unsigned int lo = 0u, hi = 0x3F800000u;
        int round = 0;
        while (hi - lo > 1u) {
            const unsigned int t = lo + ((hi - lo) >> 1);
            const float thr = __uint_as_float(t);
            float m = 0.f;
#pragma unroll
            for (int e = 0; e < SAMP_E; e++) m += (ex[e] <= thr) ? ex[e] : 0.f;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
            float* buf = wsum2 + (round & 1) * 32;
            if (lane == 0) buf[warp] = m;
            __syncthreads();
            float tot = buf[lane];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (tot <= target) lo = t; else hi = t;
            round++;
        }
        const float cut = __uint_as_float(lo);
#pragma unroll
        for (int e = 0; e < SAMP_E; e++)
            if (ex[e] <= cut) ex[e] = 0.f;
    }
    // sample: argmax e_i / q_i  (normalisation-free form of multinomial(softmax))
    float bv = -1.f; int bi = 0x7fffffff;
#pragma unroll
    for (int e = 0; e < SAMP_E; e++) {
        int i = tid + e * SAMP_T;
        if (i < V && ex[e] > 0.f) {
            float q;
            if (p.noise) q = p.noise[(long)blockIdx.x * p.noise_stride + i];
            else {
                uint32_t r4[4];
                Philox::gen(st->seed, (uint32_t)i, (uint32_t)st->step, 0x5A4Du, 0u, r4);
                q = -logf(u32_to_unit(r4[0]));
            }
            float val = ex[e] / q;
            if (val > bv || (val == bv && i < bi)) { bv = val; bi = i; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { bestv[warp] = bv; besti[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bv = bestv[lane]; bi = besti[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) besti[0] = bi;
    }
    __syncthreads();
    const int tok = besti[0];
    const int step = st->step;
    // next input embedding (both CFG rows): speech_emb[tok] + speech_pos[step+1]
    for (int d = tid; d < p.dim; d += SAMP_T) {      // dim is a multiple of 32: whole warps
        float v = p.speech_emb[(long)tok * p.dim + d] + p.speech_pos[(long)(step + 1) * p.dim + d];
        p.x[(long)(slot * 2) * p.dim + d] = v;
        p.x[(long)(slot * 2 + 1) * p.dim + d] = v;
        if (p.xb) write_handover(p.xb, p.ss, p.gain0, slot, d, p.dim, v);
    }
    __syncthreads();
    if (tid == 0) {
        p.out_tokens[(long)slot * p.out_stride + step] = tok;
        p.seen[(long)slot * p.seen_stride + tok] = 1;
        st->step = step + 1;
        st->pos = st->pos + 1;
        if (tok == p.eos || step + 1 >= st->max_new) st->done = 1;
        p.slot_pos[slot] = st->pos;
    }
}

// ------------------------------------------------------------------------------------------------
// Prefill helpers
// ------------------------------------------------------------------------------------------------
// x[b][t][:] for t < Lp: [cond prefix (Lc) | text_emb[tok]+text_pos[i] (row 1: pos only when cfg>0) | bos ...]
__global__ void assemble_embeds_kernel(const AssembleParams p) {
    pdl_prologue();
    const int t = blockIdx.x, b = blockIdx.y;
    float* out = p.x + ((long)b * (p.slab ? p.slab : p.Lp) + t) * p.dim;
    for (int d = threadIdx.x; d < p.dim; d += blockDim.x) {
        float v;
        if (t < p.Lc) v = p.prefix[(long)t * p.dim + d];
        else if (t < p.Lc + p.L) {
            int i = t - p.Lc;
            v = p.text_pos[(long)i * p.dim + d];
            if (!(b == 1 && p.cfg_on)) v += p.text_emb[(long)p.text_ids[i] * p.dim + d];
        } else v = p.speech_emb[(long)p.bos * p.dim + d] + p.speech_pos[d];
        out[d] = v;
    }
}

// rotate q,k in the fused qkv buffer (bf16 [2][Lp][3*H*64]) and write k,v into the rows' KV pages
__global__ void rope_kv_prefill_kernel(const RopeKvParams p) {
    pdl_prologue();
    const int t = blockIdx.x, b = blockIdx.y;
    bf16* row = p.qkv + ((long)b * (p.slab ? p.slab : p.Lp) + t) * (3 * p.H * HD);
    const int* pt = p.page_table + (long)(p.row0 + b) * p.max_pages;
    bf16* kpool = p.kv; bf16* vpool = p.kv + p.kv_half;
    for (int i = threadIdx.x; i < p.H * 32; i += blockDim.x) {
        int h = i >> 5, d = i & 31;
        float sn, cs;
        sincosf((float)t * p.inv_freq[d], &sn, &cs);
        bf16* q = row + h * HD; bf16* k = row + p.H * HD + h * HD; bf16* v = row + 2 * p.H * HD + h * HD;
        float q0 = __bfloat162float(q[d]), q1 = __bfloat162float(q[d + 32]);
        float k0 = __bfloat162float(k[d]), k1 = __bfloat162float(k[d + 32]);
        bf16 kr0 = __float2bfloat16(k0 * cs - k1 * sn), kr1 = __float2bfloat16(k1 * cs + k0 * sn);
        q[d] = __float2bfloat16(q0 * cs - q1 * sn); q[d + 32] = __float2bfloat16(q1 * cs + q0 * sn);
        k[d] = kr0; k[d + 32] = kr1;
        long base = (((long)pt[t / PAGE] * p.H + h) * PAGE + (t % PAGE)) * HD;
        kpool[base + d] = kr0; kpool[base + d + 32] = kr1;
        vpool[base + d] = v[d]; vpool[base + d + 32] = v[d + 32];
    }
}

__global__ void init_slot_kernel(T3SlotState* st, T3SlotState v, int* slot_pos, int slot, uint8_t* seen, int seen_stride, int bos,
                                 float* x, const float* speech_emb, const float* speech_pos, int dim, bf16* xb, float* ss, const float* gain0) {
    pdl_prologue();
    for (int i = threadIdx.x; i < seen_stride; i += blockDim.x) seen[(long)slot * seen_stride + i] = (i == bos) ? 1 : 0;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        float e = speech_emb[(long)bos * dim + d] + speech_pos[d];
        x[(long)(slot * 2) * dim + d] = e;
        x[(long)(slot * 2 + 1) * dim + d] = e;
        if (xb) write_handover(xb, ss, gain0, slot, d, dim, e);
    }
    if (threadIdx.x == 0) { st[slot] = v; slot_pos[slot] = v.pos; }
}

__global__ void prompt_embed_kernel(float* out, const float* __restrict__ emb, const float* __restrict__ pos, const int* __restrict__ ids, int n, int D) {
    pdl_prologue();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n * D) return;
    int r = i / D, c = i % D;
    out[i] = emb[(long)ids[r] * D + c] + pos[(long)r * D + c];
}
__global__ void scale_vec_kernel(float* out, const float* __restrict__ w, float s, int n) {
    pdl_prologue();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = w[i] * s;
}

// ------------------------------------------------------------------------------------------------
// Alignment-based EOS control
// ------------------------------------------------------------------------------------------------
// Attention probabilities of the conditional row's newest query over the text span at one layer, averaged over the heads
// (the analyzer's `step_attention[0].mean(0)` rows).  One CTA per stream, one warp per head: pass 1 reduces the softmax
// maximum and denominator over every cached position, pass 2 evaluates the text columns (heads summed in ascending order).
__global__ void __launch_bounds__(512) align_attn_kernel(const AlignAttnParams p) {
    __shared__ __align__(16) float qs[16 * HD];
    __shared__ float hm[16], hl[16];
    pdl_prologue();
    const int slot = p.slots ? p.slots[blockIdx.x] : p.slot;
    const AlignState as = p.state[slot];
    if (!as.on) return;
    const int tid = threadIdx.x, lane = tid & 31, h = tid >> 5, row = slot * 2;
    const int pos = p.q_b ? p.pos : p.slot_pos[slot];
    const int* pt = p.page_table + (long)row * p.max_pages;
    if (p.q_b) {
        qs[h * HD + lane] = __bfloat162float(p.q_b[h * HD + lane]) * 0.125f;
        qs[h * HD + lane + 32] = __bfloat162float(p.q_b[h * HD + lane + 32]) * 0.125f;
    } else {
        const float* q = p.q_rot + (long)row * (p.H * HD) + h * HD;
        qs[h * HD + lane] = q[lane];
        qs[h * HD + lane + 32] = q[lane + 32];
    }
    __syncthreads();
    auto score = [&](int hh, int pp) {
        const uint4* kp = reinterpret_cast<const uint4*>(p.kv + (((long)pt[pp / PAGE] * p.H + hh) * PAGE + (pp % PAGE)) * HD);
        const float* qh = qs + hh * HD;
        float sc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint4 u = kp[c];
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __bfloat1622float2(b2[e]);
                sc += f.x * qh[c * 8 + e * 2] + f.y * qh[c * 8 + e * 2 + 1];
            }
        }
        return sc;
    };
    float m = -INFINITY, l = 0.f;
    for (int pp = lane; pp <= pos; pp += 32) {
        const float sc = score(h, pp), mn = fmaxf(m, sc);
        l = l * expf(m - mn) + expf(sc - mn);
        m = mn;
    }
    const float M = warp_max(m);
    l = warp_sum(m == -INFINITY ? 0.f : l * expf(m - M));
    if (lane == 0) { hm[h] = M; hl[h] = l; }
    __syncthreads();
    float* out = p.out + (long)slot * p.ld_out;
    for (int t = tid; t < as.S; t += blockDim.x) {
        float acc = 0.f;
        for (int hh = 0; hh < 16; hh++) acc += expf(score(hh, as.i0 + t) - hm[hh]) / hl[hh];
        out[t] = acc * 0.0625f;
    }
}

// The analyzer's per-frame decision on the row(s) align_attn_kernel produced.  Quantities the analyzer recomputes from the whole
// alignment matrix every frame are kept as running values (row order = accumulation order).
__global__ void __launch_bounds__(256) align_step_kernel(const AlignStepParams p) {
    __shared__ float rv[8], rr[8];
    __shared__ int ri[8];
    pdl_prologue();
    const int slot = p.slots[blockIdx.x], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    AlignState s = p.state[slot];
    if (!s.on || (p.t3 && p.t3[slot].done)) return;
    const int S = s.S, fp = s.frame_pos;
    const float* cur = p.a_cur + (long)slot * p.ld;
    // masked row: columns above frame_pos count as 0.  Argmax (first maximum) and the maximum over columns [0, S - 5).
    float bv = -1.f, rmax = 0.f; int bi = 0x7fffffff;
    for (int c = tid; c < S; c += blockDim.x) {
        const float v = c <= fp ? cur[c] : 0.f;
        if (v > bv) { bv = v; bi = c; }
        if (c < S - 5) rmax = fmaxf(rmax, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    }
    if (lane == 0) { rv[warp] = bv; ri[warp] = bi; rr[warp] = rmax; }
    __syncthreads();
    if (tid != 0) return;
    for (int w = 1; w < 8; w++) {
        if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
        rmax = fmaxf(rmax, rr[w]);
    }
    auto at = [&](const float* r, int c) { return (c >= 0 && c <= fp) ? r[c] : 0.f; };
    const int n4 = S < 4 ? S : 4;
    float last2 = 0.f;       // max over the last two rows of the last two columns
    bool have_prev = s.T > 0;
    if (s.has_pre) {         // the first chunk holds two rows: the prefilled BOS row and this one, masked alike
        const float* pre = p.a_pre + (long)slot * p.ld;
        for (int c = 0; c < n4; c++) s.first4_max = fmaxf(s.first4_max, at(pre, c));
        s.prev_last2 = fmaxf(at(pre, S - 2), at(pre, S - 1));
        s.T += 1; s.has_pre = 0; have_prev = true;
    }
    for (int c = 0; c < n4; c++) s.first4_max = fmaxf(s.first4_max, at(cur, c));
    const float cur_last2 = fmaxf(at(cur, S - 2), at(cur, S - 1));
    last2 = have_prev ? fmaxf(s.prev_last2, cur_last2) : cur_last2;
    s.prev_last2 = cur_last2;
    s.T += 1;
    const int posn = bi;
    const int d = posn - s.text_pos;
    if (-4 < d && d < 7) s.text_pos = posn;
    const bool false_start = !s.started && (last2 > 0.1f || s.first4_max < 0.5f);
    s.started = !false_start;
    if (s.started && s.started_at < 0) s.started_at = s.T;
    const bool was_complete = s.complete && s.completed_at >= 0;
    s.complete = s.complete || s.text_pos >= S - 3;
    if (s.complete && s.completed_at < 0) s.completed_at = s.T;
    if (was_complete) {      // this row's index (T - 1) is >= completed_at
        for (int c = 0; c < 3; c++) s.tail3[c] += at(cur, S - 3 + c);
        s.rep_sum += rmax;
    }
    const bool long_tail = s.complete && fmaxf(s.tail3[0], fmaxf(s.tail3[1], s.tail3[2])) >= 10.f;
    const bool repetition = s.complete && s.rep_sum > 5.f;
    s.ctl = ((long_tail || repetition) ? 2 : 0) | (posn < S - 3 ? 1 : 0);
    s.cur_posn = posn;
    s.frame_pos = fp + 1;
    p.state[slot] = s;
    p.ctl[slot] = s.ctl;
}

__global__ void align_init_kernel(AlignState* st, int* ctl, int slot, int on, int i0, int S, int has_pre) {
    pdl_prologue();
    AlignState s{};
    s.on = on; s.i0 = i0; s.S = S; s.has_pre = has_pre; s.started_at = -1; s.completed_at = -1;
    st[slot] = s;
    ctl[slot] = 0;
}

}  // namespace

void t3_kernels_init() {
    gemv_tc_init();
    CBX_CHECK(cudaFuncSetAttribute(gemv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CBX_CHECK(cudaFuncSetAttribute(gemv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CBX_CHECK(cudaFuncSetAttribute(gemv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}
void launch_prompt_embed(float* out, const float* emb, const float* pos, const int* ids, int n, int D, cudaStream_t st) {
    launch_pdl(prompt_embed_kernel, dim3(cdiv((long)n * D, 256)), dim3(256), 0, st, out, emb, pos, ids, n, D);
    CBX_CHECK(cudaGetLastError());
}
void launch_scale_vec(float* out, const float* w, float s, int n, cudaStream_t st) {
    launch_pdl(scale_vec_kernel, dim3(cdiv(n, 256)), dim3(256), 0, st, out, w, s, n);
    CBX_CHECK(cudaGetLastError());
}

void launch_gemv(const GemvParams& p, int nwarps, cudaStream_t st) {
    CBX_REQUIRE(p.rows >= 1 && p.rows <= 32, "gemv: rows must be in [1,32]");
    CBX_REQUIRE(p.K % 16 == 0 && (p.K / 16) % nwarps == 0, "gemv: K/16 must divide by the warp count");
    const int NT = p.rows <= 8 ? 1 : p.rows <= 16 ? 2 : 4;
    const int S = p.strips_per_cta;
    size_t smem = (size_t)8 * NT * (p.K + 8) * 2 + (size_t)(nwarps + 1) * S * 16 * 8 * NT * 4 + 2 * 8 * NT * 4;
    if (NT == 4 && smem > 200 * 1024) {
        // 32 staged rows of K = 4096 (the down projection) do not fit in shared memory: two passes of 16 rows.  The second
        // pass finds the weights in L2 (8 MB against 126 MB), so HBM still streams them once.
        CBX_REQUIRE(p.epi != GEMV_STORE || p.out, "gemv: split pass needs row-addressed outputs");
        GemvParams a = p, b = p;
        a.rows = 16; b.rows = p.rows - 16; b.row_map = p.row_map + 16;
        launch_gemv(a, nwarps, st);
        launch_gemv(b, nwarps, st);
        return;
    }
    int grid = cdiv(p.n_strips, S);
    ProfScope ps(PC_GEMV, (double)p.n_strips * 16 * p.K * 2 + (double)p.rows * (p.K + p.N) * 4, st);
    CBX_REQUIRE(smem <= 200 * 1024, "gemv: staging exceeds shared memory");
    if (NT == 1) launch_pdl(gemv_kernel<1>, dim3(grid), dim3(nwarps * 32), smem, st, p);
    else if (NT == 2) launch_pdl(gemv_kernel<2>, dim3(grid), dim3(nwarps * 32), smem, st, p);
    else launch_pdl(gemv_kernel<4>, dim3(grid), dim3(nwarps * 32), smem, st, p);
    CBX_CHECK(cudaGetLastError());
}
void launch_decode_attn(const DecodeAttnParams& p, int rows, int max_pos, cudaStream_t st) {
    ProfScope ps(PC_DECODE_ATTN, 0.0, st);
    // prefetching pays while the step is latency-bound (0.78 vs 0.81 ms at 2 rows, 0.82 vs 0.89 ms at 8); at 16 rows the
    // speculative reads compete with the next projection's weight stream (1.08 vs 0.99 ms).  Both instances compute
    // bit-identical results.
    static const int pipe_env = [] { const char* v = getenv("CBX_T3_ATTN_PIPE"); return v ? atoi(v) : -1; }();
    const bool pipe = pipe_env >= 0 ? pipe_env != 0 : rows <= 8;
    // 122 registers x 256 threads = two CTAs per SM = 296 resident: from 19 rows (304 CTAs) the grid would take a second wave
    // (measured: 18 rows 1.23 ms per step, 20 rows 1.41).  Beyond that, four-warp CTAs (four per SM, 592 resident) keep every
    // (row, head) resident at once; each warp walks twice the pages.  The merge order over warps differs between the two shapes,
    // so the choice depends on the batch's row count only through this fixed threshold (results for <= 18 rows are unchanged).
    static const int nw4_rows = [] { const char* v = getenv("CBX_T3_ATTN_NW4_ROWS"); return v ? atoi(v) : 19; }();
    if (rows >= nw4_rows) launch_pdl(decode_attn_kernel<false, 4>, dim3(p.H, rows), dim3(128), 0, st, p);
    else if (pipe) launch_pdl(decode_attn_kernel<true, 8>, dim3(p.H, rows), dim3(256), 0, st, p);
    else launch_pdl(decode_attn_kernel<false, 8>, dim3(p.H, rows), dim3(256), 0, st, p);
    CBX_CHECK(cudaGetLastError());
}
void launch_sampler(const SamplerParams& p, int n_streams, cudaStream_t st) {
    CBX_REQUIRE(p.V <= SAMP_T * SAMP_E, "sampler: vocabulary too large for the register tile");
    ProfScope ps(PC_SAMPLER, (double)n_streams * 2 * p.V * 4, st);
    launch_pdl(sampler_kernel, dim3(n_streams), dim3(SAMP_T), 0, st, p);
    CBX_CHECK(cudaGetLastError());
}
void launch_assemble_embeds(const AssembleParams& p, cudaStream_t st) {
    launch_pdl(assemble_embeds_kernel, dim3(p.Lp, 2), dim3(256), 0, st, p);
    CBX_CHECK(cudaGetLastError());
}
void launch_rope_kv_prefill(const RopeKvParams& p, cudaStream_t st) {
    launch_pdl(rope_kv_prefill_kernel, dim3(p.Lp, 2), dim3(256), 0, st, p);
    CBX_CHECK(cudaGetLastError());
}
void launch_init_slot(T3SlotState* st_dev, const T3SlotState& v, int* slot_pos, int slot, uint8_t* seen, int seen_stride, int bos,
                      float* x, const float* speech_emb, const float* speech_pos, int dim, bf16* xb, float* ss, const float* gain0, cudaStream_t st) {
    CBX_REQUIRE(dim % 32 == 0, "init_slot: the hand-over rows are written by whole warps");
    launch_pdl(init_slot_kernel, dim3(1), dim3(256), 0, st, st_dev, v, slot_pos, slot, seen, seen_stride, bos, x, speech_emb, speech_pos, dim, xb, ss, gain0);
    CBX_CHECK(cudaGetLastError());
}
void launch_align_attn(const AlignAttnParams& p, int n_streams, cudaStream_t st) {
    CBX_REQUIRE(p.H == 16, "align: one warp per head of a 16-head trunk");
    launch_pdl(align_attn_kernel, dim3(n_streams), dim3(512), 0, st, p);
    CBX_CHECK(cudaGetLastError());
}
void launch_align_step(const AlignStepParams& p, int n_streams, cudaStream_t st) {
    launch_pdl(align_step_kernel, dim3(n_streams), dim3(256), 0, st, p);
    CBX_CHECK(cudaGetLastError());
}
void launch_align_init(AlignState* st_dev, int* ctl, int slot, int on, int i0, int S, int has_pre, cudaStream_t st) {
    launch_pdl(align_init_kernel, dim3(1), dim3(1), 0, st, st_dev, ctl, slot, on, i0, S, has_pre);
    CBX_CHECK(cudaGetLastError());
}
