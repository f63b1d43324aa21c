// T3 decode step as ONE persistent cooperative kernel (sm_100a): 148 CTAs, each with a producer warp that streams
// its share of the step's 1.02 GB of fragment-ordered bf16 weights through a 5 x 32 KB shared-memory ring with
// 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx), and 8 consumer warps that run the layer phases
//   QKV (RMSNorm fused) | RoPE + paged-KV attention | O-proj | gate/up + SwiGLU (RMSNorm fused) | down
// separated by grid barriers.  The weight stream never waits for a barrier (weights do not depend on
// activations), so HBM stays busy while the dependent part of a phase (barrier, activation reload, a few
// mma.m16n8k16, store) is in flight.  Work items (16-row strips or K-quarters of strips) are dealt round-robin
// over the whole step so every CTA streams the same number of bytes; K-split partial sums are reduced
// deterministically by the next phase's prologue (no atomics).
#include <cooperative_groups.h>
#include "common.cuh"
#include "t3_kernels.cuh"

namespace {

constexpr int D = 1024, FFN = 4096, H = 16, HD = 64, PAGE = 16;
constexpr int SLOT = 32768, NSLOTS = 5, CONS = 256, THREADS = 288, LDX = D + 8;
__device__ unsigned long long g_mega_trace[64];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define MTRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && l == 1) g_mega_trace[i] = gtime(); } while (0)
constexpr int QKV_ITEMS = 3 * D / 16, OP_ITEMS = (D / 16) * 4, GU_ITEMS = 2 * FFN / 32, DN_ITEMS = (D / 16) * 4;

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (long spin = 0; spin < (1L << 28); spin++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct Ring {
    uint32_t base, full0, empty0; int slot; uint32_t phase;
    __device__ void advance() { if (++slot == NSLOTS) { slot = 0; phase ^= 1; } }
};

// items of a phase dealt round-robin: item i belongs to CTA (i + off) % G
struct Deal {
    int off, G, cta;
    __device__ int first() const { int f = cta - off; return f < 0 ? f + G : f; }
    __device__ void next_phase(int n_items) { off = (off + n_items) % G; }
};

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
    cons_sync();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int v;
        long spin = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (++spin > (1L << 28)) __trap();
        } while (v < target);
    }
    cons_sync();
}

// ------------------------------------------------------------------------------------------------ producer
__device__ void producer(const MegaParams& p, Ring ring, Deal deal) {
    auto push = [&](const void* src, uint32_t bytes) {
        const uint32_t full = ring.full0 + 8 * ring.slot, empty = ring.empty0 + 8 * ring.slot;
        mbar_wait(empty, ring.phase ^ 1);
        mbar_expect_tx(full, bytes);
        bulk_load(ring.base + ring.slot * SLOT, src, bytes, full);
        ring.advance();
    };
    for (int l = 0; l < p.n_layers; l++) {
        const MegaLayer& L = p.layers[l];
        for (int i = deal.first(); i < QKV_ITEMS; i += deal.G) push(reinterpret_cast<const char*>(L.wqkv_f) + (size_t)i * SLOT, SLOT);
        deal.next_phase(QKV_ITEMS);
        for (int i = deal.first(); i < OP_ITEMS; i += deal.G) push(reinterpret_cast<const char*>(L.wo_f) + (size_t)i * (SLOT / 4), SLOT / 4);
        deal.next_phase(OP_ITEMS);
        for (int i = deal.first(); i < GU_ITEMS; i += deal.G) {
            push(reinterpret_cast<const char*>(L.wgu_f) + (size_t)(2 * i) * SLOT, SLOT);
            push(reinterpret_cast<const char*>(L.wgu_f) + (size_t)(2 * i + 1) * SLOT, SLOT);
        }
        deal.next_phase(GU_ITEMS);
        for (int i = deal.first(); i < DN_ITEMS; i += deal.G) push(reinterpret_cast<const char*>(L.wd_f) + (size_t)i * SLOT, SLOT);
        deal.next_phase(DN_ITEMS);
    }
    for (int i = deal.first(); i < p.head_items; i += deal.G) push(reinterpret_cast<const char*>(p.head_f) + (size_t)i * SLOT, SLOT);
}

// ------------------------------------------------------------------------------------------------ consumer pieces
// xs[r][k] (bf16) <- normalised (xin[r] + sum of the 4 K-quarter partial rows); the writer CTA of row r also stores the
// summed row to xout.  All 256 consumer threads take one float4 column of every row, so the loads of a phase are one
// L2 round trip; per-row sum of squares is reduced through shared memory (red: [16][8]).
template <int NT, bool PARTS>
__device__ void stage_norm_rows(const MegaParams& p, bf16* xs, const float* xin, const float* part, float* xout, const float* gain, int cta, int G, float* red) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = tid * 4;
    constexpr int R = 8 * NT;
    float4 v[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (r < p.rows) {
            const int row = p.row_map[r];
            float4 a = __ldcg(reinterpret_cast<const float4*>(xin + (size_t)row * D + k));
            if (PARTS) {
                float4 b[4];
#pragma unroll
                for (int q = 0; q < 4; q++) b[q] = __ldcg(reinterpret_cast<const float4*>(part + ((size_t)q * p.rows_total + row) * D + k));
#pragma unroll
                for (int q = 0; q < 4; q++) { a.x += b[q].x; a.y += b[q].y; a.z += b[q].z; a.w += b[q].w; }
            }
            v[r] = a;
        } else v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        float ss = v[r].x * v[r].x + v[r].y * v[r].y + v[r].z * v[r].z + v[r].w * v[r].w;
        ss = warp_sum(ss);
        if (lane == 0) red[r * 8 + warp] = ss;
    }
    cons_sync();
    const float4 gg = *reinterpret_cast<const float4*>(gain + k);
#pragma unroll
    for (int r = 0; r < R; r++) {
        bf16* dst = xs + (size_t)r * LDX + k;
        if (r < p.rows) {
            float ss = 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) ss += red[r * 8 + w];
            const float scale = rsqrtf(ss / D + p.eps);
            const int row = p.row_map[r];
            if (xout && (row % G) == cta) *reinterpret_cast<float4*>(xout + (size_t)row * D + k) = v[r];
            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(v[r].x * scale * gg.x, v[r].y * scale * gg.y), pack_bf16(v[r].z * scale * gg.z, v[r].w * scale * gg.w));
        } else *reinterpret_cast<uint2*>(dst) = make_uint2(0u, 0u);
    }
    cons_sync();
}

// xs[r][0..1024) <- src[row][k0 .. k0+1024) converted to bf16 (no norm); one float4 column per thread
template <int NT>
__device__ void stage_plain_rows(const MegaParams& p, bf16* xs, const float* src, int ld, int k0) {
    const int k = threadIdx.x * 4;
    constexpr int R = 8 * NT;
    float4 v[R];
#pragma unroll
    for (int r = 0; r < R; r++)
        v[r] = r < p.rows ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)p.row_map[r] * ld + k0 + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; r++) *reinterpret_cast<uint2*>(xs + (size_t)r * LDX + k) = make_uint2(pack_bf16(v[r].x, v[r].y), pack_bf16(v[r].z, v[r].w));
    cons_sync();
}

// one ring slot = `kt_item` k-tiles of one 16-row strip; the 8 consumer warps split them; partial sums -> part[warp][sub]
template <int NT>
__device__ __forceinline__ void strip_mma(const uint8_t* slot, const bf16* xs, int kt_item, int xk0, float* part, int sub) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    const int per = kt_item >> 3;
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; j++)
#pragma unroll
        for (int r = 0; r < 4; r++) acc[j][r] = 0.f;
    for (int t = 0; t < per; t++) {
        const int kt = warp * per + t;
        const uint4 w = *reinterpret_cast<const uint4*>(slot + ((size_t)kt * 32 + lane) * 16);
        const uint32_t a[4] = {w.x, w.y, w.z, w.w};
        const int k0 = xk0 + (kt << 4);
#pragma unroll
        for (int j = 0; j < NT; j++) {
            const bf16* xr = xs + (size_t)(j * 8 + g) * LDX + k0 + tg * 2;
            mma_bf16(acc[j], a, *reinterpret_cast<const uint32_t*>(xr), *reinterpret_cast<const uint32_t*>(xr + 8));
        }
    }
    float* pw = part + ((size_t)(warp * 2 + sub) * 16) * (8 * NT);
#pragma unroll
    for (int j = 0; j < NT; j++) {
        pw[g * (8 * NT) + j * 8 + tg * 2] = acc[j][0];
        pw[g * (8 * NT) + j * 8 + tg * 2 + 1] = acc[j][1];
        pw[(g + 8) * (8 * NT) + j * 8 + tg * 2] = acc[j][2];
        pw[(g + 8) * (8 * NT) + j * 8 + tg * 2 + 1] = acc[j][3];
    }
}

template <int NT>
__device__ __forceinline__ float part_sum(const float* part, int sub, int f, int r) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) s += part[((size_t)(w * 2 + sub) * 16 + f) * (8 * NT) + r];
    return s;
}

// Flash-decoding style attention item: (compact row r, head h, KV split s of ns).  Each warp streams whole KV pages
// (K and V of a page are fetched together), keeps an online-softmax partial (m, l, o[64]); the 8 warps are merged in
// shared memory and the CTA writes one partial {m, l, o[64]} to apart; the O-proj staging merges the ns partials.
constexpr int AP = 66;   // floats per partial
__device__ void attention_item(const MegaParams& p, int l, int r, int h, int s, int ns, float* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* qs = scratch;             // [64]
    float* wm = scratch + 64;        // [8] per-warp max
    float* wl = scratch + 72;        // [8] per-warp sum
    float* wo = scratch + 80;        // [8][64] per-warp output
    const int row = p.row_map[r], slot = row >> 1;
    const int pos = p.slot_pos[slot];
    const int* pt = p.page_table + (size_t)row * p.max_pages;
    const float* qkv = p.qkv + (size_t)row * (3 * D);
    bf16* kpool = p.kv + (size_t)l * p.kv_layer_stride;
    bf16* vpool = kpool + p.kv_half;
    const int n = pos + 1, npages = (n + PAGE - 1) / PAGE;
    const int per = (npages + ns - 1) / ns, pg0 = s * per, pg1 = min(npages, pg0 + per);
    if (tid < 32) {
        float sn, cs;
        sincosf((float)pos * p.inv_freq[tid], &sn, &cs);
        const float q0 = __ldcg(qkv + h * HD + tid), q1 = __ldcg(qkv + h * HD + tid + 32);
        qs[tid] = (q0 * cs - q1 * sn) * 0.125f;
        qs[tid + 32] = (q1 * cs + q0 * sn) * 0.125f;
        if (pos / PAGE >= pg0 && pos / PAGE < pg1) {   // the split that owns the newest position appends k/v to the cache
            const float k0 = __ldcg(qkv + D + h * HD + tid), k1 = __ldcg(qkv + D + h * HD + tid + 32);
            const size_t base = (((size_t)pt[pos / PAGE] * H + h) * PAGE + (pos % PAGE)) * HD;
            kpool[base + tid] = __float2bfloat16(k0 * cs - k1 * sn);
            kpool[base + tid + 32] = __float2bfloat16(k1 * cs + k0 * sn);
            vpool[base + tid] = __float2bfloat16(__ldcg(qkv + 2 * D + h * HD + tid));
            vpool[base + tid + 32] = __float2bfloat16(__ldcg(qkv + 2 * D + h * HD + tid + 32));
        }
    }
    cons_sync();
    const int pp = lane >> 1, half = lane & 1;
    float m = -INFINITY, lsum = 0.f;
    float acc[32];
#pragma unroll
    for (int d = 0; d < 32; d++) acc[d] = 0.f;
    for (int pg = pg0 + warp; pg < pg1; pg += 8) {
        const size_t off = (((size_t)pt[pg] * H + h) * PAGE + pp) * HD + half * 32;
        const uint4* kp = reinterpret_cast<const uint4*>(kpool + off);
        const uint4* vp = reinterpret_cast<const uint4*>(vpool + off);
        uint4 ku[4], vu[4];
#pragma unroll
        for (int c = 0; c < 4; c++) { ku[c] = kp[c]; vu[c] = vp[c]; }
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
        if (pg * PAGE + pp >= n) sc = -INFINITY;
        const float mnew = fmaxf(m, warp_max(sc));
        const float corr = expf(m - mnew), pj = expf(sc - mnew);   // mnew is finite: every page holds >= 1 valid position
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
    }
    // lanes of one `half` hold different positions: sum them (lsum is duplicated over the two halves)
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
        for (int d = 0; d < 32; d++) wo[warp * 64 + lane * 32 + d] = acc[d];
        if (lane == 0) { wm[warp] = m; wl[warp] = lsum; }
    }
    cons_sync();
    if (tid < HD) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; w++) M = fmaxf(M, wm[w]);
        float Lt = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const float e = wm[w] == -INFINITY ? 0.f : expf(wm[w] - M);
            Lt += wl[w] * e; O += wo[w * 64 + tid] * e;
        }
        float* out = p.apart + ((size_t)(row * H + h) * 8 + s) * AP;
        out[2 + tid] = O;
        if (tid == 0) { out[0] = M; out[1] = Lt; }
    }
    cons_sync();
}

// xs[r][:] <- merged attention output of row r (all heads), from the ns split partials
template <int NT>
__device__ void stage_attn_rows(const MegaParams& p, bf16* xs, int ns) {
    const int k = threadIdx.x * 4, h = k >> 6, d = k & 63;
    constexpr int R = 8 * NT;
#pragma unroll 2
    for (int r = 0; r < R; r++) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < p.rows) {
            const float* base = p.apart + (size_t)(p.row_map[r] * H + h) * 8 * AP;
            float M = -INFINITY;
            for (int s = 0; s < ns; s++) M = fmaxf(M, __ldcg(base + s * AP));
            float Lt = 0.f;
            for (int s = 0; s < ns; s++) {
                const float ms = __ldcg(base + s * AP);
                const float e = ms == -INFINITY ? 0.f : expf(ms - M);
                Lt += __ldcg(base + s * AP + 1) * e;
                const float* ov = base + s * AP + 2 + d;
                o.x += __ldcg(ov) * e; o.y += __ldcg(ov + 1) * e; o.z += __ldcg(ov + 2) * e; o.w += __ldcg(ov + 3) * e;
            }
            const float inv = 1.f / Lt;
            o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
        }
        *reinterpret_cast<uint2*>(xs + (size_t)r * LDX + k) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
    cons_sync();
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int NT>
__global__ void __launch_bounds__(THREADS, 1) t3_mega_kernel(const MegaParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring_mem = smem;
    bf16* xs = reinterpret_cast<bf16*>(smem + NSLOTS * SLOT);
    float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xs) + (size_t)8 * NT * LDX * 2);
    float* sc = part + 8 * 2 * 16 * 8 * NT;
    float* scratch = sc + p.max_seq + PAGE;
    float* red16 = scratch + 80 + 8 * 2 * 32;   // [16][8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red16 + 128);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    Ring ring;
    ring.base = smem_u32(ring_mem); ring.full0 = smem_u32(bars); ring.empty0 = ring.full0 + 8 * NSLOTS; ring.slot = 0; ring.phase = 0;
    if (tid == 0) {
        for (int s = 0; s < NSLOTS; s++) { mbar_init(ring.full0 + 8 * s, 1); mbar_init(ring.empty0 + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Deal deal{0, G, cta};
    if (warp == 8) {
        if ((tid & 31) == 0) producer(p, ring, deal);
        return;
    }
    // ---------------------------------------------------------------- consumers (256 threads)
    unsigned int bar_target = 0;
    auto gbar = [&]() { bar_target += G; grid_barrier(p.bar, bar_target); };
    auto wait_slot = [&]() -> const uint8_t* {
        mbar_wait(ring.full0 + 8 * ring.slot, ring.phase);
        return ring_mem + (size_t)ring.slot * SLOT;
    };
    auto release_slot = [&]() {   // call after every consumer warp is done reading the slot (i.e. after a cons_sync)
        if (tid == 0) mbar_arrive(ring.empty0 + 8 * ring.slot);
        ring.advance();
    };
    const int per_strip = 16 * 8 * NT;
    int ns = G / (p.rows * H);   // KV splits per (row, head): fill the grid when few rows are active
    ns = ns < 1 ? 1 : (ns > 8 ? 8 : ns);
    const float* xin = p.x;   // residual stream entering the layer
    for (int l = 0; l < p.n_layers; l++) {
        const MegaLayer& L = p.layers[l];
        // ---- P1: x = xin (+ down partials of the previous layer); RMSNorm; QKV strips
        MTRACE(0);
        if (l == 0) stage_norm_rows<NT, false>(p, xs, xin, nullptr, nullptr, L.ln1, cta, G, red16);
        else stage_norm_rows<NT, true>(p, xs, xin, p.dpart, p.xa, L.ln1, cta, G, red16);
        if (l > 0) xin = p.xa;
        MTRACE(1);
        for (int i = deal.first(); i < QKV_ITEMS; i += G) {
            const uint8_t* s = wait_slot();
            strip_mma<NT>(s, xs, 64, 0, part, 0);
            cons_sync();
            release_slot();
            for (int e = tid; e < per_strip; e += CONS) {
                const int f = e / (8 * NT), r = e % (8 * NT);
                if (r < p.rows) p.qkv[(size_t)p.row_map[r] * (3 * D) + i * 16 + f] = part_sum<NT>(part, 0, f, r);
            }
            cons_sync();
        }
        deal.next_phase(QKV_ITEMS);
        MTRACE(2);
        gbar();
        MTRACE(3);
        // ---- P2: attention, one (row, head) per CTA turn
        for (int it = cta; it < p.rows * H * ns; it += G) attention_item(p, l, (it / ns) / H, (it / ns) % H, it % ns, ns, scratch);
        MTRACE(4);
        gbar();
        MTRACE(5);
        // ---- P3: O-proj K-quarters -> opart[q]
        stage_attn_rows<NT>(p, xs, ns);
        MTRACE(6);
        for (int i = deal.first(); i < OP_ITEMS; i += G) {
            const int strip = i >> 2, q = i & 3;
            const uint8_t* s = wait_slot();
            strip_mma<NT>(s, xs, 16, q * 256, part, 0);
            cons_sync();
            release_slot();
            for (int e = tid; e < per_strip; e += CONS) {
                const int f = e / (8 * NT), r = e % (8 * NT);
                if (r < p.rows) p.opart[((size_t)q * p.rows_total + p.row_map[r]) * D + strip * 16 + f] = part_sum<NT>(part, 0, f, r);
            }
            cons_sync();
        }
        deal.next_phase(OP_ITEMS);
        MTRACE(7);
        gbar();
        MTRACE(8);
        // ---- P4: x2 = xin + sum opart -> xb; RMSNorm; gate/up strip pairs + SwiGLU -> act
        stage_norm_rows<NT, true>(p, xs, xin, p.opart, p.xb, L.ln2, cta, G, red16);
        MTRACE(9);
        for (int i = deal.first(); i < GU_ITEMS; i += G) {
            const uint8_t* s0 = wait_slot();
            strip_mma<NT>(s0, xs, 64, 0, part, 0);
            cons_sync();
            release_slot();
            const uint8_t* s1 = wait_slot();
            strip_mma<NT>(s1, xs, 64, 0, part, 1);
            cons_sync();
            release_slot();
            for (int e = tid; e < per_strip; e += CONS) {
                const int f = e / (8 * NT), r = e % (8 * NT);
                if (r < p.rows) {
                    const float gt = part_sum<NT>(part, 0, f, r), up = part_sum<NT>(part, 1, f, r);
                    p.act[(size_t)p.row_map[r] * FFN + i * 16 + f] = gt / (1.f + expf(-gt)) * up;
                }
            }
            cons_sync();
        }
        deal.next_phase(GU_ITEMS);
        MTRACE(10);
        gbar();
        MTRACE(11);
        // ---- P5: down-proj K-quarters -> dpart[q]
        for (int i = deal.first(); i < DN_ITEMS; i += G) {
            const int strip = i >> 2, q = i & 3;
            stage_plain_rows<NT>(p, xs, p.act, FFN, q * 1024);
            const uint8_t* s = wait_slot();
            strip_mma<NT>(s, xs, 64, 0, part, 0);
            cons_sync();
            release_slot();
            for (int e = tid; e < per_strip; e += CONS) {
                const int f = e / (8 * NT), r = e % (8 * NT);
                if (r < p.rows) p.dpart[((size_t)q * p.rows_total + p.row_map[r]) * D + strip * 16 + f] = part_sum<NT>(part, 0, f, r);
            }
            cons_sync();
        }
        deal.next_phase(DN_ITEMS);
        MTRACE(12);
        gbar();
        MTRACE(13);
        xin = p.xb;
    }
    // ---- head: x = xb + sum dpart; final RMSNorm; logits
    stage_norm_rows<NT, true>(p, xs, xin, p.dpart, nullptr, p.final_norm, cta, G, red16);
    for (int i = deal.first(); i < p.head_items; i += G) {
        const uint8_t* s = wait_slot();
        strip_mma<NT>(s, xs, 64, 0, part, 0);
        cons_sync();
        release_slot();
        for (int e = tid; e < per_strip; e += CONS) {
            const int f = e / (8 * NT), r = e % (8 * NT);
            const int col = i * 16 + f;
            if (r < p.rows && col < p.vocab) p.logits[(size_t)p.row_map[r] * p.ld_logits + col] = part_sum<NT>(part, 0, f, r);
        }
        cons_sync();
    }
}

size_t mega_smem(int NT, int max_seq) {
    return (size_t)NSLOTS * SLOT + (size_t)8 * NT * LDX * 2 + (size_t)8 * 2 * 16 * 8 * NT * 4 + (size_t)(max_seq + PAGE) * 4 + (80 + 8 * 2 * 32 + 128) * 4 + 2 * NSLOTS * 8 + 64;
}

}  // namespace

int t3_mega_grid = 0;

bool t3_mega_init(int max_seq) {
    int dev = 0, sms = 0, coop = 0;
    CBX_CHECK(cudaGetDevice(&dev));
    CBX_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CBX_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return false;
    CBX_CHECK(cudaFuncSetAttribute(t3_mega_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega_smem(1, max_seq)));
    CBX_CHECK(cudaFuncSetAttribute(t3_mega_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega_smem(2, max_seq)));
    int occ = 0;
    CBX_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, t3_mega_kernel<2>, THREADS, mega_smem(2, max_seq)));
    if (occ < 1) return false;
    t3_mega_grid = sms;
    return true;
}

void launch_t3_mega(const MegaParams& p, int max_seq, cudaStream_t st) {
    CBX_REQUIRE(t3_mega_grid > 0, "t3 megakernel not initialised");
    CBX_REQUIRE(p.rows >= 1 && p.rows <= 16, "t3 megakernel: rows must be in [1,16]");
    ProfScope ps(PC_GEMV, 2.0 * (p.n_layers * 16777216.0 + (double)p.head_items * 16 * D), st);
    CBX_CHECK(cudaMemsetAsync(p.bar, 0, sizeof(unsigned int), st));
    void* args[] = {(void*)&p};
    if (p.rows <= 8) CBX_CHECK(cudaLaunchCooperativeKernel((void*)t3_mega_kernel<1>, dim3(t3_mega_grid), dim3(THREADS), args, mega_smem(1, max_seq), st));
    else CBX_CHECK(cudaLaunchCooperativeKernel((void*)t3_mega_kernel<2>, dim3(t3_mega_grid), dim3(THREADS), args, mega_smem(2, max_seq), st));
}

extern "C" int cbx_t3_mega_trace(unsigned long long* out_h) { return cudaMemcpyFromSymbol(out_h, g_mega_trace, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : 1; }
