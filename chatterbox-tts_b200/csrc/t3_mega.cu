// T3 decode step as ONE persistent kernel (sm_100a): one CTA per SM; thread 0 of each CTA streams the CTA's
// share of the step's 1.02 GB of fragment-ordered bf16 weights through a ring of 32 KB shared-memory slots with bulk
// TMA copies (cp.async.bulk + mbarrier complete_tx), and the 8 warps run the layer phases
//   P1 RMSNorm + QKV | P2 RoPE + paged-KV attention | P3 O-proj | P4 RMSNorm + gate/up + SwiGLU | P5 down
// The weight stream never waits for the phases (weights do not depend on activations), so HBM stays busy while the
// dependent part of a phase is in flight.
//
// Phases are NOT separated by grid barriers.  Activations travel between CTAs through "flagged words": every value is
// published as one 8-byte store {payload, tag} and consumers spin on the words they need until the tag of the current
// (step, layer, phase) shows up -- value and flag arrive in the same L2 transaction, so a phase boundary costs one
// store + one load round trip instead of store / fence / atomic / poll / reload.  The residual stream lives in the
// consumers' registers (every CTA keeps its own fp32 copy of all rows; thread t owns columns 4t..4t+3), so the only
// data exchanged per layer are qkv (per head), the attention partials, the O-proj output, the SwiGLU activations
// (bf16 pairs) and two K-half partials of the down projection.  All reductions have a fixed order (no atomics):
// results do not depend on timing or on which other rows are in the batch.
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "t3_kernels.cuh"

namespace {

constexpr int D = 1024, FFN = 4096, H = 16, HD = 64, PAGE = 16;
constexpr int SLOT = 32768, THREADS = 256;
constexpr int LDX = D + 8, LDX2 = 2 * D + 8;
constexpr int QKV_ITEMS = 3 * D / 16, OP_ITEMS = D / 16, GU_ITEMS = FFN / 16, DN_ITEMS = (D / 16) * 2;
constexpr int NS_MAX = 4, AP = 66;   // attention KV splits per (row, head); words per partial {max, sum, o[64]}
constexpr int PHASES = 5, MAX_LAYERS = 32;

// RA = row capacity of the instance (2: one stream, 8, 16); XR = rows of the bf16 staging tile that really exist in
// shared memory (the MMA's other row slots read as zero), which buys the single-stream instance a sixth ring slot
template <int NT, int RA> struct Cfg { static constexpr int XR = RA == 2 ? 2 : 8 * NT, NSLOTS = RA == 2 ? 6 : (NT == 1 ? 5 : 4), R = 8 * NT, NPART = 2; };

__device__ unsigned long long g_mega_trace[64];
__device__ long long g_mega_prof[256 * 4];   // per CTA (thread 0): cycles waiting for weights, for arrival counters, total

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define MTRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && l == 1) g_mega_trace[i] = gtime(); } while (0)

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (long spin = 0; spin < (1L << 26); spin++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cons_sync() { __syncthreads(); }

__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------------------ flagged words
typedef unsigned long long u64;
__device__ __forceinline__ void ll_store(u64* p, uint32_t payload, uint32_t tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_store2(u64* p, uint32_t a, uint32_t b, uint32_t tag) {   // two adjacent words (16 B aligned)
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(p), "r"(a), "r"(tag), "r"(b) : "memory");
}
__device__ __forceinline__ void ll_load2(const u64* p, uint32_t (&v)[4]) {   // {payload0, tag0, payload1, tag1}
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p) : "memory");
}
// Arrival counters (one set per layer, zeroed before every step) are only a HINT that a phase's words are probably there:
// one thread per CTA spins on a counter instead of 256 threads sweeping the data lines (which floods the L2 slices the
// producers are writing to).  Validity still comes from the tags, so no fences or release/acquire are needed.
constexpr int CNT_STRIDE = 32, CNT_QKV = 0, CNT_ATT = 16, CNT_Y = 17, CNT_ACT = 18, CNT_Z = 20;
__device__ __forceinline__ void warp_signal(bool active, unsigned int* c) {   // every warp with a publishing lane adds 1
    const unsigned m = __ballot_sync(0xffffffffu, active);
    if (m != 0u && (threadIdx.x & 31) == __ffs(m) - 1) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
}
struct Spin {   // bounded spinning: a word that never arrives is a bug, trap instead of hanging the GPU
    unsigned int n = 0;
    __device__ __forceinline__ void tick() { if (++n > (1u << 22)) __trap(); }
};
__device__ __forceinline__ void wait_cnt(const unsigned int* c, unsigned int target, long long& t_acc) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        Spin sp;
        while (true) {
            unsigned int v;
            asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            if (v >= target) break;
            sp.tick();
        }
        t_acc += clock64() - t0;
    }
    __syncthreads();
}
// N groups of 4 consecutive words (32 B each, 16 B aligned) at p + g * stride
template <int N>
__device__ __forceinline__ void ll_poll4(const u64* p, long stride, int n_valid, uint32_t tag, uint32_t (&out)[N][4]) {
    Spin sp;
    while (true) {
        uint32_t a[N][4], b[N][4];
#pragma unroll
        for (int g = 0; g < N; g++)
            if (g < n_valid) { ll_load2(p + g * stride, a[g]); ll_load2(p + g * stride + 2, b[g]); }
        bool ok = true;
#pragma unroll
        for (int g = 0; g < N; g++)
            if (g < n_valid) ok = ok && a[g][1] == tag && a[g][3] == tag && b[g][1] == tag && b[g][3] == tag;
        if (ok) {
#pragma unroll
            for (int g = 0; g < N; g++) { out[g][0] = a[g][0]; out[g][1] = a[g][2]; out[g][2] = b[g][0]; out[g][3] = b[g][2]; }
            return;
        }
        sp.tick();
    }
}

// items of a phase dealt round-robin: item i belongs to CTA (i + off) % G
struct Deal {
    int off, G, cta;
    __device__ int first() const { int f = cta - off; return f < 0 ? f + G : f; }
    __device__ void next_phase(int n_items) { off = (off + n_items) % G; }
};

// ------------------------------------------------------------------------------------------------ weight feed
// A CTA's weight slots in consumption order: per layer its QKV strips, O strip, (gate, up) strip pairs and down K-halves
// (2 slots each), then its head strips.  The order depends only on the grid size and the model, so the host builds it
// once (t3_mega_init) as a table of source addresses [grid][SCHED_MAX]; the kernel copies its row to shared memory.
// The ring is primed at kernel start and every slot is re-armed with the next chunk the moment its round is over, so
// the stream only stalls when the ring is full of data the phases have not reached yet; a second cursor issues L2
// bulk prefetches `l2_ahead` slots further down the list so that HBM -> L2 traffic is decoupled from the phases.
constexpr int SCHED_MAX = 256;

// ------------------------------------------------------------------------------------------------ per-step constants
// Everything the phases would otherwise fetch through chains of dependent global loads is copied to shared memory once
// per step: row map, KV lengths, split counts, page tables and the RoPE rotation of each row's new position.
struct StepConst {
    int* row; int* pos; int* ns; int* pt; float* cs; float* sn; int max_pages;
};

// ------------------------------------------------------------------------------------------------ consumer pieces
// RMSNorm of the register-resident rows -> xs (bf16, [R][LDX])
template <int NT, int RA>
__device__ __forceinline__ void norm_to_xs(const float4 (&xr)[RA], int rows, bf16* xs, const float4 gg, float eps, float* red) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, k = tid * 4;
#pragma unroll
    for (int r = 0; r < RA; r++) {
        if (r < rows) {   // uniform
            float ss = xr[r].x * xr[r].x + xr[r].y * xr[r].y + xr[r].z * xr[r].z + xr[r].w * xr[r].w;
            ss = warp_sum(ss);
            if (lane == 0) red[r * 8 + warp] = ss;
        }
    }
    cons_sync();
#pragma unroll
    for (int r = RA; r < Cfg<NT, RA>::XR; r++) *reinterpret_cast<uint2*>(xs + (size_t)r * LDX + k) = make_uint2(0u, 0u);
#pragma unroll
    for (int r = 0; r < RA; r++) {
        uint2 o = make_uint2(0u, 0u);
        if (r < rows) {
            const float4 a = *reinterpret_cast<const float4*>(red + r * 8), b = *reinterpret_cast<const float4*>(red + r * 8 + 4);
            const float ss = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
            const float scale = rsqrtf(ss / D + eps);
            o = make_uint2(pack_bf16(xr[r].x * scale * gg.x, xr[r].y * scale * gg.y), pack_bf16(xr[r].z * scale * gg.z, xr[r].w * scale * gg.w));
        }
        *reinterpret_cast<uint2*>(xs + (size_t)r * LDX + k) = o;
    }
    cons_sync();
}

// xr[r] += the flagged rows src[r][4t .. 4t+3] (fp32 payload) of NSRC sources `src_stride` words apart; rows handled 4
// at a time to bound registers.  Columns 4t..4t+3 belong to strip t/4.
template <int NT, int RA, int NSRC>
__device__ __forceinline__ void add_rows(float4 (&xr)[RA], int rows, const u64* src, long src_stride, uint32_t tag) {
    const int k = threadIdx.x * 4;
    if (RA == 2 && NSRC == 2) {   // the single-stream case: both partials of both rows in one round trip
        uint32_t a[4][4], b[4][4];
        Spin sp;
        while (true) {
            bool ok = true;
#pragma unroll
            for (int g = 0; g < 4; g++)
                if ((g & 1) < rows) { const u64* w = src + (g >> 1) * src_stride + (size_t)(g & 1) * D + k; ll_load2(w, a[g]); ll_load2(w + 2, b[g]); }
#pragma unroll
            for (int g = 0; g < 4; g++)
                if ((g & 1) < rows) ok = ok && a[g][1] == tag && a[g][3] == tag && b[g][1] == tag && b[g][3] == tag;
            if (ok) break;
            sp.tick();
        }
#pragma unroll
        for (int g = 0; g < 4; g++)
            if ((g & 1) < rows) {
                float4& x = xr[g & 1];
                x.x += __uint_as_float(a[g][0]); x.y += __uint_as_float(a[g][2]); x.z += __uint_as_float(b[g][0]); x.w += __uint_as_float(b[g][2]);
            }
        return;
    }
#pragma unroll
    for (int q = 0; q < NSRC; q++) {
#pragma unroll
        for (int r0 = 0; r0 < RA; r0 += 4) {
            if (r0 >= rows) break;
            constexpr int GN = RA < 4 ? RA : 4;
            uint32_t v[GN][4];
            ll_poll4<GN>(src + q * src_stride + (size_t)r0 * D + k, D, rows - r0, tag, v);
#pragma unroll
            for (int g = 0; g < GN; g++)
                if (r0 + g < rows) {
                    xr[r0 + g].x += __uint_as_float(v[g][0]); xr[r0 + g].y += __uint_as_float(v[g][1]);
                    xr[r0 + g].z += __uint_as_float(v[g][2]); xr[r0 + g].w += __uint_as_float(v[g][3]);
                }
        }
    }
}

// One round = up to two ring slots (64 k-tiles of one 16-row strip each).  The 8 warps split the k-tiles of both slots
// (two independent accumulator chains per warp); partial sums -> part[warp][sub].  s1 == nullptr: one slot only.
// One round = NS ring slots (1, 2 or 4; 64 k-tiles of one 16-row strip each) multiplied against the staged rows.  The 8
// warps are split BY SLOT (8/NS warps per slot, each taking 8*NS consecutive k-tiles) and a warp waits only for its own
// slot; partial sums -> part[warp][16][R].  After the block barrier the value (slot n, strip row f, row r) is the sum of
// the 8/NS warp partials of that slot.
struct RoundCtx {
    uint8_t* ring_mem; uint32_t full0; int slot; uint32_t phase; float* part; int pbuf; long long t_mbar, t_cnt;
};
template <int NT, int NS, int NSLOTS, int XR>
__device__ __forceinline__ const float* round_mma(RoundCtx& c, const bf16* xs, int ldx, int xk_step) {
    constexpr int W = 8 / NS, KT = 64 / W, R = 8 * NT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    const int n = warp / W, ws = warp % W;
    int sl = c.slot + n; uint32_t ph = c.phase;
    if (sl >= NSLOTS) { sl -= NSLOTS; ph ^= 1; }
    const long long t0 = clock64();
    mbar_wait(c.full0 + 8 * sl, ph);
    c.t_mbar += clock64() - t0;
    const uint8_t* slot = c.ring_mem + (size_t)sl * SLOT;
    float acc[2][NT][4];   // two independent accumulator chains (even / odd k-tiles)
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[q][j][r] = 0.f;
#pragma unroll
    for (int t0 = 0; t0 < KT; t0 += 8) {
        uint4 w[8];
#pragma unroll
        for (int t = 0; t < 8; t++) w[t] = *reinterpret_cast<const uint4*>(slot + ((size_t)(ws * KT + t0 + t) * 32 + lane) * 16);
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int kk = n * xk_step + ((ws * KT + t0 + t) << 4) + tg * 2;
            const uint32_t a[4] = {w[t].x, w[t].y, w[t].z, w[t].w};
#pragma unroll
            for (int j = 0; j < NT; j++) {
                uint32_t b0 = 0u, b1 = 0u;
                if (j * 8 + g < XR) {
                    const bf16* xr = xs + (size_t)(j * 8 + g) * ldx + kk;
                    b0 = *reinterpret_cast<const uint32_t*>(xr); b1 = *reinterpret_cast<const uint32_t*>(xr + 8);
                }
                mma_bf16(acc[t & 1][j], a, b0, b1);
            }
        }
    }
    float* pb = c.part + c.pbuf * (8 * 16 * R);
    float* pw = pb + (size_t)warp * 16 * R;
#pragma unroll
    for (int j = 0; j < NT; j++) {
        *reinterpret_cast<float2*>(pw + g * R + j * 8 + tg * 2) = make_float2(acc[0][j][0] + acc[1][j][0], acc[0][j][1] + acc[1][j][1]);
        *reinterpret_cast<float2*>(pw + (g + 8) * R + j * 8 + tg * 2) = make_float2(acc[0][j][2] + acc[1][j][2], acc[0][j][3] + acc[1][j][3]);
    }
    cons_sync();
    c.pbuf ^= 1;   // the next round writes the other buffer; the one after is separated from our readers by its barrier
    return pb;
}
template <int NT, int NS>
__device__ __forceinline__ float part_sum(const float* pb, int n, int f, int r) {
    constexpr int W = 8 / NS, R = 8 * NT;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < W; w++) s += pb[((size_t)(n * W + w) * 16 + f) * R + r];
    return s;
}

__device__ __forceinline__ int kv_splits(int pos) {   // depends on the row's own length only (batch invariant)
    const int npages = (pos + PAGE - 1) / PAGE;
    const int ns = (npages + 7) / 8;
    return ns < 1 ? 1 : (ns > NS_MAX ? NS_MAX : ns);
}

// Attention item (compact row r, head h, KV split s): positions [0, pos) come from the paged cache; the new position's
// q/k/v come from the flagged qkv words of this layer.  The last split also appends k/v to the cache and folds the new
// position into its partial.  Output: one flagged partial {m, l, o[64]} (fp32).  The same pages of the next layer are
// prefetched into L2 so that only layer 0 pays HBM latency on the dependent path.
__device__ __forceinline__ void attention_item(const MegaParams& p, const StepConst& sc_, int l, int r, int h, int s, const u64* qkv_ll, u64* ap_ll, uint32_t tag_in, uint32_t tag_out, float* scratch, unsigned int* cnt, unsigned int qkv_target, long long& t_cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* qs = scratch;              // [64] rotated, scaled query
    float* raw = scratch + 64;        // [192] raw q | k | v of this head
    float* kn = scratch + 256;        // [64] new key (bf16-rounded), [64] new value
    float* wm = scratch + 384;        // [9] per-warp max (+ new position)
    float* wl = scratch + 400;        // [9]
    float* wo = scratch + 416;        // [9][64]
    const int ns = sc_.ns[r];
    if (s >= ns) return;
    const int pos = sc_.pos[r];
    const int* pt = sc_.pt + r * sc_.max_pages;
    bf16* kpool = p.kv + (size_t)l * p.kv_layer_stride;
    bf16* vpool = kpool + p.kv_half;
    const int npages = (pos + PAGE - 1) / PAGE;
    const int per = (npages + ns - 1) / ns, pg0 = s * per, pg1 = min(npages, pg0 + per);
    // four lanes per cached position (16 of the 64 dims each), eight positions per warp pass: a page is two "units"
    const int pp = lane >> 2, quarter = lane & 3;
    const int u1 = 2 * pg1;
    int u = 2 * pg0 + warp;
    auto unit_off = [&](int uu) { return (((size_t)pt[uu >> 1] * H + h) * PAGE + (uu & 1) * 8 + pp) * HD + quarter * 16; };
    // the cached pages do not depend on this step: get this warp's first two units moving before waiting for q
    uint4 ku[2][2], vu[2][2];
    auto load_unit = [&](int uu, uint4 (&kk)[2], uint4 (&vv)[2]) {
        if (uu < u1) {
            const size_t off = unit_off(uu);
            const uint4* kp = reinterpret_cast<const uint4*>(kpool + off);
            const uint4* vp = reinterpret_cast<const uint4*>(vpool + off);
            kk[0] = __ldcg(kp); kk[1] = __ldcg(kp + 1); vv[0] = __ldcg(vp); vv[1] = __ldcg(vp + 1);
        }
    };
    load_unit(u, ku[0], vu[0]);
    load_unit(u + 8, ku[1], vu[1]);
    if (l + 1 < p.n_layers && quarter == 0) {   // one 128-byte line per position: next layer's K and V -> L2
        for (int q = u; q < u1; q += 8) {
            const size_t off = unit_off(q) + p.kv_layer_stride;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(kpool + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vpool + off));
        }
    }
    wait_cnt(cnt + CNT_QKV + h, qkv_target, t_cnt);
    if (tid < 96) {   // q | k | v of head h: 3 x 64 words, two per thread
        const int part = tid >> 5, d = (tid & 31) * 2;
        const u64* src = qkv_ll + (size_t)r * (3 * D) + part * D + h * HD + d;
        uint32_t v[4];
        Spin sp;
        while (true) { ll_load2(src, v); if (v[1] == tag_in && v[3] == tag_in) break; sp.tick(); }
        *reinterpret_cast<float2*>(raw + part * 64 + d) = make_float2(__uint_as_float(v[0]), __uint_as_float(v[2]));
    }
    MTRACE(8);
    cons_sync();
    MTRACE(9);
    if (tid < 32) {   // rotate_half RoPE: pairs (d, d+32)
        const float cs = sc_.cs[r * 32 + tid], sn = sc_.sn[r * 32 + tid];
        const float q0 = raw[tid], q1 = raw[tid + 32], k0 = raw[64 + tid], k1 = raw[64 + tid + 32];
        qs[tid] = (q0 * cs - q1 * sn) * 0.125f;
        qs[tid + 32] = (q1 * cs + q0 * sn) * 0.125f;
        const bf16 kr0 = __float2bfloat16(k0 * cs - k1 * sn), kr1 = __float2bfloat16(k1 * cs + k0 * sn);
        const bf16 v0 = __float2bfloat16(raw[128 + tid]), v1 = __float2bfloat16(raw[128 + tid + 32]);
        kn[tid] = __bfloat162float(kr0); kn[tid + 32] = __bfloat162float(kr1);
        kn[64 + tid] = __bfloat162float(v0); kn[64 + tid + 32] = __bfloat162float(v1);
        if (s == ns - 1) {
            const size_t base = (((size_t)pt[pos / PAGE] * H + h) * PAGE + (pos % PAGE)) * HD;
            kpool[base + tid] = kr0; kpool[base + tid + 32] = kr1;
            vpool[base + tid] = v0; vpool[base + tid + 32] = v1;
        }
    }
    cons_sync();
    MTRACE(10);
    float m = -INFINITY, lsum = 0.f;
    float acc[16];
#pragma unroll
    for (int d = 0; d < 16; d++) acc[d] = 0.f;
    auto consume = [&](int uu, const uint4 (&kk)[2], const uint4 (&vv)[2]) {
        float sc = 0.f;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&kk[c]);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __bfloat1622float2(b2[e]);
                sc += f.x * qs[quarter * 16 + c * 8 + e * 2] + f.y * qs[quarter * 16 + c * 8 + e * 2 + 1];
            }
        }
        sc += __shfl_xor_sync(0xffffffffu, sc, 1);
        sc += __shfl_xor_sync(0xffffffffu, sc, 2);
        if ((uu >> 1) * PAGE + (uu & 1) * 8 + pp >= pos) sc = -INFINITY;   // the rest of the newest page is not cached yet
        const float mnew = fmaxf(m, warp_max(sc));
        // a unit can be entirely beyond pos (second half of the newest page): then mnew may still be -inf
        const float corr = mnew == -INFINITY ? 1.f : __expf(m - mnew), pj = mnew == -INFINITY ? 0.f : __expf(sc - mnew);
        m = mnew;
        lsum = lsum * corr + pj;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&vv[c]);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __bfloat1622float2(b2[e]);
                acc[c * 8 + e * 2] = acc[c * 8 + e * 2] * corr + (pj > 0.f ? pj * f.x : 0.f);
                acc[c * 8 + e * 2 + 1] = acc[c * 8 + e * 2 + 1] * corr + (pj > 0.f ? pj * f.y : 0.f);
            }
        }
    };
    while (u < u1) {
        consume(u, ku[0], vu[0]);
        if (u + 8 < u1) consume(u + 8, ku[1], vu[1]);
        u += 16;
        load_unit(u, ku[0], vu[0]);
        load_unit(u + 8, ku[1], vu[1]);
    }
    MTRACE(11);
    // lanes of one `quarter` hold different positions: sum them (lsum is duplicated over the four quarters)
#pragma unroll
    for (int d = 0; d < 16; d++) {
        float v = acc[d];
        v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[d] = v;
    }
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 4); lsum += __shfl_xor_sync(0xffffffffu, lsum, 8); lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
    if (lane < 4) {
#pragma unroll
        for (int d = 0; d < 16; d++) wo[warp * 64 + lane * 16 + d] = acc[d];
        if (lane == 0) { wm[warp] = m; wl[warp] = lsum; }
    }
    if (warp == 0) {   // the new position as a ninth partial (owner split only)
        float sc = qs[lane] * kn[lane] + qs[lane + 32] * kn[lane + 32];
        sc = warp_sum(sc);
        const bool own = s == ns - 1;
        wo[8 * 64 + lane] = own ? kn[64 + lane] : 0.f;
        wo[8 * 64 + lane + 32] = own ? kn[64 + lane + 32] : 0.f;
        if (lane == 0) { wm[8] = own ? sc : -INFINITY; wl[8] = own ? 1.f : 0.f; }
    }
    MTRACE(12);
    cons_sync();
    MTRACE(13);
    if (tid < HD) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 9; w++) M = fmaxf(M, wm[w]);
        float Lt = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < 9; w++) {
            const float e = wm[w] == -INFINITY ? 0.f : __expf(wm[w] - M);
            Lt += wl[w] * e; O += wo[w * 64 + tid] * e;
        }
        u64* out = ap_ll + ((size_t)(r * H + h) * NS_MAX + s) * AP;
        ll_store(out + 2 + tid, __float_as_uint(O), tag_out);
        if (tid == 0) ll_store2(out, __float_as_uint(M), __float_as_uint(Lt), tag_out);
    }
    warp_signal(tid < HD, cnt + CNT_ATT);
    // no trailing barrier: the next user of `scratch` (another item, or nothing) starts with a barrier after its poll
}

// xs[r][:] <- merged attention output of compact row r (all heads) from the flagged split partials
template <int NT, int XR>
__device__ __forceinline__ void stage_attn_rows(const MegaParams& p, const StepConst& sc_, bf16* xs, const u64* ap_ll, uint32_t tag) {
    const int k = threadIdx.x * 4, h = k >> 6, d = k & 63;
    constexpr int R = 8 * NT;
    for (int r = 0; r < R; r++) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < p.rows) {
            const int ns = sc_.ns[r];
            const u64* base = ap_ll + (size_t)(r * H + h) * NS_MAX * AP;
            uint32_t ml[NS_MAX][4], a[NS_MAX][4], b[NS_MAX][4];
            Spin sp;
            while (true) {
                bool ok = true;
#pragma unroll
                for (int s = 0; s < NS_MAX; s++)
                    if (s < ns) { ll_load2(base + s * AP, ml[s]); ll_load2(base + s * AP + 2 + d, a[s]); ll_load2(base + s * AP + 4 + d, b[s]); }
#pragma unroll
                for (int s = 0; s < NS_MAX; s++)
                    if (s < ns) ok = ok && ml[s][1] == tag && ml[s][3] == tag && a[s][1] == tag && a[s][3] == tag && b[s][1] == tag && b[s][3] == tag;
                if (ok) break;
                sp.tick();
            }
            float M = -INFINITY;
#pragma unroll
            for (int s = 0; s < NS_MAX; s++) if (s < ns) M = fmaxf(M, __uint_as_float(ml[s][0]));
            float Lt = 0.f;
#pragma unroll
            for (int s = 0; s < NS_MAX; s++)
                if (s < ns) {
                    const float ms = __uint_as_float(ml[s][0]);
                    const float e = ms == -INFINITY ? 0.f : __expf(ms - M);
                    Lt += __uint_as_float(ml[s][2]) * e;
                    o.x += __uint_as_float(a[s][0]) * e; o.y += __uint_as_float(a[s][2]) * e;
                    o.z += __uint_as_float(b[s][0]) * e; o.w += __uint_as_float(b[s][2]) * e;
                }
            const float inv = 1.f / Lt;
            o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
        }
        *reinterpret_cast<uint2*>(xs + (size_t)r * LDX + k) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
    cons_sync();
}

// xs[r][0..2048) <- flagged bf16-pair activations act[r][kh*2048 ..]; thread t owns words 4t..4t+3 (8 values) of every row
template <int NT, int XR>
__device__ __forceinline__ void stage_act_rows(const MegaParams& p, bf16* xs, const u64* act_ll, int kh, uint32_t tag) {
    constexpr int R = 8 * NT;
    const int w0 = threadIdx.x * 4;
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += 4) {
        uint32_t v[4][4];
        if (r0 < p.rows) ll_poll4<4>(act_ll + (size_t)r0 * (FFN / 2) + kh * D + w0, FFN / 2, p.rows - r0, tag, v);
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (r0 + g < p.rows) o = make_uint4(v[g][0], v[g][1], v[g][2], v[g][3]);
            if (r0 + g < XR) *reinterpret_cast<uint4*>(xs + (size_t)(r0 + g) * LDX2 + w0 * 2) = o;
        }
    }
    cons_sync();
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int NT, int RA>
__global__ void __launch_bounds__(THREADS, 1) t3_mega_kernel(const MegaParams p) {
    constexpr int R = Cfg<NT, RA>::R, NSLOTS = Cfg<NT, RA>::NSLOTS, NPART = Cfg<NT, RA>::NPART, XR = Cfg<NT, RA>::XR;
    constexpr int PART = 8 * 16 * R;   // floats per partial-sum buffer
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring_mem = smem;
    bf16* xs = reinterpret_cast<bf16*>(smem + NSLOTS * SLOT);                                     // [XR][LDX2]
    float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xs) + (size_t)XR * LDX2 * 2);   // [NPART][8][16][R]
    float* scratch = part + NPART * PART;                                                         // attention scratch [1024]
    float* red = scratch + 1024;                                                                  // [R][8]
    StepConst sc_;
    sc_.cs = red + R * 8; sc_.sn = sc_.cs + R * 32;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sc_.sn + R * 32);
    MegaLayer* s_layers = reinterpret_cast<MegaLayer*>(bars + NSLOTS);                          // [MAX_LAYERS]
    unsigned long long* sched = reinterpret_cast<unsigned long long*>(s_layers + MAX_LAYERS);    // [SCHED_MAX] slot source addresses
    sc_.row = reinterpret_cast<int*>(sched + SCHED_MAX); sc_.pos = sc_.row + R; sc_.ns = sc_.pos + R; sc_.pt = sc_.ns + R; sc_.max_pages = p.max_pages;
    const int tid = threadIdx.x;
    const int G = gridDim.x, cta = blockIdx.x;
    RoundCtx rc;
    rc.ring_mem = ring_mem; rc.full0 = smem_u32(bars); rc.slot = 0; rc.phase = 0; rc.part = part; rc.pbuf = 0; rc.t_mbar = 0; rc.t_cnt = 0;
    const long long t_start = clock64();
    const uint32_t ring_base = smem_u32(ring_mem);
    constexpr int ARM_TID = 224;   // warp 7: publishes nothing for rows < 14, so re-arming is off the publishers' path
    const int n_sched = p.sched_count[cta];
    int cursor = 0;   // next entry of the schedule to fetch (uniform; only ARM_TID issues)
    auto arm = [&](int slot, int k) {   // one thread only
        if (k < n_sched) {
            const uint32_t full = rc.full0 + 8 * slot;
            mbar_expect_tx(full, SLOT);
            bulk_load(ring_base + slot * SLOT, reinterpret_cast<const void*>(sched[k]), SLOT, full);
        }
        if (p.l2_ahead > 0) {   // wraps into the next step, which starts with the same slots
            int kp = k + p.l2_ahead;
            if (kp >= n_sched) kp -= n_sched;
            bulk_prefetch_l2(reinterpret_cast<const void*>(sched[kp]), SLOT);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < NSLOTS; s++) mbar_init(rc.full0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < p.n_layers * (int)(sizeof(MegaLayer) / 8); i += THREADS)
        reinterpret_cast<unsigned long long*>(s_layers)[i] = reinterpret_cast<const unsigned long long*>(p.layers)[i];
    for (int i = tid; i < n_sched; i += THREADS) sched[i] = p.sched[(size_t)cta * SCHED_MAX + i];
    __syncthreads();
    if (tid == ARM_TID) {
        for (int k = NSLOTS; k < NSLOTS + p.l2_ahead - 1 && k < n_sched; k++) bulk_prefetch_l2(reinterpret_cast<const void*>(sched[k]), SLOT);
        for (int s = 0; s < NSLOTS; s++) arm(s, s);
    }
    cursor = NSLOTS;
    if (tid < p.rows) {
        const int row = p.row_map[tid], pos = p.slot_pos[row >> 1];
        sc_.row[tid] = row; sc_.pos[tid] = pos; sc_.ns[tid] = kv_splits(pos);
    }
    __syncthreads();
    for (int i = tid; i < p.rows * p.max_pages; i += THREADS) {
        const int r = i / p.max_pages, j = i % p.max_pages;
        sc_.pt[i] = j * PAGE <= sc_.pos[r] ? p.page_table[(size_t)sc_.row[r] * p.max_pages + j] : 0;
    }
    for (int i = tid; i < p.rows * 32; i += THREADS) {
        float sn, cs;
        sincosf((float)sc_.pos[i >> 5] * p.inv_freq[i & 31], &sn, &cs);
        sc_.cs[i] = cs; sc_.sn[i] = sn;
    }
    Deal deal{0, G, cta};
    const uint32_t tag0 = *p.epoch;   // advanced by CTA 0 at the end of the step; unique per (step, layer, phase)
    const int k4 = tid * 4;
    float4 xr[RA];   // residual stream, columns 4*tid .. 4*tid+3 of every active row (RA = row capacity of this instance)
#pragma unroll
    for (int r = 0; r < RA; r++)
        xr[r] = r < p.rows ? *reinterpret_cast<const float4*>(p.x + (size_t)sc_.row[r] * D + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // after a round's values are published: re-arm its slots with the next chunks of the feed and move on
    auto retire = [&](int ns) {
        if (tid == ARM_TID)
            for (int j = 0; j < ns; j++) { int sl = rc.slot + j; if (sl >= NSLOTS) sl -= NSLOTS; arm(sl, cursor + j); }
        cursor += ns;
        rc.slot += ns;
        if (rc.slot >= NSLOTS) { rc.slot -= NSLOTS; rc.phase ^= 1; }
    };
    // arrival-counter targets: every warp with a publishing lane signals once per item
    const unsigned int W2 = (p.rows + 1) / 2, W4 = (p.rows + 3) / 4;
    unsigned int att_target = 0;
    for (int r = 0; r < p.rows; r++) att_target += 2u * H * sc_.ns[r];
    const int pf = tid & 15, pr = tid >> 4;   // publisher mapping: strip row f, compact row r (coalesced 128-byte stores)
    const int n_virtual = p.rows * H * NS_MAX;
    for (int l = 0; l < p.n_layers; l++) {
        const MegaLayer& L = s_layers[l];
        const int par = l & 1;
        const uint32_t tg = tag0 + (uint32_t)l * PHASES;   // tags tg+0 .. tg+4: qkv, attention partials, y, act, z
        u64* qkv_ll = p.ll_qkv + (size_t)par * 16 * 3 * D;
        u64* ap_ll = p.ll_ap + (size_t)par * 16 * H * NS_MAX * AP;
        u64* y_ll = p.ll_y + (size_t)par * 16 * D;
        u64* act_ll = p.ll_act + (size_t)par * 16 * (FFN / 2);
        u64* z_ll = p.ll_z + (size_t)par * 2 * 16 * D;
        unsigned int* cnt = p.cnt + l * CNT_STRIDE;
        // ---- P1: x += the two K-half partials of the previous layer's down projection; RMSNorm; QKV strips
        MTRACE(0);
        float4 gg = *reinterpret_cast<const float4*>(L.ln1 + k4);   // in flight while waiting
        if (l > 0) {
            wait_cnt(cnt - CNT_STRIDE + CNT_Z, DN_ITEMS * W2, rc.t_cnt);
            add_rows<NT, RA, 2>(xr, p.rows, p.ll_z + (size_t)(par ^ 1) * 2 * 16 * D, 16 * D, tg - 1);
        }
        MTRACE(16);
        norm_to_xs<NT, RA>(xr, p.rows, xs, gg, p.eps, red);
        MTRACE(1);
        {
            const int i = deal.first();   // one or two strips per CTA (192 strips)
            if (i + G < QKV_ITEMS) {
                const float* pb = round_mma<NT, 2, NSLOTS, XR>(rc, xs, LDX, 0);
                if (pr < p.rows) {
                    ll_store(qkv_ll + (size_t)pr * (3 * D) + i * 16 + pf, __float_as_uint(part_sum<NT, 2>(pb, 0, pf, pr)), tg);
                    ll_store(qkv_ll + (size_t)pr * (3 * D) + (i + G) * 16 + pf, __float_as_uint(part_sum<NT, 2>(pb, 1, pf, pr)), tg);
                }
                warp_signal(pr < p.rows, cnt + CNT_QKV + (i & 63) / 4);
                warp_signal(pr < p.rows, cnt + CNT_QKV + ((i + G) & 63) / 4);
                retire(2);
            } else if (i < QKV_ITEMS) {
                const float* pb = round_mma<NT, 1, NSLOTS, XR>(rc, xs, LDX, 0);
                if (pr < p.rows) ll_store(qkv_ll + (size_t)pr * (3 * D) + i * 16 + pf, __float_as_uint(part_sum<NT, 1>(pb, 0, pf, pr)), tg);
                warp_signal(pr < p.rows, cnt + CNT_QKV + (i & 63) / 4);
                retire(1);
            }
        }
        deal.next_phase(QKV_ITEMS);
        MTRACE(2);
        // ---- P2: attention; virtual item it = (s * rows + r) * H + h
        for (int it = cta; it < n_virtual; it += G) attention_item(p, sc_, l, (it / H) % p.rows, it % H, it / (H * p.rows), qkv_ll, ap_ll, tg, tg + 1, scratch, cnt, 12 * W2, rc.t_cnt);
        MTRACE(3);
        // ---- P3: O-proj strips (full K) -> y
        {
            const int i = deal.first();
            if (i < OP_ITEMS) {   // at most one per CTA (64 strips)
                wait_cnt(cnt + CNT_ATT, att_target, rc.t_cnt);
                stage_attn_rows<NT, XR>(p, sc_, xs, ap_ll, tg + 1);
                MTRACE(14);
                const float* pb = round_mma<NT, 1, NSLOTS, XR>(rc, xs, LDX, 0);
                if (pr < p.rows) ll_store(y_ll + (size_t)pr * D + i * 16 + pf, __float_as_uint(part_sum<NT, 1>(pb, 0, pf, pr)), tg + 2);
                warp_signal(pr < p.rows, cnt + CNT_Y);
                retire(1);
            }
            deal.next_phase(OP_ITEMS);
        }
        MTRACE(4);
        // ---- P4: x += y; RMSNorm; gate/up strip pairs + SwiGLU -> act (bf16 pairs)
        gg = *reinterpret_cast<const float4*>(L.ln2 + k4);
        wait_cnt(cnt + CNT_Y, OP_ITEMS * W2, rc.t_cnt);
        add_rows<NT, RA, 1>(xr, p.rows, y_ll, 0, tg + 2);
        MTRACE(5);
        norm_to_xs<NT, RA>(xr, p.rows, xs, gg, p.eps, red);
        {
            const int i = deal.first();   // one or two (gate, up) strip pairs per CTA (256 pairs)
            auto swiglu = [](float g, float u) { return g / (1.f + expf(-g)) * u; };
            if (i + G < GU_ITEMS) {   // slots: gate_i, up_i, gate_{i+G}, up_{i+G}
                const float* pb = round_mma<NT, 4, NSLOTS, XR>(rc, xs, LDX, 0);
                {   // 2 * 8 * R <= 256 threads: item tid / (8R), strip rows (f, f+1), compact row r
                    const int it = tid / (8 * R), f = (tid & 7) * 2, r = (tid % (8 * R)) >> 3;
                    const bool act = tid < 2 * 8 * R && r < p.rows;
                    if (act) {
                        const float a0 = swiglu(part_sum<NT, 4>(pb, 2 * it, f, r), part_sum<NT, 4>(pb, 2 * it + 1, f, r));
                        const float a1 = swiglu(part_sum<NT, 4>(pb, 2 * it, f + 1, r), part_sum<NT, 4>(pb, 2 * it + 1, f + 1, r));
                        ll_store(act_ll + (size_t)r * (FFN / 2) + ((i + it * G) * 16 + f) / 2, pack_bf16(a0, a1), tg + 3);
                    }
                    warp_signal(act, cnt + CNT_ACT + ((i + (it & 1) * G) >> 7));
                }
                retire(4);
            } else if (i < GU_ITEMS) {
                const float* pb = round_mma<NT, 2, NSLOTS, XR>(rc, xs, LDX, 0);
                {
                    const int f = (tid & 7) * 2, r = tid >> 3;
                    const bool act = tid < 8 * R && r < p.rows;
                    if (act) {
                        const float a0 = swiglu(part_sum<NT, 2>(pb, 0, f, r), part_sum<NT, 2>(pb, 1, f, r));
                        const float a1 = swiglu(part_sum<NT, 2>(pb, 0, f + 1, r), part_sum<NT, 2>(pb, 1, f + 1, r));
                        ll_store(act_ll + (size_t)r * (FFN / 2) + (i * 16 + f) / 2, pack_bf16(a0, a1), tg + 3);
                    }
                    warp_signal(act, cnt + CNT_ACT + (i >> 7));
                }
                retire(2);
            }
        }
        deal.next_phase(GU_ITEMS);
        MTRACE(6);
        // ---- P5: down-proj K-halves -> z[half]
        {
            const int i = deal.first();
            if (i < DN_ITEMS) {   // at most one per CTA (128 items)
                const int strip = i >> 1, kh = i & 1;
                wait_cnt(cnt + CNT_ACT + kh, (GU_ITEMS / 2) * W4, rc.t_cnt);
                stage_act_rows<NT, XR>(p, xs, act_ll, kh, tg + 3);
                MTRACE(15);
                const float* pb = round_mma<NT, 2, NSLOTS, XR>(rc, xs, LDX2, D);
                if (pr < p.rows) ll_store(z_ll + ((size_t)kh * 16 + pr) * D + strip * 16 + pf, __float_as_uint(part_sum<NT, 2>(pb, 0, pf, pr) + part_sum<NT, 2>(pb, 1, pf, pr)), tg + 4);
                warp_signal(pr < p.rows, cnt + CNT_Z);
                retire(2);
            }
            deal.next_phase(DN_ITEMS);
        }
        MTRACE(7);
    }
    // ---- head: x += last down partials; final RMSNorm; logits (plain stores: consumed by the next kernel)
    {
        const int lastpar = (p.n_layers - 1) & 1;
        const uint32_t tg = tag0 + (uint32_t)p.n_layers * PHASES;
        const float4 gg = *reinterpret_cast<const float4*>(p.final_norm + k4);
        wait_cnt(p.cnt + (p.n_layers - 1) * CNT_STRIDE + CNT_Z, DN_ITEMS * W2, rc.t_cnt);
        add_rows<NT, RA, 2>(xr, p.rows, p.ll_z + (size_t)lastpar * 2 * 16 * D, 16 * D, tg - 1);
        norm_to_xs<NT, RA>(xr, p.rows, xs, gg, p.eps, red);
        float* out = p.logits + (size_t)sc_.row[pr < p.rows ? pr : 0] * p.ld_logits;
        int i = deal.first();
        for (; i + 3 * G < p.head_items; i += 4 * G) {
            const float* pb = round_mma<NT, 4, NSLOTS, XR>(rc, xs, LDX, 0);
            if (pr < p.rows) {
#pragma unroll
                for (int n = 0; n < 4; n++) { const int c = (i + n * G) * 16 + pf; if (c < p.vocab) out[c] = part_sum<NT, 4>(pb, n, pf, pr); }
            }
            retire(4);
        }
        for (; i < p.head_items; i += G) {
            const float* pb = round_mma<NT, 1, NSLOTS, XR>(rc, xs, LDX, 0);
            const int c = i * 16 + pf;
            if (pr < p.rows && c < p.vocab) out[c] = part_sum<NT, 1>(pb, 0, pf, pr);
            retire(1);
        }
        // every CTA has read the epoch long before CTA 0 can get here (CTA 0 needed all of their layer outputs)
        if (cta == 0 && tid == 0) *p.epoch = tag0 + (uint32_t)(p.n_layers + 1) * PHASES;
        if (tid == 0 && cta < 256) { g_mega_prof[cta * 4] = rc.t_mbar; g_mega_prof[cta * 4 + 1] = rc.t_cnt; g_mega_prof[cta * 4 + 2] = clock64() - t_start; }
    }
}

template <int NT, int RA> size_t mega_smem(int max_pages) {
    constexpr int R = 8 * NT;
    typedef Cfg<NT, RA> C;
    return (size_t)C::NSLOTS * SLOT + (size_t)C::XR * LDX2 * 2 + (size_t)C::NPART * 8 * 16 * R * 4 + 1024 * 4 + R * 8 * 4 + 2 * R * 32 * 4 +
           C::NSLOTS * 8 + MAX_LAYERS * sizeof(MegaLayer) + SCHED_MAX * 8 + 3 * R * 4 + (size_t)R * max_pages * 4 + 64;
}

}  // namespace



// host mirror of the kernel's Deal walk: the slot addresses of every CTA in consumption order
static void build_schedule(int G, const std::vector<MegaLayer>& layers, const bf16* head_f, int head_items, std::vector<unsigned long long>& sched, std::vector<int>& count) {
    sched.assign((size_t)G * SCHED_MAX, 0ull);
    count.assign(G, 0);
    for (int cta = 0; cta < G; cta++) {
        int off = 0, n = 0;
        auto first = [&]() { int f = cta - off; return f < 0 ? f + G : f; };
        auto put = [&](const void* base, size_t slot_index) {
            CBX_REQUIRE(n < SCHED_MAX, "t3 megakernel: weight schedule too long for one CTA");
            sched[(size_t)cta * SCHED_MAX + n++] = (unsigned long long)(reinterpret_cast<const char*>(base) + slot_index * SLOT);
        };
        for (const MegaLayer& L : layers) {
            for (int i = first(); i < QKV_ITEMS; i += G) put(L.wqkv_f, i);
            off = (off + QKV_ITEMS) % G;
            for (int i = first(); i < OP_ITEMS; i += G) put(L.wo_f, i);
            off = (off + OP_ITEMS) % G;
            for (int i = first(); i < GU_ITEMS; i += G) { put(L.wgu_f, 2 * (size_t)i); put(L.wgu_f, 2 * (size_t)i + 1); }   // strips 2i (gate), 2i+1 (up)
            off = (off + GU_ITEMS) % G;
            for (int i = first(); i < DN_ITEMS; i += G) {   // strip i/2 (4 slots), K-half i%2 (2 slots)
                const size_t s0 = (size_t)(i >> 1) * 4 + (size_t)(i & 1) * 2;
                put(L.wd_f, s0); put(L.wd_f, s0 + 1);
            }
            off = (off + DN_ITEMS) % G;
        }
        for (int i = first(); i < head_items; i += G) put(head_f, i);
        count[cta] = n;
    }
}


// Per-ENGINE state (the schedule holds that engine's weight addresses): returned to the caller, never kept in globals.
bool t3_mega_init(int max_pages, const std::vector<MegaLayer>& layers, const bf16* head_f, int head_items, MegaState* out) {
    *out = MegaState{};
    int dev = 0, sms = 0, coop = 0;
    CBX_CHECK(cudaGetDevice(&dev));
    CBX_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CBX_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return false;
    if (mega_smem<2, 16>(max_pages) > 227 * 1024 || mega_smem<1, 2>(max_pages) > 227 * 1024 || mega_smem<1, 8>(max_pages) > 227 * 1024) return false;
    out->max_pages = max_pages;
    CBX_CHECK(cudaFuncSetAttribute(t3_mega_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega_smem<1, 2>(max_pages)));
    CBX_CHECK(cudaFuncSetAttribute(t3_mega_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega_smem<1, 8>(max_pages)));
    CBX_CHECK(cudaFuncSetAttribute(t3_mega_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega_smem<2, 16>(max_pages)));
    int occ0 = 0, occ1 = 0, occ2 = 0;
    CBX_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, t3_mega_kernel<1, 2>, THREADS, mega_smem<1, 2>(max_pages)));
    CBX_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, t3_mega_kernel<1, 8>, THREADS, mega_smem<1, 8>(max_pages)));
    CBX_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, t3_mega_kernel<2, 16>, THREADS, mega_smem<2, 16>(max_pages)));
    if (occ0 < 1 || occ1 < 1 || occ2 < 1) return false;
    out->grid = sms;
    std::vector<unsigned long long> sched; std::vector<int> count;
    build_schedule(sms, layers, head_f, head_items, sched, count);
    CBX_CHECK(cudaMalloc(&out->sched, sched.size() * 8));
    CBX_CHECK(cudaMalloc(&out->sched_count, count.size() * 4));
    CBX_CHECK(cudaMemcpy(out->sched, sched.data(), sched.size() * 8, cudaMemcpyHostToDevice));
    CBX_CHECK(cudaMemcpy(out->sched_count, count.data(), count.size() * 4, cudaMemcpyHostToDevice));
    return true;
}

size_t t3_mega_ll_words(int which) {   // sizes (in 8-byte words) of the double-buffered exchange areas
    switch (which) {
        case 0: return 2ull * 16 * 3 * D;               // qkv
        case 1: return 2ull * 16 * H * NS_MAX * AP;     // attention partials
        case 2: return 2ull * 16 * D;                   // y
        case 3: return 2ull * 16 * (FFN / 2);           // act (bf16 pairs)
        case 4: return 2ull * 2 * 16 * D;               // z K-half partials
        default: return (size_t)MAX_LAYERS * CNT_STRIDE / 2;   // arrival counters (u32)
    }
}

void t3_mega_free(MegaState* ms) {
    if (ms->sched) cudaFree(ms->sched);
    if (ms->sched_count) cudaFree(ms->sched_count);
    *ms = MegaState{};
}

void launch_t3_mega(const MegaParams& p_in, const MegaState& ms, cudaStream_t st) {
    MegaParams p = p_in;
    p.sched = ms.sched; p.sched_count = ms.sched_count;
    static const int l2_ahead = [] { const char* e = getenv("CBX_T3_L2_AHEAD"); return e ? atoi(e) : 0; }();
    p.l2_ahead = l2_ahead;
    CBX_REQUIRE(ms.grid > 0, "t3 megakernel not initialised");
    CBX_REQUIRE(p.rows >= 1 && p.rows <= 16, "t3 megakernel: rows must be in [1,16]");
    CBX_REQUIRE(p.n_layers <= MAX_LAYERS, "t3 megakernel: too many layers");
    CBX_REQUIRE(p.max_pages == ms.max_pages, "t3 megakernel: page-table width changed after init");
    ProfScope ps(PC_GEMV, 2.0 * (p.n_layers * 16777216.0 + (double)p.head_items * 16 * D), st);
    CBX_CHECK(cudaMemsetAsync(p.cnt, 0, (size_t)MAX_LAYERS * CNT_STRIDE * 4, st));
    void* args[] = {(void*)&p};
    // cooperative launch: the flagged-word exchange needs every CTA of the grid to be resident
    if (p.rows <= 2) CBX_CHECK(cudaLaunchCooperativeKernel((void*)t3_mega_kernel<1, 2>, dim3(ms.grid), dim3(THREADS), args, mega_smem<1, 2>(ms.max_pages), st));
    else if (p.rows <= 8) CBX_CHECK(cudaLaunchCooperativeKernel((void*)t3_mega_kernel<1, 8>, dim3(ms.grid), dim3(THREADS), args, mega_smem<1, 8>(ms.max_pages), st));
    else CBX_CHECK(cudaLaunchCooperativeKernel((void*)t3_mega_kernel<2, 16>, dim3(ms.grid), dim3(THREADS), args, mega_smem<2, 16>(ms.max_pages), st));
}

extern "C" int cbx_t3_mega_prof(long long* out_h) { return cudaMemcpyFromSymbol(out_h, g_mega_prof, sizeof(long long) * 256 * 4) == cudaSuccess ? 0 : 1; }
extern "C" int cbx_t3_mega_trace(unsigned long long* out_h) { return cudaMemcpyFromSymbol(out_h, g_mega_trace, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : 1; }
