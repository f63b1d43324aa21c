// Fused softmax(QK^T)V for head_dim 64 (sm_100a): CFM transformer blocks (full attention),
// conformer encoder (full attention + ESPnet relative-position bias) and T3 prefill (causal).
// One CTA = 64 queries of one (batch, head); K/V streamed in 64-key tiles through a cp.async
// double buffer; scores and the output accumulator never leave registers.
#include "common.cuh"

namespace {

constexpr int D = 64, BQ = 64, BKV = 128, NT = 256, KVS = 3;   // 8 warps: 4 query sub-tiles x 2 key halves of a 128-key stage
constexpr int TILE_BYTES = 64 * 128;   // 64 rows x 128 B
constexpr int STAGE_BYTES = 2 * TILE_BYTES;

__device__ __forceinline__ uint32_t swz128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// rows [row0, row0 + nrows) of a [*][64] bf16 matrix -> 128-byte swizzled rows (zero-filled past nrows_valid)
template <int NROWS>
__device__ __forceinline__ void load_rows(uint32_t sdst, const bf16* base, long ld, int row0, int nrows_valid, int tid) {
#pragma unroll
    for (int i = 0; i < NROWS * 8 / NT; i++) {
        int c = tid + i * NT, row = c >> 3, ch = c & 7;
        bool ok = (row0 + row) < nrows_valid;
        const bf16* src = base + (long)(ok ? row0 + row : 0) * ld + ch * 8;
        cp_async16(sdst + swz128(row, ch), src, ok ? 16 : 0);
    }
}

// mma.sync issue rate on sm_100a is the bound of this kernel (~32 cycles per m16n8k16 per scheduler), so every scheduler
// gets two warps: warp w handles query rows (w&3)*16.. and keys (w>>2)*64.. of each 128-key stage; the two key halves
// are merged through shared memory at the end.
__global__ void __launch_bounds__(NT) attn_kernel(const AttnParams p) {
    pdl_prologue();
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qw = warp & 3, kvh = warp >> 2;
    const int g = lane >> 2, tg = lane & 3;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * BQ;
    const int klen_raw = p.kv_len[(b / p.kv_div) & 15];
    const int Tk = klen_raw > 0 ? klen_raw : p.T;   // valid keys (and valid queries) of this sequence
    if (q0 >= Tk) return;                           // padded query tile: its rows are never consumed
    const bf16* Q = p.q + (long)b * p.q_bs + h * D;
    const bf16* K = p.k + (long)b * p.k_bs + h * D;
    const bf16* V = p.v + (long)b * p.v_bs + h * D;
    const uint32_t sQ = smem_u32(smem), sK = sQ + TILE_BYTES, sV = sK + KVS * STAGE_BYTES;

    int n_stages = (Tk + BKV - 1) / BKV;
    if (p.causal) n_stages = (min(Tk, q0 + BQ) + BKV - 1) / BKV;
    load_rows<64>(sQ, Q, p.ldq, q0, p.T, tid);
#pragma unroll
    for (int s = 0; s < KVS - 1; s++) {
        if (s < n_stages) {
            load_rows<128>(sK + s * STAGE_BYTES, K, p.ldk, s * BKV, Tk, tid);
            load_rows<128>(sV + s * STAGE_BYTES, V, p.ldv, s * BKV, Tk, tid);
        }
        cp_async_commit();
    }

    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int r = 0; r < 4; r++) o[i][r] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    uint32_t qf[4][4];
    const float sl2 = p.scale * 1.4426950408889634f;  // scores kept in log2 domain
    const int qi0 = q0 + qw * 16 + g, qi1 = qi0 + 8;

    for (int st = 0; st < n_stages; st++) {
        const int buf = st % KVS;
        {
            const int nk = st + KVS - 1;
            if (nk < n_stages) {
                load_rows<128>(sK + (nk % KVS) * STAGE_BYTES, K, p.ldk, nk * BKV, Tk, tid);
                load_rows<128>(sV + (nk % KVS) * STAGE_BYTES, V, p.ldv, nk * BKV, Tk, tid);
            }
            cp_async_commit();
        }
        cp_async_wait<KVS - 1>();
        __syncthreads();
        if (st == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) ldmatrix_x4(qf[ks], sQ + swz128(qw * 16 + (lane & 15), ks * 2 + (lane >> 4)));
        }
        const int kbase = st * BKV + kvh * 64;
        const bool live = kbase < Tk && !(p.causal && kbase > q0 + qw * 16 + 15);   // warp-uniform: any key of this half visible?
        if (live) {
            const uint32_t sk = sK + buf * STAGE_BYTES + kvh * TILE_BYTES, sv = sV + buf * STAGE_BYTES + kvh * TILE_BYTES;
            float s[8][4];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int r = 0; r < 4; r++) s[i][r] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    uint32_t kf[4];
                    int row = j * 8 + (lane & 7) + ((lane >> 4) << 3);
                    ldmatrix_x4(kf, sk + swz128(row, ks * 2 + ((lane >> 3) & 1)));
                    mma_bf16(s[j], qf[ks], kf[0], kf[1]);
                    mma_bf16(s[j + 1], qf[ks], kf[2], kf[3]);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    int kj = kbase + j * 8 + tg * 2 + (r & 1);
                    int qi = (r < 2) ? qi0 : qi1;
                    float v = s[j][r];
                    if (p.relbias && kj < Tk && qi < p.T)
                        v += p.relbias[(long)b * p.rb_bs + (long)h * p.rb_hs + (long)qi * p.rb_ld + (p.T - 1 - qi + kj)];
                    v *= sl2;
                    if (kj >= Tk || (p.causal && kj > qi)) v = -INFINITY;
                    s[j][r] = v;
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 8; j++) mx = fmaxf(mx, fmaxf(s[j][rr * 2], s[j][rr * 2 + 1]));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                float mnew = fmaxf(mrow[rr], mx);
                float msafe = (mnew == -INFINITY) ? 0.f : mnew;
                float corr = exp2f(mrow[rr] - msafe);
                mrow[rr] = mnew;
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float e0 = exp2f(s[j][rr * 2] - msafe), e1 = exp2f(s[j][rr * 2 + 1] - msafe);
                    s[j][rr * 2] = e0; s[j][rr * 2 + 1] = e1;
                    sum += e0 + e1;
                }
                lrow[rr] = lrow[rr] * corr + sum;
#pragma unroll
                for (int i = 0; i < 8; i++) { o[i][rr * 2] *= corr; o[i][rr * 2 + 1] *= corr; }
            }
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                uint32_t pf[4];
                pf[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
                pf[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
                pf[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
                pf[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
                for (int dj = 0; dj < 8; dj += 2) {
                    uint32_t vf[4];
                    int row = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                    ldmatrix_x4_trans(vf, sv + swz128(row, dj + (lane >> 4)));
                    mma_bf16(o[dj], pf, vf[0], vf[1]);
                    mma_bf16(o[dj + 1], pf, vf[2], vf[3]);
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    // quad-reduce the row sums, then merge the two key halves (warps w and w+4) through shared memory
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        lrow[rr] += __shfl_xor_sync(0xffffffffu, lrow[rr], 1);
        lrow[rr] += __shfl_xor_sync(0xffffffffu, lrow[rr], 2);
    }
    float* mg = reinterpret_cast<float*>(smem + TILE_BYTES) + (qw * 32 + lane) * 37;   // 36 floats per thread (+1 pad)
    if (kvh == 1) {
        mg[0] = mrow[0]; mg[1] = mrow[1]; mg[2] = lrow[0]; mg[3] = lrow[1];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int r = 0; r < 4; r++) mg[4 + i * 4 + r] = o[i][r];
    }
    __syncthreads();
    if (kvh == 0) {
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const float m1 = mg[rr], l1 = mg[2 + rr];
            const float M = fmaxf(mrow[rr], m1);
            const float Ms = (M == -INFINITY) ? 0.f : M;
            const float c0 = exp2f(mrow[rr] - Ms), c1 = exp2f(m1 - Ms);
            const float l = lrow[rr] * c0 + l1 * c1;
            const float inv = l > 0.f ? 1.f / l : 0.f;
            const int qi = rr ? qi1 : qi0;
            if (qi < p.T) {
                bf16* orow = p.o + (long)b * p.o_bs + (long)qi * p.ldo + h * D;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float a0 = (o[i][rr * 2] * c0 + mg[4 + i * 4 + rr * 2] * c1) * inv;
                    const float a1 = (o[i][rr * 2 + 1] * c0 + mg[4 + i * 4 + rr * 2 + 1] * c1) * inv;
                    *reinterpret_cast<uint32_t*>(orow + i * 8 + tg * 2) = pack_bf16(a0, a1);
                }
            }
        }
    }
}

}  // namespace

void attention_init() { attention_fa_init(); attention_tc_init(); CBX_CHECK(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES + 2 * KVS * STAGE_BYTES)); }

void launch_attention(const AttnParams& p, cudaStream_t st) {
    CBX_REQUIRE(p.T > 0 && p.H > 0 && p.batch > 0, "attention: empty problem");
    CBX_REQUIRE(p.ldq % 8 == 0 && p.ldk % 8 == 0 && p.ldv % 8 == 0 && p.ldo % 2 == 0, "attention: row strides must keep 16B alignment");
    if (launch_attention_fa(p, st)) return;   // tcgen05 kernel, second generation: full / causal attention without bias (CFM blocks, T3 prefill)
    if (launch_attention_tc(p, st)) return;   // first-generation tcgen05 kernel (CBX_DISABLE_ATTN_FA=1)
    const int smem = TILE_BYTES + 2 * KVS * STAGE_BYTES;
    ProfScope ps(PC_ATTN, 4.0 * p.T * p.T * D * p.H * p.batch * (p.causal ? 0.5 : 1.0), st);
    dim3 grid(cdiv(p.T, BQ), p.H, p.batch);
    launch_pdl(attn_kernel, dim3(grid), dim3(NT), smem, st, p);
    CBX_CHECK(cudaGetLastError());
}
