// S3Gen token->mel: UpsampleConformerEncoder + CausalConditionalCFM (10 Euler steps, CFG-batched
// ConditionalDecoder).  Activations are time-major channels-last; every conv is an implicit GEMM over
// zero-haloed bf16 buffers; residual streams stay fp32.  Host orchestration only.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "engine.h"

using namespace dims;

static Lin reg_lin(cbx_engine* e, const std::string& n, int N, int K, bool bias = true) {
    Lin l; l.N = N; l.K = K;
    l.w = e->reg<bf16>(n + ".w", DT_BF16, (long)N * K);
    l.b = bias ? e->reg<float>(n + ".b", DT_F32, N) : nullptr;
    return l;
}
static LNp reg_ln(cbx_engine* e, const std::string& n, int C) {
    LNp l; l.g = e->reg<float>(n + ".g", DT_F32, C); l.b = e->reg<float>(n + ".b", DT_F32, C); return l;
}

void flow_build(cbx_engine* e) {
    FlowModel& f = e->flow;
    const cbx_config& c = e->cfg;
    f.tok_emb = e->reg<float>("flow.tok_emb", DT_F32, (long)F_V * F_D);
    f.spk_w = e->reg<float>("flow.spk.w", DT_F32, (long)MEL * F_SPK);
    f.spk_b = e->reg<float>("flow.spk.b", DT_F32, MEL);
    f.enc_proj = reg_lin(e, "flow.enc_proj", MEL, F_D);
    f.embed = reg_lin(e, "flow.embed", F_D, F_D); f.embed_ln = reg_ln(e, "flow.embed_ln", F_D);
    f.up_embed = reg_lin(e, "flow.up_embed", F_D, F_D); f.up_embed_ln = reg_ln(e, "flow.up_embed_ln", F_D);
    f.pl1 = reg_lin(e, "flow.pl1", F_D, 4 * F_D); f.pl2 = reg_lin(e, "flow.pl2", F_D, 3 * F_D); f.upconv = reg_lin(e, "flow.upconv", F_D, 5 * F_D);
    f.after_norm = reg_ln(e, "flow.after_norm", F_D);
    auto conf = [&](const std::string& p) {
        ConformerLayer l;
        l.nm = reg_ln(e, p + "nm", F_D); l.nf = reg_ln(e, p + "nf", F_D);
        l.qkv4 = reg_lin(e, p + "qkv4", 4 * F_D, F_D); l.pos = reg_lin(e, p + "pos", F_D, F_D, false); l.out = reg_lin(e, p + "out", F_D, F_D);
        l.w1 = reg_lin(e, p + "w1", F_FFN, F_D); l.w2 = reg_lin(e, p + "w2", F_D, F_FFN);
        return l;
    };
    for (int i = 0; i < c.enc_blocks; i++) f.enc.push_back(conf("flow.enc" + std::to_string(i) + "."));
    for (int i = 0; i < c.up_blocks; i++) f.up.push_back(conf("flow.up" + std::to_string(i) + "."));
    f.t1 = reg_lin(e, "cfm.t1", C_TDIM, C_IN); f.t2 = reg_lin(e, "cfm.t2", C_TDIM, C_TDIM);
    const int nres = c.cfm_mid + 2;
    for (int r = 0; r < nres; r++) {
        std::string p = "cfm.r" + std::to_string(r) + ".";
        int cin = r == 0 ? C_IN : (r == nres - 1 ? 2 * C_CH : C_CH);
        ResnetP rp; rp.cin = cin;
        f.tmlp.push_back(reg_lin(e, p + "tmlp", C_CH, C_TDIM));
        rp.c1 = reg_lin(e, p + "c1", C_CH, 3 * cin); rp.n1 = reg_ln(e, p + "n1", C_CH);
        rp.c2 = reg_lin(e, p + "c2", C_CH, 3 * C_CH); rp.n2 = reg_ln(e, p + "n2", C_CH);
        rp.res = reg_lin(e, p + "res", C_CH, cin);
        f.resnets.push_back(rp);
        for (int j = 0; j < c.cfm_blocks; j++) {
            std::string q = p + "t" + std::to_string(j) + ".";
            TfmP t; t.n1 = reg_ln(e, q + "n1", C_CH); t.qkv = reg_lin(e, q + "qkv", 3 * C_INNER, C_CH, false);
            t.out = reg_lin(e, q + "out", C_CH, C_INNER); t.n3 = reg_ln(e, q + "n3", C_CH);
            t.ff0 = reg_lin(e, q + "ff0", C_FF, C_CH); t.ff2 = reg_lin(e, q + "ff2", C_CH, C_FF);
            f.tfms.push_back(t);
        }
    }
    f.down_conv = reg_lin(e, "cfm.down_conv", C_CH, 3 * C_CH); f.up_conv = reg_lin(e, "cfm.up_conv", C_CH, 3 * C_CH);
    f.final_conv = reg_lin(e, "cfm.final_conv", C_CH, 3 * C_CH); f.final_ln = reg_ln(e, "cfm.final_ln", C_CH);
    f.final_proj = reg_lin(e, "cfm.final_proj", MEL, C_CH);
    f.noise = e->reg<float>("cfm.noise", DT_F32, (long)NOISE_LEN * MEL);
    f.tproj = e->scratch<float>((long)nres * c.cfm_steps * C_CH);
}

// plain linear / conv GEMM helper.  A rows are `lda` apart; `taps` taps of `kc` channels `tap_stride` apart.
static GemmParams mk(const Lin& l, const bf16* A, long lda, int M, int kc, long tap_stride) {
    GemmParams g; g.A = A; g.lda = lda; g.kc = kc; g.tap_stride = tap_stride; g.W = l.w; g.ldw = l.K; g.M = M; g.N = l.N; g.K = l.K; g.bias = l.b;
    return g;
}

void flow_finalize(cbx_engine* e, cudaStream_t st) {
    FlowModel& f = e->flow;
    const int n = e->cfg.cfm_steps, half = C_IN / 2;
    // cosine schedule and the fp32 recurrence of solve_euler (t += dt; dt = t_span[k+1] - t)
    std::vector<float> ts(n + 1);
    for (int i = 0; i <= n; i++) ts[i] = 1.f - cosf((float)i / n * 0.5f * 3.14159265358979323846f);
    f.t_span.assign(2 * n, 0.f);   // [t_k..., dt_k...]
    float t = ts[0], dt = ts[1] - ts[0];
    for (int k = 0; k < n; k++) {
        f.t_span[k] = t; f.t_span[n + k] = dt;
        t = t + dt;
        if (k + 1 < n) dt = ts[k + 2] - t;
    }
    std::vector<float> sin_h((size_t)n * C_IN);
    for (int k = 0; k < n; k++)
        for (int i = 0; i < half; i++) {
            float fr = expf((float)i * -(logf(10000.f) / (half - 1)));
            float a = 1000.f * f.t_span[k] * fr;
            sin_h[(size_t)k * C_IN + i] = sinf(a);
            sin_h[(size_t)k * C_IN + half + i] = cosf(a);
        }
    float* sin_d = e->scratch<float>((long)n * C_IN);
    bf16* sin_b = e->scratch<bf16>((long)n * C_IN);
    bf16* h1 = e->scratch<bf16>((long)n * C_TDIM);
    bf16* h2 = e->scratch<bf16>((long)n * C_TDIM);
    CBX_CHECK(cudaMemcpyAsync(sin_d, sin_h.data(), sin_h.size() * 4, cudaMemcpyHostToDevice, st));
    launch_f32_to_bf16_rows(sin_d, C_IN, sin_b, C_IN, n, C_IN, ACT_NONE, 0.f, st);
    GemmParams g = mk(f.t1, sin_b, C_IN, n, C_IN, 0); g.act = ACT_SILU; g.outB = h1; g.ldc = C_TDIM; launch_gemm(g, st);
    g = mk(f.t2, h1, C_TDIM, n, C_TDIM, 0); g.act = ACT_MISH; g.outB = h2; g.ldc = C_TDIM; launch_gemm(g, st);   // resnets consume mish(temb)
    for (size_t r = 0; r < f.tmlp.size(); r++) {
        g = mk(f.tmlp[r], h2, C_TDIM, n, C_TDIM, 0); g.outF = f.tproj + (long)r * n * C_CH; g.ldc = C_CH; launch_gemm(g, st);
    }
    CBX_CHECK(cudaStreamSynchronize(st));
    if (cfm_tail_available())
        for (auto& t : f.tfms) cfm_tail_weights(t.tw, t.out.w, t.ff0.w, t.ff2.w, t.qkv.w);     // TMA descriptors of the static weights
}

static constexpr int CH = 2;   // causal halo rows (k=3)
static constexpr int EP = 8;   // pad rows per sequence slab of the encoder conv inputs (look-ahead 3 / left pads 2, 4)

// A lane runs `nb` calls at once (nb = 1: the classic single call).  Sequences are right-padded to the longest one:
// Ttm tokens / T = 2*Ttm frames.  Padding is exact: the estimator's convolutions are causal, LayerNorm / linears are
// per frame, attention masks keys beyond each sequence's length, and the one look-ahead conv of the encoder reads
// rows that are explicitly zeroed behind each sequence (the reference right-pads with zeros).
void lane_alloc(cbx_engine* e, Lane& L, int bmax) {
    const cbx_config& c = e->cfg;
    const long B = bmax;
    L.bmax = bmax;
    const long Tt = c.max_prompt_tokens + c.max_s3_tokens, T = 2 * Tt, Tg = 2L * c.max_s3_tokens;
    {   // CBX_S3GEN_PRIORITY=1: S3Gen lanes at the highest stream priority (their CTAs are placed before pending T3 ones)
        int lo = 0, hi = 0;
        CBX_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const char* pr = getenv("CBX_S3GEN_PRIORITY");
        CBX_CHECK(cudaStreamCreateWithPriority(&L.st, cudaStreamNonBlocking, (pr && pr[0] == '1') ? hi : lo));
    }
    CBX_CHECK(cudaEventCreateWithFlags(&L.ev_in, cudaEventDisableTiming));
    CBX_CHECK(cudaEventCreateWithFlags(&L.ev_out, cudaEventDisableTiming));
    if (bmax > 0) {   // bmax == 0: a vocoder-only workspace (parallel HiFT streams of a batch lane)
        L.tok = e->scratch<int>(B * Tt);
        L.e_in = e->scratch<bf16>(B * T * F_D); L.e_xb = e->scratch<bf16>(B * (T + EP) * F_D); L.e_y1 = e->scratch<bf16>(B * (T + EP) * F_D);
        L.e_xn = e->scratch<bf16>(B * T * F_D); L.e_qkv = e->scratch<bf16>(B * T * 4 * F_D); L.e_pos = e->scratch<bf16>(2 * T * F_D); L.e_p = e->scratch<bf16>(2 * T * F_D);
        L.e_o = e->scratch<bf16>(B * T * F_D); L.e_ff = e->scratch<bf16>(B * T * F_FFN); L.e_up = e->scratch<bf16>(B * (T + EP) * F_D); L.e_upc = e->scratch<bf16>(B * T * F_D);
        L.e_tmp = e->scratch<float>(B * T * F_D); L.e_x = e->scratch<float>(B * T * F_D); L.e_bd = e->scratch<float>(B * F_H * T * 2 * T);
        L.mu = e->scratch<float>(B * T * MEL); L.cond = e->scratch<float>(B * T * MEL); L.x = e->scratch<float>(B * T * MEL); L.v = e->scratch<float>(2 * B * T * MEL);
        const long TH = T + CH;
        L.c_in = e->scratch<bf16>(2 * B * TH * C_IN); L.c_hb = e->scratch<bf16>(2 * B * TH * C_CH); L.c_inb = e->scratch<bf16>(2 * B * TH * C_CH);
        L.c_upin = e->scratch<bf16>(2 * B * TH * 2 * C_CH); L.c_xn = e->scratch<bf16>(2 * B * T * C_CH); L.c_qkv = e->scratch<bf16>(2 * B * T * 3 * C_INNER);
        L.c_o = e->scratch<bf16>(2 * B * T * C_INNER); L.c_ff = e->scratch<bf16>(2 * B * T * C_FF); L.c_fb = e->scratch<bf16>(2 * B * T * C_CH);
        L.c_tmp = e->scratch<float>(2 * B * T * C_CH); L.c_tmp2 = e->scratch<float>(2 * B * T * C_CH); L.c_h = e->scratch<float>(2 * B * T * C_CH);
        L.melb = e->scratch<float>(B * Tg * MEL);
    }
    // hift (one call at a time)
    const long HH = H_HALO;
    L.mel = e->scratch<float>(Tg * MEL); L.h_mel = e->scratch<bf16>((Tg + 2 * HH) * MEL_PAD);     // channels 80..127 stay zero: K tiles of 64 for the tcgen05 conv
    L.h_f0a = e->scratch<bf16>((Tg + 2 * HH) * H_F0CH); L.h_f0b = e->scratch<bf16>((Tg + 2 * HH) * H_F0CH);
    L.h_f0 = e->scratch<float>(Tg); L.h_cum = e->scratch<double>(Tg * H_NHARM); L.h_s = e->scratch<float>(Tg * H_UP);
    const long F = 120 * Tg + 1;
    L.h_stft = e->scratch<bf16>((F + 2 * HH) * H_NSRC_PAD);
    long tlen[4] = {Tg, 8 * Tg, 40 * Tg, F};
    for (int i = 0; i < 4; i++) L.h_xb[i] = e->scratch<bf16>((tlen[i] + 2 * HH) * (H_BASE >> i));
    for (int i = 0; i < 3; i++) {
        long t = tlen[i + 1], ch = H_BASE >> (i + 1);
        L.h_x[i] = e->scratch<float>(t * ch); L.h_r[i] = e->scratch<float>(t * ch); L.h_acc[i] = e->scratch<float>(t * ch); L.h_si[i] = e->scratch<float>(t * ch);
        L.h_a[i] = e->scratch<bf16>((t + 2 * HH) * ch); L.h_b[i] = e->scratch<bf16>((t + 2 * HH) * ch);
    }
    L.h_post = e->scratch<float>(F * H_NSRC);
    L.h_phase = e->scratch<float>(16);
    L.g_wav = e->scratch<float>(Tg * H_UP); L.g_src = e->scratch<float>(Tg * H_UP); L.g_cache = e->scratch<float>(Tg * H_UP);
    L.g_dyn = e->scratch<SourceDyn>(1);
    CBX_CHECK(cudaMallocHost(&L.g_dyn_h, sizeof(SourceDyn)));
}

// per-sequence key lengths for the attention kernel (nothing to mask when every sequence fills the padded length)
static void set_kv_len(AttnParams& a, const Lane& L, int frames_per_token, int div) {
    bool ragged = false;
    for (int b = 0; b < L.nb; b++) ragged = ragged || L.call[b].Tt != L.Ttm;
    a.kv_div = div;
    if (ragged) for (int b = 0; b < L.nb; b++) a.kv_len[b] = frames_per_token * L.call[b].Tt;
}

// ---------------------------------------------------------------------------------------------- encoder
// rows: padded sequence length at this stage (Ttm before the upsample, 2*Ttm after); all activations are [nb][rows][C]
static void conformer_layer(cbx_engine* e, Lane& L, const ConformerLayer& l, int rows, int fpt, cudaStream_t st) {
    const int B = L.nb, M = B * rows;
    NormParams n; n.in = L.e_x; n.ld_in = F_D; n.rows = M; n.C = F_D; n.gain = l.nm.g; n.bias = l.nm.b; n.eps = 1e-12f; n.outB = L.e_xn; n.ld_outB = F_D;
    launch_norm(n, st);
    GemmParams g = mk(l.qkv4, L.e_xn, F_D, M, F_D, 0); g.outB = L.e_qkv; g.ldc = 4 * F_D; launch_gemm(g, st);        // [q+u | q+v | k | v]
    g = mk(l.pos, L.e_pos, F_D, 2 * rows - 1, F_D, 0); g.outB = L.e_p; g.ldc = F_D; launch_gemm(g, st);
    const long ldbd = 2L * rows, bd_bs = (long)F_H * rows * ldbd;
    for (int b = 0; b < B; b++) {   // (q+v) P^T per head: the head index is the GEMM batch, so calls are looped
        GemmParams bd; bd.A = L.e_qkv + (long)b * rows * 4 * F_D + F_D; bd.lda = 4 * F_D; bd.kc = 64; bd.a_bs = 64; bd.W = L.e_p; bd.ldw = F_D; bd.w_bs = 64;
        bd.M = rows; bd.N = 2 * rows - 1; bd.K = 64; bd.batch = F_H; bd.outF = L.e_bd + b * bd_bs; bd.ldc = ldbd; bd.c_bs = (long)rows * ldbd;
        launch_gemm(bd, st);
    }
    AttnParams a; a.q = L.e_qkv; a.k = L.e_qkv + 2 * F_D; a.v = L.e_qkv + 3 * F_D; a.ldq = a.ldk = a.ldv = 4 * F_D; a.o = L.e_o; a.ldo = F_D;
    a.q_bs = a.k_bs = a.v_bs = (long)rows * 4 * F_D; a.o_bs = (long)rows * F_D;
    a.T = rows; a.H = F_H; a.batch = B; a.scale = 0.125f; a.relbias = L.e_bd; a.rb_ld = ldbd; a.rb_hs = (long)rows * ldbd; a.rb_bs = bd_bs;
    set_kv_len(a, L, fpt, 1);
    launch_attention(a, st);
    g = mk(l.out, L.e_o, F_D, M, F_D, 0); g.res = L.e_x; g.ldr = F_D; g.outF = L.e_x; g.ldc = F_D; launch_gemm(g, st);
    n.gain = l.nf.g; n.bias = l.nf.b; launch_norm(n, st);
    g = mk(l.w1, L.e_xn, F_D, M, F_D, 0); g.act = ACT_SILU; g.outB = L.e_ff; g.ldc = F_FFN; launch_gemm(g, st);
    g = mk(l.w2, L.e_ff, F_FFN, M, F_FFN, 0); g.res = L.e_x; g.ldr = F_D; g.outF = L.e_x; g.ldc = F_D; launch_gemm(g, st);
    e->gpu_launches += 8 + B;
}

// Linear + LayerNorm + sqrt(d) scale: in [nb][rows][512] (contiguous) -> e_x fp32 (contiguous) and e_xb bf16 slabs of rows+EP
static void embed_stage(cbx_engine* e, Lane& L, const Lin& lin, const LNp& ln, const bf16* in, int rows, cudaStream_t st) {
    const int B = L.nb;
    GemmParams g = mk(lin, in, F_D, B * rows, F_D, 0); g.outF = L.e_tmp; g.ldc = F_D; launch_gemm(g, st);
    NormParams n; n.in = L.e_tmp; n.ld_in = F_D; n.in_bs = (long)rows * F_D; n.rows = rows; n.batch = B; n.C = F_D; n.gain = ln.g; n.bias = ln.b; n.eps = 1e-5f;
    n.out_scale = sqrtf((float)F_D);
    n.outF = L.e_x; n.ld_outF = F_D; n.outF_bs = (long)rows * F_D; n.outB = L.e_xb; n.ld_outB = F_D; n.outB_bs = (long)(rows + EP) * F_D;
    launch_norm(n, st);
    launch_relpos_table(L.e_pos, rows, F_D, st);
    e->gpu_launches += 3;
}

static void encoder(cbx_engine* e, Lane& L, cudaStream_t st) {
    FlowModel& f = e->flow;
    const int B = L.nb, Tt = L.Ttm, T = 2 * Tt;
    const long sl1 = (long)(Tt + EP) * F_D, sl2 = (long)(T + EP) * F_D;   // slab strides of the conv inputs
    launch_gather_rows_bf16(f.tok_emb, L.tok, B * Tt, F_D, L.e_in, F_D, st);
    embed_stage(e, L, f.embed, f.embed_ln, L.e_in, Tt, st);
    // PreLookaheadLayer: right-pad 3 -> conv k4 -> leaky_relu -> left-pad 2 -> conv k3 -> + x
    for (int b = 0; b < B; b++) CBX_CHECK(cudaMemsetAsync(L.e_xb + b * sl1 + (long)L.call[b].Tt * F_D, 0, 3L * F_D * 2, st));
    GemmParams g = mk(f.pl1, L.e_xb, F_D, Tt, F_D, F_D); g.batch = B; g.a_bs = sl1; g.act = ACT_LRELU; g.act_param = 0.01f; g.outB = L.e_y1 + 2 * F_D; g.ldc = F_D; g.c_bs = sl1;
    launch_gemm(g, st);
    for (int b = 0; b < B; b++) CBX_CHECK(cudaMemsetAsync(L.e_y1 + b * sl1, 0, 2L * F_D * 2, st));
    g = mk(f.pl2, L.e_y1, F_D, Tt, F_D, F_D); g.batch = B; g.a_bs = sl1; g.res = L.e_x; g.ldr = F_D; g.r_bs = (long)Tt * F_D; g.outF = L.e_x; g.ldc = F_D; g.c_bs = (long)Tt * F_D;
    launch_gemm(g, st);
    for (auto& l : f.enc) conformer_layer(e, L, l, Tt, 1, st);
    // Upsample1D: nearest x2 -> left-pad 4 -> conv k5
    for (int b = 0; b < B; b++) {
        CBX_CHECK(cudaMemsetAsync(L.e_up + b * sl2, 0, 4L * F_D * 2, st));
        launch_upsample2(L.e_x + (long)b * Tt * F_D, L.e_up + b * sl2 + 4 * F_D, F_D, Tt, F_D, st);
    }
    g = mk(f.upconv, L.e_up, F_D, T, F_D, F_D); g.batch = B; g.a_bs = sl2; g.outB = L.e_upc; g.ldc = F_D; g.c_bs = (long)T * F_D; launch_gemm(g, st);
    embed_stage(e, L, f.up_embed, f.up_embed_ln, L.e_upc, T, st);
    for (auto& l : f.up) conformer_layer(e, L, l, T, 2, st);
    NormParams n; n.in = L.e_x; n.ld_in = F_D; n.rows = B * T; n.C = F_D; n.gain = f.after_norm.g; n.bias = f.after_norm.b; n.eps = 1e-5f; n.outB = L.e_xn; n.ld_outB = F_D;
    launch_norm(n, st);
    g = mk(f.enc_proj, L.e_xn, F_D, B * T, F_D, 0); g.outF = L.mu; g.ldc = MEL; launch_gemm(g, st);
    e->gpu_launches += 6 + B;
}

// ---------------------------------------------------------------------------------------------- estimator
// all estimator tensors are [2*nb][T(+halo)][C] (rows 2b, 2b+1 = conditional / unconditional pass of call b)
// The two LayerNorms of a block CAN ride in the epilogue of the GEMM that produces their input (cluster-fused LN of
// gemm_tc.cu): `pre_normed` = c_xn already holds LN1(c_h) (written by the previous block's ff2), `next_ln` = the LN1 of
// the following block, applied by this block's ff2.  Measured on B200 it does not pay -- one call 37.1 ms fused vs 35.4 ms
// with the separate norm kernels, 8 x 140 tokens 127.6 vs 124.9 ms: under PDL the norm launches overlap their neighbours,
// while two cluster barriers and a second epilogue pass lengthen every residual GEMM -- so it is opt-in (CBX_FUSE_LN=1).
static bool fuse_ln() {
    const char* v = getenv("CBX_FUSE_LN");
    const char* tc = getenv("CBX_DISABLE_TC");
    return v && v[0] == '1' && !(tc && tc[0] == '1');   // the fused epilogue lives in the tcgen05 kernel
}
static void tfm_block(cbx_engine* e, Lane& L, const TfmP& t, int T, bool pre_normed, const LNp* next_ln, cudaStream_t st) {
    const int NB = 2 * L.nb;
    const long bs = (long)T * C_CH;
    NormParams n; n.in = L.c_h; n.ld_in = C_CH; n.in_bs = bs; n.rows = T; n.batch = NB; n.C = C_CH; n.gain = t.n1.g; n.bias = t.n1.b; n.eps = 1e-5f;
    n.outB = L.c_xn; n.ld_outB = C_CH; n.outB_bs = bs;
    if (!pre_normed) { launch_norm(n, st); e->gpu_launches += 1; }
    GemmParams g = mk(t.qkv, L.c_xn, C_CH, NB * T, C_CH, 0); g.outB = L.c_qkv; g.ldc = 3 * C_INNER; launch_gemm(g, st);
    AttnParams a; a.q = L.c_qkv; a.k = L.c_qkv + C_INNER; a.v = L.c_qkv + 2 * C_INNER; a.ldq = a.ldk = a.ldv = 3 * C_INNER;
    a.q_bs = a.k_bs = a.v_bs = (long)T * 3 * C_INNER; a.o = L.c_o; a.ldo = C_INNER; a.o_bs = (long)T * C_INNER; a.T = T; a.H = 8; a.batch = NB; a.scale = 0.125f;
    set_kv_len(a, L, 2, 2);
    launch_attention(a, st);
    g = mk(t.out, L.c_o, C_INNER, NB * T, C_INNER, 0); g.res = L.c_h; g.ldr = C_CH; g.outF = L.c_h; g.ldc = C_CH;
    if (fuse_ln()) { g.ln_gamma = t.n3.g; g.ln_beta = t.n3.b; g.ln_eps = 1e-5f; g.outB2 = L.c_xn; g.ldc2 = C_CH; }
    launch_gemm(g, st);
    if (!fuse_ln()) { n.gain = t.n3.g; n.bias = t.n3.b; launch_norm(n, st); e->gpu_launches += 1; }
    g = mk(t.ff0, L.c_xn, C_CH, NB * T, C_CH, 0); g.act = ACT_GELU; g.outB = L.c_ff; g.ldc = C_FF; launch_gemm(g, st);
    g = mk(t.ff2, L.c_ff, C_FF, NB * T, C_FF, 0); g.res = L.c_h; g.ldr = C_CH; g.outF = L.c_h; g.ldc = C_CH;
    if (next_ln) { g.ln_gamma = next_ln->g; g.ln_beta = next_ln->b; g.ln_eps = 1e-5f; g.outB2 = L.c_xn; g.ldc2 = C_CH; }
    launch_gemm(g, st);
    e->gpu_launches += 5;
}

// Rows from which the fused block tail (cfm_tail.cu: one kernel for out-proj + LN + GELU-FF + next LN + QKV per 128-row tile)
// replaces the six separate launches.  A tile is a ~40 us serial chain on ONE SM, so a single call (M ~ 900 rows = 8 tiles)
// is faster through the wide separate GEMMs; batches fill the SMs with tiles.
static int fused_tail_min_rows() {      // read per call (once per estimator stage): tests switch paths at run time
    const char* e = getenv("CBX_CFM_TAIL_MIN_ROWS");
    return e ? atoi(e) : 4096;
}

// does a call of `rows` estimator rows (2 x calls x frames) take the fused-tail kernels?  (part of the CUDA-graph key of a
// single call: the setting can change at run time)
bool flow_tail_path(long rows) { return cfm_tail_available() && !fuse_ln() && rows >= fused_tail_min_rows(); }

// the transformer blocks [j0, j0 + nb) of one estimator stage
static void tfm_stage(cbx_engine* e, Lane& L, int j0, int nb, int T, cudaStream_t st) {
    FlowModel& f = e->flow;
    const int M = 2 * L.nb * T;
    if (flow_tail_path(M)) {
        // stage opener: LayerNorm1 + QKV of the first block; then per block: attention, fused tail (+ next block's LN1 / QKV)
        CfmTailArgs a; a.M = M; a.h = L.c_h; a.qkv = L.c_qkv;
        bool ragged = false;
        for (int b = 0; b < L.nb; b++) ragged = ragged || L.call[b].Tt != L.Ttm;
        if (ragged && L.nb <= 16) { a.seq_T = T; for (int b = 0; b < L.nb; b++) a.seq_len[b] = 2 * L.call[b].Tt; }
        a.mode = CFM_TAIL_QKV; a.ln1_g = f.tfms[j0].n1.g; a.ln1_b = f.tfms[j0].n1.b;
        launch_cfm_tail(a, nullptr, nullptr, &f.tfms[j0].tw, st);
        for (int j = 0; j < nb; j++) {
            const TfmP& t = f.tfms[j0 + j];
            AttnParams at; at.q = L.c_qkv; at.k = L.c_qkv + C_INNER; at.v = L.c_qkv + 2 * C_INNER; at.ldq = at.ldk = at.ldv = 3 * C_INNER;
            at.q_bs = at.k_bs = at.v_bs = (long)T * 3 * C_INNER; at.o = L.c_o; at.ldo = C_INNER; at.o_bs = (long)T * C_INNER; at.T = T; at.H = 8; at.batch = 2 * L.nb; at.scale = 0.125f;
            set_kv_len(at, L, 2, 2);
            launch_attention(at, st);
            const bool more = j + 1 < nb;
            CfmTailArgs b; b.M = M; b.h = L.c_h; b.qkv = L.c_qkv; b.seq_T = a.seq_T; for (int i = 0; i < 16; i++) b.seq_len[i] = a.seq_len[i];
            b.mode = CFM_TAIL_OUT | CFM_TAIL_FF | (more ? CFM_TAIL_QKV : 0);
            b.b_out = t.out.b; b.ln3_g = t.n3.g; b.ln3_b = t.n3.b; b.b0 = t.ff0.b; b.b2 = t.ff2.b;
            if (more) { b.ln1_g = f.tfms[j0 + j + 1].n1.g; b.ln1_b = f.tfms[j0 + j + 1].n1.b; }
            launch_cfm_tail(b, L.c_o, &t.tw, more ? &f.tfms[j0 + j + 1].tw : nullptr, st);
        }
        e->gpu_launches += 1 + 2 * nb;
        return;
    }
    for (int j = 0; j < nb; j++)
        tfm_block(e, L, f.tfms[j0 + j], T, fuse_ln() && j > 0, (fuse_ln() && j + 1 < nb) ? &f.tfms[j0 + j + 1].n1 : nullptr, st);
}

// slab GEMMs of the estimator (batch = 2 x calls slabs of T frames): row tiles in a slab's padding are skipped
static void set_row_len(GemmParams& g, const Lane& L) {
    bool ragged = false;
    for (int b = 0; b < L.nb; b++) ragged = ragged || L.call[b].Tt != L.Ttm;
    if (!ragged || L.nb > 16) return;
    g.row_div = 2;
    for (int b = 0; b < L.nb; b++) g.row_len[b] = 2 * L.call[b].Tt;
}

// resnet: in = haloed bf16 [2nb][CH+T][cin] -> L.c_h fp32 [2nb][T][256]
static void resnet(cbx_engine* e, Lane& L, const ResnetP& r, const float* tproj, const bf16* in, int T, cudaStream_t st) {
    const int NB = 2 * L.nb;
    const long TH = T + CH, bs = (long)T * C_CH;
    GemmParams g = mk(r.c1, in, r.cin, T, r.cin, r.cin); g.batch = NB; g.a_bs = TH * r.cin; g.outF = L.c_tmp; g.ldc = C_CH; g.c_bs = bs; set_row_len(g, L); launch_gemm(g, st);
    NormParams n; n.in = L.c_tmp; n.ld_in = C_CH; n.in_bs = bs; n.rows = T; n.batch = NB; n.C = C_CH; n.gain = r.n1.g; n.bias = r.n1.b; n.eps = 1e-5f; n.act = ACT_MISH;
    n.add = tproj; n.add_bs = 0; n.outB = L.c_hb + CH * C_CH; n.ld_outB = C_CH; n.outB_bs = TH * C_CH;
    launch_norm(n, st);
    g = mk(r.c2, L.c_hb, C_CH, T, C_CH, C_CH); g.batch = NB; g.a_bs = TH * C_CH; g.outF = L.c_tmp; g.ldc = C_CH; g.c_bs = bs; set_row_len(g, L); launch_gemm(g, st);
    NormParams n2; n2.in = L.c_tmp; n2.ld_in = C_CH; n2.in_bs = bs; n2.rows = T; n2.batch = NB; n2.C = C_CH; n2.gain = r.n2.g; n2.bias = r.n2.b; n2.eps = 1e-5f; n2.act = ACT_MISH;
    n2.outF = L.c_tmp2; n2.ld_outF = C_CH; n2.outF_bs = bs;
    launch_norm(n2, st);
    g = mk(r.res, in + CH * r.cin, r.cin, T, r.cin, 0); g.batch = NB; g.a_bs = TH * r.cin; g.res = L.c_tmp2; g.ldr = C_CH; g.r_bs = bs; g.outF = L.c_h; g.ldc = C_CH; g.c_bs = bs; set_row_len(g, L);
    launch_gemm(g, st);
    e->gpu_launches += 5;
}

static void to_bf16_haloed(cbx_engine* e, Lane& L, bf16* dst, int ld, int col0, int T, cudaStream_t st) {
    // L.c_h fp32 [2nb][T][256] -> dst [2nb][CH+T][ld] columns [col0, col0+256)
    launch_f32_to_bf16_slabs(L.c_h, C_CH, (long)T * C_CH, dst + (long)CH * ld + col0, ld, (long)(T + CH) * ld, T, C_CH, 2 * L.nb, st);
    e->gpu_launches += 1;
}

static void estimator(cbx_engine* e, Lane& L, int T, int step, cudaStream_t st) {
    FlowModel& f = e->flow;
    const int NB = 2 * L.nb;
    const int nb = e->cfg.cfm_blocks, nmid = e->cfg.cfm_mid, nsteps = e->cfg.cfm_steps, nres = nmid + 2;
    const long TH = T + CH, bs = (long)T * C_CH;
    auto tp = [&](int r) { return f.tproj + ((long)r * nsteps + step) * C_CH; };
    // down stage
    resnet(e, L, f.resnets[0], tp(0), L.c_in, T, st);
    tfm_stage(e, L, 0, nb, T, st);
    to_bf16_haloed(e, L, L.c_upin, 2 * C_CH, C_CH, T, st);          // skip connection -> channels [256,512) of the up-stage input
    to_bf16_haloed(e, L, L.c_hb, C_CH, 0, T, st);
    GemmParams g = mk(f.down_conv, L.c_hb, C_CH, T, C_CH, C_CH); g.batch = NB; g.a_bs = TH * C_CH; g.outB = L.c_inb + CH * C_CH; g.ldc = C_CH; g.c_bs = TH * C_CH;
    launch_gemm(g, st);
    for (int i = 0; i < nmid; i++) {
        resnet(e, L, f.resnets[1 + i], tp(1 + i), L.c_inb, T, st);
        tfm_stage(e, L, (1 + i) * nb, nb, T, st);
        if (i + 1 < nmid) to_bf16_haloed(e, L, L.c_inb, C_CH, 0, T, st);
        else to_bf16_haloed(e, L, L.c_upin, 2 * C_CH, 0, T, st);
    }
    resnet(e, L, f.resnets[nres - 1], tp(nres - 1), L.c_upin, T, st);
    tfm_stage(e, L, (nres - 1) * nb, nb, T, st);
    to_bf16_haloed(e, L, L.c_hb, C_CH, 0, T, st);
    g = mk(f.up_conv, L.c_hb, C_CH, T, C_CH, C_CH); g.batch = NB; g.a_bs = TH * C_CH; g.outB = L.c_inb + CH * C_CH; g.ldc = C_CH; g.c_bs = TH * C_CH;
    launch_gemm(g, st);
    g = mk(f.final_conv, L.c_inb, C_CH, T, C_CH, C_CH); g.batch = NB; g.a_bs = TH * C_CH; g.outF = L.c_tmp; g.ldc = C_CH; g.c_bs = bs; launch_gemm(g, st);
    NormParams n; n.in = L.c_tmp; n.ld_in = C_CH; n.in_bs = bs; n.rows = T; n.batch = NB; n.C = C_CH; n.gain = f.final_ln.g; n.bias = f.final_ln.b; n.eps = 1e-5f; n.act = ACT_MISH;
    n.outB = L.c_fb; n.ld_outB = C_CH; n.outB_bs = bs;
    launch_norm(n, st);
    g = mk(f.final_proj, L.c_fb, C_CH, NB * T, C_CH, 0); g.outF = L.v; g.ldc = MEL; launch_gemm(g, st);
    e->gpu_launches += 5;
}

// host -> device staging of the batch's tokens (not graph-captured): L.call[0..nb) must be set
void flow_stage(cbx_engine* e, Lane& L, const int* const* tokens_h, cudaStream_t st) {
    const cbx_config& c = e->cfg;
    CBX_REQUIRE(L.nb >= 1 && L.nb <= L.bmax, "s3gen: batch exceeds the lane's capacity");
    L.Ttm = 0;
    for (int b = 0; b < L.nb; b++) {
        FlowCall& k = L.call[b];
        CBX_REQUIRE(k.n >= 1 && k.n <= c.max_s3_tokens, "s3gen: token count out of range");
        k.Tt = k.v->n_prompt + k.n;
        CBX_REQUIRE(k.v->n_feat <= 2 * k.Tt, "s3gen: prompt_feat longer than the encoded sequence");
        CBX_REQUIRE(2 * k.Tt <= NOISE_LEN, "s3gen: sequence exceeds the CFM noise buffer");
        for (int i = 0; i < k.n; i++) CBX_REQUIRE(tokens_h[b][i] >= 0 && tokens_h[b][i] < F_V, "s3gen: token id out of range");
        L.Ttm = std::max(L.Ttm, k.Tt);
    }
    if (L.nb > 1) CBX_CHECK(cudaMemsetAsync(L.tok, 0, (size_t)L.nb * L.Ttm * 4, st));   // pad ids (their rows are never consumed)
    for (int b = 0; b < L.nb; b++) {
        const FlowCall& k = L.call[b];
        CBX_CHECK(cudaMemcpyAsync(L.tok + (long)b * L.Ttm, k.v->prompt_token, (size_t)k.v->n_prompt * 4, cudaMemcpyDeviceToDevice, st));
        CBX_CHECK(cudaMemcpyAsync(L.tok + (long)b * L.Ttm + k.v->n_prompt, tokens_h[b], (size_t)k.n * 4, cudaMemcpyHostToDevice, st));
    }
    CBX_CHECK(cudaStreamSynchronize(st));   // tokens_h may be transient host buffers
}

// device-only part (CUDA-graph capturable for a fixed batch shape): L.tok -> L.melb[b] = [2 n_b][80] at stride 2*max_s3_tokens*80
void flow_run(cbx_engine* e, Lane& L, cudaStream_t st) {
    FlowModel& f = e->flow;
    const cbx_config& c = e->cfg;
    const int B = L.nb, T = 2 * L.Ttm;
    const long xs = (long)T * MEL;
    encoder(e, L, st);
    CBX_CHECK(cudaMemsetAsync(L.cond, 0, (size_t)B * xs * 4, st));
    for (int b = 0; b < B; b++) {
        const Voice& v = *L.call[b].v;
        CBX_CHECK(cudaMemcpyAsync(L.cond + b * xs, v.prompt_feat, (size_t)v.n_feat * MEL * 4, cudaMemcpyDeviceToDevice, st));
        CBX_CHECK(cudaMemcpyAsync(L.x + b * xs, f.noise, (size_t)xs * 4, cudaMemcpyDeviceToDevice, st));
    }
    // causal halos of the estimator inputs
    const long TH = T + CH;
    for (int b = 0; b < 2 * B; b++) {
        CBX_CHECK(cudaMemsetAsync(L.c_in + b * TH * C_IN, 0, (size_t)CH * C_IN * 2, st));
        CBX_CHECK(cudaMemsetAsync(L.c_hb + b * TH * C_CH, 0, (size_t)CH * C_CH * 2, st));
        CBX_CHECK(cudaMemsetAsync(L.c_inb + b * TH * C_CH, 0, (size_t)CH * C_CH * 2, st));
        CBX_CHECK(cudaMemsetAsync(L.c_upin + b * TH * 2 * C_CH, 0, (size_t)CH * 2 * C_CH * 2, st));
    }
    const int ns = c.cfm_steps;
    for (int k = 0; k < ns; k++) {
        for (int b = 0; b < B; b++)
            launch_pack_cfm_input(L.x + b * xs, L.mu + b * xs, L.call[b].v->spks, L.cond + b * xs, L.c_in + ((long)2 * b * TH + CH) * C_IN, TH * C_IN, T, MEL, st);
        estimator(e, L, T, k, st);
        for (int b = 0; b < B; b++) launch_euler_update(L.x + b * xs, L.v + (long)2 * b * xs, xs, xs, f.t_span[ns + k], c.cfm_cfg_rate, st);
        e->gpu_launches += 2 * B;
    }
    const long mel_bs = 2L * c.max_s3_tokens * MEL;
    for (int b = 0; b < B; b++) {
        const FlowCall& k = L.call[b];
        CBX_CHECK(cudaMemcpyAsync(L.melb + b * mel_bs, L.x + b * xs + (long)k.v->n_feat * MEL, (size_t)(2 * k.Tt - k.v->n_feat) * MEL * 4, cudaMemcpyDeviceToDevice, st));
    }
}

void flow_infer(cbx_engine* e, Lane& L, const Voice& v, const int* tokens_h, int n, cudaStream_t st) {
    L.nb = 1; L.call[0].v = &v; L.call[0].n = n;
    const int* toks[1] = {tokens_h};
    flow_stage(e, L, toks, st);
    flow_run(e, L, st);
    CBX_CHECK(cudaMemcpyAsync(L.mel, L.melb, (size_t)2 * n * MEL * 4, cudaMemcpyDeviceToDevice, st));
}
