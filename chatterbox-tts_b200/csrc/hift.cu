// HiFT vocoder orchestration: F0 predictor -> harmonic/noise source -> STFT -> conv_pre ->
// 3 x (transposed-conv upsample, source fusion, 3 Snake resblocks averaged) -> conv_post -> iSTFT.
// Every conv / transposed conv is an implicit GEMM over zero-haloed, channels-last bf16 buffers.
#include "engine.h"

using namespace dims;

static const int UPS_U[3] = {8, 5, 3}, UPS_K[3] = {16, 11, 7}, RES_K[3] = {3, 7, 11}, SRES_K[3] = {7, 7, 11}, DIL[3] = {1, 3, 5};
static const int SD_STRIDE[3] = {15, 3, 1}, SD_K[3] = {30, 6, 1}, SD_PAD[3] = {7, 1, 0};

static Lin reg_lin(cbx_engine* e, const std::string& n, int N, int K) {
    Lin l; l.N = N; l.K = K;
    l.w = e->reg<bf16>(n + ".w", DT_BF16, (long)N * K);
    l.b = e->reg<float>(n + ".b", DT_F32, N);
    return l;
}
static ResBlockP reg_resblock(cbx_engine* e, const std::string& p, int ch, int k) {
    ResBlockP r; r.ch = ch; r.k = k;
    for (int j = 0; j < 3; j++) {
        std::string s = std::to_string(j);
        r.c1[j] = reg_lin(e, p + "c1_" + s, ch, k * ch);
        r.c2[j] = reg_lin(e, p + "c2_" + s, ch, k * ch);
        r.a1[j] = e->reg<float>(p + "a1_" + s, DT_F32, ch);
        r.a2[j] = e->reg<float>(p + "a2_" + s, DT_F32, ch);
    }
    return r;
}

void hift_build(cbx_engine* e) {
    HiftModel& h = e->hift;
    h.conv_pre = reg_lin(e, "hift.conv_pre", H_BASE, 7 * MEL_PAD);
    for (int i = 0; i < 3; i++) {
        int cin = H_BASE >> i, ch = H_BASE >> (i + 1), taps = (UPS_K[i] + UPS_U[i] - 1) / UPS_U[i];
        std::string s = std::to_string(i);
        h.ups[i] = reg_lin(e, "hift.ups" + s, UPS_U[i] * ch, taps * cin);
        h.sdown[i] = reg_lin(e, "hift.sdown" + s, ch, SD_K[i] * H_NSRC_PAD);
        h.sres[i] = reg_resblock(e, "hift.sres" + s + ".", ch, SRES_K[i]);
        for (int k = 0; k < 3; k++) h.res[i * 3 + k] = reg_resblock(e, "hift.res" + std::to_string(i * 3 + k) + ".", ch, RES_K[k]);
    }
    h.conv_post = reg_lin(e, "hift.conv_post", H_NSRC, 7 * (H_BASE >> 3));
    for (int l = 0; l < 5; l++) h.f0c[l] = reg_lin(e, "hift.f0c" + std::to_string(l), H_F0CH, 3 * (l == 0 ? MEL_PAD : H_F0CH));
    h.f0w = e->reg<float>("hift.f0w", DT_F32, H_F0CH);
    h.f0b = e->reg<float>("hift.f0b", DT_F32, 1);
    h.lw = e->reg<float>("hift.lw", DT_F32, H_NHARM);
    h.lb = e->reg<float>("hift.lb", DT_F32, 1);
    h.fade = e->reg<float>("hift.fade", DT_F32, 960);
    if (!e->dry) hift_init_constants();
}

// 'same' conv over a haloed buffer: in = buffer base (row 0 = first halo row), T valid rows starting at row H_HALO
static GemmParams conv(const Lin& l, const bf16* in, int cin, int k, int dil, int T) {
    GemmParams g;
    const int pad = dil * (k - 1) / 2;
    g.A = in + (long)(H_HALO - pad) * cin; g.lda = cin; g.kc = cin; g.tap_stride = (long)dil * cin;
    g.W = l.w; g.ldw = l.K; g.M = T; g.N = l.N; g.K = l.K; g.bias = l.b;
    return g;
}

struct RbOut { float* dst; int accumulate; float scale; bf16* outB2; int act2; float act2_param; };

static void resblock(cbx_engine* e, const ResBlockP& r, const float* x_in, float* x_work, bf16* ha, bf16* hb, int T, const RbOut& fin, cudaStream_t st) {
    const int ch = r.ch;
    bf16* a_valid = ha + (long)H_HALO * ch;
    bf16* b_valid = hb + (long)H_HALO * ch;
    launch_snake_rows(x_in, ch, a_valid, ch, T, ch, r.a1[0], st);
    for (int j = 0; j < 3; j++) {
        GemmParams g = conv(r.c1[j], ha, ch, r.k, DIL[j], T);
        g.act = ACT_SNAKE; g.act_alpha = r.a2[j]; g.outB = b_valid; g.ldc = ch;
        launch_gemm(g, st);
        g = conv(r.c2[j], hb, ch, r.k, 1, T);
        g.res = (j == 0) ? x_in : x_work; g.ldr = ch;
        if (j < 2) {
            g.outF = x_work; g.ldc = ch; g.outB2 = a_valid; g.ldc2 = ch; g.act2 = ACT_SNAKE; g.act2_alpha = r.a1[j + 1];
        } else {
            g.outF = fin.dst; g.ldc = ch; g.accumulate = fin.accumulate; g.out_scale = fin.scale;
            if (fin.outB2) { g.outB2 = fin.outB2; g.ldc2 = ch; g.act2 = fin.act2; g.act2_param = fin.act2_param; }
        }
        launch_gemm(g, st);
    }
    e->gpu_launches += 7;
}

// mel (L.mel fp32 [Tg][80]) -> L.h_f0 [Tg]; also stages the bf16 mel for conv_pre
void hift_f0(cbx_engine* e, Lane& L, int Tg, cudaStream_t st) {
    HiftModel& h = e->hift;
    auto zero_tail = [&](bf16* buf, long T, int C) { CBX_CHECK(cudaMemsetAsync(buf + (H_HALO + T) * C, 0, (size_t)H_HALO * C * 2, st)); };
    zero_tail(L.h_mel, Tg, MEL_PAD); zero_tail(L.h_f0a, Tg, H_F0CH); zero_tail(L.h_f0b, Tg, H_F0CH);
    launch_f32_to_bf16_rows(L.mel, MEL, L.h_mel + (long)H_HALO * MEL_PAD, MEL_PAD, Tg, MEL, ACT_NONE, 0.f, st);     // 80 of 128 columns; the rest stay zero
    // F0 predictor: 5 x (conv k3 + ELU) -> |linear|
    const bf16* fin = L.h_mel; int cin = MEL_PAD;
    bf16* pp[2] = {L.h_f0a, L.h_f0b};
    for (int l = 0; l < 5; l++) {
        GemmParams g = conv(h.f0c[l], fin, cin, 3, 1, Tg);
        g.act = ACT_ELU; g.outB = pp[l & 1] + (long)H_HALO * H_F0CH; g.ldc = H_F0CH;
        launch_gemm(g, st);
        fin = pp[l & 1]; cin = H_F0CH;
    }
    launch_f0_classifier(fin + (long)H_HALO * H_F0CH, H_F0CH, h.f0w, h.f0b, L.h_f0, Tg, H_F0CH, st);
    e->gpu_launches += 7;
}

// f0 [Tg] -> source [480*Tg] (SineGen + SourceModuleHnNSF), first m samples taken from cache_source
void hift_source(cbx_engine* e, Lane& L, const float* f0, int Tg, const float* cache_src_dev, long m, float* src_out,
                 const float* phase_h, const float* noise_dev, unsigned long long seed, cudaStream_t st, const SourceDyn* dyn) {
    HiftModel& h = e->hift;
    SourceParams sp; sp.f0 = f0; sp.cum = L.h_cum; sp.s = src_out; sp.L = (long)Tg * H_UP; sp.up = H_UP; sp.sr = 24000.f; sp.n_harm = H_NHARM;
    sp.lw = h.lw; sp.lb = h.lb; sp.noise = noise_dev; sp.cache = cache_src_dev; sp.cache_len = m; sp.seed = seed; sp.dyn = dyn;
    if (phase_h) {
        CBX_CHECK(cudaMemcpyAsync(L.h_phase, phase_h, H_NHARM * 4, cudaMemcpyHostToDevice, st));
        CBX_CHECK(cudaStreamSynchronize(st));
        sp.phase = L.h_phase;
    }
    launch_source(sp, Tg, st);
    e->gpu_launches += 2;
}

void hift_infer(cbx_engine* e, Lane& L, int Tg, const float* cache_src_dev, long m, float* wav_out, float* src_out,
                const float* phase_h, const float* noise_dev, unsigned long long seed, cudaStream_t st, const SourceDyn* dyn) {
    HiftModel& h = e->hift;
    CBX_REQUIRE(Tg >= 1 && Tg <= 2 * e->cfg.max_s3_tokens, "hift: mel length out of range");
    const long Ls = (long)Tg * H_UP, F = 120L * Tg + 1;
    CBX_REQUIRE(m >= 0, "hift: negative cache_source length");
    if (m > Ls) m = Ls;
    const long tlen[4] = {Tg, 8L * Tg, 40L * Tg, F};
    auto zero_tail = [&](bf16* buf, long T, int C) { CBX_CHECK(cudaMemsetAsync(buf + (H_HALO + T) * C, 0, (size_t)H_HALO * C * 2, st)); };
    zero_tail(L.h_stft, F, H_NSRC_PAD);
    for (int i = 0; i < 4; i++) zero_tail(L.h_xb[i], tlen[i], H_BASE >> i);
    for (int i = 0; i < 3; i++) { zero_tail(L.h_a[i], tlen[i + 1], H_BASE >> (i + 1)); zero_tail(L.h_b[i], tlen[i + 1], H_BASE >> (i + 1)); }
    hift_f0(e, L, Tg, st);
    hift_source(e, L, L.h_f0, Tg, cache_src_dev, m, src_out, phase_h, noise_dev, seed, st, dyn);
    launch_stft16(src_out, Ls, L.h_stft + (long)H_HALO * H_NSRC_PAD, H_NSRC_PAD, (int)F, st);
    // ---- conv_pre (+ the leaky_relu that precedes ups[0])
    {
        GemmParams g = conv(h.conv_pre, L.h_mel, MEL_PAD, 7, 1, Tg);
        g.act = ACT_LRELU; g.act_param = 0.1f; g.outB = L.h_xb[0] + (long)H_HALO * H_BASE; g.ldc = H_BASE;
        launch_gemm(g, st);
    }
    e->gpu_launches += 2;
    for (int i = 0; i < 3; i++) {
        const int cin_i = H_BASE >> i, ch = H_BASE >> (i + 1), u = UPS_U[i], k = UPS_K[i], p = (k - u) / 2, taps = (k + u - 1) / u;
        const long Tin = tlen[i], Tout = tlen[i + 1];
        const int shift = (i == 2) ? 1 : 0;   // ReflectionPad1d((1,0)) after the last upsample
        // transposed conv as one GEMM over all u phases; row q covers outputs q*u - p + [0,u)
        GemmParams g;
        g.A = L.h_xb[i] + (long)(H_HALO - (taps - 1)) * cin_i; g.lda = cin_i; g.kc = cin_i; g.tap_stride = cin_i;
        g.W = h.ups[i].w; g.ldw = h.ups[i].K; g.M = (int)Tin + 1; g.N = u * ch; g.K = taps * cin_i; g.bias = h.ups[i].b;
        g.outF = L.h_x[i] + (long)(shift - p) * ch; g.ldc = (long)u * ch;
        g.ct_u = u; g.ct_cout = ch; g.ct_pad = p; g.ct_len = (int)(Tin * u);
        launch_gemm(g, st);
        if (shift) launch_copy_row(L.h_x[i], L.h_x[i] + 2L * ch, ch, st);
        // source fusion: strided conv of the source STFT, source resblock, x += si
        GemmParams s;
        s.A = L.h_stft + (long)(H_HALO - SD_PAD[i]) * H_NSRC_PAD; s.lda = (long)SD_STRIDE[i] * H_NSRC_PAD; s.kc = H_NSRC_PAD; s.tap_stride = H_NSRC_PAD;
        s.W = h.sdown[i].w; s.ldw = h.sdown[i].K; s.M = (int)Tout; s.N = ch; s.K = h.sdown[i].K; s.bias = h.sdown[i].b; s.outF = L.h_si[i]; s.ldc = ch;
        launch_gemm(s, st);
        resblock(e, h.sres[i], L.h_si[i], L.h_si[i], L.h_a[i], L.h_b[i], (int)Tout, RbOut{L.h_x[i], 1, 1.f, nullptr, 0, 0.f}, st);
        for (int kk = 0; kk < 3; kk++) {
            RbOut fo{L.h_acc[i], kk > 0, 1.f / 3.f, nullptr, 0, 0.f};
            if (kk == 2) { fo.outB2 = L.h_xb[i + 1] + (long)H_HALO * ch; fo.act2 = ACT_LRELU; fo.act2_param = (i == 2) ? 0.01f : 0.1f; }
            resblock(e, h.res[i * 3 + kk], L.h_x[i], L.h_r[i], L.h_a[i], L.h_b[i], (int)Tout, fo, st);
        }
        e->gpu_launches += 3;
    }
    {
        GemmParams g = conv(h.conv_post, L.h_xb[3], H_BASE >> 3, 7, 1, (int)F);
        g.outF = L.h_post; g.ldc = H_NSRC;
        launch_gemm(g, st);
    }
    launch_istft16(L.h_post, H_NSRC, (int)F, wav_out, Ls, 0.99f, h.fade, 960, st);
    e->gpu_launches += 2;
}
