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

// Decode window.  The reference's "full" overlap strategy re-synthesises every accumulated token of a text chunk for each slice and
// then emits only wav[previous_length:] (src/tts_streaming.py:694-699).  The vocoder is LOCAL: apart from the sine source (whose
// phase accumulates from the first frame -- F0 predictor, source and its STFT are therefore always computed in full, they are
// cheap), an output sample depends on mel frames within the receptive field of conv_pre, the three upsample stages and their
// resblocks (kernel 11, dilations 1 / 3 / 5: 60 samples at the stage's rate) -- under 20 mel frames in total.  With `w0` > 0 the
// convolution stack runs on mel frames [w0, Tg) only; rows left of the window are read as context (real data for conv_pre, stale
// but finite data deeper in the stack), so samples from (w0 + HIFT_WINDOW_MARGIN) * 480 on are IDENTICAL to the full decode
// (tests/test_gpu_parity.py::test_hift_window_is_exact); earlier samples of wav_out are left untouched.
void hift_infer(cbx_engine* e, Lane& L, int Tg, const float* cache_src_dev, long m, float* wav_out, float* src_out,
                const float* phase_h, const float* noise_dev, unsigned long long seed, cudaStream_t st, const SourceDyn* dyn, int w0) {
    HiftModel& h = e->hift;
    CBX_REQUIRE(Tg >= 1 && Tg <= 2 * e->cfg.max_s3_tokens, "hift: mel length out of range");
    CBX_REQUIRE(w0 >= 0 && w0 < Tg, "hift: decode window starts beyond the mel");
    const long Ls = (long)Tg * H_UP, F = 120L * Tg + 1;
    CBX_REQUIRE(m >= 0, "hift: negative cache_source length");
    if (m > Ls) m = Ls;
    const long tlen[4] = {Tg, 8L * Tg, 40L * Tg, F};
    const long off[4] = {w0, 8L * w0, 40L * w0, 120L * w0};        // first row of the window at every stage's rate
    auto zero_tail = [&](bf16* buf, long T, int C) { CBX_CHECK(cudaMemsetAsync(buf + (H_HALO + T) * C, 0, (size_t)H_HALO * C * 2, st)); };
    zero_tail(L.h_stft, F, H_NSRC_PAD);
    for (int i = 0; i < 4; i++) zero_tail(L.h_xb[i], tlen[i], H_BASE >> i);
    for (int i = 0; i < 3; i++) { zero_tail(L.h_a[i], tlen[i + 1], H_BASE >> (i + 1)); zero_tail(L.h_b[i], tlen[i + 1], H_BASE >> (i + 1)); }
    hift_f0(e, L, Tg, st);
    hift_source(e, L, L.h_f0, Tg, cache_src_dev, m, src_out, phase_h, noise_dev, seed, st, dyn);
    launch_stft16(src_out, Ls, L.h_stft + (long)H_HALO * H_NSRC_PAD, H_NSRC_PAD, (int)F, st);
    // ---- conv_pre (+ the leaky_relu that precedes ups[0])
    {
        GemmParams g = conv(h.conv_pre, L.h_mel + off[0] * MEL_PAD, MEL_PAD, 7, 1, (int)(Tg - off[0]));
        g.act = ACT_LRELU; g.act_param = 0.1f; g.outB = L.h_xb[0] + (H_HALO + off[0]) * H_BASE; g.ldc = H_BASE;
        launch_gemm(g, st);
    }
    e->gpu_launches += 2;
    for (int i = 0; i < 3; i++) {
        const int cin_i = H_BASE >> i, ch = H_BASE >> (i + 1), u = UPS_U[i], k = UPS_K[i], p = (k - u) / 2, taps = (k + u - 1) / u;
        const long Tin = tlen[i] - off[i], Tout = tlen[i + 1] - off[i + 1], oi = off[i], oo = off[i + 1];
        const int shift = (i == 2) ? 1 : 0;   // ReflectionPad1d((1,0)) after the last upsample
        // transposed conv as one GEMM over all u phases; row q covers outputs q*u - p + [0,u)
        GemmParams g;
        g.A = L.h_xb[i] + (H_HALO - (taps - 1) + oi) * cin_i; g.lda = cin_i; g.kc = cin_i; g.tap_stride = cin_i;
        g.W = h.ups[i].w; g.ldw = h.ups[i].K; g.M = (int)Tin + 1; g.N = u * ch; g.K = taps * cin_i; g.bias = h.ups[i].b;
        g.outF = L.h_x[i] + (shift - p + oo) * ch; g.ldc = (long)u * ch;
        g.ct_u = u; g.ct_cout = ch; g.ct_pad = p; g.ct_len = (int)(Tin * u);
        launch_gemm(g, st);
        if (shift && w0 == 0) launch_copy_row(L.h_x[i], L.h_x[i] + 2L * ch, ch, st);
        // source fusion: strided conv of the source STFT, source resblock, x += si
        GemmParams s;
        s.A = L.h_stft + (H_HALO - SD_PAD[i] + oo * SD_STRIDE[i]) * H_NSRC_PAD; s.lda = (long)SD_STRIDE[i] * H_NSRC_PAD; s.kc = H_NSRC_PAD; s.tap_stride = H_NSRC_PAD;
        s.W = h.sdown[i].w; s.ldw = h.sdown[i].K; s.M = (int)Tout; s.N = ch; s.K = h.sdown[i].K; s.bias = h.sdown[i].b; s.outF = L.h_si[i] + oo * ch; s.ldc = ch;
        launch_gemm(s, st);
        bf16 *ha = L.h_a[i] + oo * ch, *hb = L.h_b[i] + oo * ch;      // haloed buffers: the "halo" left of the window is earlier rows
        float *x = L.h_x[i] + oo * ch, *si = L.h_si[i] + oo * ch, *acc = L.h_acc[i] + oo * ch, *r = L.h_r[i] + oo * ch;
        resblock(e, h.sres[i], si, si, ha, hb, (int)Tout, RbOut{x, 1, 1.f, nullptr, 0, 0.f}, st);
        for (int kk = 0; kk < 3; kk++) {
            RbOut fo{acc, kk > 0, 1.f / 3.f, nullptr, 0, 0.f};
            if (kk == 2) { fo.outB2 = L.h_xb[i + 1] + (H_HALO + oo) * ch; fo.act2 = ACT_LRELU; fo.act2_param = (i == 2) ? 0.01f : 0.1f; }
            resblock(e, h.res[i * 3 + kk], x, r, ha, hb, (int)Tout, fo, st);
        }
        e->gpu_launches += 3;
    }
    {
        GemmParams g = conv(h.conv_post, L.h_xb[3] + off[3] * (H_BASE >> 3), H_BASE >> 3, 7, 1, (int)(F - off[3]));
        g.outF = L.h_post + off[3] * H_NSRC; g.ldc = H_NSRC;
        launch_gemm(g, st);
    }
    // frames left of the window hold stale rows: start far enough inside it that every frame a sample reads has been computed
    launch_istft16(L.h_post, H_NSRC, (int)F, wav_out, Ls, 0.99f, h.fade, 960, w0 > 0 ? 4 * (off[3] + 4) : 0, st);
    e->gpu_launches += 2;
}
