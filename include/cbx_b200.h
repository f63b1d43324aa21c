/* C-ABI of the B200-native Chatterbox hot path (libcbx_b200.so).
 *
 * The reference (akashdeep000/chatterbox-tts) has no FFI: its engine calls a Python object surface
 * (`chatterbox.*`) from src/tts_streaming.py.  Each entry point below states the reference call
 * site it replaces; chatterbox-tts_b200/chatterbox/ is the Python shim that binds them with ctypes
 * so the reference's worker.py / tts_streaming.py keep their interface.
 *
 * Conventions: plain pointers and sizes only.  `*_h` pointers are host memory, `*_d` pointers are
 * device memory on the engine's GPU.  `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy
 * default stream); work is ordered after everything already enqueued on it, and everything the call
 * enqueues is visible to later work on it.  Every function returns 0 on success and a non-zero code
 * on failure; cbx_last_error() then returns a thread-local message.  All functions are thread-safe
 * and do not touch the Python GIL.  There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef CBX_B200_H
#define CBX_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CBX_ABI_VERSION 2

typedef struct cbx_engine cbx_engine;

typedef struct cbx_config {
    int t3_layers;      /* 30 */
    int enc_blocks;     /* 6  conformer blocks before the x2 upsample */
    int up_blocks;      /* 4  conformer blocks after it */
    int cfm_blocks;     /* 4  transformer blocks per estimator stage */
    int cfm_mid;        /* 12 mid stages */
    int cfm_steps;      /* 10 Euler steps */
    float cfm_cfg_rate; /* 0.7 */
    int max_streams;    /* concurrent T3 streams (each = 2 CFG rows) */
    int max_seq;        /* KV positions per row */
    int max_text;       /* text tokens per T3 call incl. SOT/EOT */
    int max_s3_tokens;  /* speech tokens per s3gen call (excluding the voice prompt) */
    int max_prompt_tokens; /* voice prompt tokens (<= 250) */
    int n_voices;       /* voice-conditioning cache slots */
    int n_lanes;        /* concurrent s3gen calls */
} cbx_config;

int cbx_abi_version(void);
const char* cbx_last_error(void);

/* ChatterboxTTS.from_local(ckpt_dir, device)            -- src/tts_streaming.py:252-258 */
int cbx_engine_create(const cbx_config* cfg, int device, cbx_engine** out);
void cbx_engine_destroy(cbx_engine* e);
/* registry only, no device needed: lets a host-side packer be checked against cbx_tensor_info (tests) */
int cbx_manifest_create(const cbx_config* cfg, cbx_engine** out);
/* checkpoint upload: the engine publishes the packed tensors it wants (name, dtype 0=f32 1=bf16, numel) */
int cbx_tensor_count(cbx_engine* e);
int cbx_tensor_info(cbx_engine* e, int idx, char* name, int name_cap, int64_t* numel, int* dtype);
int cbx_tensor_upload(cbx_engine* e, const char* name, const void* data_h, int64_t nbytes);
int cbx_finalize(cbx_engine* e); /* fails if any tensor is missing; precomputes CFM time embeddings */

/* voice-conditioning cache: Conditionals(t3_cond, ref_dict) -- src/tts_streaming.py:106-118, :357-384, :386-406 */
int cbx_voice_put(cbx_engine* e, int voice, const float* speaker_emb_h /*256*/, const int32_t* cond_tokens_h, int n_cond,
                  float emotion_adv, const int32_t* prompt_token_h, int n_prompt, const float* prompt_feat_h /*n_feat*80*/,
                  int n_feat, const float* xvector_h /*192*/, void* stream);
/* clear_voice_cache(voice_id)                            -- src/tts_streaming.py:349-355 */
int cbx_voice_drop(cbx_engine* e, int voice);

/* T3.inference_stream(t3_cond, text_tokens, max_new_tokens, temperature, cfg_weight)
 *                                                        -- src/tts_streaming.py:287-292, :420-435
 * text_ids_h: one row of text ids already framed with SOT/EOT (the engine duplicates it for the CFG row,
 * :475-478).  Runs the prefill and returns a stream slot. */
int cbx_t3_open(cbx_engine* e, int voice, const int32_t* text_ids_h, int n_text, float cfg_weight, float temperature,
                float repetition_penalty, float min_p, float top_p, uint64_t seed, int max_new_tokens, int* slot_out,
                void* stream);
/* Several cbx_t3_open calls as ONE prefill pass (requests that arrive together: concurrent streams, the text chunks of one
 * request): sequences padded to the longest, causal attention with a key length per sequence.  Every stream's KV cache, first
 * logits and sampled ids equal what its own cbx_t3_open yields.  n <= 8. */
typedef struct cbx_t3_open_req {
    int voice; const int32_t* text_ids_h; int n_text;
    float cfg_weight, temperature, repetition_penalty, min_p, top_p;
    uint64_t seed; int max_new_tokens;
} cbx_t3_open_req;
int cbx_t3_open_batch(cbx_engine* e, const cbx_t3_open_req* reqs, int n, int* slots_out, void* stream);

/* next() on up to max_streams generators at once: n_steps decode steps for the given slots (batched rows).
 * noise_d: optional explicit Exp(1) sampling noise [n_steps][n_slots][8194] (parity tests); NULL = Philox(seed). */
int cbx_t3_step(cbx_engine* e, const int32_t* slots_h, int n_slots, int n_steps, const float* noise_d, void* stream);
/* selects the decode-step implementation: 0 = per-projection GEMV kernels (default), 1 = one persistent kernel per step
 * (lower latency for one or two streams, but it owns every SM while it runs; also enabled by CBX_T3_MEGA=1) */
int cbx_t3_set_persistent(cbx_engine* e, int on);
/* Moves the engine's T3 work to its high- (1) or low-priority (0) CUDA stream; work already queued stays ordered before what follows.
 * The decode step is a chain of ~150 short kernels that queue behind the wide S3Gen grids of the other streams: high priority buys
 * throughput (T3 runs ahead, S3Gen batches fill), low priority keeps a request's first S3Gen call fast.  The host scheduler switches
 * per decode batch (low while a first slice is pending).  No reference counterpart (torch streams there have one priority). */
int cbx_t3_set_priority(cbx_engine* e, int high);
/* Alignment-based EOS control of the generators opened AFTER this call (off by default).  Reference side: the model package's
 * AlignmentStreamAnalyzer, hooked on the attention of trunk layer `layer` (upstream: 9) inside the T3.inference_stream generator
 * that src/tts_streaming.py:420-435 primes -- the attention of the conditional row's newest query over the text span, averaged
 * over the heads, is tracked frame by frame; EOS is suppressed while the alignment is short of the last three text tokens and
 * forced after a long tail (>= 10 summed over a final column) or a repetition (> 5 summed row maxima over earlier columns).
 * The decision is applied to the CFG-mixed logits in the sampler kernel.  Uses the per-projection decode kernels. */
int cbx_t3_set_alignment_eos(cbx_engine* e, int on, int layer);
/* blocking read of a stream's analyzer: state_out[13] = {on, i0, S, frame_pos, text_pos, rows, started, started_at, complete,
 * completed_at, has_pre, ctl (bit 1 force, bit 0 suppress), cur_posn}; fstate_out[6] = {first4_max, prev_last2, tail3[3], rep_sum};
 * row_out / pre_out (optional, >= S floats): the newest alignment row and the prefilled BOS row (unmasked) */
int cbx_t3_alignment_peek(cbx_engine* e, int slot, int32_t* state_out, float* fstate_out, float* row_out, float* pre_out, void* stream);
/* test hook: overwrites a stream's analyzer state (same layout as the peek) -- lets a test place a stream just short of a forced EOS */
int cbx_t3_alignment_poke(cbx_engine* e, int slot, const int32_t* state, const float* fstate, void* stream);
/* parity-test entry: the analyzer alone over given alignment rows (device, [n_rows][S]; the first n_pre of them belong to the
 * first frame together with the row after them).  steps_out[frames][4] = {ctl, text_pos, started, complete} per frame. */
int cbx_op_alignment_run(const float* rows_d, int n_rows, int S, int n_pre, int32_t* steps_out_h, void* stream);
/* blocking reads of a stream's progress / tokens / last-step logits (2 x 8194: cond row, uncond row) */
int cbx_t3_poll(cbx_engine* e, int slot, int* n_generated, int* done, void* stream);
int cbx_t3_tokens(cbx_engine* e, int slot, int from, int count, int32_t* out_h, void* stream);
int cbx_t3_logits(cbx_engine* e, int slot, float* out_h, void* stream);
/* generator close()/GC: releases the KV pages             -- cancel path, src/tts_streaming.py:505-519 */
int cbx_t3_close(cbx_engine* e, int slot);
/* allocator state (tests / health checks): free KV pages and open stream slots */
int cbx_t3_stats(cbx_engine* e, int* free_pages, int* open_slots);
/* Non-blocking health probe: 0 while the engine's CUDA context is usable.  A device-side fault (a kernel trap on a stuck barrier,
 * an illegal address) is sticky for the whole process: every later call fails, and the only recovery is a new worker process.
 * The reference worker keeps serving after a failed request (src/worker.py:54-56 logs and goes on); the host engine uses this
 * probe to refuse new requests with a clear error instead of failing them one by one. */
int cbx_engine_health(cbx_engine* e);

/* S3Gen.inference(speech_tokens, ref_dict, cache_source) -> (wav, source)
 *                                                        -- src/tts_streaming.py:316-320, :583-590
 * tokens_h: n >= 3 ids < 6561.  cache_source_d: m >= 0 samples.  Outputs: 960*n samples each.
 * mel_out_d (optional, [2n][80]) exposes the CFM result; phase_h[9] / noise_d[9][960n] optionally replace the
 * SineGen randomness (parity tests); NULL = Philox(seed). */
int cbx_s3gen_infer(cbx_engine* e, int voice, const int32_t* tokens_h, int n, const float* cache_source_d, int64_t m,
                    float* wav_out_d, float* source_out_d, float* mel_out_d, const float* phase_h, const float* noise_d,
                    uint64_t seed, void* stream);
/* The same call for up to 16 requests at once (concurrent streams, or the text chunks of one request): the token -> mel
 * part runs as ONE batch padded to the longest sequence (exact: causal convolutions, per-frame norms / linears, attention
 * masked per sequence), the vocoder per call.  Results equal cbx_s3gen_infer call by call. */
typedef struct cbx_s3gen_call {
    int voice; const int32_t* tokens_h; int n;
    const float* cache_source_d; int64_t m;
    float* wav_out_d; float* source_out_d; float* mel_out_d /* optional */;
    uint64_t seed;
    /* The caller only reads wav_out_d[emit_from:] (the reference's "full" overlap emits wav[previous_length:], :694-699): the
     * vocoder's convolution stack then runs on a window that starts a receptive field before that sample; samples from emit_from
     * on equal the full decode exactly, earlier ones are left untouched.  source_out_d is always complete.  0 = everything. */
    int64_t emit_from;
} cbx_s3gen_call;
int cbx_s3gen_infer_batch(cbx_engine* e, const cbx_s3gen_call* calls, int n_calls, void* stream);
/* the two halves of the call above, for teacher-forced parity checks */
int cbx_flow_infer(cbx_engine* e, int voice, const int32_t* tokens_h, int n, float* mel_out_d, void* stream);
int cbx_hift_infer(cbx_engine* e, const float* mel_d /*[frames][80]*/, int frames, const float* cache_source_d, int64_t m,
                   float* wav_out_d, float* source_out_d, const float* phase_h, const float* noise_d, uint64_t seed, void* stream);

/* cbx_hift_infer with the convolution stack restricted to mel frames [w0, frames): samples from (w0 + 20) * 480 on equal the full decode */
int cbx_hift_infer_window(cbx_engine* e, const float* mel_d, int frames, const float* cache_source_d, int64_t m, float* wav_out_d,
                          float* source_out_d, uint64_t seed, int w0, void* stream);

/* HiFT stages for parity tests: mel -> f0 (ConvRNNF0Predictor) and f0 -> source (SineGen + SourceModuleHnNSF) */
int cbx_hift_f0(cbx_engine* e, const float* mel_d, int frames, float* f0_out_d, void* stream);
int cbx_hift_source(cbx_engine* e, const float* f0_d, int frames, const float* phase_h, const float* noise_d, uint64_t seed,
                    float* source_out_d, void* stream);

/* equal-power crossfade + clamp + int16 conversion        -- src/tts_streaming.py:710-746 and :149-155
 * out[i] = int16(clamp(x,-1,1) * 32767) with x = prev_tail[i]*fade_out[i] + cur[i]*fade_in[i] for i < fade_len (when
 * prev_tail_d), x = cur[i] otherwise, for i in [0, n_out).  fade_in_d / fade_out_d [fade_len]: the request's curves, built
 * by the caller exactly as the reference builds them (:867-871); products and sum round separately, as torch's do, so the
 * PCM is bit-exact. */
int cbx_crossfade_pcm(cbx_engine* e, const float* cur_d, int64_t n_out, const float* prev_tail_d, int fade_len,
                      const float* fade_in_d, const float* fade_out_d, int16_t* out_d, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Voice-conditioning encoders (reference src/tts_streaming.py:357-384: `s3gen.embed_ref` :366, `s3gen.tokenizer.forward`
 * :370-372, `ve.embeds_from_wavs` :374).  fp32 building blocks on device pointers (time-major, channels-last); the host
 * side (cbx_b200/conditioning.py) strings them together as the upstream modules do.  All take the caller's stream.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct cbx_sgemm_args {
    /* C[b][m][n] = epilogue(alpha * sum_kk A'(b,m,kk) * W(b,n,kk)), kk = tap * kc + c:
     *   A'(b,m,kk) = in(A[b*a_bs + t*lda + c]) with t = m*a_stride + tap*a_dil - a_pad (0 outside [0, a_rows));
     *   in(v) = relu?(v * a_scale[c] + a_shift[c]) when a_scale is given (BatchNorm + ReLU in front of a conv);
     *   W(b,n,kk) = W[b*w_bs + n*ldw + kk], or W[b*w_bs + kk*ldw + n] when w_trans;
     *   epilogue: + bias[n]; * o_scale[n] + o_shift[n]; act (0 none, 1 relu, 2 erf-gelu, 3 sigmoid, 4 tanh);
     *             * mul[b*mul_bs + (m / mul_div)*ldm + n]; + res[b*res_bs + m*ldr + n] (+ res2, same strides). */
    const float* A; int64_t lda; int64_t a_bs; int kc; int a_stride; int a_dil; int a_pad; int64_t a_rows;
    const float* a_scale; const float* a_shift; int a_relu;
    const float* W; int64_t ldw; int64_t w_bs; int w_trans;
    int M, N, K, batch;
    float alpha;
    const float* bias; const float* o_scale; const float* o_shift; int act;
    const float* mul; int64_t ldm; int64_t mul_bs; int mul_div;
    const float* res; const float* res2; int64_t ldr; int64_t res_bs;
    float* C; int64_t ldc; int64_t c_bs;
} cbx_sgemm_args;
int cbx_cond_sgemm(const cbx_sgemm_args* a, void* stream);
/* framing + window + |DFT| by direct summation: frame f = samples f*hop - pad .. (+ frame_len), reflected at the clip edges;
 * remove_dc / preemph: Kaldi's per-frame mean removal and pre-emphasis; mode 0 magnitude, 1 power, 2 sqrt(power + 1e-9);
 * out[f*ld_out + bin], bins = n_fft/2 + 1 */
int cbx_cond_frames_dft(const float* wav, int64_t n, int n_fft, int hop, int pad, int frame_len, const float* window, int remove_dc, float preemph,
                        int mode, int n_frames, float* out, int ld_out, void* stream);
/* polyphase windowed-sinc resampler (torchaudio.functional.resample): kern [up][klen] */
int cbx_cond_resample(const float* x, int64_t n_in, float* y, int64_t n_out, int down, int up, const float* kern, int klen, int width, void* stream);
int cbx_cond_layernorm(const float* x, int64_t ld_in, float* y, int64_t ld_out, int rows, int C, const float* g, const float* b, float eps, void* stream);
int cbx_cond_softmax(float* x, int64_t ld, int64_t bs, int rows, int cols, int batch, void* stream);
/* rotate-half rotary embedding (theta 10000, angles repeated over both halves) in place on [T][H][hd], then * scale */
int cbx_cond_rotary(float* x, int64_t ld, int T, int H, int hd, float scale, void* stream);
/* y = x + depthwise_conv_k(x) along time (zero padding (k-1)/2): the FSMN memory block of S3Tokenizer-v2 */
int cbx_cond_dwconv_add(const float* x, int64_t ld, const float* w, int k, float* y, int64_t ld_out, int T, int C, void* stream);
/* FSQ codebook of S3Tokenizer-v2: h [T][8] -> tanh * 0.999 -> round -> + 1 -> base-3 digits -> ids in [0, 6561) */
int cbx_cond_fsq(const float* h, int* out, int T, void* stream);
/* mode 0: x = log(max(x, floor_v)); mode 1: whisper log-mel (log10, clip to global max - 8, (x + 4) / 4; scratch1 = 1 float) */
int cbx_cond_mel_log(float* x, int64_t n, int mode, float floor_v, float* scratch1, void* stream);
/* per-column statistics over time of [T][C]: mode 0 subtract the mean in place; 1: out = [mean | unbiased std] of
 * relu(x*a_scale + a_shift) (statistics pooling); 2: out = mean, seg_out[i][c] = mean of segment i (seg frames) + mean */
int cbx_cond_col_stats(float* x, int64_t ld, int T, int C, int mode, const float* a_scale, const float* a_shift, float* out, int seg, float* seg_out, void* stream);
int cbx_cond_l2norm_rows(float* x, int rows, int C, int relu, void* stream);
int cbx_cond_mean_rows(const float* x, int rows, int C, float* out, void* stream);
/* 3x3 (pad 1) / 1x1 conv2d over [Cin][F][T], stride (stride_f, 1), folded BatchNorm, optional residual and ReLU; t_major
 * writes [T][Cout * F'] (CAMPPlus FCM head -> TDNN) */
int cbx_cond_conv2d(const float* x, const float* w, const float* scale, const float* shift, const float* res, float* y, int Cin, int Cout, int F, int T,
                    int ks, int stride_f, int relu, int t_major, void* stream);
/* one LSTM layer over B sequences: xp [B][T][4H] = x W_ih^T + b_ih + b_hh, w_hh_t [H][4H]; h_seq [B][T][H] and / or h_last [B][H] */
int cbx_cond_lstm_layer(const float* xp, const float* w_hh_t, float* h_seq, float* h_last, int B, int T, int H, void* stream);

/* counters for bench.py: kernels launched by this library since engine creation */
int64_t cbx_gpu_launches(cbx_engine* e);
/* GEMM launches that took the tcgen05/TMA path (process-wide; 0 means the mma.sync fallback served everything) */
long long cbx_gemm_tc_launches(void);
/* GEMM launches that fell back to the mma.sync kernel (shapes TMA cannot describe); 0 on the T3 / S3Gen paths */
long long cbx_gemm_mma_launches(void);
/* attention launches served by the tcgen05 flash-attention kernels (process-wide; both generations) */
long long cbx_attn_tc_launches(void);
/* of those, the launches of the second-generation kernel (P and O in tensor memory, attention_fa.cu) */
long long cbx_attn_fa_launches(void);
/* debug (CBX_ATTN_FA_DBG=4): %globaltimer stamps of the first CTA of the last launch (MMA issuer and softmax row 0 per key block) */
int cbx_attn_fa_trace(unsigned long long* out_h);
/* T3 decode projections launched on the tcgen05 path (batched rows, t3_gemv_tc.cu; process-wide) */
long long cbx_t3_tc_launches(void);
/* debug: %globaltimer stamps (ns) of the last tcgen05 GEMM's CTA 0: start, setup done, first TMA landed, MMAs issued,
 * accumulator ready, epilogue done, teardown */
int cbx_gemm_tc_trace(unsigned long long* out_h);
/* debug: %globaltimer stamps (ns) of CTA 0 through the five phases of layer 1 of the last T3 megakernel step */
int cbx_t3_mega_trace(unsigned long long* out_h);
/* debug: per CTA (256 x 4 int64) SM cycles thread 0 spent waiting for weight slots, waiting for arrival counters, and in total
 * during the last T3 megakernel step */
int cbx_t3_mega_prof(long long* out_h);

/* per-launch profiler for bench.py's roofline pass: between begin and end every kernel launch is bracketed by
 * CUDA events on its stream (T3 steps run un-graphed); end() returns, per kernel class (0 gemm, 1 attention,
 * 2 t3 gemv, 3 t3 decode attention, 4 sampler, 5 norm, 6 elementwise, 7 hift misc), the launch count, summed
 * device milliseconds and summed algorithmic work (FLOPs for classes 0-1, bytes otherwise). */
int cbx_profile_begin(void);
int cbx_profile_end(int64_t* counts, double* ms, double* work, int n_classes);

/* single-op hooks (kernel unit tests): C[M][N] = A[M][K] * W[N][K]^T (+bias), bf16 in, fp32 out; attention over
 * fused q|k|v rows; all pointers device. */
int cbx_op_gemm(const void* a_bf16_d, const void* w_bf16_d, const float* bias_d, float* out_d, int M, int N, int K, void* stream);
/* the same with the fused epilogue the model uses: act (0 none, 1 GELU, 2 SiLU, 3 Mish, 4 LeakyReLU, 5 ELU), optional fp32
 * residual [M][N] added after the activation, fp32 and / or bf16 output */
int cbx_op_gemm_ex(const void* a_bf16_d, const void* w_bf16_d, const float* bias_d, const float* res_d, float* out_f32_d,
                   void* out_bf16_d, int M, int N, int K, int act, void* stream);
/* fused tail of a CFM transformer block (kernel unit tests): mode bit 1 = out projection + residual (attn_o [M][512] bf16,
 * wout [256][512]), bit 2 = LayerNorm3 + GELU feed-forward + residual (w0 [1024][256], w2 [256][1024]), bit 4 = LayerNorm1 +
 * QKV projection of the next block (wqkv [1536][256]) -> qkv_out [M][1536] bf16; h [M][256] fp32 is updated in place */
int cbx_op_cfm_tail(int mode, int M, const void* attn_o_bf16_d, float* h_d, const void* wout_d, const float* b_out_d, const float* ln3_g_d,
                    const float* ln3_b_d, const void* w0_d, const float* b0_d, const void* w2_d, const float* b2_d, const float* ln1_g_d,
                    const float* ln1_b_d, const void* wqkv_d, void* qkv_out_bf16_d, void* stream);
/* launches served by the fused CFM block-tail kernel (process-wide) */
long long cbx_cfm_tail_launches(void);
/* debug: %globaltimer stamps (ns) of CTA 0 of the last fused-tail launch (9 phase boundaries, see cfm_tail.cu) */
int cbx_cfm_tail_trace(unsigned long long* out_h);
int cbx_op_attention(const void* qkv_bf16_d, void* out_bf16_d, int T, int H, int batch, int causal, void* stream);
/* the same with the flow encoder's additive relative-position bias: bias fp32 [batch][H][T][2T], entry (i, j) read at column T - 1 - i + j */
int cbx_op_attention_bias(const void* qkv_bf16_d, const float* bias_d, void* out_bf16_d, int T, int H, int batch, void* stream);

#ifdef __cplusplus
}
#endif
#endif
