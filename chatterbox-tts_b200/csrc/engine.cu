// C-ABI of libcbx_b200.so (see include/cbx_b200.h): engine lifecycle, checkpoint upload,
// voice cache, T3 streams, S3Gen calls, PCM conversion.  No torch types, no CPU fallback.
#include <cstring>
#include <cstdlib>
#include "engine.h"

using namespace dims;

static thread_local std::string g_err;
#define CBX_API_BEGIN try {
#define CBX_API_END                                  \
    return 0;                                        \
    } catch (const std::exception& ex) {             \
        g_err = ex.what();                           \
        return 1;                                    \
    } catch (...) {                                  \
        g_err = "unknown error";                     \
        return 1;                                    \
    }

void cbx_set_error(const char* msg) { g_err = msg ? msg : ""; }     // other translation units of the C-ABI (cond.cu)

namespace {

struct StreamBridge {   // order engine-internal stream `in` after the caller's stream and back
    cudaStream_t user, in; cudaEvent_t ev_in, ev_out;
    StreamBridge(cudaStream_t u, cudaStream_t i, cudaEvent_t a, cudaEvent_t b) : user(u), in(i), ev_in(a), ev_out(b) {
        CBX_CHECK(cudaEventRecord(ev_in, user));
        CBX_CHECK(cudaStreamWaitEvent(in, ev_in, 0));
    }
    void finish() {
        CBX_CHECK(cudaEventRecord(ev_out, in));
        CBX_CHECK(cudaStreamWaitEvent(user, ev_out, 0));
    }
};

// fade_in / fade_out: the request's curves built the reference's way (torch.sin / torch.cos of linspace(0, 1, fade_len) * pi/2
// on the device, src/tts_streaming.py:867-871); the mix rounds like torch's three separate elementwise ops (no FMA
// contraction), so the int16 output is bit-exact with `(prev * fade_out) + (head * fade_in)` -> clamp -> * 32767 -> int16.
__global__ void crossfade_pcm_kernel(const float* __restrict__ cur, long n, const float* __restrict__ prev, int fade_len,
                                     const float* __restrict__ fade_in, const float* __restrict__ fade_out, short* out) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = cur[i];
    if (prev && i < fade_len) x = __fadd_rn(__fmul_rn(prev[i], fade_out[i]), __fmul_rn(x, fade_in[i]));
    x = __fmul_rn(fminf(fmaxf(x, -1.f), 1.f), 32767.f);
    out[i] = (short)x;   // truncation toward zero, as torch .to(int16)
}

// Lanes are sticky per calling thread: a request thread keeps re-using one lane (its captured graphs and workspaces),
// different request threads spread over the lanes.  Rotating a single request over all lanes doubled its latency.
Lane& pick_lane(cbx_engine* e) {
    static thread_local cbx_engine* t_engine = nullptr;
    static thread_local int t_lane = -1;
    if (t_engine != e || t_lane < 0 || t_lane >= (int)e->lanes.size()) {
        std::lock_guard<std::mutex> g(e->lane_pick_mu);
        t_lane = e->lane_rr % (int)e->lanes.size();
        e->lane_rr++;
        t_engine = e;
    }
    return *e->lanes[t_lane];
}

}  // namespace

extern "C" {

int cbx_abi_version(void) { return CBX_ABI_VERSION; }
const char* cbx_last_error(void) { return g_err.c_str(); }

int cbx_engine_create(const cbx_config* cfg, int device, cbx_engine** out) {
    CBX_API_BEGIN
    CBX_REQUIRE(cfg && out, "null argument");
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    CBX_REQUIRE(err == cudaSuccess && ndev > 0, "no CUDA device: libcbx_b200 has no CPU path");
    CBX_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    CBX_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CBX_CHECK(cudaGetDeviceProperties(&prop, device));
    CBX_REQUIRE(prop.major == 10, "libcbx_b200 is built for sm_100a only");
    CBX_REQUIRE(cfg->t3_layers >= 1 && cfg->cfm_steps >= 1 && cfg->cfm_mid >= 1 && cfg->cfm_blocks >= 1 && cfg->enc_blocks >= 1 && cfg->up_blocks >= 1, "bad depth config");
    CBX_REQUIRE(cfg->max_streams >= 1 && cfg->max_streams <= 64 && cfg->n_lanes >= 1 && cfg->n_voices >= 1, "bad capacity config");
    CBX_REQUIRE(cfg->max_prompt_tokens >= 1 && cfg->max_s3_tokens >= 3 && 2 * (cfg->max_prompt_tokens + cfg->max_s3_tokens) <= NOISE_LEN, "bad s3gen capacity");
    cbx_engine* e = new cbx_engine();
    e->cfg = *cfg; e->device = device;
    gemm_init(); attention_init();
    t3_build(e); flow_build(e); hift_build(e);
    t3_alloc(e);
    {   // T3's ~150 short dependent kernels per step queue behind the wide S3Gen launches of the other streams.  Two streams, lowest
        // and highest priority; the host moves T3 between them (cbx_t3_set_priority): high while no first slice of a request is
        // waiting for S3Gen (throughput), low while one is (first-chunk latency).  CBX_T3_PRIORITY=1 / 0 pins the initial choice.
        int lo = 0, hi = 0;
        CBX_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CBX_CHECK(cudaStreamCreateWithPriority(&e->t3_st_prio[0], cudaStreamNonBlocking, lo));
        CBX_CHECK(cudaStreamCreateWithPriority(&e->t3_st_prio[1], cudaStreamNonBlocking, hi));
        CBX_CHECK(cudaEventCreateWithFlags(&e->t3_ev_sw, cudaEventDisableTiming));
        const char* pr = getenv("CBX_T3_PRIORITY");
        e->t3_prio = (pr && pr[0] == '1') ? 1 : 0;
        e->t3_st = e->t3_st_prio[e->t3_prio];
    }
    CBX_CHECK(cudaEventCreateWithFlags(&e->t3_ev_in, cudaEventDisableTiming));
    CBX_CHECK(cudaEventCreateWithFlags(&e->t3_ev_out, cudaEventDisableTiming));
    e->voices.resize(cfg->n_voices);
    for (auto& v : e->voices) {
        v.prefix = e->scratch<float>((long)T3_COND * T3_D);
        v.prompt_token = e->scratch<int>(cfg->max_prompt_tokens);
        v.prompt_feat = e->scratch<float>(2L * cfg->max_prompt_tokens * MEL);
        v.spks = e->scratch<float>(MEL);
    }
    for (int i = 0; i < cfg->n_lanes; i++) { Lane* L = new Lane(); lane_alloc(e, *L, 1); e->lanes.push_back(L); }
    for (int i = 0; i < 2; i++) {
        Lane* L = new Lane(); lane_alloc(e, *L, FLOW_MAXB); e->batch_lanes.push_back(L);
        for (int k = 0; k < 4; k++) { Lane* S = new Lane(); lane_alloc(e, *S, 0); L->sub.push_back(S); }
        CBX_CHECK(cudaEventCreateWithFlags(&L->ev_flow, cudaEventDisableTiming));
        L->ev_call.resize(FLOW_MAXB);
        for (auto& ev : L->ev_call) CBX_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    CBX_CHECK(cudaDeviceSynchronize());
    *out = e;
    CBX_API_END
}

int cbx_manifest_create(const cbx_config* cfg, cbx_engine** out) {
    CBX_API_BEGIN
    CBX_REQUIRE(cfg && out, "null argument");
    cbx_engine* e = new cbx_engine();
    e->cfg = *cfg; e->dry = true;
    t3_build(e); flow_build(e); hift_build(e);
    *out = e;
    CBX_API_END
}

void cbx_engine_destroy(cbx_engine* e) {
    if (!e) return;
    if (e->dry) { delete e; return; }
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (auto& kv : e->t3.step_graphs) cudaGraphExecDestroy(kv.second);
    t3_mega_free(&e->t3.mega_state);
    for (auto& t : e->tensors) cudaFree(t.ptr);
    for (void* p : e->scratch_allocs) cudaFree(p);
    for (Lane* L : e->batch_lanes) {
        e->lanes.push_back(L);
        for (Lane* S : L->sub) e->lanes.push_back(S);
        if (L->ev_flow) cudaEventDestroy(L->ev_flow);
        for (auto& ev : L->ev_call) cudaEventDestroy(ev);
    }
    for (Lane* L : e->lanes) { for (auto& kv : L->graphs) cudaGraphExecDestroy(kv.second); cudaFreeHost(L->g_dyn_h); cudaStreamDestroy(L->st); cudaEventDestroy(L->ev_in); cudaEventDestroy(L->ev_out); delete L; }
    cudaStreamDestroy(e->t3_st_prio[0]); cudaStreamDestroy(e->t3_st_prio[1]); cudaEventDestroy(e->t3_ev_sw); cudaEventDestroy(e->t3_ev_in); cudaEventDestroy(e->t3_ev_out);
    if (e->t3.align_st) { cudaStreamDestroy(e->t3.align_st); cudaEventDestroy(e->t3.align_fork); cudaEventDestroy(e->t3.align_join); }
    delete e;
}

int cbx_tensor_count(cbx_engine* e) { return e ? (int)e->tensors.size() : -1; }

int cbx_tensor_info(cbx_engine* e, int idx, char* name, int name_cap, int64_t* numel, int* dtype) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && idx >= 0 && idx < (int)e->tensors.size(), "tensor index out of range");
    const TensorRec& t = e->tensors[idx];
    CBX_REQUIRE((int)t.name.size() + 1 <= name_cap, "name buffer too small");
    std::strcpy(name, t.name.c_str());
    *numel = t.numel; *dtype = t.dtype;
    CBX_API_END
}

int cbx_tensor_upload(cbx_engine* e, const char* name, const void* data_h, int64_t nbytes) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && name && data_h, "null argument");
    auto it = e->index.find(name);
    CBX_REQUIRE(it != e->index.end(), std::string("unknown tensor ") + name);
    TensorRec& t = e->tensors[it->second];
    const int64_t want = t.numel * (t.dtype == DT_F32 ? 4 : 2);
    CBX_REQUIRE(nbytes == want, std::string("size mismatch for ") + name + ": got " + std::to_string(nbytes) + " want " + std::to_string(want));
    CBX_CHECK(cudaSetDevice(e->device));
    CBX_CHECK(cudaMemcpy(t.ptr, data_h, nbytes, cudaMemcpyHostToDevice));
    t.loaded = true;
    CBX_API_END
}

int cbx_finalize(cbx_engine* e) {
    CBX_API_BEGIN
    CBX_REQUIRE(e, "null engine");
    for (auto& t : e->tensors) CBX_REQUIRE(t.loaded, "tensor not uploaded: " + t.name);
    CBX_CHECK(cudaSetDevice(e->device));
    flow_finalize(e, e->t3_st);
    e->finalized = true;
    CBX_API_END
}

int cbx_voice_put(cbx_engine* e, int voice, const float* speaker_emb_h, const int32_t* cond_tokens_h, int n_cond, float emotion_adv,
                  const int32_t* prompt_token_h, int n_prompt, const float* prompt_feat_h, int n_feat, const float* xvector_h, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized, "engine not finalized");
    CBX_REQUIRE(voice >= 0 && voice < e->cfg.n_voices, "voice slot out of range");
    CBX_REQUIRE(n_prompt >= 1 && n_prompt <= e->cfg.max_prompt_tokens && n_feat == 2 * n_prompt, "prompt_feat must hold 2 mel frames per prompt token");
    for (int i = 0; i < n_cond; i++) CBX_REQUIRE(cond_tokens_h[i] >= 0 && cond_tokens_h[i] < T3_V, "cond token out of range");
    for (int i = 0; i < n_prompt; i++) CBX_REQUIRE(prompt_token_h[i] >= 0 && prompt_token_h[i] < F_V, "prompt token out of range");
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->voice_mu);
    std::lock_guard<std::mutex> g2(e->t3_mu);
    Voice& v = e->voices[voice];
    v.valid = false;
    StreamBridge br((cudaStream_t)stream, e->t3_st, e->t3_ev_in, e->t3_ev_out);
    t3_voice_prefix(e, v, speaker_emb_h, cond_tokens_h, n_cond, emotion_adv, e->t3_st);
    float* xv; CBX_CHECK(cudaMalloc(&xv, F_SPK * 4));
    CBX_CHECK(cudaMemcpyAsync(xv, xvector_h, F_SPK * 4, cudaMemcpyHostToDevice, e->t3_st));
    CBX_CHECK(cudaMemcpyAsync(v.prompt_token, prompt_token_h, (size_t)n_prompt * 4, cudaMemcpyHostToDevice, e->t3_st));
    CBX_CHECK(cudaMemcpyAsync(v.prompt_feat, prompt_feat_h, (size_t)n_feat * MEL * 4, cudaMemcpyHostToDevice, e->t3_st));
    launch_spk_affine(xv, F_SPK, e->flow.spk_w, e->flow.spk_b, v.spks, MEL, e->t3_st);
    CBX_CHECK(cudaStreamSynchronize(e->t3_st));
    cudaFree(xv);
    v.n_prompt = n_prompt; v.n_feat = n_feat; v.version++; v.valid = true;
    br.finish();
    CBX_API_END
}

int cbx_voice_drop(cbx_engine* e, int voice) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && voice >= 0 && voice < e->cfg.n_voices, "voice slot out of range");
    std::lock_guard<std::mutex> g(e->voice_mu);
    e->voices[voice].valid = false;
    CBX_API_END
}

int cbx_t3_open(cbx_engine* e, int voice, const int32_t* text_ids_h, int n_text, float cfg_weight, float temperature, float repetition_penalty,
                float min_p, float top_p, uint64_t seed, int max_new_tokens, int* slot_out, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && text_ids_h && slot_out, "bad argument");
    for (int i = 0; i < n_text; i++) CBX_REQUIRE(text_ids_h[i] >= 0 && text_ids_h[i] < T3_TEXT_V, "text id out of range");
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->t3_mu);
    StreamBridge br((cudaStream_t)stream, e->t3_st, e->t3_ev_in, e->t3_ev_out);
    *slot_out = t3_open(e, voice, text_ids_h, n_text, cfg_weight, temperature, repetition_penalty, min_p, top_p, seed, max_new_tokens, e->t3_st);
    br.finish();
    CBX_API_END
}

int cbx_t3_open_batch(cbx_engine* e, const cbx_t3_open_req* reqs, int n, int* slots_out, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && reqs && slots_out && n >= 1 && n <= T3_PREFILL_BATCH, "bad argument");
    T3OpenReq rq[T3_PREFILL_BATCH];
    for (int i = 0; i < n; i++) {
        CBX_REQUIRE(reqs[i].text_ids_h, "null text");
        for (int k = 0; k < reqs[i].n_text; k++) CBX_REQUIRE(reqs[i].text_ids_h[k] >= 0 && reqs[i].text_ids_h[k] < T3_TEXT_V, "text id out of range");
        rq[i] = T3OpenReq{reqs[i].voice, reqs[i].text_ids_h, reqs[i].n_text, reqs[i].cfg_weight, reqs[i].temperature, reqs[i].repetition_penalty,
                          reqs[i].min_p, reqs[i].top_p, reqs[i].seed, reqs[i].max_new_tokens};
    }
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->t3_mu);
    StreamBridge br((cudaStream_t)stream, e->t3_st, e->t3_ev_in, e->t3_ev_out);
    t3_open_batch(e, rq, n, slots_out, e->t3_st);
    br.finish();
    CBX_API_END
}

int cbx_t3_step(cbx_engine* e, const int32_t* slots_h, int n_slots, int n_steps, const float* noise_d, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && slots_h && n_steps >= 1, "bad argument");
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->t3_mu);
    StreamBridge br((cudaStream_t)stream, e->t3_st, e->t3_ev_in, e->t3_ev_out);
    t3_step(e, slots_h, n_slots, n_steps, noise_d, e->t3_st);
    br.finish();
    CBX_API_END
}

int cbx_t3_set_persistent(cbx_engine* e, int on) {
    CBX_API_BEGIN
    CBX_REQUIRE(e, "null engine");
    std::lock_guard<std::mutex> g(e->t3_mu);
    CBX_REQUIRE(!on || e->t3.mega_ok, "persistent T3 kernel is not available on this device");
    e->t3.mega = on != 0;
    CBX_API_END
}

int cbx_t3_set_priority(cbx_engine* e, int high) {
    CBX_API_BEGIN
    CBX_REQUIRE(e, "null engine");
    std::lock_guard<std::mutex> g(e->t3_mu);
    const int want = high ? 1 : 0;
    if (want != e->t3_prio) {   // later T3 work is ordered after everything already queued on the other stream
        CBX_CHECK(cudaSetDevice(e->device));
        CBX_CHECK(cudaEventRecord(e->t3_ev_sw, e->t3_st));
        CBX_CHECK(cudaStreamWaitEvent(e->t3_st_prio[want], e->t3_ev_sw, 0));
        e->t3_st = e->t3_st_prio[want];
        e->t3_prio = want;
    }
    CBX_API_END
}

int cbx_t3_set_alignment_eos(cbx_engine* e, int on, int layer) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && layer >= 0, "bad argument");
    std::lock_guard<std::mutex> g(e->t3_mu);
    e->t3.align = on != 0;
    e->t3.align_layer = layer;
    CBX_API_END
}

int cbx_t3_alignment_peek(cbx_engine* e, int slot, int32_t* state_out, float* fstate_out, float* row_out, float* pre_out, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && slot >= 0 && slot < e->cfg.max_streams && state_out && fstate_out, "bad argument");
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->t3_mu);
    AlignState s;
    CBX_CHECK(cudaMemcpyAsync(&s, e->t3.align_state + slot, sizeof(s), cudaMemcpyDeviceToHost, e->t3_st));
    CBX_CHECK(cudaStreamSynchronize(e->t3_st));
    const int32_t iv[13] = {s.on, s.i0, s.S, s.frame_pos, s.text_pos, s.T, s.started, s.started_at, s.complete, s.completed_at, s.has_pre, s.ctl, s.cur_posn};
    const float fv[6] = {s.first4_max, s.prev_last2, s.tail3[0], s.tail3[1], s.tail3[2], s.rep_sum};
    std::copy(iv, iv + 13, state_out);
    std::copy(fv, fv + 6, fstate_out);
    if (row_out && s.S > 0) CBX_CHECK(cudaMemcpyAsync(row_out, e->t3.align_cur + (long)slot * e->t3.align_ld, (size_t)s.S * 4, cudaMemcpyDeviceToHost, e->t3_st));
    if (pre_out && s.S > 0) CBX_CHECK(cudaMemcpyAsync(pre_out, e->t3.align_pre + (long)slot * e->t3.align_ld, (size_t)s.S * 4, cudaMemcpyDeviceToHost, e->t3_st));
    CBX_CHECK(cudaStreamSynchronize(e->t3_st));
    CBX_API_END
}

int cbx_t3_alignment_poke(cbx_engine* e, int slot, const int32_t* iv, const float* fv, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && slot >= 0 && slot < e->cfg.max_streams && iv && fv, "bad argument");
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->t3_mu);
    AlignState s{};
    s.on = iv[0]; s.i0 = iv[1]; s.S = iv[2]; s.frame_pos = iv[3]; s.text_pos = iv[4]; s.T = iv[5]; s.started = iv[6]; s.started_at = iv[7]; s.complete = iv[8];
    s.completed_at = iv[9]; s.has_pre = iv[10]; s.ctl = iv[11]; s.cur_posn = iv[12];
    s.first4_max = fv[0]; s.prev_last2 = fv[1]; s.tail3[0] = fv[2]; s.tail3[1] = fv[3]; s.tail3[2] = fv[4]; s.rep_sum = fv[5];
    CBX_REQUIRE(s.S >= 1 && s.S <= e->cfg.max_text && s.i0 >= 0, "alignment poke: text span out of range");
    CBX_CHECK(cudaMemcpyAsync(e->t3.align_state + slot, &s, sizeof(s), cudaMemcpyHostToDevice, e->t3_st));
    CBX_CHECK(cudaStreamSynchronize(e->t3_st));
    CBX_API_END
}

int cbx_op_alignment_run(const float* rows_d, int n_rows, int S, int n_pre, int32_t* steps_out_h, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(rows_d && steps_out_h && S >= 1 && (n_pre == 0 || n_pre == 1) && n_rows > n_pre, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    AlignState* sd; int *ctl, *slots;
    CBX_CHECK(cudaMalloc(&sd, sizeof(AlignState)));
    CBX_CHECK(cudaMalloc(&ctl, 4)); CBX_CHECK(cudaMalloc(&slots, 4));
    CBX_CHECK(cudaMemsetAsync(slots, 0, 4, st));
    launch_align_init(sd, ctl, 0, 1, 0, S, n_pre, st);
    for (int r = n_pre, f = 0; r < n_rows; r++, f++) {
        AlignStepParams p; p.slots = slots; p.state = sd; p.a_cur = rows_d + (long)r * S; p.a_pre = rows_d; p.ld = S; p.ctl = ctl;
        launch_align_step(p, 1, st);
        AlignState s;
        CBX_CHECK(cudaMemcpyAsync(&s, sd, sizeof(s), cudaMemcpyDeviceToHost, st));
        CBX_CHECK(cudaStreamSynchronize(st));
        steps_out_h[f * 4 + 0] = s.ctl; steps_out_h[f * 4 + 1] = s.text_pos; steps_out_h[f * 4 + 2] = s.started; steps_out_h[f * 4 + 3] = s.complete;
    }
    cudaFree(sd); cudaFree(ctl); cudaFree(slots);
    CBX_API_END
}

int cbx_t3_poll(cbx_engine* e, int slot, int* n_generated, int* done, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && slot >= 0 && slot < e->cfg.max_streams, "slot out of range");
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->t3_mu);
    T3SlotState s;
    CBX_CHECK(cudaMemcpyAsync(&s, e->t3.slot_state + slot, sizeof(s), cudaMemcpyDeviceToHost, e->t3_st));
    CBX_CHECK(cudaStreamSynchronize(e->t3_st));
    *n_generated = s.step; *done = s.done;
    CBX_API_END
}

int cbx_t3_tokens(cbx_engine* e, int slot, int from, int count, int32_t* out_h, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && slot >= 0 && slot < e->cfg.max_streams && from >= 0 && count >= 0 && from + count <= e->t3.out_stride, "range error");
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->t3_mu);
    if (count) CBX_CHECK(cudaMemcpyAsync(out_h, e->t3.out_tokens + (long)slot * e->t3.out_stride + from, (size_t)count * 4, cudaMemcpyDeviceToHost, e->t3_st));
    CBX_CHECK(cudaStreamSynchronize(e->t3_st));
    CBX_API_END
}

int cbx_t3_logits(cbx_engine* e, int slot, float* out_h, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && slot >= 0 && slot < e->cfg.max_streams && out_h, "bad argument");
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->t3_mu);
    CBX_CHECK(cudaMemcpy2DAsync(out_h, T3_V * 4, e->t3.logits + (long)slot * 2 * T3_VPAD, T3_VPAD * 4, T3_V * 4, 2, cudaMemcpyDeviceToHost, e->t3_st));
    CBX_CHECK(cudaStreamSynchronize(e->t3_st));
    CBX_API_END
}

int cbx_t3_close(cbx_engine* e, int slot) {
    CBX_API_BEGIN
    CBX_REQUIRE(e, "null engine");
    CBX_CHECK(cudaSetDevice(e->device));
    std::lock_guard<std::mutex> g(e->t3_mu);
    CBX_CHECK(cudaStreamSynchronize(e->t3_st));
    t3_close(e, slot);
    CBX_API_END
}

int cbx_engine_health(cbx_engine* e) {
    CBX_API_BEGIN
    CBX_REQUIRE(e, "null engine");
    CBX_CHECK(cudaSetDevice(e->device));
    const cudaError_t q = cudaStreamQuery(e->t3_st);      // sticky faults surface on any call; pending work is not an error
    if (q != cudaSuccess && q != cudaErrorNotReady) CBX_CHECK(q);
    CBX_API_END
}

int cbx_t3_stats(cbx_engine* e, int* free_pages, int* open_slots) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && free_pages && open_slots, "null argument");
    std::lock_guard<std::mutex> g(e->t3_mu);
    *free_pages = (int)e->t3.free_pages.size();
    int n = 0;
    for (int u : e->t3.slot_used) n += u ? 1 : 0;
    *open_slots = n;
    CBX_API_END
}

int cbx_flow_infer(cbx_engine* e, int voice, const int32_t* tokens_h, int n, float* mel_out_d, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && tokens_h && mel_out_d, "bad argument");
    CBX_REQUIRE(voice >= 0 && voice < e->cfg.n_voices && e->voices[voice].valid, "voice slot is empty");
    CBX_CHECK(cudaSetDevice(e->device));
    Lane& L = pick_lane(e);
    std::lock_guard<std::mutex> g(L.lock);
    StreamBridge br((cudaStream_t)stream, L.st, L.ev_in, L.ev_out);
    const Voice& v = e->voices[voice];
    flow_infer(e, L, v, tokens_h, n, L.st);
    CBX_CHECK(cudaMemcpyAsync(mel_out_d, L.mel, (size_t)2 * n * MEL * 4, cudaMemcpyDeviceToDevice, L.st));
    br.finish();
    CBX_API_END
}

int cbx_hift_infer(cbx_engine* e, const float* mel_d, int frames, const float* cache_source_d, int64_t m, float* wav_out_d, float* source_out_d,
                   const float* phase_h, const float* noise_d, uint64_t seed, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && mel_d && wav_out_d && source_out_d, "bad argument");
    CBX_CHECK(cudaSetDevice(e->device));
    Lane& L = pick_lane(e);
    std::lock_guard<std::mutex> g(L.lock);
    StreamBridge br((cudaStream_t)stream, L.st, L.ev_in, L.ev_out);
    CBX_REQUIRE(frames >= 1 && frames <= 2 * e->cfg.max_s3_tokens, "hift: mel length out of range");
    CBX_CHECK(cudaMemcpyAsync(L.mel, mel_d, (size_t)frames * MEL * 4, cudaMemcpyDeviceToDevice, L.st));
    hift_infer(e, L, frames, cache_source_d, m, wav_out_d, source_out_d, phase_h, noise_d, seed, L.st);
    br.finish();
    CBX_API_END
}

int cbx_hift_f0(cbx_engine* e, const float* mel_d, int frames, float* f0_out_d, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && mel_d && f0_out_d && frames >= 1 && frames <= 2 * e->cfg.max_s3_tokens, "bad argument");
    CBX_CHECK(cudaSetDevice(e->device));
    Lane& L = pick_lane(e);
    std::lock_guard<std::mutex> g(L.lock);
    StreamBridge br((cudaStream_t)stream, L.st, L.ev_in, L.ev_out);
    CBX_CHECK(cudaMemcpyAsync(L.mel, mel_d, (size_t)frames * MEL * 4, cudaMemcpyDeviceToDevice, L.st));
    hift_f0(e, L, frames, L.st);
    CBX_CHECK(cudaMemcpyAsync(f0_out_d, L.h_f0, (size_t)frames * 4, cudaMemcpyDeviceToDevice, L.st));
    br.finish();
    CBX_API_END
}

int cbx_hift_source(cbx_engine* e, const float* f0_d, int frames, const float* phase_h, const float* noise_d, uint64_t seed,
                    float* source_out_d, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && f0_d && source_out_d && frames >= 1 && frames <= 2 * e->cfg.max_s3_tokens, "bad argument");
    CBX_CHECK(cudaSetDevice(e->device));
    Lane& L = pick_lane(e);
    std::lock_guard<std::mutex> g(L.lock);
    StreamBridge br((cudaStream_t)stream, L.st, L.ev_in, L.ev_out);
    hift_source(e, L, f0_d, frames, nullptr, 0, source_out_d, phase_h, noise_d, seed, L.st);
    br.finish();
    CBX_API_END
}

int cbx_s3gen_infer(cbx_engine* e, int voice, const int32_t* tokens_h, int n, const float* cache_source_d, int64_t m, float* wav_out_d,
                    float* source_out_d, float* mel_out_d, const float* phase_h, const float* noise_d, uint64_t seed, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && tokens_h && wav_out_d && source_out_d, "bad argument");
    CBX_REQUIRE(voice >= 0 && voice < e->cfg.n_voices && e->voices[voice].valid, "voice slot is empty");
    CBX_REQUIRE(n >= 3, "s3gen: needs at least 3 tokens (reference pads, src/tts_streaming.py:675-677)");
    CBX_CHECK(cudaSetDevice(e->device));
    Lane& L = pick_lane(e);
    std::lock_guard<std::mutex> g(L.lock);
    StreamBridge br((cudaStream_t)stream, L.st, L.ev_in, L.ev_out);
    const Voice& v = e->voices[voice];
    const long Ls = 960L * n;
    CBX_REQUIRE(m >= 0, "s3gen: negative cache_source length");
    // upstream raises a shape error when the cache is longer than the new source (it can happen when an SOS id inside the
    // accumulated tokens makes drop_invalid_tokens shorten the sequence); here the cache is clipped to the new length
    if (m > Ls) m = Ls;
    L.nb = 1; L.call[0].v = &v; L.call[0].n = n;
    const int* toks[1] = {tokens_h};
    flow_stage(e, L, toks, L.st);
    const bool direct = phase_h || noise_d || prof_enabled();
    if (direct) {
        // explicit SineGen randomness (parity tests) or per-launch profiling: plain launches
        flow_run(e, L, L.st);
        CBX_CHECK(cudaMemcpyAsync(L.mel, L.melb, (size_t)2 * n * MEL * 4, cudaMemcpyDeviceToDevice, L.st));
        hift_infer(e, L, 2 * n, cache_source_d, m, wav_out_d, source_out_d, phase_h, noise_d, seed, L.st);
    } else {
        // steady state: the whole device-side call (~5k launches) is one CUDA graph per (voice, prompt, n) shape
        if (m) CBX_CHECK(cudaMemcpyAsync(L.g_cache, cache_source_d, (size_t)m * 4, cudaMemcpyDeviceToDevice, L.st));
        CBX_CHECK(cudaStreamSynchronize(L.st));          // previous call's read of the pinned params has retired
        L.g_dyn_h->cache_len = m; L.g_dyn_h->seed = seed;
        CBX_CHECK(cudaMemcpyAsync(L.g_dyn, L.g_dyn_h, sizeof(SourceDyn), cudaMemcpyHostToDevice, L.st));
        const unsigned long long key = ((unsigned long long)voice << 48) | ((unsigned long long)(v.version & 0xFFFF) << 32) | ((unsigned long long)v.n_prompt << 16) | (unsigned long long)n |
                                       (flow_tail_path(4L * (v.n_prompt + n)) ? 1ull << 63 : 0ull);     // which kernels the capture holds
        auto it = L.graphs.find(key);
        if (it == L.graphs.end()) {
            const long before = e->gpu_launches.load();
            cudaGraph_t graph;
            CBX_CHECK(cudaStreamBeginCapture(L.st, cudaStreamCaptureModeThreadLocal));
            try {
                flow_run(e, L, L.st);
                CBX_CHECK(cudaMemcpyAsync(L.mel, L.melb, (size_t)2 * n * MEL * 4, cudaMemcpyDeviceToDevice, L.st));
                hift_infer(e, L, 2 * n, L.g_cache, 0, L.g_wav, L.g_src, nullptr, nullptr, 0, L.st, L.g_dyn);
            } catch (...) {
                cudaGraph_t junk; cudaStreamEndCapture(L.st, &junk);
                throw;
            }
            CBX_CHECK(cudaStreamEndCapture(L.st, &graph));
            cudaGraphExec_t exec;
            CBX_CHECK(cudaGraphInstantiate(&exec, graph, 0));
            CBX_CHECK(cudaGraphDestroy(graph));
            if (L.graphs.size() >= 64) { for (auto& kv : L.graphs) cudaGraphExecDestroy(kv.second); L.graphs.clear(); }
            it = L.graphs.emplace(key, exec).first;
            L.launches_per_graph[key] = e->gpu_launches.load() - before;
            e->gpu_launches = before;
        }
        CBX_CHECK(cudaGraphLaunch(it->second, L.st));
        e->gpu_launches += L.launches_per_graph[key];
        CBX_CHECK(cudaMemcpyAsync(wav_out_d, L.g_wav, (size_t)Ls * 4, cudaMemcpyDeviceToDevice, L.st));
        CBX_CHECK(cudaMemcpyAsync(source_out_d, L.g_src, (size_t)Ls * 4, cudaMemcpyDeviceToDevice, L.st));
    }
    if (mel_out_d) CBX_CHECK(cudaMemcpyAsync(mel_out_d, L.mel, (size_t)2 * n * MEL * 4, cudaMemcpyDeviceToDevice, L.st));
    br.finish();
    CBX_API_END
}

int cbx_s3gen_infer_batch(cbx_engine* e, const cbx_s3gen_call* calls, int n_calls, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && calls && n_calls >= 1 && n_calls <= FLOW_MAXB, "s3gen batch: between 1 and 16 calls");
    CBX_CHECK(cudaSetDevice(e->device));
    // take whichever batch workspace is free (two batches may overlap on the GPU: their kernels are latency-bound)
    Lane* Lp = e->batch_lanes[0];
    std::unique_lock<std::mutex> g(Lp->lock, std::try_to_lock);
    if (!g.owns_lock()) { Lp = e->batch_lanes[1]; g = std::unique_lock<std::mutex>(Lp->lock, std::try_to_lock); }
    if (!g.owns_lock()) { Lp = e->batch_lanes[0]; g = std::unique_lock<std::mutex>(Lp->lock); }
    Lane& L = *Lp;
    StreamBridge br((cudaStream_t)stream, L.st, L.ev_in, L.ev_out);
    const int* toks[FLOW_MAXB];
    L.nb = n_calls;
    for (int b = 0; b < n_calls; b++) {
        const cbx_s3gen_call& c = calls[b];
        CBX_REQUIRE(c.tokens_h && c.wav_out_d && c.source_out_d, "s3gen batch: null pointer in call");
        CBX_REQUIRE(c.voice >= 0 && c.voice < e->cfg.n_voices && e->voices[c.voice].valid, "voice slot is empty");
        CBX_REQUIRE(c.n >= 3, "s3gen: needs at least 3 tokens (reference pads, src/tts_streaming.py:675-677)");
        CBX_REQUIRE(c.m >= 0, "s3gen: negative cache_source length");
        L.call[b].v = &e->voices[c.voice]; L.call[b].n = c.n; toks[b] = c.tokens_h;
    }
    flow_stage(e, L, toks, L.st);
    // the token -> mel part (encoder + 10 CFM steps, ~95 % of the launches) runs once for the whole batch: every GEMM /
    // norm / attention launch covers all calls, so the per-launch latency that bounds a single call is shared
    flow_run(e, L, L.st);
    const long mel_bs = 2L * e->cfg.max_s3_tokens * MEL;
    // the vocoder runs per call (different lengths, source caches and seeds) on the lane's vocoder streams: calls are dealt
    // round-robin over them, a call whose cache_source is an earlier call's source output waits for that call
    CBX_CHECK(cudaEventRecord(L.ev_flow, L.st));
    const int nsub = prof_enabled() ? 1 : (int)L.sub.size();   // per-launch profiling brackets launches with events on ONE stream
    for (int b = 0; b < n_calls; b++) {
        const cbx_s3gen_call& c = calls[b];
        const long Ls = 960L * c.n;
        Lane& S = *L.sub[b % nsub];
        cudaStream_t hs = prof_enabled() ? L.st : S.st;
        if (!prof_enabled()) {
            CBX_CHECK(cudaStreamWaitEvent(hs, L.ev_flow, 0));
            for (int a = 0; a < b; a++)
                if (c.cache_source_d && c.cache_source_d == calls[a].source_out_d) CBX_CHECK(cudaStreamWaitEvent(hs, L.ev_call[a], 0));
        }
        CBX_CHECK(cudaMemcpyAsync(S.mel, L.melb + b * mel_bs, (size_t)2 * c.n * MEL * 4, cudaMemcpyDeviceToDevice, hs));
        // decode window: the caller only consumes wav_out[emit_from:] ("full" overlap: the audio of earlier slices has been sent)
        int w0 = 0;
        if (c.emit_from > 0) { w0 = (int)(std::min<int64_t>(c.emit_from, Ls) / H_UP) - HIFT_WINDOW_MARGIN; if (w0 < 0) w0 = 0; if (w0 >= 2 * c.n) w0 = 2 * c.n - 1; }
        hift_infer(e, S, 2 * c.n, c.cache_source_d, c.m > Ls ? Ls : c.m, c.wav_out_d, c.source_out_d, nullptr, nullptr, c.seed, hs, nullptr, w0);
        if (c.mel_out_d) CBX_CHECK(cudaMemcpyAsync(c.mel_out_d, S.mel, (size_t)2 * c.n * MEL * 4, cudaMemcpyDeviceToDevice, hs));
        if (!prof_enabled()) CBX_CHECK(cudaEventRecord(L.ev_call[b], hs));
    }
    if (!prof_enabled())
        for (int b = 0; b < n_calls; b++) CBX_CHECK(cudaStreamWaitEvent(L.st, L.ev_call[b], 0));
    br.finish();
    CBX_API_END
}

int cbx_crossfade_pcm(cbx_engine* e, const float* cur_d, int64_t n_out, const float* prev_tail_d, int fade_len, const float* fade_in_d,
                      const float* fade_out_d, int16_t* out_d, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && cur_d && out_d && n_out >= 0, "bad argument");
    CBX_REQUIRE(!prev_tail_d || fade_len == 0 || (fade_in_d && fade_out_d), "crossfade: a previous tail needs both fade curves");
    CBX_CHECK(cudaSetDevice(e->device));
    if (n_out > 0) {
        crossfade_pcm_kernel<<<cdiv(n_out, 256), 256, 0, (cudaStream_t)stream>>>(cur_d, n_out, fade_len > 0 ? prev_tail_d : nullptr, fade_len, fade_in_d, fade_out_d, out_d);
        CBX_CHECK(cudaGetLastError());
        e->gpu_launches += 1;
    }
    CBX_API_END
}

int64_t cbx_gpu_launches(cbx_engine* e) { return e ? (int64_t)e->gpu_launches.load() : (int64_t)-1; }

static void ops_init_once() {
    static std::once_flag once;
    std::call_once(once, [] { gemm_init(); attention_init(); });
}

int cbx_op_gemm(const void* a, const void* w, const float* bias, float* out, int M, int N, int K, void* stream) {
    CBX_API_BEGIN
    ops_init_once();
    GemmParams g; g.A = (const bf16*)a; g.lda = K; g.kc = K; g.W = (const bf16*)w; g.ldw = K; g.M = M; g.N = N; g.K = K; g.bias = bias; g.outF = out; g.ldc = N;
    launch_gemm(g, (cudaStream_t)stream);
    CBX_API_END
}

int cbx_op_gemm_ex(const void* a, const void* w, const float* bias, const float* res, float* out_f, void* out_b, int M, int N, int K, int act, void* stream) {
    CBX_API_BEGIN
    ops_init_once();
    CBX_REQUIRE(out_f || out_b, "gemm_ex: no output");
    GemmParams g; g.A = (const bf16*)a; g.lda = K; g.kc = K; g.W = (const bf16*)w; g.ldw = K; g.M = M; g.N = N; g.K = K; g.bias = bias;
    g.act = act; g.res = res; g.ldr = N; g.outF = out_f; g.outB = (bf16*)out_b; g.ldc = N;
    launch_gemm(g, (cudaStream_t)stream);
    CBX_API_END
}

int cbx_op_cfm_tail(int mode, int M, const void* attn_o, float* h, const void* wout, const float* b_out, const float* ln3_g, const float* ln3_b,
                    const void* w0, const float* b0, const void* w2, const float* b2, const float* ln1_g, const float* ln1_b, const void* wqkv,
                    void* qkv_out, void* stream) {
    CBX_API_BEGIN
    ops_init_once();
    CBX_REQUIRE(cfm_tail_available(), "cfm_tail kernel is not available on this device");
    CfmTailWeights blk, nxt;
    cfm_tail_weights(blk, (const bf16*)wout, (const bf16*)w0, (const bf16*)w2, nullptr);
    cfm_tail_weights(nxt, nullptr, nullptr, nullptr, (const bf16*)wqkv);
    CfmTailArgs a; a.M = M; a.mode = mode; a.h = h; a.b_out = b_out; a.ln3_g = ln3_g; a.ln3_b = ln3_b; a.b0 = b0; a.b2 = b2; a.ln1_g = ln1_g; a.ln1_b = ln1_b;
    a.qkv = (bf16*)qkv_out;
    launch_cfm_tail(a, (const bf16*)attn_o, &blk, &nxt, (cudaStream_t)stream);
    CBX_API_END
}

int cbx_op_attention(const void* qkv, void* out, int T, int H, int batch, int causal, void* stream) {
    CBX_API_BEGIN
    ops_init_once();
    const long ld = 3L * H * 64;
    AttnParams a; a.q = (const bf16*)qkv; a.k = a.q + H * 64; a.v = a.q + 2 * H * 64; a.ldq = a.ldk = a.ldv = ld; a.q_bs = a.k_bs = a.v_bs = (long)T * ld;
    a.o = (bf16*)out; a.ldo = H * 64; a.o_bs = (long)T * H * 64; a.T = T; a.H = H; a.batch = batch; a.causal = causal; a.scale = 0.125f;
    launch_attention(a, (cudaStream_t)stream);
    CBX_API_END
}

// windowed vocoder decode (tests): as cbx_hift_infer, the convolution stack over mel frames [w0, frames) only
int cbx_hift_infer_window(cbx_engine* e, const float* mel_d, int frames, const float* cache_source_d, int64_t m, float* wav_out_d, float* source_out_d,
                          uint64_t seed, int w0, void* stream) {
    CBX_API_BEGIN
    CBX_REQUIRE(e && e->finalized && mel_d && wav_out_d && source_out_d, "bad argument");
    CBX_CHECK(cudaSetDevice(e->device));
    Lane& L = pick_lane(e);
    std::lock_guard<std::mutex> g(L.lock);
    StreamBridge br((cudaStream_t)stream, L.st, L.ev_in, L.ev_out);
    CBX_REQUIRE(frames >= 1 && frames <= 2 * e->cfg.max_s3_tokens, "hift: mel length out of range");
    CBX_CHECK(cudaMemcpyAsync(L.mel, mel_d, (size_t)frames * MEL * 4, cudaMemcpyDeviceToDevice, L.st));
    hift_infer(e, L, frames, cache_source_d, m, wav_out_d, source_out_d, nullptr, nullptr, seed, L.st, nullptr, w0);
    br.finish();
    CBX_API_END
}

// attention with an additive relative-position bias (flow encoder): bias fp32 [batch][H][T][2T], entry (i, j) at column T - 1 - i + j
int cbx_op_attention_bias(const void* qkv, const float* bias, void* out, int T, int H, int batch, void* stream) {
    CBX_API_BEGIN
    ops_init_once();
    const long ld = 3L * H * 64;
    AttnParams a; a.q = (const bf16*)qkv; a.k = a.q + H * 64; a.v = a.q + 2 * H * 64; a.ldq = a.ldk = a.ldv = ld; a.q_bs = a.k_bs = a.v_bs = (long)T * ld;
    a.o = (bf16*)out; a.ldo = H * 64; a.o_bs = (long)T * H * 64; a.T = T; a.H = H; a.batch = batch; a.causal = 0; a.scale = 0.125f;
    a.relbias = bias; a.rb_ld = 2L * T; a.rb_hs = (long)T * 2 * T; a.rb_bs = (long)H * T * 2 * T;
    launch_attention(a, (cudaStream_t)stream);
    CBX_API_END
}

}  // extern "C"
