// Internal engine structures: weight registry, T3 / flow / HiFT models and per-lane workspaces.
#pragma once
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "t3_kernels.cuh"
#include "hift_kernels.cuh"
#include "cfm_tail.cuh"
#include "../../include/cbx_b200.h"

enum DType { DT_F32 = 0, DT_BF16 = 1 };

struct TensorRec { std::string name; int dtype; long numel; void* ptr; bool loaded; };

// fixed dimensions of the Chatterbox path; depth / step counts come from cbx_config
namespace dims {
constexpr int T3_D = 1024, T3_H = 16, T3_FFN = 4096, T3_V = 8194, T3_VPAD = 8208, T3_TEXT_V = 704, T3_TEXT_POS = 2050,
              T3_SPEECH_POS = 4100, T3_SPK = 256, T3_PQ = 32, T3_PH = 4, T3_COND = 34, T3_BOS = 6561, T3_EOS = 6562, PAGE = 16;
constexpr int F_V = 6561, F_D = 512, F_H = 8, F_FFN = 2048, F_SPK = 192, MEL = 80, C_IN = 320, C_CH = 256, C_INNER = 512,
              C_FF = 1024, C_TDIM = 1024, NOISE_LEN = 15000;
constexpr int H_BASE = 512, H_NSRC = 18, H_NSRC_PAD = 64, MEL_PAD = 128, H_HALO = 32, H_UP = 480, H_NHARM = 9, H_F0CH = 512;
}  // namespace dims

struct Lin { bf16* w = nullptr; float* b = nullptr; int N = 0, K = 0; };
struct LNp { float* g = nullptr; float* b = nullptr; };

struct T3Layer {
    float *ln1, *ln2; bf16 *wqkv, *wo, *wgu, *wd; bf16 *wqkv_f, *wo_f, *wgu_f, *wd_f;
    alignas(64) unsigned char tm_qkv[128], tm_o[128], tm_gu[128], tm_d[128];     // TMA descriptors of the row-major weights (batched decode on tcgen05)
};
struct T3Model {
    float *text_emb, *speech_emb, *text_pos, *speech_pos, *final_norm, *inv_freq;
    bf16* head_f;
    Lin spkr, pq, pkv, po; float *emo_w, *perc_query; LNp pnorm;
    std::vector<T3Layer> layers;
    // pools / decode state (device)
    bf16* kv = nullptr; long kv_layer_stride = 0, kv_half = 0; int max_pages = 0, total_pages = 0;
    int* page_table = nullptr; T3SlotState* slot_state = nullptr; int* slot_pos = nullptr; uint8_t* seen = nullptr;
    int* out_tokens = nullptr; int out_stride = 0; float *x, *qkv, *attn, *act, *logits;
    int *d_slots = nullptr, *d_rowmap = nullptr;   // active set staging [max_streams], [2*max_streams]
    // alignment-based EOS control (off by default; cbx_t3_set_alignment_eos): per-slot analyzer state, decision word and the
    // newest alignment row(s) over the text span
    bool align = false, align_joined = false; int align_layer = 9; AlignState* align_state = nullptr; int* align_ctl = nullptr;
    float *align_cur = nullptr, *align_pre = nullptr, *align_q = nullptr; long align_ld = 0;
    cudaStream_t align_st = nullptr; cudaEvent_t align_fork = nullptr, align_join = nullptr;   // the analyzer runs beside the layers after the probe
    // megakernel state
    bool mega = false, mega_ok = false; MegaState mega_state; MegaLayer* d_layers = nullptr; unsigned long long* ll[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; unsigned int* epoch = nullptr;
    // prefill workspace
    bf16 *xb = nullptr, *attn_b = nullptr, *act_b = nullptr; float* ss = nullptr;   // bf16 hand-over buffers of the decode step
    float* pf_x; bf16 *pf_xn, *pf_qkv, *pf_att, *pf_act; int* pf_text; int pf_max = 0;
    // host side
    std::vector<int> free_pages; std::vector<int> slot_used; std::vector<std::vector<int>> slot_pages; std::vector<int> slot_maxnew; std::vector<long> slot_pos_h;
    std::vector<int> h_active;   // last active set uploaded
    std::unordered_map<int, cudaGraphExec_t> step_graphs;   // keyed by n_streams
};

struct ConformerLayer { LNp nm, nf; Lin qkv4, pos, out, w1, w2; };
struct ResnetP { Lin c1, c2, res; LNp n1, n2; int cin; };
struct TfmP { LNp n1, n3; Lin qkv, out, ff0, ff2; CfmTailWeights tw; };
struct FlowModel {
    float* tok_emb; Lin spk_aff_dummy; float *spk_w, *spk_b; Lin enc_proj;
    Lin embed, up_embed; LNp embed_ln, up_embed_ln, after_norm; Lin pl1, pl2, upconv;
    std::vector<ConformerLayer> enc, up;
    // estimator
    Lin t1, t2; std::vector<Lin> tmlp;          // per-resnet time projections (used at finalize only)
    std::vector<ResnetP> resnets; std::vector<TfmP> tfms;   // resnets: down, mid..., up ; tfms: n_blocks per resnet
    Lin down_conv, up_conv, final_conv, final_proj; LNp final_ln;
    float* noise;       // [NOISE_LEN][80] time-major
    float* tproj;       // [n_resnets][n_steps][256] precomputed time-embedding projections
    std::vector<float> t_span;
};

struct ResBlockP { Lin c1[3], c2[3]; float *a1[3], *a2[3]; int ch, k; };
struct HiftModel {
    Lin conv_pre, conv_post; Lin ups[3]; Lin sdown[3]; ResBlockP sres[3]; ResBlockP res[9];
    Lin f0c[5]; float *f0w, *f0b, *lw, *lb; float* fade;
};

struct Voice {
    bool valid = false; int version = 0; float* prefix = nullptr;   // [34][1024]
    int n_prompt = 0, n_feat = 0; int* prompt_token = nullptr; float* prompt_feat = nullptr; float* spks = nullptr;
};

struct FlowCall { const Voice* v = nullptr; int n = 0; int Tt = 0; };   // one S3Gen call of a batch: voice, new tokens, prompt + new tokens
constexpr int FLOW_MAXB = 16;

struct Lane {   // S3Gen workspace: one batch of up to `bmax` calls at a time
    std::mutex lock; cudaStream_t st; cudaEvent_t ev_in, ev_out;
    int bmax = 1, nb = 1, Ttm = 0; FlowCall call[FLOW_MAXB];   // current batch; sequences are right-padded to Ttm tokens
    float* melb = nullptr;   // [bmax][2*max_s3_tokens][80] CFM output per call
    std::vector<Lane*> sub;  // batch lanes: vocoder-only workspaces with their own streams (the calls' HiFT passes run in parallel)
    cudaEvent_t ev_flow = nullptr; std::vector<cudaEvent_t> ev_call;
    // encoder
    int* tok; bf16 *e_in, *e_xb, *e_y1, *e_xn, *e_qkv, *e_pos, *e_p, *e_o, *e_ff, *e_up, *e_upc; float *e_tmp, *e_x, *e_bd;
    // cfm
    float *mu, *cond, *x, *v; bf16 *c_in, *c_hb, *c_inb, *c_upin, *c_xn, *c_qkv, *c_o, *c_ff, *c_fb; float *c_tmp, *c_tmp2, *c_h;
    // hift
    float* mel; bf16 *h_mel, *h_f0a, *h_f0b, *h_stft; float *h_f0, *h_s; double* h_cum;
    bf16* h_xb[4];          // halo'd bf16 inputs of conv_pre output / ups outputs (lrelu'd)
    float *h_x[3], *h_r[3], *h_acc[3], *h_si[3]; bf16 *h_a[3], *h_b[3]; float* h_post; float* h_phase;
    // graph-replayed s3gen calls: lane-owned I/O buffers + per-call dynamic params, graphs keyed by (voice, n_prompt, n)
    float *g_wav, *g_src, *g_cache; SourceDyn* g_dyn; SourceDyn* g_dyn_h;
    std::unordered_map<unsigned long long, cudaGraphExec_t> graphs;
    std::unordered_map<unsigned long long, long> launches_per_graph;
};

struct cbx_engine {
    cbx_config cfg; int device = 0;
    std::vector<TensorRec> tensors; std::unordered_map<std::string, int> index;
    bool finalized = false;
    bool dry = false;   // manifest-only engine (no device): lets CPU tests check the packer against the registry
    T3Model t3; FlowModel flow; HiftModel hift;
    std::vector<Voice> voices; std::mutex voice_mu;
    std::mutex t3_mu; cudaStream_t t3_st; cudaEvent_t t3_ev_in, t3_ev_out;
    cudaStream_t t3_st_prio[2] = {nullptr, nullptr}; cudaEvent_t t3_ev_sw = nullptr; int t3_prio = 0;   // T3 work moves between a low- and a high-priority stream (cbx_t3_set_priority)
    std::vector<Lane*> lanes; std::mutex lane_pick_mu; int lane_rr = 0;
    std::vector<Lane*> batch_lanes;   // workspaces of cbx_s3gen_infer_batch (FLOW_MAXB calls each): two batches can be in flight
    std::atomic<long> gpu_launches{0};   // statistics only, bumped from the T3, S3Gen and request threads

    template <typename T> T* reg(const std::string& name, int dtype, long numel);
    template <typename T> T* scratch(long numel);
    std::vector<void*> scratch_allocs;
};

// model construction (registers tensors) and execution
void t3_build(cbx_engine* e);
void t3_alloc(cbx_engine* e);
void t3_voice_prefix(cbx_engine* e, Voice& v, const float* speaker_emb_h, const int* cond_tokens_h, int n_cond, float emotion, cudaStream_t st);
int t3_open(cbx_engine* e, int voice, const int* text_ids_h, int L, float cfg_w, float temp, float rep, float min_p, float top_p,
            unsigned long long seed, int max_new, cudaStream_t st);
struct T3OpenReq { int voice; const int* text_ids_h; int L; float cfg_w, temp, rep, min_p, top_p; unsigned long long seed; int max_new; };
constexpr int T3_PREFILL_BATCH = 8;      // requests prefilled in one pass (workspace is sized for this many)
void t3_open_batch(cbx_engine* e, const T3OpenReq* reqs, int n, int* slots_out, cudaStream_t st);
void t3_step(cbx_engine* e, const int* slots, int n, int n_steps, const float* noise_dev, cudaStream_t st);
void t3_close(cbx_engine* e, int slot);

void flow_build(cbx_engine* e);
void flow_finalize(cbx_engine* e, cudaStream_t st);
void flow_infer(cbx_engine* e, Lane& L, const Voice& v, const int* tokens_h, int n, cudaStream_t st);   // single call -> L.mel [2n][80]

void hift_build(cbx_engine* e);
void hift_infer(cbx_engine* e, Lane& L, int Tg, const float* cache_src_dev, long m, float* wav_out, float* src_out,
                const float* phase_h, const float* noise_dev, unsigned long long seed, cudaStream_t st, const SourceDyn* dyn = nullptr, int w0 = 0);
constexpr int HIFT_WINDOW_MARGIN = 20;     // mel frames between the start of a decode window and the first sample that equals the full decode
void flow_stage(cbx_engine* e, Lane& L, const int* const* tokens_h, cudaStream_t st);   // L.nb / L.call[] set by the caller
bool flow_tail_path(long rows);                                                          // fused block-tail kernels for a call of this many estimator rows?
void flow_run(cbx_engine* e, Lane& L, cudaStream_t st);                                  // -> L.melb
void hift_f0(cbx_engine* e, Lane& L, int Tg, cudaStream_t st);
void hift_source(cbx_engine* e, Lane& L, const float* f0, int Tg, const float* cache_src_dev, long m, float* src_out,
                 const float* phase_h, const float* noise_dev, unsigned long long seed, cudaStream_t st, const SourceDyn* dyn = nullptr);
void lane_alloc(cbx_engine* e, Lane& L, int bmax);

template <typename T> T* cbx_engine::reg(const std::string& name, int dtype, long numel) {
    CBX_REQUIRE(index.find(name) == index.end(), "duplicate tensor " + name);
    void* p = nullptr;
    size_t bytes = (size_t)numel * (dtype == DT_F32 ? 4 : 2);
    if (!dry) CBX_CHECK(cudaMalloc(&p, bytes < 16 ? 16 : bytes));
    index[name] = (int)tensors.size();
    tensors.push_back({name, dtype, numel, p, false});
    return reinterpret_cast<T*>(p);
}
template <typename T> T* cbx_engine::scratch(long numel) {
    void* p = nullptr;
    if (dry) return nullptr;
    size_t bytes = (size_t)(numel > 0 ? numel : 1) * sizeof(T);
    CBX_CHECK(cudaMalloc(&p, bytes));
    CBX_CHECK(cudaMemset(p, 0, bytes));
    scratch_allocs.push_back(p);
    return reinterpret_cast<T*>(p);
}
