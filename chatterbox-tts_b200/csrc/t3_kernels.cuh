// Parameter blocks for the T3 kernels (t3_kernels.cu).
#pragma once
#include <vector>
#include "common.cuh"

enum GemvEpi { GEMV_STORE = 0, GEMV_RESID = 1, GEMV_GLU = 2 };

struct GemvParams {
    const bf16* Wf = nullptr;     // fragment-ordered weights [n_strips][K/16][32][8]
    int N = 0, K = 0, n_strips = 0, strips_per_cta = 1;
    const float* x = nullptr; long ldx_in = 0;    // fp32 input rows, indexed through row_map
    const int* row_map = nullptr; int rows = 0;   // compact row -> global row (slot*2 + cfg_row)
    const float* gain = nullptr; float eps = 1e-5f;   // non-null: RMSNorm the input rows first
    float* out = nullptr; long ld_out = 0; int epi = GEMV_STORE;
    // bf16 activations between the kernels of a step (rows > 2 make the per-CTA input staging the dominant L2 traffic):
    const bf16* xb = nullptr; long ldxb = 0;      // bf16 input rows (replaces x).  With ss_in they hold x * norm gain and
    const float* ss_in = nullptr; int n_ss = 0;   // the outputs are scaled by rsqrt(sum(ss_in[row][0..n_ss)) / K + eps)
    bf16* out_b = nullptr; long ld_out_b = 0;     // GLU: bf16 activations (instead of out); RESID: bf16(x_new * next_gain)
    const float* next_gain = nullptr;             // RESID: gain of the RMSNorm that consumes the updated residual row
    float* ss_out = nullptr;                      // RESID: [rows_total][n_strips] sum of squares of this strip's 16 new values
    // weights of a LATER kernel of the step, pulled into L2 by this one before it waits for its predecessor: HBM keeps streaming
    // through the dependency latency of the chain, the later kernel's own loads hit L2
    const void* pf_ptr = nullptr; size_t pf_bytes = 0;
};
void launch_gemv(const GemvParams& p, int nwarps, cudaStream_t st);
// tcgen05 form for batched rows (t3_gemv_tc.cu): row-major weights through a TMA descriptor built once per weight
void gemv_tc_init();
bool gemv_tc_available();
void gemv_tc_weight_map(unsigned char (&map)[128], const bf16* w, int N, int K);
bool launch_gemv_tc(const GemvParams& p, const unsigned char (&map)[128], cudaStream_t st);

struct T3SlotState {   // device-resident per-stream decode state
    int pos;        // KV length (positions already cached) of both rows
    int step;       // tokens generated so far
    int max_new;
    int done;
    float cfg_w, temp, rep_pen, min_p, top_p;
    int _pad;
    unsigned long long seed;
};

struct DecodeAttnParams {
    const float* qkv = nullptr;          // [rows_total][3*H*64] fp32 (raw projections)
    float* out = nullptr;                // [rows_total][H*64] fp32 (or out_b: the same as bf16)
    bf16* out_b = nullptr;
    bf16* kv = nullptr; long kv_half = 0;   // this layer's pool: K at kv, V at kv + kv_half; [page][H][16][64]
    const int* page_table = nullptr; int max_pages = 0;   // [rows_total][max_pages]
    const int* slot_pos = nullptr;       // [slots]
    const int* row_map = nullptr;
    const float* inv_freq = nullptr;     // [32]
    int H = 16;
    float* q_save = nullptr;             // optional [rows_total][H*64]: the rotated, scaled queries (alignment probe layer)
    const void* pf_ptr = nullptr; size_t pf_bytes = 0;   // as in GemvParams
};
void launch_decode_attn(const DecodeAttnParams& p, int rows, int max_pos, cudaStream_t st);

struct SamplerParams {
    const int* slots = nullptr;          // [n_streams] active slot ids
    T3SlotState* state = nullptr;
    int* slot_pos = nullptr;
    const float* logits = nullptr; long ld_logits = 0;   // [slots*2][ld]
    uint8_t* seen = nullptr; int seen_stride = 0;
    int* out_tokens = nullptr; int out_stride = 0;
    const float* noise = nullptr; long noise_stride = 0;  // optional explicit Exp(1) noise [n_streams][V]
    float* x = nullptr;                  // [slots*2][dim] next-step input embeddings
    const float* speech_emb = nullptr; const float* speech_pos = nullptr;
    int V = 0, dim = 0, eos = 0;
    // the same rows in the decode step's bf16 hand-over form (layer 0 of the NEXT step on the tcgen05 path): bf16(x * gain0) and
    // per-16-column sums of squares
    bf16* xb = nullptr; float* ss = nullptr; const float* gain0 = nullptr;
    const int* eos_ctl = nullptr;        // optional [slots]: bit 1 forces EOS, bit 0 suppresses it (alignment control, applied after the CFG mix)
};
void launch_sampler(const SamplerParams& p, int n_streams, cudaStream_t st);

// ---- alignment-based EOS control.  Reference side: the model package's AlignmentStreamAnalyzer, hooked on the attention of one
// trunk layer inside T3.inference_stream (the generator src/tts_streaming.py:420-435 primes); SURVEY section 8(f).3.
struct AlignState {     // device-resident per-stream analyzer state
    int on;             // 0: the stream does not use alignment control
    int i0, S;          // text span [i0, i0 + S) of the conditional row's sequence
    int frame_pos;      // curr_frame_pos: columns above it are masked
    int text_pos;       // text_position
    int T;              // alignment rows seen so far
    int started, started_at, complete, completed_at;
    int has_pre;        // the prefilled BOS row waits in a_pre (consumed by the first step)
    int ctl;            // last decision: bit 1 force EOS, bit 0 suppress EOS
    int cur_posn;       // argmax of the last masked row
    float first4_max;   // max over all rows of columns [0, 4)
    float prev_last2;   // max of the previous row's last two columns
    float tail3[3];     // column sums of the last three columns over rows >= completed_at
    float rep_sum;      // sum over rows >= completed_at of the row maximum over columns [0, S - 5)
};
struct AlignAttnParams {
    const int* slots = nullptr; int slot = -1;      // decode: active slot list (one CTA each); prefill: one slot
    AlignState* state = nullptr;
    const float* q_rot = nullptr;                   // decode: rotated queries x 1/8 saved by the decode attention [rows_total][H*64]
    const bf16* q_b = nullptr; int pos = 0;         // prefill: rotated bf16 query row of position pos
    const bf16* kv = nullptr;                       // this layer's K pool [page][H][16][64]
    const int* page_table = nullptr; int max_pages = 0;
    const int* slot_pos = nullptr;
    float* out = nullptr; long ld_out = 0;          // [slots][ld_out] head-averaged probabilities over the text span
    int H = 16;
};
void launch_align_attn(const AlignAttnParams& p, int n_streams, cudaStream_t st);
struct AlignStepParams {
    const int* slots = nullptr; AlignState* state = nullptr; const T3SlotState* t3 = nullptr;
    const float* a_cur = nullptr; const float* a_pre = nullptr; long ld = 0;
    int* ctl = nullptr;                             // [slots]
};
void launch_align_step(const AlignStepParams& p, int n_streams, cudaStream_t st);
void launch_align_init(AlignState* st_dev, int* ctl, int slot, int on, int i0, int S, int has_pre, cudaStream_t st);

struct AssembleParams {
    float* x = nullptr;                  // [2][Lp][dim]
    const float* prefix = nullptr;       // [Lc][dim]
    const float* text_emb = nullptr; const float* text_pos = nullptr;
    const float* speech_emb = nullptr; const float* speech_pos = nullptr;
    const int* text_ids = nullptr;       // [L] device
    int Lc = 0, L = 0, Lp = 0, dim = 0, bos = 0, cfg_on = 0;
    int slab = 0;                        // rows between the two CFG rows' sequences (0: Lp; a batched prefill pads every sequence to the longest)
};
void launch_assemble_embeds(const AssembleParams& p, cudaStream_t st);

struct RopeKvParams {
    bf16* qkv = nullptr; int Lp = 0; int H = 16; int slab = 0;     // slab: as in AssembleParams
    bf16* kv = nullptr; long kv_half = 0;
    const int* page_table = nullptr; int max_pages = 0; int row0 = 0;
    const float* inv_freq = nullptr;
};
void launch_rope_kv_prefill(const RopeKvParams& p, cudaStream_t st);
void launch_init_slot(T3SlotState* st_dev, const T3SlotState& v, int* slot_pos, int slot, uint8_t* seen, int seen_stride, int bos,
                      float* x, const float* speech_emb, const float* speech_pos, int dim, bf16* xb, float* ss, const float* gain0, cudaStream_t st);
void t3_kernels_init();
void launch_prompt_embed(float* out, const float* emb, const float* pos, const int* ids, int n, int D, cudaStream_t st);
void launch_scale_vec(float* out, const float* w, float s, int n, cudaStream_t st);

// ---- persistent decode-step megakernel (t3_mega.cu)
struct MegaLayer { const bf16 *wqkv_f, *wo_f, *wgu_f, *wd_f; const float *ln1, *ln2; };
struct MegaParams {
    const MegaLayer* layers = nullptr; int n_layers = 0;
    const bf16* head_f = nullptr; int head_items = 0; int vocab = 0; const float* final_norm = nullptr; float eps = 1e-5f;
    const float* x = nullptr;            // [rows_total][1024] step input (written by the sampler / slot init)
    float* logits = nullptr; long ld_logits = 0;
    // flagged-word exchange areas (8-byte words {payload, tag}), double-buffered by layer parity, indexed by compact row
    unsigned long long *ll_qkv = nullptr, *ll_ap = nullptr, *ll_y = nullptr, *ll_act = nullptr, *ll_z = nullptr;
    unsigned int* cnt = nullptr;         // arrival counters [layers][32] (zeroed before every step)
    unsigned int* epoch = nullptr;       // tag base of the next step (device counter, advanced by the kernel)
    bf16* kv = nullptr; long kv_layer_stride = 0, kv_half = 0;
    const int* page_table = nullptr; int max_pages = 0; const int* slot_pos = nullptr; const int* row_map = nullptr;
    const float* inv_freq = nullptr;
    int rows = 0;
    const unsigned long long* sched = nullptr; const int* sched_count = nullptr; int l2_ahead = 0;   // filled by launch_t3_mega
};
struct MegaState { unsigned long long* sched = nullptr; int* sched_count = nullptr; int grid = 0, max_pages = 0; };   // per engine
bool t3_mega_init(int max_pages, const std::vector<MegaLayer>& layers, const bf16* head_f, int head_items, MegaState* out);
void t3_mega_free(MegaState* ms);
size_t t3_mega_ll_words(int which);   // 0 qkv, 1 attention partials, 2 y, 3 act, 4 z, 5 arrival counters
void launch_t3_mega(const MegaParams& p, const MegaState& ms, cudaStream_t st);
