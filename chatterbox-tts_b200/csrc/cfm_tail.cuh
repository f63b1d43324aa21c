// Fused tail of a CFM transformer block (cfm_tail.cu): out projection + residual, LayerNorm, GELU feed-forward + residual,
// next block's LayerNorm + QKV projection -- one kernel per 128-row tile.
#pragma once
#include "common.cuh"

enum { CFM_TAIL_OUT = 1, CFM_TAIL_FF = 2, CFM_TAIL_QKV = 4 };

struct CfmTailArgs {
    int M = 0, mode = 0;
    float* h = nullptr;                                      // [M][256] fp32 residual stream, updated in place
    const float* b_out = nullptr;                            // [256]
    const float *ln3_g = nullptr, *ln3_b = nullptr;          // [256]
    const float *b0 = nullptr, *b2 = nullptr;                // [1024], [256]
    const float *ln1_g = nullptr, *ln1_b = nullptr;          // [256] LayerNorm1 of the NEXT block
    bf16* qkv = nullptr;                                     // [M][1536] bf16 out
    // ragged batches: rows are [slab][seq_T] (two slabs per call); a tile that lies wholly in a slab's padding (row index inside
    // the slab >= seq_len[slab / 2]) is skipped -- its rows are never consumed (causal convs, masked attention keys).  0 = off
    int seq_T = 0; int seq_len[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

// TMA descriptors of one block's weights (static for the life of the engine): 128-byte opaque CUtensorMaps
struct alignas(64) CfmTailWeights {
    unsigned char out[128], w0[128], w2[128], qkv[128];
    bool has_out = false, has_ff = false, has_qkv = false;
};

void cfm_tail_init();
bool cfm_tail_available();
void cfm_tail_weights(CfmTailWeights& w, const bf16* wout /*[256][512]*/, const bf16* w0 /*[1024][256]*/, const bf16* w2 /*[256][1024]*/,
                      const bf16* wqkv /*[1536][256]*/);
void launch_cfm_tail(const CfmTailArgs& a, const bf16* attn_o /*[M][512]*/, const CfmTailWeights* blk, const CfmTailWeights* nxt, cudaStream_t st);
