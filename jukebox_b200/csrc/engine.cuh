// Engine descriptor shared by the decode kernel (decode_engine.cu) and the chunked prefill (prefill.cu).
#pragma once
#include "common.cuh"
#include "../../include/jkb200.h"
#include <vector>

namespace jk {

struct LayerDev {
    int attn_func;
    int rows;                       // cache rows per (b, h)
    __half* kc;
    __half* vc;                     // [B][H][rows][dh_pad]
    const float *ln0_g, *ln0_b, *ln1_g, *ln1_b;
    const float *b_qkv, *b_o, *b_1, *b_2;   // fp32 holding fp16-rounded biases
    const __half* enc_w;            // [W][2S] fp16 copy of c_enc_kv.w (attn_func 6)
    const float* enc_b;
};

struct EngineDev {
    int W, S, M, H, dh, dh_pad, L, blocks, bc, bins, prime_pad, enc_dims, Bmax, add_cond_after, depth, G;
    int RC;                         // rows of one shared-memory K (or V) tile of the attention phase
    int ks_shift;                   // log2(KS)
    int KS, U;                      // K-split factor of every Conv1D and the number of column units (G = U * KS)
    int nslot, uni_bytes, kvpre_bytes, small_bytes, prof_on, kv_prefetch;
    float scale2;
    const ushort2* cols;            // [U][depth][4] : (first 8-column group, number of groups) of a unit
    const uint8_t* streams;
    unsigned long long stream_stride;
    // activations between phases travel as "LL" words (NCCL's low-latency protocol): 8 bytes = {half2 data, u32 flag},
    // written with one 8-byte store and polled by the consumer - no separate flag, no grid barrier
    unsigned long long *ll_h, *ll_x1, *ll_qkv, *ll_a, *ll_g;   // [16][N/2]
    unsigned long long* xp[4];      // K-split partial sums {fp32, flag}: [G][16][64] per Conv1D of a layer
    float* part;                    // split-KV partials [Bmax*H*kMaxSplit][dh_pad + 2]
    unsigned* acnt;                 // [Bmax*H] merge tickets
    long long* lnacc;               // [2*depth][16][2] fixed-point LayerNorm accumulators (sum, sumsq), 128-B apart
    long long* prof2;               // [kProfSlots][8] intra-phase clock64 stamps of CTA 0 (tuning aid)
    unsigned long long* prof3;      // [5][256][2] per-CTA phase entry / exit times of layer 1
    unsigned long long* prof;       // [kProfSlots] phase timestamps of CTA 0 (globaltimer ns)
    unsigned* sync;                 // [0] LN0 arrivals, [32] LN1 arrivals, [64] steps executed (one 128-B line each)
    int* t;
    const float *x_emb, *pos_emb, *x_out, *start_token;
    const int* lrow0;               // [G+1] logits rows per CTA (prefix)
    // logits as a fifth Conv1D on the tensor cores (decode_engine.cu "logits GEMM"): 0 when the configuration keeps
    // the fp32 FMA path; column groups per unit [U] and stream offsets per CTA [G] (16-byte units)
    int lg_on;
    const ushort2* lg_cols;
    const uint32_t* lg_goff;
    LayerDev layer[JK_MAX_DEPTH];
};

// prefill_gemm.cu: Y = epi(X . W^T + bias [, res]) on tcgen05; w_t is [N, K] fp16.
//   epi 0: fp16(acc + b)   1: fp16(quick_gelu(fp16(acc + b)))   2: fp16(res + fp16(acc + b))
int gemm_f16_tc(const void* x, const void* w_t, const float* bias, const void* res, void* y, int M, int N, int K,
                int epi, cudaStream_t stream);

}  // namespace jk

struct jk_prior {
    jk_prior_config cfg;
    jk::EngineDev host;            // host mirror of the device struct
    jk::EngineDev* dev;            // in arena
    uint8_t* arena;
    size_t arena_bytes;
    int G;
    int smem_bytes;
    int t_host;
    std::vector<ushort2> cols;       // [U][depth][4]
    std::vector<uint32_t> goff;      // [G][depth][4] per-GEMM stream offsets (16-B units)
    int lg_on;                       // logits GEMM planned (engine.cuh)
    uint32_t* d_goff;
    ushort2* d_cols;
    // arena sub-allocations for per-layer small params
    std::vector<float*> bias_ptr[4];
    std::vector<float*> ln_ptr[4];
    std::vector<__half*> enc_w;      // [2S][W] fp16, transposed copy of c_enc_kv.w
    std::vector<float*> enc_b;
    __half* enc_x16;                 // [max_batch*enc_dims][W] fp16 scratch
    __half* enc_y16;                 // [max_batch*enc_dims][2S] fp16 scratch
    // chunked prefill (prefill.cu): K-major fp16 copies of the four Conv1D weights of every layer and the
    // [pf_rows x .] activation workspace; pf_rows = 0 when the configuration cannot use the tensor-core path
    std::vector<__half*> wt[4];      // [N][K]
    int pf_rows, pf_len;             // workspace rows (= max_batch * pf_len), positions per prefill
    __half *pf_x, *pf_xn, *pf_qkv, *pf_a, *pf_x1, *pf_g;
};
