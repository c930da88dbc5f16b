/*
 * jkb200.h - C ABI of libjkb200.so: B200 (sm_100a) kernels for Jukebox's sampling hot path.
 *
 * Boundary contract:
 *   - plain C: pointers, sizes, cudaStream_t (passed as void*); no torch types
 *   - every call enqueues on the given stream and returns; nothing synchronises, nothing
 *     allocates or frees caller memory.  Engines live inside a caller-provided arena whose
 *     size is reported by the *_arena_bytes call.
 *   - return value 0 = ok, negative = error; jk_last_error() gives the message
 *     (thread-local).  The Python host turns a non-zero code into RuntimeError, the same
 *     way the reference's optional native plug-ins surface C++ exceptions
 *     (apex/csrc/layer_norm_cuda.cpp:138-159 -> RuntimeError in jukebox/transformer/ops.py:8-24).
 *
 * Each entry point cites the reference interface it replaces (paths under /root/reference/jukebox).
 */
#ifndef JKB200_H
#define JKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JK_MAX_DEPTH 96
#define JK_MAX_BATCH 16

typedef void* jk_stream_t;          /* cudaStream_t */

const char* jk_last_error(void);
int jk_version(void);
/* number of SMs of the current device (grid size of the persistent decode kernel) */
int jk_device_sm_count(int* out);

/* ------------------------------------------------------------------------------------------
 * Autoregressive prior decode engine.
 *
 * Replaces the per-token body of ConditionalAutoregressive2D.sample / primed_sample
 * (prior/autoregressive.py:199-359): get_emb (:177-197) -> Transformer.forward(sample=True)
 * (transformer/transformer.py:169-192) -> ResAttnBlock sample branch (:62-65,82-86) ->
 * LayerNorm (transformer/ops.py:14-24), Conv1D (ops.py:83-101), FactoredAttention.forward with
 * its KV cache (transformer/factored_attention.py:230-301, 328-373), MLP + quick_gelu
 * (transformer.py:19-30, ops.py:33-35) -> +cond -> x_out logits (autoregressive.py:226-229).
 *
 * One jk_prior_step call = one token position for up to 16 samples, executed by ONE persistent
 * kernel (grid = #SMs) that streams every layer's fp16 weights once through a TMA bulk-copy
 * ring in shared memory.  The position is read from device memory, so the call sequence is
 * CUDA-graph capturable.
 * ---------------------------------------------------------------------------------------- */
typedef struct jk_prior_config {
    int32_t width;            /* prior_width                                   */
    int32_t depth;            /* prior_depth                                   */
    int32_t heads;
    int32_t n_state;          /* int(m_attn * width)  (transformer.py:42)      */
    int32_t mlp_width;        /* int(m_mlp * width)                            */
    int32_t n_ctx;            /* full input_dims of the CA2D (incl. lyric tokens for single_enc_dec) */
    int32_t blocks;           /* 0 if none (dense only)                        */
    int32_t bins;             /* rows of x_out; 0 when only_encode             */
    int32_t prime_len;        /* raw prime_len for attn_func 7, else 0         */
    int32_t encoder_dims;     /* rows of the encoder K/V for attn_func 6       */
    int32_t max_batch;        /* <= JK_MAX_BATCH                                */
    int32_t add_cond_after;   /* 1 unless merged_decoder (autoregressive.py:87-93) */
    int32_t attn_func[JK_MAX_DEPTH];  /* per layer: 0,1,2,3,6,7 (transformer.py:110-124) */
} jk_prior_config;

/* Reference-layout weights of one ResAttnBlock, device pointers.  *_w are Conv1D.w
 * [n_in, n_out] row-major (ops.py:89-95) in fp32 (w_dtype 0) or fp16 (w_dtype 1, fp16_params);
 * biases and LayerNorm parameters are fp32 (biases may be fp16 when b_dtype is 1). */
typedef struct jk_layer_weights {
    const void* c_attn_w;  const void* c_attn_b;      /* [W, 3S] ([W, S] for attn_func 6)  */
    const void* c_enc_kv_w; const void* c_enc_kv_b;   /* [W, 2S], attn_func 6 only, else NULL */
    const void* c_proj_w;  const void* c_proj_b;      /* [S, W]                              */
    const void* fc_w;      const void* fc_b;          /* [W, M]                              */
    const void* proj2_w;   const void* proj2_b;       /* [M, W]                              */
    const float* ln0_g; const float* ln0_b;           /* [W]                                 */
    const float* ln1_g; const float* ln1_b;           /* [W]                                 */
    int32_t w_dtype;       /* 0 = fp32, 1 = fp16 */
    int32_t b_dtype;       /* 0 = fp32, 1 = fp16 */
} jk_layer_weights;

typedef struct jk_prior jk_prior;   /* opaque; lives in the caller's arena */

/* How the engine would lay a configuration out on a device with `n_sms` SMs - pure host arithmetic, no device needed
 * (the CPU tests use it; the engine itself always plans for the current device).  cols (optional, uint16 pairs
 * [units][depth][4][2]) receives (first 8-column group, number of groups) of every unit for the four Conv1Ds of a layer. */
typedef struct jk_prior_plan_info {
    int32_t k_split;          /* CTAs per unit: they share the unit's columns and split K                 */
    int32_t units;            /* n_sms / k_split                                                          */
    int32_t ring_slots;       /* 16 KB weight-ring slots per SM                                           */
    int32_t smem_bytes;       /* dynamic shared memory of the decode kernel                               */
    int32_t tile_rows;        /* K/V rows per attention tile                                              */
    uint64_t arena_bytes;
    uint64_t stream_stride;   /* bytes of the longest per-SM weight stream (+ padding)                    */
} jk_prior_plan_info;
int jk_prior_plan(const jk_prior_config* cfg, int n_sms, jk_prior_plan_info* out, uint16_t* cols, size_t cols_len);

/* bytes of device memory the engine needs for packed weights, KV caches and activations */
int jk_prior_arena_bytes(const jk_prior_config* cfg, size_t* bytes);
/* `arena` is device memory (256-B aligned) of at least jk_prior_arena_bytes; it is zeroed here */
int jk_prior_create(const jk_prior_config* cfg, void* arena, size_t arena_bytes,
                    jk_prior** out, jk_stream_t stream);
int jk_prior_destroy(jk_prior* p);
/* pack one layer's weights into the per-SM stream layout (device -> device) */
int jk_prior_load_layer(jk_prior* p, int layer, const jk_layer_weights* w, jk_stream_t stream);
/* embeddings are used in place (fp32, reference layout): x_emb [bins_in, W], pos_emb [n_ctx, W],
 * x_out [bins, W] (= x_emb when tied), start_token [W] or NULL */
int jk_prior_set_embeddings(jk_prior* p, const float* x_emb, const float* pos_emb,
                            const float* x_out, const float* start_token);
/* position <- t0 (usually 0); KV caches are logically emptied (FactoredAttention.del_cache,
 * factored_attention.py:375-381) */
int jk_prior_reset(jk_prior* p, int t0, jk_stream_t stream);
/* encoder K/V for attn_func 6 layers: c_enc_kv(encoder_kv) computed once per window
 * (factored_attention.py:273-287).  encoder_kv: fp32 [n, encoder_dims, W] */
/* 1 if this engine multiplies the logits on the tensor cores (hi / lo fp16 split of x_out; needs an even K split and
 * bins > 0), i.e. if jk_step_args.logit_bias is worth computing */
int jk_prior_has_logits_gemm(const jk_prior* p, int* on);

int jk_prior_set_encoder_kv(jk_prior* p, const float* encoder_kv, int n_samples, jk_stream_t stream);

typedef struct jk_step_args {
    int32_t n_samples;            /* <= max_batch */
    /* input: either an embedded activation (Transformer.forward boundary) ... */
    const float* x_in;            /* fp32 [n, W] or NULL */
    /* ... or tokens (CA2D.sample boundary): token fed at position t is tokens[b*tok_stride + t-1] */
    const int64_t* tokens;        /* int64, or NULL */
    int64_t tok_stride;
    const float* y_cond;          /* fp32 [n, W]: input at t == 0 when the prior is y-conditioned, else NULL -> start_token */
    const float* x_cond;          /* fp32 [n, x_cond_len, W] or NULL (treated as zeros) */
    int64_t x_cond_len;           /* 1 or n_ctx */
    /* outputs (any may be NULL) */
    float* h_out;                 /* fp32 [n, W]: Transformer.forward output (before +cond) */
    float* logits;                /* fp32: logits[b*logits_bstride + t*logits_tstride + v] */
    int64_t logits_bstride;
    int64_t logits_tstride;       /* 0 to overwrite the same [n, bins] buffer every step */
    /* optional: x_cond . x_out^T of every position, computed once per window by the caller (the logits are linear in the
     * activation: (h + x_cond) . x_out^T = h . x_out^T + logit_bias).  With it, priors that add x_cond behind the stack
     * (autoregressive.py:226-227: every label-conditioned prior) still take the tensor-core logits product, whose
     * activation operand must be an fp16 value; NULL keeps the fp32 product of h + x_cond.
     * logit_bias[b*logit_bias_bstride + t*logit_bias_tstride + v]; ignored when the engine has no logits GEMM
     * (jk_prior_has_logits_gemm). */
    const float* logit_bias;
    int64_t logit_bias_bstride;
    int64_t logit_bias_tstride;
} jk_step_args;

/* Chunked prefill of the given (prime) tokens: positions 0 .. n_positions-1 of every sample through all
 * layers in one call - the chunked half of ConditionalAutoregressive2D.primed_sample
 * (prior/autoregressive.py:251-359), whose own check_chunks asserts it equals stepping token by token.
 * The four Conv1Ds of each layer run as [n_samples * n_positions, K] x [K, N] GEMMs on tcgen05.  On return
 * the engine stands at position n_positions (K/V caches filled), exactly as after that many jk_prior_step
 * calls.  tokens[b * tok_stride + t] is the token AT position t (the input of position t+1), as in
 * jk_step_args; h_out (optional, fp32 [n_samples, n_positions, width]) receives the transformer output -
 * the `only_encode` forward of the lyric encoder (prior/prior.py:285-301). */
typedef struct jk_prefill_args {
    int32_t n_samples;
    int32_t n_positions;
    const int64_t* tokens;
    int64_t tok_stride;
    const float* y_cond;      /* [n_samples, width] or NULL (start token) */
    const float* x_cond;      /* [n_samples, x_cond_len, width] or NULL */
    int64_t x_cond_len;       /* 1 or n_ctx */
    float* h_out;
} jk_prefill_args;
/* positions one prefill call can take; 0 when the configuration has no tensor-core prefill (a GEMM K that
 * is not a multiple of 64, or encoder-decoder layers): step the given tokens instead */
int jk_prior_prefill_capacity(const jk_prior* p, int* max_positions);
int jk_prior_prefill(jk_prior* p, const jk_prefill_args* args, jk_stream_t stream);

/* one token position; increments the device-side position counter */
int jk_prior_step(jk_prior* p, const jk_step_args* a, jk_stream_t stream);
/* current position (host copy of the device counter as tracked by the calls made so far) */
int jk_prior_position(const jk_prior* p, int* t);
/* debug / test access to the fp16 intermediates of the LAST layer executed:
 * which: 0 = h, 1 = qkv, 2 = attention out, 3 = x1 (x + a), 4 = gelu out.  Returns device ptr. */
int jk_prior_debug_buffer(const jk_prior* p, int which, const void** ptr, size_t* n_halfs);

/* Conv1D at prefill / training shape on the tensor cores (tcgen05 + TMA): y[M, N] = x[M, K] . w + b, fp16 in,
 * fp32 accumulate, fp16 out (transformer/ops.py:83-101).  w_t is the weight TRANSPOSED: [N, K] row-major fp16;
 * bias fp32 [N] or NULL; K must be a multiple of 64.  Used for c_enc_kv(encoder_kv)
 * (factored_attention.py:273-287) inside jk_prior_set_encoder_kv. */
int jk_conv1d_prefill_f16(const void* x, const void* w_t, const float* bias, void* y, int M, int N, int K,
                          jk_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * VQ-VAE.  Tensors are channels-last: [N, T, C] fp32.
 * ---------------------------------------------------------------------------------------- */
/* BottleneckBlock.quantise (vqvae/bottleneck.py:112-119): idx = argmin_j |x|^2 - 2 x.k_j + |k_j|^2
 * in fp32, lowest index on ties.  x [n, width], codebook [k_bins, width], idx int64 [n]. */
int jk_vq_argmin(const float* x, const float* codebook, int64_t* idx, float* min_dist /*nullable*/,
                 int64_t n, int k_bins, int width, jk_stream_t stream);
/* BottleneckBlock.dequantise (bottleneck.py:121-123): out[n, :] = codebook[idx[n], :] */
int jk_vq_gather(const int64_t* idx, const float* codebook, float* out, int64_t n, int k_bins,
                 int width, jk_stream_t stream);

/* Generic channels-last 1-D convolution used for every conv of Encoder/Decoder/Resnet1D
 * (vqvae/encdec.py:6-131, vqvae/resnet.py:27-44) and the upsampler Conditioner
 * (prior/conditioners.py:8-48):
 *   out[n, t, co] = (res ? res[n, t, co] : 0)
 *                 + scale * ( bias[co] + sum_tap sum_ci w[tap, ci, co] * pre(in[n, t*in_stride + tap_off[tap], ci]) )
 * pre = ReLU when relu_in.  Out-of-range input positions read as zero (conv padding).
 * w is packed [n_taps, c_in, c_out] (see jk_pack_conv_weight).  For a transposed conv the caller
 * issues one call per output phase with out_stride 2 / out_offset r. */
typedef struct jk_conv_args {
    const float* in;  int64_t t_in;  int32_t c_in;
    float* out;       int64_t t_out; int32_t c_out;     /* t_out counts positions written by THIS call */
    const float* w;   const float* bias;
    const float* res;                 /* nullable; indexed like out */
    int32_t n_taps;   int32_t tap_off[4];
    int32_t in_stride;                /* input positions per output position (1, or 2 for the strided encoder conv) */
    int32_t out_stride; int32_t out_offset;  /* output row = t*out_stride + out_offset (within a buffer of t_out*out_stride rows) */
    int32_t relu_in;
    float scale;
    int32_t n;                        /* batch */
    int32_t tensor_cores;             /* 0: exact fp32 FMAs in a fixed order (the encoder: its output feeds the bit-exact argmin);
                                         1: decoder side - c_in, c_out in {32, 64} run on mma.sync with the fp16 x 3 split
                                         (fp32-level accuracy, free summation order); other shapes ignore the flag */
} jk_conv_args;
int jk_conv1d_cl(const jk_conv_args* a, jk_stream_t stream);

/* ResConv1DBlock (vqvae/resnet.py:27-44): out = x + res_scale * (W2.relu(W1 *_dil relu(x) + b1) + b2)
 * x, out [n, T, C] (x != out); w1 packed [3, C, Cs]; w2 packed [1, Cs, C].  For C == Cs in {32, 64} (every
 * ResConv1DBlock of the reference's VQ-VAEs) this is ONE launch with the hidden activation kept in shared memory and
 * tmp may be NULL; other shapes run as two jk_conv1d_cl launches through tmp [n, T, Cs]. */
int jk_resblock_cl(const float* x, float* out, float* tmp, const float* w1, const float* b1, const float* w2,
                   const float* b2, int n, int64_t T, int C, int Cs, int dilation, float res_scale,
                   jk_stream_t stream);

/* The same ResConv1DBlock on the tensor cores for the DECODER side (Decoder stacks of vqvae/encdec.py:87-131 and the
 * upsampler Conditioner, prior/conditioners.py:8-48), C == Cs in {32, 64}: 3xTF32 split (hi/lo operands, fp32 accumulate in
 * mma.sync m16n8k8), i.e. fp32 accuracy up to the 2^-22 lo.lo term but NOT the FMA order of jk_resblock_cl.  The encoder,
 * whose output feeds the bit-exact codebook argmin, must keep jk_resblock_cl. */
int jk_resblock_tc(const float* x, float* out, const float* w1, const float* b1, const float* w2, const float* b2,
                   int n, int64_t T, int C, int dilation, float res_scale, jk_stream_t stream);

/* Token sampling of the autoregressive loop (prior/autoregressive.py:233-235, 343-345):
 *   tokens[r, position] ~ Categorical(logits = logits[r, :] / temp),  r = 0..n-1
 * one launch per position.  logits: fp32 rows `logits_stride` floats apart (entries of -inf carry no
 * mass, so rows already passed through filter_logits are valid input); tokens: int64 [n, tok_stride].
 * The uniform behind row r at `position` is Philox4x32-10(key = seed, counter = (position, r)), so a
 * (seed, position, row) triple always draws the same token for the same logits. */
int jk_sample_categorical(const float* logits, int64_t logits_stride, int n, int bins, float temp,
                          uint64_t seed, int position, int64_t* tokens, int64_t tok_stride,
                          jk_stream_t stream);

/* top-k / nucleus filtering in front of the sampler (transformer/ops.py:113-142 `filter_logits`, applied to
 * logits / temp as autoregressive.py:232-234 does): out[r, v] = logits[r, v] / temp if v stays, else -inf.
 * top_k > 0: the k largest stay; top_p > 0: the smallest prefix of the sorted row whose softmax mass exceeds top_p
 * stays (the entry that crosses the threshold included).  Exactly one of the two must be set; bins <= 4096.
 * Ties at the cut stay together (the reference's scatter-by-index keeps an arbitrary subset of equal values). */
int jk_filter_logits(const float* logits, int64_t logits_stride, int n, int bins, float temp, int top_k,
                     float top_p, float* out, int64_t out_stride, jk_stream_t stream);

/* torch Conv1d weight [c_out, c_in, k] (transposed = 0) or ConvTranspose1d weight
 * [c_in, c_out, k] (transposed = 1) -> packed [k, c_in, c_out] */
int jk_pack_conv_weight(const float* w, float* packed, int c_out, int c_in, int k, int transposed,
                        jk_stream_t stream);

/* LayerNorm over the last dim of fp32 rows (Conditioner.ln, prior/conditioners.py:47;
 * transformer/ops.py:14-24): y = (x-mu)/sqrt(var+eps)*g + b */
int jk_layernorm_f32(const float* x, const float* g, const float* b, float* y, int64_t rows, int width,
                     float eps, jk_stream_t stream);
/* out[n, :] = table[idx[n], :] (+ add[n, :] if add) - embedding lookups of Conditioner /
 * LabelConditioner (prior/conditioners.py:40-43, 57-68) */
int jk_embedding_f32(const int64_t* idx, const float* table, const float* add, float* out, int64_t n,
                     int rows, int width, jk_stream_t stream);

/* ---- fp32 transformer path (not the hot path: exactness against the reference's fp32 outputs) ------------------------
 * Transformer.forward in fp32 (transformer/transformer.py:169-192): sample = forward mode over a whole sequence
 * (training-shaped calls, alignment with record_attn - prior/prior.py:327-344) and sampling with fp32 K/V caches
 * (ConditionalAutoregressive2D.sample(fp16=False), prior/autoregressive.py:199-249).  One structure per layer holds the
 * reference's parameter tensors (fp32, Conv1D layout [n_in, n_out], transformer/ops.py:83-101) and the layer's caches. */
typedef struct jk_f32_layer {
    const float *ln0_g, *ln0_b, *ln1_g, *ln1_b;
    const float *c_attn_w, *c_attn_b;       /* [width, 3 n_state]; [width, n_state] (c_attn of an enc-dec layer) */
    const float *c_enc_kv_w, *c_enc_kv_b;   /* [width, 2 n_state], attn_func 6 only */
    const float *c_proj_w, *c_proj_b;       /* [n_state, width] */
    const float *fc_w, *fc_b;               /* [width, mlp_width] */
    const float *proj2_w, *proj2_b;         /* [mlp_width, width] */
    float *k_cache, *v_cache;               /* [n, n_ctx, n_state] ([n, encoder_dims, n_state] for attn_func 6); rows are
                                               absolute positions.  Forward mode passes scratch of the same shape. */
    float *attn_w;                          /* optional: attention weights [n, heads, P, n_ctx | encoder_dims] of this call
                                               (record_attn, factored_attention.py:100-102), or NULL */
    int32_t attn_func;                      /* 0, 1, 2, 3, 6, 7 (factored_attention.py:49-58) */
} jk_f32_layer;

typedef struct jk_f32_args {
    int32_t n, P, p0;              /* samples, positions in this call, absolute position of the first one */
    int32_t width, n_state, mlp_width, heads, n_ctx, blocks, prime_len, encoder_dims, depth;
    float* x;                      /* [n, P, width]: residual stream, transformed in place */
    const float* encoder_kv;       /* [n, encoder_dims, width] or NULL (read when p0 == 0) */
    float* work;                   /* jk_f32_workspace_floats() floats */
} jk_f32_args;

int jk_f32_workspace_floats(const jk_f32_args* a, size_t* floats);
/* Positions [p0, p0 + P) of every sample through all `depth` layers.  K / V of the new positions are written to the
 * caches first, each query then attends the cache rows of its pattern; a call with p0 = 0, P = n_ctx is the reference's
 * forward mode, P = 1 its per-token sampling step. */
int jk_f32_forward(const jk_f32_args* a, const jk_f32_layer* layers, jk_stream_t stream);
/* x[b, i, :] = (position p0+i == 0 ? y_cond[b] or start_token : x_emb[tokens[b, p0+i-1]]) + pos_emb[p0+i] (+ x_cond)
 * (prior/autoregressive.py:115-123 shifted input in forward mode, :176-191 get_emb in sampling). x_cond_len: 0 = none,
 * 1 = one broadcast row per sample, else rows indexed by position. */
int jk_f32_embed(float* x, const int64_t* tokens, int64_t tok_stride, const float* y_cond, const float* x_cond,
                 int64_t x_cond_len, const float* x_emb, const float* pos_emb, const float* start_token, int n, int P,
                 int p0, int width, jk_stream_t stream);
/* y[M, N] = x[M, K] . w + b;  w is [K, N] (Conv1D) or, with w_is_nk, [N, K] (nn.Linear: x_out, autoregressive.py:86) */
int jk_f32_linear(const float* x, const float* w, const float* b, float* y, int M, int N, int K, int w_is_nk,
                  jk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* JKB200_H */
