// Chunked prefill: all given (prime) positions of a window through every layer at once.
//
// Reference: ConditionalAutoregressive2D.primed_sample (prior/autoregressive.py:251-359) runs the given
// tokens through the transformer in chunks before sampling, and its own check (:330-338, check_chunks)
// asserts chunked == token-by-token.  The decode kernel (decode_engine.cu) is the token-by-token form; this
// file is the chunked form: M = n_samples x P rows per GEMM, so the four Conv1Ds of a layer run on the
// tcgen05 GEMM (prefill_gemm.cu) instead of streaming 1.8 GB of weights once per position.
//
// Per layer (rows m = b*P + p, fp16 activations, the decode kernel's rounding points):
//   xn  = LN0(x)                      ln_rows_kernel            (ops.py:14-24)
//   qkv = xn . Wqkv + b               gemm_f16_tc epi 0         (factored_attention.py:289-301)
//   a   = attention(q, K, V)          attn_fwd_kernel           (factored_attention.py:82-228, per pattern)
//   K, V -> the engine's caches       kv_scatter_kernel         (the layouts decode_engine.cu attends)
//   x1  = x + (a . Wo + b)            gemm_f16_tc epi 2         (transformer.py:82)
//   g   = quick_gelu(LN1(x1) . W1 + b)   ln_rows_kernel + gemm_f16_tc epi 1
//   x   = x1 + (g . W2 + b)           gemm_f16_tc epi 2         (transformer.py:83)
// Afterwards the engine stands at position P exactly as if P decode steps had run: only the K/V caches and
// the position carry over between steps.
#include "engine.cuh"
#include <algorithm>

using namespace jk;

namespace {

__device__ __forceinline__ float ldh(const __half* p) { return __half2float(*p); }

// ---- embedding (autoregressive.py:177-197; decode_engine.cu phase P0) ---------------------------------
__global__ void embed_rows_kernel(__half* __restrict__ x, const long long* __restrict__ tokens, long long tok_stride,
                                  const float* __restrict__ y_cond, const float* __restrict__ x_cond, long long x_cond_len,
                                  const float* __restrict__ x_emb, const float* __restrict__ pos_emb,
                                  const float* __restrict__ start_token, int n, int P, int W) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * P * W) return;
    const int col = (int)(i % W);
    const int m = (int)(i / W), b = m / P, t = m % P;
    float v;
    if (t == 0) v = y_cond ? y_cond[(size_t)b * W + col] : start_token[col];
    else v = x_emb[(size_t)tokens[(size_t)b * tok_stride + t - 1] * W + col];
    v += pos_emb[(size_t)t * W + col];
    if (x_cond) v += x_cond[((size_t)b * x_cond_len + (x_cond_len > 1 ? t : 0)) * W + col];
    x[i] = __float2half_rn(v);
}

// ---- LayerNorm of fp16 rows (fp32 math, eps 1e-5), one warp per row ------------------------------------
// same formulas as the decode kernel's staging: mean, var = E[x^2] - mean^2 (double for the cancellation),
// y = fp16(fma(fma(x, rstd, -mean*rstd), g, b))
__global__ void ln_rows_kernel(const __half* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                               __half* __restrict__ y, int rows, int W) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const __half* xr = x + (size_t)row * W;
    double s1 = 0.0, s2 = 0.0;
    for (int c = lane * 8; c < W; c += 256) {
        const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
        const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float f = __half2float(h[e]); s1 += (double)f; s2 += (double)f * (double)f; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const double rk = (double)(1.0f / (float)W);
    const double m = s1 * rk;
    double var = s2 * rk - m * m;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = 1.0f / sqrtf((float)var + 1e-5f);
    const float nmr = -(float)m * rstd;
    for (int c = lane * 8; c < W; c += 256) {
        const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
        const __half* h = reinterpret_cast<const __half*>(&v);
        __half o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
            o[e] = __float2half_rn(fmaf(fmaf(__half2float(h[e]), rstd, nmr), g[c + e], bta[c + e]));
        *reinterpret_cast<uint4*>(y + (size_t)row * W + c) = *reinterpret_cast<const uint4*>(o);
    }
}

// ---- attention, forward mode over the P given positions ------------------------------------------------
// One CTA per (position p, head h, sample b).  Keys of p by pattern (all inside [0, P)):
//   0 dense: 0..p     1 block: block start..p     2 transpose: p % bc + j*bc, j = 0..p/bc
//   3 previous block: (p/bc - 1)*bc .. +bc-1 (none in the first block -> output 0)     7 prime: 0..p (p < prime)
//   6 encoder-decoder: every encoder row; K / V come from the layer's cache (jk_prior_set_encoder_kv), q from c_attn
// Scores fp16(fp16(q.k) * dh^-1/2), softmax fp32, P rounded to fp16 (unnormalised), P.V fp32, / sum - the
// decode kernel's order of roundings.
struct AttnFwd {
    const __half* qkv;   // [n*P][q_stride]: q | k | v per row (q only for an encoder-decoder layer)
    __half* a;           // [n*P][S]
    const __half *kc, *vc;   // encoder-decoder layers: the layer's K / V cache [n][H][enc_rows][dhp]
    int P, S, H, dh, bc, attn_func, prime, q_stride, enc_rows, dhp;
    float scale2;
};

__device__ __forceinline__ int fwd_nkeys(const AttnFwd& A, int p) {
    switch (A.attn_func) {
        case 0: return p + 1;
        case 1: return p % A.bc + 1;
        case 2: return p / A.bc + 1;
        case 3: return p >= A.bc ? A.bc : 0;
        case 6: return A.enc_rows;
        case 7: return p < A.prime ? p + 1 : A.prime;
    }
    return 0;
}
__device__ __forceinline__ int fwd_key(const AttnFwd& A, int p, int j) {
    switch (A.attn_func) {
        case 1: return p - p % A.bc + j;
        case 2: return p % A.bc + j * A.bc;
        case 3: return (p / A.bc - 1) * A.bc + j;
    }
    return j;   // 0, 7
}

constexpr int kFwdThreads = 128;

__global__ void __launch_bounds__(kFwdThreads) attn_fwd_kernel(AttnFwd A) {
    extern __shared__ float fsm[];
    float* qs = fsm;                 // [dh]
    float* sc = fsm + A.dh;          // [nk]
    __shared__ float red[kFwdThreads / 32];
    const int p = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dh = A.dh, S = A.S;
    const size_t row = (size_t)b * A.P + p;
    const __half* q = A.qkv + row * A.q_stride + h * dh;
    __half* out = A.a + row * S + h * dh;
    const bool enc = A.attn_func == 6;
    const __half* kbase = enc ? A.kc + ((size_t)b * A.H + h) * A.enc_rows * A.dhp : A.qkv + (size_t)b * A.P * 3 * S + S + h * dh;
    const __half* vbase = enc ? A.vc + ((size_t)b * A.H + h) * A.enc_rows * A.dhp : kbase + S;
    const size_t kstride = enc ? (size_t)A.dhp : (size_t)3 * S;
    const int nk = fwd_nkeys(A, p);
    if (nk == 0) {
        for (int d = tid; d < dh; d += kFwdThreads) out[d] = __float2half_rn(0.f);
        return;
    }
    for (int d = tid; d < dh; d += kFwdThreads) qs[d] = ldh(q + d);
    __syncthreads();
    for (int j = warp; j < nk; j += kFwdThreads / 32) {
        const __half* k = kbase + (size_t)fwd_key(A, p, j) * kstride;
        float dot = 0.f;
        for (int d = lane; d < dh; d += 32) dot = fmaf(qs[d], ldh(k + d), dot);
        dot = warp_sum(dot);
        if (lane == 0) sc[j] = h2f_round(h2f_round(dot) * A.scale2);
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = tid; j < nk; j += kFwdThreads) mx = fmaxf(mx, sc[j]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float l = 0.f;
    for (int j = tid; j < nk; j += kFwdThreads) {
        const float e = expf(sc[j] - mx);
        l += e;
        sc[j] = h2f_round(e);
    }
    l = warp_sum(l);
    if (lane == 0) red[warp] = l;
    __syncthreads();
    const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
    for (int d = tid; d < dh; d += kFwdThreads) {
        float o = 0.f;
        for (int j = 0; j < nk; ++j) {
            o = fmaf(sc[j], ldh(vbase + (size_t)fwd_key(A, p, j) * kstride + d), o);
        }
        out[d] = __float2half_rn(o * inv);
    }
}

// ---- K, V of the given positions -> the caches the decode kernel attends -------------------------------
// cache row of position p (decode_engine.cu attn_geom.wrow); ring layouts keep only the last writer.
__global__ void kv_scatter_kernel(const __half* __restrict__ qkv, __half* __restrict__ kc, __half* __restrict__ vc, int n,
                                  int P, int S, int H, int dh, int dhp, int rows, int attn_func, int bc, int blocks,
                                  int prime) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * P * S) return;
    const int cs = (int)(i % S);
    const int m = (int)(i / S), b = m / P, p = m % P;
    const int h = cs / dh, d = cs % dh;
    int wrow = -1;
    switch (attn_func) {
        case 0: wrow = p; break;
        case 1: wrow = (p + bc >= P) ? p % bc : -1; break;
        case 2: wrow = (p % bc) * blocks + p / bc; break;
        case 3: wrow = (p + 2 * bc >= P) ? ((p / bc) & 1) * bc + p % bc : -1; break;
        case 7: wrow = (p < prime) ? p : -1; break;
    }
    if (wrow < 0) return;
    const size_t dst = (((size_t)b * H + h) * rows + wrow) * dhp + d;
    const __half* src = qkv + (size_t)m * 3 * S + cs;
    kc[dst] = src[S];
    vc[dst] = src[2 * S];
}

__global__ void rows_to_float_kernel(const __half* __restrict__ x, float* __restrict__ y, size_t cnt) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) y[i] = __half2float(x[i]);
}

__global__ void set_position_kernel(int* t, int v) { *t = v; }

}  // namespace

extern "C" int jk_prior_prefill_capacity(const jk_prior* p, int* max_positions) {
    JK_REQUIRE(p && max_positions, "null argument");
    *max_positions = p->pf_len;
    return 0;
}

extern "C" int jk_prior_prefill(jk_prior* p, const jk_prefill_args* a, jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(p && a, "null argument");
    const jk_prior_config& c = p->cfg;
    const EngineDev& E = p->host;
    JK_REQUIRE(p->pf_len > 0, "this configuration has no chunked prefill (needs width, n_state, mlp_width >= 64 and %% 8 == 0): "
                              "step the given tokens through jk_prior_step");
    JK_REQUIRE(p->t_host == 0, "prefill starts at position 0 (engine is at %d)", p->t_host);
    const int n = a->n_samples, P = a->n_positions;
    JK_REQUIRE(n >= 1 && n <= c.max_batch, "n_samples %d out of range (max_batch %d)", n, c.max_batch);
    JK_REQUIRE(P >= 1 && P <= p->pf_len && P <= c.n_ctx, "n_positions %d out of range (capacity %d)", P, p->pf_len);
    JK_REQUIRE(P == 1 || a->tokens, "tokens required");
    JK_REQUIRE(E.pos_emb && E.x_emb, "embeddings not set (jk_prior_set_embeddings)");
    JK_REQUIRE(a->x_cond_len == 0 || a->x_cond_len == 1 || a->x_cond_len == c.n_ctx, "x_cond_len must be 1 or n_ctx");
    const int W = c.width, S = c.n_state, Mw = c.mlp_width, H = c.heads;
    const int rows = n * P;
    {
        const size_t cnt = (size_t)rows * W;
        embed_rows_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(
            p->pf_x, (const long long*)a->tokens, a->tok_stride, a->y_cond, a->x_cond, a->x_cond_len ? a->x_cond_len : 1,
            E.x_emb, E.pos_emb, E.start_token, n, P, W);
        JK_CHECK_CUDA(cudaGetLastError());
    }
    const unsigned ln_grid = (unsigned)((rows + 7) / 8);
    static bool attr_set[64] = {};         // per device: the attribute belongs to the device's copy of the function
    const size_t fwd_smem = (size_t)(E.dh + std::max(P, E.enc_dims)) * 4;
    int dev = 0;
    JK_CHECK_CUDA(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        JK_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set[dev & 63] = true;
    }
    JK_REQUIRE(fwd_smem <= 64 * 1024, "prefill attention tile too large");
    for (int l = 0; l < c.depth; ++l) {
        const LayerDev& LD = E.layer[l];
        ln_rows_kernel<<<ln_grid, 256, 0, stream>>>(p->pf_x, LD.ln0_g, LD.ln0_b, p->pf_xn, rows, W);
        JK_CHECK_CUDA(cudaGetLastError());
        const bool enc = LD.attn_func == 6;       // c_attn gives q only; K / V are the encoder's (already in the cache)
        const int q_stride = enc ? S : 3 * S;
        int rc = gemm_f16_tc(p->pf_xn, p->wt[0][l], LD.b_qkv, nullptr, p->pf_qkv, rows, q_stride, W, 0, stream);
        if (rc) return rc;
        AttnFwd A;
        A.qkv = p->pf_qkv; A.a = p->pf_a; A.P = P; A.S = S; A.H = H; A.dh = E.dh; A.bc = E.bc; A.attn_func = LD.attn_func;
        A.prime = E.prime_pad; A.scale2 = E.scale2; A.q_stride = q_stride; A.kc = LD.kc; A.vc = LD.vc; A.enc_rows = E.enc_dims;
        A.dhp = E.dh_pad;
        attn_fwd_kernel<<<dim3(P, H, n), kFwdThreads, fwd_smem, stream>>>(A);
        JK_CHECK_CUDA(cudaGetLastError());
        if (!enc) {
            const size_t cnt = (size_t)rows * S;
            kv_scatter_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(p->pf_qkv, LD.kc, LD.vc, n, P, S, H, E.dh, E.dh_pad,
                                                                                  LD.rows, LD.attn_func, E.bc, E.blocks, E.prime_pad);
            JK_CHECK_CUDA(cudaGetLastError());
        }
        rc = gemm_f16_tc(p->pf_a, p->wt[1][l], LD.b_o, p->pf_x, p->pf_x1, rows, W, S, 2, stream);
        if (rc) return rc;
        ln_rows_kernel<<<ln_grid, 256, 0, stream>>>(p->pf_x1, LD.ln1_g, LD.ln1_b, p->pf_xn, rows, W);
        JK_CHECK_CUDA(cudaGetLastError());
        rc = gemm_f16_tc(p->pf_xn, p->wt[2][l], LD.b_1, nullptr, p->pf_g, rows, Mw, W, 1, stream);
        if (rc) return rc;
        rc = gemm_f16_tc(p->pf_g, p->wt[3][l], LD.b_2, p->pf_x1, p->pf_x, rows, W, Mw, 2, stream);
        if (rc) return rc;
    }
    if (a->h_out) {
        const size_t cnt = (size_t)rows * W;
        rows_to_float_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(p->pf_x, a->h_out, cnt);
        JK_CHECK_CUDA(cudaGetLastError());
    }
    set_position_kernel<<<1, 1, 0, stream>>>(E.t, P);
    JK_CHECK_CUDA(cudaGetLastError());
    p->t_host = P;
    return 0;
}
