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

// ---- the same attention on the tensor cores (mma.sync.m16n8k16), flash style --------------------------------
// Every pattern is a set of independent SEQUENCES: queries q_pos(i) = q0 + i*qs (i < nq) attend keys k_pos(j) = k0 + j*ks
// (j < nk) with k_pos <= q_pos:
//   block: one sequence per block (q = k = the block's rows)       transpose: one per residue r (rows r, r+bc, ...)
//   previous block: q = block s, k = block s-1 (none for s = 0)    dense: one sequence of all rows
//   prime: q = all rows, k = the first min(prime, P) rows          encoder-decoder: q = all rows, k = every cache row (no mask)
// so that QK^T and PV are dense [64 x dh].[dh x 32] / [64 x 32].[32 x dh] tiles.  One CTA = 64 queries of one (sequence,
// head, sample): 4 warps x 16 query rows, Q fragments in registers, K / V tiles of 32 keys staged with cp.async (gathered
// rows), online softmax in fp32 with the decode kernel's roundings (score = fp16(fp16(q.k) * dh^-1/2); P rounded to fp16
// for P.V, the row sum kept in fp32 from the unrounded exponentials).  dh <= DH (zero padded), dh even.
struct AttnSeqs {
    int attn_func, bc, P, prime, enc_rows, tiles_per_seq;
};
struct SeqGeom { int q0, qs, nq, k0, ks, nk; };
__device__ __forceinline__ SeqGeom seq_geom(const AttnSeqs& Q, int s) {
    SeqGeom g;
    switch (Q.attn_func) {
        case 1: g.q0 = s * Q.bc; g.qs = 1; g.nq = min(Q.bc, Q.P - g.q0); g.k0 = g.q0; g.ks = 1; g.nk = g.nq; break;
        case 2: g.q0 = s; g.qs = Q.bc; g.nq = s < Q.P ? (Q.P - s + Q.bc - 1) / Q.bc : 0; g.k0 = s; g.ks = Q.bc; g.nk = g.nq; break;
        case 3: g.q0 = s * Q.bc; g.qs = 1; g.nq = min(Q.bc, Q.P - g.q0); g.k0 = (s - 1) * Q.bc; g.ks = 1; g.nk = s ? Q.bc : 0; break;
        case 6: g.q0 = 0; g.qs = 1; g.nq = Q.P; g.k0 = 0; g.ks = 1; g.nk = Q.enc_rows; break;
        case 7: g.q0 = 0; g.qs = 1; g.nq = Q.P; g.k0 = 0; g.ks = 1; g.nk = min(Q.prime, Q.P); break;
        default: g.q0 = 0; g.qs = 1; g.nq = Q.P; g.k0 = 0; g.ks = 1; g.nk = Q.P; break;
    }
    return g;
}

__device__ __forceinline__ void cp16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void ldsm_t4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// W16: head rows are 16-byte aligned (dh % 8 == 0) -> 16-byte cp.async chunks; else (5b_lyrics: dh 150) 4-byte words
template <int DH, bool W16>
__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(AttnFwd A, AttnSeqs Q) {
    constexpr int XS = DH + 8, BQ = 64, BK = 32;
    extern __shared__ __align__(16) __half asm_[];
    __half* qs = asm_;                       // [BQ][XS]
    __half* ks = qs + BQ * XS;               // [BK][XS]
    __half* vs = ks + BK * XS;               // [BK][XS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int seq = blockIdx.x / Q.tiles_per_seq, qt = blockIdx.x - seq * Q.tiles_per_seq;
    const int h = blockIdx.y, b = blockIdx.z;
    const SeqGeom G = seq_geom(Q, seq);
    const int i0 = qt * BQ;
    if (i0 >= G.nq) return;
    const int dh = A.dh, S = A.S, nv = dh >> 3;              // 16-byte chunks per row
    const bool enc = A.attn_func == 6;
    const size_t rowbase = (size_t)b * A.P;
    // ---- Q tile -> shared memory (zero rows / columns beyond nq / dh) ----
    if (W16) {
        for (int i = tid; i < BQ * (DH / 8); i += 128) {
            const int r = i / (DH / 8), c = i % (DH / 8);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (i0 + r < G.nq && c < nv)
                v = *reinterpret_cast<const uint4*>(A.qkv + (rowbase + G.q0 + (size_t)(i0 + r) * G.qs) * A.q_stride + h * dh + c * 8);
            *reinterpret_cast<uint4*>(qs + r * XS + c * 8) = v;
        }
    } else {
        for (int i = tid; i < BQ * (DH / 2); i += 128) {
            const int r = i / (DH / 2), c = i % (DH / 2);
            uint32_t v = 0;
            if (i0 + r < G.nq && 2 * c < dh)
                v = *reinterpret_cast<const uint32_t*>(A.qkv + (rowbase + G.q0 + (size_t)(i0 + r) * G.qs) * A.q_stride + h * dh + c * 2);
            *reinterpret_cast<uint32_t*>(qs + r * XS + c * 2) = v;
        }
    }
    __syncthreads();
    uint32_t qf[DH / 16][4];
#pragma unroll
    for (int k = 0; k < DH / 16; ++k) ldmatrix_x4(qf[k], qs + (warp * 16 + (lane & 15)) * XS + k * 16 + (lane >> 4) * 8);
    float o[DH / 8][4];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;
    const int qi0 = i0 + warp * 16 + g, qi1 = qi0 + 8;                  // this lane's two query rows (sequence indices)
    const long long qp0 = G.q0 + (long long)qi0 * G.qs, qp1 = G.q0 + (long long)qi1 * G.qs;
    // keys beyond the last query of this tile are never needed
    int nk = G.nk;
    if (!enc && nk > 0) {
        const long long qmax = G.q0 + (long long)(min(i0 + BQ, G.nq) - 1) * G.qs;
        const long long jm = qmax >= G.k0 ? (qmax - G.k0) / G.ks + 1 : 0;
        nk = (int)min((long long)nk, jm);
    }
    const __half* kbase;
    const __half* vbase;
    size_t kstride;
    if (enc) {
        kbase = A.kc + ((size_t)b * A.H + h) * A.enc_rows * A.dhp; vbase = A.vc + ((size_t)b * A.H + h) * A.enc_rows * A.dhp;
        kstride = (size_t)A.dhp;
    } else {
        kbase = A.qkv + (rowbase + (size_t)(nk > 0 ? G.k0 : 0)) * 3 * S + S + h * dh; vbase = kbase + S; kstride = (size_t)3 * S * G.ks;
    }
    for (int j0 = 0; j0 < nk; j0 += BK) {
        __syncthreads();                                                // the previous tile's fragment reads are done
        if (W16) {
            for (int i = tid; i < BK * (DH / 8); i += 128) {
                const int r = i / (DH / 8), c = i % (DH / 8);
                if (j0 + r < nk && c < nv) {
                    cp16(ks + r * XS + c * 8, kbase + (size_t)(j0 + r) * kstride + c * 8);
                    cp16(vs + r * XS + c * 8, vbase + (size_t)(j0 + r) * kstride + c * 8);
                } else {
                    *reinterpret_cast<uint4*>(ks + r * XS + c * 8) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(vs + r * XS + c * 8) = make_uint4(0, 0, 0, 0);
                }
            }
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        } else {
#pragma unroll 4
            for (int i = tid; i < BK * (DH / 2); i += 128) {
                const int r = i / (DH / 2), c = i % (DH / 2);
                uint32_t kv = 0, vv = 0;
                if (j0 + r < nk && 2 * c < dh) {
                    kv = *reinterpret_cast<const uint32_t*>(kbase + (size_t)(j0 + r) * kstride + c * 2);
                    vv = *reinterpret_cast<const uint32_t*>(vbase + (size_t)(j0 + r) * kstride + c * 2);
                }
                *reinterpret_cast<uint32_t*>(ks + r * XS + c * 2) = kv;
                *reinterpret_cast<uint32_t*>(vs + r * XS + c * 2) = vv;
            }
        }
        __syncthreads();
        // ---- S = Q K^T for this warp's 16 rows x 32 keys ----
        float sc[BK / 8][4];
#pragma unroll
        for (int n = 0; n < BK / 8; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
#pragma unroll
            for (int np = 0; np < BK / 16; ++np) {
                uint32_t kf[4];
                ldmatrix_x4(kf, ks + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * XS + k * 16 + ((lane >> 3) & 1) * 8);
                mma_16816(sc[2 * np], qf[k], kf[0], kf[1]);
                mma_16816(sc[2 * np + 1], qf[k], kf[2], kf[3]);
            }
        }
        // ---- mask, the reference's roundings, online softmax ----
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int n = 0; n < BK / 8; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + n * 8 + 2 * t4 + (e & 1);
                const long long kp = G.k0 + (long long)j * G.ks;
                const bool row1 = e >= 2;
                const bool ok = j < nk && (enc || kp <= (row1 ? qp1 : qp0));
                const float v = ok ? h2f_round(h2f_round(sc[n][e]) * A.scale2) : -INFINITY;
                sc[n][e] = v;
                if (row1) mx1 = fmaxf(mx1, v); else mx0 = fmaxf(mx0, v);
            }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = expf(m0 - mx0), c1 = expf(m1 - mx1);
        m0 = mx0; m1 = mx1;
        float r0 = 0.f, r1 = 0.f;
        uint32_t pf[BK / 16][4];
#pragma unroll
        for (int n = 0; n < BK / 8; ++n) {
            const float e0 = expf(sc[n][0] - m0), e1 = expf(sc[n][1] - m0), e2 = expf(sc[n][2] - m1), e3 = expf(sc[n][3] - m1);
            r0 += e0 + e1; r1 += e2 + e3;
            pf[n >> 1][(n & 1) * 2] = pack_h2(e0, e1);
            pf[n >> 1][(n & 1) * 2 + 1] = pack_h2(e2, e3);
        }
        r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
        l0 = l0 * c0 + r0; l1 = l1 * c1 + r1;
#pragma unroll
        for (int n = 0; n < DH / 8; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
        // ---- O += P V ----
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
#pragma unroll
            for (int np = 0; np < DH / 16; ++np) {
                uint32_t vf[4];
                ldsm_t4(vf, vs + (kk * 16 + (lane & 15)) * XS + np * 16 + (lane >> 4) * 8);
                mma_16816(o[2 * np], pf[kk], vf[0], vf[1]);
                mma_16816(o[2 * np + 1], pf[kk], vf[2], vf[3]);
            }
        }
    }
    // ---- out = O / l (a row without keys - previous-block attention inside the first block - is 0) ----
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
        const int d = n * 8 + 2 * t4;
        if (d < dh) {
            if (qi0 < G.nq) *reinterpret_cast<uint32_t*>(A.a + (rowbase + qp0) * S + h * dh + d) = pack_h2(o[n][0] * inv0, o[n][1] * inv0);
            if (qi1 < G.nq) *reinterpret_cast<uint32_t*>(A.a + (rowbase + qp1) * S + h * dh + d) = pack_h2(o[n][2] * inv1, o[n][3] * inv1);
        }
    }
}

template <int DH, bool W16>
int launch_attn_mma(const AttnFwd& A, const AttnSeqs& Q, int nseq, int n, cudaStream_t stream) {
    constexpr size_t smem = (size_t)(64 + 2 * 32) * (DH + 8) * 2;
    static bool attr_set[64] = {};
    int dev = 0;
    JK_CHECK_CUDA(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        JK_CHECK_CUDA(cudaFuncSetAttribute((attn_fwd_mma_kernel<DH, W16>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev & 63] = true;
    }
    attn_fwd_mma_kernel<DH, W16><<<dim3((unsigned)(nseq * Q.tiles_per_seq), A.H, n), 128, smem, stream>>>(A, Q);
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// returns 1 if the tensor-core kernel took the layer, 0 if the shape is left to attn_fwd_kernel, < 0 on error
int attn_forward_mma(const AttnFwd& A, int n, cudaStream_t stream) {
    static const bool off = getenv("JK_PREFILL_SCALAR_ATTN") != nullptr;
    if (off || A.dh % 2 != 0 || A.dh > 256) return 0;
    const bool w16 = A.dh % 8 == 0 && A.S % 8 == 0 && A.dhp % 8 == 0;
    AttnSeqs Q;
    Q.attn_func = A.attn_func; Q.bc = A.bc; Q.P = A.P; Q.prime = A.prime; Q.enc_rows = A.enc_rows;
    int nseq = 1, maxq = A.P;
    switch (A.attn_func) {
        case 1: case 3: nseq = (A.P + A.bc - 1) / A.bc; maxq = std::min(A.bc, A.P); break;
        case 2: nseq = std::min(A.bc, A.P); maxq = (A.P + A.bc - 1) / A.bc; break;
        default: break;
    }
    Q.tiles_per_seq = (maxq + 63) / 64;
    int rc;
    if (!w16) rc = A.dh <= 160 ? launch_attn_mma<160, false>(A, Q, nseq, n, stream) : launch_attn_mma<256, false>(A, Q, nseq, n, stream);
    else if (A.dh <= 32) rc = launch_attn_mma<32, true>(A, Q, nseq, n, stream);
    else if (A.dh <= 64) rc = launch_attn_mma<64, true>(A, Q, nseq, n, stream);
    else if (A.dh <= 128) rc = launch_attn_mma<128, true>(A, Q, nseq, n, stream);
    else rc = launch_attn_mma<256, true>(A, Q, nseq, n, stream);
    return rc ? rc : 1;
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
        rc = attn_forward_mma(A, n, stream);            // tensor cores when the head geometry allows (dh % 8 == 0, dh <= 256)
        if (rc < 0) return rc;
        if (rc == 0) {
            attn_fwd_kernel<<<dim3(P, H, n), kFwdThreads, fwd_smem, stream>>>(A);
            JK_CHECK_CUDA(cudaGetLastError());
        }
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
