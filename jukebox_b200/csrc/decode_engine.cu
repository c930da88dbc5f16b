// Persistent decode-step engine for Jukebox's autoregressive priors on B200 (sm_100a).
//
// One launch = one token position for up to 16 samples through the WHOLE transformer stack
// (reference: ConditionalAutoregressive2D.sample loop body, prior/autoregressive.py:222-237,
//  -> Transformer.forward(sample=True), transformer/transformer.py:169-192).
//
// Why this shape (DESIGN.md has the numbers): at n_samples <= 16 the step is HBM-bound on
// weight streaming (83-99 % of the bytes), so the design goal is "every SM streams its private,
// contiguous slice of every weight matrix exactly once per step, never stalling on the
// layer-to-layer dependency chain":
//   * grid = #SMs persistent CTAs (cooperative launch), 8 consumer warps + 1 producer warp
//   * weights are pre-packed (jk_prior_load_layer) into one contiguous byte stream per CTA, in
//     the order the CTA consumes them, already in mma.sync B-fragment order.  The producer warp
//     walks that stream with 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP) into a
//     shared-memory ring guarded by full/empty mbarriers.  It is decoupled from the compute
//     phases, so it keeps prefetching the next GEMMs / next layer (ring ~10 x 16 KB per SM) while
//     the consumers sit in a grid barrier or in the attention phase.
//   * every Conv1D at decode is [16 x K] x [K x N]: M = 16 is exactly the m16n8k16 tensor-core
//     tile, so consumers use warp-level mma.sync with fp32 accumulation; each CTA owns 8-column
//     groups of N and the full K (no cross-CTA split-K, deterministic).  tcgen05 needs M >= 64
//     and would re-read a 4x zero-padded A tile from shared memory for every 8-16 weight
//     columns; it is used where tiles are >= 128 rows (prefill GEMM, VQ-VAE).
//   * LayerNorm is fused into the GEMM's activation staging, bias / quick_gelu / residual adds
//     into its epilogue; fp16 rounding points follow the reference exactly.
//   * the layer-to-layer dependency is a grid barrier through one L2 counter.
//   * KV caches are laid out per attention pattern so that the rows a token attends are one
//     contiguous run (transpose-block layers store position p at row (p % bc)*blocks + p / bc).
//
// Numerics: activations fp16, accumulation fp32, LayerNorm/softmax fp32 - see oracle/transformer_np.py.
#include "common.cuh"
#include "../../include/jkb200.h"
#include <cooperative_groups.h>
#include <vector>
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include <math.h>

using namespace jk;

namespace {

constexpr int kConsumers = 256;
constexpr int kThreads = 288;          // 8 consumer warps + 1 producer warp
constexpr int kSlotBytes = 16384;
constexpr int kMaxSlots = 12;
constexpr int kHeaderBytes = 1024;     // barriers + LN statistics
constexpr int kLogitKT = 1024;         // K tile (floats) of the fp32 logits product
constexpr int kLogitRowsPerChunk = 4;
constexpr int kLogitRowsPerPass = 16;
constexpr int kMaxSplit = 8;

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
    int nslot, uni_bytes;
    float scale2;
    const ushort2* cols;            // [G][depth][4] : (first 8-column group, number of groups)
    const uint32_t* soff;           // [G][depth+1]  : stream offset of each layer, in 16-B units
    const uint8_t* streams;
    unsigned long long stream_stride;
    __half *h, *qkv, *a, *x1, *g;   // [16][.] fp16 activations
    float* part;                    // split-KV partials [Bmax*H*kMaxSplit][dh_pad + 2]
    unsigned* bar;
    unsigned* epoch;
    int* t;
    const float *x_emb, *pos_emb, *x_out, *start_token;
    const int* lrow0;               // [G+1] logits rows per CTA (prefix)
    LayerDev layer[JK_MAX_DEPTH];
};

struct StepArgs {
    int n;
    const float* x_in;
    const long long* tokens;
    long long tok_stride;
    const float* y_cond;
    const float* x_cond;
    long long x_cond_len;
    float* h_out;
    float* logits;
    long long logits_bstride, logits_tstride;
};

struct Ring {
    uint64_t* full;
    uint64_t* empty;
    uint8_t* base;
    int nslot;
    int slot;
    uint32_t phase;
    __device__ __forceinline__ void advance() {
        if (++slot == nslot) { slot = 0; phase ^= 1u; }
    }
};

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ int kpc_of(int ncg) {
    int k = (64 / ncg) & ~7;
    return k < 8 ? 8 : k;
}

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
    consumer_sync();
    if (threadIdx.x == 0) {
        __threadfence();
        red_release_add(bar, 1u);
        unsigned spins = 0;
        while ((int)(ld_acquire_u32(bar) - target) < 0) {
            if (++spins > (1u << 28)) __trap();
        }
        __threadfence();
    }
    consumer_sync();
}

__device__ __forceinline__ float ld_half_cg(const __half* p) {
    return __half2float(__ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

// quick_gelu(x) = x * sigmoid(1.702 x) (transformer/ops.py:33-35).  The reference's eager fp16 path
// rounds after each of its three elementwise ops; restated exactly so (x is already an fp16 value).
__device__ __forceinline__ float quick_gelu_f(float x) {
    const float z = h2f_round(1.702f * x);
    const float s = h2f_round(1.0f / (1.0f + expf(-z)));
    return x * s;
}

// ---------------------------------------------------------------------------------------
// activation staging: global fp16 [16][K] -> shared fp16 [16][K+8] (ldmatrix friendly),
// optionally through LayerNorm (fp32 math, eps 1e-5; reference transformer/ops.py:14-24)
// ---------------------------------------------------------------------------------------
template <bool LN>
__device__ void stage_acts(uint8_t* acts, float* stats, const __half* in, int K, int B,
                           const float* gamma, const float* beta) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nvec = K >> 3;
    const int astride = (K + 8) * 2;
    for (int idx = tid; idx < 16 * nvec; idx += kConsumers) {
        int r = idx / nvec, v = idx - r * nvec;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (r < B) val = ldcg_u4(in + (size_t)r * K + v * 8);
        *reinterpret_cast<uint4*>(acts + r * astride + v * 16) = val;
    }
    if (!LN) return;
    consumer_sync();
    for (int rr = 0; rr < 2; ++rr) {
        int r = warp * 2 + rr;
        if (r >= B) continue;
        const uint8_t* row = acts + r * astride;
        float s = 0.f;
        for (int v = lane; v < nvec; v += 32) {
            uint4 q = *reinterpret_cast<const uint4*>(row + v * 16);
            const __half2* hp = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) { float2 f = __half22float2(hp[e]); s += f.x + f.y; }
        }
        float mean = warp_sum(s) / (float)K;
        float ss = 0.f;
        for (int v = lane; v < nvec; v += 32) {
            uint4 q = *reinterpret_cast<const uint4*>(row + v * 16);
            const __half2* hp = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 f = __half22float2(hp[e]);
                float dx = f.x - mean, dy = f.y - mean;
                ss += dx * dx + dy * dy;
            }
        }
        float var = warp_sum(ss) / (float)K;
        if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = 1.0f / sqrtf(var + 1e-5f); }
    }
    consumer_sync();
    for (int v = tid; v < nvec; v += kConsumers) {
        float gm[8], bt[8];
        *reinterpret_cast<float4*>(gm) = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
        *reinterpret_cast<float4*>(gm + 4) = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        *reinterpret_cast<float4*>(bt) = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
        *reinterpret_cast<float4*>(bt + 4) = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
        for (int r = 0; r < B; ++r) {
            uint4* p = reinterpret_cast<uint4*>(acts + r * astride + v * 16);
            uint4 q = *p;
            __half2* hp = reinterpret_cast<__half2*>(&q);
            float mean = stats[2 * r], rstd = stats[2 * r + 1];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 f = __half22float2(hp[e]);
                f.x = (f.x - mean) * rstd * gm[2 * e] + bt[2 * e];
                f.y = (f.y - mean) * rstd * gm[2 * e + 1] + bt[2 * e + 1];
                hp[e] = __floats2half2_rn(f.x, f.y);
            }
            *p = q;
        }
    }
}

// ---------------------------------------------------------------------------------------
// one Conv1D at decode: out[b, cols of this CTA] = epilogue( acts[16,K] . Wslice[K, 8*ncg] )
// ---------------------------------------------------------------------------------------
enum { EPI_QKV = 0, EPI_PROJ = 1, EPI_FC = 2, EPI_PROJ2 = 3 };

template <bool LN, int EPI>
__device__ void gemm_phase(const EngineDev* E, Ring& ring, uint8_t* uni, float* stats, const __half* in,
                           int K, int N, int g0, int ncg, int B, const float* gamma, const float* beta,
                           const float* bias) {
    if (ncg == 0) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    stage_acts<LN>(uni, stats, in, K, B, gamma, beta);
    consumer_sync();

    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
    const int nkk = K >> 4;
    const int kpc = kpc_of(ncg);
    const int astride = (K + 8) * 2;
    const uint8_t* arow = uni + (lane & 15) * astride + (lane >> 4) * 16;
    for (int kk0 = 0; kk0 < nkk; kk0 += kpc) {
        const int nk = min(kpc, nkk - kk0);
        mbar_wait(&ring.full[ring.slot], ring.phase);
        const uint8_t* sl = ring.base + ring.slot * kSlotBytes;
        for (int i = warp; i < nk; i += 8) {
            uint32_t a[4];
            ldmatrix_x4(a, arow + (kk0 + i) * 32);
            const uint2* bp = reinterpret_cast<const uint2*>(sl + (size_t)(i * ncg) * 256) + lane;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < ncg) {
                    uint2 b = bp[j * 32];
                    mma_16816(acc[j], a, b.x, b.y);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[ring.slot]);
        ring.advance();
    }
    consumer_sync();                       // everyone is done reading the staged activations
    float* red = reinterpret_cast<float*>(uni);   // [8 warps][ncg][16][8]
    {
        const int r0 = lane >> 2, c0 = (lane & 3) * 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < ncg) {
                float* d = red + ((warp * ncg + j) * 16) * 8;
                d[r0 * 8 + c0] = acc[j][0];
                d[r0 * 8 + c0 + 1] = acc[j][1];
                d[(r0 + 8) * 8 + c0] = acc[j][2];
                d[(r0 + 8) * 8 + c0 + 1] = acc[j][3];
            }
        }
    }
    consumer_sync();
    const int nc = ncg * 8;
    for (int e = tid; e < B * nc; e += kConsumers) {
        const int b = e / nc, cc = e - b * nc;
        const int j = cc >> 3, col = cc & 7;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[((w * ncg + j) * 16 + b) * 8 + col];
        const int gc = g0 * 8 + cc;
        const float y = h2f_round(s + bias[gc]);          // Conv1D output, rounded once to fp16
        if (EPI == EPI_QKV) {
            E->qkv[(size_t)b * N + gc] = __float2half_rn(y);
        } else if (EPI == EPI_PROJ) {                      // x1 = fp16(h + a)
            float hv = ld_half_cg(E->h + (size_t)b * N + gc);
            E->x1[(size_t)b * N + gc] = __float2half_rn(hv + y);
        } else if (EPI == EPI_FC) {                        // quick_gelu (transformer/ops.py:33-35)
            E->g[(size_t)b * N + gc] = __float2half_rn(quick_gelu_f(y));
        } else {                                           // h = fp16(x1 + m)
            float xv = ld_half_cg(E->x1 + (size_t)b * N + gc);
            E->h[(size_t)b * N + gc] = __float2half_rn(xv + y);
        }
    }
    consumer_sync();                       // red region is reused by the next phase's staging
}

// ---------------------------------------------------------------------------------------
// attention for one (sample, head, kv-split) work item; q_len == 1
// (reference: factored_attention.py:82-133 + per-pattern sample branches :135-228)
// ---------------------------------------------------------------------------------------
struct AttnGeom {
    int R;        // rows attended
    int base;     // first cache row of the attended run
    int cur;      // 1 if the current token is the last attended row
    int wrow;     // cache row the current token's k/v is written to (-1: none)
};

__device__ __forceinline__ AttnGeom attn_geom(const EngineDev* E, const LayerDev& LD, int p) {
    AttnGeom g;
    const int bc = E->bc;
    switch (LD.attn_func) {
        case 0: g.R = p + 1; g.base = 0; g.cur = 1; g.wrow = p; break;
        case 1: g.R = p % bc + 1; g.base = 0; g.cur = 1; g.wrow = p % bc; break;
        case 2: g.base = (p % bc) * E->blocks; g.R = p / bc + 1; g.cur = 1; g.wrow = g.base + p / bc; break;
        case 3:
            g.R = (p >= bc) ? bc : 0; g.base = ((p / bc + 1) & 1) * bc; g.cur = 0;
            g.wrow = ((p / bc) & 1) * bc + p % bc; break;
        case 7:
            g.R = min(p + 1, E->prime_pad); g.base = 0; g.cur = (p < E->prime_pad) ? 1 : 0;
            g.wrow = (p < E->prime_pad) ? p : -1; break;
        default: g.R = E->enc_dims; g.base = 0; g.cur = 0; g.wrow = -1; break;   // 6
    }
    return g;
}

__device__ __forceinline__ int attn_nsplit(const EngineDev* E, const LayerDev& LD, int B, int R) {
    if (LD.attn_func == 1 || LD.attn_func == 2 || LD.attn_func == 3) return 1;
    if (R <= 256) return 1;
    int ns = E->G / (B * E->H);
    ns = max(1, min(kMaxSplit, ns));
    return ns;
}

__device__ void attn_item(const EngineDev* E, const LayerDev& LD, uint8_t* uni, float* stats, int b, int h,
                          int s, int ns, const AttnGeom& G) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dh = E->dh, dhp = E->dh_pad, S = E->S;
    const int nvec = dhp >> 3;
    int LPR = 1;
    while (LPR < nvec && LPR < 32) LPR <<= 1;
    const int RPW = 32 / LPR;
    const int grp = lane / LPR, li = lane % LPR;

    float* qs = reinterpret_cast<float*>(uni);          // [dhp]
    float* ks = qs + dhp;                               // [dhp] current token's k
    float* vs = ks + dhp;                               // [dhp]
    float* red = vs + dhp;                              // [8][dhp]
    float* sc = red + 8 * dhp;                          // scores of this split
    const int qkv_stride = (LD.attn_func == 6) ? S : 3 * S;
    const __half* qrow = E->qkv + (size_t)b * qkv_stride + h * dh;
    const size_t cbase = ((size_t)(b * E->H + h)) * LD.rows;
    const bool writer = (s == ns - 1);

    for (int d = tid; d < dhp; d += kConsumers) {
        float q = 0.f, k = 0.f, v = 0.f;
        if (d < dh) {
            q = ld_half_cg(qrow + d);
            if (LD.attn_func != 6) {
                k = ld_half_cg(qrow + S + d);
                v = ld_half_cg(qrow + 2 * S + d);
                if (writer && G.wrow >= 0) {
                    LD.kc[(cbase + G.wrow) * dhp + d] = __float2half_rn(k);
                    LD.vc[(cbase + G.wrow) * dhp + d] = __float2half_rn(v);
                }
            }
        }
        qs[d] = q; ks[d] = k; vs[d] = v;
    }
    consumer_sync();
    const int R = G.R;
    if (R == 0) {   // prev-block attention inside the first block: keys/values are zeros -> output 0
        for (int d = tid; d < dh; d += kConsumers) E->a[(size_t)b * S + h * dh + d] = __float2half_rn(0.f);
        consumer_sync();
        return;
    }
    const int i0 = (int)(((long long)R * s) / ns), i1 = (int)(((long long)R * (s + 1)) / ns);
    const int n = i1 - i0;
    const int cur_idx = G.cur ? R - 1 : -1;
    const __half* kbase = LD.kc + (cbase + G.base) * dhp;
    const __half* vbase = LD.vc + (cbase + G.base) * dhp;

    // ---- scores: s = fp16(fp16(q.k) * dh^-1/2) --------------------------------------------
    for (int base = i0 + warp * RPW; base < i1; base += 8 * RPW) {
        const int idx = base + grp;
        float dot = 0.f;
        if (idx < i1) {
            if (idx == cur_idx) {
                for (int v = li; v < nvec; v += LPR) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) dot += qs[v * 8 + e] * ks[v * 8 + e];
                }
            } else {
                const __half* row = kbase + (size_t)idx * dhp;
                for (int v = li; v < nvec; v += LPR) {
                    uint4 q4 = __ldg(reinterpret_cast<const uint4*>(row + v * 8));
                    const __half2* hp = reinterpret_cast<const __half2*>(&q4);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float2 f = __half22float2(hp[e]);
                        dot += qs[v * 8 + 2 * e] * f.x + qs[v * 8 + 2 * e + 1] * f.y;
                    }
                }
            }
        }
        for (int o = LPR >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (idx < i1 && li == 0) sc[idx - i0] = h2f_round(h2f_round(dot) * E->scale2);
    }
    consumer_sync();
    // ---- softmax statistics (fp32) ----------------------------------------------------------
    float m = -INFINITY;
    for (int i = tid; i < n; i += kConsumers) m = fmaxf(m, sc[i]);
    m = warp_max(m);
    if (lane == 0) stats[32 + warp] = m;
    consumer_sync();
    m = stats[32];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, stats[32 + w]);
    float l = 0.f;
    for (int i = tid; i < n; i += kConsumers) {
        float e = expf(sc[i] - m);
        sc[i] = e;
        l += e;
    }
    l = warp_sum(l);
    if (lane == 0) stats[40 + warp] = l;
    consumer_sync();
    l = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) l += stats[40 + w];
    if (ns == 1) {   // reference rounding: P = fp16(softmax)
        for (int i = tid; i < n; i += kConsumers) sc[i] = h2f_round(sc[i] / l);
        consumer_sync();
    }
    // ---- P.V ------------------------------------------------------------------------------
    float acc[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[u][e] = 0.f;
    for (int base = i0 + warp * RPW; base < i1; base += 8 * RPW) {
        const int idx = base + grp;
        if (idx < i1) {
            const float pw = sc[idx - i0];
            if (idx == cur_idx) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    int v = li + u * LPR;
                    if (v < nvec) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[u][e] += pw * vs[v * 8 + e];
                    }
                }
            } else {
                const __half* row = vbase + (size_t)idx * dhp;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    int v = li + u * LPR;
                    if (v < nvec) {
                        uint4 q4 = __ldg(reinterpret_cast<const uint4*>(row + v * 8));
                        const __half2* hp = reinterpret_cast<const __half2*>(&q4);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float2 f = __half22float2(hp[e]);
                            acc[u][2 * e] += pw * f.x;
                            acc[u][2 * e + 1] += pw * f.y;
                        }
                    }
                }
            }
        }
    }
    for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[u][e] += __shfl_xor_sync(0xffffffffu, acc[u][e], o);
    }
    if (grp == 0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            int v = li + u * LPR;
            if (v < nvec) {
#pragma unroll
                for (int e = 0; e < 8; ++e) red[warp * dhp + v * 8 + e] = acc[u][e];
            }
        }
    }
    consumer_sync();
    if (ns == 1) {
        for (int d = tid; d < dh; d += kConsumers) {
            float o = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) o += red[w * dhp + d];
            E->a[(size_t)b * S + h * dh + d] = __float2half_rn(o);
        }
    } else {
        float* part = E->part + ((size_t)((b * E->H + h) * kMaxSplit + s)) * (dhp + 2);
        if (tid == 0) { part[0] = m; part[1] = l; }
        for (int d = tid; d < dhp; d += kConsumers) {
            float o = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) o += red[w * dhp + d];
            part[2 + d] = o;
        }
    }
    consumer_sync();
}

__device__ void attn_merge(const EngineDev* E, int b, int h, int ns) {
    const int dh = E->dh, dhp = E->dh_pad;
    const float* part = E->part + ((size_t)((b * E->H + h) * kMaxSplit)) * (dhp + 2);
    float M = -INFINITY;
    for (int s = 0; s < ns; ++s) M = fmaxf(M, __ldcg(part + (size_t)s * (dhp + 2)));
    float Lsum = 0.f;
    for (int s = 0; s < ns; ++s)
        Lsum += __ldcg(part + (size_t)s * (dhp + 2) + 1) * expf(__ldcg(part + (size_t)s * (dhp + 2)) - M);
    for (int d = threadIdx.x; d < dh; d += kConsumers) {
        float o = 0.f;
        for (int s = 0; s < ns; ++s)
            o += __ldcg(part + (size_t)s * (dhp + 2) + 2 + d) * expf(__ldcg(part + (size_t)s * (dhp + 2)) - M);
        E->a[(size_t)b * E->S + h * dh + d] = __float2half_rn(o / Lsum);
    }
}

// ---------------------------------------------------------------------------------------
// producer warp: walks this CTA's weight stream (and the logits rows) in consumption order
// ---------------------------------------------------------------------------------------
__device__ void producer_loop(const EngineDev* E, Ring ring, bool do_logits, int c) {
    if ((threadIdx.x & 31) != 0) return;
    const uint8_t* src = E->streams + (size_t)c * E->stream_stride + (size_t)E->soff[(size_t)c * (E->depth + 1)] * 16;
    for (int l = 0; l < E->depth; ++l) {
        const ushort2* cl = E->cols + ((size_t)c * E->depth + l) * 4;
        const int Ks[4] = {E->W, E->S, E->W, E->M};
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
            const int ncg = cl[gi].y;
            if (ncg == 0) continue;
            const int nkk = Ks[gi] >> 4, kpc = kpc_of(ncg);
            for (int kk0 = 0; kk0 < nkk; kk0 += kpc) {
                const int nk = min(kpc, nkk - kk0);
                const uint32_t bytes = (uint32_t)nk * ncg * 256u;
                mbar_wait(&ring.empty[ring.slot], ring.phase ^ 1u);
                mbar_expect_tx(&ring.full[ring.slot], bytes);
                tma_bulk_g2s(ring.base + ring.slot * kSlotBytes, src, bytes, &ring.full[ring.slot]);
                src += bytes;
                ring.advance();
            }
        }
    }
    if (do_logits) {
        const int r0 = E->lrow0[c], r1 = E->lrow0[c + 1];
        const int W = E->W;
        for (int pr = r0; pr < r1; pr += kLogitRowsPerPass) {
            const int pe = min(r1, pr + kLogitRowsPerPass);
            for (int k0 = 0; k0 < W; k0 += kLogitKT) {
                const int kt = min(kLogitKT, W - k0);
                for (int r = pr; r < pe; r += kLogitRowsPerChunk) {
                    const int nr = min(kLogitRowsPerChunk, pe - r);
                    mbar_wait(&ring.empty[ring.slot], ring.phase ^ 1u);
                    mbar_expect_tx(&ring.full[ring.slot], (uint32_t)(nr * kt * 4));
                    for (int i = 0; i < nr; ++i)
                        tma_bulk_g2s(ring.base + ring.slot * kSlotBytes + i * kt * 4,
                                     E->x_out + (size_t)(r + i) * W + k0, (uint32_t)(kt * 4), &ring.full[ring.slot]);
                    ring.advance();
                }
            }
        }
    }
}

// fp32 logits: logits[b, r] = sum_k y[b, k] * x_out[r, k],  y = float(h) (+ cond)
// (reference autoregressive.py:226-229: fp32 nn.Linear on the fp32 transformer output)
__device__ void logits_phase(const EngineDev* E, const StepArgs& A, Ring& ring, uint8_t* uni, int c, int t) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = E->lrow0[c], r1 = E->lrow0[c + 1];
    const int W = E->W, B = A.n;
    float* ys = reinterpret_cast<float*>(uni);     // [16][kt]
    for (int pr = r0; pr < r1; pr += kLogitRowsPerPass) {
        const int pe = min(r1, pr + kLogitRowsPerPass);
        float acc[kLogitRowsPerPass][2];
#pragma unroll
        for (int i = 0; i < kLogitRowsPerPass; ++i) acc[i][0] = acc[i][1] = 0.f;
        for (int k0 = 0; k0 < W; k0 += kLogitKT) {
            const int kt = min(kLogitKT, W - k0);
            consumer_sync();
            for (int idx = tid; idx < 16 * kt; idx += kConsumers) {
                int b = idx / kt, k = idx - b * kt;
                float y = 0.f;
                if (b < B) {
                    y = ld_half_cg(E->h + (size_t)b * W + k0 + k);
                    if (E->add_cond_after && A.x_cond)
                        y += A.x_cond[((size_t)b * A.x_cond_len + (A.x_cond_len > 1 ? t : 0)) * W + k0 + k];
                }
                ys[b * kt + k] = y;
            }
            consumer_sync();
            const float* y0 = ys + (warp * 2) * kt;
            const float* y1 = y0 + kt;
#pragma unroll
            for (int rc = 0; rc < kLogitRowsPerPass / kLogitRowsPerChunk; ++rc) {
                const int r = pr + rc * kLogitRowsPerChunk;
                if (r < pe) {
                    const int nr = min(kLogitRowsPerChunk, pe - r);
                    mbar_wait(&ring.full[ring.slot], ring.phase);
                    const float* wsl = reinterpret_cast<const float*>(ring.base + ring.slot * kSlotBytes);
                    for (int k = lane * 4; k < kt; k += 128) {
                        float4 a0 = *reinterpret_cast<const float4*>(y0 + k);
                        float4 a1 = *reinterpret_cast<const float4*>(y1 + k);
#pragma unroll
                        for (int i = 0; i < kLogitRowsPerChunk; ++i) {
                            if (i < nr) {
                                float4 w4 = *reinterpret_cast<const float4*>(wsl + i * kt + k);
                                acc[rc * kLogitRowsPerChunk + i][0] += a0.x * w4.x + a0.y * w4.y + a0.z * w4.z + a0.w * w4.w;
                                acc[rc * kLogitRowsPerChunk + i][1] += a1.x * w4.x + a1.y * w4.y + a1.z * w4.z + a1.w * w4.w;
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ring.empty[ring.slot]);
                    ring.advance();
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kLogitRowsPerPass; ++i) {
            float v0 = warp_sum(acc[i][0]), v1 = warp_sum(acc[i][1]);
            const int r = pr + i;
            if (lane == 0 && r < pe) {
                const int b0 = warp * 2;
                if (b0 < B) A.logits[(size_t)b0 * A.logits_bstride + (size_t)t * A.logits_tstride + r] = v0;
                if (b0 + 1 < B) A.logits[(size_t)(b0 + 1) * A.logits_bstride + (size_t)t * A.logits_tstride + r] = v1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) jk_decode_step_kernel(const EngineDev* __restrict__ E, StepArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kMaxSlots;
    float* stats = reinterpret_cast<float*>(smem + 256);    // [32] LN stats + [16] softmax scratch
    uint8_t* uni = smem + kHeaderBytes;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int c = blockIdx.x;
    Ring ring;
    ring.full = full; ring.empty = empty; ring.base = uni + E->uni_bytes; ring.nslot = E->nslot;
    ring.slot = 0; ring.phase = 0;
    if (tid == 0) {
        for (int i = 0; i < E->nslot; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        mbar_fence_init();
    }
    __syncthreads();
    const bool do_logits = (A.logits != nullptr) && E->bins > 0;
    if (warp == 8) {
        producer_loop(E, ring, do_logits, c);
        return;
    }
    const int t = *reinterpret_cast<volatile const int*>(E->t);
    const unsigned epoch0 = *reinterpret_cast<volatile const unsigned*>(E->epoch);
    unsigned nbar = 0;
    const int B = A.n, W = E->W, S = E->S, M = E->M, G = E->G;
#define GRID_BARRIER() do { ++nbar; grid_barrier(E->bar, epoch0 + nbar * (unsigned)G); } while (0)

    // ---- P0: embedding (autoregressive.py:177-197) or an externally embedded activation ------
    for (int e = c * kConsumers + tid; e < B * W; e += G * kConsumers) {
        const int b = e / W, col = e - b * W;
        float x;
        if (A.x_in) {
            x = A.x_in[e];
        } else {
            if (t == 0) x = A.y_cond ? A.y_cond[e] : E->start_token[col];
            else x = E->x_emb[(size_t)A.tokens[(size_t)b * A.tok_stride + t - 1] * W + col];
            x += E->pos_emb[(size_t)t * W + col];
            if (A.x_cond) x += A.x_cond[((size_t)b * A.x_cond_len + (A.x_cond_len > 1 ? t : 0)) * W + col];
        }
        E->h[e] = __float2half_rn(x);
    }
    GRID_BARRIER();

    for (int l = 0; l < E->depth; ++l) {
        const LayerDev& LD = E->layer[l];
        const ushort2* cl = E->cols + ((size_t)c * E->depth + l) * 4;
        const int Nqkv = (LD.attn_func == 6) ? S : 3 * S;
        gemm_phase<true, EPI_QKV>(E, ring, uni, stats, E->h, W, Nqkv, cl[0].x, cl[0].y, B, LD.ln0_g, LD.ln0_b, LD.b_qkv);
        GRID_BARRIER();
        {
            const AttnGeom geo = attn_geom(E, LD, t);
            const int ns = attn_nsplit(E, LD, B, geo.R);
            for (int it = c; it < B * E->H * ns; it += G) {
                const int s = it % ns, bh = it / ns;
                attn_item(E, LD, uni, stats, bh / E->H, bh % E->H, s, ns, geo);
            }
            GRID_BARRIER();
            if (ns > 1) {
                for (int it = c; it < B * E->H; it += G) attn_merge(E, it / E->H, it % E->H, ns);
                GRID_BARRIER();
            }
        }
        gemm_phase<false, EPI_PROJ>(E, ring, uni, stats, E->a, S, W, cl[1].x, cl[1].y, B, nullptr, nullptr, LD.b_o);
        GRID_BARRIER();
        gemm_phase<true, EPI_FC>(E, ring, uni, stats, E->x1, W, M, cl[2].x, cl[2].y, B, LD.ln1_g, LD.ln1_b, LD.b_1);
        GRID_BARRIER();
        gemm_phase<false, EPI_PROJ2>(E, ring, uni, stats, E->g, M, W, cl[3].x, cl[3].y, B, nullptr, nullptr, LD.b_2);
        GRID_BARRIER();
    }
    if (A.h_out) {
        for (int e = c * kConsumers + tid; e < B * W; e += G * kConsumers)
            A.h_out[e] = ld_half_cg(E->h + e);
    }
    if (do_logits) logits_phase(E, A, ring, uni, c, t);
    if (c == 0 && tid == 0) {
        *E->t = t + 1;
        *E->epoch = epoch0 + nbar * (unsigned)G;
    }
#undef GRID_BARRIER
}

// ---------------------------------------------------------------------------------------
// packing kernels (one-time, at weight load)
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ __half to_half(T v);
template <>
__device__ __forceinline__ __half to_half<float>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __half to_half<__half>(__half v) { return v; }

// src: Conv1D.w [K][N] row-major.  dst: per-CTA streams; this kernel fills GEMM `gi` of layer `l`.
// grid.x = CTA index c, threads loop over this CTA's (kk, j, lane) fragment slots.
template <typename T>
__global__ void pack_gemm_kernel(const T* __restrict__ src, int K, int N, uint8_t* streams,
                                 unsigned long long stream_stride, const ushort2* cols, const uint32_t* goff,
                                 int depth, int l, int gi) {
    const int c = blockIdx.x;
    const ushort2 cg = cols[((size_t)c * depth + l) * 4 + gi];
    const int g0 = cg.x, ncg = cg.y;
    if (ncg == 0) return;
    uint8_t* dst = streams + (size_t)c * stream_stride + (size_t)goff[((size_t)c * depth + l) * 4 + gi] * 16;
    const int nkk = K >> 4;
    const int total = nkk * ncg * 32;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int lane = i & 31, u = i >> 5;
        const int j = u % ncg, kk = u / ncg;
        const int n = (g0 + j) * 8 + (lane >> 2);
        const int k = kk * 16 + (lane & 3) * 2;
        __half v[4];
        v[0] = to_half<T>(src[(size_t)k * N + n]);
        v[1] = to_half<T>(src[(size_t)(k + 1) * N + n]);
        v[2] = to_half<T>(src[(size_t)(k + 8) * N + n]);
        v[3] = to_half<T>(src[(size_t)(k + 9) * N + n]);
        *reinterpret_cast<uint2*>(dst + (size_t)u * 256 + lane * 8) = *reinterpret_cast<uint2*>(v);
    }
}

template <typename T>
__global__ void round_bias_kernel(const T* __restrict__ src, float* dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __half2float(to_half<T>(src[i]));
}
template <typename T>
__global__ void to_half_kernel(const T* __restrict__ src, __half* dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = to_half<T>(src[i]);
}

// encoder K/V for attn_func 6: kv[b, e, :] = fp16(enc[b, e, :] . Wkv + b)   (once per window)
// simple tiled fp16 GEMM with fp32 accumulation; M = n*enc_dims rows.
__global__ void enc_kv_kernel(const float* __restrict__ enc, const __half* __restrict__ w, const float* __restrict__ bias,
                              __half* kc, __half* vc, int rows_total, int E_dims, int W, int S, int H, int dh, int dhp) {
    __shared__ float xs[16][33];
    __shared__ float ws[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;       // 32 x 16
    const int row0 = blockIdx.y * 16, col0 = blockIdx.x * 32;
    float acc = 0.f;
    for (int k0 = 0; k0 < W; k0 += 32) {
        int r = row0 + ty;
        xs[ty][tx] = (r < rows_total && k0 + tx < W) ? __half2float(__float2half_rn(enc[(size_t)r * W + k0 + tx])) : 0.f;
        for (int kk = ty; kk < 32; kk += 16)
            ws[kk][tx] = (k0 + kk < W && col0 + tx < 2 * S) ? __half2float(w[(size_t)(k0 + kk) * 2 * S + col0 + tx]) : 0.f;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) acc += xs[ty][kk] * ws[kk][tx];
        __syncthreads();
    }
    const int r = row0 + ty, cidx = col0 + tx;
    if (r < rows_total && cidx < 2 * S) {
        const float y = acc + bias[cidx];
        const int b = r / E_dims, e = r % E_dims;
        const int which = cidx / S, cs = cidx % S;
        const int h = cs / dh, d = cs % dh;
        __half* dst = which ? vc : kc;
        dst[(((size_t)(b * H + h)) * E_dims + e) * dhp + d] = __float2half_rn(y);
    }
}

}  // namespace

// =========================================================================================
// host side
// =========================================================================================
struct jk_prior {
    jk_prior_config cfg;
    EngineDev host;            // host mirror of the device struct
    EngineDev* dev;            // in arena
    uint8_t* arena;
    size_t arena_bytes;
    int G;
    int smem_bytes;
    int t_host;
    std::vector<ushort2> cols;
    std::vector<uint32_t> goff;      // [G][depth][4] per-GEMM stream offsets (16-B units)
    uint32_t* d_goff;
    ushort2* d_cols;
    // arena sub-allocations for per-layer small params
    std::vector<float*> bias_ptr[4];
    std::vector<float*> ln_ptr[4];
    std::vector<__half*> enc_w;
    std::vector<float*> enc_b;
};

namespace {

struct Layout {
    size_t off_dev, off_cols, off_soff, off_goff, off_lrow, off_streams, off_small, off_cache, off_h, off_x1, off_qkv, off_a, off_g, off_part, off_sync, total;
    size_t stream_stride;
    std::vector<ushort2> cols;
    std::vector<uint32_t> soff, goff;
    std::vector<int> lrow;
    std::vector<size_t> cache_off;      // per layer (K); V follows
    std::vector<size_t> cache_bytes;
    std::vector<int> cache_rows;
    size_t small_per_layer;
    int dh, dh_pad, bc, prime_pad, uni_bytes, nslot, smem_bytes;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int cache_rows_for(const jk_prior_config& c, int af, int bc, int prime_pad) {
    switch (af) {
        case 0: return c.n_ctx;
        case 1: return bc;
        case 2: return c.n_ctx;
        case 3: return 2 * bc;
        case 6: return c.encoder_dims;
        case 7: return prime_pad;
    }
    return -1;
}

int compute_layout(const jk_prior_config& c, int G, Layout& L) {
    JK_REQUIRE(c.depth >= 1 && c.depth <= JK_MAX_DEPTH, "depth %d out of range", c.depth);
    JK_REQUIRE(c.max_batch >= 1 && c.max_batch <= JK_MAX_BATCH, "max_batch %d out of range (<= %d)", c.max_batch, JK_MAX_BATCH);
    JK_REQUIRE(c.width % 16 == 0 && c.n_state % 16 == 0 && c.mlp_width % 16 == 0,
               "width/n_state/mlp_width must be multiples of 16 (got %d/%d/%d)", c.width, c.n_state, c.mlp_width);
    JK_REQUIRE(c.n_state % c.heads == 0, "n_state %% heads != 0");
    L.dh = c.n_state / c.heads;
    L.dh_pad = (int)align_up(L.dh, 8);
    JK_REQUIRE(L.dh_pad <= 512, "head_dim %d > 512 unsupported", L.dh);
    L.bc = c.blocks > 0 ? c.n_ctx / c.blocks : c.n_ctx;
    JK_REQUIRE(c.blocks == 0 || c.n_ctx % c.blocks == 0, "n_ctx %% blocks != 0");
    L.prime_pad = c.blocks > 0 ? (c.prime_len / c.blocks + 1) * c.blocks : 0;
    const int depth = c.depth;
    L.cols.assign((size_t)G * depth * 4, make_ushort2(0, 0));
    L.goff.assign((size_t)G * depth * 4, 0);
    L.soff.assign((size_t)G * (depth + 1), 0);
    std::vector<unsigned long long> cum(G, 0);
    std::vector<int> order(G);
    for (int l = 0; l < depth; ++l) {
        const int af = c.attn_func[l];
        JK_REQUIRE(af == 0 || af == 1 || af == 2 || af == 3 || af == 6 || af == 7, "attn_func %d has no decode path", af);
        JK_REQUIRE(af == 0 || c.blocks > 0 || af == 6, "attn_func %d needs blocks", af);
        const int Ks[4] = {c.width, c.n_state, c.width, c.mlp_width};
        const int Ns[4] = {af == 6 ? c.n_state : 3 * c.n_state, c.width, c.mlp_width, c.width};
        for (int gi = 0; gi < 4; ++gi) {
            JK_REQUIRE(Ns[gi] % 8 == 0, "n_out %d not a multiple of 8", Ns[gi]);
            const int groups = Ns[gi] / 8, base = groups / G, extra = groups % G;
            JK_REQUIRE(base + (extra ? 1 : 0) <= 8, "n_out %d too wide for %d CTAs (max 64 columns per CTA)", Ns[gi], G);
            for (int i = 0; i < G; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cum[a] < cum[b]; });
            std::vector<int> n(G, base);
            for (int i = 0; i < extra; ++i) n[order[i]] += 1;
            int g0 = 0;
            for (int cta = 0; cta < G; ++cta) {
                L.cols[((size_t)cta * depth + l) * 4 + gi] = make_ushort2((unsigned short)g0, (unsigned short)n[cta]);
                L.goff[((size_t)cta * depth + l) * 4 + gi] = (uint32_t)(cum[cta] / 16);
                if (gi == 0) L.soff[(size_t)cta * (depth + 1) + l] = (uint32_t)(cum[cta] / 16);
                cum[cta] += (unsigned long long)n[cta] * (Ks[gi] / 16) * 256ull;
                g0 += n[cta];
            }
        }
    }
    unsigned long long mx = 0;
    for (int cta = 0; cta < G; ++cta) {
        L.soff[(size_t)cta * (depth + 1) + depth] = (uint32_t)(cum[cta] / 16);
        mx = std::max(mx, cum[cta]);
    }
    JK_REQUIRE(mx / 16 < 0xffffffffull, "stream too long");
    L.stream_stride = align_up((size_t)mx + 256, 256);
    L.lrow.assign(G + 1, 0);
    for (int cta = 0; cta <= G; ++cta) L.lrow[cta] = (int)((long long)c.bins * cta / G);

    const int Kmax = std::max(c.width, std::max(c.n_state, c.mlp_width));
    size_t uni = (size_t)16 * (Kmax + 8) * 2;
    uni = std::max(uni, (size_t)16 * kLogitKT * 4);
    uni = std::max(uni, (size_t)8 * 8 * 16 * 8 * 4);                              // cross-warp reduction
    size_t attn = (size_t)(3 + 8) * L.dh_pad * 4 + (size_t)std::max(c.n_ctx, std::max(c.encoder_dims, 1)) * 4;
    uni = std::max(uni, attn);
    L.uni_bytes = (int)align_up(uni, 1024);
    const int max_smem = 232448;
    int nslot = (max_smem - kHeaderBytes - L.uni_bytes) / kSlotBytes;
    nslot = std::min(nslot, kMaxSlots);
    JK_REQUIRE(nslot >= 2, "not enough shared memory for the weight ring (uni %d bytes)", L.uni_bytes);
    L.nslot = nslot;
    L.smem_bytes = kHeaderBytes + L.uni_bytes + nslot * kSlotBytes;

    size_t off = 0;
    L.off_dev = off; off = align_up(off + sizeof(EngineDev), 256);
    L.off_cols = off; off = align_up(off + L.cols.size() * sizeof(ushort2), 256);
    L.off_soff = off; off = align_up(off + L.soff.size() * 4, 256);
    L.off_goff = off; off = align_up(off + L.goff.size() * 4, 256);
    L.off_lrow = off; off = align_up(off + L.lrow.size() * 4, 256);
    L.off_streams = off; off = align_up(off + (size_t)G * L.stream_stride, 256);
    // per layer small params: 4 biases + 4 LN vectors (+ enc kv weights/bias for type 6)
    L.small_per_layer = align_up((size_t)(3 * c.n_state + c.width + c.mlp_width + c.width + 4 * c.width) * 4, 256);
    L.off_small = off; off += L.small_per_layer * depth;
    for (int l = 0; l < depth; ++l)
        if (c.attn_func[l] == 6) off = align_up(off + (size_t)c.width * 2 * c.n_state * 2 + 2 * c.n_state * 4 + 512, 256);
    L.off_cache = off;
    L.cache_off.resize(depth); L.cache_bytes.resize(depth); L.cache_rows.resize(depth);
    for (int l = 0; l < depth; ++l) {
        int rows = cache_rows_for(c, c.attn_func[l], L.bc, L.prime_pad);
        L.cache_rows[l] = rows;
        size_t bytes = align_up((size_t)c.max_batch * c.heads * rows * L.dh_pad * 2, 256);
        L.cache_off[l] = off; L.cache_bytes[l] = bytes;
        off += 2 * bytes;
    }
    L.off_h = off;   off = align_up(off + (size_t)16 * c.width * 2, 256);
    L.off_x1 = off;  off = align_up(off + (size_t)16 * c.width * 2, 256);
    L.off_qkv = off; off = align_up(off + (size_t)16 * 3 * c.n_state * 2, 256);
    L.off_a = off;   off = align_up(off + (size_t)16 * c.n_state * 2, 256);
    L.off_g = off;   off = align_up(off + (size_t)16 * c.mlp_width * 2, 256);
    L.off_part = off; off = align_up(off + (size_t)c.max_batch * c.heads * kMaxSplit * (L.dh_pad + 2) * 4, 256);
    L.off_sync = off; off += 256;
    L.total = off;
    return 0;
}

int device_sms() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms;
}

}  // namespace

extern "C" int jk_device_sm_count(int* out) {
    int s = device_sms();
    JK_REQUIRE(s > 0, "no CUDA device");
    *out = s;
    return 0;
}

extern "C" int jk_prior_arena_bytes(const jk_prior_config* cfg, size_t* bytes) {
    JK_REQUIRE(cfg && bytes, "null argument");
    int G = device_sms();
    JK_REQUIRE(G > 0, "no CUDA device (the decode engine has no CPU path)");
    Layout L;
    int rc = compute_layout(*cfg, G, L);
    if (rc) return rc;
    *bytes = L.total;
    return 0;
}

extern "C" int jk_prior_create(const jk_prior_config* cfg, void* arena, size_t arena_bytes, jk_prior** out,
                               jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(cfg && arena && out, "null argument");
    int G = device_sms();
    JK_REQUIRE(G > 0, "no CUDA device (the decode engine has no CPU path)");
    Layout L;
    int rc = compute_layout(*cfg, G, L);
    if (rc) return rc;
    JK_REQUIRE(arena_bytes >= L.total, "arena too small: %zu < %zu", arena_bytes, L.total);
    JK_REQUIRE(((uintptr_t)arena & 255) == 0, "arena must be 256-byte aligned");
    JK_CHECK_CUDA(cudaMemsetAsync(arena, 0, L.total, stream));
    jk_prior* p = new jk_prior();
    p->cfg = *cfg; p->arena = (uint8_t*)arena; p->arena_bytes = arena_bytes; p->G = G; p->t_host = 0;
    p->smem_bytes = L.smem_bytes;
    p->cols = L.cols; p->goff = L.goff;
    uint8_t* A = p->arena;
    EngineDev& E = p->host;
    memset(&E, 0, sizeof(E));
    E.W = cfg->width; E.S = cfg->n_state; E.M = cfg->mlp_width; E.H = cfg->heads; E.dh = L.dh; E.dh_pad = L.dh_pad;
    E.L = cfg->n_ctx; E.blocks = cfg->blocks; E.bc = L.bc; E.bins = cfg->bins; E.prime_pad = L.prime_pad;
    E.enc_dims = cfg->encoder_dims; E.Bmax = cfg->max_batch; E.add_cond_after = cfg->add_cond_after;
    E.depth = cfg->depth; E.G = G; E.nslot = L.nslot; E.uni_bytes = L.uni_bytes;
    {   // reference: scale = 1/sqrt(sqrt(dh)); w.mul_(scale*scale)  (factored_attention.py:83-88)
        double sc = 1.0 / sqrt(sqrt((double)L.dh));
        E.scale2 = (float)(sc * sc);
    }
    E.cols = (const ushort2*)(A + L.off_cols);
    E.soff = (const uint32_t*)(A + L.off_soff);
    p->d_cols = (ushort2*)(A + L.off_cols);
    p->d_goff = (uint32_t*)(A + L.off_goff);
    E.lrow0 = (const int*)(A + L.off_lrow);
    E.streams = A + L.off_streams; E.stream_stride = L.stream_stride;
    E.h = (__half*)(A + L.off_h); E.x1 = (__half*)(A + L.off_x1); E.qkv = (__half*)(A + L.off_qkv);
    E.a = (__half*)(A + L.off_a); E.g = (__half*)(A + L.off_g); E.part = (float*)(A + L.off_part);
    E.bar = (unsigned*)(A + L.off_sync); E.epoch = E.bar + 16; E.t = (int*)(E.bar + 32);
    size_t enc_off = L.off_small + L.small_per_layer * cfg->depth;
    for (int i = 0; i < 4; ++i) { p->bias_ptr[i].resize(cfg->depth); p->ln_ptr[i].resize(cfg->depth); }
    p->enc_w.assign(cfg->depth, nullptr); p->enc_b.assign(cfg->depth, nullptr);
    for (int l = 0; l < cfg->depth; ++l) {
        LayerDev& LD = E.layer[l];
        LD.attn_func = cfg->attn_func[l];
        LD.rows = L.cache_rows[l];
        LD.kc = (__half*)(A + L.cache_off[l]);
        LD.vc = (__half*)(A + L.cache_off[l] + L.cache_bytes[l]);
        float* s = (float*)(A + L.off_small + L.small_per_layer * l);
        p->bias_ptr[0][l] = s; s += 3 * cfg->n_state;
        p->bias_ptr[1][l] = s; s += cfg->width;
        p->bias_ptr[2][l] = s; s += cfg->mlp_width;
        p->bias_ptr[3][l] = s; s += cfg->width;
        for (int i = 0; i < 4; ++i) { p->ln_ptr[i][l] = s; s += cfg->width; }
        LD.b_qkv = p->bias_ptr[0][l]; LD.b_o = p->bias_ptr[1][l]; LD.b_1 = p->bias_ptr[2][l]; LD.b_2 = p->bias_ptr[3][l];
        LD.ln0_g = p->ln_ptr[0][l]; LD.ln0_b = p->ln_ptr[1][l]; LD.ln1_g = p->ln_ptr[2][l]; LD.ln1_b = p->ln_ptr[3][l];
        if (cfg->attn_func[l] == 6) {
            p->enc_w[l] = (__half*)(A + enc_off);
            p->enc_b[l] = (float*)(A + enc_off + (size_t)cfg->width * 2 * cfg->n_state * 2);
            enc_off = align_up(enc_off + (size_t)cfg->width * 2 * cfg->n_state * 2 + 2 * cfg->n_state * 4 + 512, 256);
            LD.enc_w = p->enc_w[l]; LD.enc_b = p->enc_b[l];
        }
    }
    p->dev = (EngineDev*)(A + L.off_dev);
    JK_CHECK_CUDA(cudaMemcpyAsync(A + L.off_cols, L.cols.data(), L.cols.size() * sizeof(ushort2), cudaMemcpyHostToDevice, stream));
    JK_CHECK_CUDA(cudaMemcpyAsync(A + L.off_soff, L.soff.data(), L.soff.size() * 4, cudaMemcpyHostToDevice, stream));
    JK_CHECK_CUDA(cudaMemcpyAsync(A + L.off_goff, L.goff.data(), L.goff.size() * 4, cudaMemcpyHostToDevice, stream));
    JK_CHECK_CUDA(cudaMemcpyAsync(A + L.off_lrow, L.lrow.data(), L.lrow.size() * 4, cudaMemcpyHostToDevice, stream));
    JK_CHECK_CUDA(cudaMemcpyAsync(p->dev, &p->host, sizeof(EngineDev), cudaMemcpyHostToDevice, stream));
    JK_CHECK_CUDA(cudaStreamSynchronize(stream));      // the host vectors above go out of scope
    JK_CHECK_CUDA(cudaFuncSetAttribute(jk_decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem_bytes));
    *out = p;
    return 0;
}

extern "C" int jk_prior_destroy(jk_prior* p) {
    delete p;
    return 0;
}

template <typename T>
static int pack_one(jk_prior* p, const void* w, int K, int N, int l, int gi, cudaStream_t stream) {
    pack_gemm_kernel<T><<<p->G, 256, 0, stream>>>((const T*)w, K, N, (uint8_t*)p->host.streams, p->host.stream_stride,
                                                   p->d_cols, p->d_goff, p->cfg.depth, l, gi);
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}
template <typename T>
static int round_bias(const void* src, float* dst, int n, cudaStream_t stream) {
    round_bias_kernel<T><<<(n + 255) / 256, 256, 0, stream>>>((const T*)src, dst, n);
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int jk_prior_load_layer(jk_prior* p, int l, const jk_layer_weights* w, jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(p && w, "null argument");
    JK_REQUIRE(l >= 0 && l < p->cfg.depth, "layer %d out of range", l);
    const jk_prior_config& c = p->cfg;
    const int af = c.attn_func[l];
    const int Ks[4] = {c.width, c.n_state, c.width, c.mlp_width};
    const int Ns[4] = {af == 6 ? c.n_state : 3 * c.n_state, c.width, c.mlp_width, c.width};
    const void* ws[4] = {w->c_attn_w, w->c_proj_w, w->fc_w, w->proj2_w};
    const void* bs[4] = {w->c_attn_b, w->c_proj_b, w->fc_b, w->proj2_b};
    for (int gi = 0; gi < 4; ++gi) {
        JK_REQUIRE(ws[gi] && bs[gi], "layer %d: missing weight %d", l, gi);
        int rc = w->w_dtype ? pack_one<__half>(p, ws[gi], Ks[gi], Ns[gi], l, gi, stream)
                            : pack_one<float>(p, ws[gi], Ks[gi], Ns[gi], l, gi, stream);
        if (rc) return rc;
        rc = w->b_dtype ? round_bias<__half>(bs[gi], p->bias_ptr[gi][l], Ns[gi], stream)
                        : round_bias<float>(bs[gi], p->bias_ptr[gi][l], Ns[gi], stream);
        if (rc) return rc;
    }
    const float* lns[4] = {w->ln0_g, w->ln0_b, w->ln1_g, w->ln1_b};
    for (int i = 0; i < 4; ++i) {
        JK_REQUIRE(lns[i], "layer %d: missing LayerNorm parameter %d", l, i);
        JK_CHECK_CUDA(cudaMemcpyAsync(p->ln_ptr[i][l], lns[i], (size_t)c.width * 4, cudaMemcpyDeviceToDevice, stream));
    }
    if (af == 6) {
        JK_REQUIRE(w->c_enc_kv_w && w->c_enc_kv_b, "layer %d: attn_func 6 needs c_enc_kv", l);
        size_t n = (size_t)c.width * 2 * c.n_state;
        if (w->w_dtype) to_half_kernel<__half><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const __half*)w->c_enc_kv_w, p->enc_w[l], n);
        else to_half_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const float*)w->c_enc_kv_w, p->enc_w[l], n);
        JK_CHECK_CUDA(cudaGetLastError());
        int rc = w->b_dtype ? round_bias<__half>(w->c_enc_kv_b, p->enc_b[l], 2 * c.n_state, stream)
                            : round_bias<float>(w->c_enc_kv_b, p->enc_b[l], 2 * c.n_state, stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int jk_prior_set_embeddings(jk_prior* p, const float* x_emb, const float* pos_emb, const float* x_out,
                                       const float* start_token) {
    JK_REQUIRE(p, "null engine");
    p->host.x_emb = x_emb; p->host.pos_emb = pos_emb; p->host.x_out = x_out; p->host.start_token = start_token;
    JK_CHECK_CUDA(cudaMemcpy(p->dev, &p->host, sizeof(EngineDev), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int jk_prior_reset(jk_prior* p, int t0, jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(p, "null engine");
    JK_REQUIRE(t0 >= 0 && t0 <= p->cfg.n_ctx, "t0 out of range");
    JK_CHECK_CUDA(cudaMemcpyAsync(p->host.t, &t0, sizeof(int), cudaMemcpyHostToDevice, stream));
    JK_CHECK_CUDA(cudaStreamSynchronize(stream));
    p->t_host = t0;
    return 0;
}

extern "C" int jk_prior_set_encoder_kv(jk_prior* p, const float* encoder_kv, int n, jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(p && encoder_kv, "null argument");
    const jk_prior_config& c = p->cfg;
    JK_REQUIRE(n >= 1 && n <= c.max_batch, "n_samples out of range");
    const int rows = n * c.encoder_dims;
    for (int l = 0; l < c.depth; ++l) {
        if (c.attn_func[l] != 6) continue;
        dim3 grid((2 * c.n_state + 31) / 32, (rows + 15) / 16), block(32, 16);
        enc_kv_kernel<<<grid, block, 0, stream>>>(encoder_kv, p->enc_w[l], p->enc_b[l], p->host.layer[l].kc,
                                                  p->host.layer[l].vc, rows, c.encoder_dims, c.width, c.n_state,
                                                  c.heads, p->host.dh, p->host.dh_pad);
        JK_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int jk_prior_step(jk_prior* p, const jk_step_args* a, jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(p && a, "null argument");
    JK_REQUIRE(a->n_samples >= 1 && a->n_samples <= p->cfg.max_batch, "n_samples %d out of range (max_batch %d)",
               a->n_samples, p->cfg.max_batch);
    JK_REQUIRE(a->x_in || a->tokens || p->t_host == 0, "tokens required for t > 0");
    JK_REQUIRE(a->x_in || (p->host.pos_emb && p->host.x_emb), "embeddings not set (jk_prior_set_embeddings)");
    JK_REQUIRE(!a->logits || p->host.x_out, "x_out not set");
    JK_REQUIRE(a->x_cond_len == 0 || a->x_cond_len == 1 || a->x_cond_len == p->cfg.n_ctx, "x_cond_len must be 1 or n_ctx");
    StepArgs A;
    A.n = a->n_samples; A.x_in = a->x_in; A.tokens = (const long long*)a->tokens; A.tok_stride = a->tok_stride;
    A.y_cond = a->y_cond; A.x_cond = a->x_cond; A.x_cond_len = a->x_cond_len ? a->x_cond_len : 1;
    A.h_out = a->h_out; A.logits = a->logits; A.logits_bstride = a->logits_bstride; A.logits_tstride = a->logits_tstride;
    const EngineDev* E = p->dev;
    void* args[2] = {(void*)&E, (void*)&A};
    JK_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)jk_decode_step_kernel, dim3(p->G), dim3(kThreads), args,
                                              (size_t)p->smem_bytes, stream));
    p->t_host += 1;
    return 0;
}

extern "C" int jk_prior_position(const jk_prior* p, int* t) {
    JK_REQUIRE(p && t, "null argument");
    *t = p->t_host;
    return 0;
}

extern "C" int jk_prior_debug_buffer(const jk_prior* p, int which, const void** ptr, size_t* n) {
    JK_REQUIRE(p && ptr && n, "null argument");
    const jk_prior_config& c = p->cfg;
    switch (which) {
        case 0: *ptr = p->host.h; *n = (size_t)16 * c.width; break;
        case 1: *ptr = p->host.qkv; *n = (size_t)16 * 3 * c.n_state; break;
        case 2: *ptr = p->host.a; *n = (size_t)16 * c.n_state; break;
        case 3: *ptr = p->host.x1; *n = (size_t)16 * c.width; break;
        case 4: *ptr = p->host.g; *n = (size_t)16 * c.mlp_width; break;
        default: JK_REQUIRE(false, "unknown buffer %d", which);
    }
    return 0;
}
