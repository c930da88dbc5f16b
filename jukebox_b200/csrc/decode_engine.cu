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
//     phases, so it keeps prefetching the next GEMMs / next layer (ring 6 x 16 KB per SM) while
//     the consumers sit in a grid barrier or in the attention phase.
//   * every Conv1D at decode is [16 x K] x [K x N]: M = 16 is exactly the m16n8k16 tensor-core
//     tile, so consumers use warp-level mma.sync with fp32 accumulation; each CTA owns 8-column
//     groups of N and the full K (no cross-CTA split-K, deterministic).  tcgen05 needs M >= 64
//     and would re-read a 4x zero-padded A tile from shared memory for every 8-16 weight
//     columns; it is used where tiles are >= 128 rows (chunked prefill: prefill.cu / prefill_gemm.cu).
//   * LayerNorm is fused into the GEMM's activation staging, bias / quick_gelu / residual adds
//     into its epilogue; fp16 rounding points follow the reference exactly.
//   * the layer-to-layer dependency is a grid barrier through one L2 counter.
//   * KV caches are laid out per attention pattern so that the rows a token attends are one
//     contiguous run (transpose-block layers store position p at row (p % bc)*blocks + p / bc).
//
// Numerics: activations fp16, accumulation fp32, LayerNorm/softmax fp32 - see oracle/transformer_np.py.
#include "engine.cuh"
#include <cooperative_groups.h>
#include <vector>
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include <math.h>

using namespace jk;

namespace {

constexpr int kConsumers = 256;
constexpr int kThreads = 384;          // 2 consumer warpgroups (8 warps) + 1 producer warpgroup (warp 8 works, 9-11 exit)
constexpr int kSlotBytes = 16384;
constexpr int kMaxSlots = 12;
constexpr int kHeaderBytes = 2048;     // barriers, LN statistics, shared copies of the descriptor / layer records
constexpr int kLogitKT = 1024;         // K tile (floats) of the fp32 logits product
constexpr int kLogitRowsPerChunk = 4;
constexpr int kLogitRowsPerPass = 8;
constexpr int kMaxSplit = 8;
constexpr int kProfSlots = 1024;

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

// The one dynamic shared-memory block of the decode kernel.  Every device function derives its
// pointers from this symbol (never from pointer parameters): that is what lets the compiler emit
// LDS/STS/ATOMS instead of generic LD/ST (measured: generic loads of the B fragments made the MMA loop
// 6x slower than the tensor pipe allows).
extern __shared__ __align__(1024) uint8_t jk_smem[];
__device__ __forceinline__ uint64_t* sm_full() { return reinterpret_cast<uint64_t*>(jk_smem); }
__device__ __forceinline__ uint64_t* sm_empty() { return reinterpret_cast<uint64_t*>(jk_smem) + kMaxSlots; }
__device__ __forceinline__ float* sm_stats() { return reinterpret_cast<float*>(jk_smem + 256); }
__device__ __forceinline__ uint8_t* sm_uni() { return jk_smem + kHeaderBytes; }
// The engine descriptor lives in global memory; with the shared-memory carve-out at its maximum there is
// no L1 to cache it, so every `E->field` was an L2 round trip (~300 cycles) on the dependency chain.
// The head of the descriptor (everything before the per-layer array) and the current / next layer
// records are therefore copied into shared memory once and read with LDS.
__device__ __forceinline__ const EngineDev* sm_E() { return reinterpret_cast<const EngineDev*>(jk_smem + 512); }
__device__ __forceinline__ const LayerDev* sm_layer(int i) { return reinterpret_cast<const LayerDev*>(jk_smem + 1024 + 256 * (i & 1)); }

struct Ring {
    int base_off;          // byte offset of slot 0 inside jk_smem
    int nslot;
    int slot;
    uint32_t phase;
    __device__ __forceinline__ uint64_t* full() const { return sm_full() + slot; }
    __device__ __forceinline__ uint64_t* empty() const { return sm_empty() + slot; }
    __device__ __forceinline__ uint8_t* data() const { return jk_smem + base_off + slot * kSlotBytes; }
    __device__ __forceinline__ void advance() {
        if (++slot == nslot) { slot = 0; phase ^= 1u; }
    }
};

#define STAMP(E_, slot_, i_)                                                                \
    do {                                                                                    \
        if (blockIdx.x == 0 && threadIdx.x == 0 && (E_)->prof_on && (slot_) < kProfSlots)      \
            (E_)->prof2[(size_t)(slot_) * 8 + (i_)] = clock64();                             \
    } while (0)

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ int kpc_of(int ncg) {
    int k = (64 / ncg) & ~7;
    return k < 8 ? 8 : k;
}

// Grid-wide barrier through one L2 counter: release-add by one thread after a CTA barrier, acquire
// polling, CTA barrier.  The CTA barriers make the pattern cumulative for the whole block, no separate
// membar is needed.  Measured on B200 (tools/micro/ubench.cu): 2320 cycles = 1.18 us for 148 CTAs; a
// two-level variant (group counters + top counter) measured 1.8-2.3 us in situ (two dependent
// release/acquire round trips), so the flat counter stays.  `k` = global index of this barrier.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned k, int cta, int G) {
    consumer_sync();
    if (threadIdx.x == 0) {
        red_release_add(bar, 1u);
        const unsigned target = k * (unsigned)G;
        unsigned spins = 0;
        while ((int)(ld_acquire_u32(bar) - target) < 0) {
            if (++spins > (1u << 28)) __trap();
        }
    }
    consumer_sync();
}

// The attention -> proj boundary is not a full barrier: proj needs every (sample, head) output, and each of
// those is completed by exactly one CTA (the unsplit item's owner or the merger of its parts).  Completers add
// to a counter (release), everybody polls it (acquire): B*H arrivals instead of G on the critical path.
__device__ __forceinline__ unsigned* attn_done_counter(const EngineDev* E) { return E->bar + 64; }
__device__ __forceinline__ void attn_done_signal(const EngineDev* E) {      // after a consumer_sync that follows the writes of `a`
    if (threadIdx.x == 0) red_release_add(attn_done_counter(E), 1u);
}
__device__ __forceinline__ void attn_done_wait(const EngineDev* E, unsigned target) {
    consumer_sync();
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while ((int)(ld_acquire_u32(attn_done_counter(E)) - target) < 0) {
            if (++spins > (1u << 28)) __trap();
        }
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
// LayerNorm statistics travel with the activations: whoever WRITES a row block of the residual
// stream also adds sum(x) and sum(x^2) of its columns into per-row 64-bit fixed-point accumulators
// (exact integer adds => order independent => bit-reproducible), so the consuming GEMM can
// normalise while it stages - no extra passes over the row, no extra grid barrier.
//   sum  : x * 2^24 is an exact integer for every fp16 value
//   sumsq: x^2 is exact in fp32; scaled by 2^16 and rounded per element (a pure function of x)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long long fx_sum(float x) { return __float2ll_rn(x * 16777216.0f); }
__device__ __forceinline__ long long fx_sq(float x) { return __float2ll_rn(x * x * 65536.0f); }
__device__ __forceinline__ void red_add_s64(long long* p, long long v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}

// activation staging: global fp16 [16][K] -> shared fp16 [16][K+8] (ldmatrix friendly), optionally
// through LayerNorm (fp32 math, eps 1e-5; reference transformer/ops.py:14-24).  Each thread owns 8-column
// vectors: 8 independent 16-byte loads in flight per batch, raw store, then (LN) a rolled in-place pass
// over its own vectors.  Loops are deliberately NOT unrolled beyond that: the whole per-layer code must
// stay inside the 32 KB instruction cache - an earlier fully unrolled build measured ~2.5 us of
// instruction-fetch stalls in EVERY phase (profiles/ phase_profile_r01d.txt).
__device__ __noinline__ void stage_acts(const __half* in, int K, int B, int ln,
                                        const float* gamma, const float* beta, const long long* lnacc) {
    const int tid = threadIdx.x;
    uint8_t* acts = sm_uni();
    float* stats = sm_stats();
    const int nvec = K >> 3;
    const int astride = (K + 8) * 2;
    if (ln && tid < 16) {
        float mean = 0.f, rstd = 0.f;
        if (tid < B) {
            // double only for the cancellation in E[x^2] - mean^2 (adds / muls; no double div or sqrt:
            // those are kilobytes of library code in the instruction cache)
            const double rk = (double)(1.0f / (float)K);    // K is a multiple of 16: exact for powers of two, 1e-7 rel otherwise
            const double m = (double)__ldcg(lnacc + 16 * (2 * tid)) * (1.0 / 16777216.0) * rk;
            double var = (double)__ldcg(lnacc + 16 * (2 * tid + 1)) * (1.0 / 65536.0) * rk - m * m;
            var = var < 0.0 ? 0.0 : var;
            rstd = 1.0f / sqrtf((float)var + 1e-5f);
            mean = -(float)m * rstd;                        // staged as x * rstd + (-mean * rstd), then * gamma + beta
        }
        stats[2 * tid] = mean;
        stats[2 * tid + 1] = rstd;
    }
#pragma unroll 1
    for (int v0 = 0; v0 < nvec; v0 += kConsumers) {          // uniform trip count: the barrier below is CTA-wide
        const int v = v0 + tid;
        const bool act = v < nvec;
        uint4 x[16];                                          // 16 independent 16-byte loads in flight
        float gm[8], bt[8];
        if (act) {
            if (ln) {      // issued first: their latency overlaps the 16-load batch below
                *reinterpret_cast<float4*>(gm) = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
                *reinterpret_cast<float4*>(gm + 4) = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
                *reinterpret_cast<float4*>(bt) = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
                *reinterpret_cast<float4*>(bt + 4) = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
            }
#pragma unroll
            for (int r = 0; r < 16; ++r)
                x[r] = (r < B) ? ldcg_u4(in + (size_t)r * K + v * 8) : make_uint4(0, 0, 0, 0);
        }
        if (ln && v0 == 0) consumer_sync();                  // row statistics are in shared memory (loads in flight)
        if (act) {
            if (ln) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    __half2* hp = reinterpret_cast<__half2*>(&x[r]);
                    const float rstd = stats[2 * r + 1], nmr = stats[2 * r];     // nmr = -mean * rstd
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float2 f = __half22float2(hp[e]);
                        f.x = fmaf(fmaf(f.x, rstd, nmr), gm[2 * e], bt[2 * e]);
                        f.y = fmaf(fmaf(f.y, rstd, nmr), gm[2 * e + 1], bt[2 * e + 1]);
                        hp[e] = __floats2half2_rn(f.x, f.y);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) *reinterpret_cast<uint4*>(acts + r * astride + v * 16) = x[r];
        }
    }
}

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// one ring slot worth of k-steps for this warp.  NCG (8-column groups of this CTA) is a template
// parameter: with a run-time count the compiler serialised every LDS -> HMMA pair through one
// register pair (tools/micro/ubench.cu: 4070 vs 1170 cycles for the K = 2048 loop).
template <int NCG>
__device__ __forceinline__ void mma_chunk(float (&acc)[8][4], uint32_t arow, uint32_t sl, int kk0, int nk, int warp) {
#pragma unroll 4
    for (int i = warp; i < nk; i += 8) {
        uint32_t a[4];
        ldsm4(a, arow + (kk0 + i) * 32);
        uint2 b[NCG];
#pragma unroll
        for (int j = 0; j < NCG; ++j) b[j] = lds64(sl + ((i * NCG + j) << 8));
#pragma unroll
        for (int j = 0; j < NCG; ++j) mma_16816(acc[j], a, b[j].x, b[j].y);
    }
}

// ---------------------------------------------------------------------------------------
// one Conv1D at decode: out[b, cols of this CTA] = epilogue( acts[16,K] . Wslice[K, 8*ncg] )
// ---------------------------------------------------------------------------------------
enum { EPI_QKV = 0, EPI_PROJ = 1, EPI_FC = 2, EPI_PROJ2 = 3 };

struct GemmArgs {
    const __half* in;
    int K, N, g0, ncg, ln, epi, pslot;
    const float *gamma, *beta, *bias;
    const long long* ln_in;
    long long* ln_out;
};

__device__ __noinline__ Ring gemm_phase(Ring ring, int B, const GemmArgs g) {   // by value: registers, not local memory
    const EngineDev* E = sm_E();
    uint8_t* uni = sm_uni();
    if (g.ncg == 0) return ring;
    const int ncg = g.ncg;
    STAMP(E, g.pslot, 0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = g.K, N = g.N, epi = g.epi;
    const int nc = ncg * 8;
    const bool residual = (epi == EPI_PROJ || epi == EPI_PROJ2);
    const __half* res_src = (epi == EPI_PROJ) ? E->h : E->x1;
    __half* res_dst = (epi == EPI_PROJ) ? E->x1 : E->h;
    // epilogue operands of this thread's first output element: issue the loads now, use them at the end
    const bool has_e = tid < B * nc;
    const int eb = has_e ? tid / nc : 0, ecc = has_e ? tid - eb * nc : 0;
    float pre_bias = 0.f, pre_res = 0.f;
    if (has_e) {
        pre_bias = g.bias[g.g0 * 8 + ecc];
        if (residual) pre_res = ld_half_cg(res_src + (size_t)eb * N + g.g0 * 8 + ecc);
    }
    stage_acts(g.in, K, B, g.ln, g.gamma, g.beta, g.ln_in);
    consumer_sync();
    STAMP(E, g.pslot, 1);

    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
    const int nkk = K >> 4;
    const int kpc = kpc_of(ncg);
    const int astride = (K + 8) * 2;
    const uint32_t arow = smem_u32(uni + (lane & 15) * astride + (lane >> 4) * 16);
#define JK_MMA_LOOP(NCG)                                                                      \
    _Pragma("unroll 1") for (int kk0 = 0; kk0 < nkk; kk0 += kpc) {                            \
        const int nk = min(kpc, nkk - kk0);                                                    \
        mbar_wait(ring.full(), ring.phase);                                                    \
        mma_chunk<NCG>(acc, arow, smem_u32(ring.data()) + lane * 8, kk0, nk, warp);            \
        __syncwarp();                                                                          \
        if (lane == 0) mbar_arrive(ring.empty());                                              \
        ring.advance();                                                                        \
    }
    switch (ncg) {
        case 1: { JK_MMA_LOOP(1) } break;
        case 2: { JK_MMA_LOOP(2) } break;
        case 3: { JK_MMA_LOOP(3) } break;
        case 4: { JK_MMA_LOOP(4) } break;
        case 5: { JK_MMA_LOOP(5) } break;
        case 6: { JK_MMA_LOOP(6) } break;
        case 7: { JK_MMA_LOOP(7) } break;
        default: { JK_MMA_LOOP(8) } break;
    }
#undef JK_MMA_LOOP
    STAMP(E, g.pslot, 2);
    consumer_sync();                       // everyone is done reading the staged activations
    float* red = reinterpret_cast<float*>(uni);   // [8 warps][ncg][16][8]  (<= 32 KB)
    float* ov = reinterpret_cast<float*>(uni + 32768);   // [16][64] residual-stream outputs of this CTA
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
#pragma unroll 1
    for (int e = tid; e < B * nc; e += kConsumers) {
        const int b = e / nc, cc = e - b * nc;
        const int j = cc >> 3, col = cc & 7;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[((w * ncg + j) * 16 + b) * 8 + col];
        const int gc = g.g0 * 8 + cc;
        const bool first = (e == tid);
        const float y = h2f_round(s + (first ? pre_bias : g.bias[gc]));     // Conv1D output, rounded once to fp16
        if (epi == EPI_QKV) {
            E->qkv[(size_t)b * N + gc] = __float2half_rn(y);
        } else if (epi == EPI_FC) {                        // quick_gelu (transformer/ops.py:33-35)
            E->g[(size_t)b * N + gc] = __float2half_rn(quick_gelu_f(y));
        } else {
            // EPI_PROJ : x1 = fp16(h + a)      EPI_PROJ2 : h = fp16(x1 + m)   (transformer.py:82-83)
            const float base = first ? pre_res : ld_half_cg(res_src + (size_t)b * N + gc);
            const float o = h2f_round(base + y);
            res_dst[(size_t)b * N + gc] = __float2half_rn(o);
            ov[b * 64 + cc] = o;
        }
    }
    STAMP(E, g.pslot, 3);
    consumer_sync();                       // red region is reused by the next phase's staging
    if (residual && g.ln_out) {
        // statistics for the LayerNorm that will read these rows: warp w reduces rows 2w, 2w+1 of this
        // CTA's columns in a fixed order (exact integer sums), one 64-bit red per row and moment
#pragma unroll 1
        for (int r = 2 * warp; r < 2 * warp + 2 && r < B; ++r) {
            long long s1 = 0, s2 = 0;
            for (int cc = lane; cc < nc; cc += 32) { const float o = ov[r * 64 + cc]; s1 += fx_sum(o); s2 += fx_sq(o); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            if (lane == 0) { red_add_s64(g.ln_out + 16 * (2 * r), s1); red_add_s64(g.ln_out + 16 * (2 * r + 1), s2); }
        }
        consumer_sync();                   // ov is reused
    }
    return ring;
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

// rows of one shared-memory K (or V) tile: a multiple of 16 (the MMA row block), ~24 KB per tile
__device__ __host__ __forceinline__ int attn_tile_rows(int dhp) {
    int r = (12288 / dhp) & ~15;
    return r < 16 ? 16 : (r > 64 ? 64 : r);
}

// how many CTAs share one (sample, head): as few as keep every part inside ONE shared-memory tile
// (RC - 1 cached rows + the current token's row), bounded by the grid
__device__ __forceinline__ int attn_nsplit(const EngineDev* E, int B, int ncache) {
    const int cap = attn_tile_rows(E->dh_pad) - 1;
    int ns = (ncache + cap - 1) / cap;
    ns = min(ns, E->G / (B * E->H));
    return max(1, min(kMaxSplit, ns));
}

__device__ __forceinline__ void ldsm4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// K/V tiles in shared memory: row r = dhp halves, its 16-byte chunks XOR-swizzled with (r & 7) so that
// ldmatrix (8 rows, same chunk) touches 8 different bank groups.  Only when a row has a multiple of 8
// chunks; other head sizes stay linear (correct, bank-conflicted).
__device__ __forceinline__ uint32_t kv_chunk_off(int r, int chunk, int dhp, int swz) {
    return (uint32_t)(r * dhp * 2 + ((chunk ^ (r & swz)) << 4));
}

// cp.async of rows [0, nr) of K and V into swizzled tiles.  Thread -> (first row, chunk) is fixed, so the
// loop body is two cp.async and two adds: no per-chunk division (the generic i / nvec form cost 1.4 us of
// issue time per 47-row tile).
__device__ __forceinline__ void kv_copy_tile(uint32_t kd, uint32_t vd, const __half* ks, const __half* vs, int nr,
                                             int dhp, int swz) {
    const int nvec = dhp >> 3, tid = threadIdx.x;
    if (nvec <= kConsumers && kConsumers % nvec == 0) {
        const int rstep = kConsumers / nvec, c = tid % nvec;
#pragma unroll 2
        for (int r = tid / nvec; r < nr; r += rstep) {
            const uint32_t o = kv_chunk_off(r, c, dhp, swz);
            const size_t g = (size_t)r * dhp + c * 8;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(kd + o), "l"(ks + g));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vd + o), "l"(vs + g));
        }
    } else {
        for (int i = tid; i < nr * nvec; i += kConsumers) {
            const int r = i / nvec, c = i - r * nvec;
            const uint32_t o = kv_chunk_off(r, c, dhp, swz);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(kd + o), "l"(ks + i * 8));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vd + o), "l"(vs + i * 8));
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __noinline__ void attn_scores(uint32_t kt, int dhp, int swz, int nrb, int nr, uint32_t qh_s, float* sc, float scale2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npair = dhp >> 4;
        if (warp < nrb) {
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
            const int mi = lane >> 3;
            const int arow_i = warp * 16 + (lane & 7) + ((mi & 1) << 3);
            const uint32_t arow = kt + arow_i * dhp * 2;
            const int axor = arow_i & swz, ahi = mi >> 1;
#pragma unroll 2
            for (int ks = 0; ks < npair; ++ks) {
                uint32_t a[4];
                ldsm4(a, arow + (((2 * ks + ahi) ^ axor) << 4));
                uint32_t b0 = 0u, b1 = 0u;
                if (lane < 4) {
                    asm("ld.shared.u32 %0, [%1];" : "=r"(b0) : "r"(qh_s + (ks * 16 + 2 * lane) * 2));
                    asm("ld.shared.u32 %0, [%1];" : "=r"(b1) : "r"(qh_s + (ks * 16 + 8 + 2 * lane) * 2));
                }
                if (ks & 1) mma_16816(c1, a, b0, b1); else mma_16816(c0, a, b0, b1);
            }
            if ((lane & 3) == 0) {
                const int rlo = warp * 16 + (lane >> 2), rhi = rlo + 8;
                sc[rlo] = (rlo < nr) ? h2f_round(h2f_round(c0[0] + c1[0]) * scale2) : -INFINITY;
                sc[rhi] = (rhi < nr) ? h2f_round(h2f_round(c0[2] + c1[2]) * scale2) : -INFINITY;
            }
        }
}

// flash-decoding merge of the ns partials of one (sample, head), run by the CTA that finished last.
// Own function: its registers must not add to attn_item's (see "Register regime" in DESIGN.md).
__device__ __noinline__ void attn_merge(int item, int ns, int b, int h) {
    const EngineDev* E = sm_E();
    const int tid = threadIdx.x, dh = E->dh, dhp = E->dh_pad, S = E->S;
        const float* p0 = E->part + ((size_t)(item * kMaxSplit)) * (dhp + 2);
        const int st = dhp + 2;
        if (ns <= 4) {
            // every load of the merge is issued before the first use: ONE L2 round trip instead of three
            // dependent ones (max pass, sum pass, value pass) on the critical path of the slowest CTAs
            float m[4], l[4], v0[4], v1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool on = q < ns;
                m[q] = on ? __ldcg(p0 + (size_t)q * st) : -INFINITY;
                l[q] = on ? __ldcg(p0 + (size_t)q * st + 1) : 0.f;
                v0[q] = (on && tid < dh) ? __ldcg(p0 + (size_t)q * st + 2 + tid) : 0.f;
                v1[q] = (on && tid + kConsumers < dh) ? __ldcg(p0 + (size_t)q * st + 2 + tid + kConsumers) : 0.f;
            }
            float M = -INFINITY;
#pragma unroll
            for (int q = 0; q < 4; ++q) M = fmaxf(M, m[q]);
            float Lsum = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q < ns) {
                    const float w = expf(m[q] - M);
                    Lsum += l[q] * w; o0 += v0[q] * w; o1 += v1[q] * w;
                }
            }
            if (tid < dh) E->a[(size_t)b * S + h * dh + tid] = __float2half_rn(o0 / Lsum);
            if (tid + kConsumers < dh) E->a[(size_t)b * S + h * dh + tid + kConsumers] = __float2half_rn(o1 / Lsum);
        } else {
            float M = -INFINITY;
            for (int q = 0; q < ns; ++q) M = fmaxf(M, __ldcg(p0 + (size_t)q * st));
            float Lsum = 0.f;
            for (int q = 0; q < ns; ++q) Lsum += __ldcg(p0 + (size_t)q * st + 1) * expf(__ldcg(p0 + (size_t)q * st) - M);
            for (int d = tid; d < dh; d += kConsumers) {
                float o = 0.f;
                for (int q = 0; q < ns; ++q) o += __ldcg(p0 + (size_t)q * st + 2 + d) * expf(__ldcg(p0 + (size_t)q * st) - M);
                E->a[(size_t)b * S + h * dh + d] = __float2half_rn(o / Lsum);
            }
        }
}

// One (sample, head, part) work item; q_len == 1 (reference factored_attention.py:82-133 and the
// per-pattern sample branches :135-228).
//   * the part's K and V rows are staged with cp.async into swizzled tiles; a part that fits one tile
//     (RC-1 cached rows + the current token) is ONE tile - and may have been prefetched before the QKV
//     GEMM - longer parts (dense / prime / enc-dec layers) run double-buffered through both tile regions
//   * scores on the tensor cores: A = 16 key rows x 16 dims (ldmatrix), B = q in column 0, fp32
//     accumulate; s = fp16(fp16(q.k) * dh^-1/2) exactly as the reference rounds it
//   * softmax is flash-style in fp32 (running max / sum), every warp redundantly; P rounded to fp16 as
//     the reference's w.half(), P.V on the tensor cores (A = P in row 0, B = V via ldmatrix.trans), each
//     warp owning 16-dim output slices
//   * parts of one (sample, head) are merged by the last CTA to finish (atomic ticket), so the phase
//     needs no extra grid barrier
__device__ __noinline__ void attn_item(const LayerDev& LD_ref, int b, int h, int s, int ns,
                                       const AttnGeom G, int pslot, int pre) {
    const EngineDev* E = sm_E();
    uint8_t* uni = sm_uni();
    float* stats = sm_stats();
    const LayerDev LD = LD_ref;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dh = E->dh, dhp = E->dh_pad, S = E->S;
    const int nvec = dhp >> 3, npair = dhp >> 4;
    const int swz = (nvec & 7) ? 0 : 7;
    const int RC = attn_tile_rows(dhp);
    const int tileB = RC * dhp * 2;
    const uint32_t regA = smem_u32(uni), regB = smem_u32(jk_smem + kHeaderBytes + E->uni_bytes);
    __half* qh = reinterpret_cast<__half*>(uni + 2 * tileB);          // [dhp]
    float* sc = reinterpret_cast<float*>(uni + 2 * tileB + dhp * 2);  // [64] scores of the tile
    const int qkv_stride = (LD.attn_func == 6) ? S : 3 * S;
    const __half* qrow = E->qkv + (size_t)b * qkv_stride + h * dh;
    const size_t cbase = ((size_t)(b * E->H + h)) * LD.rows;
    const bool last_part = (s == ns - 1);
    const int R = G.R;

    if (R == 0) {   // prev-block attention inside the first block: keys/values are zeros -> output 0
        for (int d = tid; d < dh; d += kConsumers) {
            E->a[(size_t)b * S + h * dh + d] = __float2half_rn(0.f);
            if (G.wrow >= 0) {
                LD.kc[(cbase + G.wrow) * dhp + d] = __float2half_rn(ld_half_cg(qrow + S + d));
                LD.vc[(cbase + G.wrow) * dhp + d] = __float2half_rn(ld_half_cg(qrow + 2 * S + d));
            }
        }
        consumer_sync();
        attn_done_signal(E);
        return;
    }
    STAMP(E, pslot, 0);
    const int ncache = R - (G.cur ? 1 : 0);               // rows that come from the cache
    const int i0 = (int)(((long long)ncache * s) / ns), i1 = (int)(((long long)ncache * (s + 1)) / ns);
    const __half* kbase = LD.kc + (cbase + G.base) * dhp;
    const __half* vbase = LD.vc + (cbase + G.base) * dhp;
    const int trows = RC - 1;
    const int ntiles = max(1, (i1 - i0 + trows - 1) / trows);

    auto issue_tile = [&](int ti) {
        const int r0 = i0 + ti * trows, nr = max(0, min(trows, i1 - r0));
        const uint32_t kd = (ti & 1) ? regB : regA, vd = kd + tileB;
        kv_copy_tile(kd, vd, kbase + (size_t)r0 * dhp, vbase + (size_t)r0 * dhp, nr, dhp, swz);
    };
    const bool use_pre = pre && ntiles == 1;
    if (!use_pre) issue_tile(0);
    for (int d = tid; d < dhp; d += kConsumers) qh[d] = (d < dh) ? __float2half_rn(ld_half_cg(qrow + d)) : __float2half_rn(0.f);
    if (!G.cur && G.wrow >= 0 && last_part) {   // patterns that do not attend the current token still cache it
        for (int d = tid; d < dh; d += kConsumers) {
            LD.kc[(cbase + G.wrow) * dhp + d] = __float2half_rn(ld_half_cg(qrow + S + d));
            LD.vc[(cbase + G.wrow) * dhp + d] = __float2half_rn(ld_half_cg(qrow + 2 * S + d));
        }
    }

    float m_run = -INFINITY, l_run = 0.f;
    float* osm = sc + 64;                                  // [dhp] running output of multi-tile parts (owner-private)
    const uint32_t qh_s = smem_u32(qh);
#pragma unroll 1
    for (int ti = 0; ti < ntiles; ++ti) {
        if (ti + 1 < ntiles) {
            issue_tile(ti + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        const int r0 = i0 + ti * trows;
        int nr = max(0, min(trows, i1 - r0));
        const uint32_t kt = (use_pre || (ti & 1)) ? regB : regA, vt = kt + tileB;
        if (ti == ntiles - 1 && last_part && G.cur) {       // append the current token's k, v (from the QKV GEMM)
            for (int d = tid; d < dhp; d += kConsumers) {
                __half kh = __float2half_rn(0.f), vh = kh;
                if (d < dh) {
                    kh = __float2half_rn(ld_half_cg(qrow + S + d));
                    vh = __float2half_rn(ld_half_cg(qrow + 2 * S + d));
                    if (G.wrow >= 0) {
                        LD.kc[(cbase + G.wrow) * dhp + d] = kh;
                        LD.vc[(cbase + G.wrow) * dhp + d] = vh;
                    }
                }
                const uint32_t o = kv_chunk_off(nr, d >> 3, dhp, swz) + (d & 7) * 2;
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(kt + o), "h"(__half_as_ushort(kh)));
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(vt + o), "h"(__half_as_ushort(vh)));
            }
            nr += 1;
        }
        const int nrb = (nr + 15) >> 4;                     // 16-row blocks of this tile
        // V rows past nr inside the last row block meet P = 0 in the MMA: they must be finite
        for (int i = tid; i < (nrb * 16 - nr) * nvec; i += kConsumers) {
            const int r = nr + i / nvec, c = i % nvec;
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(vt + kv_chunk_off(r, c, dhp, swz)), "r"(0u));
        }
        consumer_sync();
        STAMP(E, pslot, 1);
        // ---- scores: warp w < nrb owns key rows [16w, 16w+16) ----------------------------------------
        attn_scores(kt, dhp, swz, nrb, nr, qh_s, sc, E->scale2);
        consumer_sync();
        STAMP(E, pslot, 2);
        // ---- softmax of the tile (every warp, redundantly): lane r holds rows r and r + 32 -------------
        const float s0 = (lane < nrb * 16) ? sc[lane] : -INFINITY;
        const float s1 = (lane + 32 < nrb * 16) ? sc[lane + 32] : -INFINITY;
        const float m_new = fmaxf(m_run, warp_max(fmaxf(s0, s1)));
        const float corr = expf(m_run - m_new);              // exp(-inf) = 0 on the first tile
        const float p0 = expf(s0 - m_new), p1 = expf(s1 - m_new);
        l_run = l_run * corr + warp_sum(p0 + p1);
        m_run = m_new;
        // ---- P.V: this warp's 16-dim output slices, all row blocks.  A = P (fp16) in row 0 --------------
        {
            uint32_t pa0[4], pa2[4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const float src = (ks < 2) ? p0 : p1;
                const int l0 = (ks & 1) * 16 + 2 * (lane & 3);
                const float pa = __shfl_sync(0xffffffffu, src, l0), pb = __shfl_sync(0xffffffffu, src, l0 + 1);
                const float pc = __shfl_sync(0xffffffffu, src, l0 + 8), pd = __shfl_sync(0xffffffffu, src, l0 + 9);
                const __half2 h0 = __floats2half2_rn(pa, pb), h2 = __floats2half2_rn(pc, pd);
                pa0[ks] = (lane < 4) ? *reinterpret_cast<const uint32_t*>(&h0) : 0u;
                pa2[ks] = (lane < 4) ? *reinterpret_cast<const uint32_t*>(&h2) : 0u;
            }
            const int mi = lane >> 3;
            const int vrow_l = (lane & 7) + ((mi & 1) << 3), vhi = mi >> 1;
            const bool last_tile = (ti == ntiles - 1);
#pragma unroll 1
            for (int np = warp; np < npair; np += 8) {
                float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (ks < nrb) {
                        const int vrow_i = ks * 16 + vrow_l;
                        uint32_t bf[4];
                        ldsm4_trans(bf, vt + vrow_i * dhp * 2 + (((2 * np + vhi) ^ (vrow_i & swz)) << 4));
                        const uint32_t a[4] = {pa0[ks], 0u, pa2[ks], 0u};
                        mma_16816(o0, a, bf[0], bf[1]);
                        mma_16816(o1, a, bf[2], bf[3]);
                    }
                }
                if (lane < 4) {      // row 0: dims np*16 + {0, 8} + 2*lane, +1
                    const int d = np * 16 + 2 * lane;
                    float2 r0 = make_float2(o0[0], o0[1]), r1 = make_float2(o1[0], o1[1]);
                    if (ti > 0) {
                        const float2 q0 = *reinterpret_cast<const float2*>(osm + d), q1 = *reinterpret_cast<const float2*>(osm + d + 8);
                        r0.x += q0.x * corr; r0.y += q0.y * corr; r1.x += q1.x * corr; r1.y += q1.y * corr;
                    }
                    if (!last_tile) {
                        *reinterpret_cast<float2*>(osm + d) = r0;
                        *reinterpret_cast<float2*>(osm + d + 8) = r1;
                    } else if (ns == 1) {
                        const float inv = 1.f / l_run;
                        __half* ao = E->a + (size_t)b * S + h * dh;
                        if (d < dh) ao[d] = __float2half_rn(r0.x * inv);
                        if (d + 1 < dh) ao[d + 1] = __float2half_rn(r0.y * inv);
                        if (d + 8 < dh) ao[d + 8] = __float2half_rn(r1.x * inv);
                        if (d + 9 < dh) ao[d + 9] = __float2half_rn(r1.y * inv);
                    } else {
                        float* part = E->part + ((size_t)((b * E->H + h) * kMaxSplit + s)) * (dhp + 2) + 2;
                        *reinterpret_cast<float2*>(part + d) = r0;
                        *reinterpret_cast<float2*>(part + d + 8) = r1;
                    }
                }
            }
        }
        if (ti + 1 < ntiles) consumer_sync();                 // tile buffers and sc are reused
    }
    STAMP(E, pslot, 3);
    if (ns == 1) {
        consumer_sync();
        attn_done_signal(E);
        return;
    }
    // ---- split parts: the partial is published, the last finisher merges (flash-decoding merge) ------
    const int item = b * E->H + h;
    if (tid == 0) {
        float* part = E->part + ((size_t)(item * kMaxSplit + s)) * (dhp + 2);
        part[0] = m_run; part[1] = l_run;
    }
    consumer_sync();
    if (tid == 0) {      // acq_rel ticket: publishes this CTA's partial, acquires the others' for the merger
        unsigned ticket;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(E->acnt + item) : "memory");
        stats[48] = (ticket == (unsigned)(ns - 1)) ? 1.f : 0.f;
        if (ticket == (unsigned)(ns - 1)) E->acnt[item] = 0u;
    }
    consumer_sync();
    if (stats[48] != 0.f) attn_merge(item, ns, b, h);
    consumer_sync();
    if (stats[48] != 0.f) attn_done_signal(E);
    STAMP(E, pslot, 6);
}

// Prefetch of the cached K/V rows of this CTA's first attention work item, issued BEFORE the QKV GEMM of
// the layer: cached rows do not depend on the current token, so their HBM latency hides behind the
// whole QKV phase and its barrier.  Only single-tile parts are prefetched; returns 1 if so.
__device__ __noinline__ int attn_prefetch(const LayerDev& LD_ref, int B, int c, int t) {
    const EngineDev* E = sm_E();
    if (!E->kv_prefetch) return 0;
    const LayerDev LD = LD_ref;
    const AttnGeom G = attn_geom(E, LD, t);
    if (G.R == 0) return 0;
    const int ncache = G.R - (G.cur ? 1 : 0);
    const int ns = attn_nsplit(E, B, ncache);
    if (c >= B * E->H * ns) return 0;
    const int s = c % ns, bh = c / ns, b = bh / E->H, h = bh % E->H;
    const int dhp = E->dh_pad, nvec = dhp >> 3, RC = attn_tile_rows(dhp);
    const int swz = (nvec & 7) ? 0 : 7;
    const int i0 = (int)(((long long)ncache * s) / ns), i1 = (int)(((long long)ncache * (s + 1)) / ns);
    if (i1 - i0 > RC - 1) return 0;
    const size_t cbase = ((size_t)(b * E->H + h)) * LD.rows;
    const uint32_t kd = smem_u32(jk_smem + kHeaderBytes + E->uni_bytes), vd = kd + RC * dhp * 2;
    kv_copy_tile(kd, vd, LD.kc + (cbase + G.base + i0) * dhp, LD.vc + (cbase + G.base + i0) * dhp, i1 - i0, dhp, swz);
    return 1;
}

// ---------------------------------------------------------------------------------------
// producer warp: walks this CTA's weight stream (and the logits rows) in consumption order
// ---------------------------------------------------------------------------------------
__device__ __noinline__ void producer_loop(const EngineDev* E, Ring ring, bool do_logits, int c) {
    if ((threadIdx.x & 31) != 0) return;
    const uint8_t* src = E->streams + (size_t)c * E->stream_stride + (size_t)E->soff[(size_t)c * (E->depth + 1)] * 16;
    for (int l = 0; l < E->depth; ++l) {
        const ushort2* cl = E->cols + ((size_t)c * E->depth + l) * 4;
        const int Ks[4] = {E->W, E->S, E->W, E->M};
        if (c == (l % E->G)) {
            // biases + LayerNorm parameters of this layer (one contiguous block, ~57 KB for 1b_lyrics) are
            // shared by every CTA and evicted from L2 between steps: pull them into L2 ahead of the consumers
            // (this producer runs about a layer ahead of them)
            const char* p0 = reinterpret_cast<const char*>(E->layer[l].b_qkv);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"((uint32_t)E->small_bytes) : "memory");
        }
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
            const int ncg = cl[gi].y;
            if (ncg == 0) continue;
            const int nkk = Ks[gi] >> 4, kpc = kpc_of(ncg);
            for (int kk0 = 0; kk0 < nkk; kk0 += kpc) {
                const int nk = min(kpc, nkk - kk0);
                const uint32_t bytes = (uint32_t)nk * ncg * 256u;
                mbar_wait(ring.empty(), ring.phase ^ 1u);
                mbar_expect_tx(ring.full(), bytes);
                tma_bulk_g2s(ring.data(), src, bytes, ring.full());
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
                    mbar_wait(ring.empty(), ring.phase ^ 1u);
                    mbar_expect_tx(ring.full(), (uint32_t)(nr * kt * 4));
                    for (int i = 0; i < nr; ++i)
                        tma_bulk_g2s(ring.data() + i * kt * 4, E->x_out + (size_t)(r + i) * W + k0,
                                     (uint32_t)(kt * 4), ring.full());
                    ring.advance();
                }
            }
        }
    }
}

// fp32 logits: logits[b, r] = sum_k y[b, k] * x_out[r, k],  y = float(h) (+ cond)
// (reference autoregressive.py:226-229: fp32 nn.Linear on the fp32 transformer output)
__device__ __noinline__ void logits_phase(const StepArgs& A_ref, Ring& ring_ref, int c, int t) {
    const EngineDev* E = sm_E();
    uint8_t* uni = sm_uni();
    const StepArgs A = A_ref;
    Ring ring = ring_ref;
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
            {   // y = float(h) (+ cond): 8 halves per thread-iteration, loads batched 4 deep
                const int nv = kt >> 3;
                for (int idx0 = tid; idx0 < 16 * nv; idx0 += 4 * kConsumers) {
                    uint4 hv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int idx = idx0 + u * kConsumers;
                        const int b = idx / nv, v = idx - b * nv;
                        hv[u] = (idx < 16 * nv && b < B) ? ldcg_u4(E->h + (size_t)b * W + k0 + v * 8) : make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int idx = idx0 + u * kConsumers;
                        if (idx >= 16 * nv) continue;
                        const int b = idx / nv, v = idx - b * nv;
                        const __half2* hp = reinterpret_cast<const __half2*>(&hv[u]);
                        float y[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) { float2 f = __half22float2(hp[e]); y[2 * e] = f.x; y[2 * e + 1] = f.y; }
                        if (b < B && E->add_cond_after && A.x_cond) {
                            const float* cp = A.x_cond + ((size_t)b * A.x_cond_len + (A.x_cond_len > 1 ? t : 0)) * W + k0 + v * 8;
                            const float4 c0 = *reinterpret_cast<const float4*>(cp), c1 = *reinterpret_cast<const float4*>(cp + 4);
                            y[0] += c0.x; y[1] += c0.y; y[2] += c0.z; y[3] += c0.w;
                            y[4] += c1.x; y[5] += c1.y; y[6] += c1.z; y[7] += c1.w;
                        }
                        *reinterpret_cast<float4*>(ys + b * kt + v * 8) = make_float4(y[0], y[1], y[2], y[3]);
                        *reinterpret_cast<float4*>(ys + b * kt + v * 8 + 4) = make_float4(y[4], y[5], y[6], y[7]);
                    }
                }
            }
            consumer_sync();
            const float* y0 = ys + (warp * 2) * kt;
            const float* y1 = y0 + kt;
#pragma unroll
            for (int rc = 0; rc < kLogitRowsPerPass / kLogitRowsPerChunk; ++rc) {
                const int r = pr + rc * kLogitRowsPerChunk;
                if (r < pe) {
                    const int nr = min(kLogitRowsPerChunk, pe - r);
                    mbar_wait(ring.full(), ring.phase);
                    const float* wsl = reinterpret_cast<const float*>(ring.data());
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
                    if (lane == 0) mbar_arrive(ring.empty());
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
__global__ void __launch_bounds__(kThreads, 1) jk_decode_step_kernel(const EngineDev* __restrict__ Eg, StepArgs A) {
    const int tid = threadIdx.x, warp = tid >> 5;
    const int c = blockIdx.x;
    static_assert(offsetof(EngineDev, layer) <= 512, "descriptor head must fit its shared-memory slot");
    static_assert(sizeof(LayerDev) <= 256, "layer record must fit its shared-memory slot");
    for (int i = tid; i < (int)(offsetof(EngineDev, layer) / 4); i += kThreads)
        reinterpret_cast<uint32_t*>(jk_smem + 512)[i] = reinterpret_cast<const uint32_t*>(Eg)[i];
    if (tid < (int)(sizeof(LayerDev) / 4))
        reinterpret_cast<uint32_t*>(jk_smem + 1024)[tid] = reinterpret_cast<const uint32_t*>(&Eg->layer[0])[tid];
    if (tid >= 32 && tid < 36)
        reinterpret_cast<uint32_t*>(jk_smem + 1024 + 128)[tid - 32] =
            reinterpret_cast<const uint32_t*>(Eg->cols + ((size_t)c * Eg->depth + 0) * 4)[tid - 32];
    __syncthreads();
    const EngineDev* E = sm_E();
    Ring ring;
    ring.base_off = kHeaderBytes + E->uni_bytes + E->kvpre_bytes; ring.nslot = E->nslot; ring.slot = 0; ring.phase = 0;
    if (tid == 0) {
        for (int i = 0; i < E->nslot; ++i) { mbar_init(sm_full() + i, 1); mbar_init(sm_empty() + i, 8); }
        mbar_fence_init();
    }
    __syncthreads();
    const bool do_logits = (A.logits != nullptr) && E->bins > 0;
    // Register reallocation between warpgroups (setmaxnreg, sm_90a+): the block launches with 168 registers per
    // thread (65536 / 384); the producer warpgroup keeps 40 and hands the rest to the two consumer warpgroups.
    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 8) producer_loop(Eg, ring, do_logits, c);
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int t = *reinterpret_cast<volatile const int*>(E->t);
    const unsigned epoch0 = *reinterpret_cast<volatile const unsigned*>(E->epoch);
    unsigned nbar = 0, nreal = 0, ndone = 0;    // phase index (profiling slots), real grid barriers, attention items
    const unsigned done0 = *reinterpret_cast<volatile const unsigned*>(E->bar + 1600);
    const int B = A.n, W = E->W, S = E->S, M = E->M, G = E->G;
#define GRID_BARRIER()                                                                     \
    do {                                                                                   \
        STAMP(E, (int)nbar, 4);                                                            \
        if (tid == 0 && E->prof_on && nbar >= 6 && nbar < 11) {                            \
            unsigned long long now_;                                                       \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                       \
            E->prof3[((nbar - 6) * 256 + c) * 2] = now_;                                   \
        }                                                                                  \
        ++nbar;                                                                            \
        ++nreal;                                                                           \
        grid_barrier(E->bar, epoch0 + nreal, c, G);                                        \
        if (tid == 0 && E->prof_on && nbar >= 7 && nbar < 12) {                            \
            unsigned long long now_;                                                       \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                       \
            E->prof3[((nbar - 7) * 256 + c) * 2 + 1] = now_;                               \
        }                                                                                  \
        STAMP(E, (int)nbar - 1, 5);                                                        \
        if (c == 0 && tid == 0 && E->prof_on && nbar < (unsigned)kProfSlots) {             \
            unsigned long long now;                                                        \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));                        \
            E->prof[nbar] = now;                                                           \
        }                                                                                  \
    } while (0)
    if (c == 0 && tid == 0 && E->prof_on) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        E->prof[0] = now;
    }

    // ---- P0: embedding (autoregressive.py:177-197) or an externally embedded activation ------
    for (int e0 = c * kConsumers; e0 < B * W; e0 += G * kConsumers) {
        const int e = e0 + tid;
        const bool act = e < B * W;
        const int b = act ? e / W : 0, col = act ? e - b * W : 0;
        float hv = 0.f;
        if (act) {
            float x;
            if (A.x_in) {
                x = A.x_in[e];
            } else {
                if (t == 0) x = A.y_cond ? A.y_cond[e] : E->start_token[col];
                else x = E->x_emb[(size_t)A.tokens[(size_t)b * A.tok_stride + t - 1] * W + col];
                x += E->pos_emb[(size_t)t * W + col];
                if (A.x_cond) x += A.x_cond[((size_t)b * A.x_cond_len + (A.x_cond_len > 1 ? t : 0)) * W + col];
            }
            const __half hh = __float2half_rn(x);
            E->h[e] = hh;
            hv = __half2float(hh);
        }
        // LayerNorm statistics of layer 0's input: warp-reduce when the warp sits in one row
        long long s1 = act ? fx_sum(hv) : 0, s2 = act ? fx_sq(hv) : 0;
        const int b0 = __shfl_sync(0xffffffffu, b, 0);
        if (__all_sync(0xffffffffu, (!act) || b == b0)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if ((tid & 31) == 0 && (s1 != 0 || s2 != 0)) {
                red_add_s64(E->lnacc + 16 * (2 * b0), s1);
                red_add_s64(E->lnacc + 16 * (2 * b0 + 1), s2);
            }
        } else if (act) {
            red_add_s64(E->lnacc + 16 * (2 * b), s1);
            red_add_s64(E->lnacc + 16 * (2 * b + 1), s2);
        }
    }
    GRID_BARRIER();

#pragma unroll 1
    for (int l = 0; l < E->depth; ++l) {
        const LayerDev& LD = *sm_layer(l);
        const ushort2* cl = reinterpret_cast<const ushort2*>(jk_smem + 1024 + 256 * (l & 1) + 128);
        const int Nqkv = (LD.attn_func == 6) ? S : 3 * S;
        const int pre_ok = attn_prefetch(LD, B, c, t);
        // a fresh argument record per phase: nothing of it stays live across the calls in between
        {
            GemmArgs ga;
            ga.in = E->h; ga.K = W; ga.N = Nqkv; ga.g0 = cl[0].x; ga.ncg = cl[0].y; ga.ln = 1; ga.epi = EPI_QKV;
            ga.pslot = (int)nbar; ga.gamma = LD.ln0_g; ga.beta = LD.ln0_b; ga.bias = LD.b_qkv;
            ga.ln_in = E->lnacc + (size_t)(2 * l) * 512; ga.ln_out = nullptr;
            ring = gemm_phase(ring, B, ga);
        }
        // next layer's record + column assignment -> the other shared-memory slot.  The descriptor is in
        // HBM (the weight stream evicts it from L2 every step): issue the loads here so their latency hides
        // behind the barrier instead of sitting on the dependency chain.
        if (l + 1 < E->depth) {
            if (tid < (int)(sizeof(LayerDev) / 4))
                reinterpret_cast<uint32_t*>(jk_smem + 1024 + 256 * ((l + 1) & 1))[tid] =
                    reinterpret_cast<const uint32_t*>(&Eg->layer[l + 1])[tid];
            if (tid >= 32 && tid < 36)
                reinterpret_cast<uint32_t*>(jk_smem + 1024 + 256 * ((l + 1) & 1) + 128)[tid - 32] =
                    reinterpret_cast<const uint32_t*>(E->cols + ((size_t)c * E->depth + l + 1) * 4)[tid - 32];
        }
        GRID_BARRIER();
        {
            const AttnGeom geo = attn_geom(E, LD, t);
            const int ns = attn_nsplit(E, B, geo.R - (geo.cur ? 1 : 0));
            for (int it = c; it < B * E->H * ns; it += G) {
                const int s = it % ns, bh = it / ns;
                attn_item(LD, bh / E->H, bh % E->H, s, ns, geo, (int)nbar, pre_ok && it == c);
            }
            STAMP(E, (int)nbar, 4);
            ++nbar;
            ndone += (unsigned)(B * E->H);
            attn_done_wait(E, done0 + ndone);
            STAMP(E, (int)nbar - 1, 5);
            if (c == 0 && tid == 0 && E->prof_on && nbar < (unsigned)kProfSlots) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                E->prof[nbar] = now;
            }
        }
        if (c == 0 && tid < 32) E->lnacc[(size_t)(2 * l) * 512 + 16 * tid] = 0;      // LN0 statistics of this layer are consumed
        {
            GemmArgs ga;
            ga.in = E->a; ga.K = S; ga.N = W; ga.g0 = cl[1].x; ga.ncg = cl[1].y; ga.ln = 0; ga.epi = EPI_PROJ;
            ga.pslot = (int)nbar; ga.gamma = nullptr; ga.beta = nullptr; ga.bias = LD.b_o; ga.ln_in = nullptr;
            ga.ln_out = E->lnacc + (size_t)(2 * l + 1) * 512;
            ring = gemm_phase(ring, B, ga);
        }
        GRID_BARRIER();
        {
            GemmArgs ga;
            ga.in = E->x1; ga.K = W; ga.N = M; ga.g0 = cl[2].x; ga.ncg = cl[2].y; ga.ln = 1; ga.epi = EPI_FC;
            ga.pslot = (int)nbar; ga.gamma = LD.ln1_g; ga.beta = LD.ln1_b; ga.bias = LD.b_1;
            ga.ln_in = E->lnacc + (size_t)(2 * l + 1) * 512; ga.ln_out = nullptr;
            ring = gemm_phase(ring, B, ga);
        }
        GRID_BARRIER();
        if (c == 0 && tid < 32) E->lnacc[(size_t)(2 * l + 1) * 512 + 16 * tid] = 0;  // LN1 statistics are consumed
        {
            GemmArgs ga;
            ga.in = E->g; ga.K = M; ga.N = W; ga.g0 = cl[3].x; ga.ncg = cl[3].y; ga.ln = 0; ga.epi = EPI_PROJ2;
            ga.pslot = (int)nbar; ga.gamma = nullptr; ga.beta = nullptr; ga.bias = LD.b_2; ga.ln_in = nullptr;
            ga.ln_out = (l + 1 < E->depth) ? E->lnacc + (size_t)(2 * l + 2) * 512 : nullptr;
            ring = gemm_phase(ring, B, ga);
        }
        GRID_BARRIER();
    }
    if (A.h_out) {
        for (int e = c * kConsumers + tid; e < B * W; e += G * kConsumers)
            A.h_out[e] = ld_half_cg(E->h + e);
    }
    if (do_logits) logits_phase(A, ring, c, t);
    if (c == 0 && tid == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (E->prof_on && nbar + 1 < (unsigned)kProfSlots) E->prof[nbar + 1] = now;
        *E->t = t + 1;
        *E->epoch = epoch0 + nreal;
        *(E->bar + 1600) = done0 + ndone;
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

// encoder K/V for attn_func 6: kv = fp16(fp16(enc) . Wkv + b), once per window.  The product itself runs on
// the tcgen05 prefill GEMM (prefill_gemm.cu); these kernels only convert / lay out its operands and result.
template <typename T>
__global__ void transpose_to_half_kernel(const T* __restrict__ src, __half* __restrict__ dst, int K, int N) {
    // src [K][N] row-major -> dst [N][K] row-major (the K-major layout the tensor core reads)
    __shared__ __half tile[32][33];
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, n = n0 + threadIdx.x;
        tile[i][threadIdx.x] = (k < K && n < N) ? to_half<T>(src[(size_t)k * N + n]) : __float2half_rn(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int n = n0 + i, k = k0 + threadIdx.x;
        if (n < N && k < K) dst[(size_t)n * K + k] = tile[threadIdx.x][i];
    }
}

__global__ void enc_kv_scatter_kernel(const __half* __restrict__ y, __half* kc, __half* vc, int rows_total, int E_dims,
                                      int S, int H, int dh, int dhp) {
    // y [rows_total][2S] -> K, V caches [b][h][e][dh_pad]
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows_total * 2 * S) return;
    const int r = (int)(i / (2 * S)), cidx = (int)(i % (2 * S));
    const int b = r / E_dims, e = r % E_dims;
    const int which = cidx / S, cs = cidx % S;
    const int h = cs / dh, d = cs % dh;
    __half* dst = which ? vc : kc;
    dst[(((size_t)(b * H + h)) * E_dims + e) * dhp + d] = y[i];
}

}  // namespace

// =========================================================================================
// host side
// =========================================================================================
namespace {

struct Layout {
    size_t off_dev, off_cols, off_soff, off_goff, off_lrow, off_streams, off_small, off_cache, off_h, off_x1, off_qkv, off_a, off_g, off_part, off_acnt, off_prof, off_prof2, off_prof3, off_lnacc, off_encx, off_ency, off_wt, off_pf, off_sync, total;
    size_t wt_per_layer;
    int pf_len, pf_rows;
    size_t stream_stride;
    std::vector<ushort2> cols;
    std::vector<uint32_t> soff, goff;
    std::vector<int> lrow;
    std::vector<size_t> cache_off;      // per layer (K); V follows
    std::vector<size_t> cache_bytes;
    std::vector<int> cache_rows;
    size_t small_per_layer;
    int dh, dh_pad, bc, prime_pad, uni_bytes, kvpre_bytes, kv_prefetch, nslot, smem_bytes;
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
    L.dh_pad = (int)align_up(L.dh, 16);      // MMA k-steps / output pairs of 16 dims
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
    const int RC = attn_tile_rows(L.dh_pad);
    const size_t kv_stage = (size_t)2 * RC * L.dh_pad * 2;                        // one K tile + one V tile
    size_t attn = kv_stage + (size_t)L.dh_pad * 2 + 64 * 4 + (size_t)L.dh_pad * 4 + 64;   // tiles, q, scores, running output
    uni = std::max(uni, attn);
    L.uni_bytes = (int)align_up(uni, 1024);
    const int max_smem = 232448;
    L.kvpre_bytes = (int)align_up(kv_stage, 1024);      // second K/V stage: prefetch target / double buffer
    // prefetching the first attention tile before the QKV GEMM hides its latency but costs more issue time
    // than it saves (measured 2703 vs 2653 us / step at position 4000): off unless JK_KV_PREFETCH is set
    L.kv_prefetch = getenv("JK_KV_PREFETCH") ? 1 : 0;
    int nslot = (max_smem - kHeaderBytes - L.uni_bytes - L.kvpre_bytes) / kSlotBytes;
    nslot = std::min(nslot, kMaxSlots);
    JK_REQUIRE(nslot >= 2, "not enough shared memory for the weight ring (uni %d bytes)", L.uni_bytes);
    L.nslot = nslot;
    L.smem_bytes = kHeaderBytes + L.uni_bytes + L.kvpre_bytes + nslot * kSlotBytes;

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
    L.off_acnt = off; off = align_up(off + (size_t)c.max_batch * c.heads * 4, 256);
    L.off_prof = off; off = align_up(off + (size_t)kProfSlots * 8, 256);
    L.off_prof2 = off; off = align_up(off + (size_t)kProfSlots * 8 * 8, 256);
    L.off_prof3 = off; off = align_up(off + (size_t)5 * 256 * 2 * 8, 256);
    L.off_lnacc = off; off = align_up(off + (size_t)2 * depth * 512 * 8, 256);
    {
        bool any6 = false;
        for (int l = 0; l < depth; ++l) any6 = any6 || (c.attn_func[l] == 6);
        L.off_encx = off; if (any6) off = align_up(off + (size_t)c.max_batch * c.encoder_dims * c.width * 2, 1024);
        L.off_ency = off; if (any6) off = align_up(off + (size_t)c.max_batch * c.encoder_dims * 2 * c.n_state * 2, 1024);
    }
    {   // chunked prefill: K-major fp16 weight copies + activation workspace (prefill.cu).  Needs every GEMM K to
        // be a multiple of the tcgen05 K block (64) and no encoder-decoder layer; else pf_rows = 0.
        bool ok = (c.width % 64 == 0) && (c.n_state % 64 == 0) && (c.mlp_width % 64 == 0) && !getenv("JK_NO_PREFILL");
        for (int l = 0; l < depth; ++l) ok = ok && (c.attn_func[l] != 6);
        L.pf_len = ok ? std::min(c.n_ctx, 512) : 0;
        L.pf_rows = c.max_batch * L.pf_len;
        L.wt_per_layer = align_up((size_t)(3 * c.n_state * c.width + c.width * c.n_state + 2 * c.mlp_width * c.width) * 2, 1024);
        L.off_wt = off; if (ok) off = align_up(off + L.wt_per_layer * depth, 1024);
        L.off_pf = off;
        if (ok) off = align_up(off + (size_t)L.pf_rows * (3 * c.width + 3 * c.n_state + c.n_state + c.mlp_width) * 2, 1024);
    }
    L.off_sync = off; off += 8192;
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
    E.depth = cfg->depth; E.G = G; E.nslot = L.nslot; E.uni_bytes = L.uni_bytes; E.kvpre_bytes = L.kvpre_bytes; E.kv_prefetch = L.kv_prefetch; E.small_bytes = (int)L.small_per_layer; E.prof_on = getenv("JK_PROFILE") ? 1 : 0;
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
    E.acnt = (unsigned*)(A + L.off_acnt); E.prof = (unsigned long long*)(A + L.off_prof);
    E.lnacc = (long long*)(A + L.off_lnacc);
    E.prof2 = (long long*)(A + L.off_prof2);
    E.prof3 = (unsigned long long*)(A + L.off_prof3);
    {
        const char* sr = getenv("JK_ATTN_SPLIT_ROWS");
        E.split_rows = sr ? atoi(sr) : 96;
        if (E.split_rows < 8) E.split_rows = 8;
    }
    E.bar = (unsigned*)(A + L.off_sync); E.epoch = E.bar + 1536; E.t = (int*)(E.bar + 1568);
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
    p->enc_x16 = (__half*)(A + L.off_encx); p->enc_y16 = (__half*)(A + L.off_ency);
    p->pf_rows = L.pf_rows; p->pf_len = L.pf_len;
    for (int i = 0; i < 4; ++i) p->wt[i].assign(cfg->depth, nullptr);
    if (L.pf_rows) {
        for (int l = 0; l < cfg->depth; ++l) {
            __half* w = (__half*)(A + L.off_wt + L.wt_per_layer * l);
            p->wt[0][l] = w; w += (size_t)3 * cfg->n_state * cfg->width;
            p->wt[1][l] = w; w += (size_t)cfg->width * cfg->n_state;
            p->wt[2][l] = w; w += (size_t)cfg->mlp_width * cfg->width;
            p->wt[3][l] = w;
        }
        __half* a = (__half*)(A + L.off_pf);
        const size_t R = (size_t)L.pf_rows;
        p->pf_x = a; a += R * cfg->width;
        p->pf_xn = a; a += R * cfg->width;
        p->pf_x1 = a; a += R * cfg->width;
        p->pf_qkv = a; a += R * 3 * cfg->n_state;
        p->pf_a = a; a += R * cfg->n_state;
        p->pf_g = a;
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
    if (p->pf_rows) {      // K-major fp16 copies for the tensor-core prefill
        for (int gi = 0; gi < 4; ++gi) {
            dim3 tg((Ns[gi] + 31) / 32, (Ks[gi] + 31) / 32), tb(32, 8);
            if (w->w_dtype) transpose_to_half_kernel<__half><<<tg, tb, 0, stream>>>((const __half*)ws[gi], p->wt[gi][l], Ks[gi], Ns[gi]);
            else transpose_to_half_kernel<float><<<tg, tb, 0, stream>>>((const float*)ws[gi], p->wt[gi][l], Ks[gi], Ns[gi]);
            JK_CHECK_CUDA(cudaGetLastError());
        }
    }
    const float* lns[4] = {w->ln0_g, w->ln0_b, w->ln1_g, w->ln1_b};
    for (int i = 0; i < 4; ++i) {
        JK_REQUIRE(lns[i], "layer %d: missing LayerNorm parameter %d", l, i);
        JK_CHECK_CUDA(cudaMemcpyAsync(p->ln_ptr[i][l], lns[i], (size_t)c.width * 4, cudaMemcpyDeviceToDevice, stream));
    }
    if (af == 6) {
        JK_REQUIRE(w->c_enc_kv_w && w->c_enc_kv_b, "layer %d: attn_func 6 needs c_enc_kv", l);
        dim3 tg((2 * c.n_state + 31) / 32, (c.width + 31) / 32), tb(32, 8);
        if (w->w_dtype) transpose_to_half_kernel<__half><<<tg, tb, 0, stream>>>((const __half*)w->c_enc_kv_w, p->enc_w[l], c.width, 2 * c.n_state);
        else transpose_to_half_kernel<float><<<tg, tb, 0, stream>>>((const float*)w->c_enc_kv_w, p->enc_w[l], c.width, 2 * c.n_state);
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
    JK_REQUIRE(c.width % 64 == 0, "encoder-decoder layers need width %% 64 == 0 (tcgen05 K block)");
    {   // encoder_kv.type_as(x): fp32 -> fp16 once, shared by every enc-dec layer
        const size_t cnt = (size_t)rows * c.width;
        to_half_kernel<float><<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(encoder_kv, p->enc_x16, cnt);
        JK_CHECK_CUDA(cudaGetLastError());
    }
    for (int l = 0; l < c.depth; ++l) {
        if (c.attn_func[l] != 6) continue;
        int rc = jk_conv1d_prefill_f16(p->enc_x16, p->enc_w[l], p->enc_b[l], p->enc_y16, rows, 2 * c.n_state, c.width, stream_);
        if (rc) return rc;
        const size_t cnt = (size_t)rows * 2 * c.n_state;
        enc_kv_scatter_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(p->enc_y16, p->host.layer[l].kc, p->host.layer[l].vc,
                                                                                  rows, c.encoder_dims, c.n_state, c.heads,
                                                                                  p->host.dh, p->host.dh_pad);
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
        case 5: *ptr = p->host.prof; *n = (size_t)kProfSlots * 4; break;     /* uint64 timestamps */
        case 7: *ptr = p->host.prof3; *n = (size_t)5 * 256 * 2 * 4; break;
        case 6: *ptr = p->host.prof2; *n = (size_t)kProfSlots * 8 * 4; break; /* int64 clock64 stamps [slot][8] */
        default: JK_REQUIRE(false, "unknown buffer %d", which);
    }
    return 0;
}
