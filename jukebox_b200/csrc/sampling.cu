// Token sampling for the autoregressive loop: x / temp -> Categorical(logits=x).sample()
// (reference jukebox/prior/autoregressive.py:233-235 and :343-345) as ONE launch per position instead
// of the ~10 elementwise/reduction launches the torch expression costs between two decode steps.
//
// One CTA per sample row.  Each thread owns a contiguous run of bins so that the inclusive scan of
// exp(v - max) is the CDF in bin order; the token is the first bin whose CDF reaches u * total,
// u in (0, 1] from Philox4x32-10 keyed by (seed) and countered by (position, row) - the draw for a
// given (seed, position, row) does not depend on launch order or on the other rows.
#include "common.cuh"
#include "../../include/jkb200.h"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint32_t c0, uint32_t c1) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c[4] = {c0, c1, 0x6a6b3230u, 0u};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c[0];
}

// max_per: bins each thread owns (compile-time bound keeps e[] in registers)
template <int MAX_PER>
__global__ void __launch_bounds__(kThreads)
sample_categorical_kernel(const float* __restrict__ logits, long long lstride, int bins, float temp,
                          unsigned long long seed, int position, long long* __restrict__ tokens,
                          long long tok_stride) {
    __shared__ float s_red[kThreads / 32];
    __shared__ float s_scan[kThreads / 32];
    __shared__ float s_bcast[2];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* l = logits + (long long)row * lstride;
    const int per = (bins + kThreads - 1) / kThreads;
    const int b0 = tid * per;
    float e[MAX_PER];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAX_PER; ++j) {
        const int b = b0 + j;
        e[j] = (j < per && b < bins) ? __ldcg(l + b) / temp : -INFINITY;
        mx = fmaxf(mx, e[j]);
    }
    mx = jk::warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) mx = fmaxf(mx, s_red[w]);
    float local = 0.f;
#pragma unroll
    for (int j = 0; j < MAX_PER; ++j) {
        e[j] = (e[j] == -INFINITY) ? 0.f : __expf(e[j] - mx);
        local += e[j];
    }
    // inclusive scan of the per-thread sums: warp shuffle scan, then the warp totals
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    float woff = 0.f, total = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        if (w < warp) woff += s_scan[w];
        total += s_scan[w];
    }
    incl += woff;
    const float excl = incl - local;
    if (tid == 0) {
        const uint32_t r = philox_u32(seed, (uint32_t)position, (uint32_t)row);
        s_bcast[0] = (float)((r >> 8) + 1u) * (1.0f / 16777216.0f) * total;    // u in (2^-24, 1]
    }
    if (tid == 0) s_bcast[1] = __int_as_float(0x7fffffff);
    __syncthreads();
    const float target = s_bcast[0];
    // the owning thread: first whose inclusive sum reaches the target (ties between threads with an
    // empty run cannot win: local > 0 is required)
    const bool mine = local > 0.f && excl < target && incl >= target;
    const bool last_resort = (tid == kThreads - 1);
    int pick = -1;
    if (mine) {
        float run = excl;
        int fallback = -1;
#pragma unroll
        for (int j = 0; j < MAX_PER; ++j) {
            if (e[j] > 0.f) {
                run += e[j];
                fallback = b0 + j;
                if (pick < 0 && run >= target) pick = b0 + j;
            }
        }
        if (pick < 0) pick = fallback;
        tokens[(long long)row * tok_stride + position] = pick;
        s_bcast[1] = 0.f;
    }
    __syncthreads();
    // rounding of woff/incl can leave no owner in a pathological row: take the last bin with mass
    if (s_bcast[1] != 0.f) {
        __shared__ int s_last;
        if (tid == 0) s_last = -1;
        __syncthreads();
        int lastb = -1;
#pragma unroll
        for (int j = 0; j < MAX_PER; ++j)
            if (e[j] > 0.f) lastb = b0 + j;
        if (lastb >= 0) atomicMax(&s_last, lastb);
        __syncthreads();
        if (last_resort) tokens[(long long)row * tok_stride + position] = s_last < 0 ? 0 : s_last;
    }
}

}  // namespace

extern "C" int jk_sample_categorical(const float* logits, int64_t logits_stride, int n, int bins, float temp,
                                     uint64_t seed, int position, int64_t* tokens, int64_t tok_stride,
                                     jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(logits && tokens, "null argument");
    JK_REQUIRE(bins >= 1 && bins <= 32 * kThreads, "bins must be in [1, %d]", 32 * kThreads);
    JK_REQUIRE(temp > 0.f, "temp must be positive");
    JK_REQUIRE(position >= 0, "negative position");
    if (n == 0) return 0;
    const int per = (bins + kThreads - 1) / kThreads;
#define JK_LAUNCH(MP)                                                                              \
    sample_categorical_kernel<MP><<<n, kThreads, 0, stream>>>(logits, (long long)logits_stride, bins, temp, \
                                                              (unsigned long long)seed, position,  \
                                                              (long long*)tokens, (long long)tok_stride)
    if (per <= 4) JK_LAUNCH(4);
    else if (per <= 8) JK_LAUNCH(8);
    else if (per <= 16) JK_LAUNCH(16);
    else JK_LAUNCH(32);
#undef JK_LAUNCH
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}
