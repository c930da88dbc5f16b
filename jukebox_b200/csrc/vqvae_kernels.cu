// VQ-VAE kernels (fp32, channels-last [N, T, C]) and small fp32 helpers used once per window.
//
// Reference (under /root/reference/jukebox/vqvae/): bottleneck.py:112-123 (quantise / dequantise),
// encdec.py:6-131 and resnet.py:27-75 (Conv1d / ConvTranspose1d / ResConv1DBlock stacks).
// The encoder feeds an argmin whose indices must be bit-exact, so these kernels keep true fp32 FMA
// arithmetic (no TF32): see DESIGN.md "VQ-VAE numerics".
#include "common.cuh"
#include "../../include/jkb200.h"

using namespace jk;

namespace {

// ---------------------------------------------------------------------------------------
// codebook argmin.  64 rows per CTA, x rows live in registers, the codebook streams through
// shared memory in tiles of 128 codes; the distance matrix is never materialised.
// dist = (|x|^2 - 2 x.k) + |k|^2 evaluated in fp32 exactly as bottleneck.py:115-117 associates it.
// ---------------------------------------------------------------------------------------
template <int WIDTH>
__global__ void __launch_bounds__(256) vq_argmin_kernel(const float* __restrict__ x, const float* __restrict__ cb,
                                                        long long* __restrict__ idx, float* __restrict__ mind,
                                                        long long n, int kbins) {
    constexpr int TILE = 128;
    __shared__ __align__(16) float cs[TILE][WIDTH];
    __shared__ float kk[TILE];
    __shared__ float bd[4][64];
    __shared__ int bi[4][64];
    const int tid = threadIdx.x, r = tid & 63, q = tid >> 6;
    const long long row = (long long)blockIdx.x * 64 + r;
    float xr[WIDTH];
    float xx = 0.f;
    if (row < n) {
#pragma unroll
        for (int d = 0; d < WIDTH; d += 4) {
            float4 v = *reinterpret_cast<const float4*>(x + row * WIDTH + d);
            xr[d] = v.x; xr[d + 1] = v.y; xr[d + 2] = v.z; xr[d + 3] = v.w;
        }
#pragma unroll
        for (int d = 0; d < WIDTH; ++d) xx += xr[d] * xr[d];
    } else {
#pragma unroll
        for (int d = 0; d < WIDTH; ++d) xr[d] = 0.f;
    }
    float best = INFINITY;
    int besti = 0;
    for (int c0 = 0; c0 < kbins; c0 += TILE) {
        __syncthreads();
        for (int i = tid; i < TILE * WIDTH / 4; i += 256) {
            int j = i / (WIDTH / 4), d4 = i % (WIDTH / 4);
            float4 v = make_float4(0, 0, 0, 0);
            if (c0 + j < kbins) v = *reinterpret_cast<const float4*>(cb + (size_t)(c0 + j) * WIDTH + d4 * 4);
            *reinterpret_cast<float4*>(&cs[j][d4 * 4]) = v;
        }
        __syncthreads();
        if (tid < TILE) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < WIDTH; ++d) s += cs[tid][d] * cs[tid][d];
            kk[tid] = s;
        }
        __syncthreads();
        const int jend = min(TILE, kbins - c0);
        for (int j = q * (TILE / 4); j < (q + 1) * (TILE / 4) && j < jend; ++j) {
            float dot = 0.f;
#pragma unroll
            for (int d = 0; d < WIDTH; d += 4) {
                float4 k4 = *reinterpret_cast<const float4*>(&cs[j][d]);
                dot = fmaf(xr[d], k4.x, dot);
                dot = fmaf(xr[d + 1], k4.y, dot);
                dot = fmaf(xr[d + 2], k4.z, dot);
                dot = fmaf(xr[d + 3], k4.w, dot);
            }
            float dist = (xx - 2.0f * dot) + kk[j];
            if (dist < best) { best = dist; besti = c0 + j; }
        }
    }
    bd[q][r] = best;
    bi[q][r] = besti;
    __syncthreads();
    if (q == 0 && row < n) {
        float b = bd[0][r];
        int bidx = bi[0][r];
#pragma unroll
        for (int qq = 1; qq < 4; ++qq) {
            float d = bd[qq][r];
            int i = bi[qq][r];
            if (d < b || (d == b && i < bidx)) { b = d; bidx = i; }
        }
        idx[row] = bidx;
        if (mind) mind[row] = b;
    }
}

// generic width (multiple of 4, <= 512): one warp per row, lanes over codes
__global__ void vq_argmin_generic_kernel(const float* __restrict__ x, const float* __restrict__ cb,
                                         long long* __restrict__ idx, float* __restrict__ mind, long long n, int kbins,
                                         int width) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* xr = x + row * width;
    float xx = 0.f;
    for (int d = 0; d < width; ++d) xx += xr[d] * xr[d];
    float best = INFINITY;
    int besti = 0x7fffffff;
    for (int j = lane; j < kbins; j += 32) {
        const float* k = cb + (size_t)j * width;
        float dot = 0.f, ks = 0.f;
        for (int d = 0; d < width; ++d) { dot = fmaf(xr[d], k[d], dot); ks += k[d] * k[d]; }
        float dist = (xx - 2.0f * dot) + ks;
        if (dist < best) { best = dist; besti = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) { idx[row] = besti; if (mind) mind[row] = best; }
}

__global__ void vq_gather_kernel(const long long* __restrict__ idx, const float* __restrict__ cb, float* __restrict__ out,
                                 long long n, int kbins, int width4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * width4) return;
    const long long r = i / width4;
    const int d = (int)(i % width4);
    long long j = idx[r];
    j = j < 0 ? 0 : (j >= kbins ? kbins - 1 : j);
    reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(cb)[j * width4 + d];
}

// ---------------------------------------------------------------------------------------
// generic channels-last conv: tile of 64 positions x 64 output channels per CTA,
// K = taps x c_in walked in slabs of 32 input channels through shared memory.
// ---------------------------------------------------------------------------------------
struct ConvP {
    const float* in; long long t_in; int c_in;
    float* out; long long t_out; int c_out;
    const float* w; const float* bias; const float* res;
    int n_taps; int tap_off[4]; int in_stride; int out_stride; int out_offset; int relu_in; float scale;
};

__global__ void __launch_bounds__(256) conv1d_cl_kernel(ConvP P) {
    constexpr int TT = 64, CT = 64, KC = 32;
    __shared__ float xs[KC][TT + 1];
    __shared__ __align__(16) float ws[KC][CT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long t0 = (long long)blockIdx.x * TT;
    const int co0 = blockIdx.y * CT;
    const int nb = blockIdx.z;
    const float* in = P.in + (size_t)nb * P.t_in * P.c_in;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int tap = 0; tap < P.n_taps; ++tap) {
        const int off = P.tap_off[tap];
        for (int ci0 = 0; ci0 < P.c_in; ci0 += KC) {
            __syncthreads();
            for (int i = tid; i < KC * TT; i += 256) {
                const int k = i % KC, tt = i / KC;
                const long long tp = (t0 + tt) * P.in_stride + off;
                float v = 0.f;
                if (ci0 + k < P.c_in && tp >= 0 && tp < P.t_in && t0 + tt < P.t_out) {
                    v = in[(size_t)tp * P.c_in + ci0 + k];
                    if (P.relu_in) v = fmaxf(v, 0.f);
                }
                xs[k][tt] = v;
            }
            for (int i = tid; i < KC * CT; i += 256) {
                const int c = i % CT, k = i / CT;
                float v = 0.f;
                if (ci0 + k < P.c_in && co0 + c < P.c_out) v = P.w[((size_t)tap * P.c_in + ci0 + k) * P.c_out + co0 + c];
                ws[k][c] = v;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < KC; ++k) {
                float a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = xs[k][ty * 4 + i];
                const float4 b = *reinterpret_cast<const float4*>(&ws[k][tx * 4]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                    acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                    acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                    acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
                }
            }
        }
    }
    const long long rows_out = P.t_out * P.out_stride;
    float* out = P.out + (size_t)nb * rows_out * P.c_out;
    const float* res = P.res ? P.res + (size_t)nb * rows_out * P.c_out : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long t = t0 + ty * 4 + i;
        if (t >= P.t_out) continue;
        const long long orow = t * P.out_stride + P.out_offset;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tx * 4 + j;
            if (co >= P.c_out) continue;
            float v = P.scale * (acc[i][j] + (P.bias ? P.bias[co] : 0.f));
            if (res) v += res[(size_t)orow * P.c_out + co];
            out[(size_t)orow * P.c_out + co] = v;
        }
    }
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ packed, int c_out, int c_in,
                                        int k, int transposed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = c_out * c_in * k;
    if (i >= total) return;
    const int co = i % c_out, ci = (i / c_out) % c_in, kk = i / (c_out * c_in);
    const size_t src = transposed ? ((size_t)ci * c_out + co) * k + kk : ((size_t)co * c_in + ci) * k + kk;
    packed[i] = w[src];
}

__global__ void layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                     float* __restrict__ y, long long rows, int width, float eps) {
    __shared__ float sh[8];
    const long long r = blockIdx.x;
    const float* xr = x + r * width;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float s = 0.f;
    for (int i = tid; i < width; i += 256) s += xr[i];
    s = warp_sum(s);
    if (lane == 0) sh[warp] = s;
    __syncthreads();
    float mean = 0.f;
    for (int w = 0; w < 8; ++w) mean += sh[w];
    mean /= (float)width;
    __syncthreads();
    float ss = 0.f;
    for (int i = tid; i < width; i += 256) { float d = xr[i] - mean; ss += d * d; }
    ss = warp_sum(ss);
    if (lane == 0) sh[warp] = ss;
    __syncthreads();
    float var = 0.f;
    for (int w = 0; w < 8; ++w) var += sh[w];
    const float rstd = 1.0f / sqrtf(var / (float)width + eps);
    for (int i = tid; i < width; i += 256) y[r * width + i] = (xr[i] - mean) * rstd * g[i] + b[i];
}

__global__ void embedding_f32_kernel(const long long* __restrict__ idx, const float* __restrict__ table,
                                     const float* __restrict__ add, float* __restrict__ out, long long n, int rows,
                                     int width) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * width) return;
    const long long r = i / width;
    const int d = (int)(i % width);
    long long j = idx[r];
    j = j < 0 ? 0 : (j >= rows ? rows - 1 : j);
    float v = table[j * width + d];
    if (add) v += add[i];
    out[i] = v;
}

}  // namespace

extern "C" int jk_vq_argmin(const float* x, const float* codebook, int64_t* idx, float* min_dist, int64_t n, int k_bins,
                            int width, jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(x && codebook && idx, "null argument");
    JK_REQUIRE(n >= 0 && k_bins >= 1 && width >= 1, "bad sizes");
    if (n == 0) return 0;
    if (width == 64) {
        vq_argmin_kernel<64><<<(unsigned)((n + 63) / 64), 256, 0, stream>>>(x, codebook, (long long*)idx, min_dist, n, k_bins);
    } else {
        vq_argmin_generic_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(x, codebook, (long long*)idx, min_dist, n, k_bins, width);
    }
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int jk_vq_gather(const int64_t* idx, const float* codebook, float* out, int64_t n, int k_bins, int width,
                            jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(idx && codebook && out, "null argument");
    JK_REQUIRE(width % 4 == 0, "width must be a multiple of 4");
    if (n == 0) return 0;
    const long long total = n * (width / 4);
    vq_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const long long*)idx, codebook, out, n, k_bins, width / 4);
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int jk_conv1d_cl(const jk_conv_args* a, jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(a && a->in && a->out && a->w, "null argument");
    JK_REQUIRE(a->n_taps >= 1 && a->n_taps <= 4, "n_taps must be 1..4");
    JK_REQUIRE(a->n >= 1 && a->n <= 65535, "batch out of range");
    JK_REQUIRE(a->in_stride >= 1 && a->out_stride >= 1, "bad strides");
    if (a->t_out == 0) return 0;
    ConvP P;
    P.in = a->in; P.t_in = a->t_in; P.c_in = a->c_in; P.out = a->out; P.t_out = a->t_out; P.c_out = a->c_out;
    P.w = a->w; P.bias = a->bias; P.res = a->res; P.n_taps = a->n_taps;
    for (int i = 0; i < 4; ++i) P.tap_off[i] = a->tap_off[i];
    P.in_stride = a->in_stride; P.out_stride = a->out_stride; P.out_offset = a->out_offset; P.relu_in = a->relu_in;
    P.scale = a->scale;
    dim3 grid((unsigned)((a->t_out + 63) / 64), (unsigned)((a->c_out + 63) / 64), (unsigned)a->n);
    conv1d_cl_kernel<<<grid, 256, 0, stream>>>(P);
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int jk_resblock_cl(const float* x, float* out, float* tmp, const float* w1, const float* b1, const float* w2,
                              const float* b2, int n, int64_t T, int C, int Cs, int dilation, float res_scale,
                              jk_stream_t stream) {
    JK_REQUIRE(x && out && tmp && w1 && w2, "null argument");
    jk_conv_args a;
    a.in = x; a.t_in = T; a.c_in = C; a.out = tmp; a.t_out = T; a.c_out = Cs; a.w = w1; a.bias = b1; a.res = nullptr;
    a.n_taps = 3; a.tap_off[0] = -dilation; a.tap_off[1] = 0; a.tap_off[2] = dilation; a.tap_off[3] = 0;
    a.in_stride = 1; a.out_stride = 1; a.out_offset = 0; a.relu_in = 1; a.scale = 1.0f; a.n = n;
    int rc = jk_conv1d_cl(&a, stream);
    if (rc) return rc;
    a.in = tmp; a.c_in = Cs; a.out = out; a.c_out = C; a.w = w2; a.bias = b2; a.res = x; a.n_taps = 1; a.tap_off[0] = 0;
    a.scale = res_scale;
    return jk_conv1d_cl(&a, stream);
}

extern "C" int jk_pack_conv_weight(const float* w, float* packed, int c_out, int c_in, int k, int transposed,
                                   jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(w && packed, "null argument");
    const int total = c_out * c_in * k;
    pack_conv_weight_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, packed, c_out, c_in, k, transposed);
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int jk_layernorm_f32(const float* x, const float* g, const float* b, float* y, int64_t rows, int width,
                                float eps, jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(x && g && b && y, "null argument");
    if (rows == 0) return 0;
    JK_REQUIRE(rows < (1ll << 31), "too many rows");
    layernorm_f32_kernel<<<(unsigned)rows, 256, 0, stream>>>(x, g, b, y, rows, width, eps);
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int jk_embedding_f32(const int64_t* idx, const float* table, const float* add, float* out, int64_t n, int rows,
                                int width, jk_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    JK_REQUIRE(idx && table && out, "null argument");
    if (n == 0) return 0;
    const long long total = n * width;
    embedding_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const long long*)idx, table, add, out, n, rows, width);
    JK_CHECK_CUDA(cudaGetLastError());
    return 0;
}
