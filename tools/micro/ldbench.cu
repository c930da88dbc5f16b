// Do a thread's independent L2 loads overlap?  One CTA per SM, 256 threads, every thread issues N independent 16-byte
// loads of L2-resident lines written by ANOTHER kernel (so nothing is in L1), in four forms:
//   relaxed : ld.relaxed.gpu.global.v2.u64   (SASS LDG.E.128.STRONG.GPU) - what a polled LL word needs
//   weak    : ld.global.v2.u64               (SASS LDG.E.128, may hit L1)
//   nc      : ld.global.nc.v2.u64            (SASS LDG.E.128.CONSTANT)
//   cpasync : cp.async.cg.shared.global 16   (SASS LDGSTS, L2 only) + wait_group + LDS.128
// Prints cycles from first issue to last value consumed, for N = 1, 2, 4, 8, 16.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int MODE, int N>
__global__ void ld_kernel(const unsigned long long* p, unsigned long long* out, long long* cyc, int stride_words) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int tid = threadIdx.x;
    const unsigned long long* src = p + (size_t)blockIdx.x * 65536 + tid * 2;
    unsigned long long acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 3) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm) + tid * 16;
#pragma unroll
        for (int i = 0; i < N; ++i)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 4096), "l"(src + (size_t)i * stride_words));
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sm + tid * 16 + i * 4096);
            acc += v.x ^ v.y;
        }
    } else {
        ulonglong2 v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const unsigned long long* a = src + (size_t)i * stride_words;
            if (MODE == 0) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(v[i].x), "=l"(v[i].y) : "l"(a) : "memory");
            if (MODE == 1) asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(v[i].x), "=l"(v[i].y) : "l"(a) : "memory");
            if (MODE == 2) asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(v[i].x), "=l"(v[i].y) : "l"(a) : "memory");
        }
#pragma unroll
        for (int i = 0; i < N; ++i) acc += v[i].x ^ v[i].y;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void fill_kernel(unsigned long long* p, size_t n, unsigned long long v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v + i;
}

template <int MODE, int N>
void run(const char* name, unsigned long long* buf, size_t words, unsigned long long* out, long long* cyc, int sms) {
    long long best = 1ll << 60;
    cudaFuncSetAttribute(ld_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int rep = 0; rep < 5; ++rep) {
        fill_kernel<<<sms * 4, 256>>>(buf, words, rep);      // lines are rewritten by another kernel: L2-resident, not in any L1
        ld_kernel<MODE, N><<<sms, 256, 65536>>>(buf, out, cyc, 512);
        cudaDeviceSynchronize();
        if (cyc[0] < best) best = cyc[0];
    }
    printf("  %-8s N=%2d : %6lld cycles  [%s]\n", name, N, best, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);
    const size_t words = (size_t)sms * 65536;
    unsigned long long *buf, *out; long long* cyc;
    cudaMalloc(&buf, words * 8); cudaMalloc(&out, (size_t)sms * 256 * 8); cudaMallocManaged(&cyc, 64);
    printf("cycles from first issue to last use, 256 threads x N independent 16-byte loads per thread (L2 hits):\n");
#define ROW(N) run<0, N>("relaxed", buf, words, out, cyc, sms); run<1, N>("weak", buf, words, out, cyc, sms); \
               run<2, N>("nc", buf, words, out, cyc, sms); run<3, N>("cpasync", buf, words, out, cyc, sms);
    ROW(1) ROW(2) ROW(4) ROW(8) ROW(16)
    return 0;
}
