// Operand-pattern microbenchmarks for IDP.4A (register-file bandwidth vs pipe rate), sm_100a.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__device__ __forceinline__ uint32_t dp(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
struct Params { uint32_t k[24]; };

// MODE 0: a[i]=dp(x,y,a[i]); 1: dp(b[i],y,a[i]); 2: dp(b[i],c[i],a[i]);
// 3: sample pattern, 4 words u[4], 12 coefficient regs k, 12 accumulators: acc[p][..]
// 4: as 3 but coefficients from kernel params (uniform/constant operands)
// 5: as 3 but order by coefficient (k reused across 4 samples)
template <int MODE>
__global__ void __launch_bounds__(256) kk(uint32_t *out, uint32_t seed, Params prm) {
    uint32_t a[12], b[12], c[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { a[i] = seed + i * 77 + threadIdx.x; b[i] = seed * 5 + i + threadIdx.x; c[i] = seed * 9 + i * 3 + threadIdx.x; }
    uint32_t x = seed ^ threadIdx.x, y = seed * 3 + 1;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] = dp(x, y, a[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] = dp(b[i], y, a[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] = dp(b[i], c[i], a[i]);
        } else if (MODE == 3) {
            // 4 samples x 1 word x 3 planes: data b[s], coefficients c[0..2], acc a[s*3+p]
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int p = 0; p < 3; ++p) a[s * 3 + p] = dp(b[s], c[p], a[s * 3 + p]);
        } else if (MODE == 4) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int p = 0; p < 3; ++p) a[s * 3 + p] = dp(b[s], prm.k[p], a[s * 3 + p]);
        } else if (MODE == 5) {
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int s = 0; s < 4; ++s) a[s * 3 + p] = dp(b[s], c[p], a[s * 3 + p]);
        } else if (MODE == 6) {
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int s = 0; s < 4; ++s) a[s * 3 + p] = dp(b[s], prm.k[p], a[s * 3 + p]);
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) r ^= a[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, uint32_t *d) {
    Params prm;
    for (int i = 0; i < 24; ++i) prm.k[i] = 0x01020304u * (i + 1);
    const int blocks = 148 * 4, threads = 256;
    kk<MODE><<<blocks, threads>>>(d, 1, prm);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kk<MODE><<<blocks, threads>>>(d, 2, prm);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double winstr = (double)blocks * threads / 32 * ITER * 12;
    printf("%-60s %7.3f ms  %5.2f dp4a/clk/SM\n", name, ms, winstr / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    uint32_t *d;
    cudaMalloc(&d, 4096);
    run<0>("dp(x,y,a[i])   1 varying", d);
    run<1>("dp(b[i],y,a[i]) 2 varying", d);
    run<2>("dp(b[i],c[i],a[i]) 3 varying", d);
    run<3>("sample-major: dp(b[s],c[p],a[s][p])", d);
    run<4>("sample-major, coefficients from params (uniform)", d);
    run<5>("plane-major: k reused over 4 samples", d);
    run<6>("plane-major, coefficients from params", d);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
