// dp4a dependency latency / per-warp issue rate on B200: C independent chains per warp, W warps per SM sub-partition.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mb_dp4a_latency tools/mb_dp4a_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void chains(unsigned *out, int iters, unsigned a0, unsigned b0) {
    unsigned acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = threadIdx.x + c;
    unsigned a = a0 + threadIdx.x, b = b0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = __dp4a(a, b, acc[c]);
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[c];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (unsigned)(t1 - t0);
    if (s == 0x12345678u) out[0] = s;
}
template <int C>
void run(int warps_per_sm, unsigned *d) {
    const int iters = 20000;
    chains<C><<<148, warps_per_sm * 32>>>(d, iters, 3, 5);
    cudaDeviceSynchronize();
    chains<C><<<148, warps_per_sm * 32>>>(d, iters, 3, 5);
    cudaDeviceSynchronize();
    unsigned h[2];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const double per = (double)h[1] / (iters * 8.0 * C);
    printf("chains %d  warps/SM %2d (per scheduler %d): %.2f cycles per dp4a per warp, %.2f dp4a/clk/scheduler\n", C, warps_per_sm,
           warps_per_sm / 4, per, (warps_per_sm / 4.0) / per);
}
int main() {
    unsigned *d;
    cudaMalloc(&d, 8);
    for (int w : {4, 8, 16}) {
        run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
