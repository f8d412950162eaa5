// Shared-memory atomic throughput on sm_100a: how many warp-wide ATOMS an SM retires per clock for
//   mode 0: conflict-free addresses (word = bin * 32 + lane: bank = lane), histogram shared by the CTA's warps
//   mode 1: random bins of a warp-private 768-word histogram (what hist_rgb_kernel does on noisy input)
//   mode 2: mode 0 with a plain load / add / store instead of the atomic (LSU reference, racy across warps)
//   mode 3: conflict-free, but two lanes per bank (16 copies)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mb_atoms tools/mb_atoms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(unsigned *out, int iters) {
    extern __shared__ unsigned h[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 768 * 32; i += blockDim.x) h[i] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    unsigned *mine = h + warp * 768;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const unsigned b0 = (x >> 24), b1 = 256 + ((x >> 16) & 255), b2 = 512 + ((x >> 8) & 255);
        if (MODE == 0) {
            atomicAdd(&h[b0 * 32 + lane], 1u); atomicAdd(&h[b1 * 32 + lane], 1u); atomicAdd(&h[b2 * 32 + lane], 1u);
        } else if (MODE == 1) {
            atomicAdd(&mine[b0], 1u); atomicAdd(&mine[b1], 1u); atomicAdd(&mine[b2], 1u);
        } else if (MODE == 2) {
            volatile unsigned *v = h;
            v[b0 * 32 + lane] = v[b0 * 32 + lane] + 1; v[b1 * 32 + lane] = v[b1 * 32 + lane] + 1; v[b2 * 32 + lane] = v[b2 * 32 + lane] + 1;
        } else {
            const int l2 = lane >> 1;
            atomicAdd(&h[b0 * 16 + l2], 1u); atomicAdd(&h[b1 * 16 + l2], 1u); atomicAdd(&h[b2 * 16 + l2], 1u);
        }
    }
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i < 768 * 32; i += blockDim.x) s += h[i];
    if (s == 0xffffffffu) out[0] = s;
}

template <int MODE>
void run(const char *name, int ctas_per_sm) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned *d;
    cudaMalloc(&d, 4);
    const int iters = 20000;
    const size_t smem = 768 * 32 * 4;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * ctas_per_sm, 512, smem>>>(d, 100);
    cudaEventRecord(e0);
    k<MODE><<<sms * ctas_per_sm, 512, smem>>>(d, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clocks = ms * 1e-3 * khz * 1e3;
    const double warp_atoms_per_sm = (double)ctas_per_sm * 16 * iters * 3;
    printf("%-58s CTAs/SM %d: %.3f warp-wide updates per clock per SM (%.1f lane updates/clk/SM), %.2f ms, err %s\n", name, ctas_per_sm,
           warp_atoms_per_sm / clocks, 32 * warp_atoms_per_sm / clocks, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    for (int c = 1; c <= 2; ++c) {
        run<0>("atomics, conflict-free (bank = lane)", c);
        run<1>("atomics, random bins of a warp-private histogram", c);
        run<2>("plain load/add/store, conflict-free", c);
        run<3>("atomics, 16 copies (two lanes per bank)", c);
    }
    return 0;
}
