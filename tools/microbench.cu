// Instruction-throughput microbenchmarks that size the tile kernel's inner loops (sm_100a).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

__device__ __forceinline__ int32_t dp4a_us(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t cvtpack(int32_t a, int32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// mode 0: 12 independent dp4a chains; 1: dp4a + funnel shift 3:1; 2: cvt.pack chains; 3: shf chains; 4: imad chains
// 5: dp4a 6 + LDS 3 + SHF 2 + 4 alu (the aligned inner loop mix); 6: LDS.32 only; 7: dp4a 9 + LDS 3 + 4 alu
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 2654435761u + seed;
    __syncthreads();
    uint32_t a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = seed + i * 77 + threadIdx.x;
    uint32_t x = seed ^ threadIdx.x, y = seed * 3 + 1, z = threadIdx.x & 24;
    const uint32_t *p = sm + (threadIdx.x & 31) + ((threadIdx.x >> 5) << 6);
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] = dp4a_uu(x, y, a[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t w = __funnelshift_r(a[8 + i], x, z);
                a[i] = dp4a_uu(w, y, a[i]);
                a[4 + i] = dp4a_uu(w, x, a[4 + i]);
                a[8 + i] = dp4a_us(w, y, a[8 + i]) | 1;
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] = cvtpack((int32_t)a[i], (int32_t)x, y);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] = __funnelshift_r(a[i], x, z);
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] = a[i] * x + y;
        } else if (MODE == 5 || MODE == 7) {
            // two samples per iteration
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const uint32_t w0 = p[(it + s * 40) & 1023], w1 = p[((it + s * 40) & 1023) + 32], w2 = p[((it + s * 40) & 1023) + 64];
                uint32_t a0 = 1u << 21, a1 = 0;
                int32_t a2 = 0;
                if (MODE == 5) {
                    const uint32_t u0 = __funnelshift_r(w0, w1, z), u1 = __funnelshift_r(w1, w2, z);
                    a0 = dp4a_uu(u0, x, a0); a1 = dp4a_uu(u0, y, a1); a2 = dp4a_us(u0, a[11], a2);
                    a0 = dp4a_uu(u1, a[10], a0); a1 = dp4a_uu(u1, a[9], a1); a2 = dp4a_us(u1, a[8], a2);
                } else {
                    a0 = dp4a_uu(w0, x, a0); a1 = dp4a_uu(w0, y, a1); a2 = dp4a_us(w0, a[11], a2);
                    a0 = dp4a_uu(w1, a[10], a0); a1 = dp4a_uu(w1, a[9], a1); a2 = dp4a_us(w1, a[8], a2);
                    a0 = dp4a_uu(w2, a[7], a0); a1 = dp4a_uu(w2, a[6], a1); a2 = dp4a_us(w2, a[5], a2);
                }
                const int32_t acc = (int32_t)(a0 + (a1 << 8) + ((uint32_t)a2 << 16));
                a[s] = cvtpack(acc >> 22, (int32_t)a[s], a[s + 2]);
            }
        } else if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] += p[(it + i * 32) & 1023];
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) r ^= a[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}


// ---- second family: one "sample" = NW aligned words of 4 taps, 3 byte planes -------------------
// VAR bit0: data from shared memory (else registers), bit1: funnel-shift alignment, bit2: epilogue (combine, shift, pack)
template <int NW, int VAR>
__global__ void __launch_bounds__(256) ks(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u + seed;
    __syncthreads();
    uint32_t k0[NW], k1[NW], k2[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) { k0[i] = seed + i; k1[i] = seed * 3 + i; k2[i] = seed * 7 + i; }
    const uint32_t z = (threadIdx.x & 3) * 8;
    const uint32_t *p = sm + (threadIdx.x & 31) + ((threadIdx.x >> 5) << 7);
    uint32_t acc = 0, r0 = seed, r1 = seed + 1;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        const uint32_t *q = p + (it & 7) * 128 * 0 + ((it & 15) << 4);
        uint32_t o[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {  // 4 samples (4 rows of one unit)
            uint32_t w[NW + 1];
#pragma unroll
            for (int i = 0; i <= NW; ++i) w[i] = (VAR & 1) ? q[s * 512 + i * 32] : (r0 + s * 3 + i) ^ r1;
            uint32_t a0 = 1u << 21, a1 = 0;
            int32_t a2 = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const uint32_t u = (VAR & 2) ? __funnelshift_r(w[i], w[i + 1], z) : w[i];
                a0 = dp4a_uu(u, k0[i], a0);
                a1 = dp4a_uu(u, k1[i], a1);
                a2 = dp4a_us(u, k2[i], a2);
            }
            if (VAR & 4) o[s] = (uint32_t)((int32_t)(a0 + (a1 << 8) + ((uint32_t)a2 << 16)) >> 22);
            else o[s] = a0 ^ a1 ^ (uint32_t)a2;
        }
        if (VAR & 4) acc += cvtpack((int32_t)o[1], (int32_t)o[0], cvtpack((int32_t)o[3], (int32_t)o[2], 0));
        else acc += o[0] ^ o[1] ^ o[2] ^ o[3];
        r0 += acc; r1 ^= r0;
    }
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

template <int NW, int VAR>
void runs(const char *name, uint32_t *d, int bps) {
    const int blocks = 148 * bps, threads = 256;
    ks<NW, VAR><<<blocks, threads>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    ks<NW, VAR><<<blocks, threads>>>(d, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double samples = (double)blocks * threads / 32 * ITER * 4;
    printf("%-52s bps=%d %8.3f ms  %6.2f SM-cycles per warp-sample @1.965GHz\n", name, bps, ms,
           (ms * 1e-3 * 1.965e9) / (samples / 148));
}

// pure LDS with immediate offsets
__global__ void __launch_bounds__(256) klds(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u + seed;
    __syncthreads();
    const uint32_t *p = sm + (threadIdx.x & 31) + ((threadIdx.x >> 5) << 7);
    uint32_t a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        const uint32_t *q = p + ((it & 15) << 4);
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] ^= q[i * 160];
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, double warp_instr_per_iter, uint32_t *d) {
    const int blocks = 148 * 4, threads = 256;
    k<MODE><<<blocks, threads>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32;
    const double winstr = warps * ITER * warp_instr_per_iter;
    // per SM per cycle at 1.965 GHz nominal (report both raw rate and assumed-clock figure)
    printf("%-44s %8.3f ms  %7.2f Gwarp-instr/s  = %5.2f warp-instr/clk/SM @1.965GHz (counted ops only)\n", name, ms,
           winstr / ms * 1e-6, winstr / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    uint32_t *d;
    cudaMalloc(&d, 4096);
    run<0>("dp4a x12 independent", 12, d);
    run<1>("4x(shf + 3 dp4a) [count 16]", 16, d);
    run<2>("cvt.pack.sat x12", 12, d);
    run<3>("shf.r x12", 12, d);
    run<4>("imad x12", 12, d);
    run<6>("lds.32 x12 (+12 iadd)", 12, d);
    run<5>("aligned sample x2 (6 dp4a,3 lds,2 shf,~5) [count 2]", 2, d);
    run<7>("misaligned sample x2 (9 dp4a,3 lds,~5) [count 2]", 2, d);

    for (int bps = 4; bps <= 8; bps += 4) {
        runs<2, 0>("NW2 regs, no funnel, no epilogue (6 dp4a)", d, bps);
        runs<2, 1>("NW2 lds(3), no funnel, no epilogue", d, bps);
        runs<2, 3>("NW2 lds(3), funnel, no epilogue", d, bps);
        runs<2, 7>("NW2 lds(3), funnel, epilogue  [aligned sample]", d, bps);
        runs<2, 4>("NW2 regs, no funnel, epilogue", d, bps);
        runs<3, 1>("NW3 lds(4), no funnel, no epilogue", d, bps);
        runs<3, 5>("NW3 lds(4), no funnel, epilogue [misaligned sample]", d, bps);
        runs<3, 7>("NW3 lds(4), funnel, epilogue", d, bps);
    }
    {
        klds<<<148 * 4, 256>>>(d, 1);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        klds<<<148 * 4, 256>>>(d, 2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("pure LDS.32 x16: %.3f ms -> %.2f LDS wavefronts/clk/SM\n", ms, 148.0 * 4 * 8 * ITER * 16 / (ms * 1e-3 * 1.965e9) / 148);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
