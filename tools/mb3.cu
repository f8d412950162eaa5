// Co-issue microbenchmarks: 12 independent dp4a + N other ops per iteration (sm_100a).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__device__ __forceinline__ uint32_t dp(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t shf(uint32_t lo, uint32_t hi, uint32_t s) {
    uint32_t d;
    asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(s));
    return d;
}
__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lds(uint32_t addr) {
    uint32_t d;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(d) : "r"(addr));
    return d;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// KIND 0 none, 1 shf (3 reg inputs), 2 lop3, 3 lds, 4 imad, 5 shf with 2 distinct regs
template <int KIND, int N, int NDP>
__global__ void __launch_bounds__(256) kk(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i + seed;
    __syncthreads();
    uint32_t a[12], e[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { a[i] = seed + i * 77 + threadIdx.x; e[i] = seed * 5 + i + threadIdx.x; }
    uint32_t x = seed ^ threadIdx.x, y = seed * 3 + 1, z = threadIdx.x & 24;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * 4;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            if (i < NDP) a[i] = dp(x, y, a[i]);
            if (i < N) {
                if (KIND == 1) e[i] = shf(e[i], x, z);
                if (KIND == 2) e[i] = lop(e[i], x, y);
                if (KIND == 3) e[i] ^= lds(base + i * 1024 % 4096);
                if (KIND == 4) e[i] = imad(e[i], x, y);
                if (KIND == 5) e[i] = shf(e[i], e[i], z);
                if (KIND == 6) e[i] = (e[i] << 8) + x;                       // LEA
                if (KIND == 7) e[i] = e[i] + x + y;                          // IADD3
                if (KIND == 8) e[i] = __byte_perm(e[i], x, 0x3240);          // PRMT
                if (KIND == 9) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(e[i]) : "r"(e[i]), "r"(x), "r"(y));
                if (KIND == 10) e[i] = (uint32_t)__vimin_s32_relu((int32_t)e[i], 255) + 1;  // VIMNMX (+1 to keep a chain)
                if (KIND == 11) e[i] = (uint32_t)((int32_t)e[i] >> 6) ^ x;   // SHF.R.S32.HI imm + lop
                if (KIND == 12) e[i] = (uint32_t)((int32_t)(e[i] ^ x) >> 6); // same
                if (KIND == 13) e[i] = __funnelshift_r(e[i], x, 8);          // shf with immediate shift
                if (KIND == 14) e[i] = __byte_perm(e[i], x, z);              // PRMT with register selector
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) r ^= a[i] ^ e[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}
template <int KIND, int N, int NDP>
void run(const char *name, uint32_t *d) {
    const int blocks = 148 * 4, threads = 256;
    kk<KIND, N, NDP><<<blocks, threads>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kk<KIND, N, NDP><<<blocks, threads>>>(d, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double iters = (double)blocks * threads / 32 * ITER / 148;
    printf("%-34s dp=%2d other=%2d  %6.2f SM-cycles/iter\n", name, NDP, N, ms * 1e-3 * 1.965e9 / iters);
}
int main() {
    uint32_t *d;
    cudaMalloc(&d, 4096);
    run<0, 0, 12>("dp4a only", d);
    run<6, 12, 12>("+LEA", d); run<6, 12, 0>("LEA only", d);
    run<7, 12, 12>("+IADD3", d); run<7, 12, 0>("IADD3 only", d);
    run<8, 12, 12>("+PRMT imm", d); run<8, 12, 0>("PRMT only", d);
    run<14, 12, 12>("+PRMT reg", d);
    run<9, 12, 12>("+I2IP", d); run<9, 12, 0>("I2IP only", d);
    run<10, 12, 12>("+VIMNMX(+iadd)", d); run<10, 12, 0>("VIMNMX(+iadd) only", d);
    run<11, 12, 12>("+SHF.S32 imm (+lop)", d); run<11, 12, 0>("SHF.S32 imm (+lop) only", d);
    run<13, 12, 12>("+SHF funnel imm", d); run<13, 12, 0>("SHF funnel imm only", d);
    run<1, 4, 12>("+shf(3 regs)", d); run<1, 8, 12>("+shf(3 regs)", d); run<1, 12, 12>("+shf(3 regs)", d);
    run<5, 12, 12>("+shf(2 regs)", d);
    run<2, 4, 12>("+lop3", d); run<2, 8, 12>("+lop3", d); run<2, 12, 12>("+lop3", d);
    run<3, 4, 12>("+lds", d); run<3, 6, 12>("+lds", d); run<3, 12, 12>("+lds", d);
    run<4, 4, 12>("+imad", d); run<4, 12, 12>("+imad", d);
    run<1, 12, 0>("shf only", d); run<2, 12, 0>("lop3 only", d); run<3, 12, 0>("lds only", d);
    run<1, 12, 6>("6 dp + 12 shf", d); run<2, 12, 6>("6 dp + 12 lop3", d);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
