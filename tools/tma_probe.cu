// Probe of the TMA pieces the tile kernel relies on (sm_100a): bulk-group bookkeeping with no copies,
// 2-D loads with the 128-byte swizzle (checks the ct_off() address formula), 2-D stores, partial boxes.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../image_transformation_b200/csrc/kernels.cuh"
#include "../image_transformation_b200/csrc/tile_kernel.cuh"
using namespace b200comp;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_groups() {
    if (threadIdx.x == 0) {
        bulk_wait_read<1>();
        bulk_commit();
        bulk_wait_read<3>();
        bulk_commit();
        bulk_wait_all();
    }
}

// load one tile (two swizzled halves) at (x0, y0), copy it out through ct_off(), store it to `out` with TMA
__global__ void k_tile(const CUtensorMap *in_map, const CUtensorMap *out_map, int x0, int y0, int tw, uint32_t *dump) {
    extern __shared__ uint32_t raw[];
    uint32_t *ct = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(raw) + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    uint64_t *bar = reinterpret_cast<uint64_t *>(ct + kTileWords);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_async_smem();
        mbar_expect_tx(bar, tw > 32 ? 8192u : 4096u);
        tma_load_2d(ct, in_map, x0, y0, bar);
        if (tw > 32) tma_load_2d(ct + 1024, in_map, x0 + 32, y0, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < kTileWords; i += blockDim.x) {
        const int r = i >> 6, x = i & 63;
        dump[i] = (x < tw) ? ct[ct_off(r, x)] : 0u;
    }
    // modify through the generic proxy, then store
    for (int i = threadIdx.x; i < kTileWords; i += blockDim.x) {
        const int r = i >> 6, x = i & 63;
        if (x < tw) ct[ct_off(r, x)] ^= 0x01010101u;
    }
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_store_2d(out_map, x0, y0, ct);
        if (tw > 32) tma_store_2d(out_map, x0 + 32, y0, ct + 1024);
        bulk_commit();
        bulk_wait_all();
    }
}

// unswizzled 64x32 box at signed coordinates (identity overlays), into a 128-byte aligned buffer
__global__ void k_overlay(const CUtensorMap *map, int x0, int y0, uint32_t *dump) {
    extern __shared__ uint32_t raw[];
    uint32_t *P = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(raw) + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    uint64_t *bar = reinterpret_cast<uint64_t *>(P + kTileWords);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_async_smem();
        mbar_expect_tx(bar, 8192u);
        tma_load_2d(P, map, x0, y0, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < kTileWords; i += blockDim.x) dump[i] = P[i];
}

// the command ring: cp.async 16-byte quarters, wait_group bookkeeping
__global__ void k_ring(const Cmd *stream, int n, uint32_t *sum) {
    __shared__ __align__(16) Cmd ring[kRing];
    const int tid = threadIdx.x;
    if (tid < 4) {
        for (int i = 0; i < kRingAhead; ++i) {
            cp_async16(reinterpret_cast<uint8_t *>(ring + i) + 16 * tid, reinterpret_cast<const uint8_t *>(stream + i) + 16 * tid);
            cp_async_commit();
        }
    }
    uint32_t acc = 0;
    for (int pos = 0; pos < n; ++pos) {
        if (tid < 4) {
            cp_async_wait<kRingAhead - 1 - kLook>();
            cp_async16(reinterpret_cast<uint8_t *>(ring + ((pos + kRingAhead) & (kRing - 1))) + 16 * tid,
                       reinterpret_cast<const uint8_t *>(stream + pos + kRingAhead) + 16 * tid);
            cp_async_commit();
        }
        __syncthreads();
        for (int a = 0; a <= kLook; ++a) acc += ring[(pos + a) & (kRing - 1)].w[tid & 15] * (uint32_t)(a + 1);
    }
    if (tid < 4) cp_async_wait<0>();
    atomicAdd(sum, acc);
}

int main() {
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    k_groups<<<1, 32>>>();
    CK(cudaDeviceSynchronize());
    printf("ok   empty bulk groups\n");

    const int W = 492, H = 200;
    const size_t pitch = (size_t)W * 4;  // 1968 = 16 * 123
    std::vector<uint32_t> h((size_t)W * H);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[(size_t)y * W + x] = (uint32_t)(y * 1000 + x);
    uint32_t *d_in, *d_out, *d_dump;
    CUtensorMap *d_maps;
    CK(cudaMalloc(&d_in, pitch * H)); CK(cudaMalloc(&d_out, pitch * H)); CK(cudaMalloc(&d_dump, kTileWords * 4));
    CK(cudaMalloc(&d_maps, 2 * sizeof(CUtensorMap)));
    CK(cudaMemcpy(d_in, h.data(), pitch * H, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0, pitch * H));
    CUtensorMap hm[2];
    for (int i = 0; i < 2; ++i) {
        const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H};
        const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
        const cuuint32_t box[2] = {32, 32};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&hm[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, i ? (void *)d_out : (void *)d_in, gdim, gstride, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d -> %d\n", i, (int)r);
    }
    CK(cudaMemcpy(d_maps, hm, sizeof hm, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    const int cases[3][3] = {{64, 32, 64}, {448, 192, 44}, {128, 0, 64}};  // (x0, y0, tw): interior, right/bottom edge, top
    for (auto &c : cases) {
        k_tile<<<1, 128, kTileWords * 4 + 1024 + 64>>>(d_maps, d_maps + 1, c[0], c[1], c[2], d_dump);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAIL tile (%d,%d): %s\n", c[0], c[1], cudaGetErrorString(e)); return 1; }
        std::vector<uint32_t> dump(kTileWords), out((size_t)W * H);
        CK(cudaMemcpy(dump.data(), d_dump, kTileWords * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.data(), d_out, pitch * H, cudaMemcpyDeviceToHost));
        int bad_l = 0, bad_s = 0;
        for (int r = 0; r < 32; ++r) for (int x = 0; x < c[2]; ++x) {
            const int gy = c[1] + r, gx = c[0] + x;
            const uint32_t want = (gy < H && gx < W) ? h[(size_t)gy * W + gx] : 0u;
            if (dump[r * 64 + x] != want) ++bad_l;
            if (gy < H && gx < W && out[(size_t)gy * W + gx] != (want ^ 0x01010101u)) ++bad_s;
        }
        printf("%s tile (%d,%d) tw=%d: load mismatches %d, store mismatches %d\n", (bad_l || bad_s) ? "FAIL" : "ok  ", c[0], c[1], c[2], bad_l, bad_s);
    }
    {   // overlay boxes: 131x32 overlay, pitch 528; boxes at negative / overhanging coordinates; 2x2 overlay
        const int ow = 131, oh = 32;
        const size_t op = 528;
        std::vector<uint32_t> ho(op / 4 * oh);
        for (int y = 0; y < oh; ++y) for (int x = 0; x < ow; ++x) ho[(size_t)y * (op / 4) + x] = (uint32_t)(y * 1000 + x + 7);
        uint32_t *d_o;
        CK(cudaMalloc(&d_o, op * oh));
        CK(cudaMemcpy(d_o, ho.data(), op * oh, cudaMemcpyHostToDevice));
        CUtensorMap om[2];
        const cuuint32_t box[2] = {64, 32};
        const cuuint32_t es[2] = {1, 1};
        {
            const cuuint64_t gdim[2] = {(cuuint64_t)ow, (cuuint64_t)oh};
            const cuuint64_t gstride[1] = {(cuuint64_t)op};
            printf("encode overlay -> %d\n", (int)enc(&om[0], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_o, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
            const cuuint64_t gdim2[2] = {2, 2};
            const cuuint64_t gstride2[1] = {16};
            printf("encode 2x2 overlay -> %d\n", (int)enc(&om[1], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_o, gdim2, gstride2, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
        }
        CK(cudaMemcpy(d_maps, om, sizeof om, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(k_overlay, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
        const int oc[4][3] = {{0, 0, 0}, {-20, -10, 0}, {100, 16, 0}, {-4, -4, 1}};
        for (auto &c : oc) {
            k_overlay<<<1, 128, kTileWords * 4 + 1024 + 64>>>(d_maps + c[2], c[0], c[1], d_dump);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("FAIL overlay (%d,%d) map %d: %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
            std::vector<uint32_t> dump(kTileWords);
            CK(cudaMemcpy(dump.data(), d_dump, kTileWords * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            const int w_ = c[2] ? 2 : ow, h_ = c[2] ? 2 : oh;
            const size_t p_ = c[2] ? 4 : op / 4;
            for (int r = 0; r < 32; ++r) for (int x = 0; x < 64; ++x) {
                const int gy = c[1] + r, gx = c[0] + x;
                const uint32_t want = (gy >= 0 && gx >= 0 && gy < h_ && gx < w_) ? ho[(size_t)gy * p_ + gx] : 0u;
                if (dump[r * 64 + x] != want) ++bad;
            }
            printf("%s overlay (%d,%d) map %d: mismatches %d\n", bad ? "FAIL" : "ok  ", c[0], c[1], c[2], bad);
        }
    }
    {   // ring
        const int n = 100;
        std::vector<Cmd> hs(n + kRing);
        for (size_t i = 0; i < hs.size(); ++i) for (int k = 0; k < 16; ++k) hs[i].w[k] = (uint32_t)(i * 16 + k);
        Cmd *d_s; uint32_t *d_sum;
        CK(cudaMalloc(&d_s, hs.size() * sizeof(Cmd))); CK(cudaMalloc(&d_sum, 4));
        CK(cudaMemcpy(d_s, hs.data(), hs.size() * sizeof(Cmd), cudaMemcpyHostToDevice));
        CK(cudaMemset(d_sum, 0, 4));
        k_ring<<<1, 128>>>(d_s, n, d_sum);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAIL ring: %s\n", cudaGetErrorString(e)); return 1; }
        uint32_t got = 0, want = 0;
        CK(cudaMemcpy(&got, d_sum, 4, cudaMemcpyDeviceToHost));
        for (int tid = 0; tid < 128; ++tid) for (int pos = 0; pos < n; ++pos) for (int a = 0; a <= kLook; ++a) want += hs[pos + a].w[tid & 15] * (uint32_t)(a + 1);
        printf("%s ring: got %u want %u\n", got == want ? "ok  " : "FAIL", got, want);
    }
    return 0;
}
