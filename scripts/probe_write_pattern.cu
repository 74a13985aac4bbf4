// Micro-probe (not product): DRAM write efficiency of the NCHW tile pattern vs contiguous tiles.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
// mode 0: tile = (b, ix): 64 channel rows of 800 B, 160000 B apart.  mode 1: 51200 B contiguous per CTA.
// mode 2: tile = (b, 8 channels, 8 ix rows): 8 runs of 6400 B.  mode 3: (b, c) plane strips: 1 channel x 64 ix rows = 51200 B contiguous (same as 1 but plane-major order)
__global__ void __launch_bounds__(256) k_write(float *out, int mode, int ntiles) {
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (mode == 0) {
            const int b = tile / 200, ix = tile % 200;
            float *g = out + (size_t)b * 64 * 40000 + (size_t)ix * 200;
            for (int i = threadIdx.x; i < 64 * 50; i += 256) { const int r = i / 50, v = i % 50; *reinterpret_cast<float4 *>(g + (size_t)r * 40000 + v * 4) = make_float4(0, 0, 0, 0); }
        } else if (mode == 1) {
            float *g = out + (size_t)tile * 12800;
            for (int i = threadIdx.x; i < 3200; i += 256) *reinterpret_cast<float4 *>(g + i * 4) = make_float4(0, 0, 0, 0);
        } else if (mode == 2) {
            const int b = tile / 200, r = tile % 200, cg = r / 25, xg = r % 25;     // 8 channel groups x 25 ix groups
            float *g = out + ((size_t)b * 64 + cg * 8) * 40000 + (size_t)xg * 1600;
            for (int i = threadIdx.x; i < 8 * 400; i += 256) { const int c = i / 400, v = i % 400; *reinterpret_cast<float4 *>(g + (size_t)c * 40000 + v * 4) = make_float4(0, 0, 0, 0); }
        }
    }
}
int main() {
    const size_t N = (size_t)8 * 64 * 200 * 200;
    float *buf[4];
    for (int i = 0; i < 4; ++i) CK(cudaMalloc(&buf[i], N * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode)
        for (int grid : {1600, 592, 296, 148}) {
            for (int i = 0; i < 8; ++i) k_write<<<grid, 256>>>(buf[i % 4], mode, 1600);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            for (int i = 0; i < 200; ++i) k_write<<<grid, 256>>>(buf[i % 4], mode, 1600);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("mode %d grid %4d: %.2f us per 82 MB  (%.0f GB/s)\n", mode, grid, ms * 5, N * 4 / (ms * 5e-6) / 1e9);
        }
    return 0;
}
