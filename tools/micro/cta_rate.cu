// CTA dispatch rate micro-benchmark: how many 64-thread CTAs per microsecond can one kernel retire?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 4) k(float* out, int work) {
    extern __shared__ float sm[];
    float x = threadIdx.x;
    for (int i = 0; i < work; ++i) x = x * 1.0001f + 0.5f;
    if (x == 123.456f) out[0] = x + sm[0];
}
int main() {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads[] = {64, 128, 256};
    for (int ti = 0; ti < 3; ++ti)
    for (int smem = 0; smem <= 8192; smem += 8192)
    for (int work = 0; work <= 4000; work += 2000) {
        const int T = threads[ti];
        const int grid = 4096 * 64 / T * 16;
        for (int r = 0; r < 3; ++r) k<<<grid, T, smem>>>(d, work);
        cudaEventRecord(e0);
        for (int r = 0; r < 10; ++r) k<<<grid, T, smem>>>(d, work);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("threads %3d smem %5d work %4d: %.1f CTAs/us  (%.1f warps/us)\n", T, smem, work, 10.0 * grid / (ms * 1e3), 10.0 * grid * (T / 32) / (ms * 1e3));
    }
    return 0;
}
