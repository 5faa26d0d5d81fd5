// f32x2_rate.cu - does packed fp32 (FFMA2 / FMUL2 / FADD2, sm_100) save ISSUE slots?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o f32x2_rate f32x2_rate.cu && ./f32x2_rate
//
// The race step kernel is issue-bound (profiles/README.md).  Three loop bodies doing the same arithmetic on 16
// independent fp32 chains per thread, 8 warps per SM sub-partition:
//   scalar : 16 FMUL + 16 FADD (separately rounded, what --fmad=false code issues)
//   packed : 8 FMUL2 + 8 FADD2
//   mixed  : the same plus 16 integer (LOP3/IADD) instructions per iteration, scalar vs packed
// Prints ns per iteration per warp and warp-instructions per cycle per sub-partition.
#include <cuda_runtime.h>
#include <stdio.h>

// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with --fmad=false; the fma identities below
// keep the two roundings apart (a*b + -0 == rn(a*b); a*1 + b == rn(a+b)) and still cost one packed issue each.
__device__ __forceinline__ unsigned long long pk(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 up(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
    return up(r);
}
__device__ __forceinline__ float2 xmul2(float2 a, float2 b) { return fma2(a, b, make_float2(-0.f, -0.f)); }
// (a * 1 + b with a LITERAL 1 is turned back into an add and contracted with the mul before it: the 1 must be opaque)
__device__ __forceinline__ float2 xadd2(float2 a, float2 b, float one) { return fma2(a, make_float2(one, one), b); }

template <int MODE>
__global__ void __launch_bounds__(256) body(float* out, int iters, float m, float a, float one)
{
    float2 x[8];
    unsigned u[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = threadIdx.x + i;
    const float2 m2 = make_float2(m, m), a2 = make_float2(a, a);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE & 1) {
                x[i] = xadd2(xmul2(x[i], m2), a2, one);
            } else {
                x[i].x = __fadd_rn(__fmul_rn(x[i].x, m), a);
                x[i].y = __fadd_rn(__fmul_rn(x[i].y, m), a);
            }
        }
        if (MODE & 4) {     // 12 integer instructions: neither the fp32 pipes nor the integer pipe saturate - issue decides
#pragma unroll
            for (int i = 0; i < 4; ++i) u[i] = (u[i] ^ (u[(i + 1) & 3] + it)) + (u[i] >> 3);
        }
        if (MODE & 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 4; ++i) u[i] = (u[i] ^ (u[(i + 1) & 3] + it)) + (u[i] >> 3);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    unsigned v = u[0] ^ u[1] ^ u[2] ^ u[3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)v;
}

template <int MODE>
static void run(const char* name, float* out, int iters, int fp_instr, int int_instr)
{
    const int blocks = 148 * 4;     // 4 CTAs of 256 threads per SM = 8 warps per sub-partition
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    body<MODE><<<blocks, 256>>>(out, iters, 0.999f, 0.001f, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    body<MODE><<<blocks, 256>>>(out, iters, 0.999f, 0.001f, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    const double per_iter_cycles = cycles / iters;                       // 8 warps per sub-partition share it
    const double ipc = 8.0 * (fp_instr + int_instr) / per_iter_cycles;
    printf("%-28s %8.3f ms  %7.1f cycles/iter (8 warps/SMSP)  issued/cycle/SMSP ~ %.2f  (%d fp + %d int instr per iter)\n",
           name, ms, per_iter_cycles, ipc, fp_instr, int_instr);
}

int main()
{
    float* out;
    cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
    const int iters = 20000;
    run<0>("scalar fp only", out, iters, 32, 0);
    run<1>("packed fp only", out, iters, 16, 0);
    run<2>("scalar fp + int", out, iters, 32, 48);
    run<3>("packed fp + int", out, iters, 16, 48);
    run<4>("scalar fp + few int", out, iters, 32, 12);
    run<5>("packed fp + few int", out, iters, 16, 12);
    printf("(clock = nominal max; ratios between the lines are what matters)\n");
    return 0;
}
