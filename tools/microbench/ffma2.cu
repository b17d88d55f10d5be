// Microbenchmark: issue throughput of scalar FFMA/FADD/FMNMX vs packed FFMA2/FADD2 (sm_100 f32x2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b)
{
    float x[ILP]; float2 y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = make_float2(x[i], x[i] + 0.5f); }
    const float2 a2 = make_float2(a, a * 1.01f), b2 = make_float2(b, b * 0.99f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);                 // FFMA
            if (MODE == 1) y[i] = __ffma2_rn(y[i], a2, b2);         // FFMA2
            if (MODE == 2) x[i] = x[i] + a;                         // FADD
            if (MODE == 3) y[i] = __fadd2_rn(y[i], a2);             // FADD2
            if (MODE == 4) x[i] = fmaxf(x[i], a) * b;               // FMNMX + FMUL
            if (MODE == 5) { x[i] = fmaf(x[i], a, b); y[i].x = fmaxf(y[i].x, x[i]); }   // FFMA + FMNMX (two pipes)
            if (MODE == 6) y[i] = __fmul2_rn(y[i], a2);             // FMUL2
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i].x + y[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int flops_per_inst, int inst_per_iter)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8;
    float* out; cudaMalloc(&out, grid * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double warp_inst = (double)grid * 8 * ITERS * ILP * inst_per_iter;
    double per_clk_sm = warp_inst / (ms * 1e-3) / sms / 1.9e9;
    printf("%-14s %8.3f ms  %6.2f warp-inst/clk/SM (at 1.9 GHz)  %7.1f TFLOP/s\n", name, ms, per_clk_sm,
           warp_inst * 32 * flops_per_inst / inst_per_iter / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main()
{
    run<0>("FFMA", 2, 1);
    run<1>("FFMA2", 4, 1);
    run<2>("FADD", 1, 1);
    run<3>("FADD2", 2, 1);
    run<4>("FMNMX+FMUL", 2, 2);
    run<5>("FFMA+FMNMX", 3, 2);
    run<6>("FMUL2", 2, 1);
    return 0;
}
