// fp64_peak.cu — measures the FP64 DFMA roof of this GPU (BASELINE.md §2 asks for it: it
// is not in MEASURED_PEAKS.json).  Prints TFLOP/s at full occupancy, and the issue interval
// of dependent / independent DFMA chains in a single warp (latency numbers used in DESIGN.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-9 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == 123.456) out[0] = s;
}

template <int CHAINS>
double run(int blocks, int threads, int iters) {
    double* d;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_kernel<CHAINS><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dfma_kernel<CHAINS><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(d);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double clk = p.clockRate * 1e3;
    printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, clk / 1e6);
    {
        const int iters = 20000, threads = 1024, blocks = sms * 2;
        double ms = run<8>(blocks, threads, iters);
        double flops = 2.0 * 8 * iters * (double)threads * blocks;
        printf("full occupancy: %.2f TFLOP/s FP64 (DFMA), %.1f FMA/clk/SM at the max clock\n", flops / ms / 1e9,
               flops / 2 / (ms * 1e-3) / sms / clk);
    }
    {
        const int iters = 200000;
        double ms1 = run<1>(sms, 32, iters), ms8 = run<8>(sms, 32, iters);
        printf("one warp per SM: dependent DFMA every %.1f cycles; 8 independent chains: one DFMA every %.1f cycles\n",
               ms1 * 1e-3 * clk / iters, ms8 * 1e-3 * clk / iters / 8);
        double ms4w = run<8>(sms, 128, iters);
        printf("four warps per SM (one per sub-partition), 8 chains: one DFMA per warp every %.1f cycles\n",
               ms4w * 1e-3 * clk / iters / 8);
        double ms8w = run<8>(sms, 256, iters);
        printf("eight warps per SM, 8 chains: one DFMA per warp every %.1f cycles\n", ms8w * 1e-3 * clk / iters / 8);
    }
    return 0;
}
