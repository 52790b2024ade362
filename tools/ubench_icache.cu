// ubench_icache.cu — how fast does an SM run straight-line code that does not fit its instruction
// caches?  N_BODY independent FFMA instructions (320 KB of SASS at 20480) executed once per outer
// iteration, with 1..32 warps per SM (one CTA per SM).  Reports instructions per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_icache ubench_icache.cu && ./ubench_icache
#include <cstdio>
#include <cuda_runtime.h>

template <int BODY>
__global__ void k_line(float* out, int outer, float a, float b) {
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = threadIdx.x * 1e-6f + c;
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int i = 0; i < BODY / 16; ++i) {
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = fmaf(acc[c], a, b + (float)(i * 16 + c));  // distinct immediates: no loop rolling
        }
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) s += acc[c];
    if (s == 123.456f) out[0] = s;
}

template <int BODY>
static void run(int sms, double clkhz, float* d) {
    for (int wps : {1, 2, 4, 8, 16, 32}) {
        const int outer = 64;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k_line<BODY><<<sms, 32 * wps>>>(d, outer, 1.0000001f, 1e-9f);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_line<BODY><<<sms, 32 * wps>>>(d, outer, 1.0000001f, 1e-9f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double cyc = ms * 1e-3 * clkhz, inst = (double)BODY * outer * wps;
        printf("body %6d instructions (%4d KB), %2d warps/SM: %.3f warp-instructions per cycle per SM (%.1f B of code per cycle)\n", BODY,
               BODY * 16 / 1024, wps, inst / cyc, inst / cyc * 16.0);
    }
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* d;
    cudaMalloc(&d, 64);
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    run<512>(p.multiProcessorCount, p.clockRate * 1e3, d);
    run<4096>(p.multiProcessorCount, p.clockRate * 1e3, d);
    run<20480>(p.multiProcessorCount, p.clockRate * 1e3, d);
    return 0;
}
