// tools/ubench_flag.cu - latency of a producer -> consumer hand-over between two warps of one CTA
// through a shared-memory flag (the chain link of kernels_dep.cuh), in cycles per hop:
//   mode 0: flag only (volatile store / volatile poll)
//   mode 1: + __threadfence_block() on both sides
//   mode 2: + a shared-memory value written before the flag and read after it
//   mode 3: + a GLOBAL value (st.global, fence, flag | poll, fence, ld.global) - the refactorization's link
//   mode 4: mode 3 with the value read through ld.global.cg (L2)
//   mode 5: mode 2 + warp-shuffle reduction over 8 lanes + dependent DFMA (the solve's link)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_flag.cu -o /tmp/ubench_flag
#include <cstdio>
#include <cuda_runtime.h>

__global__ void pingpong(int mode, int iters, int nspin, unsigned sleep_ns, double* g, long long* out) {
    __shared__ volatile unsigned flagA, flagB;
    __shared__ double sval[2];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { flagA = 0; flagB = 0; sval[0] = 1.0; sval[1] = 1.0; }
    __syncthreads();
    long long t0 = clock64();
    double acc = 1.0;
    if (w == 0 || w == 1) {
        volatile unsigned* mine = w == 0 ? &flagB : &flagA;   // the flag I wait for
        volatile unsigned* other = w == 0 ? &flagA : &flagB;  // the flag I set
        for (int it = 1; it <= iters; ++it) {
            if (w == 1 || it > 1) {
                while (*mine < (unsigned)(w == 0 ? it - 1 : it)) { if (sleep_ns) __nanosleep(sleep_ns); }
                if (mode >= 1) __threadfence_block();
                if (mode == 2 || mode == 5) acc += sval[w ^ 1];
                if (mode == 3) acc += g[(w ^ 1) * 32 + lane];
                if (mode == 4) acc += __ldcg(g + (w ^ 1) * 32 + lane);
                if (mode == 5) {
                    acc = fma(acc, 1.0000001, 0.5);
                    for (int o = 4; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, 8);
                }
            }
            if (mode == 2 || mode == 5) { if (lane == 0) sval[w] = acc * 0.5; }
            if (mode == 3 || mode == 4) g[w * 32 + lane] = acc * 0.5;
            if (mode >= 1) __threadfence_block();
            __syncwarp();
            if (lane == 0) *other = (unsigned)it;
        }
    } else {
        // bystanders: spin on a flag that never changes, like the waiting warps of the kernel
        if (w < 2 + nspin) {
            while (flagB < (unsigned)iters) { if (sleep_ns) __nanosleep(sleep_ns); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; g[100] = acc; }
}

int main() {
    double* g; long long* out;
    cudaMalloc(&g, 4096); cudaMemset(g, 0, 4096); cudaMalloc(&out, 64);
    const int iters = 20000;
    for (int mode = 0; mode <= 5; ++mode)
        for (int nspin : {0, 6, 30})
            for (unsigned sl : {0u, 40u, 200u}) {
                pingpong<<<1, 1024>>>(mode, iters, nspin, sl, g, out);
                cudaDeviceSynchronize();
                long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
                printf("mode %d bystanders %2d sleep %3u ns: %.0f cycles per hop\n", mode, nspin, sl, (double)c / (2.0 * iters));
            }
    return 0;
}
