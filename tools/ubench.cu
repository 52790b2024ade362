// ubench.cu — micro-benchmarks behind the design of the register-resident kernel:
// FP64 issue rate with partly active warps, shared-memory broadcast loads, single-lane stores,
// shuffles, and the complex multiply-add + broadcast-load mix of the generated code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long clk() { return clock64(); }

// ---- 1. DFMA with a lane mask --------------------------------------------------------------
__global__ void k_dfma_mask(double* out, int iters, unsigned mask, double a, double b) {
    const int lane = threadIdx.x & 31;
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = threadIdx.x * 1e-9 + c;
    if ((mask >> lane) & 1u) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fma(acc[c], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[c];
    if (s == 123.456) out[0] = s;
}

// ---- 2. LDS.128: uniform address / per-lane address; STS.128 from one lane; SHFL ---------------
template <int MODE>
__global__ void k_lsu(double* out, int iters) {
    extern __shared__ double2 sm[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int idx = (MODE == 1) ? lane : 0;
    unsigned u = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (MODE == 0 || MODE == 1) {  // LDS.128 uniform / lane-distinct
                double2 v;
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sm + idx + c * 32)));
                acc.x += v.x;
                acc.y += v.y;
            } else if (MODE == 2) {  // STS.128 from lane 0 only
                if (lane == 0)
                    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"((unsigned)__cvta_generic_to_shared(sm + c * 32 + (threadIdx.x >> 5))), "d"(acc.x), "d"(acc.y) : "memory");
            } else if (MODE == 3) {  // SHFL.IDX from a uniform lane
                u = __shfl_sync(0xffffffffu, u, c) + 1;
            }
        }
    }
    if (acc.x + acc.y + u == 123.456) out[0] = acc.x;
}

// ---- 3. the mix: one uniform LDS.128 + complex multiply-add into NACC accumulators ------------
template <int NACC, bool WITH_LDS, bool WITH_STS>
__global__ void k_mix(double* out, int iters) {
    extern __shared__ double2 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(1e-3 * i, -1e-3 * i);
    __syncthreads();
    double2 acc[NACC];
#pragma unroll
    for (int c = 0; c < NACC; ++c) acc[c] = make_double2(lane + c, lane - c);
    double2 l = make_double2(1e-3 * lane, 1e-4);
    const double2* row = sm + warp * 64;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < NACC; ++c) {
            double2 u;
            if (WITH_LDS) {
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(u.x), "=d"(u.y) : "r"((unsigned)__cvta_generic_to_shared(row + c)));
            } else {
                u = make_double2(1.0 + c, 0.5);
            }
            acc[c].x = fma(-l.x, u.x, acc[c].x);
            acc[c].x = fma(l.y, u.y, acc[c].x);
            acc[c].y = fma(-l.x, u.y, acc[c].y);
            acc[c].y = fma(-l.y, u.x, acc[c].y);
            if (WITH_STS) {
                if (lane == (i & 31))
                    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"((unsigned)__cvta_generic_to_shared(sm + 2048 + warp * 64 + c)), "d"(acc[c].x), "d"(acc[c].y) : "memory");
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < NACC; ++c) s += acc[c].x + acc[c].y;
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const double clkhz = p.clockRate * 1e3;
    double* d;
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k_lsu<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_lsu<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_lsu<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_lsu<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_mix<16, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    cudaFuncSetAttribute(k_mix<16, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    cudaFuncSetAttribute(k_mix<16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    printf("device %s, %d SMs, %.0f MHz\n", p.name, sms, clkhz / 1e6);
    {
        const int iters = 20000;
        const unsigned masks[] = {0xffffffffu, 0x0000ffffu, 0x000000ffu, 0x0000000fu, 0x55555555u, 0x00ff00ffu, 0xffff0000u};
        for (int wps : {4, 8, 16}) {
            for (unsigned m : masks) {
                double ms = time_ms([&] { k_dfma_mask<<<sms, 32 * wps>>>(d, iters, m, 1.0000001, 1e-9); });
                double cyc = ms * 1e-3 * clkhz;
                printf("dfma mask %08x warps/SM %2d: %.2f cycles per warp-DFMA per sub-partition\n", m, wps,
                       cyc / (iters * 8.0 * wps / 4.0));
            }
        }
    }
    {
        const int iters = 20000;
        const char* names[] = {"LDS.128 uniform address", "LDS.128 lane-distinct", "STS.128 one lane", "SHFL.IDX"};
        for (int wps : {4, 8, 16}) {
            double ms[4];
            ms[0] = time_ms([&] { k_lsu<0><<<sms, 32 * wps, 65536>>>(d, iters); });
            ms[1] = time_ms([&] { k_lsu<1><<<sms, 32 * wps, 65536>>>(d, iters); });
            ms[2] = time_ms([&] { k_lsu<2><<<sms, 32 * wps, 65536>>>(d, iters); });
            ms[3] = time_ms([&] { k_lsu<3><<<sms, 32 * wps, 65536>>>(d, iters); });
            for (int q = 0; q < 4; ++q)
                printf("%-26s warps/SM %2d: %.2f cycles per warp-instruction per SM\n", names[q], wps,
                       ms[q] * 1e-3 * clkhz / (iters * 8.0 * wps));
        }
    }
    {
        const int iters = 5000;
        for (int wps : {4, 8, 12, 16}) {
            double a = time_ms([&] { k_mix<16, false, false><<<sms, 32 * wps, 98304>>>(d, iters); });
            double b = time_ms([&] { k_mix<16, true, false><<<sms, 32 * wps, 98304>>>(d, iters); });
            double c = time_ms([&] { k_mix<16, true, true><<<sms, 32 * wps, 98304>>>(d, iters); });
            const double per = iters * 16.0 * wps / 4.0;  // complex multiply-adds per sub-partition
            printf("complex multiply-add, warps/SM %2d: alone %.2f | +uniform LDS.128 %.2f | +LDS+one-lane STS.128 %.2f cycles per sub-partition (floor 4 DFMA)\n",
                   wps, a * 1e-3 * clkhz / per, b * 1e-3 * clkhz / per, c * 1e-3 * clkhz / per);
        }
    }
    cudaFree(d);
    return 0;
}
