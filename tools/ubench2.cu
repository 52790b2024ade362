// ubench2.cu — issue-slot accounting of the FP64 pipe on B200: does a DFMA leave room for other
// instructions, what do LDS.128 (broadcast) / STS.128 (one lane) / SHFL cost next to it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu && ./ubench2
#include <cstdio>
#include <cuda_runtime.h>

// 8 independent DFMA chains + NALU integer ops per DFMA
template <int NALU>
__global__ void k_dfma_alu(double* out, int iters, double a, double b) {
    double acc[8];
    unsigned x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { acc[c] = threadIdx.x * 1e-9 + c; x[c] = threadIdx.x + c; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            acc[c] = fma(acc[c], a, b);
#pragma unroll
            for (int q = 0; q < NALU; ++q) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(x[(c + 1) & 7]), "r"(i));
        }
    }
    double s = 0;
    unsigned xs = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) { s += acc[c]; xs ^= x[c]; }
    if (s == 123.456 || xs == 0x12345u) out[0] = s;
}

// LSU-only loops (integer accumulation so that the FP64 pipe is not involved)
template <int MODE>
__global__ void k_lsu(unsigned* out, int iters) {
    extern __shared__ uint4 sm4[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm4[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    unsigned acc = 0, u = threadIdx.x;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm4) + ((MODE == 1) ? lane * 16 : 0);
    const uint4 val = make_uint4(lane, 1, 2, 3);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (MODE == 0 || MODE == 1) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + c * 512));
                acc ^= v.x;
            } else if (MODE == 2) {
                if (lane == 0)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(base + c * 512 + (threadIdx.x >> 5) * 16), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w));
            } else if (MODE == 3) {
                asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(u) : "r"(c));
            } else if (MODE == 4) {
                uint2 v;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(base + c * 512));
                acc ^= v.x;
            }
        }
    }
    if (acc + u == 0x12345u) out[0] = acc;
}

// complex multiply-add stream + a broadcast LDS.128 per multiply-add (+ an independent one-lane STS.128)
template <int MODE>  // 0 alone, 1 +LDS, 2 +LDS +STS (independent data), 3 +LDS.64 x2
__global__ void k_mix(double* out, int iters) {
    extern __shared__ double2 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(1e-3 * i, -1e-3 * i);
    __syncthreads();
    double2 acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = make_double2(lane + c, lane - c);
    const double2 l = make_double2(1e-3 * lane, 1e-4);
    const unsigned rowa = (unsigned)__cvta_generic_to_shared(sm + warp * 64);
    const unsigned pub = (unsigned)__cvta_generic_to_shared(sm + 2048 + warp * 64);
    const double2 pv = make_double2(lane, 2.0);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            double2 u;
            if (MODE >= 1) {
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(u.x), "=d"(u.y) : "r"(rowa + c * 16));
            } else {
                u = make_double2(1.0 + c, 0.5);
            }
            acc[c].x = fma(-l.x, u.x, acc[c].x);
            acc[c].x = fma(l.y, u.y, acc[c].x);
            acc[c].y = fma(-l.x, u.y, acc[c].y);
            acc[c].y = fma(-l.y, u.x, acc[c].y);
            if (MODE == 2) {
                if (lane == 7)
                    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(pub + c * 16), "d"(pv.x), "d"(pv.y));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) s += acc[c].x + acc[c].y;
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
    cudaFuncSetAttribute(k_lsu<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_mix<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    cudaFuncSetAttribute(k_mix<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    cudaFuncSetAttribute(k_mix<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    printf("device %s, %d SMs, %.0f MHz\n", p.name, sms, clkhz / 1e6);
    for (int wps : {4, 8, 16}) {
        const int iters = 20000;
        double t0 = time_ms([&] { k_dfma_alu<0><<<sms, 32 * wps>>>(d, iters, 1.0000001, 1e-9); });
        double t1 = time_ms([&] { k_dfma_alu<1><<<sms, 32 * wps>>>(d, iters, 1.0000001, 1e-9); });
        double t2 = time_ms([&] { k_dfma_alu<2><<<sms, 32 * wps>>>(d, iters, 1.0000001, 1e-9); });
        double t4 = time_ms([&] { k_dfma_alu<4><<<sms, 32 * wps>>>(d, iters, 1.0000001, 1e-9); });
        const double per = iters * 8.0 * wps / 4.0;
        printf("warps/SM %2d: cycles per DFMA per sub-partition: alone %.2f | +1 LOP3 %.2f | +2 LOP3 %.2f | +4 LOP3 %.2f\n", wps,
               t0 * 1e-3 * clkhz / per, t1 * 1e-3 * clkhz / per, t2 * 1e-3 * clkhz / per, t4 * 1e-3 * clkhz / per);
    }
    {
        const int iters = 20000;
        const char* names[] = {"LDS.128 uniform address", "LDS.128 lane-distinct", "STS.128 one lane", "SHFL.IDX", "LDS.64 uniform address"};
        for (int wps : {4, 8, 16}) {
            double ms[5];
            ms[0] = time_ms([&] { k_lsu<0><<<sms, 32 * wps, 65536>>>((unsigned*)d, iters); });
            ms[1] = time_ms([&] { k_lsu<1><<<sms, 32 * wps, 65536>>>((unsigned*)d, iters); });
            ms[2] = time_ms([&] { k_lsu<2><<<sms, 32 * wps, 65536>>>((unsigned*)d, iters); });
            ms[3] = time_ms([&] { k_lsu<3><<<sms, 32 * wps, 65536>>>((unsigned*)d, iters); });
            ms[4] = time_ms([&] { k_lsu<4><<<sms, 32 * wps, 65536>>>((unsigned*)d, iters); });
            for (int q = 0; q < 5; ++q)
                printf("%-26s warps/SM %2d: %.2f cycles per warp-instruction per SM\n", names[q], wps,
                       ms[q] * 1e-3 * clkhz / (iters * 8.0 * wps));
        }
    }
    {
        const int iters = 5000;
        for (int wps : {4, 8, 12, 16}) {
            double a = time_ms([&] { k_mix<0><<<sms, 32 * wps, 98304>>>(d, iters); });
            double b = time_ms([&] { k_mix<1><<<sms, 32 * wps, 98304>>>(d, iters); });
            double c = time_ms([&] { k_mix<2><<<sms, 32 * wps, 98304>>>(d, iters); });
            const double per = iters * 16.0 * wps / 4.0;
            printf("complex multiply-add, warps/SM %2d: alone %.2f | +uniform LDS.128 %.2f | +LDS +independent one-lane STS.128 %.2f cycles per sub-partition\n",
                   wps, a * 1e-3 * clkhz / per, b * 1e-3 * clkhz / per, c * 1e-3 * clkhz / per);
        }
    }
    cudaFree(d);
    return 0;
}
