// ubench3.cu — LSU costs that decide how a pivot row is broadcast inside a warp:
// SHFL.IDX, STS.{32,64,128} from one lane / four lanes / all lanes, LDS.128 broadcast / per-class.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench3 ubench3.cu && ./ubench3
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k_lsu(unsigned* out, int iters) {
    extern __shared__ uint4 sm4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm4[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    unsigned acc = 0, u0 = threadIdx.x, u1 = threadIdx.x * 3, u2 = 7, u3 = 11;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm4) + warp * 2048;
    if (MODE == 1) base += lane * 16;           // lane-distinct, conflict-free
    if (MODE == 2) base += (lane & 3) * 16;     // four classes (2-D broadcast)
    for (int i = 0; i < iters; ++i) {
        const unsigned a = base + ((i & 7) << 8);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (MODE <= 2) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a + c * 16 * (MODE == 1 ? 0 : 1)) : "memory");
                acc ^= v.x ^ v.w;
            } else if (MODE == 3) {  // STS.128 one lane
                if (lane == 5) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a + c * 16), "r"(u0), "r"(u1), "r"(u2), "r"(u3) : "memory");
            } else if (MODE == 4) {  // STS.64 one lane
                if (lane == 5) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a + c * 16), "r"(u0), "r"(u1) : "memory");
            } else if (MODE == 5) {  // STS.32 one lane
                if (lane == 5) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + c * 16), "r"(u0) : "memory");
            } else if (MODE == 6) {  // STS.128 all lanes
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a + lane * 16), "r"(u0), "r"(u1), "r"(u2), "r"(u3) : "memory");
            } else if (MODE == 7) {  // STS.128 from four lanes (one per class)
                if ((lane >> 2) == 3) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a + c * 64 + (lane & 3) * 16), "r"(u0), "r"(u1), "r"(u2), "r"(u3) : "memory");
            } else if (MODE == 8) {  // 4 independent SHFL.IDX from a uniform lane (one complex value)
                unsigned t0, t1, t2, t3;
                asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(t0) : "r"(u0), "r"((i + c) & 31));
                asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(t1) : "r"(u1), "r"((i + c) & 31));
                asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(t2) : "r"(u2), "r"((i + c) & 31));
                asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(t3) : "r"(u3), "r"((i + c) & 31));
                acc ^= t0 ^ t1 ^ t2 ^ t3;
            } else if (MODE == 9) {  // SHFL.IDX with a lane-dependent source (2-D broadcast)
                unsigned t0;
                asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(t0) : "r"(u0), "r"(((i + c) & 7) * 4 + (lane & 3)));
                acc ^= t0;
            }
        }
    }
    if (acc == 0x12345u) out[0] = acc;
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

#define RUN(M, NAME, PER)                                                                              \
    {                                                                                                  \
        cudaFuncSetAttribute(k_lsu<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);            \
        for (int wps : {4, 8, 16}) {                                                                   \
            double ms = time_ms([&] { k_lsu<M><<<sms, 32 * wps, 65536>>>((unsigned*)d, iters); });     \
            printf("%-44s warps/SM %2d: %.2f cycles per warp-instruction per SM\n", NAME, wps,         \
                   ms * 1e-3 * clkhz / (iters * 8.0 * PER * wps));                                     \
        }                                                                                              \
    }

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const double clkhz = p.clockRate * 1e3;
    double* d;
    cudaMalloc(&d, 64);
    const int iters = 20000;
    printf("device %s, %d SMs, %.0f MHz\n", p.name, sms, clkhz / 1e6);
    RUN(0, "LDS.128 one address for the warp", 1)
    RUN(1, "LDS.128 lane-distinct (512 B)", 1)
    RUN(2, "LDS.128 four addresses (lane & 3)", 1)
    RUN(3, "STS.128 one lane", 1)
    RUN(4, "STS.64 one lane", 1)
    RUN(5, "STS.32 one lane", 1)
    RUN(6, "STS.128 all lanes (512 B)", 1)
    RUN(7, "STS.128 four lanes", 1)
    RUN(8, "SHFL.IDX uniform source (x4 per value)", 4)
    RUN(9, "SHFL.IDX lane-dependent source", 1)
    cudaFree(d);
    return 0;
}
