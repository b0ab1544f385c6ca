// Micro-benchmark: issue rate of scalar vs packed (f32x2) FP32 instructions on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
// Prints warp-instructions per cycle per SM sub-partition for chains of independent operations.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
#define CHAINS 8
#define ITERS 2048

template <int OP>
__global__ void k(float *out, float seed, long long *cyc)
{
    float s[CHAINS];
    u64 p[CHAINS];
    for (int i = 0; i < CHAINS; ++i) {
        s[i] = seed + i + threadIdx.x;
        float2 t = make_float2(s[i], s[i] + 0.5f);
        p[i] = *reinterpret_cast<u64 *>(&t);
    }
    float c1 = seed * 0.999f, c2 = seed * 1e-3f;
    float2 t1 = make_float2(c1, c1), t2 = make_float2(c2, c2);
    u64 q1 = *reinterpret_cast<u64 *>(&t1), q2 = *reinterpret_cast<u64 *>(&t2);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(c1), "f"(c2));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(q1), "l"(q2));
            if (OP == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(c2));
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q2));
            if (OP == 4) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(c1));
            if (OP == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q1));
            if (OP == 6) {  // the recursion's mix, scalar: mul, add, fma
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(c1));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(c2));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(c1), "f"(c2));
            }
            if (OP == 7) {  // the same mix, packed
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q1));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q2));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(q1), "l"(q2));
            }
            if (OP == 8) {  // packed fma next to scalar integer work (alu pipe)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(q1), "l"(q2));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(c2));
            }
        }
    }
    long long t1c = clock64();
    float acc = 0.f;
    for (int i = 0; i < CHAINS; ++i) {
        float2 t = *reinterpret_cast<float2 *>(&p[i]);
        acc += s[i] + t.x + t.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1c - t0;
}

template <int OP>
void run(const char *name, int per_iter, float *d, long long *dc)
{
    for (int warps_per_smsp : {1, 2, 4}) {
        int threads = 128 * warps_per_smsp;
        k<OP><<<1, threads>>>(d, 1.0f, dc);
        k<OP><<<1, threads>>>(d, 1.0f, dc);
        long long c;
        cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        double inst = (double)ITERS * CHAINS * per_iter * warps_per_smsp;  // warp-instructions per sub-partition
        printf("%-28s warps/SMSP %d  cycles %8lld  inst/cycle/SMSP %.3f\n", name, warps_per_smsp, c, inst / c);
    }
}

int main()
{
    float *d;
    long long *dc;
    cudaMalloc(&d, 1 << 20);
    cudaMalloc(&dc, 8);
    run<0>("fma.f32 (3-reg)", 1, d, dc);
    run<1>("fma.f32x2", 1, d, dc);
    run<2>("add.f32", 1, d, dc);
    run<3>("add.f32x2", 1, d, dc);
    run<4>("mul.f32", 1, d, dc);
    run<5>("mul.f32x2", 1, d, dc);
    run<6>("mul+add+fma scalar", 3, d, dc);
    run<7>("mul+add+fma packed", 3, d, dc);
    run<8>("fma.f32x2 + add.f32", 2, d, dc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
