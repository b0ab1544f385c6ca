// Micro-benchmark: HBM read bandwidth when a frame is read in column stripes of P bytes per row
// (the columns pass reads 128-byte pieces of rows that lie one pitch apart), against a linear read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stripe_bw stripe_bw.cu && ./stripe_bw
#include <cstdio>
#include <cuda_runtime.h>

// one warp per (plane, stripe): streams its stripe top to bottom, UNROLL independent 16-byte loads per lane
template <int P, int UNROLL>
__global__ void k_stripe(const float4 *base, int pitch16, int h, int stripes, int planes, float *sink)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int plane = warp / stripes, stripe = warp % stripes;
    if (plane >= planes) return;
    constexpr int LPR = P / 16;            // lanes per row
    constexpr int RPI = 32 / LPR;          // rows per instruction
    const float4 *p = base + (size_t)plane * pitch16 * h + (size_t)stripe * LPR + (lane % LPR) + (size_t)(lane / LPR) * pitch16;
    float acc = 0.f;
    for (int r = 0; r + RPI * UNROLL <= h; r += RPI * UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(p + (size_t)(r + u * RPI) * pitch16);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].w;
    }
    if (acc == 12345.678f) sink[0] = acc;
}

__global__ void k_linear(const float4 *base, size_t n16, float *sink)
{
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(base + i);
        acc += v.x + v.w;
    }
    if (acc == 12345.678f) sink[0] = acc;
}

// rows-pass pattern: one warp per (plane, 32-row block) walks along the rows; per step it reads a
// 32-row x 128-byte tile (4 rows per instruction, 8 lanes per row) and writes WOUT such tiles to WOUT
// other planes as whole 128-byte lines (streaming stores), UNROLL steps in flight.
template <int WOUT, int UNROLL>
__global__ void k_rows(const float4 *src, float4 *dst, int pitch16, int h, int planes)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int blocks = h / 32, plane = warp / blocks, rb = warp % blocks;
    if (plane >= planes) return;
    const size_t pbytes16 = (size_t)pitch16 * h;
    const float4 *p = src + plane * pbytes16 + (size_t)(rb * 32 + (lane >> 3)) * pitch16 + (lane & 7);
    float4 *q = dst + (size_t)plane * WOUT * pbytes16 + (size_t)(rb * 32 + (lane >> 3)) * pitch16 + (lane & 7);
    for (int c = 0; c + 8 * UNROLL <= pitch16; c += 8 * UNROLL) {
        float4 v[UNROLL][8];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[u][i] = __ldcs(p + c + 8 * u + (size_t)(4 * i) * pitch16);
#pragma unroll
        for (int o = 0; o < WOUT; ++o)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) __stcs(q + o * pbytes16 + c + 8 * u + (size_t)(4 * i) * pitch16, v[u][i]);
    }
}

// the same walk, but tiles live in HBM as contiguous 4 KB blocks ([row block][column block][32 rows][128 B]):
// BIN: the source plane is blocked too; the destination always is.  One instruction moves 512 contiguous bytes.
template <int WOUT, int UNROLL, bool BIN>
__global__ void k_rows_blocked(const float4 *src, float4 *dst, int pitch16, int h, int planes)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int blocks = h / 32, plane = warp / blocks, rb = warp % blocks;
    if (plane >= planes) return;
    const size_t pbytes16 = (size_t)pitch16 * h;
    const int tiles_x = pitch16 / 8;
    const float4 *p = src + plane * pbytes16 + (size_t)(rb * 32 + (lane >> 3)) * pitch16 + (lane & 7);
    const float4 *pb = src + plane * pbytes16 + (size_t)rb * tiles_x * 256 + lane;
    float4 *q = dst + (size_t)plane * WOUT * pbytes16 + (size_t)rb * tiles_x * 256 + lane;
    for (int c = 0; c + UNROLL <= tiles_x; c += UNROLL) {
        float4 v[UNROLL][8];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                v[u][i] = BIN ? __ldcs(pb + (size_t)(c + u) * 256 + 32 * i) : __ldcs(p + 8 * (c + u) + (size_t)(4 * i) * pitch16);
#pragma unroll
        for (int o = 0; o < WOUT; ++o)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) __stcs(q + o * pbytes16 + (size_t)(c + u) * 256 + 32 * i, v[u][i]);
    }
}

template <int WOUT, int UNROLL, bool BIN>
void run_rows_blocked(const float4 *src, float4 *dst, int w, int h, int planes)
{
    const int pitch16 = w / 4, warps = planes * (h / 32), threads = 128;
    const int blocks = (warps * 32 + threads - 1) / threads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_rows_blocked<WOUT, UNROLL, BIN><<<blocks, threads>>>(src, dst, pitch16, h, planes);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k_rows_blocked<WOUT, UNROLL, BIN><<<blocks, threads>>>(src, dst, pitch16, h, planes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 5.0 * planes * (1 + WOUT) * (double)w * h * 4;
    printf("rows pattern, 4 KB-blocked %s: read 1 plane, write %d, unroll %d, warps %5d : %7.1f GB/s (read + write)\n",
           BIN ? "in+out" : "out   ", WOUT, UNROLL, warps, bytes / ms / 1e6);
}

template <int WOUT, int UNROLL>
void run_rows(const float4 *src, float4 *dst, int w, int h, int planes)
{
    const int pitch16 = w / 4, warps = planes * (h / 32), threads = 128;
    const int blocks = (warps * 32 + threads - 1) / threads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_rows<WOUT, UNROLL><<<blocks, threads>>>(src, dst, pitch16, h, planes);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k_rows<WOUT, UNROLL><<<blocks, threads>>>(src, dst, pitch16, h, planes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 5.0 * planes * (1 + WOUT) * (double)w * h * 4;
    printf("rows pattern: read 1 plane, write %d, unroll %d, warps %5d : %7.1f GB/s (read + write)\n", WOUT, UNROLL, warps,
           bytes / ms / 1e6);
}

template <int P, int UNROLL>
void run(const float4 *d, int w, int h, int planes, float *sink, int threads)
{
    const int pitch16 = w / 4, stripes = w * 4 / P;
    const int warps = planes * stripes;
    const int blocks = (warps * 32 + threads - 1) / threads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_stripe<P, UNROLL><<<blocks, threads>>>(d, pitch16, h, stripes, planes, sink);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k_stripe<P, UNROLL><<<blocks, threads>>>(d, pitch16, h, stripes, planes, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 5.0 * planes * (double)w * h * 4;
    printf("stripe %4d B/row  unroll %2d  block %4d  warps %6d : %7.1f GB/s\n", P, UNROLL, threads, warps, bytes / ms / 1e6);
}

int main()
{
    const int w = 3840, h = 2160, planes = 15;          // 15 f32 planes = 498 MB > L2
    const size_t n16 = (size_t)planes * w * h / 4;
    float4 *d;
    float *sink;
    cudaMalloc(&d, n16 * 16);
    cudaMalloc(&sink, 4);
    cudaMemset(d, 0, n16 * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_linear<<<148 * 8, 512>>>(d, n16, sink);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k_linear<<<148 * 8, 512>>>(d, n16, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("linear read                                          : %7.1f GB/s\n", 5.0 * n16 * 16 / ms / 1e6);
    run<128, 4>(d, w, h, planes, sink, 32);
    run<128, 8>(d, w, h, planes, sink, 32);
    run<128, 8>(d, w, h, planes, sink, 128);
    run<128, 16>(d, w, h, planes, sink, 128);
    run<256, 4>(d, w, h, planes, sink, 32);
    run<256, 8>(d, w, h, planes, sink, 32);
    run<256, 8>(d, w, h, planes, sink, 128);
    run<256, 16>(d, w, h, planes, sink, 128);
    run<512, 8>(d, w, h, planes, sink, 32);
    run<512, 8>(d, w, h, planes, sink, 128);
    run<512, 16>(d, w, h, planes, sink, 128);
    {   // rows-pass pattern: 6 source planes -> 6 * WOUT destination planes
        float4 *dst;
        const int rp = 6;
        cudaMalloc(&dst, (size_t)rp * 3 * w * h * 4);
        run_rows<1, 1>(d, dst, w, 2144, rp);
        run_rows<1, 2>(d, dst, w, 2144, rp);
        run_rows<2, 1>(d, dst, w, 2144, rp);
        run_rows<2, 2>(d, dst, w, 2144, rp);
        run_rows<3, 1>(d, dst, w, 2144, rp);
        run_rows<1, 4>(d, dst, w, 2144, rp);
        run_rows<2, 4>(d, dst, w, 2144, rp);
        run_rows<3, 3>(d, dst, w, 2144, rp);
        run_rows_blocked<2, 4, false>(d, dst, w, 2144, rp);
        run_rows_blocked<2, 4, true>(d, dst, w, 2144, rp);
        run_rows_blocked<3, 3, false>(d, dst, w, 2144, rp);
        run_rows_blocked<3, 3, true>(d, dst, w, 2144, rp);
        run_rows_blocked<1, 4, true>(d, dst, w, 2144, rp);
        cudaFree(dst);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
