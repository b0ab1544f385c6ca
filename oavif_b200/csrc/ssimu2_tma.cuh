// ssimu2_tma.cuh — Blackwell tile movement for the recursive-blur kernels: TMA (cp.async.bulk.tensor) descriptors
// built on the host, and the mbarrier / bulk-group PTX the kernels synchronise with.
//
// Why: in the cp.async versions of k_iir_rows / k_iir_cols two of a CTA's four to eight warps did nothing but
// address arithmetic, LDGSTS and STG, and every warp met in one block barrier per chunk — where ncu put 42 % / 38 %
// of the stall samples (profiles/r1_final_ncu_*).  With TMA one elected lane issues a whole tile, out-of-bounds
// elements arrive as zeros (the filter's zero padding, for free), a recursion warp waits on the mbarrier of the
// ring slot it is about to read and on nothing else, and a finished tile leaves through a TMA store that clips
// itself at the image edge.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace oavif {

// ---- device side ------------------------------------------------------------------------------------------
constexpr uint32_t kMbarSuspendNs = 1000000u;   // upper bound of one parked wait, ns
constexpr uint32_t kMbarMaxTries = 1u << 24;    // failed attempts before a wait gives up (__trap)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// makes the initialised barriers visible to the async proxy (TMA completes transactions on them)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Blocks until the barrier's phase with the given parity has completed.  try_wait parks the warp in hardware until
// the phase completes or a time limit passes; with the default limit a waiting warp came back every few hundred
// cycles and re-issued the test (SYNCS + BRA were a quarter of the columns kernel's executed instructions,
// profiles/r2_final_ncu_cols_hot.txt).  The explicit limit is far above any wait seen here, so a waiting warp costs
// no issue slots; completion still wakes it at once.
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
        : "memory");
    return done != 0;
}
// A wait that can never be satisfied (a protocol error, a copy that faulted) must not hang the device: after
// kMbarMaxTries failed attempts — seconds at the very least, every attempt parks first — the kernel traps and the
// host sees a launch failure instead of a stuck GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t tries = 0;
    while (!mbar_try_wait(bar, parity))
        if (++tries > kMbarMaxTries) __trap();
}

// global -> shared tile; completion is counted in bytes on `bar`.  Coordinates are element indices, innermost
// first; anything outside the tensor reads as zero.
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
// global -> shared, a plain run of bytes (16-byte aligned, a multiple of 16 long); completion counted on `bar`
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global tile (bulk async-group completion); elements outside the tensor are not written
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N bulk groups of this thread have finished READING shared memory (the staging may be reused)
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory (st.shared) become visible to the async proxy (a TMA store that follows)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library links cudart only
inline EncodeTiledFn tma_encoder()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// A 4-D f32 tensor {inner, rows, planes, images} with byte strides for dimensions 1..3 and a {box0, box1, 1, 1} box.
inline bool tma_make_4d(CUtensorMap *out, const float *base, uint64_t inner, uint64_t rows, uint64_t planes, uint64_t images,
                        uint64_t row_bytes, uint64_t plane_bytes, uint64_t image_bytes, uint32_t box0, uint32_t box1,
                        bool swizzle128, bool promote256 = false)
{
    EncodeTiledFn enc = tma_encoder();
    if (!enc) return false;
    const cuuint64_t dim[4] = {inner, rows, planes, images};
    // a stride that is never stepped over (dimension of size 1) still has to be a legal value
    const cuuint64_t str[3] = {row_bytes, plane_bytes ? plane_bytes : row_bytes * rows,
                               image_bytes ? image_bytes : (plane_bytes ? plane_bytes : row_bytes * rows) * planes};
    const cuuint32_t box[4] = {box0, box1, 1, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), dim, str, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
               promote256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace oavif
