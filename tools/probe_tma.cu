// Probe: DRAM bytes fetched by TMA *tensor* box loads of a 16-byte column slice out of 528-byte rows
// (net_contact viewed as [N/2 env pairs][132 floats]) for the L2 promotion modes, and by small 1-D
// cp.async.bulk copies.  Run under ncu --metrics dram__bytes_read.sum.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int ROWS = 64;

__global__ void k_tma_box(const __grid_constant__ CUtensorMap tm, float* out, int col0, int box_cols) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(ROWS * box_cols * 4));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(sm)), "l"(&tm), "r"(col0), "r"((int)(blockIdx.x * ROWS)), "r"(s32(&bar)) : "memory");
    }
    __syncthreads();
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    float s = 0.f;
    for (int i = threadIdx.x; i < ROWS * box_cols; i += blockDim.x) s += sm[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// per-thread small 1-D bulk copies: `bytes` from base + row*stride + off
__global__ void k_bulk_small(const char* base, float* out, int64_t stride, int64_t off, int bytes) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"((int)blockDim.x));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    char* dst = reinterpret_cast<char*>(sm) + threadIdx.x * bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(base + row * stride + off), "r"(bytes), "r"(s32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    out[row] = reinterpret_cast<float*>(dst)[0];
}

int main() {
    const int64_t rows = 1 << 20;            // env pairs
    const int cols = 132;                    // floats per row (528 B)
    float *a, *out;
    cudaMalloc(&a, rows * cols * 4 + 4096); cudaMalloc(&out, rows * 4 * 4);
    cudaMemset(a, 0, rows * cols * 4 + 4096);
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMapL2promotion promos[4] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
    for (int p = 0; p < 4; ++p) {
        for (int bc = 4; bc <= 8; bc += 4) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
            cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
            cuuint32_t box[2] = {(cuuint32_t)bc, ROWS};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, promos[p], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            k_tma_box<<<(unsigned)(rows / ROWS), 64, ROWS * bc * 4>>>(tm, out, 36, bc);
        }
    }
    // small 1-D bulk copies: 16 B at stride 528 off 144; 48 B at stride 1144*? use stride 1152 (16-aligned) off 64
    k_bulk_small<<<(unsigned)(rows / 128), 128, 128 * 16>>>((const char*)a, out, 528, 144, 16);
    k_bulk_small<<<(unsigned)(rows / 128 / 4), 128, 128 * 48>>>((const char*)a, out, 1152, 64, 48);
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
