// Probe: how many DRAM bytes does B200 fetch for small strided reads (Isaac Gym AoS gathers)?
// usage: probe_fetch <l2_fetch_granularity or 0>; run under `ncu --metrics dram__bytes_read.sum`.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

template <int MODE>
__global__ void k_gather(const float* __restrict__ base, float* out, int64_t n, int stride_f, int off_f, int nf) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = base + i * stride_f + off_f;
    float s = 0.f;
    for (int k = 0; k < nf; ++k) {
        float v;
        if (MODE == 0) v = p[k];
        else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p + k));
        else if (MODE == 2) asm volatile("ld.global.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p + k));
        else if (MODE == 3) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p + k));
        else asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p + k));
        s += v;
    }
    out[i] = s;
}

int main(int argc, char** argv) {
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("set granularity %d -> %s, effective %zu\n", gran, cudaGetErrorString(e), got);
    }
    const int64_t n = 1 << 21;
    float *a, *out;
    cudaMalloc(&a, n * 1144 + 4096); cudaMalloc(&out, n * 4);
    cudaMemset(a, 0, n * 1144 + 4096);
    dim3 b(256), g((unsigned)((n + 255) / 256));
    // contact-like: 3 floats at stride 66 floats (264 B), offset 36
    k_gather<0><<<g, b>>>(a, out, n, 66, 36, 3);
    k_gather<1><<<g, b>>>(a, out, n, 66, 36, 3);
    k_gather<2><<<g, b>>>(a, out, n, 66, 36, 3);
    k_gather<3><<<g, b>>>(a, out, n, 66, 36, 3);
    k_gather<4><<<g, b>>>(a, out, n, 66, 36, 3);
    // rigid-like: 10 floats at stride 286 floats (1144 B), offset 16
    k_gather<0><<<g, b>>>(a, out, n, 286, 16, 10);
    k_gather<1><<<g, b>>>(a, out, n, 286, 16, 10);
    k_gather<3><<<g, b>>>(a, out, n, 286, 16, 10);
    // single float at stride 286 (1 sector per env), and at stride 64 floats (256 B)
    k_gather<0><<<g, b>>>(a, out, n, 286, 16, 1);
    k_gather<0><<<g, b>>>(a, out, n, 64, 0, 1);
    k_gather<0><<<g, b>>>(a, out, n, 32, 0, 1);
    k_gather<0><<<g, b>>>(a, out, n, 16, 0, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s (n=%lld)\n", cudaGetErrorString(e), (long long)n);
    return e != cudaSuccess;
}
