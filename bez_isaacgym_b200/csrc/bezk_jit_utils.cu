// The generic tensor helpers the north_star names and the reference keeps next to its tasks (bez_isaacgym/utils/torch_jit_utils.py,
// which star-imports isaacgym.torch_utils): quat_rotate / quat_rotate_inverse (projected gravity = quat_rotate_inverse(q, g)),
// scale_transform / unscale_transform / saturate.  KickEnv itself never calls them (its only quat_rotate_inverse calls
// are commented out, DOF values are concatenated unscaled); they ship so that a task written against the reference's helper
// module finds them on the B200 path.  Plain streaming elementwise kernels, arithmetic order as the Python expressions.
#include "bezk_common.cuh"
#include "bezk_internal.h"

namespace bezk {

// isaacgym.torch_utils.quat_rotate / quat_rotate_inverse (xyzw quaternion):
//   a = v * (2 w^2 - 1);  b = cross(q_xyz, v) * w * 2;  c = q_xyz * dot(q_xyz, v) * 2;   rotate: a + b + c, inverse: a - b + c
__global__ void __launch_bounds__(256) quat_rotate_kernel(const float* __restrict__ q, const float* __restrict__ v, float* __restrict__ out,
                                                          int inverse, int64_t n, int q_vec4) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float x, y, z, w;
        if (q_vec4) { const float4 t = ldg_stream4(reinterpret_cast<const float4*>(q) + i); x = t.x; y = t.y; z = t.z; w = t.w; }
        else { x = q[4 * i]; y = q[4 * i + 1]; z = q[4 * i + 2]; w = q[4 * i + 3]; }
        const float vx = ldg_stream(v + 3 * i), vy = ldg_stream(v + 3 * i + 1), vz = ldg_stream(v + 3 * i + 2);
        const float s = 2.0f * (w * w) - 1.0f;
        const float ax = vx * s, ay = vy * s, az = vz * s;
        const float bx = ((y * vz - z * vy) * w) * 2.0f, by = ((z * vx - x * vz) * w) * 2.0f, bz = ((x * vy - y * vx) * w) * 2.0f;
        const float d = (x * vx + y * vy) + z * vz;                         // bmm(q_vec, v)
        const float cx = (x * d) * 2.0f, cy = (y * d) * 2.0f, cz = (z * d) * 2.0f;
        float ox, oy, oz;
        if (inverse) { ox = (ax - bx) + cx; oy = (ay - by) + cy; oz = (az - bz) + cz; }
        else { ox = (ax + bx) + cx; oy = (ay + by) + cy; oz = (az + bz) + cz; }
        __stcs(out + 3 * i, ox); __stcs(out + 3 * i + 1, oy); __stcs(out + 3 * i + 2, oz);
    }
}

// utils/torch_jit_utils.py:78-134.  mode 0: scale_transform = 2 * (x - (lower + upper) * 0.5) / (upper - lower)
//                                   mode 1: unscale_transform = x * (upper - lower) * 0.5 + (lower + upper) * 0.5
//                                   mode 2: saturate = max(min(x, upper), lower)
__global__ void __launch_bounds__(256) scale_transform_kernel(const float* __restrict__ x, const float* __restrict__ lower,
                                                              const float* __restrict__ upper, float* __restrict__ y, int mode,
                                                              int64_t total, int dims) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % dims);
        const float lo = lower[c], hi = upper[c], xv = ldg_stream(x + i);
        float o;
        if (mode == 0) o = (2.0f * (xv - (lo + hi) * 0.5f)) / (hi - lo);
        else if (mode == 1) o = (xv * (hi - lo)) * 0.5f + (lo + hi) * 0.5f;
        else o = tensor_clamp(xv, lo, hi);
        __stcs(y + i, o);
    }
}

static inline unsigned jit_blocks(int64_t items) {
    int64_t b = (items + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148LL * 16) b = 148LL * 16;
    return (unsigned)b;
}

cudaError_t launch_quat_rotate(const float* q, const float* v, float* out, int inverse, int64_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int q_vec4 = (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
    return launch_ex(quat_rotate_kernel, dim3(jit_blocks(n)), dim3(256), 0, st, q, v, out, inverse, n, q_vec4);
}

cudaError_t launch_scale_transform(const float* x, const float* lower, const float* upper, float* y, int mode, int64_t n, int dims,
                                   cudaStream_t st) {
    if (n == 0 || dims == 0) return cudaSuccess;
    return launch_ex(scale_transform_kernel, dim3(jit_blocks(n * dims)), dim3(256), 0, st, x, lower, upper, y, mode, n * dims, dims);
}

}  // namespace bezk
