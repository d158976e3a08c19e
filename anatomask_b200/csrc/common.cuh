// Common device/host helpers for libanatomask_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/anatomask_b200.h"

namespace amb {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 or a negative code; message via amb_last_error()
// ------------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern int g_launch_count;                      // kernels launched by this library (bench.py's gpu_launches)
extern const char* g_last_conv_kernel;          // kernel that served the last amb_conv / amb_conv_wgrad call (bench.py's roofline)

#define AMB_CHECK(cond, code, ...)            \
    do {                                      \
        if (!(cond)) {                        \
            amb::set_error(__VA_ARGS__);      \
            return (code);                    \
        }                                     \
    } while (0)

#define AMB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            amb::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,                \
                           cudaGetErrorString(_e));                                             \
            return AMB_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define AMB_LAUNCH_CHECK()                  \
    do {                                    \
        amb::g_launch_count++;              \
        AMB_CUDA(cudaGetLastError());       \
    } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
int num_sms();

// ------------------------------------------------------------------------------------------------------------
// geometry shared by the HBM-bound kernels: a channels-last activation seen as patches of edge P
// (P = 16 >> stage in the encoder, 2 << level in the decoder) so that "active patches only" and "dense"
// iterate the same way.
// ------------------------------------------------------------------------------------------------------------
struct Geo {
    int N, D, H, W, C;        // tensor (N, D, H, W, C), C fastest
    int P, lgP;               // patch edge in voxels at this resolution (power of two)
    int fd, fh, fw;           // patches per axis (D = fd*P ...)
    const int* list;          // active patch ids n*L + l (L = fd*fh*fw) or nullptr = all patches
    const int* count;         // device count of list entries (nullptr with list == nullptr)
    const uint8_t* active;    // (N, fd, fh, fw) bytes, may be nullptr when dense
};

#ifdef __CUDACC__

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }

struct __align__(16) bf16x8 {
    __nv_bfloat162 a, b, c, d;
};

__device__ __forceinline__ void unpack8(const bf16x8& v, float* f) {
    float2 t;
    t = __bfloat1622float2(v.a); f[0] = t.x; f[1] = t.y;
    t = __bfloat1622float2(v.b); f[2] = t.x; f[3] = t.y;
    t = __bfloat1622float2(v.c); f[4] = t.x; f[5] = t.y;
    t = __bfloat1622float2(v.d); f[6] = t.x; f[7] = t.y;
}

__device__ __forceinline__ bf16x8 pack8(const float* f) {
    bf16x8 v;
    v.a = __floats2bfloat162_rn(f[0], f[1]);
    v.b = __floats2bfloat162_rn(f[2], f[3]);
    v.c = __floats2bfloat162_rn(f[4], f[5]);
    v.d = __floats2bfloat162_rn(f[6], f[7]);
    return v;
}

// explicit 128-bit accesses: the struct-typed form above was split by nvcc into four 32-bit LDGs, which left every
// elementwise kernel at a fraction of HBM speed
__device__ __forceinline__ void unpack_u4(const uint4& v, float* f) {
    float2 t;
    t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x)); f[0] = t.x; f[1] = t.y;
    t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y)); f[2] = t.x; f[3] = t.y;
    t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.z)); f[4] = t.x; f[5] = t.y;
    t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.w)); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void load8(const bf16* p, float* f) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    unpack_u4(v, f);
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void store8(bf16* p, const float* f) {
    uint4 v;
    v.x = pack2(f[0], f[1]); v.y = pack2(f[2], f[3]); v.z = pack2(f[4], f[5]); v.w = pack2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = v;
}

// fused output transform of the conv epilogues: act(acc * scale + shift)
__device__ __forceinline__ float ep_apply(float acc, float scale, float shift, int act) {
    float f = fmaf(acc, scale, shift);
    if (act == AMB_ACT_RELU6) f = fminf(fmaxf(f, 0.f), 6.f);
    else if (act == AMB_ACT_LRELU) f = f > 0.f ? f : 0.01f * f;
    return f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// voxel-run iterator used by all HBM-bound kernels: item = (run, channel-group of 8)
// run = P consecutive voxels along W inside one patch row.
struct RunPos {
    long voxel;       // linear voxel index of the first voxel of the run ((n*D+z)*H+y)*W+x
    int n, patch;     // sample and linear patch id (n*L + l)
};

__device__ __forceinline__ RunPos decode_run(const Geo& g, long run) {
    // run = (entry * P + rz) * P + ry
    int ry = (int)(run & (g.P - 1));
    long t = run >> g.lgP;
    int rz = (int)(t & (g.P - 1));
    long entry = t >> g.lgP;
    int pid = g.list ? g.list[entry] : (int)entry;
    int L = g.fd * g.fh * g.fw;
    int n = pid / L, l = pid - n * L;
    int pz = l / (g.fh * g.fw);
    int r2 = l - pz * g.fh * g.fw;
    int py = r2 / g.fw, px = r2 - py * g.fw;
    RunPos r;
    r.n = n;
    r.patch = pid;
    r.voxel = (((long)n * g.D + (pz << g.lgP) + rz) * g.H + (py << g.lgP) + ry) * g.W + ((long)px << g.lgP);
    return r;
}

__device__ __forceinline__ long geo_num_runs(const Geo& g) {
    long entries = g.list ? (long)(*g.count) : (long)g.N * g.fd * g.fh * g.fw;
    return entries << (2 * g.lgP);
}

__device__ __forceinline__ bool voxel_active(const Geo& g, long voxel) {
    // voxel linear → patch flag (32-bit arithmetic: voxel counts on this path are far below 2^31)
    uint32_t t = (uint32_t)voxel;
    const uint32_t x = t % (uint32_t)g.W; t /= (uint32_t)g.W;
    const uint32_t y = t % (uint32_t)g.H; t /= (uint32_t)g.H;
    const uint32_t z = t % (uint32_t)g.D;
    const uint32_t n = t / (uint32_t)g.D;
    return g.active[((n * g.fd + (z >> g.lgP)) * g.fh + (y >> g.lgP)) * g.fw + (x >> g.lgP)] != 0;
}

#endif  // __CUDACC__

}  // namespace amb
