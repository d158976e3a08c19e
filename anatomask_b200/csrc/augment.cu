// Device-side input pipeline: the per-sample work of nnUNetDataLoader3D.generate_train_batch
// (N/training/dataloading/data_loader_3d.py:7-51: crop the initial patch at a bounding box, zero-pad what lies outside the
// case) followed by the training transforms of P/pretrain_AntoMask.py:79-113 as the pre-training scripts configure them:
// batchgenerators SpatialTransform (rotation / isotropic scaling about the patch centre, order-3 spline resampling,
// constant border 0, no elastic deformation, centre crop otherwise) and MirrorTransform — producing `inp` (B,1,D,H,W) fp32
// directly in HBM.  The random draws stay on the host (anatomask_b200/augment.py); these kernels do the voxel work.
//
// Order-3 resampling is scipy.ndimage.map_coordinates(order=3, mode='constant', cval=0) — what batchgenerators'
// interpolate_img calls — restated:
//   1. prefilter: the cubic B-spline coefficients of the initial patch, mirror boundary at the PATCH edges.  scipy runs the
//      recursive filter (gain 6, pole z = √3 − 2, exact mirror initialisation); its impulse response on the mirror-extended
//      signal is h[k] = −6z/(1 − z²) · z^|k|, so each axis is a symmetric FIR here, truncated at |k| <= 20 (z²¹ ≈ 1e-12):
//      embarrassingly parallel, one pass per axis, the first pass fused with the crop + zero padding.
//   2. evaluation: for an output voxel the source coordinate is c = M·(o − (O−1)/2) + (P/2 − ½); outside [0, P−1] on any
//      axis → cval; else the 4×4×4 B-spline support around ⌊c⌋ with out-of-range support indices mirrored.
// The mirror flips are folded into the write index.
#include "common.cuh"

namespace amb {

#define AUG_K 20                          // FIR half-width
#define AUG_R 4                           // outputs per thread along the filtered axis

struct AugFir {
    float h[AUG_K + 1];
};

__device__ __forceinline__ int mirror_idx(int j, int n) {
    // scipy's mirror extension (period 2n − 2): … 2 1 | 0 1 2 … n−1 | n−2 n−3 …
    if (n == 1) return 0;
    const int s2 = 2 * n - 2;
    j = j < 0 ? -j : j;
    j %= s2;
    return j >= n ? s2 - j : j;
}

// pass along x (contiguous), fused with the crop: the patch voxel (z, y, x) is src[lb + (z, y, x)] or 0 outside the case
__global__ void prefilter_x_kernel(const float* __restrict__ src, int sD, int sH, int sW, int lz, int ly, int lx,
                                   float* __restrict__ dst, int pD, int pH, int pW, AugFir F) {
    const int groups = (pW + AUG_R - 1) / AUG_R;
    const long total = (long)pD * pH * groups;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int gx = (int)(i % groups);
        const long row = i / groups;
        const int y = (int)(row % pH), z = (int)(row / pH);
        const int sz = z + lz, sy = y + ly;
        const bool row_in = sz >= 0 && sz < sD && sy >= 0 && sy < sH;
        const float* srow = src + ((long)sz * sH + sy) * sW;
        const int x0 = gx * AUG_R;
        float win[2 * AUG_K + AUG_R];
#pragma unroll
        for (int k = 0; k < 2 * AUG_K + AUG_R; ++k) {
            const int px = mirror_idx(x0 - AUG_K + k, pW), sx = px + lx;
            win[k] = (row_in && sx >= 0 && sx < sW) ? __ldg(srow + sx) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < AUG_R; ++r) {
            if (x0 + r >= pW) break;
            float acc = F.h[0] * win[AUG_K + r];
#pragma unroll
            for (int k = 1; k <= AUG_K; ++k) acc = fmaf(F.h[k], win[AUG_K + r - k] + win[AUG_K + r + k], acc);
            dst[row * pW + x0 + r] = acc;
        }
    }
}

// pass along y (AXIS 1) or z (AXIS 0) of a (pD, pH, pW) buffer: x is the thread-fastest index, so every tap is a coalesced load
template <int AXIS>
__global__ void prefilter_yz_kernel(const float* __restrict__ in, float* __restrict__ out, int pD, int pH, int pW, AugFir F) {
    const int n = AXIS == 1 ? pH : pD, other = AXIS == 1 ? pD : pH;
    const int groups = (n + AUG_R - 1) / AUG_R;
    const long stride = AXIS == 1 ? pW : (long)pH * pW;
    const long total = (long)other * groups * pW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % pW);
        long t = i / pW;
        const int g = (int)(t % groups), o = (int)(t / groups);
        const long base = AXIS == 1 ? (long)o * pH * pW + x : (long)o * pW + x;
        const int a0 = g * AUG_R;
        float win[2 * AUG_K + AUG_R];
#pragma unroll
        for (int k = 0; k < 2 * AUG_K + AUG_R; ++k) win[k] = __ldg(in + base + (long)mirror_idx(a0 - AUG_K + k, n) * stride);
#pragma unroll
        for (int r = 0; r < AUG_R; ++r) {
            if (a0 + r >= n) break;
            float acc = F.h[0] * win[AUG_K + r];
#pragma unroll
            for (int k = 1; k <= AUG_K; ++k) acc = fmaf(F.h[k], win[AUG_K + r - k] + win[AUG_K + r + k], acc);
            out[base + (long)(a0 + r) * stride] = acc;
        }
    }
}

struct AugXf {
    double m[9];                 // c_j = Σ_i (o_i − oc_i)·m[i*3 + j] + pc_j   (batchgenerators' row-vector convention, scale folded in)
    double oc[3], pc[3];
    int mirror[3];
};

__device__ __forceinline__ void bspline3(float t, float* w) {
    const float u = 1.f - t;
    w[0] = u * u * u * (1.f / 6.f);
    w[1] = (3.f * t * t * t - 6.f * t * t + 4.f) * (1.f / 6.f);
    w[2] = (-3.f * t * t * t + 3.f * t * t + 3.f * t + 1.f) * (1.f / 6.f);
    w[3] = t * t * t * (1.f / 6.f);
}

__global__ void resample3_kernel(const float* __restrict__ coef, int pD, int pH, int pW, float* __restrict__ out, int oD,
                                 int oH, int oW, AugXf X, float cval) {
    const long total = (long)oD * oH * oW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % oW), y = (int)((i / oW) % oH), z = (int)(i / ((long)oW * oH));
        const double d0 = z - X.oc[0], d1 = y - X.oc[1], d2 = x - X.oc[2];
        const double c0 = d0 * X.m[0] + d1 * X.m[3] + d2 * X.m[6] + X.pc[0];
        const double c1 = d0 * X.m[1] + d1 * X.m[4] + d2 * X.m[7] + X.pc[1];
        const double c2 = d0 * X.m[2] + d1 * X.m[5] + d2 * X.m[8] + X.pc[2];
        float v = cval;
        if (c0 >= 0.0 && c0 <= pD - 1 && c1 >= 0.0 && c1 <= pH - 1 && c2 >= 0.0 && c2 <= pW - 1) {
            const int f0 = (int)floor(c0), f1 = (int)floor(c1), f2 = (int)floor(c2);
            float w0[4], w1[4], w2[4];
            bspline3((float)(c0 - f0), w0);
            bspline3((float)(c1 - f1), w1);
            bspline3((float)(c2 - f2), w2);
            int i0[4], i1[4], i2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                i0[k] = mirror_idx(f0 - 1 + k, pD);
                i1[k] = mirror_idx(f1 - 1 + k, pH);
                i2[k] = mirror_idx(f2 - 1 + k, pW);
            }
            float acc = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float sa = 0.f;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float* row = coef + ((long)i0[a] * pH + i1[b]) * pW;
                    const float sb = w2[0] * __ldg(row + i2[0]) + w2[1] * __ldg(row + i2[1]) + w2[2] * __ldg(row + i2[2]) +
                                     w2[3] * __ldg(row + i2[3]);
                    sa = fmaf(w1[b], sb, sa);
                }
                acc = fmaf(w0[a], sa, acc);
            }
            v = acc;
        }
        const int mz = X.mirror[0] ? oD - 1 - z : z, my = X.mirror[1] ? oH - 1 - y : y, mx = X.mirror[2] ? oW - 1 - x : x;
        out[((long)mz * oH + my) * oW + mx] = v;
    }
}

// no rotation / scaling drawn: centre crop of the initial patch (batchgenerators center_crop_aug) straight from the case,
// zero padding outside it, mirror flips in the write index
__global__ void crop_mirror_kernel(const float* __restrict__ src, int sD, int sH, int sW, int lz, int ly, int lx,
                                   float* __restrict__ out, int oD, int oH, int oW, int fz, int fy, int fx) {
    const long total = (long)oD * oH * oW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % oW), y = (int)((i / oW) % oH), z = (int)(i / ((long)oW * oH));
        const int sz = z + lz, sy = y + ly, sx = x + lx;
        float v = 0.f;
        if (sz >= 0 && sz < sD && sy >= 0 && sy < sH && sx >= 0 && sx < sW) v = __ldg(src + ((long)sz * sH + sy) * sW + sx);
        const int mz = fz ? oD - 1 - z : z, my = fy ? oH - 1 - y : y, mx = fx ? oW - 1 - x : x;
        out[((long)mz * oH + my) * oW + mx] = v;
    }
}

static inline int aug_grid(long items, int block) {
    long b = (items + block - 1) / block;
    long cap = (long)num_sms() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace amb

using namespace amb;

extern "C" int amb_aug_spline_prefilter(const float* src, int sD, int sH, int sW, int lb_z, int lb_y, int lb_x, float* coef,
                                        float* scratch, int pD, int pH, int pW, void* stream) {
    AMB_CHECK(src && coef && scratch && pD > 1 && pH > 1 && pW > 1, AMB_ERR_ARG, "spline prefilter: bad arguments");
    AugFir F;
    const double z = sqrt(3.0) - 2.0, g = -6.0 * z / (1.0 - z * z);
    double zk = 1.0;
    for (int k = 0; k <= AUG_K; ++k) { F.h[k] = (float)(g * zk); zk *= z; }
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 256;
    prefilter_x_kernel<<<aug_grid((long)pD * pH * ((pW + AUG_R - 1) / AUG_R), block), block, 0, st>>>(
        src, sD, sH, sW, lb_z, lb_y, lb_x, coef, pD, pH, pW, F);
    AMB_LAUNCH_CHECK();
    prefilter_yz_kernel<1><<<aug_grid((long)pD * ((pH + AUG_R - 1) / AUG_R) * pW, block), block, 0, st>>>(coef, scratch, pD, pH, pW, F);
    AMB_LAUNCH_CHECK();
    prefilter_yz_kernel<0><<<aug_grid((long)pH * ((pD + AUG_R - 1) / AUG_R) * pW, block), block, 0, st>>>(scratch, coef, pD, pH, pW, F);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_aug_resample(const float* coef, int pD, int pH, int pW, const double* matrix9, const int* mirror3, float cval,
                                float* out, int oD, int oH, int oW, void* stream) {
    AMB_CHECK(coef && out && matrix9, AMB_ERR_ARG, "resample: bad arguments");
    AugXf X;
    for (int i = 0; i < 9; ++i) X.m[i] = matrix9[i];
    const int o[3] = {oD, oH, oW}, p[3] = {pD, pH, pW};
    for (int d = 0; d < 3; ++d) {
        X.oc[d] = (o[d] - 1) / 2.0;          // create_zero_centered_coordinate_mesh
        X.pc[d] = p[d] / 2.0 - 0.5;          // ctr = data.shape[d + 2] / 2. - 0.5
        X.mirror[d] = mirror3 ? mirror3[d] : 0;
    }
    const long total = (long)oD * oH * oW;
    resample3_kernel<<<aug_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(coef, pD, pH, pW, out, oD, oH, oW, X, cval);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_aug_crop_mirror(const float* src, int sD, int sH, int sW, int lb_z, int lb_y, int lb_x, const int* mirror3,
                                   float* out, int oD, int oH, int oW, void* stream) {
    AMB_CHECK(src && out, AMB_ERR_ARG, "crop: bad arguments");
    const long total = (long)oD * oH * oW;
    crop_mirror_kernel<<<aug_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
        src, sD, sH, sW, lb_z, lb_y, lb_x, out, oD, oH, oW, mirror3 ? mirror3[0] : 0, mirror3 ? mirror3[1] : 0, mirror3 ? mirror3[2] : 0);
    AMB_LAUNCH_CHECK();
    return 0;
}
