// HBM-bound kernels: pooled masked norm / BatchNorm (stats, finalize, apply, backward), densify fill, skip add,
// layout conversion and the active-patch work-list.  All activations bf16 channels-last, 128-bit accesses,
// one item = 8 channels of one voxel.
#include "common.cuh"

namespace amb {

int make_geo(const amb_geo* a, Geo& g) {
    AMB_CHECK(a != nullptr, AMB_ERR_ARG, "geo is null");
    AMB_CHECK(a->C % 8 == 0 && a->C >= 8, AMB_ERR_ARG, "C=%d must be a multiple of 8", a->C);
    AMB_CHECK(a->fd > 0 && a->D % a->fd == 0 && a->H % a->fh == 0 && a->W % a->fw == 0, AMB_ERR_ARG,
              "dims (%d,%d,%d) not divisible by mask grid (%d,%d,%d)", a->D, a->H, a->W, a->fd, a->fh, a->fw);
    int P = a->D / a->fd;
    AMB_CHECK(a->H / a->fh == P && a->W / a->fw == P && (P & (P - 1)) == 0, AMB_ERR_ARG,
              "patch edge must be a power of two and equal on all axes (got %d,%d,%d)", P, a->H / a->fh, a->W / a->fw);
    g.N = a->N; g.D = a->D; g.H = a->H; g.W = a->W; g.C = a->C;
    g.P = P; g.lgP = 0;
    while ((1 << g.lgP) < P) g.lgP++;
    g.fd = a->fd; g.fh = a->fh; g.fw = a->fw;
    g.list = a->active_list; g.count = a->active_count; g.active = a->active;
    AMB_CHECK((g.list == nullptr) == (g.count == nullptr), AMB_ERR_ARG, "active_list and active_count go together");
    return 0;
}

static int pick_block(int CG) {
    // block size: a multiple of the channel-group count so a thread keeps its channel group across the grid stride
    AMB_CHECK(CG <= 512, AMB_ERR_ARG, "C=%d too large for the elementwise kernels (max 4096)", CG * 8);
    int b = (256 / CG) * CG;
    if (b == 0) b = CG;
    return b;
}

// item → (voxel, channel-group); returns false when out of range
struct Item {
    long voxel;
    int cg;
    int patch;
};

// Iteration: a thread keeps its channel group (blockDim is a multiple of CG) and walks voxel slots with a constant
// stride, so the per-item work is an add (dense) or shifts + 32-bit divisions by the mask-grid extents (active list) —
// the first version paid two 64-bit divisions per 16-byte item and ran at a quarter of HBM speed.
struct Walk {
    long slot, step, nslots;      // voxel slots: dense = voxels; list = (entries << 3·lgP)
};
__device__ __forceinline__ Walk make_walk(const Geo& g, int CG) {
    Walk w;
    w.slot = ((long)blockIdx.x * blockDim.x + threadIdx.x) / CG;
    w.step = ((long)gridDim.x * blockDim.x) / CG;
    w.nslots = g.list ? ((long)(*g.count) << (3 * g.lgP)) : (long)g.N * g.D * g.H * g.W;
    return w;
}
template <bool LIST>
__device__ __forceinline__ long slot_voxel(const Geo& g, long slot) {
    if (!LIST) return slot;
    const uint32_t P1 = (uint32_t)g.P - 1u;
    const uint32_t s = (uint32_t)slot;                    // < 2^31 for every tensor on this path
    const uint32_t v = s & P1, ry = (s >> g.lgP) & P1, rz = (s >> (2 * g.lgP)) & P1;
    const uint32_t pid = (uint32_t)g.list[s >> (3 * g.lgP)];
    const uint32_t L = (uint32_t)(g.fd * g.fh * g.fw), hw = (uint32_t)(g.fh * g.fw);
    const uint32_t n = pid / L, l = pid - n * L;
    const uint32_t pz = l / hw, r2 = l - pz * hw;
    const uint32_t py = r2 / (uint32_t)g.fw, px = r2 - py * (uint32_t)g.fw;
    return (((long)n * g.D + (pz << g.lgP) + rz) * g.H + (py << g.lgP) + ry) * g.W + ((long)px << g.lgP) + v;
}

__device__ __forceinline__ bool get_item(const Geo& g, long item, long total, int CG, Item& it) {
    if (item >= total) return false;
    it.cg = (int)(item % CG);
    long t = item / CG;
    if (g.list == nullptr) {       // dense: flat
        it.voxel = t;
        it.patch = -1;
        return true;
    }
    int v = (int)(t & (g.P - 1));
    RunPos r = decode_run(g, t >> g.lgP);
    it.voxel = r.voxel + v;
    it.patch = r.patch;
    return true;
}

__device__ __forceinline__ long total_items(const Geo& g, int CG) {
    if (g.list == nullptr) return (long)g.N * g.D * g.H * g.W * CG;
    return (geo_num_runs(g) << g.lgP) * CG;
}

__device__ __forceinline__ float act_fwd(float u, int act) {
    if (act == AMB_ACT_LRELU) return u > 0.f ? u : 0.01f * u;
    if (act == AMB_ACT_RELU6) return fminf(fmaxf(u, 0.f), 6.f);
    return u;
}
__device__ __forceinline__ float act_grad(float u, int act) {
    if (act == AMB_ACT_LRELU) return u > 0.f ? 1.f : 0.01f;
    if (act == AMB_ACT_RELU6) return (u > 0.f && u < 6.f) ? 1.f : 0.f;
    return 1.f;
}

// ------------------------------------------------------------------------------------------------------------
// Σx, Σx² (mode 0)   |   Σg, Σg·x̂ (+ Σ_inactive dout → dtoken) (mode 1)
// ------------------------------------------------------------------------------------------------------------
// LIST: walk the active-patch work-list (else the dense tensor); FILL: densify backward (visits every voxel, routes
// masked voxels to the mask-token gradient).  Compile-time so the dense BatchNorm instantiation carries no decode code.
template <int MODE, int ACT, bool LIST, bool FILL>
__global__ void __launch_bounds__(512, 2) reduce_kernel(Geo g, const bf16* __restrict__ x, const bf16* __restrict__ dout,
                                                     const bf16* __restrict__ res, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, const float* __restrict__ saved,
                                                     int act_unused, int fill_unused, double* __restrict__ sums,
                                                     double* __restrict__ dtoken) {
    constexpr int act = ACT;
    constexpr bool fill = FILL;
    extern __shared__ float sacc[];           // [3][C]: Σ, Σ·, Σ_inactive — per-CTA fp32 partials, fp64 across CTAs
    const int CG = g.C / 8;
    for (int i = threadIdx.x; i < g.C * 3; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    float a0[8], a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = a2[j] = 0.f;
    Geo gd = g;
    if (MODE == 1 && fill) gd.list = nullptr;            // densify backward visits every voxel
    const int cg = threadIdx.x % CG;
    float sc[8], sh[8];
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = scale[cg * 8 + j]; sh[j] = shift[cg * 8 + j]; }
    }
    Walk w = make_walk(gd, CG);
#pragma unroll 2
    for (; w.slot < w.nslots; w.slot += w.step) {
        Item it;
        it.voxel = slot_voxel<LIST && !FILL>(gd, w.slot);
        it.cg = cg;
        const long off = it.voxel * g.C + it.cg * 8;
        if (MODE == 0) {
            float f[8];
            load8(x + off, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { a0[j] += f[j]; a1[j] += f[j] * f[j]; }
        } else {
            float d[8];
            load8(dout + off, d);
            if (fill && !voxel_active(g, it.voxel)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a2[j] += d[j];
                continue;
            }
            float f[8], r[8];
            load8(x + off, f);
            if (res) load8(res + off, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float u = sc[j] * f[j] + sh[j] + (res ? r[j] : 0.f);
                float gj = d[j] * act_grad(u, act);
                a0[j] += gj;
                a1[j] += gj * f[j];            // Σg·x; turned into Σg·x̂ = rstd·(Σg·x − μ·Σg) after the loop
            }
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            a1[j] = saved[g.C + cg * 8 + j] * (a1[j] - saved[cg * 8 + j] * a0[j]);
    }
    // lanes of a warp that share a channel group (lane % CG equal) are combined with shuffles first
    const bool shfl = CG < 32 && (CG & (CG - 1)) == 0 && (blockDim.x & 31) == 0;
    if (shfl) {
        for (int o = CG; o < 32; o <<= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], o);
                a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], o);
                if (MODE == 1) a2[j] += __shfl_xor_sync(0xffffffffu, a2[j], o);
            }
        }
    }
    if (!shfl || (threadIdx.x & 31) < CG) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sacc[cg * 8 + j], a0[j]);
            atomicAdd(&sacc[g.C + cg * 8 + j], a1[j]);
            if (MODE == 1 && fill) atomicAdd(&sacc[2 * g.C + cg * 8 + j], a2[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g.C; i += blockDim.x) {
        atomicAdd(&sums[i], (double)sacc[i]);
        atomicAdd(&sums[g.C + i], (double)sacc[g.C + i]);
        if (MODE == 1 && fill && dtoken) atomicAdd(&dtoken[i], (double)sacc[2 * g.C + i]);
    }
}

__global__ void finalize_kernel(Geo g, const double* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, float* __restrict__ scale,
                                float* __restrict__ shift, float* __restrict__ saved, float* running_mean,
                                float* running_var, long* nbt, float momentum, const double* n_total) {
    const double n = n_total ? *n_total
                             : (g.list ? (double)((long)(*g.count) << (3 * g.lgP)) : (double)g.N * g.D * g.H * g.W);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.C; c += gridDim.x * blockDim.x) {
        double mean = sums[c] / n;
        double var = sums[g.C + c] / n - mean * mean;
        if (var < 0.0) var = 0.0;
        float rstd = (float)(1.0 / sqrt(var + (double)eps));
        float s = gamma[c] * rstd;
        scale[c] = s;
        shift[c] = beta[c] - (float)mean * s;
        saved[c] = (float)mean;
        saved[g.C + c] = rstd;
        if (running_mean) {
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * n / (n - 1.0));
        }
    }
    if (nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
}

__global__ void eval_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                            float* scale, float* shift, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        float s = gamma[c] * rsqrtf(rv[c] + eps);
        scale[c] = s;
        shift[c] = beta[c] - rm[c] * s;
    }
}

template <int ACT, bool LIST, bool FILL>
__global__ void __launch_bounds__(512) apply_kernel(Geo g, const bf16* __restrict__ x, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, const bf16* __restrict__ res,
                                                    const float* __restrict__ token, int act_unused, bf16* __restrict__ out) {
    constexpr int act = ACT;
    const int CG = g.C / 8;
    Geo gd = g;
    if (FILL) gd.list = nullptr;
    const int cg = threadIdx.x % CG;
    float sc[8], sh[8], tk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = scale[cg * 8 + j]; sh[j] = shift[cg * 8 + j];
        tk[j] = token ? token[cg * 8 + j] : 0.f;
    }
    Walk w = make_walk(gd, CG);
#pragma unroll 2
    for (; w.slot < w.nslots; w.slot += w.step) {
        Item it;
        it.voxel = slot_voxel<LIST && !FILL>(gd, w.slot);
        it.cg = cg;
        const long off = it.voxel * g.C + it.cg * 8;
        float o[8];
        if (FILL && !voxel_active(g, it.voxel)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = tk[j];
        } else {
            float f[8], r[8];
            load8(x + off, f);
            if (res) load8(res + off, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = act_fwd(sc[j] * f[j] + sh[j] + (res ? r[j] : 0.f), act);
        }
        store8(out + off, o);
    }
}

template <int ACT, bool LIST>
__global__ void __launch_bounds__(512, 2)
bwd_apply_kernel(Geo g, const bf16* __restrict__ dout, const bf16* __restrict__ x, const bf16* __restrict__ res,
                 const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ saved,
                 const double* __restrict__ sums, int act_unused, int fill, bf16* __restrict__ dx, bf16* __restrict__ dres,
                 float* __restrict__ dgamma, float* __restrict__ dbeta, const double* n_total) {
    constexpr int act = ACT;
    const int CG = g.C / 8;
    const int cg = threadIdx.x % CG;
    const double n = n_total ? *n_total
                             : (g.list ? (double)((long)(*g.count) << (3 * g.lgP)) : (double)g.N * g.D * g.H * g.W);
    // dx = scale·(g − Σg/n − x̂·Σ(g·x̂)/n) = scale·g + ca·x + cb with per-channel constants (fewer live registers)
    float sc[8], sh[8], ca[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int c = cg * 8 + j;
        sc[j] = scale[c]; sh[j] = shift[c];
        const float mu = saved[c], rs = saved[g.C + c];
        const float m1 = (float)(sums[c] / n), m2 = (float)(sums[g.C + c] / n);
        ca[j] = -sc[j] * m2 * rs;
        cb[j] = -sc[j] * m1 - ca[j] * mu;
    }
    if (blockIdx.x == 0 && dgamma) {
        for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
            dbeta[c] = (float)sums[c];
            dgamma[c] = (float)sums[g.C + c];
        }
    }
    Walk w = make_walk(g, CG);
#pragma unroll 2
    for (; w.slot < w.nslots; w.slot += w.step) {
        Item it;
        it.voxel = slot_voxel<LIST>(g, w.slot);
        it.cg = cg;
        const long off = it.voxel * g.C + it.cg * 8;
        float d[8], f[8], r[8], o[8], gg[8];
        load8(dout + off, d);
        load8(x + off, f);
        if (res) load8(res + off, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float u = sc[j] * f[j] + sh[j] + (res ? r[j] : 0.f);
            gg[j] = d[j] * act_grad(u, act);
            o[j] = fmaf(sc[j], gg[j], fmaf(ca[j], f[j], cb[j]));
        }
        store8(dx + off, o);
        if (dres) store8(dres + off, gg);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Dense walk with batched loads (the decoder's BatchNorm passes, 0.27 - 1.6 GB each): a thread owns chunk ids
// tid + k·T (T = threads in the grid, a multiple of the channel-group count, so its channel group is fixed) and issues the
// 128-bit loads of DB consecutive k before consuming any of them.
// ------------------------------------------------------------------------------------------------------------
#define DB 4
template <int MODE, int ACT>
__global__ void __launch_bounds__(512, 1) reduce_dense_kernel(Geo g, const bf16* __restrict__ x, const bf16* __restrict__ dout,
                                                           const bf16* __restrict__ res, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, const float* __restrict__ saved,
                                                           double* __restrict__ sums) {
    extern __shared__ float sacc[];           // [2][C]
    const int CG = g.C / 8;
    for (int i = threadIdx.x; i < g.C * 2; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
    const int cg = threadIdx.x % CG;
    float sc[8], sh[8];
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = scale[cg * 8 + j]; sh[j] = shift[cg * 8 + j]; }
    }
    const long T = (long)gridDim.x * blockDim.x;
    const long total = (long)g.N * g.D * g.H * g.W * CG;
    for (long c0 = (long)blockIdx.x * blockDim.x + threadIdx.x; c0 < total; c0 += DB * T) {
        uint4 vx[DB], vd[DB], vr[DB];
#pragma unroll
        for (int k = 0; k < DB; ++k) {
            const long c = c0 + k * T;
            if (c < total) {
                vx[k] = __ldg(reinterpret_cast<const uint4*>(x) + c);
                if (MODE == 1) {
                    vd[k] = __ldg(reinterpret_cast<const uint4*>(dout) + c);
                    if (res) vr[k] = __ldg(reinterpret_cast<const uint4*>(res) + c);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < DB; ++k) {
            if (c0 + k * T >= total) break;
            float f[8];
            unpack_u4(vx[k], f);
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { a0[j] += f[j]; a1[j] += f[j] * f[j]; }
            } else {
                float d[8], rr[8];
                unpack_u4(vd[k], d);
                if (res) unpack_u4(vr[k], rr);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float uu = sc[j] * f[j] + sh[j] + (res ? rr[j] : 0.f);
                    float gj = d[j] * act_grad(uu, ACT);
                    a0[j] += gj;
                    a1[j] += gj * f[j];
                }
            }
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            a1[j] = saved[g.C + cg * 8 + j] * (a1[j] - saved[cg * 8 + j] * a0[j]);
    }
    const bool shfl = CG < 32 && (CG & (CG - 1)) == 0 && (blockDim.x & 31) == 0;
    if (shfl) {
        for (int o = CG; o < 32; o <<= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], o);
                a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], o);
            }
        }
    }
    if (!shfl || (threadIdx.x & 31) < CG) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sacc[cg * 8 + j], a0[j]);
            atomicAdd(&sacc[g.C + cg * 8 + j], a1[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g.C; i += blockDim.x) {
        atomicAdd(&sums[i], (double)sacc[i]);
        atomicAdd(&sums[g.C + i], (double)sacc[g.C + i]);
    }
}

template <int ACT>
__global__ void __launch_bounds__(512, 1) apply_dense_kernel(Geo g, const bf16* __restrict__ x, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, const bf16* __restrict__ res,
                                                          bf16* __restrict__ out) {
    const int CG = g.C / 8;
    const int cg = threadIdx.x % CG;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale[cg * 8 + j]; sh[j] = shift[cg * 8 + j]; }
    const long T = (long)gridDim.x * blockDim.x;
    const long total = (long)g.N * g.D * g.H * g.W * CG;
    for (long c0 = (long)blockIdx.x * blockDim.x + threadIdx.x; c0 < total; c0 += DB * T) {
        uint4 vx[DB], vr[DB];
#pragma unroll
        for (int k = 0; k < DB; ++k) {
            const long c = c0 + k * T;
            if (c < total) {
                vx[k] = __ldg(reinterpret_cast<const uint4*>(x) + c);
                if (res) vr[k] = __ldg(reinterpret_cast<const uint4*>(res) + c);
            }
        }
#pragma unroll
        for (int k = 0; k < DB; ++k) {
            const long c = c0 + k * T;
            if (c >= total) break;
            float f[8], rr[8], o[8];
            unpack_u4(vx[k], f);
            if (res) unpack_u4(vr[k], rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = act_fwd(sc[j] * f[j] + sh[j] + (res ? rr[j] : 0.f), ACT);
            store8(out + c * 8, o);
        }
    }
}

template <int ACT>
__global__ void __launch_bounds__(512, 1)
bwd_apply_dense_kernel(Geo g, const bf16* __restrict__ dout, const bf16* __restrict__ x, const bf16* __restrict__ res,
                       const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ saved,
                       const double* __restrict__ sums, bf16* __restrict__ dx, bf16* __restrict__ dres,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, const double* n_total) {
    const int CG = g.C / 8;
    const int cg = threadIdx.x % CG;
    const double n = n_total ? *n_total : (double)g.N * g.D * g.H * g.W;
    float sc[8], sh[8], ca[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int c = cg * 8 + j;
        sc[j] = scale[c]; sh[j] = shift[c];
        const float mu = saved[c], rs = saved[g.C + c];
        const float m1 = (float)(sums[c] / n), m2 = (float)(sums[g.C + c] / n);
        ca[j] = -sc[j] * m2 * rs;
        cb[j] = -sc[j] * m1 - ca[j] * mu;
    }
    if (blockIdx.x == 0 && dgamma) {
        for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
            dbeta[c] = (float)sums[c];
            dgamma[c] = (float)sums[g.C + c];
        }
    }
    const long T = (long)gridDim.x * blockDim.x;
    const long total = (long)g.N * g.D * g.H * g.W * CG;
    for (long c0 = (long)blockIdx.x * blockDim.x + threadIdx.x; c0 < total; c0 += DB * T) {
        uint4 vx[DB], vd[DB], vr[DB];
#pragma unroll
        for (int k = 0; k < DB; ++k) {
            const long c = c0 + k * T;
            if (c < total) {
                vd[k] = __ldg(reinterpret_cast<const uint4*>(dout) + c);
                vx[k] = __ldg(reinterpret_cast<const uint4*>(x) + c);
                if (res) vr[k] = __ldg(reinterpret_cast<const uint4*>(res) + c);
            }
        }
#pragma unroll
        for (int k = 0; k < DB; ++k) {
            const long c = c0 + k * T;
            if (c >= total) break;
            float d[8], f[8], rr[8], o[8], gg[8];
            unpack_u4(vd[k], d);
            unpack_u4(vx[k], f);
            if (res) unpack_u4(vr[k], rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float uu = sc[j] * f[j] + sh[j] + (res ? rr[j] : 0.f);
                gg[j] = d[j] * act_grad(uu, ACT);
                o[j] = fmaf(sc[j], gg[j], fmaf(ca[j], f[j], cb[j]));
            }
            store8(dx + c * 8, o);
            if (dres) store8(dres + c * 8, gg);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Row-group walk of the active-patch list (the sparse encoder's norms).  The per-chunk walk above decodes a patch id
// (three 32-bit divisions) for every 16-byte item; here one warp iteration covers G = min(4, P) consecutive y-rows of one
// patch at one z — a lane decodes once and then touches the same 16-byte column of G rows (G independent 128-bit loads in
// flight per tensor).  A patch row is P·C/8 chunks; 32 of them per warp iteration (UPR iterations per row when it is longer).
// Requires a power-of-two channel-group count; anything else takes the per-chunk kernels.
// ------------------------------------------------------------------------------------------------------------
struct RowWalk {
    uint32_t unit, step, nunits;
    int lgG, lgUPR, lgCG, G;
    int cg;                  // channel group of this lane (constant over the walk)
    bool lane_on;            // lane inside the row (rows shorter than 32 chunks leave lanes idle)
    uint32_t vox;            // voxel of this lane inside the row
    long row_stride;         // elements between consecutive y-rows
};

__device__ __forceinline__ int ilog2_dev(int v) { return 31 - __clz(v); }

__device__ __forceinline__ RowWalk make_row_walk(const Geo& g) {
    RowWalk w;
    const int CG = g.C >> 3;
    w.lgCG = ilog2_dev(CG);
    const int R = g.P * CG;                                  // chunks per patch row
    const int UPR = R > 32 ? (R >> 5) : 1;
    w.lgUPR = ilog2_dev(UPR);
    w.G = g.P < 4 ? g.P : 4;
    w.lgG = ilog2_dev(w.G);
    const uint32_t warp = ((uint32_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t nwarps = ((uint32_t)gridDim.x * blockDim.x) >> 5;
    nwarps &= ~(uint32_t)(UPR - 1);                          // stride a multiple of UPR: `part` (and so cg) stays constant
    const int lane = threadIdx.x & 31;
    const uint32_t part = warp & (uint32_t)(UPR - 1);
    const uint32_t c = part * 32u + (uint32_t)lane;
    w.lane_on = c < (uint32_t)R;
    w.cg = (int)(c & (uint32_t)(CG - 1));
    w.vox = c >> w.lgCG;
    w.unit = warp;
    w.step = nwarps;
    w.nunits = (nwarps == 0 || warp >= nwarps) ? 0u : ((uint32_t)(*g.count) << (g.lgP + (g.lgP - w.lgG) + w.lgUPR));
    w.row_stride = (long)g.W * g.C;
    return w;
}

// element offset of this lane's chunk in the first row of the group
__device__ __forceinline__ long row_unit_offset(const Geo& g, const RowWalk& w, uint32_t u) {
    uint32_t t = u >> w.lgUPR;
    const uint32_t yg = t & (((uint32_t)g.P >> w.lgG) - 1u);
    t >>= (g.lgP - w.lgG);
    const uint32_t rz = t & ((uint32_t)g.P - 1u);
    const uint32_t pid = (uint32_t)g.list[t >> g.lgP];
    const uint32_t L = (uint32_t)(g.fd * g.fh * g.fw), hw = (uint32_t)(g.fh * g.fw);
    const uint32_t n = pid / L, l = pid - n * L;
    const uint32_t pz = l / hw, r2 = l - pz * hw;
    const uint32_t py = r2 / (uint32_t)g.fw, px = r2 - py * (uint32_t)g.fw;
    const long voxel = (((long)n * g.D + (pz << g.lgP) + rz) * g.H + (py << g.lgP) + (yg << w.lgG)) * g.W +
                       ((long)px << g.lgP) + w.vox;
    return voxel * g.C + (long)w.cg * 8;
}

// G (rows per group) is a template parameter: the G·(1..3) 128-bit loads of a group are issued back to back before any of
// them is consumed (memory-level parallelism per thread instead of per-SM thread count).
template <int MODE, int ACT, int G>
__global__ void __launch_bounds__(256, 2) reduce_rows_kernel(Geo g, const bf16* __restrict__ x, const bf16* __restrict__ dout,
                                                          const bf16* __restrict__ res, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, const float* __restrict__ saved,
                                                          double* __restrict__ sums) {
    extern __shared__ float sacc[];           // [2][C]
    const int CG = g.C / 8;
    for (int i = threadIdx.x; i < g.C * 2; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
    RowWalk w = make_row_walk(g);
    const int cg = w.cg;
    float sc[8], sh[8];
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = scale[cg * 8 + j]; sh[j] = shift[cg * 8 + j]; }
    }
    if (w.lane_on) {
        for (uint32_t u = w.unit; u < w.nunits; u += w.step) {
            const long off0 = row_unit_offset(g, w, u);
            uint4 vx[G], vd[G], vr[G];
#pragma unroll
            for (int r = 0; r < G; ++r) {
                const long off = off0 + r * w.row_stride;
                vx[r] = __ldg(reinterpret_cast<const uint4*>(x + off));
                if (MODE == 1) {
                    vd[r] = __ldg(reinterpret_cast<const uint4*>(dout + off));
                    if (res) vr[r] = __ldg(reinterpret_cast<const uint4*>(res + off));
                }
            }
#pragma unroll
            for (int r = 0; r < G; ++r) {
                float f[8];
                unpack_u4(vx[r], f);
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { a0[j] += f[j]; a1[j] += f[j] * f[j]; }
                } else {
                    float d[8], rr[8];
                    unpack_u4(vd[r], d);
                    if (res) unpack_u4(vr[r], rr);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float uu = sc[j] * f[j] + sh[j] + (res ? rr[j] : 0.f);
                        float gj = d[j] * act_grad(uu, ACT);
                        a0[j] += gj;
                        a1[j] += gj * f[j];
                    }
                }
            }
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            a1[j] = saved[g.C + cg * 8 + j] * (a1[j] - saved[cg * 8 + j] * a0[j]);
    }
    if (CG < 32) {                            // lanes that share a channel group (lane % CG equal): shuffle first
        for (int o = CG; o < 32; o <<= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], o);
                a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], o);
            }
        }
    }
    if (CG >= 32 || (threadIdx.x & 31) < CG) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sacc[cg * 8 + j], a0[j]);
            atomicAdd(&sacc[g.C + cg * 8 + j], a1[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g.C; i += blockDim.x) {
        atomicAdd(&sums[i], (double)sacc[i]);
        atomicAdd(&sums[g.C + i], (double)sacc[g.C + i]);
    }
}

template <int ACT, int G>
__global__ void __launch_bounds__(256, 3) apply_rows_kernel(Geo g, const bf16* __restrict__ x, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, const bf16* __restrict__ res,
                                                         bf16* __restrict__ out) {
    RowWalk w = make_row_walk(g);
    if (!w.lane_on) return;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale[w.cg * 8 + j]; sh[j] = shift[w.cg * 8 + j]; }
    for (uint32_t u = w.unit; u < w.nunits; u += w.step) {
        const long off0 = row_unit_offset(g, w, u);
        uint4 vx[G], vr[G];
#pragma unroll
        for (int r = 0; r < G; ++r) {
            const long off = off0 + r * w.row_stride;
            vx[r] = __ldg(reinterpret_cast<const uint4*>(x + off));
            if (res) vr[r] = __ldg(reinterpret_cast<const uint4*>(res + off));
        }
#pragma unroll
        for (int r = 0; r < G; ++r) {
            float f[8], rr[8], o[8];
            unpack_u4(vx[r], f);
            if (res) unpack_u4(vr[r], rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = act_fwd(sc[j] * f[j] + sh[j] + (res ? rr[j] : 0.f), ACT);
            store8(out + off0 + r * w.row_stride, o);
        }
    }
}

template <int ACT, int G>
__global__ void __launch_bounds__(256, 2)
bwd_apply_rows_kernel(Geo g, const bf16* __restrict__ dout, const bf16* __restrict__ x, const bf16* __restrict__ res,
                      const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ saved,
                      const double* __restrict__ sums, bf16* __restrict__ dx, bf16* __restrict__ dres,
                      float* __restrict__ dgamma, float* __restrict__ dbeta, const double* n_total) {
    if (blockIdx.x == 0 && dgamma) {
        for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
            dbeta[c] = (float)sums[c];
            dgamma[c] = (float)sums[g.C + c];
        }
    }
    RowWalk w = make_row_walk(g);
    if (!w.lane_on) return;
    const double n = n_total ? *n_total : (double)((long)(*g.count) << (3 * g.lgP));
    float sc[8], sh[8], ca[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int c = w.cg * 8 + j;
        sc[j] = scale[c]; sh[j] = shift[c];
        const float mu = saved[c], rs = saved[g.C + c];
        const float m1 = (float)(sums[c] / n), m2 = (float)(sums[g.C + c] / n);
        ca[j] = -sc[j] * m2 * rs;
        cb[j] = -sc[j] * m1 - ca[j] * mu;
    }
    for (uint32_t u = w.unit; u < w.nunits; u += w.step) {
        const long off0 = row_unit_offset(g, w, u);
        uint4 vx[G], vd[G], vr[G];
#pragma unroll
        for (int r = 0; r < G; ++r) {
            const long off = off0 + r * w.row_stride;
            vd[r] = __ldg(reinterpret_cast<const uint4*>(dout + off));
            vx[r] = __ldg(reinterpret_cast<const uint4*>(x + off));
            if (res) vr[r] = __ldg(reinterpret_cast<const uint4*>(res + off));
        }
#pragma unroll
        for (int r = 0; r < G; ++r) {
            const long off = off0 + r * w.row_stride;
            float d[8], f[8], rr[8], o[8], gg[8];
            unpack_u4(vd[r], d);
            unpack_u4(vx[r], f);
            if (res) unpack_u4(vr[r], rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float uu = sc[j] * f[j] + sh[j] + (res ? rr[j] : 0.f);
                gg[j] = d[j] * act_grad(uu, ACT);
                o[j] = fmaf(sc[j], gg[j], fmaf(ca[j], f[j], cb[j]));
            }
            store8(dx + off, o);
            if (dres) store8(dres + off, gg);
        }
    }
}

// Engine mode (ops.LEAN_ZERO): a sparse tensor whose consumers all walk the active-patch list is read at most ONE voxel beyond
// a visible patch (3x3x3 halo), so instead of zero-filling the whole tensor only the 1-voxel shell of every visible patch that
// lies inside masked patches is cleared.  One CTA per visible patch.
__device__ __forceinline__ void zero_slab(const Geo& g, bf16* __restrict__ x, uint32_t n, int pz, int py, int px, int dir,
                                          int first, int stride) {
    const int P = g.P, CG = g.C / 8;
    const int dz = dir / 9 - 1, dy = (dir / 3) % 3 - 1, dx = dir % 3 - 1;
    const int lz = dz ? 0 : g.lgP, ly = dy ? 0 : g.lgP, lx = dx ? 0 : g.lgP;        // log2 extents of the slab
    const int z0 = (pz << g.lgP) + (dz < 0 ? -1 : (dz > 0 ? P : 0));
    const int y0 = (py << g.lgP) + (dy < 0 ? -1 : (dy > 0 ? P : 0));
    const int x0 = (px << g.lgP) + (dx < 0 ? -1 : (dx > 0 ? P : 0));
    const int items = CG << (lz + ly + lx);
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    for (int i = first; i < items; i += stride) {
        const int cg = i % CG;
        const int v = i / CG;
        const int vx = v & ((1 << lx) - 1), vy = (v >> lx) & ((1 << ly) - 1), vz = v >> (lx + ly);
        const long voxel = (((long)n * g.D + z0 + vz) * g.H + y0 + vy) * g.W + x0 + vx;
        *reinterpret_cast<uint4*>(x + voxel * g.C + cg * 8) = z4;
    }
}

__global__ void __launch_bounds__(256) zero_shell_kernel(Geo g, bf16* __restrict__ x) {
    const uint32_t L = (uint32_t)(g.fd * g.fh * g.fw), hw = (uint32_t)(g.fh * g.fw);
    const int n_entries = *g.count;
    // the shell splits into 26 slabs, one per neighbouring patch (face: P x P voxels, edge: P, corner: 1); a slab is cleared
    // only when that neighbour exists and is masked.  The 26 visibility bytes are fetched by 26 threads at once (26 slab loops
    // each behind its own load were latency-bound); the 6 faces are then cleared by the whole block, the 20 edges / corners by
    // one warp each.
    __shared__ unsigned char s_flag[27];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = blockIdx.x; e < n_entries; e += gridDim.x) {
        const uint32_t pid = (uint32_t)g.list[e];
        const uint32_t n = pid / L, l = pid - n * L;
        const int pz = (int)(l / hw), r2 = (int)(l - (uint32_t)pz * hw);
        const int py = r2 / g.fw, px = r2 - py * g.fw;
        __syncthreads();
        if (threadIdx.x < 27) {
            const int dir = threadIdx.x;
            const int qz = pz + dir / 9 - 1, qy = py + (dir / 3) % 3 - 1, qx = px + dir % 3 - 1;
            s_flag[dir] = dir != 13 && qz >= 0 && qz < g.fd && qy >= 0 && qy < g.fh && qx >= 0 && qx < g.fw &&
                          !g.active[((n * g.fd + qz) * g.fh + qy) * g.fw + qx];
        }
        __syncthreads();
        int k = 0;
        for (int dir = 0; dir < 27; ++dir) {
            const int nz = (dir / 9 != 1) + ((dir / 3) % 3 != 1) + (dir % 3 != 1);      // 1 = face, 2 = edge, 3 = corner
            if (nz == 1) {
                if (s_flag[dir]) zero_slab(g, x, n, pz, py, px, dir, threadIdx.x, blockDim.x);
            } else if (nz >= 2) {
                if ((k++ & 7) == warp && s_flag[dir]) zero_slab(g, x, n, pz, py, px, dir, lane, 32);
            }
        }
    }
}

// fine[n, 2z, 2y, 2x, :] += coarse[n, z, y, x, :] — the input gradient of a 1x1 stride-2 convolution lands on the even voxels
// only; adding it in place replaces a zero-filled full-resolution tensor plus a full-resolution add (the residual shortcut,
// P/STUNet_head.py:96-101).  `g` describes the COARSE tensor (walks its active-patch list when it has one).
template <bool LIST>
__global__ void __launch_bounds__(256) add_parity0_kernel(Geo g, const bf16* __restrict__ coarse, bf16* __restrict__ fine) {
    const int CG = g.C / 8;
    const int cg = threadIdx.x % CG;
    Walk w = make_walk(g, CG);
    for (; w.slot < w.nslots; w.slot += w.step) {
        const long v = slot_voxel<LIST>(g, w.slot);
        uint32_t t = (uint32_t)v;
        const uint32_t xx = t % (uint32_t)g.W; t /= (uint32_t)g.W;
        const uint32_t y = t % (uint32_t)g.H; t /= (uint32_t)g.H;
        const uint32_t z = t % (uint32_t)g.D;
        const uint32_t n = t / (uint32_t)g.D;
        const long fv = (((long)n * (2 * g.D) + 2 * z) * (2 * g.H) + 2 * y) * (2 * g.W) + 2 * xx;
        float a[8], b[8];
        load8(coarse + v * g.C + cg * 8, a);
        const uint4 old = *reinterpret_cast<const uint4*>(fine + fv * g.C + cg * 8);
        unpack_u4(old, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += b[j];
        store8(fine + fv * g.C + cg * 8, a);
    }
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, long n8) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        float fa[8], fb[8];
        load8(a + i * 8, fa);
        load8(b + i * 8, fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] += fb[j];
        store8(out + i * 8, fa);
    }
}

// NCDHW fp32 → NDHWC bf16 through a padded smem tile (coalesced on both sides)
__global__ void to_ndhwc_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int C, long S) {
    __shared__ float tile[32][33];
    const long n = blockIdx.z;
    const long s0 = (long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i;
        long s = s0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && s < S) ? src[(n * C + c) * S + s] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        long s = s0 + i;
        int c = c0 + threadIdx.x;
        if (c < C && s < S) dst[(n * S + s) * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
    }
}

__global__ void to_ncdhw_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int C, long S) {
    __shared__ float tile[32][33];
    const long n = blockIdx.z;
    const long s0 = (long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        long s = s0 + i;
        int c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && s < S) ? bf2f(src[(n * S + s) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i;
        long s = s0 + threadIdx.x;
        if (c < C && s < S) dst[(n * C + c) * S + s] = tile[threadIdx.x][i];
    }
}

// deterministic compaction of the (N·L) visibility bytes into an ascending list of active patch ids
__global__ void active_list_kernel(const uint8_t* __restrict__ active, int n, int* __restrict__ list,
                                   int* __restrict__ count) {
    __shared__ int warp_tot[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int start = 0; start < n; start += blockDim.x) {
        int i = start + threadIdx.x;
        int a = (i < n) && active[i] != 0;
        unsigned bal = __ballot_sync(0xffffffffu, a);
        int pre = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_tot[w];
        if (a) list[base + woff + pre] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < nw; ++w) t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

static int grid_for(long work_items, int block) {
    long b = (work_items + block - 1) / block;
    long cap = (long)num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

static bool rows_ok(const Geo& g) {
    const int CG = g.C / 8;
    return g.list != nullptr && (CG & (CG - 1)) == 0 && getenv("AMB_NORM_NO_ROWS") == nullptr;
}
// persistent grid of the row-group kernels: every resident CTA slot of the device once (a second, partial wave of a
// bandwidth-bound kernel runs at a fraction of the occupancy), never more warps than row groups; with more than 32 channel
// groups the warp count must cover at least one full row (UPR warps)
template <typename K>
static int rows_grid(const Geo& g, K kernel, size_t smem) {
    const int CG = g.C / 8;
    const long R = (long)g.P * CG, UPR = R > 32 ? R / 32 : 1;
    const int G = g.P < 4 ? g.P : 4;
    const long units = (long)g.N * g.fd * g.fh * g.fw * g.P * (g.P / G) * UPR;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem) != cudaSuccess || occ < 1) occ = 2;
    long ctas = (units + 7) / 8;
    const long cap = (long)num_sms() * occ;
    if (ctas > cap) ctas = cap;
    const long min_ctas = (UPR + 7) / 8;
    if (ctas < min_ctas) ctas = min_ctas;
    return (int)ctas;
}
#define AMB_ROWS_G(LAUNCH)                           \
    do {                                             \
        if (g.P >= 4) { LAUNCH(4); }                 \
        else if (g.P == 2) { LAUNCH(2); }            \
        else { LAUNCH(1); }                          \
    } while (0)

// grid of the per-chunk reduce kernels.  Every CTA ends with 2·C fp64 atomics onto the same 2·C addresses, so a small tensor
// spread over thousands of CTAs is bound by that tail (33 MB took 50 µs): at least 32 chunks per thread, and never more CTAs
// than are resident at once.
template <typename K>
static int reduce_grid(long work_items, int block, K kernel, size_t smem) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem) != cudaSuccess || occ < 1) occ = 2;
    long b = (work_items + (long)block * 32 - 1) / ((long)block * 32);
    const long cap = (long)num_sms() * occ;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

template <typename K>
static int dense_grid(long chunks, int block, K kernel, size_t smem) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem) != cudaSuccess || occ < 1) occ = 1;
    long b = (chunks + (long)block * 8 - 1) / ((long)block * 8);
    const long cap = (long)num_sms() * occ;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}
static bool dense_batched_ok() { return getenv("AMB_NORM_NO_DENSE_BATCH") == nullptr; }

static long host_items_upper(const Geo& g) { return (long)g.N * g.D * g.H * g.W * (g.C / 8); }

}  // namespace amb

using namespace amb;

extern "C" int amb_build_active_list(const uint8_t* active, int n_patches, int* list, int* count, void* stream) {
    AMB_CHECK(active && list && count && n_patches > 0, AMB_ERR_ARG, "amb_build_active_list: null argument");
    active_list_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(active, n_patches, list, count);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int N, int C, int D, int H, int W,
                                           void* stream) {
    long S = (long)D * H * W;
    dim3 grid(ceil_div(S, 32), ceil_div(C, 32), N), block(32, 8);
    to_ndhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, C, S);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int N, int C, int D, int H, int W,
                                           void* stream) {
    long S = (long)D * H * W;
    dim3 grid(ceil_div(S, 32), ceil_div(C, 32), N), block(32, 8);
    to_ncdhw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, C, S);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_norm_stats(const amb_geo* a, const void* x, double* sums, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    int CG = g.C / 8, block = pick_block(CG);
    if (block < 0) return block;
    if (rows_ok(g)) {
        const size_t sm = g.C * 2 * sizeof(float);
#define AMB_STATS_ROWS(GG)                                                                                     \
    reduce_rows_kernel<0, 0, GG><<<rows_grid(g, reduce_rows_kernel<0, 0, GG>, sm), 256, sm, (cudaStream_t)stream>>>( \
        g, (const bf16*)x, nullptr, nullptr, nullptr, nullptr, nullptr, sums)
        AMB_ROWS_G(AMB_STATS_ROWS);
#undef AMB_STATS_ROWS
    } else if (g.list)
        reduce_kernel<0, 0, true, false><<<reduce_grid(host_items_upper(g), block, reduce_kernel<0, 0, true, false>, g.C * 3 * sizeof(float)), block, g.C * 3 * sizeof(float), (cudaStream_t)stream>>>(
            g, (const bf16*)x, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, sums, nullptr);
    else if (dense_batched_ok()) {
        const size_t sm = g.C * 2 * sizeof(float);
        reduce_dense_kernel<0, 0><<<reduce_grid(host_items_upper(g), block, reduce_dense_kernel<0, 0>, sm), block, sm, (cudaStream_t)stream>>>(
            g, (const bf16*)x, nullptr, nullptr, nullptr, nullptr, nullptr, sums);
    } else
        reduce_kernel<0, 0, false, false><<<reduce_grid(host_items_upper(g), block, reduce_kernel<0, 0, false, false>, g.C * 3 * sizeof(float)), block, g.C * 3 * sizeof(float), (cudaStream_t)stream>>>(
            g, (const bf16*)x, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, sums, nullptr);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_norm_finalize(const amb_geo* a, const double* sums, const float* gamma, const float* beta,
                                 float eps, float* scale, float* shift, float* saved, float* running_mean,
                                 float* running_var, long* nbt, float momentum, const double* n_total, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    finalize_kernel<<<ceil_div(g.C, 128), 128, 0, (cudaStream_t)stream>>>(g, sums, gamma, beta, eps, scale, shift,
                                                                          saved, running_mean, running_var, nbt,
                                                                          momentum, n_total);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_norm_eval(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                             float* scale, float* shift, int C, void* stream) {
    eval_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, rm, rv, eps, scale, shift, C);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_norm_apply(const amb_geo* a, const void* x, const float* scale, const float* shift,
                              const void* residual, const float* token, int act, void* out, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    AMB_CHECK(!token || g.active, AMB_ERR_ARG, "densify fill needs the active mask");
    int CG = g.C / 8, block = pick_block(CG);
    if (block < 0) return block;
#define AMB_APPLY(A, LI, FI)                                                                                \
    apply_kernel<A, LI, FI><<<grid_for(host_items_upper(g), block), block, 0, (cudaStream_t)stream>>>(        \
        g, (const bf16*)x, scale, shift, (const bf16*)residual, token, act, (bf16*)out)
    if (token) { AMB_APPLY(0, false, true); }                       // densify fill: every voxel, no activation
    else if (rows_ok(g)) {
#define AMB_APPLY_ROWS(A, GG)                                                                                 \
    apply_rows_kernel<A, GG><<<rows_grid(g, apply_rows_kernel<A, GG>, 0), 256, 0, (cudaStream_t)stream>>>(        \
        g, (const bf16*)x, scale, shift, (const bf16*)residual, (bf16*)out)
#define AMB_APPLY_ROWS1(GG) AMB_APPLY_ROWS(1, GG)
#define AMB_APPLY_ROWS2(GG) AMB_APPLY_ROWS(2, GG)
#define AMB_APPLY_ROWS0(GG) AMB_APPLY_ROWS(0, GG)
        if (act == AMB_ACT_LRELU) AMB_ROWS_G(AMB_APPLY_ROWS1);
        else if (act == AMB_ACT_RELU6) AMB_ROWS_G(AMB_APPLY_ROWS2);
        else AMB_ROWS_G(AMB_APPLY_ROWS0);
#undef AMB_APPLY_ROWS
#undef AMB_APPLY_ROWS0
#undef AMB_APPLY_ROWS1
#undef AMB_APPLY_ROWS2
    } else if (g.list) {
        if (act == AMB_ACT_LRELU) AMB_APPLY(1, true, false);
        else if (act == AMB_ACT_RELU6) AMB_APPLY(2, true, false);
        else AMB_APPLY(0, true, false);
    } else if (dense_batched_ok()) {
#define AMB_APPLY_DENSE(A)                                                                                         \
    apply_dense_kernel<A><<<dense_grid(host_items_upper(g), block, apply_dense_kernel<A>, 0), block, 0, (cudaStream_t)stream>>>( \
        g, (const bf16*)x, scale, shift, (const bf16*)residual, (bf16*)out)
        if (act == AMB_ACT_LRELU) AMB_APPLY_DENSE(1);
        else if (act == AMB_ACT_RELU6) AMB_APPLY_DENSE(2);
        else AMB_APPLY_DENSE(0);
#undef AMB_APPLY_DENSE
    } else {
        if (act == AMB_ACT_LRELU) AMB_APPLY(1, false, false);
        else if (act == AMB_ACT_RELU6) AMB_APPLY(2, false, false);
        else AMB_APPLY(0, false, false);
    }
#undef AMB_APPLY
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_norm_bwd_reduce(const amb_geo* a, const void* dout, const void* x, const void* residual,
                                   const float* scale, const float* shift, const float* saved, int act, int fill,
                                   double* sums, double* dtoken, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    AMB_CHECK(!fill || g.active, AMB_ERR_ARG, "densify backward needs the active mask");
    int CG = g.C / 8, block = pick_block(CG);
    if (block < 0) return block;
#define AMB_RED(A, LI, FI)                                                                                                       \
    reduce_kernel<1, A, LI, FI><<<reduce_grid(host_items_upper(g), block, reduce_kernel<1, A, LI, FI>, g.C * 3 * sizeof(float)), block, g.C * 3 * sizeof(float), (cudaStream_t)stream>>>( \
        g, (const bf16*)x, (const bf16*)dout, (const bf16*)residual, scale, shift, saved, act, fill, sums, dtoken)
    if (fill) { AMB_RED(0, false, true); }
    else if (rows_ok(g)) {
        const size_t sm = g.C * 2 * sizeof(float);
#define AMB_RED_ROWS(A, GG)                                                                                     \
    reduce_rows_kernel<1, A, GG><<<rows_grid(g, reduce_rows_kernel<1, A, GG>, sm), 256, sm, (cudaStream_t)stream>>>( \
        g, (const bf16*)x, (const bf16*)dout, (const bf16*)residual, scale, shift, saved, sums)
#define AMB_RED_ROWS1(GG) AMB_RED_ROWS(1, GG)
#define AMB_RED_ROWS2(GG) AMB_RED_ROWS(2, GG)
#define AMB_RED_ROWS0(GG) AMB_RED_ROWS(0, GG)
        if (act == AMB_ACT_LRELU) AMB_ROWS_G(AMB_RED_ROWS1);
        else if (act == AMB_ACT_RELU6) AMB_ROWS_G(AMB_RED_ROWS2);
        else AMB_ROWS_G(AMB_RED_ROWS0);
#undef AMB_RED_ROWS
#undef AMB_RED_ROWS0
#undef AMB_RED_ROWS1
#undef AMB_RED_ROWS2
    } else if (g.list) {
        if (act == AMB_ACT_LRELU) AMB_RED(1, true, false);
        else if (act == AMB_ACT_RELU6) AMB_RED(2, true, false);
        else AMB_RED(0, true, false);
    } else if (dense_batched_ok()) {
        const size_t sm = g.C * 2 * sizeof(float);
#define AMB_RED_DENSE(A)                                                                                          \
    reduce_dense_kernel<1, A><<<reduce_grid(host_items_upper(g), block, reduce_dense_kernel<1, A>, sm), block, sm, (cudaStream_t)stream>>>( \
        g, (const bf16*)x, (const bf16*)dout, (const bf16*)residual, scale, shift, saved, sums)
        if (act == AMB_ACT_LRELU) AMB_RED_DENSE(1);
        else if (act == AMB_ACT_RELU6) AMB_RED_DENSE(2);
        else AMB_RED_DENSE(0);
#undef AMB_RED_DENSE
    } else {
        if (act == AMB_ACT_LRELU) AMB_RED(1, false, false);
        else if (act == AMB_ACT_RELU6) AMB_RED(2, false, false);
        else AMB_RED(0, false, false);
    }
#undef AMB_RED
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_norm_bwd_apply(const amb_geo* a, const void* dout, const void* x, const void* residual,
                                  const float* scale, const float* shift, const float* saved, const double* sums,
                                  int act, int fill, void* dx, void* dres, float* dgamma, float* dbeta,
                                  const double* n_total, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    (void)fill;   // dx is only defined on visited (active) voxels; the caller zero-fills dx when the list is sparse
    int CG = g.C / 8, block = pick_block(CG);
    if (block < 0) return block;
#define AMB_BWD(A, LI)                                                                                          \
    bwd_apply_kernel<A, LI><<<grid_for(host_items_upper(g), block), block, 0, (cudaStream_t)stream>>>(          \
        g, (const bf16*)dout, (const bf16*)x, (const bf16*)residual, scale, shift, saved, sums, act, fill, (bf16*)dx, \
        (bf16*)dres, dgamma, dbeta, n_total)
    if (rows_ok(g)) {
#define AMB_BWD_ROWS(A, GG)                                                                                    \
    bwd_apply_rows_kernel<A, GG><<<rows_grid(g, bwd_apply_rows_kernel<A, GG>, 0), 256, 0, (cudaStream_t)stream>>>( \
        g, (const bf16*)dout, (const bf16*)x, (const bf16*)residual, scale, shift, saved, sums, (bf16*)dx,      \
        (bf16*)dres, dgamma, dbeta, n_total)
#define AMB_BWD_ROWS1(GG) AMB_BWD_ROWS(1, GG)
#define AMB_BWD_ROWS2(GG) AMB_BWD_ROWS(2, GG)
#define AMB_BWD_ROWS0(GG) AMB_BWD_ROWS(0, GG)
        if (act == AMB_ACT_LRELU) AMB_ROWS_G(AMB_BWD_ROWS1);
        else if (act == AMB_ACT_RELU6) AMB_ROWS_G(AMB_BWD_ROWS2);
        else AMB_ROWS_G(AMB_BWD_ROWS0);
#undef AMB_BWD_ROWS
#undef AMB_BWD_ROWS0
#undef AMB_BWD_ROWS1
#undef AMB_BWD_ROWS2
    } else if (g.list) {
        if (act == AMB_ACT_LRELU) AMB_BWD(1, true);
        else if (act == AMB_ACT_RELU6) AMB_BWD(2, true);
        else AMB_BWD(0, true);
    } else if (dense_batched_ok()) {
#define AMB_BWD_DENSE(A)                                                                                           \
    bwd_apply_dense_kernel<A><<<dense_grid(host_items_upper(g), block, bwd_apply_dense_kernel<A>, 0), block, 0, (cudaStream_t)stream>>>( \
        g, (const bf16*)dout, (const bf16*)x, (const bf16*)residual, scale, shift, saved, sums, (bf16*)dx,          \
        (bf16*)dres, dgamma, dbeta, n_total)
        if (act == AMB_ACT_LRELU) AMB_BWD_DENSE(1);
        else if (act == AMB_ACT_RELU6) AMB_BWD_DENSE(2);
        else AMB_BWD_DENSE(0);
#undef AMB_BWD_DENSE
    } else {
        if (act == AMB_ACT_LRELU) AMB_BWD(1, false);
        else if (act == AMB_ACT_RELU6) AMB_BWD(2, false);
        else AMB_BWD(0, false);
    }
#undef AMB_BWD
    AMB_LAUNCH_CHECK();
    return 0;
}

__global__ void count_voxels_kernel(Geo g, double* out) {
    *out = g.list ? (double)((long)(*g.count) << (3 * g.lgP)) : (double)g.N * g.D * g.H * g.W;
}

extern "C" int amb_count_voxels(const amb_geo* a, double* out, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    count_voxels_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(g, out);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_add(const void* a, const void* b, void* out, long n, void* stream) {
    AMB_CHECK(n % 8 == 0, AMB_ERR_ARG, "amb_add: n must be a multiple of 8");
    add_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n / 8);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_zero_shell(const amb_geo* a, void* x, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    AMB_CHECK(g.list && g.active && x, AMB_ERR_ARG, "amb_zero_shell needs the active mask and its work-list");
    int ctas = g.N * g.fd * g.fh * g.fw;
    if (ctas > num_sms() * 8) ctas = num_sms() * 8;
    zero_shell_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(g, (bf16*)x);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_add_parity0(const amb_geo* a, const void* coarse, void* fine, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    AMB_CHECK(coarse && fine, AMB_ERR_ARG, "amb_add_parity0: null argument");
    int CG = g.C / 8, block = pick_block(CG);
    if (block < 0) return block;
    if (g.list)
        add_parity0_kernel<true><<<grid_for(host_items_upper(g), block), block, 0, (cudaStream_t)stream>>>(g, (const bf16*)coarse, (bf16*)fine);
    else
        add_parity0_kernel<false><<<grid_for(host_items_upper(g), block), block, 0, (cudaStream_t)stream>>>(g, (const bf16*)coarse, (bf16*)fine);
    AMB_LAUNCH_CHECK();
    return 0;
}
