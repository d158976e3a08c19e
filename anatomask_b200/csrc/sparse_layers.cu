// The rest of the reference's sparse-layer API (P/encoder3D.py:30-37,47-78,181-276 — SURVEY §8f row 4): the layers the
// MedNeXt / ConvNeXt heads need next to the STUNet path.  All HBM- or issue-bound CUDA-core kernels on bf16
// channels-last activations, one 16-byte item = 8 channels of one voxel.
//
//   voxel_norm      SparseGroupNorm / SparseConvNeXtLayerNorm: the reference feeds the visible voxels as an (N_active, C)
//                   matrix to nn.GroupNorm / nn.LayerNorm, i.e. EVERY VOXEL is normalised on its own over each channel
//                   group (G groups; LayerNorm = 1 group) with a per-channel affine — P/encoder3D.py:47-78,193-243
//   pool3d          SparseMaxPooling / SparseAvgPooling: pool, then multiply by the mask at the output resolution
//   masked_mean     SparseAdaptiveAvgPooling(1): Σ x·mask / (Σ mask + 1e-6) per sample and channel
//   dwconv          depthwise k³ convolution (k ∈ {3,5,7}, stride 1 / 2, pad k/2) forward, input and weight gradients —
//                   SparseConvNeXtBlock.dwconv, MedNeXtBlock.conv1 (P/encoder3D.py:259, P/MedNeXt_head.py:255-262)
//   gelu            exact (erf) GELU forward / backward
//   layer_scale     ConvNeXt tail: out = input + mask·γ_c·x (P/encoder3D.py:270-279)
#include "common.cuh"

namespace amb {

int make_geo(const amb_geo* a, Geo& g);

static inline int grid_cap(long items, int block, int per_sm = 8) {
    long b = (items + block - 1) / block;
    long cap = (long)num_sms() * per_sm;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------------------
// per-voxel group norm
// ------------------------------------------------------------------------------------------------------------
#define VN_MAXCH 4      // 16-byte chunks per thread: C <= 4 · 32 · 8 = 1024

template <bool LIST>
__device__ __forceinline__ long vn_voxel(const Geo& g, long slot) {
    if (!LIST) return slot;
    const uint32_t P1 = (uint32_t)g.P - 1u;
    const uint32_t s = (uint32_t)slot;
    const uint32_t v = s & P1, ry = (s >> g.lgP) & P1, rz = (s >> (2 * g.lgP)) & P1;
    const uint32_t pid = (uint32_t)g.list[s >> (3 * g.lgP)];
    const uint32_t L = (uint32_t)(g.fd * g.fh * g.fw), hw = (uint32_t)(g.fh * g.fw);
    const uint32_t n = pid / L, l = pid - n * L;
    const uint32_t pz = l / hw, r2 = l - pz * hw;
    const uint32_t py = r2 / (uint32_t)g.fw, px = r2 - py * (uint32_t)g.fw;
    return (((long)n * g.D + (pz << g.lgP) + rz) * g.H + (py << g.lgP) + ry) * g.W + ((long)px << g.lgP) + v;
}

// TPV threads (a power of two <= 32, lanes of one warp) share a voxel; thread t owns chunks t, t + TPV, ...
// Group statistics: one group → shuffles; groups of <= 8 channels → inside a thread; wider groups → a per-voxel
// shared-memory table.  BWD additionally keeps per-thread Σg·x̂ / Σg per owned channel (dgamma / dbeta).
template <bool LIST, bool BWD>
__global__ void __launch_bounds__(256, 1)
voxel_norm_kernel(Geo g, int TPV, int groups, float eps, const bf16* __restrict__ x, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const bf16* __restrict__ dout, bf16* __restrict__ out,
                  float* __restrict__ dgamma, float* __restrict__ dbeta) {
    extern __shared__ float vsm[];                       // [2][C] dgamma / dbeta partials (BWD), then wide groups: [voxels per block][groups][2]
    const int CG = g.C / 8, gs = g.C / groups;
    const int vl = threadIdx.x % TPV, vib = threadIdx.x / TPV, vpb = blockDim.x / TPV;
    const long nslots = g.list ? ((long)(*g.count) << (3 * g.lgP)) : (long)g.N * g.D * g.H * g.W;
    const long step = (long)gridDim.x * vpb;
    const int mode = groups == 1 ? 0 : (gs <= 8 ? 1 : 2);
    float* red = vsm;
    float* tab = vsm + (BWD ? 2 * g.C : 0) + (size_t)vib * groups * 2;
    if (BWD) {
        for (int i = threadIdx.x; i < 2 * g.C; i += blockDim.x) red[i] = 0.f;
        __syncthreads();
    }
    float ga[VN_MAXCH][8], ag[VN_MAXCH][8], ab[VN_MAXCH][8];
#pragma unroll
    for (int i = 0; i < VN_MAXCH; ++i) {
        const int cg = vl + i * TPV;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ga[i][j] = cg < CG ? gamma[cg * 8 + j] : 0.f;
            ag[i][j] = ab[i][j] = 0.f;
        }
    }
    // all lanes of a warp stay in the loop until the warp's last voxel is done (shuffles / __syncwarp below)
    for (long base = (long)blockIdx.x * vpb; base < nslots; base += step) {       // block-uniform trip count
        const long slot = base + vib;
        const bool valid = slot < nslots;
        const long voxel = valid ? vn_voxel<LIST>(g, slot) : 0;
        const bool vis = valid && (LIST || g.active == nullptr || voxel_active(g, voxel));
        float xv[VN_MAXCH][8];
#pragma unroll
        for (int i = 0; i < VN_MAXCH; ++i) {
            const int cg = vl + i * TPV;
            if (vis && cg < CG) load8(x + voxel * g.C + cg * 8, xv[i]);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) xv[i][j] = 0.f;
            }
        }
        // ---- statistics: mean / rstd per (chunk, sub-group) into mu[i][j], rs[i][j] (broadcast per channel) -------
        float mu[VN_MAXCH][8], rs[VN_MAXCH][8];
        if (mode == 0) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) { s += xv[i][j]; q += xv[i][j] * xv[i][j]; }
            for (int o = 1; o < TPV; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
            const float m = s / g.C, var = fmaxf(q / g.C - m * m, 0.f), r = rsqrtf(var + eps);
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) { mu[i][j] = m; rs[i][j] = r; }
        } else if (mode == 1) {
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i)
#pragma unroll
                for (int j0 = 0; j0 < 8; ++j0) {
                    float s = 0.f, q = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j / gs == j0 / gs) { s += xv[i][j]; q += xv[i][j] * xv[i][j]; }
                    const float m = s / gs, var = fmaxf(q / gs - m * m, 0.f);
                    mu[i][j0] = m; rs[i][j0] = rsqrtf(var + eps);
                }
        } else {
            for (int k = vl; k < groups * 2; k += TPV) tab[k] = 0.f;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i) {
                const int cg = vl + i * TPV;
                if (cg < CG) {
                    float s = 0.f, q = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s += xv[i][j]; q += xv[i][j] * xv[i][j]; }
                    const int grp = cg * 8 / gs;
                    atomicAdd(&tab[grp * 2], s); atomicAdd(&tab[grp * 2 + 1], q);
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i) {
                const int cg = vl + i * TPV;
                const int grp = cg < CG ? cg * 8 / gs : 0;
                const float m = tab[grp * 2] / gs, var = fmaxf(tab[grp * 2 + 1] / gs - m * m, 0.f), r = rsqrtf(var + eps);
#pragma unroll
                for (int j = 0; j < 8; ++j) { mu[i][j] = m; rs[i][j] = r; }
            }
            __syncwarp();
        }
        if (!BWD) {
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i) {
                const int cg = vl + i * TPV;
                if (valid && cg < CG) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        o[j] = vis ? fmaf((xv[i][j] - mu[i][j]) * rs[i][j], ga[i][j], beta[cg * 8 + j]) : 0.f;
                    store8(out + voxel * g.C + cg * 8, o);
                }
            }
            continue;
        }
        // ---- backward: gg = dout·γ; dx = rstd·(gg − mean_grp(gg) − x̂·mean_grp(gg·x̂)) -------------------------------
        float gg[VN_MAXCH][8], m1[VN_MAXCH][8], m2[VN_MAXCH][8];
#pragma unroll
        for (int i = 0; i < VN_MAXCH; ++i) {
            const int cg = vl + i * TPV;
            float d[8];
            if (vis && cg < CG) load8(dout + voxel * g.C + cg * 8, d);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) d[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (xv[i][j] - mu[i][j]) * rs[i][j];
                xv[i][j] = xh;                                 // x̂ from here on
                ag[i][j] += d[j] * xh;
                ab[i][j] += d[j];
                gg[i][j] = d[j] * ga[i][j];
            }
        }
        if (mode == 0) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) { s += gg[i][j]; q += gg[i][j] * xv[i][j]; }
            for (int o = 1; o < TPV; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) { m1[i][j] = s / g.C; m2[i][j] = q / g.C; }
        } else if (mode == 1) {
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i)
#pragma unroll
                for (int j0 = 0; j0 < 8; ++j0) {
                    float s = 0.f, q = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j / gs == j0 / gs) { s += gg[i][j]; q += gg[i][j] * xv[i][j]; }
                    m1[i][j0] = s / gs; m2[i][j0] = q / gs;
                }
        } else {
            for (int k = vl; k < groups * 2; k += TPV) tab[k] = 0.f;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i) {
                const int cg = vl + i * TPV;
                if (cg < CG) {
                    float s = 0.f, q = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s += gg[i][j]; q += gg[i][j] * xv[i][j]; }
                    const int grp = cg * 8 / gs;
                    atomicAdd(&tab[grp * 2], s); atomicAdd(&tab[grp * 2 + 1], q);
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < VN_MAXCH; ++i) {
                const int cg = vl + i * TPV;
                const int grp = cg < CG ? cg * 8 / gs : 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { m1[i][j] = tab[grp * 2] / gs; m2[i][j] = tab[grp * 2 + 1] / gs; }
            }
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < VN_MAXCH; ++i) {
            const int cg = vl + i * TPV;
            if (valid && cg < CG) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = vis ? rs[i][j] * (gg[i][j] - m1[i][j] - xv[i][j] * m2[i][j]) : 0.f;
                store8(out + voxel * g.C + cg * 8, o);
            }
        }
    }
    if (BWD) {
#pragma unroll
        for (int i = 0; i < VN_MAXCH; ++i) {
            const int cg = vl + i * TPV;
            if (cg < CG) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    atomicAdd(&red[cg * 8 + j], ag[i][j]);
                    atomicAdd(&red[g.C + cg * 8 + j], ab[i][j]);
                }
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
            atomicAdd(&dgamma[c], red[c]);
            atomicAdd(&dbeta[c], red[g.C + c]);
        }
    }
}

static int voxel_norm_launch(const amb_geo* a, bool bwd, const void* x, const float* gamma, const float* beta, int groups,
                             float eps, const void* dout, void* out, float* dgamma, float* dbeta, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    AMB_CHECK(groups >= 1 && g.C % groups == 0, AMB_ERR_ARG, "voxel norm: C=%d not divisible by groups=%d", g.C, groups);
    const int gs = g.C / groups, CG = g.C / 8;
    AMB_CHECK(g.C <= VN_MAXCH * 256, AMB_ERR_ARG, "voxel norm: C=%d > %d", g.C, VN_MAXCH * 256);
    AMB_CHECK(groups == 1 || gs == 1 || gs == 2 || gs == 4 || gs % 8 == 0, AMB_ERR_ARG,
              "voxel norm: %d channels per group (need 1, 2, 4 or a multiple of 8)", gs);
    int TPV = 1;
    while (TPV < 32 && TPV < CG) TPV <<= 1;
    const int block = 256, vpb = block / TPV;
    const size_t smem = ((groups > 1 && gs > 8) ? (size_t)vpb * groups * 2 * sizeof(float) : 0) + (bwd ? 2 * (size_t)g.C * sizeof(float) : 0);
    AMB_CHECK(smem <= 48 * 1024, AMB_ERR_ARG, "voxel norm: %d groups need too much shared memory", groups);
    const long upper = (long)g.N * g.D * g.H * g.W;
    const int grid = grid_cap(upper, vpb, 4);
    cudaStream_t st = (cudaStream_t)stream;
#define AMB_VN(LI, BW)                                                                                                       \
    voxel_norm_kernel<LI, BW><<<grid, block, smem, st>>>(g, TPV, groups, eps, (const bf16*)x, gamma, beta, (const bf16*)dout, \
                                                         (bf16*)out, dgamma, dbeta)
    if (g.list) { if (bwd) AMB_VN(true, true); else AMB_VN(true, false); }
    else { if (bwd) AMB_VN(false, true); else AMB_VN(false, false); }
#undef AMB_VN
    AMB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// pooling (+ output mask)
// ------------------------------------------------------------------------------------------------------------
struct PoolP {
    int N, D, H, W, C, oD, oH, oW, k, s, p, mode, include_pad, divisor;
    int fd, fh, fw;
    const uint8_t* active;           // (N, fd, fh, fw) at the OUTPUT resolution, nullptr = no mask
};

__device__ __forceinline__ bool pool_out_active(const PoolP& P, int n, int z, int y, int xx) {
    if (!P.active) return true;
    const int pz = z / (P.oD / P.fd), py = y / (P.oH / P.fh), px = xx / (P.oW / P.fw);
    return P.active[((n * P.fd + pz) * P.fh + py) * P.fw + px] != 0;
}

__device__ __forceinline__ float pool_div(const PoolP& P, int z0, int y0, int x0) {
    if (P.divisor > 0) return (float)P.divisor;
    // torch: window clipped to the padded extent first, then (count_include_pad ? that : the in-bounds part)
    const int z1 = min(z0 + P.k, P.D + P.p), y1 = min(y0 + P.k, P.H + P.p), x1 = min(x0 + P.k, P.W + P.p);
    if (P.include_pad) return (float)((z1 - z0) * (y1 - y0) * (x1 - x0));
    const int a = min(z1, P.D) - max(z0, 0), b = min(y1, P.H) - max(y0, 0), c = min(x1, P.W) - max(x0, 0);
    return (float)(a * b * c);
}

__global__ void pool_fwd_kernel(PoolP P, const bf16* __restrict__ x, bf16* __restrict__ y) {
    const int CG = P.C / 8;
    const long total = (long)P.N * P.oD * P.oH * P.oW * CG;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % CG);
        long t = i / CG;
        const int ox = (int)(t % P.oW); t /= P.oW;
        const int oy = (int)(t % P.oH); t /= P.oH;
        const int oz = (int)(t % P.oD);
        const int n = (int)(t / P.oD);
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
        if (pool_out_active(P, n, oz, oy, ox)) {
            const int z0 = oz * P.s - P.p, y0 = oy * P.s - P.p, x0 = ox * P.s - P.p;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = P.mode == 0 ? -INFINITY : 0.f;
            for (int dz = 0; dz < P.k; ++dz) {
                const int z = z0 + dz;
                if (z < 0 || z >= P.D) continue;
                for (int dy = 0; dy < P.k; ++dy) {
                    const int yy = y0 + dy;
                    if (yy < 0 || yy >= P.H) continue;
                    for (int dx = 0; dx < P.k; ++dx) {
                        const int xx = x0 + dx;
                        if (xx < 0 || xx >= P.W) continue;
                        float f[8];
                        load8(x + ((((long)n * P.D + z) * P.H + yy) * P.W + xx) * P.C + cg * 8, f);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] = P.mode == 0 ? fmaxf(acc[j], f[j]) : acc[j] + f[j];
                    }
                }
            }
            const float inv = P.mode == 0 ? 1.f : 1.f / pool_div(P, z0, y0, x0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = acc[j] * inv;
        }
        store8(y + ((((long)n * P.oD + oz) * P.oH + oy) * P.oW + ox) * P.C + cg * 8, o);
    }
}

// gather form: an input voxel collects from every output window that contains it (max: only where it is the window's
// first maximum in (z, y, x) scan order — torch's max_pool3d backward)
__global__ void pool_bwd_kernel(PoolP P, const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx) {
    const int CG = P.C / 8;
    const long total = (long)P.N * P.D * P.H * P.W * CG;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % CG);
        long t = i / CG;
        const int ix = (int)(t % P.W); t /= P.W;
        const int iy = (int)(t % P.H); t /= P.H;
        const int iz = (int)(t % P.D);
        const int n = (int)(t / P.D);
        float self[8], acc[8];
        load8(x + i * 8, self);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        // outputs o with o·s − p <= i < o·s − p + k
        const int oz_lo = max(0, (iz + P.p - P.k + P.s) / P.s), oz_hi = min(P.oD - 1, (iz + P.p) / P.s);
        const int oy_lo = max(0, (iy + P.p - P.k + P.s) / P.s), oy_hi = min(P.oH - 1, (iy + P.p) / P.s);
        const int ox_lo = max(0, (ix + P.p - P.k + P.s) / P.s), ox_hi = min(P.oW - 1, (ix + P.p) / P.s);
        for (int oz = oz_lo; oz <= oz_hi; ++oz)
            for (int oy = oy_lo; oy <= oy_hi; ++oy)
                for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                    if (!pool_out_active(P, n, oz, oy, ox)) continue;
                    float d[8];
                    load8(dy + ((((long)n * P.oD + oz) * P.oH + oy) * P.oW + ox) * P.C + cg * 8, d);
                    const int z0 = oz * P.s - P.p, y0 = oy * P.s - P.p, x0 = ox * P.s - P.p;
                    if (P.mode != 0) {
                        const float inv = 1.f / pool_div(P, z0, y0, x0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] += d[j] * inv;
                        continue;
                    }
                    // am I the first maximum of this window?  (an earlier element >= me, or a later one > me, beats me)
                    bool win[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) win[j] = true;
                    for (int dz = 0; dz < P.k; ++dz) {
                        const int z = z0 + dz;
                        if (z < 0 || z >= P.D) continue;
                        for (int dyy = 0; dyy < P.k; ++dyy) {
                            const int yy = y0 + dyy;
                            if (yy < 0 || yy >= P.H) continue;
                            for (int dxx = 0; dxx < P.k; ++dxx) {
                                const int xx = x0 + dxx;
                                if (xx < 0 || xx >= P.W || (z == iz && yy == iy && xx == ix)) continue;
                                const bool earlier = z < iz || (z == iz && (yy < iy || (yy == iy && xx < ix)));
                                float f[8];
                                load8(x + ((((long)n * P.D + z) * P.H + yy) * P.W + xx) * P.C + cg * 8, f);
#pragma unroll
                                for (int j = 0; j < 8; ++j) win[j] = win[j] && (earlier ? f[j] < self[j] : f[j] <= self[j]);
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += win[j] ? d[j] : 0.f;
                }
        store8(dx + i * 8, acc);
    }
}

static int make_pool(PoolP& P, int N, int D, int H, int W, int C, int k, int s, int p, int mode, int include_pad,
                     int divisor, const uint8_t* active, int fd, int fh, int fw) {
    AMB_CHECK(C % 8 == 0 && k >= 1 && s >= 1 && p >= 0 && 2 * p <= k, AMB_ERR_ARG, "pool: bad C/k/s/p %d/%d/%d/%d", C, k, s, p);
    P.N = N; P.D = D; P.H = H; P.W = W; P.C = C; P.k = k; P.s = s; P.p = p; P.mode = mode; P.include_pad = include_pad;
    P.divisor = divisor;
    P.oD = (D + 2 * p - k) / s + 1; P.oH = (H + 2 * p - k) / s + 1; P.oW = (W + 2 * p - k) / s + 1;
    AMB_CHECK(P.oD > 0 && P.oH > 0 && P.oW > 0, AMB_ERR_ARG, "pool: empty output");
    P.active = active; P.fd = fd; P.fh = fh; P.fw = fw;
    if (active)
        AMB_CHECK(fd > 0 && P.oD % fd == 0 && P.oH % fh == 0 && P.oW % fw == 0, AMB_ERR_ARG,
                  "pool: output (%d,%d,%d) not divisible by mask grid (%d,%d,%d)", P.oD, P.oH, P.oW, fd, fh, fw);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// masked global average
// ------------------------------------------------------------------------------------------------------------
__global__ void masked_mean_fwd_kernel(Geo g, const bf16* __restrict__ x, float* __restrict__ sums) {
    // grid.y = sample; thread keeps its channel chunk, walks the sample's voxels
    const int CG = g.C / 8, cg = threadIdx.x % CG, n = blockIdx.y;
    const long S = (long)g.D * g.H * g.W;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    for (long v = ((long)blockIdx.x * blockDim.x + threadIdx.x) / CG; v < S; v += ((long)gridDim.x * blockDim.x) / CG) {
        const long voxel = n * S + v;
        if (!voxel_active(g, voxel)) continue;
        float f[8];
        load8(x + voxel * g.C + cg * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sums[(long)n * g.C + cg * 8 + j], a[j]);
}

__global__ void masked_mean_finish_kernel(Geo g, float* __restrict__ sums) {
    // Σ → Σ / (visible voxels of the sample + 1e-6)
    const int n = blockIdx.x;
    const int L = g.fd * g.fh * g.fw;
    int cnt = 0;
    for (int l = 0; l < L; ++l) cnt += g.active[n * L + l] != 0;
    const float inv = 1.f / ((float)cnt * (float)(1 << (3 * g.lgP)) + 1e-6f);
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) sums[(long)n * g.C + c] *= inv;
}

__global__ void masked_mean_bwd_kernel(Geo g, const float* __restrict__ dmean, bf16* __restrict__ dx) {
    const int CG = g.C / 8;
    const long S = (long)g.D * g.H * g.W, total = (long)g.N * S * CG;
    const int L = g.fd * g.fh * g.fw;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % CG);
        const long voxel = i / CG;
        const int n = (int)(voxel / S);
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
        if (voxel_active(g, voxel)) {
            int cnt = 0;
            for (int l = 0; l < L; ++l) cnt += g.active[n * L + l] != 0;
            const float inv = 1.f / ((float)cnt * (float)(1 << (3 * g.lgP)) + 1e-6f);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = dmean[(long)n * g.C + cg * 8 + j] * inv;
        }
        store8(dx + i * 8, o);
    }
}

// ------------------------------------------------------------------------------------------------------------
// depthwise convolution
// ------------------------------------------------------------------------------------------------------------
struct DwP {
    int N, D, H, W, C, oD, oH, oW, k, s;
    int fd, fh, fw;                  // mask grid at the OUTPUT resolution
    const uint8_t* active;           // nullptr = dense
    int flip;                        // 1: taps reversed (stride-1 input gradient = the same convolution with a mirrored kernel)
    const uint8_t* in_active;        // mask applied to the INPUT tensor while it is staged (input gradient: dy counts on visible
                                     // outputs only — the backward of the forward pass's mask multiply), grid (fd, fh, fw) too
};

__device__ __forceinline__ bool dw_in_active(const DwP& P, int n, int z, int y, int xx) {
    if (!P.in_active) return true;
    const int pz = z / (P.D / P.fd), py = y / (P.H / P.fh), px = xx / (P.W / P.fw);
    return P.in_active[((n * P.fd + pz) * P.fh + py) * P.fw + px] != 0;
}

__device__ __forceinline__ bool dw_out_active(const DwP& P, int n, int z, int y, int xx) {
    if (!P.active) return true;
    const int pz = z / (P.oD / P.fd), py = y / (P.oH / P.fh), px = xx / (P.oW / P.fw);
    return P.active[((n * P.fd + pz) * P.fh + py) * P.fw + px] != 0;
}

// One CTA = an output tile TZ × TY × 16 (x) for one chunk of 8 channels; a thread = 4 consecutive x outputs.  The input
// halo tile and the chunk's weights ([tap][8] fp32) sit in shared memory; per (dz, dy) a thread reads 3·S + K inputs and
// K weight vectors for 4·K·8 FMAs.
template <int K, int S, int TZ, int TY>
__global__ void __launch_bounds__(TZ* TY * 4)
dwconv_fwd_kernel(DwP P, const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  bf16* __restrict__ y) {
    constexpr int TX = 16, IZ = (TZ - 1) * S + K, IY = (TY - 1) * S + K, IX = (TX - 1) * S + K, pad = K / 2;
    extern __shared__ __align__(16) unsigned char dsm[];
    uint4* tile = reinterpret_cast<uint4*>(dsm);                           // [IZ][IY][IX] × 8 bf16
    float* wsm = reinterpret_cast<float*>(dsm + (size_t)IZ * IY * IX * 16);   // [K³][8]
    const int cg = blockIdx.y;
    const int tilesX = (P.oW + TX - 1) / TX, tilesY = (P.oH + TY - 1) / TY, tilesZ = (P.oD + TZ - 1) / TZ;
    int t = blockIdx.x;
    const int tx = t % tilesX; t /= tilesX;
    const int ty = t % tilesY; t /= tilesY;
    const int tz = t % tilesZ;
    const int n = t / tilesZ;
    const int oz0 = tz * TZ, oy0 = ty * TY, ox0 = tx * TX;
    const int nthr = TZ * TY * 4;
    // any visible output in this tile?
    bool any = P.active == nullptr;
    if (!any) {
        const int pz = P.oD / P.fd, py = P.oH / P.fh, px = P.oW / P.fw;
        for (int a = oz0 / pz; a <= min(oz0 + TZ - 1, P.oD - 1) / pz && !any; ++a)
            for (int b = oy0 / py; b <= min(oy0 + TY - 1, P.oH - 1) / py && !any; ++b)
                for (int c = ox0 / px; c <= min(ox0 + TX - 1, P.oW - 1) / px && !any; ++c)
                    any = P.active[((n * P.fd + a) * P.fh + b) * P.fw + c] != 0;
    }
    const int lz = threadIdx.x / (TY * 4), ly = (threadIdx.x / 4) % TY, lx = (threadIdx.x % 4) * 4;
    const int oz = oz0 + lz, oy = oy0 + ly;
    float acc[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
    if (any) {
        const int iz0 = oz0 * S - pad, iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;
        for (int i = threadIdx.x; i < IZ * IY * IX; i += nthr) {
            const int xx = i % IX, yy = (i / IX) % IY, zz = i / (IX * IY);
            const int gz = iz0 + zz, gy = iy0 + yy, gx = ix0 + xx;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gz >= 0 && gz < P.D && gy >= 0 && gy < P.H && gx >= 0 && gx < P.W && dw_in_active(P, n, gz, gy, gx))
                v = __ldg(reinterpret_cast<const uint4*>(x + ((((long)n * P.D + gz) * P.H + gy) * P.W + gx) * P.C + cg * 8));
            tile[i] = v;
        }
        for (int i = threadIdx.x; i < K * K * K * 8; i += nthr) {
            const int j = i % 8, tap = i / 8;
            wsm[i] = w[(long)(cg * 8 + j) * K * K * K + (P.flip ? K * K * K - 1 - tap : tap)];
        }
        __syncthreads();
        for (int dz = 0; dz < K; ++dz)
            for (int dy = 0; dy < K; ++dy) {
                float wr[K][8];
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float4 a = *reinterpret_cast<const float4*>(wsm + ((dz * K + dy) * K + kx) * 8);
                    const float4 b = *reinterpret_cast<const float4*>(wsm + ((dz * K + dy) * K + kx) * 8 + 4);
                    wr[kx][0] = a.x; wr[kx][1] = a.y; wr[kx][2] = a.z; wr[kx][3] = a.w;
                    wr[kx][4] = b.x; wr[kx][5] = b.y; wr[kx][6] = b.z; wr[kx][7] = b.w;
                }
                const uint4* row = tile + ((size_t)(lz * S + dz) * IY + (ly * S + dy)) * IX + lx * S;
#pragma unroll
                for (int xi = 0; xi < 3 * S + K; ++xi) {
                    float f[8];
                    unpack_u4(row[xi], f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int kx = xi - q * S;
                        if (kx >= 0 && kx < K) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(f[j], wr[kx][j], acc[q][j]);
                        }
                    }
                }
            }
    }
    if (oz < P.oD && oy < P.oH) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ox = ox0 + lx + q;
            if (ox >= P.oW) continue;
            float o[8];
            const bool vis = any && dw_out_active(P, n, oz, oy, ox);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = vis ? acc[q][j] + (bias ? bias[cg * 8 + j] : 0.f) : 0.f;
            store8(y + ((((long)n * P.oD + oz) * P.oH + oy) * P.oW + ox) * P.C + cg * 8, o);
        }
    }
}

// stride-2 input gradient: dx[u] = Σ_t dy[(u + pad − t) / 2]·w[t] over the taps with matching parity
__global__ void dwconv_dgrad_s2_kernel(DwP P, const bf16* __restrict__ dy, const float* __restrict__ w, bf16* __restrict__ dx) {
    const int CG = P.C / 8, K = P.k, pad = K / 2;
    const long total = (long)P.N * P.D * P.H * P.W * CG;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % CG);
        long t = i / CG;
        const int ix = (int)(t % P.W); t /= P.W;
        const int iy = (int)(t % P.H); t /= P.H;
        const int iz = (int)(t % P.D);
        const int n = (int)(t / P.D);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int tz = (iz + pad) & 1; tz < K; tz += 2) {
            const int oz = (iz + pad - tz) / 2;
            if (iz + pad - tz < 0 || oz >= P.oD) continue;
            for (int ty = (iy + pad) & 1; ty < K; ty += 2) {
                const int oy = (iy + pad - ty) / 2;
                if (iy + pad - ty < 0 || oy >= P.oH) continue;
                for (int tx = (ix + pad) & 1; tx < K; tx += 2) {
                    const int ox = (ix + pad - tx) / 2;
                    if (ix + pad - tx < 0 || ox >= P.oW) continue;
                    if (!dw_out_active(P, n, oz, oy, ox)) continue;
                    float d[8];
                    load8(dy + ((((long)n * P.oD + oz) * P.oH + oy) * P.oW + ox) * P.C + cg * 8, d);
                    const int tap = (tz * K + ty) * K + tx;
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(d[j], w[(long)(cg * 8 + j) * K * K * K + tap], acc[j]);
                }
            }
        }
        store8(dx + i * 8, acc);
    }
}

// weight gradient: dw[c][t] += Σ_o dy[o, c]·x[o·S + t − pad, c].  One CTA = an output tile 4 × 8 × 16 and 8 channels, staged
// like the forward pass; a thread owns taps tid, tid + 128, ... and walks the tile's 512 outputs.
template <int K, int S>
__global__ void __launch_bounds__(128)
dwconv_wgrad_kernel(DwP P, const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw) {
    constexpr int TZ = 4, TY = 8, TX = 16, IZ = (TZ - 1) * S + K, IY = (TY - 1) * S + K, IX = (TX - 1) * S + K, pad = K / 2;
    extern __shared__ __align__(16) unsigned char dsm[];
    uint4* tile = reinterpret_cast<uint4*>(dsm);
    uint4* dtile = tile + (size_t)IZ * IY * IX;                            // [TZ][TY][TX]
    const int cg = blockIdx.y;
    const int tilesX = (P.oW + TX - 1) / TX, tilesY = (P.oH + TY - 1) / TY, tilesZ = (P.oD + TZ - 1) / TZ;
    int t = blockIdx.x;
    const int tx = t % tilesX; t /= tilesX;
    const int ty = t % tilesY; t /= tilesY;
    const int tz = t % tilesZ;
    const int n = t / tilesZ;
    const int oz0 = tz * TZ, oy0 = ty * TY, ox0 = tx * TX;
    bool any = P.active == nullptr;
    if (!any) {
        const int pz = P.oD / P.fd, py = P.oH / P.fh, px = P.oW / P.fw;
        for (int a = oz0 / pz; a <= min(oz0 + TZ - 1, P.oD - 1) / pz && !any; ++a)
            for (int b = oy0 / py; b <= min(oy0 + TY - 1, P.oH - 1) / py && !any; ++b)
                for (int c = ox0 / px; c <= min(ox0 + TX - 1, P.oW - 1) / px && !any; ++c)
                    any = P.active[((n * P.fd + a) * P.fh + b) * P.fw + c] != 0;
    }
    if (!any) return;                                                   // dy is zero on masked outputs
    const int iz0 = oz0 * S - pad, iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;
    for (int i = threadIdx.x; i < IZ * IY * IX; i += 128) {
        const int xx = i % IX, yy = (i / IX) % IY, zz = i / (IX * IY);
        const int gz = iz0 + zz, gy = iy0 + yy, gx = ix0 + xx;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gz >= 0 && gz < P.D && gy >= 0 && gy < P.H && gx >= 0 && gx < P.W)
            v = __ldg(reinterpret_cast<const uint4*>(x + ((((long)n * P.D + gz) * P.H + gy) * P.W + gx) * P.C + cg * 8));
        tile[i] = v;
    }
    for (int i = threadIdx.x; i < TZ * TY * TX; i += 128) {
        const int xx = i % TX, yy = (i / TX) % TY, zz = i / (TX * TY);
        const int gz = oz0 + zz, gy = oy0 + yy, gx = ox0 + xx;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gz < P.oD && gy < P.oH && gx < P.oW && dw_out_active(P, n, gz, gy, gx))
            v = __ldg(reinterpret_cast<const uint4*>(dy + ((((long)n * P.oD + gz) * P.oH + gy) * P.oW + gx) * P.C + cg * 8));
        dtile[i] = v;
    }
    __syncthreads();
    for (int tap = threadIdx.x; tap < K * K * K; tap += 128) {
        const int kx = tap % K, ky = (tap / K) % K, kz = tap / (K * K);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int zz = 0; zz < TZ; ++zz)
            for (int yy = 0; yy < TY; ++yy) {
                const uint4* xr = tile + ((size_t)(zz * S + kz) * IY + (yy * S + ky)) * IX + kx;
                const uint4* dr = dtile + (zz * TY + yy) * TX;
#pragma unroll 4
                for (int xx = 0; xx < TX; ++xx) {
                    float f[8], d[8];
                    unpack_u4(xr[xx * S], f);
                    unpack_u4(dr[xx], d);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], d[j], acc[j]);
                }
            }
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&dw[(long)(cg * 8 + j) * K * K * K + tap], acc[j]);
    }
}

static int make_dw(DwP& P, int N, int D, int H, int W, int C, int k, int s, const uint8_t* active, int fd, int fh, int fw) {
    AMB_CHECK(C % 8 == 0, AMB_ERR_ARG, "depthwise conv: C=%d must be a multiple of 8", C);
    AMB_CHECK((k == 3 || k == 5 || k == 7) && (s == 1 || s == 2), AMB_ERR_UNSUPPORTED, "depthwise conv: k=%d s=%d (k in 3/5/7, s in 1/2)", k, s);
    AMB_CHECK(s == 1 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), AMB_ERR_ARG, "depthwise conv: stride 2 needs even extents");
    P.N = N; P.D = D; P.H = H; P.W = W; P.C = C; P.k = k; P.s = s;
    P.oD = D / s; P.oH = H / s; P.oW = W / s;
    P.active = active; P.fd = fd; P.fh = fh; P.fw = fw; P.flip = 0; P.in_active = nullptr;
    if (active)
        AMB_CHECK(fd > 0 && P.oD % fd == 0 && P.oH % fh == 0 && P.oW % fw == 0, AMB_ERR_ARG,
                  "depthwise conv: output (%d,%d,%d) not divisible by mask grid (%d,%d,%d)", P.oD, P.oH, P.oW, fd, fh, fw);
    return 0;
}

template <int K, int S, int TZ, int TY>
static int dw_fwd_launch(const DwP& P, const void* x, const float* w, const float* bias, void* y, cudaStream_t st) {
    constexpr int IZ = (TZ - 1) * S + K, IY = (TY - 1) * S + K, IX = 15 * S + K;
    const size_t smem = (size_t)IZ * IY * IX * 16 + (size_t)K * K * K * 8 * 4;
    auto kern = dwconv_fwd_kernel<K, S, TZ, TY>;
    AMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ((P.oW + 15) / 16) * ((P.oH + TY - 1) / TY) * ((P.oD + TZ - 1) / TZ) * P.N;
    kern<<<dim3(tiles, P.C / 8), TZ * TY * 4, smem, st>>>(P, (const bf16*)x, w, bias, (bf16*)y);
    AMB_LAUNCH_CHECK();
    return 0;
}

template <int K, int S>
static int dw_wgrad_launch(const DwP& P, const void* x, const void* dy, float* dw, cudaStream_t st) {
    constexpr int IZ = 3 * S + K, IY = 7 * S + K, IX = 15 * S + K;
    const size_t smem = ((size_t)IZ * IY * IX + 4 * 8 * 16) * 16;
    auto kern = dwconv_wgrad_kernel<K, S>;
    AMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ((P.oW + 15) / 16) * ((P.oH + 7) / 8) * ((P.oD + 3) / 4) * P.N;
    kern<<<dim3(tiles, P.C / 8), 128, smem, st>>>(P, (const bf16*)x, (const bf16*)dy, dw);
    AMB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// GELU, layer scale
// ------------------------------------------------------------------------------------------------------------
__global__ void gelu_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dout, bf16* __restrict__ out, long n8) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        float f[8], d[8], o[8];
        load8(x + i * 8, f);
        if (dout) load8(dout + i * 8, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float cdf = 0.5f * (1.f + erff(f[j] * 0.70710678118654752f));
            if (dout) o[j] = d[j] * (cdf + f[j] * 0.3989422804014327f * __expf(-0.5f * f[j] * f[j]));
            else o[j] = f[j] * cdf;
        }
        store8(out + i * 8, o);
    }
}

// forward: out = inp + m·γ_c·x       backward (dout given): dxb = m·γ_c·dout, dgamma_c += Σ m·dout·x   (dinp = dout)
__global__ void layer_scale_kernel(Geo g, const bf16* __restrict__ inp, const bf16* __restrict__ x, const float* __restrict__ gamma,
                                   const bf16* __restrict__ dout, bf16* __restrict__ out, float* __restrict__ dgamma) {
    const int CG = g.C / 8, cg = threadIdx.x % CG;
    const long nvox = (long)g.N * g.D * g.H * g.W;
    float ga[8], ag[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = gamma ? gamma[cg * 8 + j] : 1.f; ag[j] = 0.f; }
    for (long v = ((long)blockIdx.x * blockDim.x + threadIdx.x) / CG; v < nvox; v += ((long)gridDim.x * blockDim.x) / CG) {
        const bool vis = g.active == nullptr || voxel_active(g, v);
        const long off = v * g.C + cg * 8;
        float o[8];
        if (dout) {
            float d[8], f[8];
            load8(dout + off, d);
            if (vis) load8(x + off, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o[j] = vis ? d[j] * ga[j] : 0.f;
                if (vis) ag[j] += d[j] * f[j];
            }
        } else {
            float a[8], f[8];
            load8(inp + off, a);
            if (vis) load8(x + off, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = vis ? fmaf(ga[j], f[j], a[j]) : a[j];
        }
        store8(out + off, o);
    }
    if (dout && dgamma) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&dgamma[cg * 8 + j], ag[j]);
    }
}

}  // namespace amb

using namespace amb;

extern "C" int amb_voxel_norm_fwd(const amb_geo* g, const void* x, const float* gamma, const float* beta, int groups,
                                  float eps, void* out, void* stream) {
    return voxel_norm_launch(g, false, x, gamma, beta, groups, eps, nullptr, out, nullptr, nullptr, stream);
}

extern "C" int amb_voxel_norm_bwd(const amb_geo* g, const void* dout, const void* x, const float* gamma, int groups,
                                  float eps, void* dx, float* dgamma, float* dbeta, void* stream) {
    AMB_CHECK(dgamma && dbeta, AMB_ERR_ARG, "voxel norm backward needs dgamma and dbeta (fp32[C], accumulated into)");
    return voxel_norm_launch(g, true, x, gamma, nullptr, groups, eps, dout, dx, dgamma, dbeta, stream);
}

extern "C" int amb_pool3d_fwd(const void* x, void* y, int N, int D, int H, int W, int C, int k, int stride, int pad, int mode,
                              int count_include_pad, int divisor_override, const uint8_t* active, int fd, int fh, int fw,
                              void* stream) {
    PoolP P;
    if (int e = make_pool(P, N, D, H, W, C, k, stride, pad, mode, count_include_pad, divisor_override, active, fd, fh, fw)) return e;
    const long items = (long)N * P.oD * P.oH * P.oW * (C / 8);
    pool_fwd_kernel<<<grid_cap(items, 256), 256, 0, (cudaStream_t)stream>>>(P, (const bf16*)x, (bf16*)y);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_pool3d_bwd(const void* x, const void* dy, void* dx, int N, int D, int H, int W, int C, int k, int stride,
                              int pad, int mode, int count_include_pad, int divisor_override, const uint8_t* active, int fd,
                              int fh, int fw, void* stream) {
    PoolP P;
    if (int e = make_pool(P, N, D, H, W, C, k, stride, pad, mode, count_include_pad, divisor_override, active, fd, fh, fw)) return e;
    const long items = (long)N * D * H * W * (C / 8);
    pool_bwd_kernel<<<grid_cap(items, 256), 256, 0, (cudaStream_t)stream>>>(P, (const bf16*)x, (const bf16*)dy, (bf16*)dx);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_masked_mean_fwd(const amb_geo* a, const void* x, float* mean, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    AMB_CHECK(g.active != nullptr, AMB_ERR_ARG, "masked mean needs the active mask");
    const int CG = g.C / 8;
    AMB_CHECK(CG <= 256, AMB_ERR_ARG, "masked mean: C=%d > 2048", g.C);
    const int block = (256 / CG) * CG;
    AMB_CUDA(cudaMemsetAsync(mean, 0, (size_t)g.N * g.C * sizeof(float), (cudaStream_t)stream));
    const long items = (long)g.D * g.H * g.W * CG;
    masked_mean_fwd_kernel<<<dim3(grid_cap(items, block, 2), g.N), block, 0, (cudaStream_t)stream>>>(g, (const bf16*)x, mean);
    AMB_LAUNCH_CHECK();
    masked_mean_finish_kernel<<<g.N, 128, 0, (cudaStream_t)stream>>>(g, mean);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_masked_mean_bwd(const amb_geo* a, const float* dmean, void* dx, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    AMB_CHECK(g.active != nullptr, AMB_ERR_ARG, "masked mean needs the active mask");
    const long items = (long)g.N * g.D * g.H * g.W * (g.C / 8);
    masked_mean_bwd_kernel<<<grid_cap(items, 256), 256, 0, (cudaStream_t)stream>>>(g, dmean, (bf16*)dx);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_dwconv3d(int op, const void* x, const float* w, const float* bias, void* y, int N, int D, int H, int W,
                            int C, int k, int stride, const uint8_t* active, int fd, int fh, int fw, void* stream) {
    // op AMB_OP_CONV: x (N,D,H,W,C) → y (N,D/s,H/s,W/s,C) = (dwconv(x) + bias)·mask;  op AMB_OP_CONV_DGRAD: x = dy at the output
    // resolution → y = dx (N,D,H,W,C) with dy read as dy·mask (the backward of the mask multiply); w = the layer's fp32 weight
    // (C,1,k,k,k) in both cases; `active` is the mask at the OUTPUT resolution of the layer in both cases
    DwP P;
    cudaStream_t st = (cudaStream_t)stream;
    AMB_CHECK(op == AMB_OP_CONV || op == AMB_OP_CONV_DGRAD, AMB_ERR_ARG, "depthwise conv: op %d", op);
    if (op == AMB_OP_CONV_DGRAD && stride == 2) {
        if (int e = make_dw(P, N, D, H, W, C, k, 2, active, fd, fh, fw)) return e;
        const long items = (long)N * D * H * W * (C / 8);
        dwconv_dgrad_s2_kernel<<<grid_cap(items, 256), 256, 0, st>>>(P, (const bf16*)x, w, (bf16*)y);
        AMB_LAUNCH_CHECK();
        return 0;
    }
    if (op == AMB_OP_CONV_DGRAD) {
        if (int e = make_dw(P, N, D, H, W, C, k, 1, active, fd, fh, fw)) return e;
        P.flip = 1;
        P.in_active = active;          // dy counts on visible outputs only; dx itself is not masked
        P.active = nullptr;
        bias = nullptr;
    } else if (int e = make_dw(P, N, D, H, W, C, k, stride, active, fd, fh, fw)) return e;
    if (P.s == 1) {
        if (k == 3) return dw_fwd_launch<3, 1, 4, 8>(P, x, w, bias, y, st);
        if (k == 5) return dw_fwd_launch<5, 1, 4, 8>(P, x, w, bias, y, st);
        return dw_fwd_launch<7, 1, 4, 8>(P, x, w, bias, y, st);
    }
    if (k == 3) return dw_fwd_launch<3, 2, 4, 4>(P, x, w, bias, y, st);
    if (k == 5) return dw_fwd_launch<5, 2, 4, 4>(P, x, w, bias, y, st);
    return dw_fwd_launch<7, 2, 4, 4>(P, x, w, bias, y, st);
}

extern "C" int amb_dwconv3d_wgrad(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int C, int k, int stride,
                                  const uint8_t* active, int fd, int fh, int fw, void* stream) {
    DwP P;
    if (int e = make_dw(P, N, D, H, W, C, k, stride, active, fd, fh, fw)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 1) {
        if (k == 3) return dw_wgrad_launch<3, 1>(P, x, dy, dw, st);
        if (k == 5) return dw_wgrad_launch<5, 1>(P, x, dy, dw, st);
        return dw_wgrad_launch<7, 1>(P, x, dy, dw, st);
    }
    if (k == 3) return dw_wgrad_launch<3, 2>(P, x, dy, dw, st);
    if (k == 5) return dw_wgrad_launch<5, 2>(P, x, dy, dw, st);
    return dw_wgrad_launch<7, 2>(P, x, dy, dw, st);
}

extern "C" int amb_gelu(const void* x, const void* dout, void* out, long n, void* stream) {
    AMB_CHECK(n % 8 == 0, AMB_ERR_ARG, "amb_gelu: n must be a multiple of 8");
    gelu_kernel<<<grid_cap(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dout, (bf16*)out, n / 8);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_layer_scale(const amb_geo* a, const void* inp, const void* x, const float* gamma, const void* dout, void* out,
                               float* dgamma, void* stream) {
    Geo g;
    if (int e = make_geo(a, g)) return e;
    const int CG = g.C / 8;
    AMB_CHECK(CG <= 256, AMB_ERR_ARG, "layer scale: C=%d > 2048", g.C);
    const int block = (256 / CG) * CG;
    const long items = (long)g.N * g.D * g.H * g.W * CG;
    layer_scale_kernel<<<grid_cap(items, block, 4), block, 0, (cudaStream_t)stream>>>(g, (const bf16*)inp, (const bf16*)x, gamma,
                                                                                     (const bf16*)dout, (bf16*)out, dgamma);
    AMB_LAUNCH_CHECK();
    return 0;
}
