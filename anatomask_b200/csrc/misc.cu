// Patch loss, hard-mask generation, EMA / AdamW over the flat parameter arena, weight packing, the Cin=1 stem and
// the Cout=1 reconstruction head.  All HBM-bound: 128-bit accesses, warp-shuffle / shared-memory reductions.
#include <stdarg.h>

#include "common.cuh"

namespace amb {

static thread_local char g_err[512] = "";
int g_launch_count = 0;
const char* g_last_conv_kernel = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------------------
// patch loss.  One CTA (256 threads) per 16³ patch; thread t owns the 16-float row (z = t/16, y = t%16).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i];
    return t;
}

__global__ void __launch_bounds__(256)
patch_loss_fwd_kernel(const float* __restrict__ inp, const float* __restrict__ rec, const uint8_t* __restrict__ active,
                      int N, int D, int H, int W, int normalize, float* __restrict__ per_patch,
                      float* __restrict__ loss, float* __restrict__ pstats, unsigned int* ticket) {
    __shared__ float sh[8];
    __shared__ bool is_last;
    const int fd = D / 16, fh = H / 16, fw = W / 16, L = fd * fh * fw;
    const int pid = blockIdx.x, n = pid / L, l = pid % L;
    const int pz = l / (fh * fw), py = (l / fw) % fh, px = l % fw;
    const bool masked = active[pid] == 0;
    float l2 = 0.f, mean = 0.f, rstd = 1.f;
    if (masked) {   // uniform per block
        const int z = threadIdx.x >> 4, y = threadIdx.x & 15;
        const long off = (((long)n * D + pz * 16 + z) * H + py * 16 + y) * W + px * 16;
        float t[16], r[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 a = *reinterpret_cast<const float4*>(inp + off + j * 4);
            float4 b = *reinterpret_cast<const float4*>(rec + off + j * 4);
            t[j * 4] = a.x; t[j * 4 + 1] = a.y; t[j * 4 + 2] = a.z; t[j * 4 + 3] = a.w;
            r[j * 4] = b.x; r[j * 4 + 1] = b.y; r[j * 4 + 2] = b.z; r[j * 4 + 3] = b.w;
        }
        if (normalize) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) s += t[j];
            mean = block_sum_256(s, sh) * (1.f / 4096.f);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) q += (t[j] - mean) * (t[j] - mean);
            float var = block_sum_256(q, sh) * (1.f / 4095.f);          // unbiased (torch.var default)
            rstd = 1.f / sqrtf(var + 1e-6f);
        }
        float e = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float d = r[j] - (t[j] - mean) * rstd;
            e += d * d;
        }
        l2 = block_sum_256(e, sh) * (1.f / 4096.f);
    }
    if (threadIdx.x == 0) {
        per_patch[pid] = l2;
        if (pstats) { pstats[2 * pid] = mean; pstats[2 * pid + 1] = rstd; }
        __threadfence();
        unsigned int done = atomicAdd(ticket, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {      // deterministic final reduction by the last CTA
        __threadfence();
        const int total = N * L;
        float s = 0.f, c = 0.f;
        for (int i = threadIdx.x; i < total; i += 256) {
            s += __ldcg(per_patch + i);
            c += active[i] == 0 ? 1.f : 0.f;
        }
        s = block_sum_256(s, sh);
        c = block_sum_256(c, sh);
        if (threadIdx.x == 0) {
            float denom = c + 1e-8f;
            if (loss) loss[0] = s / denom;
            if (pstats) pstats[2 * total] = denom;
            *ticket = 0;
        }
    }
}

__global__ void __launch_bounds__(256)
patch_loss_bwd_kernel(const float* __restrict__ inp, const float* __restrict__ rec, const uint8_t* __restrict__ active,
                      const float* __restrict__ pstats, const float* __restrict__ dloss, int N, int D, int H, int W,
                      float* __restrict__ drec) {
    const int fd = D / 16, fh = H / 16, fw = W / 16, L = fd * fh * fw;
    const int pid = blockIdx.x, n = pid / L, l = pid % L;
    const int pz = l / (fh * fw), py = (l / fw) % fh, px = l % fw;
    const int z = threadIdx.x >> 4, y = threadIdx.x & 15;
    const long off = (((long)n * D + pz * 16 + z) * H + py * 16 + y) * W + px * 16;
    if (active[pid] != 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(drec + off + j * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float mean = pstats[2 * pid], rstd = pstats[2 * pid + 1];
    const float k = dloss[0] * 2.f / (4096.f * pstats[2 * N * L]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 a = *reinterpret_cast<const float4*>(inp + off + j * 4);
        float4 b = *reinterpret_cast<const float4*>(rec + off + j * 4);
        float4 o;
        o.x = k * (b.x - (a.x - mean) * rstd);
        o.y = k * (b.y - (a.y - mean) * rstd);
        o.z = k * (b.z - (a.z - mean) * rstd);
        o.w = k * (b.w - (a.w - mean) * rstd);
        *reinterpret_cast<float4*>(drec + off + j * 4) = o;
    }
}

// ------------------------------------------------------------------------------------------------------------
// hard mask: one CTA per sample, in-smem bitonic sort of (key, index) pairs, L ≤ 1024
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pair_less(float ka, int ia, float kb, int ib) {
    return ka < kb || (ka == kb && ia < ib);
}

__device__ void bitonic_sort(float* key, int* idx, int n) {   // n power of two, ascending
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int p = i ^ j;
                if (p > i) {
                    bool up = (i & k) == 0;
                    bool lt = pair_less(key[p], idx[p], key[i], idx[i]);
                    if (lt == up) {
                        float tk = key[i]; key[i] = key[p]; key[p] = tk;
                        int ti = idx[i]; idx[i] = idx[p]; idx[p] = ti;
                    }
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(512)
hard_mask_kernel(const float* __restrict__ loss_pred, int L, int Lp2, int len_loss, int len_keep,
                 unsigned long long seed, unsigned long long offset, const unsigned long long* __restrict__ offset_dev,
                 int* __restrict__ hard, int* __restrict__ order, uint8_t* __restrict__ mask_out) {
    if (offset_dev) offset += *offset_dev * (unsigned long long)(gridDim.x * L);     // graph-replay safe RNG stream
    extern __shared__ unsigned char smraw[];
    float* key = reinterpret_cast<float*>(smraw);
    int* idx = reinterpret_cast<int*>(key + Lp2);
    uint8_t* is_hard = reinterpret_cast<uint8_t*>(idx + Lp2);
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < Lp2; i += blockDim.x) {
        key[i] = i < L ? loss_pred[(long)b * L + i] : __int_as_float(0x7f800000);
        idx[i] = i;
        if (i < L) is_hard[i] = 0;
    }
    bitonic_sort(key, idx, Lp2);
    if (order)
        for (int i = threadIdx.x; i < L; i += blockDim.x) order[(long)b * L + i] = idx[i];
    // hard set = the len_loss largest losses = sorted positions [L-len_loss, L)
    for (int i = threadIdx.x; i < len_loss; i += blockDim.x) {
        int id = idx[L - len_loss + i];
        if (hard) hard[(long)b * len_loss + i] = id;
        is_hard[id] = 1;
    }
    __syncthreads();
    if (mask_out == nullptr) return;
    // random fill: the len_keep smallest counter-based keys among the non-hard patches become visible
    for (int i = threadIdx.x; i < Lp2; i += blockDim.x) {
        float k = __int_as_float(0x7f800000);
        if (i < L && !is_hard[i]) {
            unsigned long long r = splitmix64(seed ^ splitmix64(offset + (unsigned long long)b * L + i));
            k = (float)(r >> 40) * (1.0f / 16777216.0f);
        }
        key[i] = k;
        idx[i] = i;
    }
    bitonic_sort(key, idx, Lp2);
    for (int i = threadIdx.x; i < L; i += blockDim.x) mask_out[(long)b * L + i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < len_keep; i += blockDim.x) mask_out[(long)b * L + idx[i]] = 1;
}

// ------------------------------------------------------------------------------------------------------------
// flat-arena EMA / AdamW / Σg²
// ------------------------------------------------------------------------------------------------------------
__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ model, long n, float d, float omd) {
    const long stride = (long)gridDim.x * blockDim.x;
    const long n4 = n >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 e = reinterpret_cast<float4*>(ema)[i];
        float4 m = reinterpret_cast<const float4*>(model)[i];
        // two separately rounded products then a rounded add — no FMA contraction (torch: ema*d + (1-d)*model)
        e.x = __fadd_rn(__fmul_rn(e.x, d), __fmul_rn(omd, m.x));
        e.y = __fadd_rn(__fmul_rn(e.y, d), __fmul_rn(omd, m.y));
        e.z = __fadd_rn(__fmul_rn(e.z, d), __fmul_rn(omd, m.z));
        e.w = __fadd_rn(__fmul_rn(e.w, d), __fmul_rn(omd, m.w));
        reinterpret_cast<float4*>(ema)[i] = e;
    }
    for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        ema[i] = __fadd_rn(__fmul_rn(ema[i], d), __fmul_rn(omd, model[i]));
}

// device-scalar variants (CUDA-graph replay: the per-step scalars live in a small device buffer refreshed by the host)
__global__ void ema_dev_kernel(float* __restrict__ ema, const float* __restrict__ model, long n,
                               const float* __restrict__ hyper) {
    const float d = hyper[0], omd = hyper[1];
    const long stride = (long)gridDim.x * blockDim.x;
    const long n4 = n >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 e = reinterpret_cast<float4*>(ema)[i];
        float4 m = reinterpret_cast<const float4*>(model)[i];
        e.x = __fadd_rn(__fmul_rn(e.x, d), __fmul_rn(omd, m.x));
        e.y = __fadd_rn(__fmul_rn(e.y, d), __fmul_rn(omd, m.y));
        e.z = __fadd_rn(__fmul_rn(e.z, d), __fmul_rn(omd, m.z));
        e.w = __fadd_rn(__fmul_rn(e.w, d), __fmul_rn(omd, m.w));
        reinterpret_cast<float4*>(ema)[i] = e;
    }
    for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        ema[i] = __fadd_rn(__fmul_rn(ema[i], d), __fmul_rn(omd, model[i]));
}

__global__ void sumsq_kernel(const float* __restrict__ g, long n, double* __restrict__ out) {
    __shared__ double sh[32];
    const long stride = (long)gridDim.x * blockDim.x;
    const long n4 = n >> 2;
    float acc = 0.f;
    double dacc = 0.0;
    int cnt = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = reinterpret_cast<const float4*>(g)[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        if (++cnt == 64) { dacc += acc; acc = 0.f; cnt = 0; }
    }
    for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += g[i] * g[i];
    dacc += acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = dacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        atomicAdd(out, t);
    }
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt, const double* __restrict__ gnorm_sq, float max_norm,
                             float gscale) {
    float coef = gscale;   // gscale = 1/world when g holds the all-reduced SUM of the ranks' gradients
    if (gnorm_sq) {   // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (norm + 1e-6), max=1)
        float norm = (float)sqrt(*gnorm_sq) * gscale;
        coef = fminf(max_norm / (norm + 1e-6f), 1.f) * gscale;
    }
    const float step_size = lr / bc1;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gi = g[i] * coef;
        float pi = p[i] * (1.f - lr * wd);
        float mi = m[i] + (gi - m[i]) * (1.f - b1);
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
        m[i] = mi;
        v[i] = vi;
    }
}

// hyper = {lr, beta1, beta2, eps, wd, bc1, sqrt(bc2), max_norm, gscale}
__global__ void adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long n, const float* __restrict__ hyper,
                                 const double* __restrict__ gnorm_sq) {
    const float lr = hyper[2 + 0], b1 = hyper[2 + 1], b2 = hyper[2 + 2], eps = hyper[2 + 3], wd = hyper[2 + 4],
                bc1 = hyper[2 + 5], bc2_sqrt = hyper[2 + 6], max_norm = hyper[2 + 7], gscale = hyper[2 + 8];
    float coef = gscale;
    if (gnorm_sq) {
        float norm = (float)sqrt(*gnorm_sq) * gscale;
        coef = fminf(max_norm / (norm + 1e-6f), 1.f) * gscale;
    }
    const float step_size = lr / bc1;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gi = g[i] * coef;
        float pi = p[i] * (1.f - lr * wd);
        float mi = m[i] + (gi - m[i]) * (1.f - b1);
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
        m[i] = mi;
        v[i] = vi;
    }
}

// ------------------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int T, int A, int B,
                                   long st, long sa, long sb) {
    const long total = (long)T * A * B;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int b = (int)(i % B);
        long r = i / B;
        int a = (int)(r % A);
        int t = (int)(r / A);
        dst[i] = __float2bfloat16(src[t * st + a * sa + b * sb]);
    }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int T, int A, int B,
                                    long st, long sa, long sb) {
    const long total = (long)T * A * B;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int b = (int)(i % B);
        long r = i / B;
        int a = (int)(r % A);
        int t = (int)(r / A);
        dst[t * st + a * sa + b * sb] = src[i];
    }
}

// ------------------------------------------------------------------------------------------------------------
// stem: masked fp32 input (Cin = 1) → conv1 k3 (+b1) and conv3 k1 (+b3) at active voxels, bf16 channels-last out
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float masked_input(const Geo& g, const float* __restrict__ inp, int n, int z, int y, int x) {
    if ((unsigned)z >= (unsigned)g.D || (unsigned)y >= (unsigned)g.H || (unsigned)x >= (unsigned)g.W) return 0.f;
    if (!g.active[((n * g.fd + (z >> g.lgP)) * g.fh + (y >> g.lgP)) * g.fw + (x >> g.lgP)]) return 0.f;
    return inp[(((long)n * g.D + z) * g.H + y) * g.W + x];
}

__global__ void __launch_bounds__(256)
stem_fwd_kernel(Geo g, const float* __restrict__ inp, const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w3, const float* __restrict__ b3, bf16* __restrict__ out1,
                bf16* __restrict__ out3) {
    extern __shared__ float sw[];      // [27][C] w1 transposed, then b1[C], w3[C], b3[C]
    const int C = g.C, CG = C / 8;
    for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = w1[(i % C) * 27 + i / C];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        sw[27 * C + i] = b1[i];
        sw[28 * C + i] = w3[i];
        sw[29 * C + i] = b3[i];
    }
    __syncthreads();
    const long runs = geo_num_runs(g);
    const long total = (runs << g.lgP) * CG;
    // 32-bit index arithmetic throughout (host checks total < 2^31): five 64-bit divisions per thread cost more than the
    // 27-tap stencil itself
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < (uint32_t)total; item += gridDim.x * blockDim.x) {
        const int cg = (int)(item % (uint32_t)CG);
        const uint32_t t = item / (uint32_t)CG;
        const int v = (int)(t & (uint32_t)(g.P - 1));
        RunPos r = decode_run(g, (long)(t >> g.lgP));
        const long voxel = r.voxel + v;
        uint32_t vq = (uint32_t)voxel;
        const int x = (int)(vq % (uint32_t)g.W); vq /= (uint32_t)g.W;
        const int y = (int)(vq % (uint32_t)g.H); vq /= (uint32_t)g.H;
        const int z = (int)(vq % (uint32_t)g.D);
        float acc[8], o3[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = sw[27 * C + cg * 8 + j];
        float center = 0.f;
        // visibility / bounds are resolved per input ROW (9 rows), not per tap: inside a patch row only the two x-edge
        // voxels can see a different patch
        const int px = x >> g.lgP, xin = x & (g.P - 1);
#pragma unroll
        for (int r9 = 0; r9 < 9; ++r9) {
            const int iz = z + r9 / 3 - 1, iy = y + r9 % 3 - 1;
            float xv[3] = {0.f, 0.f, 0.f};
            if ((unsigned)iz < (unsigned)g.D && (unsigned)iy < (unsigned)g.H) {
                const uint8_t* arow = g.active + ((r.n * g.fd + (iz >> g.lgP)) * g.fh + (iy >> g.lgP)) * g.fw;
                const float* row = inp + (((long)r.n * g.D + iz) * g.H + iy) * g.W + x;
                const bool a_mid = arow[px] != 0;
                const bool a_lo = xin == 0 ? (px > 0 && arow[px - 1] != 0) : a_mid;
                const bool a_hi = xin == g.P - 1 ? (px + 1 < g.fw && arow[px + 1] != 0) : a_mid;
                if (a_lo) xv[0] = row[-1];
                if (a_mid) xv[1] = row[0];
                if (a_hi) xv[2] = row[1];
            }
            if (r9 == 4) center = xv[1];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4 w0 = *reinterpret_cast<const float4*>(&sw[(r9 * 3 + dx) * C + cg * 8]);
                const float4 w1 = *reinterpret_cast<const float4*>(&sw[(r9 * 3 + dx) * C + cg * 8 + 4]);
                acc[0] = fmaf(xv[dx], w0.x, acc[0]); acc[1] = fmaf(xv[dx], w0.y, acc[1]);
                acc[2] = fmaf(xv[dx], w0.z, acc[2]); acc[3] = fmaf(xv[dx], w0.w, acc[3]);
                acc[4] = fmaf(xv[dx], w1.x, acc[4]); acc[5] = fmaf(xv[dx], w1.y, acc[5]);
                acc[6] = fmaf(xv[dx], w1.z, acc[6]); acc[7] = fmaf(xv[dx], w1.w, acc[7]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o3[j] = fmaf(center, sw[28 * C + cg * 8 + j], sw[29 * C + cg * 8 + j]);
        store8(out1 + voxel * C + cg * 8, acc);
        store8(out3 + voxel * C + cg * 8, o3);
    }
}

// Same stem, one thread = 8 consecutive voxels of a patch row x one group of 8 channels.  The first kernel above issued
// 27 input loads + 27 visibility tests + 54 shared-memory weight loads per VOXEL and channel group and was issue-bound
// (0.31 ms for 215 MB of output, 7 % of DRAM speed).  Here the 9 input rows are loaded once per 8 voxels as 2×float4 + 2
// scalars, visibility is resolved once per row, and every weight pair read from shared memory feeds 8 voxels:
// ~250 instead of ~550 instructions per voxel and channel group.
__global__ void __launch_bounds__(256)
stem_fwd_run8_kernel(Geo g, const float* __restrict__ inp, const float* __restrict__ w1, const float* __restrict__ b1,
                     const float* __restrict__ w3, const float* __restrict__ b3, bf16* __restrict__ out1,
                     bf16* __restrict__ out3) {
    extern __shared__ float sw[];      // [27][C] w1 transposed, then b1[C], w3[C], b3[C]
    const int C = g.C, CG = C / 8;
    for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = w1[(i % C) * 27 + i / C];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        sw[27 * C + i] = b1[i];
        sw[28 * C + i] = w3[i];
        sw[29 * C + i] = b3[i];
    }
    __syncthreads();
    const long runs = geo_num_runs(g);
    const int segs = g.P >> 3;                                  // 8-voxel segments per run (P is a multiple of 8)
    const uint32_t total = (uint32_t)(runs * segs * CG);
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
        const int cg = (int)(item % (uint32_t)CG);
        const uint32_t t = item / (uint32_t)CG;
        const int seg = (int)(t % (uint32_t)segs);
        RunPos r = decode_run(g, (long)(t / (uint32_t)segs));
        const long voxel = r.voxel + seg * 8;
        uint32_t vq = (uint32_t)voxel;
        const int x = (int)(vq % (uint32_t)g.W); vq /= (uint32_t)g.W;
        const int y = (int)(vq % (uint32_t)g.H); vq /= (uint32_t)g.H;
        const int z = (int)(vq % (uint32_t)g.D);
        const int px = x >> g.lgP, xin = x & (g.P - 1);
        float acc[8][8];
#pragma unroll
        for (int v = 0; v < 8; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[v][j] = sw[27 * C + cg * 8 + j];
        float center[8];
#pragma unroll
        for (int r9 = 0; r9 < 9; ++r9) {
            const int iz = z + r9 / 3 - 1, iy = y + r9 % 3 - 1;
            float xv[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) xv[i] = 0.f;
            if ((unsigned)iz < (unsigned)g.D && (unsigned)iy < (unsigned)g.H) {
                const uint8_t* arow = g.active + ((r.n * g.fd + (iz >> g.lgP)) * g.fh + (iy >> g.lgP)) * g.fw;
                const float* row = inp + (((long)r.n * g.D + iz) * g.H + iy) * g.W + x;
                const bool a_mid = arow[px] != 0;
                const bool a_lo = xin == 0 ? (px > 0 && arow[px - 1] != 0) : a_mid;
                const bool a_hi = xin + 8 == g.P ? (px + 1 < g.fw && arow[px + 1] != 0) : a_mid;
                if (a_mid) {
                    const float4 q0 = __ldg(reinterpret_cast<const float4*>(row));
                    const float4 q1 = __ldg(reinterpret_cast<const float4*>(row + 4));
                    xv[1] = q0.x; xv[2] = q0.y; xv[3] = q0.z; xv[4] = q0.w;
                    xv[5] = q1.x; xv[6] = q1.y; xv[7] = q1.z; xv[8] = q1.w;
                }
                if (a_lo) xv[0] = __ldg(row - 1);
                if (a_hi) xv[9] = __ldg(row + 8);
            }
            if (r9 == 4) {
#pragma unroll
                for (int v = 0; v < 8; ++v) center[v] = xv[v + 1];
            }
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4 wa = *reinterpret_cast<const float4*>(&sw[(r9 * 3 + dx) * C + cg * 8]);
                const float4 wb = *reinterpret_cast<const float4*>(&sw[(r9 * 3 + dx) * C + cg * 8 + 4]);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int v = 0; v < 8; ++v)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[v][j] = fmaf(xv[v + dx], wv[j], acc[v][j]);
            }
        }
        float w3v[8], b3v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { w3v[j] = sw[28 * C + cg * 8 + j]; b3v[j] = sw[29 * C + cg * 8 + j]; }
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            float o3[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o3[j] = fmaf(center[v], w3v[j], b3v[j]);
            store8(out1 + (voxel + v) * C + cg * 8, acc[v]);
            store8(out3 + (voxel + v) * C + cg * 8, o3);
        }
    }
}

// Tiled stem forward (the default): run8 above still issues its 9 input-row loads behind 27 visibility tests — dependent
// global loads in the middle of the FMA stream (22 % FMA utilisation, 185 µs for 2x128^3).  Here a block takes SF_VOX voxels
// = ZS whole z-slices of ONE visible patch (P x P x ZS voxels), stages their masked input neighbourhood
// (ZS + 2) x (P + 2) x (P + 2) in shared memory — visibility and zero padding applied while staging, every input voxel loaded
// once per tile instead of once per run and tap row — and thread = (8-voxel segment, channel group) then runs LDS + FMA only.
// (A first version staged 9 rows per run, 4x the loads, each behind its visibility byte: 300 µs.)
#define SF_VOX 512
__global__ void __launch_bounds__(256)
stem_fwd_tiled_kernel(Geo g, const float* __restrict__ inp, const float* __restrict__ w1, const float* __restrict__ b1,
                      const float* __restrict__ w3, const float* __restrict__ b3, bf16* __restrict__ out1,
                      bf16* __restrict__ out3) {
    extern __shared__ float sw[];      // [27][C] w1 transposed, b1[C], w3[C], b3[C]; then sx [ZS + 2][P + 2][P + 2]
    const int C = g.C, CG = C / 8, P = g.P, R = SF_VOX >> g.lgP, PX = P + 2, ZS = R >> g.lgP;
    float* sx = sw + 30 * C;
    for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = w1[(i % C) * 27 + i / C];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        sw[27 * C + i] = b1[i];
        sw[28 * C + i] = w3[i];
        sw[29 * C + i] = b3[i];
    }
    const long nruns = geo_num_runs(g);               // a multiple of R: tiles never straddle patches
    const int segs = P >> 3;                          // 8-voxel segments per run
    const int n_stage = (ZS + 2) * PX * PX;
    for (long base = (long)blockIdx.x * R; base < nruns; base += (long)gridDim.x * R) {
        const RunPos r0 = decode_run(g, base);        // first run of the tile: row 0 of z-slice rz0 of its patch
        uint32_t t = (uint32_t)r0.voxel;
        const int x0 = (int)(t % (uint32_t)g.W); t /= (uint32_t)g.W;
        const int y0 = (int)(t % (uint32_t)g.H); t /= (uint32_t)g.H;
        const int z0 = (int)(t % (uint32_t)g.D);
        const int n = r0.n;
        __syncthreads();                              // previous tile consumed (and the weights staged on the first pass)
#pragma unroll 2
        for (int i = threadIdx.x; i < n_stage; i += blockDim.x) {
            const int xi = i % PX, q = i / PX;
            const int yi = q % PX, zi = q / PX;
            const int iz = z0 + zi - 1, iy = y0 + yi - 1, x = x0 + xi - 1;
            const bool inb = (unsigned)iz < (unsigned)g.D && (unsigned)iy < (unsigned)g.H && (unsigned)x < (unsigned)g.W;
            const int cz = inb ? iz : z0, cy = inb ? iy : y0, cx = inb ? x : x0;      // clamped: both loads always legal
            const uint8_t a = g.active[((n * g.fd + (cz >> g.lgP)) * g.fh + (cy >> g.lgP)) * g.fw + (cx >> g.lgP)];
            const float v = __ldg(inp + (((long)n * g.D + cz) * g.H + cy) * g.W + cx);
            sx[i] = (inb && a) ? v : 0.f;
        }
        __syncthreads();
        for (int item = threadIdx.x; item < (SF_VOX >> 3) * CG; item += blockDim.x) {
            const int cg = item % CG, seg = item / CG;
            const int r = seg / segs, s8 = seg - r * segs;
            const int ry = r & (P - 1), rz = r >> g.lgP;
            float acc[8][8];
#pragma unroll
            for (int v = 0; v < 8; ++v)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[v][j] = sw[27 * C + cg * 8 + j];
            float center[8];
#pragma unroll
            for (int r9 = 0; r9 < 9; ++r9) {
                const float* row = sx + ((rz + r9 / 3) * PX + ry + r9 % 3) * PX + s8 * 8;
                float xv[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) xv[i] = row[i];
                if (r9 == 4) {
#pragma unroll
                    for (int v = 0; v < 8; ++v) center[v] = xv[v + 1];
                }
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 wa = *reinterpret_cast<const float4*>(&sw[(r9 * 3 + dx) * C + cg * 8]);
                    const float4 wb = *reinterpret_cast<const float4*>(&sw[(r9 * 3 + dx) * C + cg * 8 + 4]);
                    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int v = 0; v < 8; ++v)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[v][j] = fmaf(xv[v + dx], wv[j], acc[v][j]);
                }
            }
            float w3v[8], b3v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { w3v[j] = sw[28 * C + cg * 8 + j]; b3v[j] = sw[29 * C + cg * 8 + j]; }
            const long voxel = (((long)n * g.D + z0 + rz) * g.H + y0 + ry) * g.W + x0 + s8 * 8;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                float o3[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o3[j] = fmaf(center[v], w3v[j], b3v[j]);
                store8(out1 + (voxel + v) * C + cg * 8, acc[v]);
                store8(out3 + (voxel + v) * C + cg * 8, o3);
            }
        }
    }
}

// thread = (voxel lane, channel group, "slot"): slots 0..26 = conv1 taps, 27 = db1, 28 = dw3, 29 = db3
__global__ void __launch_bounds__(1024)
stem_wgrad_kernel(Geo g, const float* __restrict__ inp, const bf16* __restrict__ dy1, const bf16* __restrict__ dy3,
                  float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw3, float* __restrict__ db3,
                  int lanes) {
    extern __shared__ float sacc[];    // [32][C]
    const int C = g.C, CG = C / 8;
    for (int i = threadIdx.x; i < 32 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const int slot = threadIdx.x & 31;
    const int cg = (threadIdx.x >> 5) % CG;
    const int lane = (threadIdx.x >> 5) / CG;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (slot < 30) {
        const int dz = slot < 27 ? slot / 9 - 1 : 0, dy = slot < 27 ? (slot / 3) % 3 - 1 : 0,
                  dx = slot < 27 ? slot % 3 - 1 : 0;
        const bool needs_x = slot < 27 || slot == 28;
        const bf16* __restrict__ dsrc = slot >= 28 ? dy3 : dy1;
        const long nruns = geo_num_runs(g);
        // one run = P consecutive voxels along x inside one visible patch: all index arithmetic is per run, the inner
        // loop is one 4-byte and one 16-byte load plus 8 FMAs per voxel
        for (long run = (long)blockIdx.x * lanes + lane; run < nruns; run += (long)gridDim.x * lanes) {
            RunPos r = decode_run(g, run);
            const int x0 = (int)(r.voxel % g.W);
            const int y = (int)((r.voxel / g.W) % g.H);
            const int z = (int)((r.voxel / ((long)g.W * g.H)) % g.D);
            const int iz = z + dz, iy = y + dy;
            const bool row_ok = (unsigned)iz < (unsigned)g.D && (unsigned)iy < (unsigned)g.H;
            bool a_lo = false, a_mid = false, a_hi = false;
            const float* row = inp;
            if (needs_x && row_ok) {
                const int px = x0 >> g.lgP;
                const uint8_t* arow = g.active + ((r.n * g.fd + (iz >> g.lgP)) * g.fh + (iy >> g.lgP)) * g.fw;
                a_mid = arow[px] != 0;
                a_lo = px > 0 && arow[px - 1] != 0;
                a_hi = px + 1 < g.fw && arow[px + 1] != 0;
                row = inp + (((long)r.n * g.D + iz) * g.H + iy) * g.W;
            }
            const bf16* dptr = dsrc + r.voxel * C + cg * 8;
            for (int v = 0; v < g.P; ++v) {
                float xv = 1.f;
                if (needs_x) {
                    const int x = x0 + v + dx;
                    const bool ok = row_ok && (v + dx < 0 ? a_lo : (v + dx >= g.P ? a_hi : a_mid));
                    xv = ok ? row[x] : 0.f;
                }
                float d[8];
                load8(dptr + (long)v * C, d);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(d[j], xv, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sacc[slot * C + cg * 8 + j], acc[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 30 * C; i += blockDim.x) {
        const int s = i / C, c = i % C;
        const float v = sacc[i];
        if (s < 27) atomicAdd(&dw1[c * 27 + s], v);
        else if (s == 27) atomicAdd(&db1[c], v);
        else if (s == 28) atomicAdd(&dw3[c], v);
        else atomicAdd(&db3[c], v);
    }
}

// Tiled variant (the default): a block stages SW_VOX visible voxels (SW_VOX / P runs) of dy1 / dy3 as fp32 and the masked
// 3 × 3 × (P + 2) input neighbourhood of every run in shared memory; thread = (run group, channel group, slot) then runs a
// pure LDS + FMA loop (the kernel above issues a dependent global load per voxel and slot and reached 270 GB/s).
#define SW_VOX 128
__global__ void __launch_bounds__(256)
stem_wgrad_tiled_kernel(Geo g, const float* __restrict__ inp, const bf16* __restrict__ dy1, const bf16* __restrict__ dy3,
                        float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw3,
                        float* __restrict__ db3) {
    extern __shared__ float sm[];
    const int C = g.C, CG = C / 8, P = g.P, R = SW_VOX >> g.lgP, PX = P + 2;
    float* sd1 = sm;                         // [SW_VOX][C]
    float* sd3 = sd1 + SW_VOX * C;           // [SW_VOX][C]
    float* sx = sd3 + SW_VOX * C;            // [R][9][PX]  masked input rows (dz, dy), x0 − 1 … x0 + P
    float* sacc = sx + R * 9 * PX;           // [32][C]
    int* meta = (int*)(sacc + 32 * C);       // [R][5] first voxel (−1 = no run), n, z, y, x0
    for (int i = threadIdx.x; i < 32 * C; i += blockDim.x) sacc[i] = 0.f;
    const int slot = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int cg = wq % CG, grp = wq / CG, G = 8 / CG;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const long nruns = geo_num_runs(g);
    for (long base = (long)blockIdx.x * R; base < nruns; base += (long)gridDim.x * R) {
        __syncthreads();                     // previous tile fully consumed (and sacc zeroed on the first pass)
        if ((int)threadIdx.x < R) {
            const long run = base + threadIdx.x;
            int* m = meta + threadIdx.x * 5;
            if (run < nruns) {
                const RunPos r = decode_run(g, run);
                uint32_t t = (uint32_t)r.voxel;
                m[0] = (int)t;
                m[4] = (int)(t % (uint32_t)g.W); t /= (uint32_t)g.W;
                m[3] = (int)(t % (uint32_t)g.H); t /= (uint32_t)g.H;
                m[2] = (int)(t % (uint32_t)g.D);
                m[1] = r.n;
            } else {
                m[0] = -1;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SW_VOX * CG; i += blockDim.x) {
            const int c8 = i % CG, vox = i / CG;
            const int vfirst = meta[(vox >> g.lgP) * 5];
            float a[8], b[8];
            if (vfirst >= 0) {
                const long off = ((long)vfirst + (vox & (P - 1))) * C + c8 * 8;
                load8(dy1 + off, a);
                load8(dy3 + off, b);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
            }
            float4* o1 = reinterpret_cast<float4*>(sd1 + vox * C + c8 * 8);
            float4* o3 = reinterpret_cast<float4*>(sd3 + vox * C + c8 * 8);
            o1[0] = make_float4(a[0], a[1], a[2], a[3]); o1[1] = make_float4(a[4], a[5], a[6], a[7]);
            o3[0] = make_float4(b[0], b[1], b[2], b[3]); o3[1] = make_float4(b[4], b[5], b[6], b[7]);
        }
        for (int i = threadIdx.x; i < R * 9 * PX; i += blockDim.x) {
            const int xi = i % PX, q = i / PX;
            const int r9 = q % 9, r = q / 9;
            const int* m = meta + r * 5;
            float v = 0.f;
            if (m[0] >= 0) {
                const int iz = m[2] + r9 / 3 - 1, iy = m[3] + r9 % 3 - 1, x = m[4] - 1 + xi;
                if ((unsigned)iz < (unsigned)g.D && (unsigned)iy < (unsigned)g.H && (unsigned)x < (unsigned)g.W &&
                    g.active[((m[1] * g.fd + (iz >> g.lgP)) * g.fh + (iy >> g.lgP)) * g.fw + (x >> g.lgP)])
                    v = inp[(((long)m[1] * g.D + iz) * g.H + iy) * g.W + x];
            }
            sx[i] = v;
        }
        __syncthreads();
        if (slot < 30) {
            const float* dbase = slot >= 28 ? sd3 : sd1;
            for (int r = grp; r < R; r += G) {
                const float* xrow = slot < 27 ? sx + (r * 9 + slot / 3) * PX + slot % 3
                                              : (slot == 28 ? sx + (r * 9 + 4) * PX + 1 : nullptr);
                const float* d = dbase + (r << g.lgP) * C + cg * 8;
#pragma unroll 4
                for (int v = 0; v < P; ++v) {
                    const float xv = xrow ? xrow[v] : 1.f;
                    const float4 a = *reinterpret_cast<const float4*>(d + v * C);
                    const float4 b = *reinterpret_cast<const float4*>(d + v * C + 4);
                    acc[0] = fmaf(a.x, xv, acc[0]); acc[1] = fmaf(a.y, xv, acc[1]);
                    acc[2] = fmaf(a.z, xv, acc[2]); acc[3] = fmaf(a.w, xv, acc[3]);
                    acc[4] = fmaf(b.x, xv, acc[4]); acc[5] = fmaf(b.y, xv, acc[5]);
                    acc[6] = fmaf(b.z, xv, acc[6]); acc[7] = fmaf(b.w, xv, acc[7]);
                }
            }
        }
    }
    __syncthreads();
    if (slot < 30) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sacc[slot * C + cg * 8 + j], acc[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 30 * C; i += blockDim.x) {
        const int s = i / C, c = i % C;
        const float v = sacc[i];
        if (v == 0.f) continue;
        if (s < 27) atomicAdd(&dw1[c * 27 + s], v);
        else if (s == 27) atomicAdd(&db1[c], v);
        else if (s == 28) atomicAdd(&dw3[c], v);
        else atomicAdd(&db3[c], v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// reconstruction head (C → 1)
// ------------------------------------------------------------------------------------------------------------
__global__ void proj_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                float* __restrict__ rec, long voxels, int C) {
    extern __shared__ float sw[];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const float bias = b[0];
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < voxels; v += (long)gridDim.x * blockDim.x) {
        float acc = bias;
        for (int c = 0; c < C; c += 8) {
            float f[8];
            load8(x + v * C + c, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(f[j], sw[c + j], acc);
        }
        rec[v] = acc;
    }
}

__global__ void __launch_bounds__(512)
proj_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ drec,
                bf16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long voxels, int C) {
    extern __shared__ float sacc[];    // [C] + 1
    const int CG = C / 8;
    for (int i = threadIdx.x; i <= C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const int cg = threadIdx.x % CG;
    float wv[8], acc[8], accb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { wv[j] = w[cg * 8 + j]; acc[j] = 0.f; }
    // blockDim is a multiple of CG: a thread keeps its channel group and walks voxels with a constant stride
    const long vstride = (long)gridDim.x * (blockDim.x / CG);
    for (long v = (long)blockIdx.x * (blockDim.x / CG) + threadIdx.x / CG; v < voxels; v += vstride) {
        const float d = drec[v];
        float f[8], o[8];
        load8(x + v * C + cg * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { o[j] = d * wv[j]; acc[j] = fmaf(d, f[j], acc[j]); }
        store8(dx + v * C + cg * 8, o);
        if (cg == 0) accb += d;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sacc[cg * 8 + j], acc[j]);
    if (cg == 0) atomicAdd(&sacc[C], accb);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&dw[i], sacc[i]);
    if (threadIdx.x == 0) atomicAdd(db, sacc[C]);
}

static int grid_cap(long items, int block, int per_sm) {
    long b = (items + block - 1) / block, cap = (long)num_sms() * per_sm;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int stem_geo(Geo& g, const uint8_t* active, const int* list, const int* count, int N, int D, int H, int W,
                    int fd, int fh, int fw, int C) {
    AMB_CHECK(C % 8 == 0 && C <= 256, AMB_ERR_ARG, "stem: C=%d must be a multiple of 8 and <= 256", C);
    AMB_CHECK(active != nullptr, AMB_ERR_ARG, "stem: active mask required");
    AMB_CHECK(D % fd == 0 && H / fh == D / fd && W / fw == D / fd && H % fh == 0 && W % fw == 0, AMB_ERR_ARG,
              "stem: bad mask grid");
    g.N = N; g.D = D; g.H = H; g.W = W; g.C = C;
    g.P = D / fd; g.lgP = 0;
    while ((1 << g.lgP) < g.P) g.lgP++;
    AMB_CHECK((1 << g.lgP) == g.P, AMB_ERR_ARG, "stem: patch edge must be a power of two");
    g.fd = fd; g.fh = fh; g.fw = fw;
    g.list = list; g.count = count; g.active = active;
    return 0;
}

}  // namespace amb

using namespace amb;

extern "C" const char* amb_last_error(void) { return g_err; }
extern "C" int amb_version(void) { return AMB_VERSION; }
extern "C" int amb_sm_arch(void) { return 100; }
extern "C" long amb_launch_count(void) { return g_launch_count; }
extern "C" void amb_reset_launch_count(void) { g_launch_count = 0; }
extern "C" const char* amb_last_conv_kernel(void) { return g_last_conv_kernel; }

extern "C" int amb_patch_loss_fwd(const float* inp, const float* rec, const uint8_t* active, int N, int D, int H,
                                  int W, int normalize, float* per_patch, float* loss, float* patch_stats,
                                  unsigned int* ticket, void* stream) {
    AMB_CHECK(D % 16 == 0 && H % 16 == 0 && W % 16 == 0, AMB_ERR_ARG, "patch loss: dims must be multiples of 16");
    AMB_CHECK(inp && rec && active && per_patch && ticket, AMB_ERR_ARG, "patch loss: null argument");
    int L = (D / 16) * (H / 16) * (W / 16);
    patch_loss_fwd_kernel<<<N * L, 256, 0, (cudaStream_t)stream>>>(inp, rec, active, N, D, H, W, normalize, per_patch,
                                                                   loss, patch_stats, ticket);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_patch_loss_bwd(const float* inp, const float* rec, const uint8_t* active, const float* patch_stats,
                                  const float* dloss, int N, int D, int H, int W, float* drec, void* stream) {
    AMB_CHECK(D % 16 == 0 && H % 16 == 0 && W % 16 == 0, AMB_ERR_ARG, "patch loss: dims must be multiples of 16");
    int L = (D / 16) * (H / 16) * (W / 16);
    patch_loss_bwd_kernel<<<N * L, 256, 0, (cudaStream_t)stream>>>(inp, rec, active, patch_stats, dloss, N, D, H, W,
                                                                   drec);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_hard_mask(const float* loss_pred, int B, int L, int len_loss, int len_keep,
                             unsigned long long seed, unsigned long long offset,
                             const unsigned long long* offset_dev, int* hard, int* order, uint8_t* mask_out,
                             void* stream) {
    AMB_CHECK(L >= 1 && L <= 4096, AMB_ERR_ARG, "hard mask: L=%d out of range (1..4096)", L);
    AMB_CHECK(len_loss >= 0 && len_loss <= L && len_keep >= 0 && len_keep + len_loss <= L, AMB_ERR_ARG,
              "hard mask: len_loss=%d len_keep=%d L=%d", len_loss, len_keep, L);
    int Lp2 = 1;
    while (Lp2 < L) Lp2 <<= 1;
    size_t smem = (size_t)Lp2 * 8 + L;
    hard_mask_kernel<<<B, 512, smem, (cudaStream_t)stream>>>(loss_pred, L, Lp2, len_loss, len_keep, seed, offset,
                                                             offset_dev, hard, order, mask_out);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_ema_update(float* ema, const float* model, long n, double decay, void* stream) {
    AMB_CHECK(((uintptr_t)ema % 16 == 0) && ((uintptr_t)model % 16 == 0), AMB_ERR_ARG, "ema: 16B alignment required");
    ema_kernel<<<grid_cap(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(ema, model, n, (float)decay,
                                                                              (float)(1.0 - decay));
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_step_dev(float* ema, const float* model, long n_ema, float* p, const float* g, float* m, float* v,
                            long n_live, const float* hyper, const double* gnorm_sq, int do_adamw, int do_ema,
                            void* stream) {
    if (do_adamw) {
        adamw_dev_kernel<<<grid_cap(n_live, 256, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n_live, hyper, gnorm_sq);
        AMB_LAUNCH_CHECK();
    }
    if (do_ema) {
        AMB_CHECK(((uintptr_t)ema % 16 == 0) && ((uintptr_t)model % 16 == 0), AMB_ERR_ARG, "ema: 16B alignment required");
        ema_dev_kernel<<<grid_cap(n_ema / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(ema, model, n_ema, hyper);
        AMB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int amb_sumsq(const float* g, long n, double* out, void* stream) {
    AMB_CHECK((uintptr_t)g % 16 == 0, AMB_ERR_ARG, "sumsq: 16B alignment required");
    sumsq_kernel<<<grid_cap(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(g, n, out);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_adamw_step(float* p, const float* g, float* m, float* v, long n, double lr, double beta1,
                              double beta2, double eps, double weight_decay, int step, const double* gnorm_sq,
                              double max_norm, double gscale, void* stream) {
    AMB_CHECK(step >= 1, AMB_ERR_ARG, "adamw: step must be >= 1");
    double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
    adamw_kernel<<<grid_cap(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)lr, (float)beta1,
                                                                        (float)beta2, (float)eps, (float)weight_decay,
                                                                        (float)bc1, (float)sqrt(bc2), gnorm_sq,
                                                                        (float)max_norm, (float)gscale);
    AMB_LAUNCH_CHECK();
    return 0;
}

// Tiled variant for the layouts the path actually uses: the parameter tensor has the taps innermost (stride 1) and one of
// the two channel axes next (stride T), i.e. param[(slow·F + fast)·T + t] ↔ packed[(t·A + a)·B + b].  A block moves a tile
// of 256 (a, b) positions × ≤32 taps through shared memory so that both sides are accessed in runs of ≥ 32 B — the
// element-wise kernels above read with a stride of T floats and ran at ~1/4 of HBM speed.
//   PACK  : fp32 parameters → bf16 packed          !PACK : fp32 packed gradient → fp32 parameter layout
template <bool PACK, bool BFAST>
__device__ __forceinline__ void repack_tile(float (*tile)[33], const float* __restrict__ src, void* __restrict__ dst_, int T,
                                            int A, int B, int tiles_b, int tchunks, uint32_t blk) {
    constexpr uint32_t TA = BFAST ? 4 : 16, TB = BFAST ? 64 : 16;     // powers of two: index splits are shifts
    const int tcx = (int)(blk % (uint32_t)tchunks); blk /= (uint32_t)tchunks;
    const int a0 = (int)(blk / (uint32_t)tiles_b) * (int)TA, b0 = (int)(blk % (uint32_t)tiles_b) * (int)TB;
    const int t0 = tcx * 32;
    const uint32_t tc = (uint32_t)((T - t0) < 32 ? (T - t0) : 32);
    const float inv_tc = 1.0f / (float)tc;                              // e / tc below is exact for e < 8192, tc <= 32
    const uint32_t n = 256u * tc;
    // parameter-layout index of tile position (la, lb), tap t0
    auto param_index = [&](uint32_t la, uint32_t lb) -> long {
        const long a = a0 + (int)la, b = b0 + (int)lb;
        return (BFAST ? (a * B + b) : (b * A + a)) * T + t0;
    };
    // parameter order: taps fastest, then the fast channel axis
    auto split_param = [&](uint32_t e, uint32_t& la, uint32_t& lb, uint32_t& t) {
        const uint32_t q = (uint32_t)(((float)e + 0.5f) * inv_tc);
        t = e - q * tc;
        if (BFAST) { lb = q % TB; la = q / TB; }
        else { la = q % TA; lb = q / TA; }
    };
    if (PACK) {
        for (uint32_t e = threadIdx.x; e < n; e += 256u) {
            uint32_t la, lb, t;
            split_param(e, la, lb, t);
            float v = 0.f;
            if (a0 + (int)la < A && b0 + (int)lb < B) v = __ldg(src + param_index(la, lb) + t);
            tile[la * TB + lb][t] = v;
        }
        __syncthreads();
        bf16* dst = (bf16*)dst_;
        constexpr uint32_t hb = TB / 2u;
        for (uint32_t e = threadIdx.x; e < n / 2u; e += 256u) {     // packed order: b pairs, a, tap
            const uint32_t lb = (e % hb) * 2u, q = e / hb;
            const uint32_t la = q % TA, t = q / TA;
            if (a0 + (int)la < A && b0 + (int)lb < B) {
                const uint32_t p = la * TB + lb;
                *reinterpret_cast<uint32_t*>(dst + ((long)(t0 + (int)t) * A + a0 + (int)la) * B + b0 + (int)lb) =
                    pack2(tile[p][t], tile[p + 1][t]);
            }
        }
    } else {
        for (uint32_t e = threadIdx.x; e < n; e += 256u) {          // packed order
            const uint32_t lb = e % TB, q = e / TB;
            const uint32_t la = q % TA, t = q / TA;
            float v = 0.f;
            if (a0 + (int)la < A && b0 + (int)lb < B) v = __ldg(src + ((long)(t0 + (int)t) * A + a0 + (int)la) * B + b0 + (int)lb);
            tile[la * TB + lb][t] = v;
        }
        __syncthreads();
        float* dst = (float*)dst_;
        for (uint32_t e = threadIdx.x; e < n; e += 256u) {
            uint32_t la, lb, t;
            split_param(e, la, lb, t);
            if (a0 + (int)la < A && b0 + (int)lb < B) dst[param_index(la, lb) + t] = tile[la * TB + lb][t];
        }
    }
}

template <bool PACK, bool BFAST>
__global__ void __launch_bounds__(256) repack_tiled_kernel(const float* __restrict__ src, void* __restrict__ dst, int T,
                                                           int A, int B, int tiles_b, int tchunks) {
    __shared__ float tile[256][33];
    repack_tile<PACK, BFAST>(tile, src, dst, T, A, B, tiles_b, tchunks, blockIdx.x);
}

// Taps-major master weights (the engine's parameter arena keeps conv weights as fp32 [tap][Cout][Cin], trainer.ParamArena):
// the forward operand is a plain fp32 → bf16 conversion, the input-gradient operand [tap][Cin][Cout] a per-tap transpose.
//   linear    : block = 2048 consecutive elements, a thread converts 8 (two 128-bit loads, one 128-bit store)
__device__ __forceinline__ void pack_linear_tile(const float* __restrict__ src, bf16* __restrict__ dst, long total, uint32_t blk) {
    const long i = ((long)blk * 256 + threadIdx.x) * 8;
    if (i >= total) return;
    if (i + 8 <= total && (((uintptr_t)(src + i)) & 15) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src + i) + 1);
        uint4 o;
        o.x = pack2(a.x, a.y); o.y = pack2(a.z, a.w); o.z = pack2(b.x, b.y); o.w = pack2(b.z, b.w);
        *reinterpret_cast<uint4*>(dst + i) = o;
    } else {
        for (long j = i; j < total && j < i + 8; ++j) dst[j] = __float2bfloat16(__ldg(src + j));
    }
}

//   transpose : src [t][B][A] (a fastest) → dst [t][A][B] (b fastest); block = 32 a x 64 b of one tap through shared memory,
//               128-byte runs on both sides
__device__ __forceinline__ void pack_transpose_tile(float (*tile)[33], const float* __restrict__ src, bf16* __restrict__ dst,
                                                    int A, int B, int tiles_b, int tiles_a, uint32_t blk) {
    const int tb = (int)(blk % (uint32_t)tiles_b); blk /= (uint32_t)tiles_b;
    const int ta = (int)(blk % (uint32_t)tiles_a);
    const long t = (long)(blk / (uint32_t)tiles_a);
    const int a0 = ta * 32, b0 = tb * 64;
    const float* s = src + t * A * B;
    bf16* d = dst + t * A * B;
    for (uint32_t e = threadIdx.x; e < 2048u; e += 256u) {
        const uint32_t la = e & 31u, lb = e >> 5;
        float v = 0.f;
        if (a0 + (int)la < A && b0 + (int)lb < B) v = __ldg(s + (long)(b0 + (int)lb) * A + a0 + (int)la);
        tile[lb][la] = v;
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < 1024u; e += 256u) {
        const uint32_t lb = (e & 31u) * 2u, la = e >> 5;
        if (a0 + (int)la < A && b0 + (int)lb < B)                   // B is even (channel counts are multiples of 8)
            *reinterpret_cast<uint32_t*>(d + (long)(a0 + (int)la) * B + b0 + (int)lb) = pack2(tile[lb][la], tile[lb + 1][la]);
    }
}

// every conv weight of a step packed by ONE launch: block → job by binary search over the jobs' first tile
__global__ void __launch_bounds__(256) repack_batched_kernel(const amb_pack_job* __restrict__ jobs, int n_jobs) {
    __shared__ float tile[256][33];
    __shared__ int s_job;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_jobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        s_job = lo;
    }
    __syncthreads();
    const amb_pack_job J = jobs[s_job];
    const uint32_t blk = blockIdx.x - (uint32_t)J.tile_begin;
    if (J.b_fast == 2) pack_linear_tile(J.src, (bf16*)J.dst, (long)J.T * J.A * J.B, blk);
    else if (J.b_fast == 3) pack_transpose_tile(tile, J.src, (bf16*)J.dst, J.A, J.B, J.tiles_b, J.tchunks, blk);
    else if (J.b_fast) repack_tile<true, true>(tile, J.src, J.dst, J.T, J.A, J.B, J.tiles_b, J.tchunks, blk);
    else repack_tile<true, false>(tile, J.src, J.dst, J.T, J.A, J.B, J.tiles_b, J.tchunks, blk);
}

__global__ void __launch_bounds__(256) pack_linear_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long total) {
    pack_linear_tile(src, dst, total, blockIdx.x);
}

__global__ void __launch_bounds__(256) pack_transpose_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int A, int B,
                                                             int tiles_b, int tiles_a) {
    __shared__ float tile[64][33];
    pack_transpose_tile(tile, src, dst, A, B, tiles_b, tiles_a, blockIdx.x);
}

// 1: b is the fast parameter axis, 0: a is; 2 / 3: taps-major source, dst order / per-tap transposed (pack only);
// −1: layout not covered by a tiled kernel
static int repack_case(int T, int A, int B, long st, long sa, long sb, bool pack = false) {
    if (pack && !(B & 1) && (T == 1 || st == (long)A * B)) {
        if (sa == B && sb == 1) return 2;
        if (sa == 1 && sb == A) return 3;
    }
    if (st != 1 || (B & 1)) return -1;
    if (sb == T && sa == (long)B * T) return 1;
    if (sa == T && sb == (long)A * T) return 0;
    return -1;
}

template <bool PACK>
static void launch_repack(const float* src, void* dst, int T, int A, int B, int b_fast, cudaStream_t st) {
    const int TA = b_fast ? 4 : 16, TB = b_fast ? 64 : 16;
    const int tiles_a = (A + TA - 1) / TA, tiles_b = (B + TB - 1) / TB, tchunks = (T + 31) / 32;
    if (b_fast)
        repack_tiled_kernel<PACK, true><<<tiles_a * tiles_b * tchunks, 256, 0, st>>>(src, dst, T, A, B, tiles_b, tchunks);
    else
        repack_tiled_kernel<PACK, false><<<tiles_a * tiles_b * tchunks, 256, 0, st>>>(src, dst, T, A, B, tiles_b, tchunks);
}

extern "C" int amb_pack_weight(const float* src, void* dst, int T, int A, int B, long st, long sa, long sb,
                               void* stream) {
    const int c = repack_case(T, A, B, st, sa, sb, true);
    if (c == 2) {
        const long total = (long)T * A * B;
        pack_linear_kernel<<<(unsigned)((total + 2047) / 2048), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, total);
    } else if (c == 3) {
        const int tiles_b = (B + 63) / 64, tiles_a = (A + 31) / 32;
        pack_transpose_kernel<<<(unsigned)((long)T * tiles_a * tiles_b), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, A, B,
                                                                                                        tiles_b, tiles_a);
    } else if (c >= 0)
        launch_repack<true>(src, dst, T, A, B, c, (cudaStream_t)stream);
    else
        pack_weight_kernel<<<grid_cap((long)T * A * B, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, T, A, B,
                                                                                               st, sa, sb);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_pack_weights_batched(const amb_pack_job* jobs_dev, int n_jobs, int total_tiles, void* stream) {
    AMB_CHECK(jobs_dev != nullptr && n_jobs > 0 && total_tiles > 0, AMB_ERR_ARG, "amb_pack_weights_batched: empty job table");
    repack_batched_kernel<<<total_tiles, 256, 0, (cudaStream_t)stream>>>(jobs_dev, n_jobs);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_unpack_wgrad(const float* src, float* dst, int T, int A, int B, long st, long sa, long sb,
                                void* stream) {
    const int c = repack_case(T, A, B, st, sa, sb);
    if (c >= 0)
        launch_repack<false>(src, dst, T, A, B, c, (cudaStream_t)stream);
    else
        unpack_wgrad_kernel<<<grid_cap((long)T * A * B, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, dst, T, A, B, st,
                                                                                                sa, sb);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_stem_fwd(const float* inp, const uint8_t* active, const int* active_list, const int* active_count,
                            int N, int D, int H, int W, int fd, int fh, int fw, int C, const float* w1,
                            const float* b1, const float* w3, const float* b3, void* out1, void* out3, void* stream) {
    Geo g;
    if (int e = stem_geo(g, active, active_list, active_count, N, D, H, W, fd, fh, fw, C)) return e;
    int block = (256 / (C / 8)) * (C / 8);
    AMB_CHECK((long)N * D * H * W * (C / 8) < (1L << 31), AMB_ERR_ARG, "stem: tensor too large for 32-bit indexing");
    if (g.P % 8 == 0 && g.P * g.P <= SF_VOX && !getenv("AMB_STEM_FWD_V1") && !getenv("AMB_STEM_FWD_RUN8")) {
        const int ZS = SF_VOX / (g.P * g.P);                         // whole z-slices of a patch per tile (P = 16: 2, P = 8: 8)
        const size_t smem = ((size_t)30 * C + (size_t)(ZS + 2) * (g.P + 2) * (g.P + 2)) * sizeof(float);
        if (ZS <= g.P && smem <= 200 * 1024) {
            AMB_CUDA(cudaFuncSetAttribute(stem_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            long tiles = ((long)N * D * H * W / SF_VOX);
            int grid = (int)(tiles < (long)num_sms() * 4 ? (tiles < 1 ? 1 : tiles) : (long)num_sms() * 4);
            stem_fwd_tiled_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(g, inp, w1, b1, w3, b3, (bf16*)out1, (bf16*)out3);
            AMB_LAUNCH_CHECK();
            return 0;
        }
    }
    if (g.P % 8 == 0 && W % 4 == 0 && !getenv("AMB_STEM_FWD_V1")) {       // 8 voxels per thread (float4 input rows)
        stem_fwd_run8_kernel<<<grid_cap((long)N * D * H * W / 8 * (C / 8), block, 4), block, 30 * C * sizeof(float),
                               (cudaStream_t)stream>>>(g, inp, w1, b1, w3, b3, (bf16*)out1, (bf16*)out3);
        AMB_LAUNCH_CHECK();
        return 0;
    }
    stem_fwd_kernel<<<grid_cap((long)N * D * H * W * (C / 8), block, 8), block, 30 * C * sizeof(float),
                      (cudaStream_t)stream>>>(g, inp, w1, b1, w3, b3, (bf16*)out1, (bf16*)out3);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_stem_wgrad(const float* inp, const uint8_t* active, const int* active_list,
                              const int* active_count, int N, int D, int H, int W, int fd, int fh, int fw, int C,
                              const void* dy1, const void* dy3, float* dw1, float* db1, float* dw3, float* db3,
                              void* stream) {
    Geo g;
    if (int e = stem_geo(g, active, active_list, active_count, N, D, H, W, fd, fh, fw, C)) return e;
    int CG = C / 8;
    if ((CG == 1 || CG == 2 || CG == 4 || CG == 8) && g.P <= SW_VOX && (SW_VOX / g.P) % (8 / CG) == 0 &&
        !getenv("AMB_STEM_WGRAD_V1")) {
        const int R = SW_VOX / g.P;
        const size_t smem = ((size_t)2 * SW_VOX * C + (size_t)R * 9 * (g.P + 2) + 32 * C) * sizeof(float) + (size_t)R * 5 * sizeof(int);
        AMB_CUDA(cudaFuncSetAttribute(stem_wgrad_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stem_wgrad_tiled_kernel<<<num_sms() * 4, 256, smem, (cudaStream_t)stream>>>(g, inp, (const bf16*)dy1, (const bf16*)dy3,
                                                                                  dw1, db1, dw3, db3);
        AMB_LAUNCH_CHECK();
        return 0;
    }
    int lanes = 1024 / (32 * CG);
    AMB_CHECK(lanes >= 1, AMB_ERR_ARG, "stem wgrad: C=%d too large", C);
    int block = lanes * 32 * CG;
    stem_wgrad_kernel<<<num_sms() * 2, block, 32 * C * sizeof(float), (cudaStream_t)stream>>>(
        g, inp, (const bf16*)dy1, (const bf16*)dy3, dw1, db1, dw3, db3, lanes);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_proj_fwd(const void* x, const float* w, const float* b, float* rec, long voxels, int C,
                            void* stream) {
    AMB_CHECK(C % 8 == 0, AMB_ERR_ARG, "proj: C must be a multiple of 8");
    proj_fwd_kernel<<<grid_cap(voxels, 256, 8), 256, C * sizeof(float), (cudaStream_t)stream>>>((const bf16*)x, w, b,
                                                                                               rec, voxels, C);
    AMB_LAUNCH_CHECK();
    return 0;
}

extern "C" int amb_proj_bwd(const void* x, const float* w, const float* drec, void* dx, float* dw, float* db,
                            long voxels, int C, void* stream) {
    AMB_CHECK(C % 8 == 0 && C / 8 <= 512, AMB_ERR_ARG, "proj: C must be a multiple of 8");
    int CG = C / 8, block = (256 / CG) * CG;
    if (block == 0) block = CG;
    proj_bwd_kernel<<<grid_cap(voxels * CG, block, 4), block, (C + 1) * sizeof(float), (cudaStream_t)stream>>>(
        (const bf16*)x, w, drec, (bf16*)dx, dw, db, voxels, C);
    AMB_LAUNCH_CHECK();
    return 0;
}
