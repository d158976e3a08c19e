// CUDA-core execution of a gather-GEMM plan (conv_plan.cuh): the differential-testing partner of the tcgen05 kernel and
// the path for channel counts the tensor-core kernel does not take (C % 16 != 0).  Also hosts the amb_conv /
// amb_conv_wgrad entry points and their dispatch.
#include "conv_plan.cuh"

namespace amb {

struct DirectParams {
    Plan plan;
    const bf16* x;
    bf16* y;
    const bf16* w;
    const float* bias;
    const float* ep_scale;
    int ep_act;
    const uint8_t* active;
    const int* list;
    const int* count;
};

__device__ __forceinline__ bool out_voxel_active(const Plan& p, const uint8_t* active, int n, int z, int y, int x) {
    return active[((n * p.fd + (z >> p.lgPv)) * p.fh + (y >> p.lgPv)) * p.fw + (x >> p.lgPv)] != 0;
}

// one warp per (group, out voxel); lanes split the output channels, each lane loops taps × Cx with 128-bit loads
__global__ void __launch_bounds__(256) direct_conv_kernel(const __grid_constant__ DirectParams P) {
    const Plan& p = P.plan;
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    long vox_per_group;
    if (P.list) vox_per_group = (long)(*P.count) << (3 * p.lgPv);
    else vox_per_group = (long)p.oN * p.oD * p.oH * p.oW;
    const long total = vox_per_group * p.n_groups;
    for (long item = warp; item < total; item += nwarps) {
        const int g = (int)(item / vox_per_group);
        long v = item % vox_per_group;
        int n, z, y, x;
        if (P.list) {
            const int Pv = 1 << p.lgPv;
            int sub = (int)(v & ((1 << (3 * p.lgPv)) - 1));
            int pid = P.list[v >> (3 * p.lgPv)];
            int L = p.fd * p.fh * p.fw;
            n = pid / L;
            int l = pid % L;
            z = (l / (p.fh * p.fw)) * Pv + (sub >> (2 * p.lgPv));
            y = ((l / p.fw) % p.fh) * Pv + ((sub >> p.lgPv) & (Pv - 1));
            x = (l % p.fw) * Pv + (sub & (Pv - 1));
        } else {
            x = (int)(v % p.oW); v /= p.oW;
            y = (int)(v % p.oH); v /= p.oH;
            z = (int)(v % p.oD);
            n = (int)(v / p.oD);
        }
        const Group& G = p.groups[g];
        const View& ov = p.out_views[G.out_view];
        bf16* yrow = P.y + ov.base + n * ov.sN + z * ov.sD + y * ov.sH + x * ov.sW;
        const bool on = (p.lgPv < 0 || P.active == nullptr) ? true : out_voxel_active(p, P.active, n, z, y, x);
        for (int r = lane; r < p.Cy; r += 32) {
            float acc = 0.f;
            if (on) {
                for (int t = G.tap_begin; t < G.tap_begin + G.tap_count; ++t) {
                    const Tap& T = p.taps[t];
                    const View& iv = p.in_views[T.view];
                    const int iz = z + T.dz, iy = y + T.dy, ix = x + T.dx;
                    if ((unsigned)iz >= (unsigned)iv.D || (unsigned)iy >= (unsigned)iv.H || (unsigned)ix >= (unsigned)iv.W)
                        continue;
                    const bf16* xr = P.x + iv.base + n * iv.sN + iz * iv.sD + iy * iv.sH + ix * iv.sW;
                    const bf16* wr = P.w + ((long)T.w * p.Cy + r) * p.Cx;
                    for (int c = 0; c < p.Cx; c += 8) {
                        float a[8], b[8];
                        load8(xr + c, a);
                        load8(wr + c, b);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc = fmaf(a[j], b[j], acc);
                    }
                }
                acc = ep_apply(acc, P.ep_scale ? P.ep_scale[r] : 1.f, P.bias ? P.bias[r] : 0.f, P.ep_act);
            }
            yrow[r] = __float2bfloat16(acc);
        }
    }
}

struct DirectWgradParams {
    Plan plan;
    const bf16* x;
    const bf16* dy;
    float* dw;
    const int* list;
    const int* count;
    int vox_chunks;
};

// block = (slab row r, 8-column group) pairs for one tap; grid.y splits the voxels; fp32 atomics into dw
__global__ void __launch_bounds__(256) direct_wgrad_kernel(const __grid_constant__ DirectWgradParams P) {
    const Plan& p = P.plan;
    const int CXG = p.Cx / 8;
    const long pairs = (long)p.Cy * CXG;
    const long pid_ = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid_ >= pairs * p.n_taps) return;
    const int t = (int)(pid_ / pairs);
    const int r = (int)((pid_ % pairs) / CXG);
    const int cg = (int)(pid_ % CXG);
    const Tap& T = p.taps[t];
    int g = 0;
    for (int i = 0; i < p.n_groups; ++i)
        if (t >= p.groups[i].tap_begin && t < p.groups[i].tap_begin + p.groups[i].tap_count) g = i;
    const View& ov = p.out_views[p.groups[g].out_view];
    const View& iv = p.in_views[T.view];
    long nvox;
    if (P.list) nvox = (long)(*P.count) << (3 * p.lgPv);
    else nvox = (long)p.oN * p.oD * p.oH * p.oW;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long v0 = blockIdx.y; v0 < nvox; v0 += gridDim.y) {
        long v = v0;
        int n, z, y, x;
        if (P.list) {
            const int Pv = 1 << p.lgPv;
            int sub = (int)(v & ((1 << (3 * p.lgPv)) - 1));
            int pid = P.list[v >> (3 * p.lgPv)];
            int L = p.fd * p.fh * p.fw;
            n = pid / L;
            int l = pid % L;
            z = (l / (p.fh * p.fw)) * Pv + (sub >> (2 * p.lgPv));
            y = ((l / p.fw) % p.fh) * Pv + ((sub >> p.lgPv) & (Pv - 1));
            x = (l % p.fw) * Pv + (sub & (Pv - 1));
        } else {
            x = (int)(v % p.oW); v /= p.oW;
            y = (int)(v % p.oH); v /= p.oH;
            z = (int)(v % p.oD);
            n = (int)(v / p.oD);
        }
        const int iz = z + T.dz, iy = y + T.dy, ix = x + T.dx;
        if ((unsigned)iz >= (unsigned)iv.D || (unsigned)iy >= (unsigned)iv.H || (unsigned)ix >= (unsigned)iv.W) continue;
        const float d = bf2f(P.dy[ov.base + n * ov.sN + z * ov.sD + y * ov.sH + x * ov.sW + r]);
        float a[8];
        load8(P.x + iv.base + n * iv.sN + iz * iv.sD + iy * iv.sH + ix * iv.sW + cg * 8, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(d, a[j], acc[j]);
    }
    float* out = P.dw + ((long)T.w * p.Cy + r) * p.Cx + cg * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(out + j, acc[j]);
}

static int direct_conv(const Plan& p, const amb_conv_args* a) {
    AMB_CHECK(a->stats == nullptr, AMB_ERR_UNSUPPORTED, "conv: fused stats need the tcgen05 path");
    DirectParams P;
    P.plan = p;
    P.x = (const bf16*)a->x; P.y = (bf16*)a->y; P.w = (const bf16*)a->w; P.bias = a->bias; P.ep_scale = a->ep_scale; P.ep_act = a->ep_act;
    P.active = a->active; P.list = a->active_list; P.count = a->active_count;
    long warps = (long)p.oN * p.oD * p.oH * p.oW * p.n_groups;
    long blocks = (warps + 7) / 8;
    long cap = (long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    direct_conv_kernel<<<(int)blocks, 256, 0, (cudaStream_t)a->stream>>>(P);
    AMB_LAUNCH_CHECK();
    g_last_conv_kernel = "direct_conv_kernel";
    return 0;
}

static int direct_wgrad(const Plan& p, const amb_wgrad_args* a) {
    DirectWgradParams P;
    P.plan = p;
    P.x = (const bf16*)a->x; P.dy = (const bf16*)a->dy; P.dw = a->dw;
    P.list = a->active_list; P.count = a->active_count;
    long pairs = (long)p.Cy * (p.Cx / 8) * p.n_taps;
    long nvox = (long)p.oN * p.oD * p.oH * p.oW;
    int bx = ceil_div(pairs, 256);
    long by = (long)num_sms() * 8 / bx;
    if (by < 1) by = 1;
    if (by > nvox) by = nvox;
    if (by > 65535) by = 65535;
    dim3 grid(bx, (unsigned)by);
    direct_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)a->stream>>>(P);
    AMB_LAUNCH_CHECK();
    g_last_conv_kernel = "direct_wgrad_kernel";
    return 0;
}

}  // namespace amb

using namespace amb;

extern "C" int amb_conv(const amb_conv_args* a) {
    AMB_CHECK(a && a->x && a->y && a->w, AMB_ERR_ARG, "amb_conv: null argument");
    AMB_CHECK(a->Cin % 8 == 0 && a->Cout % 8 == 0, AMB_ERR_ARG, "amb_conv: Cin=%d / Cout=%d must be multiples of 8", a->Cin, a->Cout);
    Plan p;
    if (int e = build_plan(p, a->op, a->N, a->D, a->H, a->W, a->Cin, a->Cout, a->k, a->stride)) return e;
    if (a->active) {
        AMB_CHECK(a->op == AMB_OP_CONV || a->op == AMB_OP_CONV_DGRAD, AMB_ERR_ARG, "amb_conv: masks only on CONV / CONV_DGRAD");
        int fD = a->op == AMB_OP_CONV ? a->D / a->stride : a->D;
        int fH = a->op == AMB_OP_CONV ? a->H / a->stride : a->H;
        int fW = a->op == AMB_OP_CONV ? a->W / a->stride : a->W;
        if (int e = plan_set_mask(p, fD, fH, fW, a->fd, a->fh, a->fw)) return e;
    }
    AMB_CHECK((a->active_list == nullptr) == (a->active_count == nullptr), AMB_ERR_ARG, "amb_conv: list and count go together");
    AMB_CHECK(a->active_list == nullptr || a->active != nullptr, AMB_ERR_ARG, "amb_conv: active_list needs active");
    if (a->impl != AMB_IMPL_DIRECT) {
        // halo-plane kernel (v2) for dense 3x3x3 s1; the per-tap kernel (v1) everywhere else, and wherever the
        // active-patch work-list lets it skip masked tiles outright (patch edge >= 8 at the output resolution)
        const bool list_pays = a->active_list != nullptr && p.lgPv >= 3;
        if (a->impl != AMB_IMPL_TCGEN05_V1 && (!list_pays || p.lgPv >= 4)) {   // v3 takes the list itself at edge >= 16
            int r4 = igemm4_conv(p, a);          // halo planes, dz taps stacked along N: narrow 3x3x3 s1 layers
            if (r4 < 0) return r4;
            if (r4 == 1) return 0;
            int r3 = igemm3_conv(p, a);          // round-1 form (one MMA per tile and tap), kept as the A/B baseline
            if (r3 < 0) return r3;
            if (r3 == 1) return 0;
        }
        if (a->op == AMB_OP_CONVT && a->impl != AMB_IMPL_TCGEN05_V1) {
            int rt = igemm4t_conv(p, a);         // ConvTranspose forward, Cout <= 64: kz taps stacked along N over halo planes
            if (rt < 0) return rt;
            if (rt == 1) return 0;
        }
        int r = igemm_conv(p, a);
        if (r < 0) return r;
        if (r == 1) return 0;
        AMB_CHECK(a->impl == AMB_IMPL_AUTO, AMB_ERR_UNSUPPORTED,
                  "amb_conv: shape not supported by the tcgen05 kernel (Cx=%d Cy=%d): %s", p.Cx, p.Cy, amb_last_error());
    }
    return direct_conv(p, a);
}

extern "C" long amb_conv_workspace_bytes(const amb_conv_args* a) {
    if (!a || a->impl == AMB_IMPL_DIRECT || a->Cin % 8 != 0 || a->Cout % 8 != 0) return 0;
    Plan p;
    if (build_plan(p, a->op, a->N, a->D, a->H, a->W, a->Cin, a->Cout, a->k, a->stride)) return 0;
    if (a->active && (a->op == AMB_OP_CONV || a->op == AMB_OP_CONV_DGRAD)) {
        int fD = a->op == AMB_OP_CONV ? a->D / a->stride : a->D;
        int fH = a->op == AMB_OP_CONV ? a->H / a->stride : a->H;
        int fW = a->op == AMB_OP_CONV ? a->W / a->stride : a->W;
        if (plan_set_mask(p, fD, fH, fW, a->fd, a->fh, a->fw)) return 0;
    }
    // the active-patch work-list (patch edge >= 8) and split-K exclude each other: listed layers have plenty of tiles
    const bool sparse_list = a->active_list != nullptr && p.lgPv >= 3;
    return igemm_workspace_bytes(p, sparse_list);
}

extern "C" int amb_conv_wgrad(const amb_wgrad_args* a) {
    AMB_CHECK(a && a->x && a->dy && a->dw, AMB_ERR_ARG, "amb_conv_wgrad: null argument");
    AMB_CHECK(a->op == AMB_OP_CONV || a->op == AMB_OP_CONVT, AMB_ERR_ARG, "amb_conv_wgrad: op must be CONV or CONVT");
    AMB_CHECK(a->Cin % 8 == 0 && a->Cout % 8 == 0, AMB_ERR_ARG, "amb_conv_wgrad: channels must be multiples of 8");
    Plan p;
    if (int e = build_plan(p, a->op, a->N, a->D, a->H, a->W, a->Cin, a->Cout, a->k, a->stride)) return e;
    if (a->active_list) {
        AMB_CHECK(a->op == AMB_OP_CONV && a->active_count, AMB_ERR_ARG, "amb_conv_wgrad: active list only for CONV");
        if (int e = plan_set_mask(p, a->D / a->stride, a->H / a->stride, a->W / a->stride, a->fd, a->fh, a->fw)) return e;
    }
    if (a->impl != AMB_IMPL_DIRECT) {
        int r = igemm_wgrad(p, a);
        if (r < 0) return r;
        if (r == 1) return 0;
        AMB_CHECK(a->impl != AMB_IMPL_TCGEN05, AMB_ERR_UNSUPPORTED,
                  "amb_conv_wgrad: shape not supported by the tcgen05 kernel: %s", amb_last_error());
    }
    return direct_wgrad(p, a);
}
