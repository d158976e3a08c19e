// Gather-GEMM plan: every convolution on the path (Conv3d k1/k3 s1/s2, ConvTranspose3d k4 s2 p1, and their input
// gradients) is expressed as
//     out_view[g][o, :] = Σ_{tap ∈ group g}  W[tap.w] · in_view[tap.view][o + (tap.dz, tap.dy, tap.dx), :]
// where a *view* is a strided window of a channels-last tensor (the whole tensor, or one of its 8 stride-2 parity
// classes) and out-of-range reads are zero.  The CUDA-core gather kernel and the tcgen05/TMA implicit-GEMM kernel both
// execute the same plan, so one can check the other.
#pragma once
#include "common.cuh"

namespace amb {

struct View {
    long base;             // element offset of view voxel (0,0,0,0) channel 0 from the tensor base
    long sN, sD, sH, sW;   // element strides of the view axes
    int N, D, H, W;        // view extents
};

struct Tap {
    int8_t view, dz, dy, dx;
    int16_t w;             // weight slab index
    int16_t pad;
};

struct Group {
    int16_t tap_begin, tap_count;
    int16_t out_view, pad;
};

struct Plan {
    int n_in_views, n_out_views, n_groups, n_taps;
    int Cx, Cy;            // channels of the gathered tensor (contraction) and of the produced tensor (rows of W)
    int oN, oD, oH, oW;    // extent of every out view
    int lgPv;              // log2 patch edge in out-view coordinates (mask / active-list granularity), -1 = no mask
    int fd, fh, fw;        // mask grid
    View in_views[8];
    View out_views[8];
    Group groups[8];
    Tap taps[64];
};

static inline View full_view(int N, int D, int H, int W, int C) {
    View v;
    v.base = 0;
    v.sW = C; v.sH = (long)W * C; v.sD = (long)H * W * C; v.sN = (long)D * H * W * C;
    v.N = N; v.D = D; v.H = H; v.W = W;
    return v;
}

static inline View parity_view(int N, int D, int H, int W, int C, int pz, int py, int px) {
    View v;
    v.base = (((long)pz * H + py) * W + px) * C;
    v.sW = 2L * C; v.sH = 2L * W * C; v.sD = 2L * H * W * C; v.sN = (long)D * H * W * C;
    v.N = N; v.D = D / 2; v.H = H / 2; v.W = W / 2;
    return v;
}

// Builds the plan for amb_conv_args-style arguments.  Returns 0 or an error code (message set).
static inline int build_plan(Plan& p, int op, int N, int D, int H, int W, int Cin, int Cout, int k, int stride) {
    memset(&p, 0, sizeof(p));
    p.lgPv = -1;
    auto add_tap = [&](int view, int dz, int dy, int dx, int w) {
        Tap& t = p.taps[p.n_taps++];
        t.view = (int8_t)view; t.dz = (int8_t)dz; t.dy = (int8_t)dy; t.dx = (int8_t)dx; t.w = (int16_t)w; t.pad = 0;
    };
    if (op == AMB_OP_CONV || op == AMB_OP_CONV_DGRAD) {
        AMB_CHECK((k == 1 || k == 3) && (stride == 1 || stride == 2), AMB_ERR_ARG, "conv: k=%d stride=%d unsupported", k, stride);
        AMB_CHECK(stride == 1 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), AMB_ERR_ARG, "conv: stride 2 needs even dims");
    } else {
        AMB_CHECK(k == 4 && stride == 2, AMB_ERR_ARG, "convT: only k4 s2 p1");
    }
    const int h = k / 2;
    if (op == AMB_OP_CONV) {
        p.Cx = Cin; p.Cy = Cout;
        p.oN = N; p.oD = D / stride; p.oH = H / stride; p.oW = W / stride;
        p.n_out_views = 1; p.out_views[0] = full_view(N, p.oD, p.oH, p.oW, Cout);
        p.n_groups = 1; p.groups[0].tap_begin = 0; p.groups[0].out_view = 0;
        if (stride == 1) {
            p.n_in_views = 1; p.in_views[0] = full_view(N, D, H, W, Cin);
            for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) for (int c = 0; c < k; ++c)
                add_tap(0, a - h, b - h, c - h, (a * k + b) * k + c);
        } else {
            p.n_in_views = 8;
            for (int q = 0; q < 8; ++q) p.in_views[q] = parity_view(N, D, H, W, Cin, q >> 2, (q >> 1) & 1, q & 1);
            // input index 2o + kk - h : (parity, offset) per axis
            auto po = [&](int kk, int& par, int& off) {
                int rel = kk - h;            // -1, 0, +1 (k=3) or 0 (k=1)
                par = rel & 1;               // -1 → 1, 0 → 0, 1 → 1
                off = (rel - par) / 2;       // -1 → -1, 0 → 0, 1 → 0
            };
            for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) for (int c = 0; c < k; ++c) {
                int pa, oa, pb, ob, pc, oc;
                po(a, pa, oa); po(b, pb, ob); po(c, pc, oc);
                add_tap(pa * 4 + pb * 2 + pc, oa, ob, oc, (a * k + b) * k + c);
            }
        }
        p.groups[0].tap_count = (int16_t)p.n_taps;
    } else if (op == AMB_OP_CONV_DGRAD) {
        // gathered = dy (N, D/s, H/s, W/s, Cout) ; produced = dx (N, D, H, W, Cin) ; W slab [tap][Cin][Cout]
        p.Cx = Cout; p.Cy = Cin;
        if (stride == 1) {
            p.oN = N; p.oD = D; p.oH = H; p.oW = W;
            p.n_in_views = 1; p.in_views[0] = full_view(N, D, H, W, Cout);
            p.n_out_views = 1; p.out_views[0] = full_view(N, D, H, W, Cin);
            p.n_groups = 1; p.groups[0].tap_begin = 0; p.groups[0].out_view = 0;
            for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) for (int c = 0; c < k; ++c)
                add_tap(0, h - a, h - b, h - c, (a * k + b) * k + c);
            p.groups[0].tap_count = (int16_t)p.n_taps;
        } else {
            p.oN = N; p.oD = D / 2; p.oH = H / 2; p.oW = W / 2;
            p.n_in_views = 1; p.in_views[0] = full_view(N, D / 2, H / 2, W / 2, Cout);
            // dx[i] = Σ_{o,kk : 2o + kk - h = i} dy[o] W_kk^T.  Per axis, fine index i = 2j + par:
            //   k=3: par 0 → (kk=1, o=j) ; par 1 → (kk=0, o=j+1), (kk=2, o=j)       k=1: par 0 → (kk=0, o=j)
            struct AxTap { int kk, off; };
            auto axis = [&](int par, AxTap* out) -> int {
                if (k == 1) { if (par == 0) { out[0] = {0, 0}; return 1; } return 0; }
                if (par == 0) { out[0] = {1, 0}; return 1; }
                out[0] = {0, 1}; out[1] = {2, 0}; return 2;
            };
            for (int q = 0; q < 8; ++q) {
                AxTap ta[2], tb[2], tc[2];
                int na = axis(q >> 2, ta), nb = axis((q >> 1) & 1, tb), nc = axis(q & 1, tc);
                if (na * nb * nc == 0) continue;          // class receives no contribution: stays zero (caller zero-fills)
                Group& g = p.groups[p.n_groups];
                g.tap_begin = (int16_t)p.n_taps;
                g.out_view = (int16_t)p.n_out_views;
                p.out_views[p.n_out_views++] = parity_view(N, D, H, W, Cin, q >> 2, (q >> 1) & 1, q & 1);
                for (int a = 0; a < na; ++a) for (int b = 0; b < nb; ++b) for (int c = 0; c < nc; ++c)
                    add_tap(0, ta[a].off, tb[b].off, tc[c].off, (ta[a].kk * k + tb[b].kk) * k + tc[c].kk);
                g.tap_count = (int16_t)(p.n_taps - g.tap_begin);
                p.n_groups++;
            }
        }
    } else if (op == AMB_OP_CONVT) {
        // x (N,D,H,W,Cin) → y (N,2D,2H,2W,Cout): out index o = 2i - 1 + kk.  Per axis, o = 2j + par:
        //   par 0 → (kk=1, i=j), (kk=3, i=j-1) ; par 1 → (kk=0, i=j+1), (kk=2, i=j)
        p.Cx = Cin; p.Cy = Cout;
        p.oN = N; p.oD = D; p.oH = H; p.oW = W;
        p.n_in_views = 1; p.in_views[0] = full_view(N, D, H, W, Cin);
        const int kk_of[2][2] = {{1, 3}, {0, 2}}, off_of[2][2] = {{0, -1}, {1, 0}};
        for (int q = 0; q < 8; ++q) {
            const int pa = q >> 2, pb = (q >> 1) & 1, pc = q & 1;
            Group& g = p.groups[p.n_groups++];
            g.tap_begin = (int16_t)p.n_taps;
            g.out_view = (int16_t)q;
            p.out_views[q] = parity_view(N, 2 * D, 2 * H, 2 * W, Cout, pa, pb, pc);
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c)
                add_tap(0, off_of[pa][a], off_of[pb][b], off_of[pc][c],
                        (kk_of[pa][a] * 4 + kk_of[pb][b]) * 4 + kk_of[pc][c]);
            g.tap_count = 8;
        }
        p.n_out_views = 8;
    } else if (op == AMB_OP_CONVT_DGRAD) {
        // gathered = dy (N,2D,2H,2W,Cout) ; produced = dx (N,D,H,W,Cin): dx[i] = Σ_kk dy[2i - 1 + kk] W_kk
        //   kk=0 → (par 1, off -1) ; 1 → (par 0, 0) ; 2 → (par 1, 0) ; 3 → (par 0, +1)
        p.Cx = Cout; p.Cy = Cin;
        p.oN = N; p.oD = D; p.oH = H; p.oW = W;
        p.n_in_views = 8;
        for (int q = 0; q < 8; ++q) p.in_views[q] = parity_view(N, 2 * D, 2 * H, 2 * W, Cout, q >> 2, (q >> 1) & 1, q & 1);
        p.n_out_views = 1; p.out_views[0] = full_view(N, D, H, W, Cin);
        p.n_groups = 1; p.groups[0].tap_begin = 0; p.groups[0].out_view = 0;
        const int par_of[4] = {1, 0, 1, 0}, off_of[4] = {-1, 0, 0, 1};
        for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) for (int c = 0; c < 4; ++c)
            add_tap(par_of[a] * 4 + par_of[b] * 2 + par_of[c], off_of[a], off_of[b], off_of[c], (a * 4 + b) * 4 + c);
        p.groups[0].tap_count = 64;
    } else {
        AMB_CHECK(false, AMB_ERR_ARG, "unknown conv op %d", op);
    }
    return 0;
}

// mask granularity in out-view coordinates: the mask grid (fd,fh,fw) tiles the FULL produced tensor
static inline int plan_set_mask(Plan& p, int full_oD, int full_oH, int full_oW, int fd, int fh, int fw) {
    AMB_CHECK(fd > 0 && full_oD % fd == 0 && full_oH % fh == 0 && full_oW % fw == 0, AMB_ERR_ARG, "conv: bad mask grid");
    int P = full_oD / fd;
    AMB_CHECK(full_oH / fh == P && full_oW / fw == P, AMB_ERR_ARG, "conv: anisotropic patch edge");
    int Pv = P * p.oD / full_oD;     // parity views halve the edge
    AMB_CHECK(Pv >= 1 && (Pv & (Pv - 1)) == 0, AMB_ERR_ARG, "conv: patch edge %d (view %d) must be a power of two >= 1", P, Pv);
    p.lgPv = 0;
    while ((1 << p.lgPv) < Pv) p.lgPv++;
    p.fd = fd; p.fh = fh; p.fw = fw;
    return 0;
}

// implemented in conv_igemm.cu / conv_wgrad_tc.cu: return 1 when the tcgen05 path handled the call, 0 when the shape is
// not supported by it (caller falls back to the CUDA-core kernel), <0 on error.
int igemm_conv(const Plan& p, const amb_conv_args* a);
long igemm_workspace_bytes(const Plan& p, bool sparse_list);
int igemm3_conv(const Plan& p, const amb_conv_args* a);
int igemm4_conv(const Plan& p, const amb_conv_args* a);
int igemm4t_conv(const Plan& p, const amb_conv_args* a);
int igemm_wgrad(const Plan& p, const amb_wgrad_args* a);
int igemm_wgrad_halo(const Plan& p, const amb_wgrad_args* a);
int igemm_wgrad_ns(const Plan& p, const amb_wgrad_args* a);

}  // namespace amb
