// tcgen05 weight-gradient kernel for a gather-GEMM plan:
//     dW[tap.w][r][c] += Σ_o dY_view(tap)[o − off(tap), r] · X_view(tap)[o, c]
// i.e. per tap a GEMM with M = r (channels of dY), N = c (channels of X), K = voxels.  Both operands are read exactly as
// they lie in HBM (channels-last): a TMA box of 128 voxels × ≤64 channels lands in smem as [voxel][channel] rows, which
// is the *MN-major* canonical UMMA layout (K = voxel rows, 8-row groups SBO apart, 64/32/16-channel atoms LBO apart).
// The X box is loaded once per voxel tile and shared by all taps; the taps differ only in the shift of the dY box, so
// for narrow layers (Cy < 128) 128/Cy taps are stacked along M and one MMA computes several taps.
//
//   job      : (batch of ≤ 512/NTw accumulator units sharing one X view and N chunk) × (a K split of the voxel tiles)
//   roles    : warp 0 TMA producer (X ring ×2, dY ring ×4) · warp 1 MMA issuer · warp 2 TMEM alloc · warps 4-7 epilogue
//   epilogue : accumulators (TMEM) → fp32 atomics into dW (split-K reduction across CTAs)
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace amb {

using namespace ptx;

int encode_view_map(CUtensorMap* m, const void* tensor_base, const View& v, int C, int kc, const int box[4]);

#define WG_MAX_UNITS 64

struct WTap {
    int8_t aview, sz, sy, sx;     // dY view and the shift applied to the tile origin (= −tap offset)
    int16_t w;
    int16_t bview;
};
struct WBatch {
    int unit_begin, unit_count, bview;
};

struct WgradParams {
    CUtensorMap a_maps[8];        // dY views, box = slabW channels × 128 voxels
    CUtensorMap b_maps[8];        // X views,  box = nslabW channels × 128 voxels
    WTap taps[64];
    int8_t unit_taps[WG_MAX_UNITS][8];   // stacked mode: tap index per M slab (−1 = none)
    int view_unit_begin[9];       // units of X view v are [view_unit_begin[v], view_unit_begin[v+1])
    int view_batch_begin[9];      // batches never straddle an X view
    int per_batch;                // accumulator units per CTA = 512 / NTw
    int n_batches, n_units, n_views;
    int stacked;                  // 1: Cy < 128, units stack 128/Cy taps ; 0: unit = (tap, 128-row slab of Cy)
    int mslabs;                   // Cy / 128 when !stacked
    int Cx, Cy;
    int slabW, a_slabs;           // dY slab width (channels) and slabs per A slot (always 128 rows in total)
    int nslabW, b_slabs, NTw, n_nchunks;
    uint32_t a_slab_bytes, b_slab_bytes;
    int lgbn, lgbd, lgbh, lgbw, Tn, Tz, Ty, Tx;
    int oN, oD, oH, oW, lgPv, fd, fh, fw;
    const int* list;
    const int* count;
    int ksplit;
    float* dw;
    uint32_t idesc;
    uint32_t a_layout, b_layout, a_sbo, b_sbo, a_kstep, b_kstep;
    int a_slots;                  // dY ring depth (what fits beside the X ring)
    uint32_t b_slot_bytes, a_slot_bytes;
    int KV;                       // voxels per K tile (64 or 128): smaller tiles = deeper prefetch for the same smem
};

#define WG_A_SLOTS_MAX 12
#define WG_B_SLOTS 2
#define WG_G 4                       // accumulator units (independent MMA chains) interleaved per K tile

__device__ __forceinline__ long wg_num_ktiles(const WgradParams& P) {
    if (P.list) {
        const int Pv = 1 << P.lgPv;
        return (long)(*P.count) * (Pv >> P.lgbd) * (Pv >> P.lgbh) * (Pv >> P.lgbw);
    }
    return (long)P.Tn * P.Tz * P.Ty * P.Tx;
}

__device__ __forceinline__ void wg_decode(const WgradParams& P, long t, int& n0, int& z0, int& y0, int& x0) {
    if (P.list) {
        const int Pv = 1 << P.lgPv;
        const int sx = Pv >> P.lgbw, sy = Pv >> P.lgbh, sz = Pv >> P.lgbd;
        int ix = (int)(t % sx); t /= sx;
        int iy = (int)(t % sy); t /= sy;
        int iz = (int)(t % sz); t /= sz;
        int pid = P.list[t];
        const int L = P.fd * P.fh * P.fw;
        n0 = pid / L;
        int l = pid % L;
        z0 = ((l / (P.fh * P.fw)) << P.lgPv) + (iz << P.lgbd);
        y0 = (((l / P.fw) % P.fh) << P.lgPv) + (iy << P.lgbh);
        x0 = ((l % P.fw) << P.lgPv) + (ix << P.lgbw);
    } else {
        x0 = (int)(t % P.Tx) << P.lgbw; t /= P.Tx;
        y0 = (int)(t % P.Ty) << P.lgbh; t /= P.Ty;
        z0 = (int)(t % P.Tz) << P.lgbd; t /= P.Tz;
        n0 = (int)t << P.lgbn;
    }
}

__global__ void __launch_bounds__(256, 1) wgrad_kernel(const __grid_constant__ WgradParams P) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int A_SLOTS = P.a_slots;
    uint8_t* a_ring = smem;
    const uint32_t ASB = P.a_slot_bytes;
    uint8_t* b_ring = smem + A_SLOTS * ASB;
    uint8_t* ctrl = b_ring + WG_B_SLOTS * P.b_slot_bytes;
    uint64_t* a_full = (uint64_t*)ctrl;
    uint64_t* a_empty = a_full + WG_A_SLOTS_MAX;
    uint64_t* b_full = a_empty + WG_A_SLOTS_MAX;
    uint64_t* b_empty = b_full + WG_B_SLOTS;
    uint64_t* acc_full = b_empty + WG_B_SLOTS;
    uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
    // per-CTA table of the slab loads of each unit of this batch: {a-map index, shifts, first channel} (−1 = no load)
    int4* s_slab = (int4*)(ctrl + 256);                 // [<=32 units][8 slabs]

    if (threadIdx.x == 0) {
        for (int s = 0; s < A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < WG_B_SLOTS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // job decode: blockIdx.x = (batch, nchunk, ksplit)
    int job = blockIdx.x;
    const int ks = job % P.ksplit; job /= P.ksplit;
    const int nchunk = job % P.n_nchunks; job /= P.n_nchunks;
    WBatch B;
    {
        int v = 0;
        while (v + 1 < P.n_views && job >= P.view_batch_begin[v + 1]) ++v;
        B.bview = v;
        B.unit_begin = P.view_unit_begin[v] + (job - P.view_batch_begin[v]) * P.per_batch;
        const int left = P.view_unit_begin[v + 1] - B.unit_begin;
        B.unit_count = left < P.per_batch ? left : P.per_batch;
    }
    const long ktiles = wg_num_ktiles(P);
    const long k_begin = ktiles * ks / P.ksplit, k_end = ktiles * (ks + 1) / P.ksplit;

    for (int i = threadIdx.x; i < B.unit_count * 8; i += blockDim.x) {
        const int ul = i >> 3, j = i & 7, u = B.unit_begin + ul;
        int4 e = make_int4(-1, 0, 0, 0);
        if (j < P.a_slabs) {
            int ti, c0;
            if (P.stacked) { ti = P.unit_taps[u][j]; c0 = 0; }
            else { ti = u / P.mslabs; c0 = (u % P.mslabs) * 128 + j * P.slabW; }
            if (ti >= 0) {
                const WTap T = P.taps[ti];
                e = make_int4(T.aview, (T.sx & 0xFF) | ((T.sy & 0xFF) << 8) | ((T.sz & 0xFF) << 16), c0, 1);
            }
        }
        s_slab[i] = e;
    }
    __syncthreads();

    if (warp == 0 || warp == 2 || warp == 3) {
        // three producer warps (the per-unit TMA issue path was the bottleneck): warp 0 also feeds the X ring; unit i of
        // the CTA's stream goes to ring slot i % A_SLOTS, producer w takes units i ≡ w (mod 3)
        const int pw = warp == 0 ? 0 : warp - 1;             // 0, 1, 2
        uint32_t a_slot = (uint32_t)pw % (uint32_t)A_SLOTS, a_phase = 0, bi = 0;
        int ucur = pw;                                        // unit index inside the current K tile (may run past it)
        const uint32_t a_bytes1 = P.a_slab_bytes;
        for (long kt = k_begin; kt < k_end; ++kt, ++bi) {
            int n0, z0, y0, x0;
            wg_decode(P, kt, n0, z0, y0, x0);
            if (pw == 0) {
                const int bs = bi % WG_B_SLOTS;
                mbar_wait(&b_empty[bs], ((bi / WG_B_SLOTS) & 1) ^ 1, 11);
                if (elect_one()) {
                    mbar_expect_tx(&b_full[bs], P.b_slab_bytes * P.b_slabs);
                    for (int j = 0; j < P.b_slabs; ++j)
                        tma_load_5d(b_ring + bs * P.b_slot_bytes + j * P.b_slab_bytes, &P.b_maps[B.bview], &b_full[bs],
                                    nchunk * P.NTw + j * P.nslabW, x0, y0, z0, n0);
                }
                __syncwarp();
            }
            for (; ucur < B.unit_count; ucur += 3) {
                mbar_wait(&a_empty[a_slot], a_phase ^ 1u, 12);
                if (elect_one()) {
                    const int4* tab = s_slab + ucur * 8;
                    int real = 0;
                    for (int j = 0; j < P.a_slabs; ++j) real += tab[j].x >= 0;
                    mbar_expect_tx(&a_full[a_slot], a_bytes1 * (uint32_t)real);
                    uint8_t* dst = a_ring + a_slot * ASB;
                    for (int j = 0; j < P.a_slabs; ++j) {
                        const int4 e = tab[j];
                        if (e.x < 0) continue;
                        const int sx = (int)(int8_t)(e.y & 0xFF), sy = (int)(int8_t)((e.y >> 8) & 0xFF),
                                  sz = (int)(int8_t)((e.y >> 16) & 0xFF);
                        tma_load_5d(dst + j * a_bytes1, &P.a_maps[e.x], &a_full[a_slot], e.z, x0 + sx, y0 + sy, z0 + sz, n0);
                    }
                }
                __syncwarp();
                a_slot += 3;
                while (a_slot >= (uint32_t)A_SLOTS) { a_slot -= (uint32_t)A_SLOTS; a_phase ^= 1u; }
            }
            ucur -= B.unit_count;
        }
    } else if (warp == 1) {
        uint32_t a_slot = 0, a_phase = 0, bi = 0;
        const uint32_t a_kstep16 = P.a_kstep >> 4, b_kstep16 = P.b_kstep >> 4, NTw = (uint32_t)P.NTw, idesc = P.idesc;
        const uint32_t a_ring_u32 = smem_u32(a_ring);
        const uint64_t adesc0 = umma_desc(0, P.a_slab_bytes, P.a_sbo, P.a_layout);
        for (long kt = k_begin; kt < k_end; ++kt, ++bi) {
            const int bs = bi % WG_B_SLOTS;
            mbar_wait(&b_full[bs], (bi / WG_B_SLOTS) & 1, 13);
            tc_fence_after();
            const uint64_t bdesc = umma_desc(smem_u32(b_ring + bs * P.b_slot_bytes), P.b_slab_bytes, P.b_sbo, P.b_layout);
            const bool accum = kt != k_begin;
            // units are processed WG_G at a time: their accumulators are independent, so the 8-deep chains of dependent
            // tcgen05.mma interleave instead of each waiting out the accumulate latency
            for (int u = 0; u < B.unit_count; u += WG_G) {
                const int nu = (B.unit_count - u) < WG_G ? (B.unit_count - u) : WG_G;
                uint32_t slot[WG_G];
#pragma unroll
                for (int j = 0; j < WG_G; ++j) {
                    slot[j] = 0;
                    if (j < nu) {
                        slot[j] = a_slot;
                        mbar_wait(&a_full[a_slot], a_phase, 14);
                        if (++a_slot == (uint32_t)A_SLOTS) { a_slot = 0; a_phase ^= 1u; }
                    }
                }
                tc_fence_after();
                if (elect_one()) {
                    const int ksteps = P.KV >> 4;
                    for (int k = 0; k < ksteps; ++k) {      // KV voxels per tile = KV/16 × K16
#pragma unroll
                        for (int j = 0; j < WG_G; ++j) {
                            if (j < nu) {
                                const uint64_t ad = adesc0 + (uint64_t)(((a_ring_u32 + slot[j] * ASB) & 0x3FFFFu) >> 4) +
                                                    (uint64_t)(a_kstep16 * k);
                                mma_bf16(tmem_base + (uint32_t)(u + j) * NTw, ad, bdesc + (uint64_t)(b_kstep16 * k), idesc,
                                         accum || (k != 0));
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < WG_G; ++j)
                        if (j < nu) mma_commit(&a_empty[slot[j]]);
                }
                __syncwarp();
            }
            if (elect_one()) mma_commit(&b_empty[bs]);
            __syncwarp();
        }
        if (elect_one()) mma_commit(acc_full);
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int m = q * 32 + lane;
        mbar_wait(acc_full, 0, 15);
        tc_fence_after();
        if (k_end > k_begin) {
            for (int u = 0; u < B.unit_count; ++u) {
                const int ug = B.unit_begin + u;
                int ti, r;
                if (P.stacked) {
                    ti = P.unit_taps[ug][m / P.slabW];
                    r = m % P.slabW;
                } else {
                    ti = ug / P.mslabs;
                    r = (ug % P.mslabs) * 128 + m;
                }
                float* dst = nullptr;
                if (ti >= 0) dst = P.dw + ((long)P.taps[ti].w * P.Cy + r) * P.Cx + nchunk * P.NTw;
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * P.NTw);
                for (int col = 0; col < P.NTw; col += 16) {
                    uint32_t rr[16];
                    tmem_ld_x16(t_addr + col, rr);
                    tmem_ld_wait();
                    if (dst) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)       // 128-bit reductions: 4x fewer L2 atomic transactions
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + col + j),
                                         "f"(__uint_as_float(rr[j])), "f"(__uint_as_float(rr[j + 1])),
                                         "f"(__uint_as_float(rr[j + 2])), "f"(__uint_as_float(rr[j + 3]))
                                         : "memory");
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int ilog2w(int v) {
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}
static int pow2_ceilw(int v) { return 1 << ilog2w(v); }
static bool chan_ok(int c) { return c == 16 || c == 32 || c == 64 || c % 128 == 0; }

int igemm_wgrad(const Plan& p, const amb_wgrad_args* a) {
    // dense 3x3x3 s1 with narrow dY: halo-plane kernel (conv_wgrad_halo.cu)
    if (int r = igemm_wgrad_ns(p, a)) return r;          // dz taps stacked along N: 64->64 / 64->32 dense 3x3x3
    if (int r = igemm_wgrad_halo(p, a)) return r;
    // roles here: dY has p.Cy channels (M), X has p.Cx channels (N)
    if (!chan_ok(p.Cx) || !chan_ok(p.Cy)) {
        set_error("wgrad: channels (%d,%d) must be 16, 32, 64 or a multiple of 128", p.Cx, p.Cy);
        return 0;
    }
    static WgradParams P;
    memset(&P, 0, sizeof(P));
    P.Cx = p.Cx; P.Cy = p.Cy;
    P.stacked = p.Cy < 128;
    P.slabW = p.Cy < 64 ? p.Cy : 64;
    P.a_slabs = 128 / P.slabW;
    P.mslabs = P.stacked ? 1 : p.Cy / 128;
    P.NTw = p.Cx < 128 ? p.Cx : 128;
    P.nslabW = P.NTw < 64 ? P.NTw : 64;
    P.b_slabs = P.NTw / P.nslabW;
    P.n_nchunks = p.Cx / P.NTw;
    const char* kvenv = getenv("AMB_WG_KV");
    P.KV = (kvenv && atoi(kvenv) == 64) ? 64 : 128;      // measured: 64-voxel tiles halve throughput (per-unit issue cost)
    P.a_slab_bytes = (uint32_t)P.KV * P.slabW * 2u;
    P.b_slab_bytes = (uint32_t)P.KV * P.nslabW * 2u;
    auto layout_of = [](int w) { return w == 64 ? 2u : (w == 32 ? 4u : 6u); };
    P.a_layout = layout_of(P.slabW); P.b_layout = layout_of(P.nslabW);
    P.a_sbo = 8u * P.slabW * 2u; P.b_sbo = 8u * P.nslabW * 2u;       // 8 voxel rows
    P.a_kstep = 16u * P.slabW * 2u; P.b_kstep = 16u * P.nslabW * 2u; // 16 voxel rows per MMA
    P.idesc = umma_idesc_bf16(128, P.NTw, 1, 1);
    P.oN = p.oN; P.oD = p.oD; P.oH = p.oH; P.oW = p.oW;
    P.lgPv = p.lgPv; P.fd = p.fd; P.fh = p.fh; P.fw = p.fw;
    P.dw = a->dw;

    int bw = pow2_ceilw(p.oW) < 8 ? pow2_ceilw(p.oW) : 8;
    int bh = pow2_ceilw(p.oH) < 8 ? pow2_ceilw(p.oH) : 8;
    int rem = P.KV / (bw * bh);
    if (rem < 1) rem = 1;
    int bd = pow2_ceilw(p.oD) < rem ? pow2_ceilw(p.oD) : rem;
    int bn = rem / bd;
    const bool use_list = a->active_list != nullptr && p.lgPv >= 0 && (1 << p.lgPv) >= 8 && bn == 1 && bw == 8 &&
                          bh == 8 && bd <= 2;
    P.list = use_list ? a->active_list : nullptr;
    P.count = use_list ? a->active_count : nullptr;
    P.lgbn = ilog2w(bn); P.lgbd = ilog2w(bd); P.lgbh = ilog2w(bh); P.lgbw = ilog2w(bw);
    P.Tn = ceil_div(p.oN, bn); P.Tz = ceil_div(p.oD, bd); P.Ty = ceil_div(p.oH, bh); P.Tx = ceil_div(p.oW, bw);

    // taps sorted by X view (plan in_view); the dY view of a tap is the out view of its group
    int order[64], n = 0;
    for (int v = 0; v < p.n_in_views; ++v)
        for (int t = 0; t < p.n_taps; ++t)
            if (p.taps[t].view == v) order[n++] = t;
    for (int i = 0; i < n; ++i) {
        const Tap& T = p.taps[order[i]];
        int g = 0;
        for (int j = 0; j < p.n_groups; ++j)
            if (order[i] >= p.groups[j].tap_begin && order[i] < p.groups[j].tap_begin + p.groups[j].tap_count) g = j;
        WTap& w = P.taps[i];
        w.aview = (int8_t)p.groups[g].out_view;
        w.sz = (int8_t)-T.dz; w.sy = (int8_t)-T.dy; w.sx = (int8_t)-T.dx;
        w.w = T.w;
        w.bview = T.view;
    }
    // units (per X view, contiguous) and batches (chunks of per_batch units inside a view)
    const int per_batch = 512 / P.NTw;
    P.per_batch = per_batch;
    P.n_views = p.n_in_views;
    memset(P.unit_taps, -1, sizeof(P.unit_taps));
    int nu = 0, nb = 0, i = 0;
    for (int v = 0; v < p.n_in_views; ++v) {
        P.view_unit_begin[v] = nu;
        P.view_batch_begin[v] = nb;
        if (P.stacked) {
            while (i < n && P.taps[i].bview == v) {
                if (nu >= WG_MAX_UNITS) { set_error("wgrad: too many units"); return 0; }
                for (int j = 0; j < P.a_slabs && i < n && P.taps[i].bview == v; ++j) P.unit_taps[nu][j] = (int8_t)i++;
                nu++;
            }
        } else {
            // unit index = tap * mslabs + mslab (taps are sorted by view, so a view's units are contiguous)
            while (i < n && P.taps[i].bview == v) { ++i; nu += P.mslabs; }
        }
        nb += (nu - P.view_unit_begin[v] + per_batch - 1) / per_batch;
    }
    P.view_unit_begin[p.n_in_views] = nu;
    P.view_batch_begin[p.n_in_views] = nb;
    P.n_units = nu; P.n_batches = nb;

    const int box[4] = {bn, bd, bh, bw};
    for (int i = 0; i < p.n_out_views; ++i)
        if (int e = encode_view_map(&P.a_maps[i], a->dy, p.out_views[i], p.Cy, P.slabW, box)) return e;
    for (int i = 0; i < p.n_in_views; ++i)
        if (int e = encode_view_map(&P.b_maps[i], a->x, p.in_views[i], p.Cx, P.nslabW, box)) return e;

    long ktiles_upper = (long)P.Tn * P.Tz * P.Ty * P.Tx;
    int base_jobs = nb * P.n_nchunks;
    // split-K: enough CTAs to fill the chip, but every extra split costs a full set of fp32 reductions in the epilogue:
    // two waves only when each CTA still gets a long K range
    const char* wenv = getenv("AMB_WG_WAVES");
    int waves = wenv ? atoi(wenv) : (ktiles_upper * base_jobs >= 64L * num_sms() ? 2 : 1);
    int ksplit = (waves * num_sms() + base_jobs - 1) / base_jobs;
    if (ksplit > ktiles_upper) ksplit = (int)ktiles_upper;
    if (ksplit < 1) ksplit = 1;
    P.ksplit = ksplit;
    P.b_slot_bytes = (P.b_slab_bytes * P.b_slabs + 1023u) & ~1023u;
    P.a_slot_bytes = (uint32_t)P.KV * 256u;           // 128 M rows x KV voxels x 2 B
    P.a_slots = (int)((227u * 1024u - 1280u - 4096u - WG_B_SLOTS * P.b_slot_bytes) / P.a_slot_bytes);
    if (P.a_slots > WG_A_SLOTS_MAX) P.a_slots = WG_A_SLOTS_MAX;
    size_t smem = (size_t)P.a_slots * P.a_slot_bytes + (size_t)WG_B_SLOTS * P.b_slot_bytes + 1024 + 256 + 32 * 8 * 16;
    AMB_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_kernel<<<base_jobs * ksplit, 256, smem, (cudaStream_t)a->stream>>>(P);
    AMB_LAUNCH_CHECK();
    g_last_conv_kernel = "wgrad_kernel";
    return 1;
}

}  // namespace amb
