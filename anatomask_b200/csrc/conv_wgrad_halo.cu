// tcgen05 weight-gradient kernel, halo-plane variant: dense 3×3×3 stride-1 convolutions whose dY is narrow (Cy = 32/64)
// — LightDecoder's last two blocks at 64³/128³, where the per-tap kernel (conv_wgrad_tc.cu) re-reads dY 27× through TMA
// and is bound by the per-unit TMA issue path (measured: tensor pipe 26 %, 2.9 GB DRAM reads for 1.07 GB of operands).
//
//     dW[tap][r][c] = Σ_o dY[o − off(tap), r] · X[o, c]          M = r (dY channels), N = c (X channels), K = voxels
//
// A CTA walks z along a column of 8×8-voxel tiles.  Per z step it needs one X box [8 y][8 x][Cx] (B operand, MN-major)
// and the three dY halo planes z−1, z, z+1, each ONE TMA box [10 y][16 x][Cy] that stays in a ring while z advances:
// every plane is loaded once per column instead of 27 times per tile.  All 9 in-plane taps read the plane in place —
// the A descriptor starts ((sy+1)·16 + (sx+1)) voxel rows into it (SBO = one y row of the plane; the swizzle phase
// follows the absolute smem address, as in conv_igemm3.cu).  Taps are stacked along M through the descriptor's LBO:
//   Cy = 64 : two taps per unit (LBO = byte distance between the two tap origins inside the plane)
//   Cy = 32 : the three sx taps of one (sz, sy) per unit (LBO = one voxel row = 64 B; the 4th atom is discarded)
//
//   warp 0  dY plane producer     warps 1, 2  MMA issuers (warp 2 also owns the TMEM allocation)     warp 3  X box producer
//   warps 4-7 epilogue: accumulators (TMEM) → fp32 red.v4 into dW (split-K reduction across CTAs)
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace amb {

using namespace ptx;

int encode_view_map(CUtensorMap* m, const void* tensor_base, const View& v, int C, int kc, const int box[4]);

#define WH_MAX_UNITS 64
#define WH_MAX_BATCHES 32
#define WH_UNITS_CTA 8               // accumulator units per CTA (unrolled issue loop)
#define WH_A_SLOTS_MAX 8
#define WH_B_SLOTS 4

struct WhUnit {
    int16_t tap[4];                  // weight slab index per M atom (−1 = atom discarded)
    int32_t off;                     // byte offset of atom 0 inside its plane
    int32_t lbo;                     // bytes between consecutive atoms
    int32_t dzslot;                  // 0, 1, 2 : plane z−1, z, z+1 of the step
};

struct WgradHaloParams {
    CUtensorMap a_maps[8];           // dY views (1 for a convolution, the 8 parity classes of a ConvTranspose), box
                                     // (slabW channels, 16 x, 10 y, 1, 1)
    CUtensorMap b_map;               // X,  box (nslabW channels, 8 x, 8 y, 1, 1)
    WhUnit units[WH_MAX_UNITS];
    int n_units, n_batches;
    int16_t batch_unit_begin[WH_MAX_BATCHES], batch_unit_count[WH_MAX_BATCHES];   // a batch = the units of one CTA
    int8_t batch_class[WH_MAX_BATCHES];                                          // dY view of the batch's units
    int Cx, Cy, slabW, NTw, nslabW, b_slabs, n_nchunks;
    uint32_t plane_bytes, b_slab_bytes, b_slot_bytes;
    uint32_t a_layout, b_layout, a_sbo, b_sbo, a_kstep, b_kstep, idesc;
    int a_slots, issuers;
    int wide, mslabs, a_slabs, b_slots;   // wide: Cy >= 128, a unit is one tap x one 128-channel slab of dY (two 64-channel atoms)
    uint32_t plane_slab_bytes;
    int oN, oD, oH, oW, Ty, Tx;
    uint32_t n_steps;                // columns × oD (dense walk)
    int ksplit;
    const int* list;                 // active-patch work-list: columns are enumerated inside visible patches only and a
    const int* count;                // column is one patch deep (2^lgPv z steps + 2 halo planes); nullptr = dense
    int lgPv, fd, fh, fw;
    float* dw;
};

struct WhCol {
    int n0, y0, x0, z0;
};
__device__ __forceinline__ WhCol wh_column(const WgradHaloParams& P, uint32_t col) {
    WhCol c;
    if (P.list) {
        const uint32_t lt = (uint32_t)P.lgPv - 3u;                 // log2 of 8-voxel tiles per patch edge
        const uint32_t tx = col & ((1u << lt) - 1u), ty = (col >> lt) & ((1u << lt) - 1u);
        const uint32_t pid = (uint32_t)P.list[col >> (2u * lt)];
        const uint32_t L = (uint32_t)(P.fd * P.fh * P.fw), hw = (uint32_t)(P.fh * P.fw);
        const uint32_t n = pid / L, l = pid - n * L;
        const uint32_t pz = l / hw, r2 = l - pz * hw;
        const uint32_t py = r2 / (uint32_t)P.fw, px = r2 - py * (uint32_t)P.fw;
        c.n0 = (int)n;
        c.z0 = (int)(pz << P.lgPv);
        c.y0 = (int)((py << P.lgPv) + ty * 8u);
        c.x0 = (int)((px << P.lgPv) + tx * 8u);
        return c;
    }
    c.z0 = 0;
    c.x0 = (int)(col % (uint32_t)P.Tx) * 8; col /= (uint32_t)P.Tx;
    c.y0 = (int)(col % (uint32_t)P.Ty) * 8;
    c.n0 = (int)(col / (uint32_t)P.Ty);
    return c;
}

__global__ void __launch_bounds__(256, 1) wgrad_halo_kernel(const __grid_constant__ WgradHaloParams P) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t A_SLOTS = (uint32_t)P.a_slots;
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + A_SLOTS * P.plane_bytes;
    const uint32_t B_SLOTS = (uint32_t)P.b_slots;
    uint8_t* ctrl = b_ring + B_SLOTS * P.b_slot_bytes;
    uint64_t* a_full = (uint64_t*)ctrl;
    uint64_t* a_empty = a_full + WH_A_SLOTS_MAX;
    uint64_t* b_full = a_empty + WH_A_SLOTS_MAX;
    uint64_t* b_empty = b_full + WH_B_SLOTS;
    uint64_t* acc_full = b_empty + WH_B_SLOTS;
    uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

    if (threadIdx.x == 0) {
        const uint32_t nI = (uint32_t)P.issuers;                 // every issuing warp commits to the "empty" barriers
        for (uint32_t s = 0; s < A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], nI); }
        for (int s = 0; s < WH_B_SLOTS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], nI); }
        mbar_init(acc_full, nI);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // job decode: blockIdx.x = (batch, nchunk, ksplit)
    uint32_t job = blockIdx.x;
    const uint32_t ks = job % (uint32_t)P.ksplit; job /= (uint32_t)P.ksplit;
    const int nchunk = (int)(job % (uint32_t)P.n_nchunks); job /= (uint32_t)P.n_nchunks;
    const int mslab = (int)(job / (uint32_t)P.n_batches);
    job %= (uint32_t)P.n_batches;
    const int unit_begin = P.batch_unit_begin[job], unit_count = P.batch_unit_count[job];
    const CUtensorMap* a_map = &P.a_maps[P.batch_class[job]];
    // oD below = z steps per column: the whole depth when dense, one patch edge with the work-list
    const uint32_t n_steps = P.list ? ((uint32_t)(*P.count) << (3 * P.lgPv - 6)) : P.n_steps;
    const uint32_t s_begin = (uint32_t)((unsigned long long)n_steps * ks / (uint32_t)P.ksplit);
    const uint32_t s_end = (uint32_t)((unsigned long long)n_steps * (ks + 1) / (uint32_t)P.ksplit);
    const uint32_t oD = P.list ? (1u << P.lgPv) : (uint32_t)P.oD;

    if (warp == 0) {
        // dY planes: a segment [z, zend) of one column needs planes z−1 … zend (rows / planes outside the tensor are
        // zero-filled by TMA = the convolution's zero padding)
        uint32_t slot = 0, phase = 0, s = s_begin;
        while (s < s_end) {
            const uint32_t col = s / oD;
            const int z = (int)(s - col * oD);
            const uint32_t rem = s_end - s;
            const int zend = (oD - (uint32_t)z) < rem ? (int)oD : z + (int)rem;
            const WhCol c = wh_column(P, col);
            for (int zp = z - 1; zp <= zend; ++zp) {
                mbar_wait(&a_empty[slot], phase ^ 1u, 21);
                if (elect_one()) {
                    mbar_expect_tx(&a_full[slot], P.plane_bytes);
                    for (int j = 0; j < P.a_slabs; ++j)
                        tma_load_5d(a_ring + slot * P.plane_bytes + j * P.plane_slab_bytes, a_map, &a_full[slot],
                                    mslab * 128 + j * 64, c.x0 - 1, c.y0 - 1, c.z0 + zp, c.n0);
                }
                __syncwarp();
                if (++slot == A_SLOTS) { slot = 0; phase ^= 1u; }
            }
            s += (uint32_t)(zend - z);
        }
    } else if (warp == 3) {
        uint32_t slot = 0, phase = 0, s = s_begin;
        while (s < s_end) {
            const uint32_t col = s / oD;
            const int z = (int)(s - col * oD);
            const uint32_t rem = s_end - s;
            const int zend = (oD - (uint32_t)z) < rem ? (int)oD : z + (int)rem;
            const WhCol c = wh_column(P, col);
            for (int zp = z; zp < zend; ++zp) {
                mbar_wait(&b_empty[slot], phase ^ 1u, 22);
                if (elect_one()) {
                    mbar_expect_tx(&b_full[slot], P.b_slab_bytes * (uint32_t)P.b_slabs);
                    for (int j = 0; j < P.b_slabs; ++j)
                        tma_load_5d(b_ring + slot * P.b_slot_bytes + j * P.b_slab_bytes, &P.b_map, &b_full[slot],
                                    nchunk * P.NTw + j * P.nslabW, c.x0, c.y0, c.z0 + zp, c.n0);
                }
                __syncwarp();
                if (++slot == B_SLOTS) { slot = 0; phase ^= 1u; }
            }
            s += (uint32_t)(zend - z);
        }
    } else if (warp == 1 || (warp == 2 && P.issuers == 2)) {
        // MMA issuers.  One elected thread sustains roughly one tcgen05.mma per ~45 ns (descriptor arithmetic + R2UR moves on
        // a single warp without ILP), which is slower than the tensor pipe at N <= 64 — so the CTA's units are split across
        // TWO issuing warps (independent accumulators; every ring barrier then counts both warps' commits).
        const int iw = warp == 1 ? 0 : 1, nI = P.issuers;
        uint32_t a_lbo[WH_UNITS_CTA], a_off[WH_UNITS_CTA], a_dz[WH_UNITS_CTA];
#pragma unroll
        for (int j = 0; j < WH_UNITS_CTA; ++j) {
            const int u = j * nI + iw;
            const WhUnit& U = P.units[unit_begin + (u < unit_count ? u : 0)];
            a_lbo[j] = (((uint32_t)U.lbo >> 4) & 0x3FFFu) << 16;
            a_off[j] = (uint32_t)U.off >> 4;
            a_dz[j] = (uint32_t)U.dzslot;
        }
        const int my_units = (unit_count - iw + nI - 1) / nI;      // units iw, iw + nI, ...
        const uint32_t a_desc_hi = (uint32_t)(umma_desc(0, 0, P.a_sbo, P.a_layout) >> 32);
        const uint32_t a_kstep16 = P.a_kstep >> 4, b_kstep16 = P.b_kstep >> 4, idesc = P.idesc;
        const uint32_t d_stride = (uint32_t)(P.NTw * nI), d_base = tmem_base + (uint32_t)(iw * P.NTw);
        const uint32_t a_ring_u32 = smem_u32(a_ring), b_ring_u32 = smem_u32(b_ring), PB = P.plane_bytes;
        const uint64_t b_hi = umma_desc(0, P.b_slab_bytes, P.b_sbo, P.b_layout);
        uint32_t ws = 0, wph = 0;          // next plane to wait for
        uint32_t s0 = 0;                   // ring slot of plane z−1 of the current step
        uint32_t bs = 0, bph = 0;
        bool accum = false;
        uint32_t s = s_begin;
        while (s < s_end) {
            const uint32_t col = s / oD;
            const uint32_t z = s - col * oD;
            const uint32_t rem = s_end - s;
            const uint32_t len = (oD - z) < rem ? (oD - z) : rem;
            for (int i = 0; i < 2; ++i) {
                mbar_wait(&a_full[ws], wph, 23);
                if (++ws == A_SLOTS) { ws = 0; wph ^= 1u; }
            }
            for (uint32_t j = 0; j < len; ++j) {
                mbar_wait(&a_full[ws], wph, 24);
                if (++ws == A_SLOTS) { ws = 0; wph ^= 1u; }
                mbar_wait(&b_full[bs], bph, 25);
                tc_fence_after();
                const uint32_t s1 = (s0 + 1 == A_SLOTS) ? 0u : s0 + 1;
                const uint32_t s2 = (s1 + 1 == A_SLOTS) ? 0u : s1 + 1;
                if (elect_one()) {
                    const uint32_t p0 = ((a_ring_u32 + s0 * PB) & 0x3FFFFu) >> 4, p1 = ((a_ring_u32 + s1 * PB) & 0x3FFFFu) >> 4,
                                   p2 = ((a_ring_u32 + s2 * PB) & 0x3FFFFu) >> 4;
                    const uint64_t bdesc = b_hi | (uint64_t)(((b_ring_u32 + bs * P.b_slot_bytes) & 0x3FFFFu) >> 4);
                    uint32_t a_lo[WH_UNITS_CTA];
#pragma unroll
                    for (int u = 0; u < WH_UNITS_CTA; ++u)
                        a_lo[u] = a_lbo[u] | ((a_dz[u] == 0 ? p0 : (a_dz[u] == 1 ? p1 : p2)) + a_off[u]);
                    // 64 voxels = 4 × K16, k outermost: all units of this issuer interleave, so the dependent MMAs on one
                    // accumulator are my_units issue slots apart (measured: a lone trailing unit serialises)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int u = 0; u < WH_UNITS_CTA; ++u) {
                            if (u < my_units) {
                                const uint64_t ad = ((uint64_t)a_desc_hi << 32) | (uint64_t)(a_lo[u] + a_kstep16 * (uint32_t)k);
                                mma_bf16(d_base + (uint32_t)u * d_stride, ad, bdesc + (uint64_t)(b_kstep16 * (uint32_t)k), idesc,
                                         accum || (k != 0));
                            }
                        }
                    }
                    mma_commit(&a_empty[s0]);
                    mma_commit(&b_empty[bs]);
                }
                __syncwarp();
                accum = true;
                s0 = s1;
                if (++bs == B_SLOTS) { bs = 0; bph ^= 1u; }
            }
            // the two trailing planes of the segment
            const uint32_t t1 = (s0 + 1 == A_SLOTS) ? 0u : s0 + 1;
            if (elect_one()) {
                mma_commit(&a_empty[s0]);
                mma_commit(&a_empty[t1]);
            }
            __syncwarp();
            s0 = (t1 + 1 == A_SLOTS) ? 0u : t1 + 1;
            s += len;
        }
        if (elect_one()) mma_commit(acc_full);
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int m = q * 32 + lane;
        mbar_wait(acc_full, 0, 26);
        tc_fence_after();
        if (s_end > s_begin) {
            const int atom = P.wide ? 0 : m / P.slabW, r = P.wide ? mslab * 128 + m : m % P.slabW;
            for (int u = 0; u < unit_count; ++u) {
                const int w = P.units[unit_begin + u].tap[atom];
                float* dst = nullptr;
                if (w >= 0) dst = P.dw + ((long)w * P.Cy + r) * P.Cx + nchunk * P.NTw;
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * P.NTw);
                for (int col = 0; col < P.NTw; col += 16) {
                    uint32_t rr[16];
                    tmem_ld_x16(t_addr + col, rr);
                    tmem_ld_wait();
                    if (dst) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
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

// returns 1 when handled, 0 when the shape is outside this kernel's scope, <0 on error
int igemm_wgrad_halo(const Plan& p, const amb_wgrad_args* a) {
    if (getenv("AMB_DISABLE_WH")) return 0;
    // sparse layers: the kernel walks the active-patch list (a patch must hold whole 8x8 columns: patch edge >= 8).  Round 1
    // kept this opt-in (no step gain while the encoder weight gradients hid on the side stream); with the step's device time
    // now back to back on both streams it is worth 0.27 ms of a 21 ms step (20.97 -> 20.70, same box), so it is the default.
    // AMB_WH_NO_LIST=1 sends the sparse layers back to the per-tap kernel.
    const bool use_list = a->active_list != nullptr && p.lgPv >= 3 && getenv("AMB_WH_NO_LIST") == nullptr;
    if (a->active_list != nullptr && !use_list) return 0;
    // a convolution (27 taps, one dY view) or a ConvTranspose k4 s2 (8 parity-class views of dY with 8 taps each)
    const bool is_conv = p.n_out_views == 1 && p.n_taps == 27;
    const bool is_convT = a->op == AMB_OP_CONVT && p.n_out_views == 8 && p.n_taps == 64 && !getenv("AMB_WH_NO_CONVT");
    if (p.n_in_views != 1 || !(is_conv || is_convT)) return 0;
    for (int i = 0; i < p.n_taps; ++i) {
        const Tap& T = p.taps[i];
        if (T.dz < -1 || T.dz > 1 || T.dy < -1 || T.dy > 1 || T.dx < -1 || T.dx > 1) return 0;
    }
    if (is_convT && p.Cy < 64) return 0;
    const bool wide = p.Cy >= 128 && !getenv("AMB_WH_NO_WIDE");
    if (p.Cy != 32 && p.Cy != 64 && !(wide && p.Cy % 128 == 0)) return 0;
    if (!(p.Cx == 32 || p.Cx == 64 || p.Cx % 128 == 0)) return 0;
    if (p.oH % 8 != 0 || p.oW % 8 != 0 || p.oH < 16 || p.oW < 16) return 0;
    const View& vi = p.in_views[0];
    if (vi.D != p.oD || vi.H != p.oH || vi.W != p.oW) return 0;   // stride 1 only

    static WgradHaloParams P;
    memset(&P, 0, sizeof(P));
    P.Cx = p.Cx; P.Cy = p.Cy;
    P.slabW = wide ? 64 : p.Cy;                                    // one atom = all dY channels of one tap (64 when wide)
    P.wide = wide ? 1 : 0;
    P.mslabs = wide ? p.Cy / 128 : 1;
    P.a_slabs = wide ? 2 : 1;
    P.b_slots = wide ? 2 : WH_B_SLOTS;                             // wide planes are 40 KB: trade X ring depth for a 5th plane
    // wide layers: 128-column accumulators (4 units per CTA).  Measured against 64 columns / 8 units per CTA: 1141 vs 908
    // TFLOP/s on 128->128 @64^3 — the N = 128 MMA halves the operand reads per FLOP, which matters more than plane traffic
    const char* ntenv = getenv("AMB_WH_WIDE_NT");
    const int nt_cap = (wide && ntenv && atoi(ntenv) == 64) ? 64 : 128;
    P.NTw = p.Cx < nt_cap ? p.Cx : nt_cap;
    P.nslabW = P.NTw < 64 ? P.NTw : 64;
    P.b_slabs = P.NTw / P.nslabW;
    P.n_nchunks = p.Cx / P.NTw;
    const uint32_t a_row = (uint32_t)P.slabW * 2u, b_row = (uint32_t)P.nslabW * 2u;
    P.plane_slab_bytes = 10u * 16u * a_row;                        // 20 KB (64 channels) / 10 KB (32), 1024-aligned
    P.plane_bytes = P.plane_slab_bytes * (uint32_t)P.a_slabs;
    P.b_slab_bytes = 64u * b_row;
    P.b_slot_bytes = P.b_slab_bytes * (uint32_t)P.b_slabs;
    auto layout_of = [](int w) { return w == 64 ? 2u : (w == 32 ? 4u : 6u); };
    P.a_layout = layout_of(P.slabW); P.b_layout = layout_of(P.nslabW);
    P.a_sbo = 16u * a_row;                                         // next 8-voxel group = next y row of the plane
    P.a_kstep = 2u * P.a_sbo;                                      // K16 = two y rows
    P.b_sbo = 8u * b_row;
    P.b_kstep = 16u * b_row;
    P.idesc = umma_idesc_bf16(128, P.NTw, 1, 1);
    P.oN = p.oN; P.oD = p.oD; P.oH = p.oH; P.oW = p.oW;
    P.Ty = p.oH / 8; P.Tx = p.oW / 8;
    P.dw = a->dw;
    P.list = use_list ? a->active_list : nullptr;
    P.count = use_list ? a->active_count : nullptr;
    P.lgPv = p.lgPv; P.fd = p.fd; P.fh = p.fh; P.fw = p.fw;
    const char* ienv = getenv("AMB_WH_ISSUERS");
    P.issuers = (ienv && atoi(ienv) == 1) ? 1 : 2;

    // units: taps of one dY view and plane (same sz) stacked along M; batches = the units of one CTA, never mixing views
    const bool single = getenv("AMB_WH_SINGLE") != nullptr;       // debugging aid: one tap per unit
    int per_batch = 512 / P.NTw;
    if (per_batch > WH_UNITS_CTA) per_batch = WH_UNITS_CTA;
    int nu = 0, nb = 0;
    for (int g = 0; g < p.n_groups; ++g) {
        const Group& G = p.groups[g];
        const int cls_unit_begin = nu;
        for (int sz = -1; sz <= 1; ++sz) {
            // taps of this view and plane sorted by their byte offset in the plane
            int idx[9], offs[9], n = 0;
            for (int sy = -1; sy <= 1; ++sy)
                for (int sx = -1; sx <= 1; ++sx)
                    for (int t = G.tap_begin; t < G.tap_begin + G.tap_count; ++t)
                        if (-p.taps[t].dz == sz && -p.taps[t].dy == sy && -p.taps[t].dx == sx) {
                            idx[n] = t;
                            offs[n] = ((sy + 1) * 16 + (sx + 1)) * (int)a_row;
                            n++;
                        }
            const int per_unit = (single || wide) ? 1 : (p.Cy == 64 ? 2 : 3);
            if (per_unit == 3 && n != 9) return 0;
            for (int i = 0; i < n; i += per_unit) {
                if (nu >= WH_MAX_UNITS) { set_error("wgrad halo: too many units"); return 0; }
                WhUnit& U = P.units[nu++];
                for (int j = 0; j < 4; ++j) U.tap[j] = -1;
                U.off = offs[i];
                U.dzslot = sz + 1;
                U.lbo = (int)a_row;                                    // default: next voxel row (discarded atoms)
                if (wide) U.lbo = (int)P.plane_slab_bytes;             // second atom = channels 64..127 of the same tap
                if (per_unit >= 2 && i + 1 < n) U.lbo = offs[i + 1] - offs[i];
                for (int j = 0; j < per_unit && i + j < n; ++j) {
                    if (j >= 2 && offs[i + j] - offs[i + j - 1] != U.lbo) { set_error("wgrad halo: taps not equidistant"); return -1; }
                    U.tap[j] = p.taps[idx[i + j]].w;
                }
            }
        }
        // balanced batches over this view's units
        const int cu = nu - cls_unit_begin;
        if (cu == 0) continue;
        const int cb = (cu + per_batch - 1) / per_batch, upb = (cu + cb - 1) / cb;
        for (int u0 = 0; u0 < cu; u0 += upb) {
            if (nb >= WH_MAX_BATCHES) { set_error("wgrad halo: too many batches"); return 0; }
            P.batch_unit_begin[nb] = (int16_t)(cls_unit_begin + u0);
            P.batch_unit_count[nb] = (int16_t)((cu - u0) < upb ? (cu - u0) : upb);
            P.batch_class[nb] = (int8_t)G.out_view;
            nb++;
        }
    }
    P.n_units = nu;
    P.n_batches = nb;

    const int abox[4] = {1, 1, 10, 16}, bbox[4] = {1, 1, 8, 8};
    for (int v = 0; v < p.n_out_views; ++v)
        if (int e = encode_view_map(&P.a_maps[v], a->dy, p.out_views[v], p.Cy, P.slabW, abox)) return e;
    if (int e = encode_view_map(&P.b_map, a->x, p.in_views[0], p.Cx, P.nslabW, bbox)) return e;

    const long steps = use_list ? ((long)p.oN * p.fd * p.fh * p.fw << (3 * p.lgPv - 6)) : (long)p.oN * P.Ty * P.Tx * p.oD;
    if (steps >= (1L << 31)) return 0;
    P.n_steps = (uint32_t)steps;
    const int base_jobs = P.mslabs * P.n_batches * P.n_nchunks;
    int ksplit = num_sms() / base_jobs;
    if (ksplit < 1) ksplit = 1;
    if ((long)ksplit > steps) ksplit = (int)steps;
    P.ksplit = ksplit;
    const uint32_t budget = 227u * 1024u - 1024u - 512u - (uint32_t)P.b_slots * P.b_slot_bytes;
    P.a_slots = (int)(budget / P.plane_bytes);
    if (P.a_slots > WH_A_SLOTS_MAX) P.a_slots = WH_A_SLOTS_MAX;
    if (P.a_slots < 4) return 0;
    const size_t smem = (size_t)P.a_slots * P.plane_bytes + (size_t)P.b_slots * P.b_slot_bytes + 1024 + 512;
    AMB_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_halo_kernel<<<base_jobs * ksplit, 256, smem, (cudaStream_t)a->stream>>>(P);
    AMB_LAUNCH_CHECK();
    g_last_conv_kernel = "wgrad_halo_kernel";
    return 1;
}

}  // namespace amb
