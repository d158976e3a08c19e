// tcgen05 weight gradient of the two heaviest dense 3×3×3 stride-1 layers (LightDecoder block 3: 64→64 and 64→32 at 128³),
// round 2: the three dz taps are stacked along N.
//
//     dW[(dz,dy,dx)][co][ci] = Σ_o dY[o, co] · X[o + (dz,dy,dx), ci]
//
// conv_wgrad_halo.cu issues M128 × N64 × K16 MMAs (two taps' dY channels along M, the 64 X channels along N).  The probe
// tests/probes/mma_probe.cu shows why that tops out at 45 % of the tensor pipe: an M128 K16 SS-mode MMA costs
// ≥ 64 + N/8 cycles (the A operand is fetched from shared memory at 64 B/clk), i.e. 72 cycles at N = 64 for 32 cycles of
// math; only N ≥ 192 runs at the math rate.  Here one operand supplies halo PLANES (all 9 in-plane shifts of a z-plane are
// read in place through shifted descriptors, two of them stacked along M) and the other supplies plain 8×8 tiles of THREE
// consecutive z-planes, which sit in consecutive ring slots and form one N-stacked operand (N = 3·C_B):
//
//   role 1 (Cx = Cy = 64): A = dY plane z shifted by −(dy,dx)   B = X tiles z−1, z, z+1       D[(tap, co), (dz, ci)]   N = 192
//   role 2 (Cx = 64, Cy = 32): A = X plane z shifted by +(dy,dx)   B = dY tiles z−1, z, z+1   D[(tap, ci), (dz, co)]   N = 96
//
// TMEM holds 512 columns.  Role 1 needs 4.5 units of 192 columns, split over TWO CTA kinds of equal work so that both walk
// the same voxels in lockstep and share them through L2: {pair, pair, pair₃·(dz −1,0)} and {pair, single, pair₃·(dz +1)}
// = 192 + 192 + 128 and 192 + 192 + 64 columns, 271 and 264 tensor cycles per K16.  Role 2 fits one CTA (5 × 96 columns).
// A CTA walks z along a column of 8×8 tiles: per step ONE new plane and ONE new tile (the tile ring has two mirror slots so
// that every (z−1, z, z+1) triple is contiguous).  Split-K over CTAs, fp32 red.global into dW.
//
//   warp 0  plane producer     warp 1  MMA issuer     warp 2  TMEM allocator     warp 3  tile producer     warps 4-7 epilogue
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace amb {

using namespace ptx;

int encode_view_map(CUtensorMap* m, const void* tensor_base, const View& v, int C, int kc, const int box[4]);

#define NS_MAX_UNITS 8
#define NS_A_SLOTS 4
#define NS_B_RING 6                  // + 2 mirror slots

struct NsUnit {
    int32_t off;                     // byte offset of atom 0's window inside a plane slot
    int32_t lbo;                     // bytes between the two stacked windows
    int16_t blk0, nblk;              // N blocks (ring slots) blk0 .. blk0+nblk-1 of the (z−1, z, z+1) triple
    int16_t d_col, pad;
    int16_t w[2][3];                 // weight slab per (atom, block), −1 = discarded
};

struct WgradNsParams {
    CUtensorMap a_map;               // plane operand: box (CA channels, PW x, 10 y, 1, 1)
    CUtensorMap b_map;               // tile operand:  box (CB channels, 8 x, 8 y, 1, 1)
    NsUnit units[NS_MAX_UNITS];
    int16_t batch_begin[2], batch_count[2];
    int n_batches, role;             // role 1: A = dY (rows co, cols ci)   role 2: A = X (rows ci, cols co)
    int CA, CB, Cx, Cy;
    uint32_t a_slot_bytes, a_tx, b_slot_bytes;
    uint32_t a_layout, b_layout, a_sbo, b_sbo, a_kstep, b_kstep;
    uint32_t idesc[3];
    int oD, Ty, Tx;
    uint32_t n_steps;
    int ksplit;
    float* dw;
};

__global__ void __launch_bounds__(256, 1) wgrad_ns_kernel(const __grid_constant__ WgradNsParams P) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + NS_A_SLOTS * P.a_slot_bytes;
    uint8_t* ctrl = b_ring + (NS_B_RING + 2) * P.b_slot_bytes;
    uint64_t* a_full = (uint64_t*)ctrl;            // [4]
    uint64_t* a_empty = a_full + NS_A_SLOTS;       // [4]
    uint64_t* b_full = a_empty + NS_A_SLOTS;       // [6]
    uint64_t* b_empty = b_full + NS_B_RING;        // [6]
    uint64_t* acc_full = b_empty + NS_B_RING;
    uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS_A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NS_B_RING; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&P.a_map); prefetch_tmap(&P.b_map); }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // job: blockIdx.x = ks · n_batches + batch — the CTA kinds of one K split are neighbours and walk the same voxels
    const int batch = (int)(blockIdx.x % (uint32_t)P.n_batches);
    const uint32_t ks = blockIdx.x / (uint32_t)P.n_batches;
    const int unit_begin = P.batch_begin[batch], unit_count = P.batch_count[batch];
    const uint32_t s_begin = (uint32_t)((unsigned long long)P.n_steps * ks / (uint32_t)P.ksplit);
    const uint32_t s_end = (uint32_t)((unsigned long long)P.n_steps * (ks + 1) / (uint32_t)P.ksplit);
    const uint32_t oD = (uint32_t)P.oD;

    // a segment = the steps [z, zend) this CTA takes inside one column (n, y0, x0)
    auto column = [&](uint32_t col, int& n0, int& y0, int& x0) {
        x0 = (int)(col % (uint32_t)P.Tx) * 8; col /= (uint32_t)P.Tx;
        y0 = (int)(col % (uint32_t)P.Ty) * 8;
        n0 = (int)(col / (uint32_t)P.Ty);
    };

    if (warp == 0) {
        // planes z .. zend−1 (one per step); rows / columns outside the tensor are zero-filled by TMA = the zero padding
        uint32_t slot = 0, phase = 0, s = s_begin;
        while (s < s_end) {
            const uint32_t col = s / oD;
            const int z = (int)(s - col * oD);
            const uint32_t rem = s_end - s;
            const int zend = (oD - (uint32_t)z) < rem ? (int)oD : z + (int)rem;
            int n0, y0, x0;
            column(col, n0, y0, x0);
            for (int zp = z; zp < zend; ++zp) {
                mbar_wait(&a_empty[slot], phase ^ 1u, 51);
                if (elect_one()) {
                    mbar_expect_tx(&a_full[slot], P.a_tx);
                    tma_load_5d(a_ring + slot * P.a_slot_bytes, &P.a_map, &a_full[slot], 0, x0 - 1, y0 - 1, zp, n0);
                }
                __syncwarp();
                if (++slot == NS_A_SLOTS) { slot = 0; phase ^= 1u; }
            }
            s += (uint32_t)(zend - z);
        }
    } else if (warp == 3) {
        // tiles z−1 .. zend (len + 2 per segment); tile q lives in slot q % R and, for q % R < 2, also in mirror slot R + q % R
        uint32_t q = 0, s = s_begin;
        while (s < s_end) {
            const uint32_t col = s / oD;
            const int z = (int)(s - col * oD);
            const uint32_t rem = s_end - s;
            const int zend = (oD - (uint32_t)z) < rem ? (int)oD : z + (int)rem;
            int n0, y0, x0;
            column(col, n0, y0, x0);
            for (int zp = z - 1; zp <= zend; ++zp, ++q) {
                const uint32_t slot = q % NS_B_RING, phase = (q / NS_B_RING) & 1u;
                mbar_wait(&b_empty[slot], phase ^ 1u, 52);
                if (elect_one()) {
                    const bool mirror = slot < 2;
                    mbar_expect_tx(&b_full[slot], mirror ? 2u * P.b_slot_bytes : P.b_slot_bytes);
                    tma_load_5d(b_ring + slot * P.b_slot_bytes, &P.b_map, &b_full[slot], 0, x0, y0, zp, n0);
                    if (mirror)
                        tma_load_5d(b_ring + (NS_B_RING + slot) * P.b_slot_bytes, &P.b_map, &b_full[slot], 0, x0, y0, zp, n0);
                }
                __syncwarp();
            }
            s += (uint32_t)(zend - z);
        }
    } else if (warp == 1) {
        uint32_t u_alo[NS_MAX_UNITS], u_boff[NS_MAX_UNITS], u_id[NS_MAX_UNITS], u_d[NS_MAX_UNITS];
#pragma unroll
        for (int j = 0; j < NS_MAX_UNITS; ++j) {
            const NsUnit& U = P.units[unit_begin + (j < unit_count ? j : 0)];
            u_alo[j] = ((((uint32_t)U.lbo >> 4) & 0x3FFFu) << 16) | ((uint32_t)U.off >> 4);
            u_boff[j] = ((uint32_t)U.blk0 * P.b_slot_bytes) >> 4;
            u_id[j] = P.idesc[U.nblk - 1];
            u_d[j] = tmem_base + (uint32_t)U.d_col;
        }
        const uint64_t a_hi = (uint64_t)(uint32_t)(umma_desc(0, 0, P.a_sbo, P.a_layout) >> 32) << 32;
        const uint64_t b_hi = umma_desc(0, P.b_slot_bytes, P.b_sbo, P.b_layout);          // LBO = next N block = next ring slot
        const uint32_t a_k16 = P.a_kstep >> 4, b_k16 = P.b_kstep >> 4;
        const uint32_t a_ring_u32 = smem_u32(a_ring), b_ring_u32 = smem_u32(b_ring);
        uint32_t as = 0, aph = 0;                 // plane ring position
        uint32_t q = 0;                           // tile counter: the step's triple is tiles q, q+1, q+2
        bool accum = false;
        uint32_t s = s_begin;
        while (s < s_end) {
            const uint32_t col = s / oD;
            const uint32_t z = s - col * oD;
            const uint32_t rem = s_end - s;
            const uint32_t len = (oD - z) < rem ? (oD - z) : rem;
            for (int i = 0; i < 2; ++i) {         // the first two tiles of the segment
                const uint32_t t = q + (uint32_t)i;
                mbar_wait(&b_full[t % NS_B_RING], (t / NS_B_RING) & 1u, 53);
            }
            for (uint32_t j = 0; j < len; ++j, ++q) {
                const uint32_t t2 = q + 2u;
                mbar_wait(&b_full[t2 % NS_B_RING], (t2 / NS_B_RING) & 1u, 54);
                mbar_wait(&a_full[as], aph, 55);
                tc_fence_after();
                const uint32_t bslot = q % NS_B_RING;
                if (elect_one()) {
                    const uint32_t a0 = ((a_ring_u32 + as * P.a_slot_bytes) & 0x3FFFFu) >> 4;
                    const uint32_t b0 = ((b_ring_u32 + bslot * P.b_slot_bytes) & 0x3FFFFu) >> 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int u = 0; u < NS_MAX_UNITS; ++u) {
                            if (u < unit_count) {
                                const uint64_t ad = a_hi | (uint64_t)(u_alo[u] + a0 + a_k16 * (uint32_t)k);
                                const uint64_t bd = b_hi | (uint64_t)(b0 + u_boff[u] + b_k16 * (uint32_t)k);
                                mma_bf16(u_d[u], ad, bd, u_id[u], accum || (k != 0));
                            }
                        }
                    }
                    mma_commit(&a_empty[as]);
                    mma_commit(&b_empty[bslot]);          // tile q was the z−1 tile of this step: its last use
                }
                __syncwarp();
                accum = true;
                if (++as == NS_A_SLOTS) { as = 0; aph ^= 1u; }
            }
            if (elect_one()) {                            // the two trailing tiles of the segment
                mma_commit(&b_empty[q % NS_B_RING]);
                mma_commit(&b_empty[(q + 1u) % NS_B_RING]);
            }
            __syncwarp();
            q += 2u;
            s += len;
        }
        if (elect_one()) mma_commit(acc_full);
        __syncwarp();
    } else if (warp >= 4) {
        const int q4 = warp - 4;
        const int m = q4 * 32 + lane;
        mbar_wait(acc_full, 0, 56);
        tc_fence_after();
        if (s_end > s_begin) {
            const int atom = m / P.CA, r = m % P.CA;          // CA = 64: two atoms of 64 rows
            for (int u = 0; u < unit_count; ++u) {
                const NsUnit& U = P.units[unit_begin + u];
                for (int blk = 0; blk < U.nblk; ++blk) {
                    const int w = U.w[atom][blk];
                    const uint32_t t_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(U.d_col + blk * P.CB);
                    for (int col = 0; col < P.CB; col += 16) {
                        uint32_t rr[16];
                        tmem_ld_x16(t_addr + col, rr);
                        tmem_ld_wait();
                        if (w < 0) continue;
                        if (P.role == 1) {                    // rows co, columns ci: dW[w][co][ci .. ci+15]
                            float* dst = P.dw + ((long)w * P.Cy + r) * P.Cx + col;
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                                             "f"(__uint_as_float(rr[j])), "f"(__uint_as_float(rr[j + 1])),
                                             "f"(__uint_as_float(rr[j + 2])), "f"(__uint_as_float(rr[j + 3]))
                                             : "memory");
                        } else {                              // rows ci, columns co: dW[w][co][ci] — a warp covers 32 ci
                            float* dst = P.dw + ((long)w * P.Cy + col) * P.Cx + r;
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (long)j * P.Cx), "f"(__uint_as_float(rr[j]))
                                             : "memory");
                        }
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
int igemm_wgrad_ns(const Plan& p, const amb_wgrad_args* a) {
    if (getenv("AMB_DISABLE_WNS")) return 0;
    if (a->active_list != nullptr || a->op != AMB_OP_CONV) return 0;
    if (p.n_out_views != 1 || p.n_in_views != 1 || p.n_taps != 27) return 0;
    for (int i = 0; i < 27; ++i) {
        const Tap& T = p.taps[i];
        if (T.dz < -1 || T.dz > 1 || T.dy < -1 || T.dy > 1 || T.dx < -1 || T.dx > 1) return 0;
    }
    int role = 0;
    if (p.Cx == 64 && p.Cy == 64) role = 1;
    else if (p.Cx == 64 && p.Cy == 32) role = 2;
    if (role == 0) return 0;
    if (p.oH % 8 != 0 || p.oW % 8 != 0 || p.oH < 16 || p.oW < 16 || p.oD < 4) return 0;
    const View& vi = p.in_views[0];
    if (vi.D != p.oD || vi.H != p.oH || vi.W != p.oW) return 0;   // stride 1 only

    static WgradNsParams P;
    memset(&P, 0, sizeof(P));
    P.role = role;
    P.Cx = p.Cx; P.Cy = p.Cy;
    P.CA = 64;                                    // plane operand: dY (role 1) or X (role 2), 64 channels = one SW128 atom
    P.CB = role == 1 ? p.Cx : p.Cy;               // tile operand: X (64) or dY (32)
    const char* pwenv = getenv("AMB_WNS_PW");
    const int PW = (pwenv && atoi(pwenv) == 16) ? 16 : 10;
    const uint32_t a_row = 128u, b_row = (uint32_t)P.CB * 2u;
    P.a_tx = 10u * (uint32_t)PW * a_row;
    P.a_slot_bytes = (P.a_tx + 1023u) & ~1023u;
    P.b_slot_bytes = 64u * b_row;                 // 8 KB (64 ch) / 4 KB (32 ch)
    P.a_layout = 2u;                              // SW128
    P.b_layout = P.CB == 64 ? 2u : 4u;            // SW128 / SW64
    P.a_sbo = (uint32_t)PW * a_row;               // next 8-voxel group = next y row of the plane
    P.a_kstep = 2u * P.a_sbo;                     // K16 = two y rows
    P.b_sbo = 8u * b_row;
    P.b_kstep = 16u * b_row;
    for (int nb = 1; nb <= 3; ++nb) P.idesc[nb - 1] = umma_idesc_bf16(128, nb * P.CB, 1, 1);
    P.oD = p.oD; P.Ty = p.oH / 8; P.Tx = p.oW / 8;
    P.dw = a->dw;

    // in-plane taps sorted by window offset; sign = −1: the plane operand is dY (window shifted against the tap), +1: X
    const int sign = role == 1 ? -1 : 1;
    int t_dy[9], t_dx[9], n9 = 0;
    for (int sy = -1; sy <= 1; ++sy)
        for (int sx = -1; sx <= 1; ++sx) { t_dy[n9] = sign * sy; t_dx[n9] = sign * sx; ++n9; }   // window (sy,sx) ↔ tap (dy,dx)
    auto win_off = [&](int i) { return (int)((((sign * t_dy[i]) + 1) * PW + ((sign * t_dx[i]) + 1)) * (int)a_row); };
    auto slab = [&](int i, int dz) {
        for (int t = 0; t < 27; ++t)
            if (p.taps[t].dz == dz && p.taps[t].dy == t_dy[i] && p.taps[t].dx == t_dx[i]) return (int)p.taps[t].w;
        return -1;
    };
    // N block j of the (z−1, z, z+1) tile triple ↔ dz = j − 1 (role 1: X tile at z + dz) or 1 − j (role 2: dY tile at z − dz)
    auto dz_of_block = [&](int j) { return role == 1 ? j - 1 : 1 - j; };
    int nu = 0;
    auto add_unit = [&](int i0, int i1, int blk0, int nblk, int d_col) {
        NsUnit& U = P.units[nu++];
        U.off = win_off(i0);
        U.lbo = i1 >= 0 ? win_off(i1) - win_off(i0) : (int)a_row;
        U.blk0 = (int16_t)blk0; U.nblk = (int16_t)nblk; U.d_col = (int16_t)d_col;
        for (int at = 0; at < 2; ++at)
            for (int b = 0; b < 3; ++b) {
                const int i = at == 0 ? i0 : i1;
                U.w[at][b] = (int16_t)((i >= 0 && b < nblk) ? slab(i, dz_of_block(blk0 + b)) : -1);
            }
    };
    const int W3 = 3 * P.CB;
    if (role == 1) {
        P.n_batches = 2;
        P.batch_begin[0] = 0;
        add_unit(0, 1, 0, 3, 0); add_unit(2, 3, 0, 3, W3); add_unit(4, 5, 0, 2, 2 * W3);
        P.batch_count[0] = 3;
        P.batch_begin[1] = 3;
        add_unit(6, 7, 0, 3, 0); add_unit(8, -1, 0, 3, W3); add_unit(4, 5, 2, 1, 2 * W3);
        P.batch_count[1] = 3;
    } else {
        P.n_batches = 1;
        P.batch_begin[0] = 0;
        for (int i = 0; i < 4; ++i) add_unit(2 * i, 2 * i + 1, 0, 3, i * W3);
        add_unit(8, -1, 0, 3, 4 * W3);
        P.batch_count[0] = 5;
    }
    for (int u = 0; u < nu; ++u)
        for (int at = 0; at < 2; ++at)
            for (int b = 0; b < P.units[u].nblk; ++b)
                if (P.units[u].w[at][b] < 0 && !(at == 1 && u == (role == 1 ? 4 : 4))) { set_error("wgrad ns: tap lookup failed"); return -1; }

    const int abox[4] = {1, 1, 10, PW}, bbox[4] = {1, 1, 8, 8};
    const void* a_t = role == 1 ? a->dy : a->x;
    const void* b_t = role == 1 ? a->x : a->dy;
    const View& av = role == 1 ? p.out_views[0] : p.in_views[0];
    const View& bv = role == 1 ? p.in_views[0] : p.out_views[0];
    if (int e = encode_view_map(&P.a_map, a_t, av, P.CA, P.CA, abox)) return e;
    if (int e = encode_view_map(&P.b_map, b_t, bv, P.CB, P.CB, bbox)) return e;

    const long steps = (long)p.oN * P.Ty * P.Tx * p.oD;
    if (steps >= (1L << 31)) return 0;
    P.n_steps = (uint32_t)steps;
    int ksplit = num_sms() / P.n_batches;
    if ((long)ksplit > steps) ksplit = (int)steps;
    if (ksplit < 1) ksplit = 1;
    P.ksplit = ksplit;
    const size_t smem = (size_t)NS_A_SLOTS * P.a_slot_bytes + (size_t)(NS_B_RING + 2) * P.b_slot_bytes + 1024 + 512;
    if (smem > 227 * 1024) return 0;
    AMB_CUDA(cudaFuncSetAttribute(wgrad_ns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_ns_kernel<<<P.n_batches * ksplit, 256, smem, (cudaStream_t)a->stream>>>(P);
    AMB_LAUNCH_CHECK();
    g_last_conv_kernel = "wgrad_ns_kernel";
    return 1;
}

}  // namespace amb
