// tcgen05 implicit GEMM for 3×3×3 stride-1 convolutions with narrow outputs (one N tile, Cout ≤ 64), round 2:
// halo planes as in conv_igemm3.cu, but the three dz taps that read the SAME plane window are issued as ONE MMA.
//
// A CTA owns T z-adjacent tiles of 16 y × 8 x voxels; tile t keeps its accumulator in TMEM columns [t·NT, (t+1)·NT).
// Input plane p (z0−1+p) at in-plane tap (dy,dx) contributes to tile t = p − dz with weight W(dz,dy,dx), dz = 0..2.
// The tiles p−2, p−1, p are ADJACENT accumulators, so with the weight slab laid out as [W(dz=2) | W(dz=1) | W(dz=0)]
// (N-major blocks of NT rows) one tcgen05.mma of N = 3·NT does the work of three N = NT MMAs and reads the A window
// from shared memory once instead of three times (edge planes use the matching 1- or 2-block slice of the slab).
//   per (tap, k16): planes 0..T+1 → N = NT·(1,2,3,…,3,2,1): T+2 MMAs instead of 3T, A traffic (T+2)/3T of before.
//   M128×N64 SS-mode MMAs are shared-memory-operand bound (6 KB per 32 tensor cycles against 128 B/clk = 67 %);
//   at N = 192 it is 10 KB per 96 cycles = 81 % of the pipe, so the tensor core, not the operand fetch, is the limit.
// The very first k-step of a unit initialises each tile with its own N = NT MMA (a stacked MMA cannot mix
// "overwrite" and "accumulate" columns); everything after accumulates.
//
// Planes are TMA boxes [18 y][PW x][32 ch] (PW = 10: 11.25 KB in a 12 KB slot; 64-byte swizzle, rows 64 B): the A
// descriptor walks x inside an 8-row group and steps PW·64 B (SBO) per y — the swizzle phase follows absolute
// shared-memory address bits for both TMA and UMMA, so neither the window start nor SBO need 512 B alignment.
//
//   warp 0  plane producer (2 sets × (T+2) slots)   warp 1  MMA issuer            warp 2  TMEM allocator
//   warp 3  weight-slab producer (ring)             warps 4-7 epilogue (+bias, mask, Σ/Σ², bf16 stores)
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace amb {

using namespace ptx;

#define V4_SLOT 12288u
#define V4_B_SLOTS_MAX 8

struct Igemm4Params {
    CUtensorMap a_map;                   // dims (C, W, H, D, N), box (32, PW, 18, 1, 1), SWIZZLE_64B
    CUtensorMap w_map;                   // dims (Cx, Cy, 27), box (32, NT, 1), SWIZZLE_64B
    bf16* y;
    long sN, sD, sH, sW;
    const float* bias;
    const uint8_t* active;
    const int* list;                     // active-patch work-list (patch edge >= 16 output voxels) or nullptr = dense
    const int* count;
    double* stats;
    const float* ep_scale;               // optional fused epilogue y = act(acc·scale + bias) (inference-mode BN folded in)
    int ep_act;
    int oN, oD, oH, oW, Cy;
    int lgPv, fd, fh, fw;
    int Ty, Tx, Tzg, NT, kchunks, b_slots, order;
    uint32_t plane_tx, b_bytes, blk_bytes, tmem_cols, sbo_a;
    uint32_t idesc[3];                   // N = NT, 2·NT, 3·NT
    uint16_t row_off[9];                 // (dy·PW + dx)·64: window start inside a plane
    int16_t tap_w[3][9];                 // [plane offset dz][in-plane tap] → weight slab index
    int dbg_skip;                        // diagnostic (AMB_V4_SKIP, with AMB_V4_DBG): 1 = no weight-slab loads, 2 = no plane loads (wrong results)
    long long* dbg;                      // AMB_V4_DBG=1: per-CTA cycle counts of the MMA issuer {total, tempty, b_full, a_full, issue, units}
};

struct Unit4 {
    int n, y0, x0, z0;
};

template <int T>
__device__ __forceinline__ void v4_decode(const Igemm4Params& P, uint32_t u, Unit4& c) {
    if (P.list) {
        // visible patches only (T = 4): a patch of edge Pv holds (Pv/16) x (Pv/8) x (Pv/4) units of 16 x 8 x 4 voxels
        const uint32_t ly = (uint32_t)P.lgPv - 4u, lx = (uint32_t)P.lgPv - 3u, lz = (uint32_t)P.lgPv - 2u;
        const uint32_t ix = u & ((1u << lx) - 1u); u >>= lx;
        const uint32_t iy = u & ((1u << ly) - 1u); u >>= ly;
        const uint32_t iz = u & ((1u << lz) - 1u); u >>= lz;
        const uint32_t pid = (uint32_t)P.list[u];
        const uint32_t L = (uint32_t)(P.fd * P.fh * P.fw), hw = (uint32_t)(P.fh * P.fw);
        const uint32_t n = pid / L, l = pid - n * L;
        const uint32_t pz = l / hw, r2 = l - pz * hw;
        const uint32_t py = r2 / (uint32_t)P.fw, px = r2 - py * (uint32_t)P.fw;
        c.n = (int)n;
        c.z0 = (int)((pz << P.lgPv) + iz * 4u);
        c.y0 = (int)((py << P.lgPv) + iy * 16u);
        c.x0 = (int)((px << P.lgPv) + ix * 8u);
        return;
    }
    c.z0 = (int)(u % (uint32_t)P.Tzg) * T; u /= (uint32_t)P.Tzg;
    c.x0 = (int)(u % (uint32_t)P.Tx) * 8; u /= (uint32_t)P.Tx;
    c.y0 = (int)(u % (uint32_t)P.Ty) * 16;
    c.n = (int)(u / (uint32_t)P.Ty);
}

__device__ __forceinline__ float v4_column_sums(float* v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            float send = up ? v[i] : v[i + o];
            float keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

template <int T>
__global__ void __launch_bounds__(256, 1) igemm4_kernel(const __grid_constant__ Igemm4Params P) {
    constexpr int NP = T + 2;                      // planes per (unit, channel chunk)
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;                                        // [2 sets][NP] slots of V4_SLOT bytes
    uint8_t* b_ring = smem + 2u * NP * V4_SLOT;
    uint8_t* ctrl = b_ring + (size_t)P.b_slots * P.b_bytes;
    uint64_t* a_full = (uint64_t*)ctrl;            // [2 * NP] <= 16
    uint64_t* a_empty = a_full + 16;               // [2]
    uint64_t* b_full = a_empty + 2;                // [8]
    uint64_t* b_empty = b_full + V4_B_SLOTS_MAX;   // [8]
    uint64_t* tfull = b_empty + V4_B_SLOTS_MAX;    // [2]
    uint64_t* tempty = tfull + 2;                  // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    float* s_stats = (float*)(ctrl + 512);
    float* s_bias = s_stats + (P.stats ? 2 * P.Cy : 0);   // [Cy] (zeros without a bias): the epilogue reads it as float4
    float* s_scale = s_bias + P.Cy;                       // [Cy] (ones without ep_scale)
    const bool ep = P.ep_scale != nullptr || P.ep_act != 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * NP; ++s) mbar_init(&a_full[s], 1);
        for (int s = 0; s < 2; ++s) mbar_init(&a_empty[s], 1);
        for (int s = 0; s < V4_B_SLOTS_MAX; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&P.a_map); prefetch_tmap(&P.w_map); }
    if (P.stats) for (int i = threadIdx.x; i < 2 * P.Cy; i += blockDim.x) s_stats[i] = 0.f;
    for (int i = threadIdx.x; i < P.Cy; i += blockDim.x) { s_bias[i] = P.bias ? P.bias[i] : 0.f; s_scale[i] = P.ep_scale ? P.ep_scale[i] : 1.f; }
    if (warp == 2) tmem_alloc(tmem_slot, P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t nunits = P.list ? ((uint32_t)(*P.count) << (3 * P.lgPv - 9))
                                   : (uint32_t)(P.oN * P.Ty * P.Tx * P.Tzg);
    const uint32_t kchunks = (uint32_t)P.kchunks, NT = (uint32_t)P.NT, b_bytes = P.b_bytes;
    const uint32_t a_ring_u32 = smem_u32(a_ring), b_ring_u32 = smem_u32(b_ring);
    const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty);
    const uint32_t b_full0 = smem_u32(b_full), b_empty0 = smem_u32(b_empty);
    const uint32_t B_SLOTS = (uint32_t)P.b_slots;

    if (warp == 0) {
        // =============================== plane producer ===============================
        uint32_t cc = 0;                           // chunk counter: set = cc & 1, use count of that set = cc >> 1
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            Unit4 c;
            v4_decode<T>(P, u, c);
            for (uint32_t kc = 0; kc < kchunks; ++kc, ++cc) {
                const uint32_t set = cc & 1u, ph = (cc >> 1) & 1u;
                mbar_wait_u32(a_empty0 + set * 8u, ph ^ 1u, 41);
                if (elect_one()) {
                    // ONE barrier per plane set: the issuer polls once per chunk instead of once per plane
                    const uint32_t bar = a_full0 + set * 8u;
                    mbar_expect_tx_u32(bar, (P.dbg_skip & 2) ? 0u : (uint32_t)NP * P.plane_tx);
#pragma unroll
                    for (int pl = 0; pl < NP; ++pl)
                        if (!(P.dbg_skip & 2))
                        tma_load_5d_u32(a_ring_u32 + (set * NP + (uint32_t)pl) * V4_SLOT, &P.a_map, bar, (int)(kc * 32),
                                        c.x0 - 1, c.y0 - 1, c.z0 - 1 + pl, c.n);
                }
                __syncwarp();
            }
        }
    } else if (warp == 3) {
        // =============================== weight-slab producer ===============================
        // one slab per (channel chunk, in-plane tap): the three dz blocks in DEscending dz order = ascending tile order
        uint32_t slot = 0, phase = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            for (uint32_t kc = 0; kc < kchunks; ++kc) {
                for (int t9 = 0; t9 < 9; ++t9) {
                    mbar_wait_u32(b_empty0 + slot * 8u, phase ^ 1u, 42);
                    if (elect_one()) {
                        const uint32_t bar = b_full0 + slot * 8u;
                        mbar_expect_tx_u32(bar, (P.dbg_skip & 1) ? 0u : 3u * P.blk_bytes);
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            if (!(P.dbg_skip & 1))
                            tma_load_3d_u32(b_ring_u32 + slot * b_bytes + (uint32_t)j * P.blk_bytes, &P.w_map, bar,
                                            (int)(kc * 32), 0, P.tap_w[2 - j][t9]);
                    }
                    __syncwarp();
                    if (++slot == B_SLOTS) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        // K-major, 64-byte swizzle: layout 4.  A: 8-row groups (x) sbo_a bytes apart (one y step of the plane).  B: 512 B.
        const uint64_t a_hi = (uint64_t)(uint32_t)(umma_desc(0, 16, P.sbo_a, 4) >> 32) << 32;
        const uint64_t b_hi = (uint64_t)(uint32_t)(umma_desc(0, 16, 512, 4) >> 32) << 32;
        const uint32_t lo_const = (uint32_t)(umma_desc(0, 16, 0, 4) & 0xFFFFFFFFu);
        const uint32_t id1 = P.idesc[0], id2 = P.idesc[1], id3 = P.idesc[2];
        const uint32_t blk16 = P.blk_bytes >> 4;
        uint32_t cc = 0, b_slot = 0, b_phase = 0, iter = 0;
        const bool dbg = P.dbg != nullptr;
        long long t_start = 0, w_te = 0, w_b = 0, w_a = 0, w_i = 0, tq = 0;
        if (dbg) t_start = clock64();
        // the tensor pipe queues only a few MMAs behind the issuing thread, so every cycle this thread spends polling a barrier
        // between two taps is a cycle the pipe idles: the NEXT tap's weight slab (and at the end of a chunk the next plane set)
        // is polled non-blockingly before the current tap's MMAs are issued; the blocking wait remains as the fallback
        uint32_t b_ready = 0, a_ready = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x, ++iter) {
            const uint32_t acc = iter & 1u;
            if (dbg) tq = clock64();
            mbar_wait_u32(smem_u32(&tempty[acc]), ((iter >> 1) & 1u) ^ 1u, 43);
            if (dbg) w_te += clock64() - tq;
            tc_fence_after();
            const uint32_t d_base = tmem_base + acc * T * NT;
            uint32_t d_tile[T];
#pragma unroll
            for (int t = 0; t < T; ++t) d_tile[t] = d_base + (uint32_t)t * NT;
            for (uint32_t kc = 0; kc < kchunks; ++kc, ++cc) {
                const uint32_t set = cc & 1u, aph = (cc >> 1) & 1u;
                const uint32_t a_set = a_ring_u32 + set * NP * V4_SLOT;
#pragma unroll
                for (int t9 = 0; t9 < 9; ++t9) {
                    if (dbg) tq = clock64();
                    if (!b_ready) mbar_wait_u32(b_full0 + b_slot * 8u, b_phase, 45);
                    if (dbg) { const long long t = clock64(); w_b += t - tq; tq = t; }
                    if (t9 == 0 && !a_ready) mbar_wait_u32(a_full0 + set * 8u, aph, 44);
                    if (dbg) { const long long t = clock64(); w_a += t - tq; tq = t; }
                    tc_fence_after();
                    {
                        const uint32_t nslot = b_slot + 1 == B_SLOTS ? 0u : b_slot + 1;
                        b_ready = mbar_test_wait_u32(b_full0 + nslot * 8u, b_slot + 1 == B_SLOTS ? b_phase ^ 1u : b_phase);
                        if (t9 == 8) a_ready = mbar_test_wait_u32(a_full0 + ((cc + 1) & 1u) * 8u, ((cc + 1) >> 1) & 1u);
                    }
                    // window start inside a plane: (dy·PW + dx)·64 B with PW = 10 — a compile-time constant once t9 is unrolled
                    const uint32_t a_lo = lo_const | (((a_set + (uint32_t)(((t9 / 3) * 10 + t9 % 3) * 64)) & 0x3FFFFu) >> 4);
                    const uint32_t b_lo = lo_const | (((b_ring_u32 + b_slot * b_bytes) & 0x3FFFFu) >> 4);
                    const bool first = (kc | (uint32_t)t9) == 0u;
                    if (elect_one()) {
                        // plane pl covers tiles lo..hi (ascending) with dz = pl-lo .. pl-hi → slab blocks 2-(pl-lo) ...
                        auto issue = [&](int pl, int k, int lo, int hi, bool accumulate) {
                            const int cnt = hi - lo + 1, j0 = 2 - (pl - lo);
                            const uint64_t adesc = a_hi | (uint64_t)(a_lo + (uint32_t)pl * (V4_SLOT >> 4) + (uint32_t)(2 * k));
                            const uint64_t bdesc = b_hi | (uint64_t)(b_lo + (uint32_t)j0 * blk16 + (uint32_t)(2 * k));
                            mma_bf16(d_base + (uint32_t)lo * NT, adesc, bdesc, cnt == 1 ? id1 : (cnt == 2 ? id2 : id3), accumulate);
                        };
                        if (first) {
                            // k = 0: every tile is initialised by its own dz = 0 MMA, the dz >= 1 parts follow stacked
#pragma unroll
                            for (int t = 0; t < T; ++t) issue(t, 0, t, t, false);
#pragma unroll
                            for (int pl = 1; pl < NP; ++pl) {
                                const int lo = pl - 2 < 0 ? 0 : pl - 2, hi = pl - 1 > T - 1 ? T - 1 : pl - 1;
                                issue(pl, 0, lo, hi, true);
                            }
#pragma unroll
                            for (int pl = 0; pl < NP; ++pl) {
                                const int lo = pl - 2 < 0 ? 0 : pl - 2, hi = pl > T - 1 ? T - 1 : pl;
                                issue(pl, 1, lo, hi, true);
                            }
                        } else if (P.order == 0) {
                            // running descriptors: A walks plane by plane (+ one slot), B only moves over the first planes
                            // (slab block 2, 1, 0, 0, ...), the accumulator addresses are four values per unit
                            uint64_t ad = a_hi | (uint64_t)a_lo;
                            uint64_t bd = b_hi | (uint64_t)(b_lo + 2u * blk16);
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
#pragma unroll
                                for (int pl = 0; pl < NP; ++pl) {
                                    const int lo = pl - 2 < 0 ? 0 : pl - 2, hi = pl > T - 1 ? T - 1 : pl;
                                    const int cnt = hi - lo + 1;
                                    mma_bf16(d_tile[lo], ad, bd, cnt == 1 ? id1 : (cnt == 2 ? id2 : id3), true);
                                    if (pl + 1 < NP) {
                                        desc_advance(ad, (int32_t)(V4_SLOT >> 4));
                                        if (pl < 2) desc_advance(bd, -(int32_t)blk16);
                                    } else if (k == 0) {
                                        desc_advance(ad, 2 - (int32_t)((NP - 1) * (V4_SLOT >> 4)));
                                        desc_advance(bd, 2 + (int32_t)((NP - 1 < 2 ? NP - 1 : 2) * blk16));
                                    }
                                }
                            }
                        } else {
#pragma unroll
                            for (int pl = 0; pl < NP; ++pl) {
                                const int lo = pl - 2 < 0 ? 0 : pl - 2, hi = pl > T - 1 ? T - 1 : pl;
#pragma unroll
                                for (int k = 0; k < 2; ++k) issue(pl, k, lo, hi, true);
                            }
                        }
                        mma_commit_u32(b_empty0 + b_slot * 8u);
                        if (t9 == 8) mma_commit_u32(a_empty0 + set * 8u);      // the whole plane set is free again
                    }
                    __syncwarp();
                    if (dbg) w_i += clock64() - tq;
                    if (++b_slot == B_SLOTS) { b_slot = 0; b_phase ^= 1u; }
                }
            }
            if (elect_one()) mma_commit_u32(smem_u32(&tfull[acc]));
            __syncwarp();
        }
        if (dbg && lane == 0) {
            long long* o = P.dbg + (size_t)blockIdx.x * 8;
            o[0] = clock64() - t_start; o[1] = w_te; o[2] = w_b; o[3] = w_a; o[4] = w_i; o[5] = iter;
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===============================
        const int q = warp - 4;
        const int row = q * 32 + lane;
        uint32_t iter = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x, ++iter) {
            Unit4 c;
            v4_decode<T>(P, u, c);
            const uint32_t acc = iter & 1u;
            const int x = c.x0 + (row & 7), y = c.y0 + (row >> 3);
            const bool valid_xy = y < P.oH && x < P.oW;
            mbar_wait(&tfull[acc], (iter >> 1) & 1u, 46);
            tc_fence_after();
            for (int t = 0; t < T; ++t) {
                const int z = c.z0 + t;
                if (z >= P.oD) break;
                bool on = valid_xy;
                if (valid_xy && P.active && P.lgPv >= 0)
                    on = P.active[((c.n * P.fd + (z >> P.lgPv)) * P.fh + (y >> P.lgPv)) * P.fw + (x >> P.lgPv)] != 0;
                bf16* yrow = P.y + (long)c.n * P.sN + (long)z * P.sD + (long)y * P.sH + (long)x * P.sW;
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * T + (uint32_t)t) * NT;
                for (int col = 0; col < P.NT; col += 32) {
                    uint32_t r[32];
                    const bool wide = (P.NT - col) >= 32;
                    if (wide) tmem_ld_x32(t_addr + col, r);
                    else tmem_ld_x16(t_addr + col, r);
                    tmem_ld_wait();
                    const int ncol = wide ? 32 : 16;
                    float v[32];
                    const float4* bq = reinterpret_cast<const float4*>(s_bias + col);
                    const float4* sq4 = reinterpret_cast<const float4*>(s_scale + col);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (j < ncol) { b4 = bq[j >> 2]; if (ep) s4 = sq4[j >> 2]; }
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            float f = 0.f;
                            if (j < ncol && on) f = ep ? ep_apply(__uint_as_float(r[j + jj]), ss[jj], bb[jj], P.ep_act) : __uint_as_float(r[j + jj]) + bb[jj];
                            v[j + jj] = f;
                        }
                    }
                    if (valid_xy) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (j < ncol) {
                                uint4 o;
                                o.x = pack2(v[j], v[j + 1]); o.y = pack2(v[j + 2], v[j + 3]);
                                o.z = pack2(v[j + 4], v[j + 5]); o.w = pack2(v[j + 6], v[j + 7]);
                                *reinterpret_cast<uint4*>(yrow + col + j) = o;
                            }
                        }
                    }
                    if (P.stats) {
                        float sq[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) { if (!valid_xy) v[j] = 0.f; sq[j] = v[j] * v[j]; }
                        float s1 = v4_column_sums(v);
                        float s2 = v4_column_sums(sq);
                        if (lane < ncol) {
                            atomicAdd(&s_stats[col + lane], s1);
                            atomicAdd(&s_stats[P.Cy + col + lane], s2);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (P.stats) {
        for (int i = threadIdx.x; i < 2 * P.Cy; i += blockDim.x) {
            float v = s_stats[i];
            if (v != 0.f) atomicAdd(&P.stats[i], (double)v);
        }
    }
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, P.tmem_cols);
    }
}

typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// returns 1 when handled, 0 when the shape is not for this kernel, <0 on error
int igemm4_conv(const Plan& p, const amb_conv_args* a) {
    if (env_int("AMB_DISABLE_V4", 0) == 1) return 0;
    if (!((a->op == AMB_OP_CONV || a->op == AMB_OP_CONV_DGRAD) && a->k == 3 && a->stride == 1)) return 0;
    if (p.Cx % 32 != 0 || p.Cy % 16 != 0 || p.n_taps != 27 || p.n_in_views != 1 || p.n_groups != 1) return 0;
    // one N tile of Cy columns: T·Cy·2 <= 512 TMEM columns → Cy <= 64: T = 4 (6), Cy = 128: T = 2 (N = 128 / 256 MMAs),
    // Cy = 256: T = 1 (no stacking left, but the halo planes still cut the operand traffic of the per-tap kernel 9x)
    const int wide_max = env_int("AMB_V4_WIDE", 256);
    if (p.Cy > 64 && !((p.Cy == 128 || p.Cy == 256) && p.Cy <= wide_max)) return 0;
    if (p.oH < 16 || p.oW < 8 || p.oD < 4) return 0;
    // active-patch work-list: usable when a patch holds whole 16 x 8 x 4 units; with a finer grid the per-tap kernel's list
    // (2 x 8 x 8 tiles) still pays, so leave those layers to it
    const bool use_list = a->active_list != nullptr && p.lgPv >= 4 && !getenv("AMB_V3_NO_LIST");
    if (a->active_list != nullptr && p.lgPv >= 3 && !use_list) return 0;
    const int NT = p.Cy;
    if (a->stats && p.Cy > 2048) return 0;
    // tiles per unit: 4 (two accumulator sets of 4·NT columns = all of TMEM at NT = 64); 6 at NT <= 32 on dense layers:
    // N = 96 MMAs are A-fetch bound (76 cycles for 48 of math), the fewer edge planes the better (789 → 897 TFLOP/s on 64→32@128³)
    int T = (NT <= 32 && !use_list && p.oD >= 12) ? 6 : 4;
    const int t_env = env_int("AMB_V4_T", 0);
    if (t_env == 4) T = 4;
    if (NT == 128) T = 2;
    if (NT == 256) T = 1;
    if (use_list && T != 4) return 0;

    static Igemm4Params P;
    memset(&P, 0, sizeof(P));
    P.y = (bf16*)a->y;
    const View& ov = p.out_views[0];
    P.sN = ov.sN; P.sD = ov.sD; P.sH = ov.sH; P.sW = ov.sW;
    P.bias = a->bias; P.active = a->active; P.stats = a->stats; P.ep_scale = a->ep_scale; P.ep_act = a->ep_act;
    P.list = use_list ? a->active_list : nullptr;
    P.count = use_list ? a->active_count : nullptr;
    P.oN = p.oN; P.oD = p.oD; P.oH = p.oH; P.oW = p.oW; P.Cy = p.Cy;
    P.lgPv = p.lgPv; P.fd = p.fd; P.fh = p.fh; P.fw = p.fw;
    P.Ty = ceil_div(p.oH, 16); P.Tx = ceil_div(p.oW, 8); P.Tzg = ceil_div(p.oD, T);
    P.NT = NT; P.kchunks = p.Cx / 32;
    P.order = env_int("AMB_V4_ORDER", 0);
    const int PW = 10;                                // 8 x + halo
    const uint32_t slot_need = (uint32_t)(18 * PW * 64);
    P.plane_tx = slot_need;
    P.sbo_a = (uint32_t)PW * 64u;
    P.blk_bytes = (uint32_t)NT * 64u;
    P.b_bytes = (3u * P.blk_bytes + 1023u) & ~1023u;
    P.tmem_cols = 32;
    while (P.tmem_cols < (uint32_t)(2 * T * NT)) P.tmem_cols <<= 1;
    for (int c = 1; c <= 3; ++c) P.idesc[c - 1] = c * NT <= 256 ? umma_idesc_bf16(128, c * NT, 0, 0) : 0u;   // (c <= T)
    // in-plane taps in raster order (dy, dx) = (t9 / 3 - 1, t9 % 3 - 1) — the kernel's unrolled tap loop has the window offsets
    // as immediates; the weight slab of every (dz, dy, dx) is looked up in the plan (forward and input-gradient plans differ)
    int8_t t_dy[9], t_dx[9];
    for (int t9 = 0; t9 < 9; ++t9) { t_dy[t9] = (int8_t)(t9 / 3 - 1); t_dx[t9] = (int8_t)(t9 % 3 - 1); }
    for (int t9 = 0; t9 < 9; ++t9) {
        P.row_off[t9] = (uint16_t)(((t_dy[t9] + 1) * PW + (t_dx[t9] + 1)) * 64);
        for (int dz = 0; dz < 3; ++dz) {
            int found = -1;
            for (int t = 0; t < 27; ++t)
                if (p.taps[t].dz == dz - 1 && p.taps[t].dy == t_dy[t9] && p.taps[t].dx == t_dx[t9]) found = p.taps[t].w;
            if (found < 0) return 0;
            P.tap_w[dz][t9] = (int16_t)found;
        }
    }

    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    AMB_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) == cudaSuccess &&
                  q == cudaDriverEntryPointSuccess, AMB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    EncodeTiledFn4 enc = (EncodeTiledFn4)fnp;
    {
        const View& iv = p.in_views[0];
        cuuint64_t dims[5] = {(cuuint64_t)p.Cx, (cuuint64_t)iv.W, (cuuint64_t)iv.H, (cuuint64_t)iv.D, (cuuint64_t)iv.N};
        cuuint64_t strides[4] = {(cuuint64_t)iv.sW * 2, (cuuint64_t)iv.sH * 2, (cuuint64_t)iv.sD * 2, (cuuint64_t)iv.sN * 2};
        cuuint32_t box[5] = {32, (cuuint32_t)PW, 18, 1, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&P.a_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)((const bf16*)a->x + iv.base), dims, strides,
                         box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AMB_CHECK(r == CUDA_SUCCESS, AMB_ERR_CUDA, "cuTensorMapEncodeTiled(halo plane, v4) failed: %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Cx, (cuuint64_t)p.Cy, 27};
        cuuint64_t strides[2] = {(cuuint64_t)p.Cx * 2, (cuuint64_t)p.Cy * p.Cx * 2};
        cuuint32_t box[3] = {32, (cuuint32_t)NT, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&P.w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)a->w, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AMB_CHECK(r == CUDA_SUCCESS, AMB_ERR_CUDA, "cuTensorMapEncodeTiled(weight slab, v4) failed: %d", (int)r);
    }
    // shared memory: two plane sets (the next chunk streams in under the current one) + as many weight slabs as fit
    const size_t fixed = 2u * (size_t)(T + 2) * V4_SLOT + 1024 + 512 +
                         (a->stats ? 2 * (size_t)p.Cy * sizeof(float) : 0) + 2 * (size_t)p.Cy * sizeof(float);
    int b_slots = env_int("AMB_V4_B_SLOTS", V4_B_SLOTS_MAX);
    if (b_slots > V4_B_SLOTS_MAX) b_slots = V4_B_SLOTS_MAX;
    while (b_slots > 2 && fixed + (size_t)b_slots * P.b_bytes > 227 * 1024) b_slots--;
    if (b_slots < 3 || fixed + (size_t)b_slots * P.b_bytes > 227 * 1024) return 0;
    P.b_slots = b_slots;
    const size_t smem = fixed + (size_t)b_slots * P.b_bytes;
    long units = use_list ? ((long)p.oN * p.fd * p.fh * p.fw << (3 * p.lgPv - 9)) : (long)p.oN * P.Ty * P.Tx * P.Tzg;
    int grid = (int)(units < (long)num_sms() ? units : (long)num_sms());
    static long long* dbg_buf = nullptr;
    const bool dbg = getenv("AMB_V4_DBG") != nullptr;                    // issuer cycle accounting (diagnostic; synchronises)
    if (dbg) {
        if (!dbg_buf) AMB_CUDA(cudaMalloc(&dbg_buf, 8 * sizeof(long long) * 1024));
        AMB_CUDA(cudaMemsetAsync(dbg_buf, 0, 8 * sizeof(long long) * 1024, (cudaStream_t)a->stream));
        P.dbg = dbg_buf;
        P.dbg_skip = env_int("AMB_V4_SKIP", 0);
    }
    if (T == 6) {
        AMB_CUDA(cudaFuncSetAttribute(igemm4_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        igemm4_kernel<6><<<grid, 256, smem, (cudaStream_t)a->stream>>>(P);
    } else if (T == 2) {
        AMB_CUDA(cudaFuncSetAttribute(igemm4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        igemm4_kernel<2><<<grid, 256, smem, (cudaStream_t)a->stream>>>(P);
    } else if (T == 1) {
        AMB_CUDA(cudaFuncSetAttribute(igemm4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        igemm4_kernel<1><<<grid, 256, smem, (cudaStream_t)a->stream>>>(P);
    } else {
        AMB_CUDA(cudaFuncSetAttribute(igemm4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        igemm4_kernel<4><<<grid, 256, smem, (cudaStream_t)a->stream>>>(P);
    }
    AMB_LAUNCH_CHECK();
    if (dbg) {
        static long long host[8 * 1024];
        AMB_CUDA(cudaStreamSynchronize((cudaStream_t)a->stream));
        AMB_CUDA(cudaMemcpy(host, dbg_buf, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost));
        double s[6] = {0, 0, 0, 0, 0, 0}, mx = 0;
        for (int b = 0; b < grid; ++b) { for (int j = 0; j < 6; ++j) s[j] += (double)host[b * 8 + j]; if (host[b * 8] > mx) mx = (double)host[b * 8]; }
        fprintf(stderr, "V4DBG T=%d NT=%d Cx=%d grid=%d units/CTA=%.1f | issuer cycles/CTA: total %.0f (max %.0f)  tempty %.1f%%  b_full %.1f%%  "
                        "a_full %.1f%%  issue %.1f%%  | per (chunk, tap): %.0f cycles\n", T, NT, p.Cx, grid, s[5] / grid, s[0] / grid, mx,
                100 * s[1] / s[0], 100 * s[2] / s[0], 100 * s[3] / s[0], 100 * s[4] / s[0], s[0] / (s[5] * P.kchunks * 9));
    }
    g_last_conv_kernel = "igemm4_kernel";
    return 1;
}

}  // namespace amb
