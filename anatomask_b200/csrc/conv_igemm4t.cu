// tcgen05 implicit GEMM for the ConvTranspose3d(k4, s2, p1) FORWARD with narrow outputs (Cout <= 64): the decoder's last
// up-sampling layer (P/decoder3D.py:17; STUNet-B dec.3.up_sample 64 -> 64, 64^3 -> 128^3, 275 GF per launch, run by the
// teacher and the student).  Same halo planes and N-stacking as conv_igemm4.cu, applied to the transposed geometry.
//
// Output voxel 2j + p (parity p per axis) is a 2x2x2 convolution of the input around j:  out index o = 2i - 1 + k, so
//   p = 0: (k = 1, i = j), (k = 3, i = j - 1)        p = 1: (k = 0, i = j + 1), (k = 2, i = j)
// A CTA unit = 2 z-adjacent input-resolution tiles of 16 y x 8 x voxels and ONE in-plane parity class (py, px); it keeps the
// four accumulators a = 2·t + pz (tile t, z parity pz) in adjacent TMEM columns [a·NT, (a+1)·NT).  Input plane pl (z0 - 1 + pl),
// read through the window of in-plane tap (dy, dx), contributes
//   pl = 0 → a0 (kz 3)    pl = 1 → a0, a1, a2 (kz 1, 2, 3)    pl = 2 → a1, a2, a3 (kz 0, 1, 2)    pl = 3 → a3 (kz 0)
// so with the weight slab laid out as [W(kz=0) | W(kz=1) | W(kz=2) | W(kz=3)] the two inner planes are ONE MMA of N = 3·NT each:
// 4 MMAs per (in-plane tap, k16) instead of 8, the A window fetched from shared memory once per plane instead of per tap.
// The per-tap kernel (conv_igemm.cu) runs this layer at 560-590 TFLOP/s — every N = 64 MMA pays the 64-cycle A fetch.
//
//   warp 0  plane producer (2 sets x 4 slots)    warp 1  MMA issuer             warp 2  TMEM allocator
//   warp 3  weight-slab producer (ring)           warps 4-7 epilogue (+bias, bf16 stores to the stride-2 output positions)
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace amb {

using namespace ptx;

#define V4T_SLOT 12288u
#define V4T_B_SLOTS_MAX 8
#define V4T_NP 4

struct Igemm4tParams {
    CUtensorMap a_map;                   // input: dims (C, W, H, D, N), box (32, PW, 18, 1, 1), SWIZZLE_64B
    CUtensorMap w_map;                   // weights: dims (Cx, Cy, 64), box (32, NT, 1), SWIZZLE_64B
    bf16* y;                             // output (N, 2D, 2H, 2W, Cy)
    long sN, sD, sH, sW;                 // its element strides
    const float* bias;
    int iN, iD, iH, iW, Cy;
    int Ty, Tx, Tzg, NT, kchunks, b_slots;
    uint32_t plane_tx, b_bytes, blk_bytes, tmem_cols, sbo_a;
    uint32_t idesc[3];                   // N = NT, 2·NT, 3·NT
    uint16_t row_off[4][4];              // [class q = 2·py + px][in-plane tap] → window start inside a plane
    int16_t slab[4][4][4];               // [class][in-plane tap][kz] → weight slab index (kz·4 + ky)·4 + kx
};

struct Unit4t {
    int n, y0, x0, z0;
};

__device__ __forceinline__ void v4t_decode(const Igemm4tParams& P, uint32_t u, Unit4t& c) {
    c.z0 = (int)(u % (uint32_t)P.Tzg) * 2; u /= (uint32_t)P.Tzg;
    c.x0 = (int)(u % (uint32_t)P.Tx) * 8; u /= (uint32_t)P.Tx;
    c.y0 = (int)(u % (uint32_t)P.Ty) * 16;
    c.n = (int)(u / (uint32_t)P.Ty);
}

__global__ void __launch_bounds__(256, 1) igemm4t_kernel(const __grid_constant__ Igemm4tParams P) {
    constexpr int NP = V4T_NP;
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;                                        // [2 sets][NP] slots of V4T_SLOT bytes
    uint8_t* b_ring = smem + 2u * NP * V4T_SLOT;
    uint8_t* ctrl = b_ring + (size_t)P.b_slots * P.b_bytes;
    uint64_t* a_full = (uint64_t*)ctrl;            // [2 * NP]
    uint64_t* a_empty = a_full + 16;               // [2]
    uint64_t* b_full = a_empty + 2;                // [8]
    uint64_t* b_empty = b_full + V4T_B_SLOTS_MAX;  // [8]
    uint64_t* tfull = b_empty + V4T_B_SLOTS_MAX;   // [2]
    uint64_t* tempty = tfull + 2;                  // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    float* s_bias = (float*)(ctrl + 512);          // [Cy] (zeros without a bias)

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * NP; ++s) mbar_init(&a_full[s], 1);
        for (int s = 0; s < 2; ++s) mbar_init(&a_empty[s], 1);
        for (int s = 0; s < V4T_B_SLOTS_MAX; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&P.a_map); prefetch_tmap(&P.w_map); }
    for (int i = threadIdx.x; i < P.Cy; i += blockDim.x) s_bias[i] = P.bias ? P.bias[i] : 0.f;
    if (warp == 2) tmem_alloc(tmem_slot, P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t nunits = (uint32_t)(P.iN * P.Ty * P.Tx * P.Tzg);
    const uint32_t kchunks = (uint32_t)P.kchunks, NT = (uint32_t)P.NT, b_bytes = P.b_bytes;
    const uint32_t a_ring_u32 = smem_u32(a_ring), b_ring_u32 = smem_u32(b_ring);
    const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty);
    const uint32_t b_full0 = smem_u32(b_full), b_empty0 = smem_u32(b_empty);
    const uint32_t B_SLOTS = (uint32_t)P.b_slots;

    if (warp == 0) {
        // =============================== plane producer ===============================
        // the four in-plane classes of a spatial unit run back to back on this CTA and re-load the same planes (L2 hits)
        uint32_t cc = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            Unit4t c;
            v4t_decode(P, u, c);
            for (int q = 0; q < 4; ++q) {
                for (uint32_t kc = 0; kc < kchunks; ++kc, ++cc) {
                    const uint32_t set = cc & 1u, ph = (cc >> 1) & 1u;
                    mbar_wait_u32(a_empty0 + set * 8u, ph ^ 1u, 51);
                    if (elect_one()) {
                        const uint32_t bar = a_full0 + set * 8u;            // one barrier per plane set
                        mbar_expect_tx_u32(bar, (uint32_t)NP * P.plane_tx);
#pragma unroll
                        for (int pl = 0; pl < NP; ++pl)
                            tma_load_5d_u32(a_ring_u32 + (set * NP + (uint32_t)pl) * V4T_SLOT, &P.a_map, bar, (int)(kc * 32),
                                            c.x0 - 1, c.y0 - 1, c.z0 - 1 + pl, c.n);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 3) {
        // =============================== weight-slab producer ===============================
        // one slab per (class, channel chunk, in-plane tap): the four kz blocks in ascending kz order
        uint32_t slot = 0, phase = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            for (int q = 0; q < 4; ++q) {
                for (uint32_t kc = 0; kc < kchunks; ++kc) {
                    for (int t4 = 0; t4 < 4; ++t4) {
                        mbar_wait_u32(b_empty0 + slot * 8u, phase ^ 1u, 52);
                        if (elect_one()) {
                            const uint32_t bar = b_full0 + slot * 8u;
                            mbar_expect_tx_u32(bar, 4u * P.blk_bytes);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                tma_load_3d_u32(b_ring_u32 + slot * b_bytes + (uint32_t)j * P.blk_bytes, &P.w_map, bar,
                                                (int)(kc * 32), 0, P.slab[q][t4][j]);
                        }
                        __syncwarp();
                        if (++slot == B_SLOTS) { slot = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        const uint64_t a_hi = (uint64_t)(uint32_t)(umma_desc(0, 16, P.sbo_a, 4) >> 32) << 32;
        const uint64_t b_hi = (uint64_t)(uint32_t)(umma_desc(0, 16, 512, 4) >> 32) << 32;
        const uint32_t lo_const = (uint32_t)(umma_desc(0, 16, 0, 4) & 0xFFFFFFFFu);
        const uint32_t id1 = P.idesc[0], id3 = P.idesc[2];
        const uint32_t blk16 = P.blk_bytes >> 4;
        uint32_t cc = 0, b_slot = 0, b_phase = 0, iter = 0;
        uint32_t b_ready = 0, a_ready = 0;      // look-ahead polls of the next slab / plane set (see conv_igemm4.cu)
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
#pragma unroll
            for (int q = 0; q < 4; ++q, ++iter) {
                const uint32_t acc = iter & 1u;
                mbar_wait_u32(smem_u32(&tempty[acc]), ((iter >> 1) & 1u) ^ 1u, 53);
                tc_fence_after();
                const uint32_t d_base = tmem_base + acc * 4u * NT;
                for (uint32_t kc = 0; kc < kchunks; ++kc, ++cc) {
                    const uint32_t set = cc & 1u, aph = (cc >> 1) & 1u;
                    const uint32_t a_set = a_ring_u32 + set * NP * V4T_SLOT;
#pragma unroll
                    for (int t4 = 0; t4 < 4; ++t4) {
                        if (!b_ready) mbar_wait_u32(b_full0 + b_slot * 8u, b_phase, 55);
                        if (t4 == 0 && !a_ready) mbar_wait_u32(a_full0 + set * 8u, aph, 54);
                        tc_fence_after();
                        {
                            const uint32_t nslot = b_slot + 1 == B_SLOTS ? 0u : b_slot + 1;
                            b_ready = mbar_test_wait_u32(b_full0 + nslot * 8u, b_slot + 1 == B_SLOTS ? b_phase ^ 1u : b_phase);
                            if (t4 == 3) a_ready = mbar_test_wait_u32(a_full0 + ((cc + 1) & 1u) * 8u, ((cc + 1) >> 1) & 1u);
                        }
                        // window start (dy + 1, dx + 1) inside a plane of PW = 10 columns: an immediate once q and t4 are unrolled
                        // (parity 0 reads offsets {0, -1}, parity 1 reads {+1, 0}: row_off in igemm4t_conv)
                        const int wy = ((q >> 1) ? ((t4 >> 1) ? 0 : 1) : ((t4 >> 1) ? -1 : 0)) + 1;
                        const int wx = ((q & 1) ? ((t4 & 1) ? 0 : 1) : ((t4 & 1) ? -1 : 0)) + 1;
                        const uint32_t a_lo = lo_const | (((a_set + (uint32_t)((wy * 10 + wx) * 64)) & 0x3FFFFu) >> 4);
                        const uint32_t b_lo = lo_const | (((b_ring_u32 + b_slot * b_bytes) & 0x3FFFFu) >> 4);
                        const bool first = (kc | (uint32_t)t4) == 0u;
                        if (elect_one()) {
                            // plane pl, k16 step k: accumulators lo .. lo+cnt-1 take slab blocks blk .. blk+cnt-1
                            auto issue = [&](int pl, int k, int lo, int cnt, int blk, bool accumulate) {
                                const uint64_t adesc = a_hi | (uint64_t)(a_lo + (uint32_t)pl * (V4T_SLOT >> 4) + (uint32_t)(2 * k));
                                const uint64_t bdesc = b_hi | (uint64_t)(b_lo + (uint32_t)blk * blk16 + (uint32_t)(2 * k));
                                mma_bf16(d_base + (uint32_t)lo * NT, adesc, bdesc, cnt == 1 ? id1 : id3, accumulate);
                            };
                            if (first) {
                                // a stacked MMA cannot mix "overwrite" and "accumulate" columns: every accumulator is first
                                // written by its own N = NT MMA (its dz = 0 tap), the other contributions of k = 0 follow singly
                                issue(1, 0, 0, 1, 1, false);
                                issue(1, 0, 1, 1, 2, false);
                                issue(2, 0, 2, 1, 1, false);
                                issue(2, 0, 3, 1, 2, false);
                                issue(0, 0, 0, 1, 3, true);
                                issue(1, 0, 2, 1, 3, true);
                                issue(2, 0, 1, 1, 0, true);
                                issue(3, 0, 3, 1, 0, true);
                                issue(0, 1, 0, 1, 3, true);
                                issue(1, 1, 0, 3, 1, true);
                                issue(2, 1, 1, 3, 0, true);
                                issue(3, 1, 3, 1, 0, true);
                            } else {
#pragma unroll
                                for (int k = 0; k < 2; ++k) {
                                    issue(0, k, 0, 1, 3, true);
                                    issue(1, k, 0, 3, 1, true);
                                    issue(2, k, 1, 3, 0, true);
                                    issue(3, k, 3, 1, 0, true);
                                }
                            }
                            mma_commit_u32(b_empty0 + b_slot * 8u);
                            if (t4 == 3) mma_commit_u32(a_empty0 + set * 8u);      // the whole plane set is free again
                        }
                        __syncwarp();
                        if (++b_slot == B_SLOTS) { b_slot = 0; b_phase ^= 1u; }
                    }
                }
                if (elect_one()) mma_commit_u32(smem_u32(&tfull[acc]));
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===============================
        const int qw = warp - 4;
        const int row = qw * 32 + lane;
        uint32_t iter = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            Unit4t c;
            v4t_decode(P, u, c);
            const int xi = c.x0 + (row & 7), yi = c.y0 + (row >> 3);
            const bool valid_xy = yi < P.iH && xi < P.iW;
            for (int q = 0; q < 4; ++q, ++iter) {
                const int py = q >> 1, px = q & 1;
                const uint32_t acc = iter & 1u;
                mbar_wait(&tfull[acc], (iter >> 1) & 1u, 56);
                tc_fence_after();
#pragma unroll 1
                for (int a = 0; a < 4; ++a) {
                    const int zi = c.z0 + (a >> 1), pz = a & 1;
                    if (zi >= P.iD) break;
                    bf16* yrow = P.y + (long)c.n * P.sN + (long)(2 * zi + pz) * P.sD + (long)(2 * yi + py) * P.sH +
                                 (long)(2 * xi + px) * P.sW;
                    const uint32_t t_addr = tmem_base + ((uint32_t)(qw * 32) << 16) + (acc * 4u + (uint32_t)a) * NT;
                    for (int col = 0; col < P.NT; col += 32) {
                        uint32_t r[32];
                        const bool wide = (P.NT - col) >= 32;
                        if (wide) tmem_ld_x32(t_addr + col, r);
                        else tmem_ld_x16(t_addr + col, r);
                        tmem_ld_wait();
                        const int ncol = wide ? 32 : 16;
                        if (valid_xy) {
                            const float4* bq = reinterpret_cast<const float4*>(s_bias + col);
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                if (j < ncol) {
                                    const float4 b0 = bq[j >> 2], b1 = bq[(j >> 2) + 1];
                                    uint4 o;
                                    o.x = pack2(__uint_as_float(r[j]) + b0.x, __uint_as_float(r[j + 1]) + b0.y);
                                    o.y = pack2(__uint_as_float(r[j + 2]) + b0.z, __uint_as_float(r[j + 3]) + b0.w);
                                    o.z = pack2(__uint_as_float(r[j + 4]) + b1.x, __uint_as_float(r[j + 5]) + b1.y);
                                    o.w = pack2(__uint_as_float(r[j + 6]) + b1.z, __uint_as_float(r[j + 7]) + b1.w);
                                    *reinterpret_cast<uint4*>(yrow + col + j) = o;
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, P.tmem_cols);
    }
}

typedef CUresult (*EncodeTiledFn4t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// returns 1 when handled, 0 when the shape is not for this kernel, <0 on error
int igemm4t_conv(const Plan& p, const amb_conv_args* a) {
    if (getenv("AMB_DISABLE_V4T")) return 0;
    if (a->op != AMB_OP_CONVT || a->k != 4 || a->stride != 2) return 0;
    if (p.Cx % 32 != 0 || p.Cy % 16 != 0 || p.Cy > 64 || p.n_taps != 64 || p.n_in_views != 1 || p.n_groups != 8) return 0;
    if (a->stats || a->ep_scale || a->ep_act || a->active) return 0;
    const View& iv = p.in_views[0];
    if (iv.H < 16 || iv.W < 8 || iv.D < 2) return 0;
    const int NT = p.Cy;
    // a persistent CTA runs the four in-plane classes of a unit back to back: with fewer units than SMs the per-tap kernel's
    // finer tiles fill the machine better (measured: 64->64 at 16^3 x 2: 0.031 ms here, 0.022 ms per tap)
    const long n_units = (long)iv.N * ceil_div(iv.H, 16) * ceil_div(iv.W, 8) * ceil_div(iv.D, 2);
    if (n_units < (long)num_sms() && !getenv("AMB_V4T_ALL")) return 0;

    static Igemm4tParams P;
    memset(&P, 0, sizeof(P));
    P.y = (bf16*)a->y;
    P.sW = p.Cy; P.sH = 2L * iv.W * p.Cy; P.sD = 2L * iv.H * P.sH; P.sN = 2L * iv.D * P.sD;
    P.bias = a->bias;
    P.iN = iv.N; P.iD = iv.D; P.iH = iv.H; P.iW = iv.W; P.Cy = p.Cy;
    P.Ty = ceil_div(iv.H, 16); P.Tx = ceil_div(iv.W, 8); P.Tzg = ceil_div(iv.D, 2);
    P.NT = NT; P.kchunks = p.Cx / 32;
    const int PW = 10;                                // 8 x + halo
    P.plane_tx = (uint32_t)(18 * PW * 64);
    P.sbo_a = (uint32_t)PW * 64u;
    P.blk_bytes = (uint32_t)NT * 64u;
    P.b_bytes = (4u * P.blk_bytes + 1023u) & ~1023u;
    P.tmem_cols = 32;
    while (P.tmem_cols < (uint32_t)(2 * 4 * NT)) P.tmem_cols <<= 1;
    for (int c = 1; c <= 3; ++c) P.idesc[c - 1] = umma_idesc_bf16(128, c * NT, 0, 0);
    // out index o = 2i - 1 + k: parity 0 ← (k 1, i = j), (k 3, i = j - 1); parity 1 ← (k 0, i = j + 1), (k 2, i = j)  (conv_plan.cuh)
    const int kk_of[2][2] = {{1, 3}, {0, 2}}, off_of[2][2] = {{0, -1}, {1, 0}};
    for (int q = 0; q < 4; ++q) {
        const int py = q >> 1, px = q & 1;
        for (int t4 = 0; t4 < 4; ++t4) {
            const int b = t4 >> 1, c = t4 & 1;
            const int dy = off_of[py][b], dx = off_of[px][c], ky = kk_of[py][b], kx = kk_of[px][c];
            P.row_off[q][t4] = (uint16_t)(((dy + 1) * PW + (dx + 1)) * 64);
            for (int kz = 0; kz < 4; ++kz) P.slab[q][t4][kz] = (int16_t)((kz * 4 + ky) * 4 + kx);
        }
    }

    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    AMB_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qr) == cudaSuccess &&
                  qr == cudaDriverEntryPointSuccess, AMB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    EncodeTiledFn4t enc = (EncodeTiledFn4t)fnp;
    {
        cuuint64_t dims[5] = {(cuuint64_t)p.Cx, (cuuint64_t)iv.W, (cuuint64_t)iv.H, (cuuint64_t)iv.D, (cuuint64_t)iv.N};
        cuuint64_t strides[4] = {(cuuint64_t)iv.sW * 2, (cuuint64_t)iv.sH * 2, (cuuint64_t)iv.sD * 2, (cuuint64_t)iv.sN * 2};
        cuuint32_t box[5] = {32, (cuuint32_t)PW, 18, 1, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&P.a_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)((const bf16*)a->x + iv.base), dims, strides,
                         box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AMB_CHECK(r == CUDA_SUCCESS, AMB_ERR_CUDA, "cuTensorMapEncodeTiled(halo plane, v4t) failed: %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Cx, (cuuint64_t)p.Cy, 64};
        cuuint64_t strides[2] = {(cuuint64_t)p.Cx * 2, (cuuint64_t)p.Cy * p.Cx * 2};
        cuuint32_t box[3] = {32, (cuuint32_t)NT, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&P.w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)a->w, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AMB_CHECK(r == CUDA_SUCCESS, AMB_ERR_CUDA, "cuTensorMapEncodeTiled(weight slab, v4t) failed: %d", (int)r);
    }
    const size_t fixed = 2u * (size_t)V4T_NP * V4T_SLOT + 1024 + 512 + (size_t)p.Cy * sizeof(float);
    int b_slots = V4T_B_SLOTS_MAX;
    while (b_slots > 2 && fixed + (size_t)b_slots * P.b_bytes > 227 * 1024) b_slots--;
    if (b_slots < 3 || fixed + (size_t)b_slots * P.b_bytes > 227 * 1024) return 0;
    P.b_slots = b_slots;
    const size_t smem = fixed + (size_t)b_slots * P.b_bytes;
    const long units = (long)P.iN * P.Ty * P.Tx * P.Tzg;
    const int grid = (int)(units < (long)num_sms() ? units : (long)num_sms());
    AMB_CUDA(cudaFuncSetAttribute(igemm4t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    igemm4t_kernel<<<grid, 256, smem, (cudaStream_t)a->stream>>>(P);
    AMB_LAUNCH_CHECK();
    g_last_conv_kernel = "igemm4t_kernel";
    return 1;
}

}  // namespace amb
