// tcgen05 implicit GEMM for dense 3×3×3 stride-1 convolutions with narrow outputs (Cout tile ≤ 64) — the layers that
// dominate the step (LightDecoder block 3: 64→64 and 64→32 at 128³) and that the per-tap kernel runs at the L2→SM
// limit (17–33e-3 B/FLOP; measured 690 / 390 TFLOP/s).
//
// Both measured limits are attacked at once:
//   * operand traffic — the input is staged as halo PLANES: one TMA box [18 y][16 x][32 channels] (18 KB, 64-byte
//     swizzle) per input z-plane and 32-channel chunk.  All 9 in-plane taps read it in place: the UMMA A-descriptor
//     starts (dy·16+dx) rows into the plane (rows are 64 B, 8-row groups 1 KB apart = SBO; the swizzle phase follows the
//     absolute smem address, so unaligned starts need no base offset — measured).
//   * dependent-MMA latency — a CTA owns FOUR z-adjacent 1×16×8 tiles (4 TMEM accumulators): the 6 planes z0−1…z0+4
//     serve all of them, every weight slab is used by 4 independent MMA chains, and the planes are released in dz order
//     so the next chunk's planes stream in underneath.
// Per 4 tiles and 64 input channels: 12 planes × 18 KB + 54 slabs × 4 KB = 432 KB (1.9e-3 B/FLOP, 9× less than per-tap).
//
//   warp 0  plane producer (ring of 8 × 18 KB)      warp 1  MMA issuer            warp 2  TMEM allocator
//   warp 3  weight-slab producer (ring of 8)        warps 4-7 epilogue (+bias, mask, Σ/Σ², bf16 stores)
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace amb {

using namespace ptx;

#define V3_PLANE_BYTES (18 * 16 * 64)      // 18 KB
#define V3_A_SLOTS_MAX 12              // plane ring depth is a launch parameter (6 planes in use + prefetch)
#define V3_B_SLOTS 8
#define V3_T 4

struct Igemm3Params {
    CUtensorMap a_map;                   // dims (C, W, H, D, N), box (32, 16, 18, 1, 1), SWIZZLE_64B
    CUtensorMap w_map;                   // dims (Cx, Cy, 27), box (32, NT, 1), SWIZZLE_64B
    bf16* y;
    long sN, sD, sH, sW;
    const float* bias;
    const uint8_t* active;
    const int* list;                     // active-patch work-list (patch edge >= 16 output voxels) or nullptr = dense
    const int* count;
    double* stats;
    int oN, oD, oH, oW, Cy;
    int lgPv, fd, fh, fw;
    int Ty, Tx, Tzg, n_ntiles, NT, kchunks, issuers, a_slots;
    uint32_t b_bytes, b_tx, tmem_cols, idesc;
    int8_t tap_dy[27], tap_dx[27];       // taps sorted by dz (9 per plane offset), values 0..2
    int16_t tap_w[27];
};

struct Unit3 {
    int nt, n, y0, x0, z0;
};

__device__ __forceinline__ void v3_decode(const Igemm3Params& P, uint32_t u, Unit3& c) {
    c.nt = (int)(u % (uint32_t)P.n_ntiles); u /= (uint32_t)P.n_ntiles;
    if (P.list) {
        // visible patches only: a patch of edge Pv holds (Pv/16) x (Pv/8) x (Pv/4) units of 16 x 8 x 4 voxels
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
        c.z0 = (int)((pz << P.lgPv) + iz * V3_T);
        c.y0 = (int)((py << P.lgPv) + iy * 16u);
        c.x0 = (int)((px << P.lgPv) + ix * 8u);
        return;
    }
    c.z0 = (int)(u % (uint32_t)P.Tzg) * V3_T; u /= (uint32_t)P.Tzg;
    c.x0 = (int)(u % (uint32_t)P.Tx) * 8; u /= (uint32_t)P.Tx;
    c.y0 = (int)(u % (uint32_t)P.Ty) * 16;
    c.n = (int)(u / (uint32_t)P.Ty);
}

__device__ __forceinline__ float v3_column_sums(float* v) {
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

__global__ void __launch_bounds__(256, 1) igemm3_kernel(const __grid_constant__ Igemm3Params P) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;
    const uint32_t V3_A_SLOTS = (uint32_t)P.a_slots;
    uint8_t* b_ring = smem + V3_A_SLOTS * V3_PLANE_BYTES;
    uint8_t* ctrl = b_ring + (size_t)V3_B_SLOTS * P.b_bytes;
    uint64_t* a_full = (uint64_t*)ctrl;            // [12]
    uint64_t* a_empty = a_full + V3_A_SLOTS_MAX;   // [12]
    uint64_t* b_full = a_empty + V3_A_SLOTS_MAX;   // [8]
    uint64_t* b_empty = b_full + 8;                // [8]
    uint64_t* tfull = b_empty + 8;                 // [2]
    uint64_t* tempty = tfull + 2;                  // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    float* s_stats = (float*)(ctrl + 512);
    float* s_bias = s_stats + (P.stats ? 2 * P.Cy : 0);   // [Cy] (zeros without a bias): the epilogue reads it as float4

    if (threadIdx.x == 0) {
        const uint32_t nI = (uint32_t)P.issuers;          // every issuing warp commits to the consumer-side barriers
        for (int s = 0; s < V3_A_SLOTS_MAX; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], nI); }
        for (int s = 0; s < V3_B_SLOTS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], nI); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], nI); mbar_init(&tempty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&P.a_map); prefetch_tmap(&P.w_map); }
    if (P.stats) for (int i = threadIdx.x; i < 2 * P.Cy; i += blockDim.x) s_stats[i] = 0.f;
    for (int i = threadIdx.x; i < P.Cy; i += blockDim.x) s_bias[i] = P.bias ? P.bias[i] : 0.f;
    if (warp == 2) tmem_alloc(tmem_slot, P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t nunits = P.list ? ((uint32_t)(*P.count) << (3 * P.lgPv - 9)) * (uint32_t)P.n_ntiles
                                   : (uint32_t)(P.oN * P.Ty * P.Tx * P.Tzg * P.n_ntiles);
    const uint32_t kchunks = (uint32_t)P.kchunks, NT = (uint32_t)P.NT, b_bytes = P.b_bytes;
    const uint32_t a_ring_u32 = smem_u32(a_ring), b_ring_u32 = smem_u32(b_ring);
    const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty);
    const uint32_t b_full0 = smem_u32(b_full), b_empty0 = smem_u32(b_empty);

    if (warp == 0) {
        // =============================== plane producer ===============================
        uint32_t slot = 0, phase = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            Unit3 c;
            v3_decode(P, u, c);
            for (uint32_t kc = 0; kc < kchunks; ++kc) {
                for (int pl = 0; pl < V3_T + 2; ++pl) {
                    mbar_wait_u32(a_empty0 + slot * 8u, phase ^ 1u, 31);
                    if (elect_one()) {
                        mbar_expect_tx_u32(a_full0 + slot * 8u, V3_PLANE_BYTES);
                        tma_load_5d_u32(a_ring_u32 + slot * V3_PLANE_BYTES, &P.a_map, a_full0 + slot * 8u, (int)(kc * 32),
                                        c.x0 - 1, c.y0 - 1, c.z0 - 1 + pl, c.n);
                    }
                    __syncwarp();
                    if (++slot == V3_A_SLOTS) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        // =============================== weight-slab producer ===============================
        uint32_t slot = 0, phase = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
            Unit3 c;
            v3_decode(P, u, c);
            const int ncol = c.nt * (int)NT;
            for (uint32_t kc = 0; kc < kchunks; ++kc) {
                for (int t = 0; t < 27; ++t) {
                    mbar_wait_u32(b_empty0 + slot * 8u, phase ^ 1u, 32);
                    if (elect_one()) {
                        mbar_expect_tx_u32(b_full0 + slot * 8u, P.b_tx);
                        tma_load_3d_u32(b_ring_u32 + slot * b_bytes, &P.w_map, b_full0 + slot * 8u, (int)(kc * 32), ncol,
                                        P.tap_w[t]);
                    }
                    __syncwarp();
                    if (++slot == V3_B_SLOTS) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1 || (warp == 2 && P.issuers == 2)) {
        // =============================== MMA issuers ===============================
        // One elected thread sustains about one tcgen05.mma per 45 ns, slower than the tensor pipe at N <= 64; with two
        // issuing warps each owns half of the V3_T tiles (independent accumulators).
        const int t_begin = P.issuers == 2 ? (warp == 1 ? 0 : V3_T / 2) : 0;
        const int t_end = P.issuers == 2 ? t_begin + V3_T / 2 : V3_T;
        // K-major, 64-byte swizzle: layout 4.  A: 8-row groups (x) 1024 B apart (one y step of the plane).  B: 512 B.
        const uint32_t a_hi = (uint32_t)(umma_desc(0, 16, 1024, 4) >> 32), b_hi = (uint32_t)(umma_desc(0, 16, 512, 4) >> 32);
        const uint32_t lo_const = (uint32_t)(umma_desc(0, 16, 0, 4) & 0xFFFFFFFFu);
        uint32_t a_slot0 = 0, a_phase0 = 0;     // ring position of plane 0 of the current chunk
        uint32_t b_slot = 0, b_phase = 0, iter = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x, ++iter) {
            const uint32_t acc = iter & 1u;
            mbar_wait_u32(smem_u32(&tempty[acc]), ((iter >> 1) & 1u) ^ 1u, 33);
            tc_fence_after();
            const uint32_t d_base = tmem_base + acc * V3_T * NT;
            for (uint32_t kc = 0; kc < kchunks; ++kc) {
                int planes_ready = 0;
                for (int dz = 0; dz < 3; ++dz) {
                    // taps of this dz read planes dz .. dz+T-1 (tile t reads plane t+dz)
                    while (planes_ready < dz + V3_T) {
                        uint32_t s = a_slot0 + (uint32_t)planes_ready, ph = a_phase0;
                        if (s >= V3_A_SLOTS) { s -= V3_A_SLOTS; ph ^= 1u; }
                        mbar_wait_u32(a_full0 + s * 8u, ph, 34);
                        ++planes_ready;
                    }
                    tc_fence_after();
                    for (int t9 = 0; t9 < 9; ++t9) {
                        const int tap = dz * 9 + t9;
                        mbar_wait_u32(b_full0 + b_slot * 8u, b_phase, 35);
                        tc_fence_after();
                        const uint32_t row_off = (uint32_t)(P.tap_dy[tap] * 16 + P.tap_dx[tap]) * 64u;
                        const uint32_t b_lo = lo_const | (((b_ring_u32 + b_slot * b_bytes) & 0x3FFFFu) >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
#pragma unroll
                                for (int t = 0; t < V3_T; ++t) {
                                    if (t < t_begin || t >= t_end) continue;
                                    uint32_t s = a_slot0 + (uint32_t)(t + dz);
                                    if (s >= V3_A_SLOTS) s -= V3_A_SLOTS;
                                    const uint32_t a_addr = a_ring_u32 + s * V3_PLANE_BYTES + row_off;
                                    const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(lo_const | ((a_addr & 0x3FFFFu) >> 4)) ;
                                    const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)b_lo;
                                    mma_bf16(d_base + (uint32_t)t * NT, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), P.idesc,
                                             (kc | (uint32_t)tap | (uint32_t)k) != 0u);
                                }
                            }
                            mma_commit_u32(b_empty0 + b_slot * 8u);
                            // plane dz is dead after the dz taps (tile 0 was its last reader); after dz = 2 all remaining go
                            if (t9 == 8) {
                                const int first = dz, last = dz == 2 ? V3_T + 1 : dz;
                                for (int pl = first; pl <= last; ++pl) {
                                    uint32_t s = a_slot0 + (uint32_t)pl;
                                    if (s >= V3_A_SLOTS) s -= V3_A_SLOTS;
                                    mma_commit_u32(a_empty0 + s * 8u);
                                }
                            }
                        }
                        __syncwarp();
                        if (++b_slot == V3_B_SLOTS) { b_slot = 0; b_phase ^= 1u; }
                    }
                }
                a_slot0 += V3_T + 2;
                if (a_slot0 >= V3_A_SLOTS) { a_slot0 -= V3_A_SLOTS; a_phase0 ^= 1u; }
            }
            if (elect_one()) mma_commit_u32(smem_u32(&tfull[acc]));
            __syncwarp();
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===============================
        const int q = warp - 4;
        const int row = q * 32 + lane;
        uint32_t iter = 0;
        for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x, ++iter) {
            Unit3 c;
            v3_decode(P, u, c);
            const uint32_t acc = iter & 1u;
            const int x = c.x0 + (row & 7), y = c.y0 + (row >> 3);
            const bool valid_xy = y < P.oH && x < P.oW;
            mbar_wait(&tfull[acc], (iter >> 1) & 1u, 36);
            tc_fence_after();
            for (int t = 0; t < V3_T; ++t) {
                const int z = c.z0 + t;
                if (z >= P.oD) break;
                bool on = valid_xy;
                if (valid_xy && P.active && P.lgPv >= 0)
                    on = P.active[((c.n * P.fd + (z >> P.lgPv)) * P.fh + (y >> P.lgPv)) * P.fw + (x >> P.lgPv)] != 0;
                bf16* yrow = P.y + (long)c.n * P.sN + (long)z * P.sD + (long)y * P.sH + (long)x * P.sW + (long)c.nt * P.NT;
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * V3_T + (uint32_t)t) * NT;
                for (int col = 0; col < P.NT; col += 32) {
                    uint32_t r[32];
                    const bool wide = (P.NT - col) >= 32;
                    if (wide) tmem_ld_x32(t_addr + col, r);
                    else tmem_ld_x16(t_addr + col, r);
                    tmem_ld_wait();
                    const int ncol = wide ? 32 : 16;
                    float v[32];
                    const float4* bq = reinterpret_cast<const float4*>(s_bias + c.nt * P.NT + col);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (j < ncol) b4 = bq[j >> 2];
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            float f = 0.f;
                            if (j < ncol && on) f = __uint_as_float(r[j + jj]) + bb[jj];
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
                        float s1 = v3_column_sums(v);
                        float s2 = v3_column_sums(sq);
                        if (lane < ncol) {
                            atomicAdd(&s_stats[c.nt * P.NT + col + lane], s1);
                            atomicAdd(&s_stats[P.Cy + c.nt * P.NT + col + lane], s2);
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

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// returns 1 when handled, 0 when the shape is not for this kernel, <0 on error
int igemm3_conv(const Plan& p, const amb_conv_args* a) {
    const char* dis = getenv("AMB_DISABLE_V3");
    if (dis && atoi(dis) == 1) return 0;
    if (a->ep_scale || a->ep_act) return 0;              // fused output transforms: conv_igemm4.cu / conv_igemm.cu
    if (!((a->op == AMB_OP_CONV || a->op == AMB_OP_CONV_DGRAD) && a->k == 3 && a->stride == 1)) return 0;
    if (p.Cx % 32 != 0 || p.Cy % 16 != 0 || p.n_taps != 27 || p.n_in_views != 1 || p.n_groups != 1) return 0;
    if (p.oH < 16 || p.oW < 8 || p.oD < V3_T) return 0;
    // active-patch work-list: usable when a patch holds whole 16 x 8 x 4 units; with a finer grid the per-tap kernel's list
    // (2 x 8 x 8 tiles) still pays, so leave those layers to it
    const bool use_list = a->active_list != nullptr && p.lgPv >= 4 && !getenv("AMB_V3_NO_LIST");
    if (a->active_list != nullptr && p.lgPv >= 3 && !use_list) return 0;
    int NT = 0;
    for (int nt = 64; nt >= 16; nt -= 16) if (p.Cy % nt == 0) { NT = nt; break; }
    if (NT == 0 || p.Cy / NT > 1) return 0;          // one N tile: wider layers are already tensor-bound in the per-tap kernel
    if (a->stats && p.Cy > 2048) return 0;

    static Igemm3Params P;
    memset(&P, 0, sizeof(P));
    P.y = (bf16*)a->y;
    const View& ov = p.out_views[0];
    P.sN = ov.sN; P.sD = ov.sD; P.sH = ov.sH; P.sW = ov.sW;
    P.bias = a->bias; P.active = a->active; P.stats = a->stats;
    P.list = use_list ? a->active_list : nullptr;
    P.count = use_list ? a->active_count : nullptr;
    P.oN = p.oN; P.oD = p.oD; P.oH = p.oH; P.oW = p.oW; P.Cy = p.Cy;
    P.lgPv = p.lgPv; P.fd = p.fd; P.fh = p.fh; P.fw = p.fw;
    P.Ty = ceil_div(p.oH, 16); P.Tx = ceil_div(p.oW, 8); P.Tzg = ceil_div(p.oD, V3_T);
    P.NT = NT; P.n_ntiles = p.Cy / NT; P.kchunks = p.Cx / 32;
    const char* ienv = getenv("AMB_V3_ISSUERS");
    P.issuers = (ienv && atoi(ienv) == 1) ? 1 : 2;
    P.b_tx = (uint32_t)NT * 64u;
    P.b_bytes = (P.b_tx + 1023u) & ~1023u;
    P.tmem_cols = 32;
    while (P.tmem_cols < (uint32_t)(2 * V3_T * NT)) P.tmem_cols <<= 1;
    P.idesc = umma_idesc_bf16(128, NT, 0, 0);
    int n = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int t = 0; t < 27; ++t)
            if (p.taps[t].dz == dz) {
                P.tap_dy[n] = (int8_t)(p.taps[t].dy + 1);
                P.tap_dx[n] = (int8_t)(p.taps[t].dx + 1);
                P.tap_w[n] = p.taps[t].w;
                ++n;
            }
    if (n != 27) return 0;

    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    AMB_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) == cudaSuccess &&
                  q == cudaDriverEntryPointSuccess, AMB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    EncodeTiledFn3 enc = (EncodeTiledFn3)fnp;
    {
        const View& iv = p.in_views[0];
        cuuint64_t dims[5] = {(cuuint64_t)p.Cx, (cuuint64_t)iv.W, (cuuint64_t)iv.H, (cuuint64_t)iv.D, (cuuint64_t)iv.N};
        cuuint64_t strides[4] = {(cuuint64_t)iv.sW * 2, (cuuint64_t)iv.sH * 2, (cuuint64_t)iv.sD * 2, (cuuint64_t)iv.sN * 2};
        cuuint32_t box[5] = {32, 16, 18, 1, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&P.a_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)((const bf16*)a->x + iv.base), dims, strides,
                         box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AMB_CHECK(r == CUDA_SUCCESS, AMB_ERR_CUDA, "cuTensorMapEncodeTiled(halo plane) failed: %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Cx, (cuuint64_t)p.Cy, 27};
        cuuint64_t strides[2] = {(cuuint64_t)p.Cx * 2, (cuuint64_t)p.Cy * p.Cx * 2};
        cuuint32_t box[3] = {32, (cuuint32_t)NT, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&P.w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)a->w, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AMB_CHECK(r == CUDA_SUCCESS, AMB_ERR_CUDA, "cuTensorMapEncodeTiled(weight slab) failed: %d", (int)r);
    }
    // plane ring: 6 planes are in use per (unit, channel chunk) and four of them are released together at its end, so
    // with 8 slots the next chunk's planes 2..5 only start loading then (a ~1 us bubble per chunk); 10 slots hide it
    const char* senv = getenv("AMB_V3_A_SLOTS");
    P.a_slots = senv ? atoi(senv) : 10;
    if (P.a_slots < 7 || P.a_slots > V3_A_SLOTS_MAX) P.a_slots = 10;
    while (P.a_slots > 7 && (size_t)P.a_slots * V3_PLANE_BYTES + (size_t)V3_B_SLOTS * P.b_bytes + 1024 + 512 +
                                    (a->stats ? 2 * (size_t)p.Cy * sizeof(float) : 0) + (size_t)p.Cy * sizeof(float) > 227 * 1024)
        P.a_slots--;
    size_t smem = (size_t)P.a_slots * V3_PLANE_BYTES + (size_t)V3_B_SLOTS * P.b_bytes + 1024 + 512 +
                  (a->stats ? 2 * (size_t)p.Cy * sizeof(float) : 0) + (size_t)p.Cy * sizeof(float);
    if (smem > 227 * 1024) return 0;
    long units = use_list ? ((long)p.oN * p.fd * p.fh * p.fw << (3 * p.lgPv - 9)) * P.n_ntiles
                          : (long)p.oN * P.Ty * P.Tx * P.Tzg * P.n_ntiles;
    int grid = (int)(units < (long)num_sms() ? units : (long)num_sms());
    AMB_CUDA(cudaFuncSetAttribute(igemm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    igemm3_kernel<<<grid, 256, smem, (cudaStream_t)a->stream>>>(P);
    AMB_LAUNCH_CHECK();
    g_last_conv_kernel = "igemm3_kernel";
    return 1;
}

}  // namespace amb
