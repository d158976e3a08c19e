// tcgen05 implicit-GEMM for the dominant case — dense 3×3×3 stride-1 convolutions (forward and input-gradient) with
// Cin % 64 == 0 — built around the measured limit of v1: the L2→SM fabric (~6.3 KB/clk chip-wide, ~44 B/clk/SM).
// v1 re-fetches the 128-voxel input window for every one of the 27 taps (16 KB + 8 KB weights per 128×64×64 k-block
// ≈ 23e-3 B/FLOP against a 5.4e-3 budget → 23 % of tensor peak, which is what it measures).
//
// Here the input is staged as halo *slices*: one TMA box [18 rows y][16 columns x][64 channels] per input z-plane lands
// in smem once, and all 9 in-plane taps (dy,dx) of that plane read it in place — the UMMA A-descriptor simply starts
// (dy·16+dx) rows further into the slice (rows are 128 B, 8-row groups 2 KB apart = SBO, the 128-byte swizzle phase of
// the unaligned start is carried in the descriptor's base-offset field).  A tile is 1×16×8 output voxels (M = 128); a
// CTA walks a column of tiles along z, so with one channel chunk consecutive tiles share two of their three planes and
// only ONE new 36 KB plane is fetched per tile (27×16 KB → 36 KB of A traffic per tile).
//
//   warp 0  A producer (halo planes, ring of 4 × 36 KB)       warp 1  MMA issuer
//   warp 3  B producer (weight slabs [NT][64], own ring)      warp 2  TMEM allocator          warps 4-7 epilogue
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace amb {

using namespace ptx;

int encode_weight_map(CUtensorMap* m, const void* w, int T, int rows, int cols, int kc, int nt);

#define V2_SLICE_ROWS (18 * 16)
#define V2_SLICE_BYTES (V2_SLICE_ROWS * 128)      // 36 KB
#define V2_A_SLOTS 4

struct Igemm2Params {
    CUtensorMap a_map;
    CUtensorMap w_map;
    bf16* y;
    long sN, sD, sH, sW;
    const float* bias;
    const uint8_t* active;
    double* stats;
    int oN, oD, oH, oW, Cy;
    int lgPv, fd, fh, fw;
    int Ty, Tx, LZ, nseg, n_ntiles, NT, kchunks, b_stages;
    uint32_t b_bytes, tmem_cols, idesc;
    int bo_mode;
    int8_t tap_dy[27], tap_dx[27];      // taps sorted by dz (9 per plane), values 0..2
    int16_t tap_w[27];
};

__device__ __forceinline__ long v2_num_units(const Igemm2Params& P) {
    return (long)P.oN * P.Ty * P.Tx * P.nseg * P.n_ntiles;
}

struct Unit {
    int nt, n, y0, x0, z0, z1;
};

__device__ __forceinline__ void v2_decode(const Igemm2Params& P, long u, Unit& c) {
    c.nt = (int)(u % P.n_ntiles); u /= P.n_ntiles;
    int seg = (int)(u % P.nseg); u /= P.nseg;
    c.x0 = (int)(u % P.Tx) * 8; u /= P.Tx;
    c.y0 = (int)(u % P.Ty) * 16;
    c.n = (int)(u / P.Ty);
    c.z0 = seg * P.LZ;
    c.z1 = c.z0 + P.LZ < P.oD ? c.z0 + P.LZ : P.oD;
}

__device__ __forceinline__ float v2_warp_column_sums(float* v) {
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

__global__ void __launch_bounds__(256, 1) igemm2_kernel(const __grid_constant__ Igemm2Params P) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + V2_A_SLOTS * V2_SLICE_BYTES;
    uint8_t* ctrl = b_ring + (size_t)P.b_stages * P.b_bytes;
    uint64_t* a_full = (uint64_t*)ctrl;            // [4]
    uint64_t* a_empty = a_full + 4;                // [4]
    uint64_t* b_full = a_empty + 4;                // [8]
    uint64_t* b_empty = b_full + 8;                // [8]
    uint64_t* tfull = b_empty + 8;                 // [2]
    uint64_t* tempty = tfull + 2;                  // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    float* s_stats = (float*)(ctrl + 256);

    if (threadIdx.x == 0) {
        for (int s = 0; s < V2_A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < P.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&P.a_map); prefetch_tmap(&P.w_map); }
    if (P.stats) for (int i = threadIdx.x; i < 2 * P.Cy; i += blockDim.x) s_stats[i] = 0.f;
    if (warp == 2) tmem_alloc(tmem_slot, P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long nunits = v2_num_units(P);
    const bool rolling = P.kchunks == 1;

    if (warp == 0) {
        // =============================== A producer: halo planes ===============================
        if (lane == 0) {
            uint32_t li = 0;
            for (long u = blockIdx.x; u < nunits; u += gridDim.x) {
                Unit c;
                v2_decode(P, u, c);
                auto load = [&](int zs, int kc) {
                    const int s = li % V2_A_SLOTS;
                    mbar_wait(&a_empty[s], ((li / V2_A_SLOTS) & 1) ^ 1, 21);
                    mbar_expect_tx(&a_full[s], V2_SLICE_BYTES);
                    tma_load_5d(a_ring + (size_t)s * V2_SLICE_BYTES, &P.a_map, &a_full[s], kc * 64, c.x0 - 1, c.y0 - 1, zs, c.n);
                    ++li;
                };
                if (rolling) {
                    for (int zs = c.z0 - 1; zs <= c.z1; ++zs) load(zs, 0);
                } else {
                    for (int z = c.z0; z < c.z1; ++z)
                        for (int kc = 0; kc < P.kchunks; ++kc)
                            for (int dz = 0; dz < 3; ++dz) load(z + dz - 1, kc);
                }
            }
        }
    } else if (warp == 3) {
        // =============================== B producer: weight slabs ===============================
        if (lane == 0) {
            uint32_t bi = 0;
            for (long u = blockIdx.x; u < nunits; u += gridDim.x) {
                Unit c;
                v2_decode(P, u, c);
                for (int z = c.z0; z < c.z1; ++z)
                    for (int kc = 0; kc < P.kchunks; ++kc)
                        for (int t = 0; t < 27; ++t, ++bi) {
                            const int s = bi % P.b_stages;
                            mbar_wait(&b_empty[s], ((bi / P.b_stages) & 1) ^ 1, 22);
                            mbar_expect_tx(&b_full[s], P.b_bytes);
                            tma_load_3d(b_ring + (size_t)s * P.b_bytes, &P.w_map, &b_full[s], kc * 64, c.nt * P.NT, P.tap_w[t]);
                        }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            uint32_t bi = 0, tile_iter = 0, unit_base = 0, ready = 0;
            for (long u = blockIdx.x; u < nunits; u += gridDim.x) {
                Unit c;
                v2_decode(P, u, c);
                const int nz = c.z1 - c.z0;
                for (int z = c.z0; z < c.z1; ++z, ++tile_iter) {
                    const int acc = tile_iter & 1;
                    mbar_wait(&tempty[acc], ((tile_iter >> 1) & 1) ^ 1, 23);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P.NT);
                    bool first = true;
                    for (int kc = 0; kc < P.kchunks; ++kc) {
                        for (int dz = 0; dz < 3; ++dz) {
                            const uint32_t idx = rolling ? unit_base + (uint32_t)(z - c.z0 + dz)
                                                         : unit_base + (uint32_t)(((z - c.z0) * P.kchunks + kc) * 3 + dz);
                            while (ready <= idx) {
                                mbar_wait(&a_full[ready % V2_A_SLOTS], (ready / V2_A_SLOTS) & 1, 24);
                                ++ready;
                            }
                            tc_fence_after();
                            const uint32_t a_slot = smem_u32(a_ring + (size_t)(idx % V2_A_SLOTS) * V2_SLICE_BYTES);
                            for (int t = dz * 9; t < dz * 9 + 9; ++t, ++bi) {
                                const int s = bi % P.b_stages;
                                mbar_wait(&b_full[s], (bi / P.b_stages) & 1, 25);
                                tc_fence_after();
                                const uint32_t a_addr = a_slot + (uint32_t)(P.tap_dy[t] * 16 + P.tap_dx[t]) * 128u;
                                uint64_t adesc = umma_desc(a_addr, 16, 2048, 2);
                                if (P.bo_mode == 0) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;   // swizzle phase of the start row
                                const uint64_t bdesc = umma_desc(smem_u32(b_ring + (size_t)s * P.b_bytes), 16, 1024, 2);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), P.idesc, !first);
                                    first = false;
                                }
                                mma_commit(&b_empty[s]);
                            }
                            // a plane is dead once the last tile that reads it has issued its taps
                            if (!rolling || dz == 0 || z == c.z1 - 1) mma_commit(&a_empty[idx % V2_A_SLOTS]);
                        }
                    }
                    mma_commit(&tfull[acc]);
                }
                unit_base += rolling ? (uint32_t)(nz + 2) : (uint32_t)(nz * P.kchunks * 3);
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===============================
        const int q = warp - 4;
        const int row = q * 32 + lane;
        uint32_t tile_iter = 0;
        for (long u = blockIdx.x; u < nunits; u += gridDim.x) {
            Unit c;
            v2_decode(P, u, c);
            const int x = c.x0 + (row & 7), y = c.y0 + (row >> 3);
            const bool valid_xy = y < P.oH && x < P.oW;
            for (int z = c.z0; z < c.z1; ++z, ++tile_iter) {
                const int acc = tile_iter & 1;
                bool on = valid_xy;
                if (valid_xy && P.active && P.lgPv >= 0)
                    on = P.active[((c.n * P.fd + (z >> P.lgPv)) * P.fh + (y >> P.lgPv)) * P.fw + (x >> P.lgPv)] != 0;
                bf16* yrow = P.y + (long)c.n * P.sN + (long)z * P.sD + (long)y * P.sH + (long)x * P.sW + (long)c.nt * P.NT;
                mbar_wait(&tfull[acc], (tile_iter >> 1) & 1, 26);
                tc_fence_after();
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * P.NT);
                for (int col = 0; col < P.NT; col += 32) {
                    uint32_t r[32];
                    const bool wide = (P.NT - col) >= 32;
                    if (wide) tmem_ld_x32(t_addr + col, r);
                    else tmem_ld_x16(t_addr + col, r);
                    tmem_ld_wait();
                    const int ncol = wide ? 32 : 16;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float f = 0.f;
                        if (j < ncol) {
                            f = __uint_as_float(r[j]);
                            if (P.bias) f += __ldg(P.bias + c.nt * P.NT + col + j);
                            if (!on) f = 0.f;
                        }
                        v[j] = f;
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
                        float s1 = v2_warp_column_sums(v);
                        float s2 = v2_warp_column_sums(sq);
                        if (lane < ncol) {
                            atomicAdd(&s_stats[c.nt * P.NT + col + lane], s1);
                            atomicAdd(&s_stats[P.Cy + c.nt * P.NT + col + lane], s2);
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

static int pow2_ceil2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// returns 1 when handled, 0 when the shape is not for this kernel, <0 on error
int igemm2_conv(const Plan& p, const amb_conv_args* a) {
    // opt-in: measured no faster than the T-interleaved per-tap kernel (the limit was MMA latency, not L2 traffic)
    const char* en = getenv("AMB_ENABLE_V2");
    if (!(en && atoi(en) == 1)) return 0;
    if (!((a->op == AMB_OP_CONV || a->op == AMB_OP_CONV_DGRAD) && a->k == 3 && a->stride == 1)) return 0;
    if (p.Cx % 64 != 0 || p.Cy % 16 != 0 || p.n_taps != 27 || p.n_in_views != 1 || p.n_groups != 1) return 0;
    if (p.oH < 16 || p.oW < 8 || p.oD < 2) return 0;
    int NT = 0;
    for (int nt = 256; nt >= 16; nt -= 16) if (p.Cy % nt == 0) { NT = nt; break; }
    if (NT == 0) return 0;
    if (a->stats && p.Cy > 2048) return 0;

    static Igemm2Params P;
    memset(&P, 0, sizeof(P));
    const char* bo = getenv("AMB_V2_BO_MODE");
    P.bo_mode = bo ? atoi(bo) : 1;     // measured: the swizzle phase comes from absolute smem address bits; field stays 0
    P.y = (bf16*)a->y;
    const View& ov = p.out_views[0];
    P.sN = ov.sN; P.sD = ov.sD; P.sH = ov.sH; P.sW = ov.sW;
    P.bias = a->bias; P.active = a->active; P.stats = a->stats;
    P.oN = p.oN; P.oD = p.oD; P.oH = p.oH; P.oW = p.oW; P.Cy = p.Cy;
    P.lgPv = p.lgPv; P.fd = p.fd; P.fh = p.fh; P.fw = p.fw;
    P.Ty = ceil_div(p.oH, 16); P.Tx = ceil_div(p.oW, 8);
    P.NT = NT; P.n_ntiles = p.Cy / NT; P.kchunks = p.Cx / 64;
    P.b_bytes = (uint32_t)NT * 128u;
    P.b_bytes = (P.b_bytes + 1023u) & ~1023u;
    int bs = (int)((76u * 1024u) / P.b_bytes);
    if (bs > 8) bs = 8;
    if (bs < 2) return 0;
    P.b_stages = bs;
    P.tmem_cols = (uint32_t)pow2_ceil2(2 * NT);
    if (P.tmem_cols < 32) P.tmem_cols = 32;
    P.idesc = umma_idesc_bf16(128, NT, 0, 0);
    // column segments along z: balance SM fill against the 2 extra planes a segment costs
    long cols = (long)p.oN * P.Ty * P.Tx * P.n_ntiles;
    int best = p.oD;
    double best_eff = -1.0;
    for (int lz = 4; lz <= p.oD; lz *= 2) {
        long units = cols * ceil_div(p.oD, lz);
        long waves = (units + num_sms() - 1) / num_sms();
        double eff = (double)units / (double)(waves * num_sms()) * (P.kchunks == 1 ? (double)lz / (lz + 2) : 1.0);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = lz; }
    }
    P.LZ = best;
    P.nseg = ceil_div(p.oD, P.LZ);
    // taps: 9 per input plane, planes in ascending dz (plan offsets are −1..1)
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

    // halo-plane map: dims (C, W, H, D, N), box (64, 16, 18, 1, 1), 128-byte swizzle, OOB → 0 (the conv's zero padding)
    {
        void* fnp = nullptr;
        cudaDriverEntryPointQueryResult q;
        AMB_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) == cudaSuccess &&
                      q == cudaDriverEntryPointSuccess, AMB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        EncodeTiledFn2 enc = (EncodeTiledFn2)fnp;
        const View& iv = p.in_views[0];
        cuuint64_t dims[5] = {(cuuint64_t)p.Cx, (cuuint64_t)iv.W, (cuuint64_t)iv.H, (cuuint64_t)iv.D, (cuuint64_t)iv.N};
        cuuint64_t strides[4] = {(cuuint64_t)iv.sW * 2, (cuuint64_t)iv.sH * 2, (cuuint64_t)iv.sD * 2, (cuuint64_t)iv.sN * 2};
        cuuint32_t box[5] = {64, 16, 18, 1, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&P.a_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)((const bf16*)a->x + iv.base), dims, strides,
                         box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AMB_CHECK(r == CUDA_SUCCESS, AMB_ERR_CUDA, "cuTensorMapEncodeTiled(halo plane) failed: %d", (int)r);
    }
    if (int e = encode_weight_map(&P.w_map, a->w, 27, p.Cy, p.Cx, 64, NT)) return e;

    size_t smem = (size_t)V2_A_SLOTS * V2_SLICE_BYTES + (size_t)P.b_stages * P.b_bytes + 1024 + 256 +
                  (a->stats ? 2 * (size_t)p.Cy * sizeof(float) : 0);
    if (smem > 227 * 1024) return 0;
    long units = cols * P.nseg;
    int grid = (int)(units < (long)num_sms() ? units : (long)num_sms());
    AMB_CUDA(cudaFuncSetAttribute(igemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    igemm2_kernel<<<grid, 256, smem, (cudaStream_t)a->stream>>>(P);
    AMB_LAUNCH_CHECK();
    return 1;
}

}  // namespace amb
