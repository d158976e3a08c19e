// tcgen05 / TMEM / TMA implicit-GEMM execution of a gather-GEMM plan (conv_plan.cuh).
//
//   tile      : 128 output voxels (a box bn×bd×bh×bw of one out view) × NT output channels
//   K loop    : for every tap of the tile's group, for every KC-channel slab of the gathered tensor:
//                 A = TMA box load of the tap-shifted input window  → smem [128 rows][KC] (K-major, 128/64/32B swizzle)
//                 B = TMA load of W[tap][nt*NT .. +NT][slab]         → smem [NT rows][KC]
//                 KC/16 × tcgen05.mma (M=128, N=NT, K=16), fp32 accumulators in TMEM
//               zero padding and stride-2 / transposed-conv geometry come for free from TMA out-of-bounds fill and
//               parity-class tensor maps — there is no im2col buffer.
//   roles     : warp 0 TMA producer · warp 1 MMA issuer · warp 2 TMEM allocator · warps 4-7 epilogue
//   pipelines : smem ring (full/empty mbarriers) and a double-buffered TMEM accumulator (tmem_full/tmem_empty), so the
//               epilogue of tile i overlaps the main loop of tile i+1; CTAs are persistent (one per SM) and walk the
//               tile list with a static stride.
//   epilogue  : TMEM → registers (tcgen05.ld 32x32b) → +bias → mask → bf16 → 128-bit global stores; optional
//               per-channel Σy / Σy² (warp transpose-reduce → smem → one fp64 atomic per channel per CTA).
//   sparsity  : with an active-patch work-list only tiles inside visible patches are enumerated (stage 0/1 of the
//               encoder, where a 2×8×8 tile fits a patch); otherwise tiles are dense and the epilogue zeroes masked rows.
#include "conv_plan.cuh"
#include "ptx.cuh"

namespace amb {

using namespace ptx;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static CUtensorMapSwizzle swizzle_for(int kc) {
    return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// rank-5 map over a view: dims (C, W, H, D, N), box (kc, bw, bh, bd, bn)
int encode_view_map(CUtensorMap* m, const void* tensor_base, const View& v, int C, int kc, const int box[4]) {
    EncodeTiledFn enc = get_encode();
    AMB_CHECK(enc != nullptr, AMB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.D, (cuuint64_t)v.N};
    cuuint64_t strides[4] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sD * 2, (cuuint64_t)v.sN * 2};
    cuuint32_t bx[5] = {(cuuint32_t)kc, (cuuint32_t)box[3], (cuuint32_t)box[2], (cuuint32_t)box[1], (cuuint32_t)box[0]};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    void* base = (void*)((const bf16*)tensor_base + v.base);
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMB_CHECK(r == CUDA_SUCCESS, AMB_ERR_CUDA,
              "cuTensorMapEncodeTiled(view) failed: %d (C=%d dims %d,%d,%d,%d box %d,%d,%d,%d kc=%d)", (int)r, C, v.N,
              v.D, v.H, v.W, box[0], box[1], box[2], box[3], kc);
    return 0;
}

// rank-3 map over packed weights [T][rows][cols]: dims (cols, rows, T), box (kc, nt, 1)
int encode_weight_map(CUtensorMap* m, const void* w, int T, int rows, int cols, int kc, int nt) {
    EncodeTiledFn enc = get_encode();
    AMB_CHECK(enc != nullptr, AMB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)T};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
    cuuint32_t bx[3] = {(cuuint32_t)kc, (cuuint32_t)nt, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)w, dims, strides, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMB_CHECK(r == CUDA_SUCCESS, AMB_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d (T=%d rows=%d cols=%d)", (int)r, T,
              rows, cols);
    return 0;
}

struct IgemmParams {
    CUtensorMap in_maps[8];
    CUtensorMap w_map;
    Plan plan;
    bf16* y;
    const float* bias;
    const uint8_t* active;
    const int* list;
    const int* count;
    double* stats;
    const float* ep_scale;               // optional fused epilogue y = act(acc·scale + bias) (inference-mode BN folded in)
    int ep_act;
    int lgbn, lgbd, lgbh, lgbw;      // log2 of the tile box
    int Tn, Tz, Ty, Tx;              // dense tiling of an out view
    int n_ntiles, NT;
    int T;                           // spatial tiles (independent TMEM accumulators) interleaved per CTA iteration
    int kchunks;                     // Cx / KC
    int ksplit;                      // > 1: the taps of a tile are split over ksplit work items (small spatial extents), fp32
    float* ws;                       //      partial sums are added into ws (same element offsets as y) and finalised by split_finalize_kernel
    int stages;
    uint32_t stage_bytes, a_bytes, b_bytes, b_tx;   // b_bytes: smem footprint (1 KB aligned), b_tx: bytes the TMA delivers
    uint32_t tmem_cols;
    uint32_t idesc;
};

struct TileCoord {
    int n0, z0, y0, x0;
};

// all tile arithmetic is 32-bit (tile counts are far below 2^31); the per-k-block loops below contain no division at
// all: ring stage / phase and the smem addresses advance incrementally.  (v1 of these loops spent ~430 ns per k-block
// in runtime div/mod, 64-bit division calls and parameter reloads — more than the MMAs themselves for N <= 128.)
__device__ __forceinline__ uint32_t num_spatial(const IgemmParams& P) {
    const Plan& p = P.plan;
    if (P.list) {
        const int Pv = 1 << p.lgPv;
        return (uint32_t)(*P.count) * (uint32_t)((Pv >> P.lgbd) * (Pv >> P.lgbh) * (Pv >> P.lgbw));
    }
    return (uint32_t)(P.Tn * P.Tz * P.Ty * P.Tx);
}

__device__ __forceinline__ void decode_spatial(const IgemmParams& P, uint32_t t, TileCoord& c) {
    const Plan& p = P.plan;
    if (P.list) {
        // sub-tile counts inside a patch are powers of two
        const int lx = p.lgPv - P.lgbw, ly = p.lgPv - P.lgbh, lz = p.lgPv - P.lgbd;
        const uint32_t ix = t & ((1u << lx) - 1); t >>= lx;
        const uint32_t iy = t & ((1u << ly) - 1); t >>= ly;
        const uint32_t iz = t & ((1u << lz) - 1); t >>= lz;
        const uint32_t pid = (uint32_t)P.list[t];
        const uint32_t L = (uint32_t)(p.fd * p.fh * p.fw);
        const uint32_t n = pid / L, l = pid - n * L;
        const uint32_t hw = (uint32_t)(p.fh * p.fw);
        const uint32_t pz = l / hw, r2 = l - pz * hw;
        const uint32_t py = r2 / (uint32_t)p.fw, px = r2 - py * (uint32_t)p.fw;
        c.n0 = (int)n;
        c.z0 = (int)((pz << p.lgPv) + (iz << P.lgbd));
        c.y0 = (int)((py << p.lgPv) + (iy << P.lgbh));
        c.x0 = (int)((px << p.lgPv) + (ix << P.lgbw));
    } else {
        const uint32_t tx = t % (uint32_t)P.Tx; t /= (uint32_t)P.Tx;
        const uint32_t ty = t % (uint32_t)P.Ty; t /= (uint32_t)P.Ty;
        const uint32_t tz = t % (uint32_t)P.Tz; t /= (uint32_t)P.Tz;
        c.x0 = (int)(tx << P.lgbw);
        c.y0 = (int)(ty << P.lgbh);
        c.z0 = (int)(tz << P.lgbd);
        c.n0 = (int)(t << P.lgbn);
    }
}

// work item w → (group, N tile, first spatial tile); spatial fastest so that concurrently running CTAs share weights
struct SuperTile {
    int g, nt;
    uint32_t s0;
    int tap_begin, tap_end;          // the taps this work item contracts over (all of the group's unless split-K)
};
__device__ __forceinline__ void decode_super(const IgemmParams& P, uint32_t w, uint32_t n_super, SuperTile& st) {
    uint32_t q = w / n_super;
    st.s0 = (w - q * n_super) * (uint32_t)P.T;
    const uint32_t ksp = q % (uint32_t)P.ksplit;
    q /= (uint32_t)P.ksplit;
    st.g = (int)(q / (uint32_t)P.n_ntiles);
    st.nt = (int)(q - (uint32_t)st.g * (uint32_t)P.n_ntiles);
    const Group& G = P.plan.groups[st.g];
    const int per = (G.tap_count + P.ksplit - 1) / P.ksplit;
    const int b = (int)ksp * per, e = b + per;
    st.tap_begin = G.tap_begin + (b < G.tap_count ? b : G.tap_count);
    st.tap_end = G.tap_begin + (e < G.tap_count ? e : G.tap_count);
}

// transpose-reduce: every lane holds v[0..31] (one row, 32 columns); afterwards lane L holds Σ_rows column L in v[0]
__device__ __forceinline__ float warp_column_sums(float* v) {
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

template <int KC>
__global__ void __launch_bounds__(256, 1) igemm_kernel(const __grid_constant__ IgemmParams P) {
    constexpr uint32_t LAYOUT = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);     // SW128 / SW64 / SW32
    constexpr uint32_t SBO = 8u * KC * 2u;                                // 8 rows of KC bf16
    extern __shared__ uint8_t smem_raw[];
    const Plan& p = P.plan;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* ctrl = smem + (size_t)P.stages * P.stage_bytes;
    uint64_t* full_bar = (uint64_t*)ctrl;                 // [stages]
    uint64_t* empty_bar = full_bar + 8;                   // [stages]
    uint64_t* tfull_bar = empty_bar + 8;                  // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                 // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
    float* s_stats = (float*)(ctrl + 256);                // [2][Cy] when stats requested
    float* s_bias = s_stats + (P.stats ? 2 * p.Cy : 0);   // [Cy] (zeros without a bias): the epilogue reads it as float4
    float* s_scale = s_bias + p.Cy;                       // [Cy] (ones without ep_scale)
    const bool ep = P.ep_scale != nullptr || P.ep_act != 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.n_in_views; ++i) prefetch_tmap(&P.in_maps[i]);
        prefetch_tmap(&P.w_map);
    }
    if (P.stats) for (int i = threadIdx.x; i < 2 * p.Cy; i += blockDim.x) s_stats[i] = 0.f;
    for (int i = threadIdx.x; i < p.Cy; i += blockDim.x) { s_bias[i] = P.bias ? P.bias[i] : 0.f; s_scale[i] = P.ep_scale ? P.ep_scale[i] : 1.f; }
    if (warp == 2) tmem_alloc(tmem_slot, P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t n_spatial = num_spatial(P);
    const uint32_t Tn_ = (uint32_t)P.T;
    const uint32_t n_super = (n_spatial + Tn_ - 1) / Tn_;
    const uint32_t nwork = n_super * (uint32_t)(p.n_groups * P.n_ntiles * P.ksplit);
    // loop-invariant parameters, hoisted into registers
    const uint32_t stages = (uint32_t)P.stages, stage_bytes = P.stage_bytes, a_bytes = P.a_bytes, b_bytes = P.b_bytes;
    const uint32_t NT = (uint32_t)P.NT, kchunks = (uint32_t)P.kchunks, idesc = P.idesc;
    const uint32_t smem_base = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);

    if (warp == 0) {
        // =============================== TMA producer ===============================
        {
            uint32_t stage = 0, phase = 0;
            for (uint32_t w = blockIdx.x; w < nwork; w += gridDim.x) {
                SuperTile st;
                decode_super(P, w, n_super, st);
                TileCoord c[4];
                uint32_t nv = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    c[t].n0 = c[t].z0 = c[t].y0 = c[t].x0 = 0;
                    if ((uint32_t)t < Tn_ && st.s0 + t < n_spatial) { decode_spatial(P, st.s0 + t, c[t]); nv = t + 1; }
                }
                const uint32_t tx_bytes = nv * a_bytes + P.b_tx;
                const int ncol = st.nt * (int)NT;
                for (int ti = st.tap_begin; ti < st.tap_end; ++ti) {
                    const Tap tap = p.taps[ti];
                    const CUtensorMap* amap = &P.in_maps[tap.view];
                    for (uint32_t kc = 0; kc < kchunks; ++kc) {
                        const uint32_t fb = full0 + stage * 8u, a_dst = smem_base + stage * stage_bytes;
                        mbar_wait_u32(empty0 + stage * 8u, phase ^ 1u, 1);
                        if (elect_one()) {
                            mbar_expect_tx_u32(fb, tx_bytes);
#pragma unroll
                            for (int t = 0; t < 4; ++t)
                                if ((uint32_t)t < nv)
                                    tma_load_5d_u32(a_dst + (uint32_t)t * a_bytes, amap, fb, (int)(kc * KC), c[t].x0 + tap.dx,
                                                    c[t].y0 + tap.dy, c[t].z0 + tap.dz, c[t].n0);
                            tma_load_3d_u32(a_dst + Tn_ * a_bytes, &P.w_map, fb, (int)(kc * KC), ncol, tap.w);
                        }
                        __syncwarp();
                        if (++stage == stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        {
            uint32_t stage = 0, phase = 0, iter = 0;
            const uint32_t desc_hi = (uint32_t)(umma_desc(0, 16, SBO, LAYOUT) >> 32);
            const uint32_t desc_lo_const = (uint32_t)(umma_desc(0, 16, SBO, LAYOUT) & 0xFFFFFFFFu);
            for (uint32_t w = blockIdx.x; w < nwork; w += gridDim.x, ++iter) {
                SuperTile st;
                decode_super(P, w, n_super, st);
                const uint32_t s0 = st.s0;
                const uint32_t left = n_spatial - s0;
                const uint32_t nv = left < Tn_ ? left : Tn_;
                const uint32_t acc = iter & 1u;
                mbar_wait_u32(smem_u32(&tempty_bar[acc]), ((iter >> 1) & 1u) ^ 1u, 2);
                tc_fence_after();
                const uint32_t d_base = tmem_base + acc * Tn_ * NT;
                const uint32_t kblocks = (uint32_t)(st.tap_end - st.tap_begin) * kchunks;
                for (uint32_t kb = 0; kb < kblocks; ++kb) {
                    mbar_wait_u32(full0 + stage * 8u, phase, 3);
                    tc_fence_after();
                    const uint32_t a_lo = desc_lo_const | (((smem_base + stage * stage_bytes) & 0x3FFFFu) >> 4);
                    const uint32_t b_lo = a_lo + ((Tn_ * a_bytes) >> 4);
                    // the T accumulators are independent: consecutive MMAs never wait on each other's accumulate
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k) {
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                if ((uint32_t)t < nv) {
                                    const uint64_t adesc = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + (uint32_t)t * (a_bytes >> 4) + 2u * k);
                                    const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * k);
                                    mma_bf16(d_base + (uint32_t)t * NT, adesc, bdesc, idesc, (kb | (uint32_t)k) != 0u);
                                }
                            }
                        }
                        mma_commit_u32(empty0 + stage * 8u);        // smem slot free once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
                if (elect_one()) mma_commit_u32(smem_u32(&tfull_bar[acc]));      // accumulators complete
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===============================
        const int q4 = warp - 4;                         // TMEM lane quadrant == warp % 4
        const int row = q4 * 32 + lane;
        uint32_t iter = 0;
        for (uint32_t w = blockIdx.x; w < nwork; w += gridDim.x, ++iter) {
            SuperTile st;
            decode_super(P, w, n_super, st);
            const Group& G = p.groups[st.g];
            const View& ov = p.out_views[G.out_view];
            const uint32_t acc = iter & 1u;
            mbar_wait(&tfull_bar[acc], (iter >> 1) & 1u, 4);
            tc_fence_after();
            for (uint32_t t = 0; t < Tn_; ++t) {
                if (st.s0 + t >= n_spatial) break;
                TileCoord c;
                decode_spatial(P, st.s0 + t, c);
                // row → voxel of the out view (same linearisation as the TMA box: n, z, y, x with x fastest)
                const int x = c.x0 + (row & ((1 << P.lgbw) - 1));
                const int y = c.y0 + ((row >> P.lgbw) & ((1 << P.lgbh) - 1));
                const int z = c.z0 + ((row >> (P.lgbw + P.lgbh)) & ((1 << P.lgbd) - 1));
                const int n = c.n0 + (row >> (P.lgbw + P.lgbh + P.lgbd));
                const bool valid = n < p.oN && z < p.oD && y < p.oH && x < p.oW;
                bool on = valid;
                if (valid && P.active && p.lgPv >= 0)
                    on = P.active[((n * p.fd + (z >> p.lgPv)) * p.fh + (y >> p.lgPv)) * p.fw + (x >> p.lgPv)] != 0;
                bf16* yrow = P.y + ov.base + (long)n * ov.sN + (long)z * ov.sD + (long)y * ov.sH + (long)x * ov.sW +
                             (long)st.nt * P.NT;
                const uint32_t t_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (acc * Tn_ + t) * NT;
                if (P.ksplit > 1) {
                    // split-K: raw fp32 partial sums into the workspace (bias / mask / bf16 / Σ,Σ² happen in the finalise pass)
                    if (st.tap_end > st.tap_begin) {
                        float* wrow = P.ws + (yrow - P.y);
                        for (int col = 0; col < P.NT; col += 16) {
                            uint32_t r[16];
                            tmem_ld_x16(t_addr + col, r);
                            tmem_ld_wait();
                            if (valid && on) {
#pragma unroll
                                for (int j = 0; j < 16; j += 4)
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + col + j),
                                                 "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])),
                                                 "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                                                 : "memory");
                            }
                        }
                    }
                    continue;
                }
                for (int col = 0; col < P.NT; col += 32) {
                    uint32_t r[32];
                    const bool wide = (P.NT - col) >= 32;
                    if (wide) tmem_ld_x32(t_addr + col, r);
                    else tmem_ld_x16(t_addr + col, r);
                    tmem_ld_wait();
                    const int ncol = wide ? 32 : 16;
                    float v[32];
                    const float4* bq = reinterpret_cast<const float4*>(s_bias + st.nt * P.NT + col);
                    const float4* sq4 = reinterpret_cast<const float4*>(s_scale + st.nt * P.NT + col);
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
                    if (valid) {
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
                        for (int j = 0; j < 32; ++j) { if (!valid) v[j] = 0.f; sq[j] = v[j] * v[j]; }
                        float s1 = warp_column_sums(v);
                        float s2 = warp_column_sums(sq);
                        if (lane < ncol) {
                            atomicAdd(&s_stats[st.nt * P.NT + col + lane], s1);
                            atomicAdd(&s_stats[p.Cy + st.nt * P.NT + col + lane], s2);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (P.stats) {
        for (int i = threadIdx.x; i < 2 * p.Cy; i += blockDim.x) {
            float v = s_stats[i];
            if (v != 0.f) atomicAdd(&P.stats[i], (double)v);
        }
    }
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, P.tmem_cols);
    }
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}
static int pow2_ceil(int v) { return 1 << ilog2(v); }

static int pick_nt(int Cy) {
    for (int nt = 256; nt >= 16; nt -= 16)
        if (Cy % nt == 0) return nt;
    return 0;
}

// tile box (128 voxels) and accumulator interleave of the per-tap kernel for a plan — shared by the launch and by
// amb_conv_workspace_bytes()
struct IgemmGeo {
    int KC, NT, bw, bh, bd, bn, T;
    long items;                      // work items without split-K
    int max_taps;
};
static bool igemm_geometry(const Plan& p, IgemmGeo& g) {
    if (p.Cx % 16 != 0 || p.Cy % 16 != 0) return false;
    g.KC = p.Cx % 64 == 0 ? 64 : (p.Cx % 32 == 0 ? 32 : 16);
    g.NT = pick_nt(p.Cy);
    if (g.NT == 0) return false;
    g.bw = pow2_ceil(p.oW) < 8 ? pow2_ceil(p.oW) : 8;
    g.bh = pow2_ceil(p.oH) < 8 ? pow2_ceil(p.oH) : 8;
    const int rem = 128 / (g.bw * g.bh);
    g.bd = pow2_ceil(p.oD) < rem ? pow2_ceil(p.oD) : rem;
    g.bn = rem / g.bd;
    int T = 256 / g.NT;
    if (T > 4) T = 4;
    if (T < 1) T = 1;
    const uint32_t a_bytes = 128u * g.KC * 2u, b_bytes = ((uint32_t)g.NT * g.KC * 2u + 1023u) & ~1023u;
    while (T > 1 && 3u * (T * a_bytes + b_bytes) > 200u * 1024u) T >>= 1;
    g.T = T;
    const long tiles = (long)ceil_div(p.oN, g.bn) * ceil_div(p.oD, g.bd) * ceil_div(p.oH, g.bh) * ceil_div(p.oW, g.bw);
    g.items = tiles * p.n_groups * (p.Cy / g.NT);       // split-K runs with T = 1 (one tile per work item)
    g.max_taps = 0;
    for (int i = 0; i < p.n_groups; ++i) if (p.groups[i].tap_count > g.max_taps) g.max_taps = p.groups[i].tap_count;
    return true;
}
// taps split over this many work items (1 = no split): only when the tiles alone leave most SMs idle
static int igemm_ksplit(const Plan& p, const IgemmGeo& g, bool sparse_list) {
    if (getenv("AMB_NO_SPLITK") || sparse_list) return 1;
    if (g.items * 2 > (long)num_sms() || (long)g.max_taps * (p.Cx / g.KC) < 16) return 1;
    int ks = (int)((long)num_sms() / g.items);
    if (ks > g.max_taps) ks = g.max_taps;
    return ks < 2 ? 1 : ks;
}
long igemm_workspace_bytes(const Plan& p, bool sparse_list) {
    IgemmGeo g;
    if (!igemm_geometry(p, g) || igemm_ksplit(p, g, sparse_list) == 1) return 0;
    long elems = 0;                                     // the workspace mirrors the output tensor's element offsets
    for (int v = 0; v < p.n_out_views; ++v) {
        const View& ov = p.out_views[v];
        const long last = ov.base + (long)(ov.N - 1) * ov.sN + (long)(ov.D - 1) * ov.sD + (long)(ov.H - 1) * ov.sH +
                          (long)(ov.W - 1) * ov.sW + p.Cy;
        if (last > elems) elems = last;
    }
    return elems * 4;
}

// split-K finish: y = mask · act(ws · scale + bias) in bf16, Σy / Σy² per channel over the written values
__global__ void __launch_bounds__(256) split_finalize_kernel(const float* __restrict__ ws, bf16* __restrict__ y, long voxels, int Cy,
                                                             const float* __restrict__ bias, const float* __restrict__ ep_scale,
                                                             int ep_act, const uint8_t* __restrict__ active, int lgP, int D, int H,
                                                             int W, int fd, int fh, int fw, double* __restrict__ stats) {
    extern __shared__ float s_acc[];                    // [2][Cy] when stats
    const int CG = Cy / 8;
    if (stats) for (int i = threadIdx.x; i < 2 * Cy; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const long total = voxels * CG;
    for (long item = (long)blockIdx.x * blockDim.x + threadIdx.x; item < total; item += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(item % CG);
        const long vox = item / CG;
        bool on = true;
        if (active) {
            uint32_t t = (uint32_t)vox;
            const uint32_t x = t % (uint32_t)W; t /= (uint32_t)W;
            const uint32_t yy = t % (uint32_t)H; t /= (uint32_t)H;
            const uint32_t z = t % (uint32_t)D;
            const uint32_t n = t / (uint32_t)D;
            on = active[((n * fd + (z >> lgP)) * fh + (yy >> lgP)) * fw + (x >> lgP)] != 0;
        }
        float v[8];
        const float4 a0 = *reinterpret_cast<const float4*>(ws + vox * Cy + cg * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(ws + vox * Cy + cg * 8 + 4);
        const float in[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cg * 8 + j;
            v[j] = on ? ep_apply(in[j], ep_scale ? ep_scale[c] : 1.f, bias ? bias[c] : 0.f, ep_act) : 0.f;
        }
        store8(y + vox * Cy + cg * 8, v);
        if (stats && on) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&s_acc[cg * 8 + j], v[j]);
                atomicAdd(&s_acc[Cy + cg * 8 + j], v[j] * v[j]);
            }
        }
    }
    __syncthreads();
    if (stats)
        for (int i = threadIdx.x; i < 2 * Cy; i += blockDim.x)
            if (s_acc[i] != 0.f) atomicAdd(&stats[i], (double)s_acc[i]);
}

int igemm_conv(const Plan& p, const amb_conv_args* a) {
    // shapes the tensor-core kernel takes
    if (p.Cx % 16 != 0 || p.Cy % 16 != 0) {
        set_error("Cx=%d / Cy=%d not multiples of 16", p.Cx, p.Cy);
        return 0;
    }
    const int KC = p.Cx % 64 == 0 ? 64 : (p.Cx % 32 == 0 ? 32 : 16);
    const int NT = pick_nt(p.Cy);
    if (NT == 0) { set_error("no N tile for Cy=%d", p.Cy); return 0; }
    if (a->stats && p.Cy > 2048) { set_error("stats: Cy too large"); return 0; }

    static IgemmParams P;       // host staging (single-threaded by contract, mirrors `_cur_active`)
    memset(&P, 0, sizeof(P));
    P.plan = p;
    // tile box: 128 voxels
    int bw = pow2_ceil(p.oW) < 8 ? pow2_ceil(p.oW) : 8;
    int bh = pow2_ceil(p.oH) < 8 ? pow2_ceil(p.oH) : 8;
    int rem = 128 / (bw * bh);
    int bd = pow2_ceil(p.oD) < rem ? pow2_ceil(p.oD) : rem;
    int bn = rem / bd;
    const bool use_list = a->active_list != nullptr && p.lgPv >= 0 && (1 << p.lgPv) >= 8 && bn == 1 && bw == 8 &&
                          bh == 8 && bd == 2;
    P.lgbn = ilog2(bn); P.lgbd = ilog2(bd); P.lgbh = ilog2(bh); P.lgbw = ilog2(bw);
    P.Tn = ceil_div(p.oN, bn); P.Tz = ceil_div(p.oD, bd); P.Ty = ceil_div(p.oH, bh); P.Tx = ceil_div(p.oW, bw);
    P.NT = NT; P.n_ntiles = p.Cy / NT;
    P.kchunks = p.Cx / KC;
    P.a_bytes = 128u * KC * 2u;
    P.b_tx = (uint32_t)NT * KC * 2u;
    P.b_bytes = (P.b_tx + 1023u) & ~1023u;
    // T independent accumulators per CTA iteration: dependent tcgen05.mma on ONE accumulator are latency-bound
    // (~100 ns each, measured), so narrow-N layers interleave up to 4 tiles; the weight slab is shared by all of them.
    int T = 256 / NT;                 // 2 buffers x T x NT TMEM columns <= 512
    if (T > 4) T = 4;
    if (T < 1) T = 1;
    const char* tenv = getenv("AMB_IGEMM_T");
    if (tenv && atoi(tenv) >= 1 && atoi(tenv) <= T) T = atoi(tenv);
    while (T > 1 && 3u * (T * P.a_bytes + P.b_bytes) > 200u * 1024u) T >>= 1;     // keep >= 3 pipeline stages
    // split-K for the 8^3 / 16^3 stages: 16-64 tiles cannot fill 148 SMs and each of them walks all 27 (64) taps serially
    P.ksplit = 1;
    {
        IgemmGeo geo;
        const int ks = igemm_geometry(p, geo) ? igemm_ksplit(p, geo, use_list) : 1;
        if (ks > 1 && a->workspace != nullptr && a->workspace_bytes >= igemm_workspace_bytes(p, use_list)) {
            P.ksplit = ks;
            P.ws = (float*)a->workspace;
            T = 1;
        }
    }
    P.T = T;
    P.stage_bytes = (uint32_t)T * P.a_bytes + P.b_bytes;
    int stages = (int)((200u * 1024u) / P.stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) { set_error("tile does not fit shared memory"); return 0; }
    P.stages = stages;
    P.tmem_cols = (uint32_t)pow2_ceil(2 * T * NT);
    if (P.tmem_cols < 32) P.tmem_cols = 32;
    P.idesc = umma_idesc_bf16(128, NT, 0, 0);
    P.y = (bf16*)a->y; P.bias = a->bias; P.active = a->active;
    P.list = use_list ? a->active_list : nullptr;
    P.count = use_list ? a->active_count : nullptr;
    P.stats = a->stats; P.ep_scale = a->ep_scale; P.ep_act = a->ep_act;

    const int box[4] = {bn, bd, bh, bw};
    for (int i = 0; i < p.n_in_views; ++i)
        if (int e = encode_view_map(&P.in_maps[i], a->x, p.in_views[i], p.Cx, KC, box)) return e;
    int n_slabs = 0;
    for (int i = 0; i < p.n_taps; ++i) if (p.taps[i].w + 1 > n_slabs) n_slabs = p.taps[i].w + 1;
    // the packed weight tensor always holds k³ (or 64) slabs even when a plan uses a subset
    int slabs_full = (a->op == AMB_OP_CONVT || a->op == AMB_OP_CONVT_DGRAD) ? 64 : a->k * a->k * a->k;
    if (slabs_full > n_slabs) n_slabs = slabs_full;
    if (int e = encode_weight_map(&P.w_map, a->w, n_slabs, p.Cy, p.Cx, KC, NT)) return e;

    size_t smem = (size_t)P.stages * P.stage_bytes + 1024 + 256 + (a->stats ? 2 * (size_t)p.Cy * sizeof(float) : 0) +
                  2 * (size_t)p.Cy * sizeof(float);
    long tiles_upper = (((long)P.Tn * P.Tz * P.Ty * P.Tx + T - 1) / T) * p.n_groups * P.n_ntiles * P.ksplit;
    int grid = (int)(tiles_upper < (long)num_sms() ? tiles_upper : (long)num_sms());
    cudaStream_t st = (cudaStream_t)a->stream;
    if (KC == 64) {
        AMB_CUDA(cudaFuncSetAttribute(igemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        igemm_kernel<64><<<grid, 256, smem, st>>>(P);
    } else if (KC == 32) {
        AMB_CUDA(cudaFuncSetAttribute(igemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        igemm_kernel<32><<<grid, 256, smem, st>>>(P);
    } else {
        AMB_CUDA(cudaFuncSetAttribute(igemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        igemm_kernel<16><<<grid, 256, smem, st>>>(P);
    }
    AMB_LAUNCH_CHECK();
    if (P.ksplit > 1) {
        // finish: the whole output tensor (dims of the FULL-resolution produced tensor, mask grid at that resolution)
        int fD = a->D, fH = a->H, fW = a->W;
        if (a->op == AMB_OP_CONV) { fD /= a->stride; fH /= a->stride; fW /= a->stride; }
        if (a->op == AMB_OP_CONVT) { fD *= 2; fH *= 2; fW *= 2; }
        const long voxels = (long)a->N * fD * fH * fW;
        int lgP = 0;
        if (a->active) while ((fD / a->fd) >> (lgP + 1)) lgP++;
        const long items = voxels * (p.Cy / 8);
        long blocks = (items + 255) / 256, cap = (long)num_sms() * 8;
        if (blocks > cap) blocks = cap;
        split_finalize_kernel<<<(int)blocks, 256, a->stats ? 2 * p.Cy * sizeof(float) : 0, st>>>(
            P.ws, (bf16*)a->y, voxels, p.Cy, a->bias, a->ep_scale, a->ep_act, a->active, lgP, fD, fH, fW, a->fd, a->fh, a->fw,
            a->stats);
        AMB_LAUNCH_CHECK();
    }
    g_last_conv_kernel = "igemm_kernel";
    return 1;
}

}  // namespace amb
