// Microbenchmark (round 2, second probe): does staging the A operand in tensor memory lift the 64-cycle floor of a narrow
// SS-mode MMA?  profiles/r2_mma_probe.txt: an M128 x N x K16 bf16 MMA with both operands in shared memory costs
// max(N/2, ~64 + N/8) cycles — the A window (128 rows x 32 B) streams at 64 B/clk.  Here the same window is first copied
// smem -> TMEM (tcgen05.cp.128x256b, one K16 slab) and the MMA takes A from TMEM ([a_tmem] form).
//   modes: SS (baseline) | cp+TS (copy then MMA, rotating TMEM slots) | TS only (A stale in TMEM: the MMA's own floor)
//          | cp only (the copy's rate) | cp+TS x R (one copy feeds R MMAs with different B: weight-gradient / multi-N-tile reuse)
// Also a numerical check: cp+TS result == SS result on random data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_ts_probe mma_ts_probe.cu && ./mma_ts_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../anatomask_b200/csrc/ptx.cuh"
using namespace amb::ptx;

__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

struct Params { int mode, n, ne, reps, reuse, sbo_a, layout, same_b, zero, lds; long long* out; float* dump; };
#define A_COL0 448u            // TMEM staging for A: 8 slots x 8 columns

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Params P) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    __shared__ uint64_t s_ad[8], s_bd[8];
    __shared__ uint32_t s_dd[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pseudo-random bf16 payload (small integers: exact products)
    for (int i = threadIdx.x; i < 160 * 1024 / 2; i += blockDim.x)
        ((__nv_bfloat16*)smem)[i] = __float2bfloat16(P.zero ? 0.f : (float)(((i * 2654435761u) >> 27) & 7) - 3.f);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 112 * 1024;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tslot;
    const uint32_t idesc = umma_idesc_bf16(128, P.n, 0, 0);
    if (warp == 1) {
        long long t0 = 0, t1 = 0;
        // descriptors precomputed: the issuing thread's loop is 8 (or 8 x reuse) back-to-back tcgen05 instructions
        uint64_t ad[8], bd[8];
        uint32_t at[8], dd[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ad[i] = umma_desc(a_base + (uint32_t)((i % 6) * 18432 + (P.layout == 4 ? 640 + 64 : 0) + (i & 1) * 32), 16, P.sbo_a, P.layout);
            bd[i] = umma_desc(b_base + (uint32_t)(P.same_b ? 0 : (i & 1) * 32), 16, P.layout == 2 ? 1024 : 512, P.layout);
            at[i] = tb + A_COL0 + (uint32_t)i * 8u;
            dd[i] = tb + (uint32_t)((i & 1) * P.n);
        }
        const uint64_t bstep = (uint64_t)((P.n * 64) >> 4);
        if (lane == 0) for (int i = 0; i < 8; ++i) { s_ad[i] = ad[i]; s_bd[i] = bd[i]; s_dd[i] = dd[i]; }
        __syncwarp();
        for (int pass = 0; pass < 2; ++pass) {
            t0 = clock64();
            for (int r = 0; r < P.reps; ++r) {
                if (elect_one()) {
                    if (P.mode == 0 && P.lds) {          // descriptors re-read from shared memory before every MMA (the first probe's loop)
#pragma unroll 4
                        for (int i = 0; i < P.ne; ++i)
                            mma_bf16(((volatile uint32_t*)s_dd)[i], ((volatile uint64_t*)s_ad)[i], ((volatile uint64_t*)s_bd)[i], idesc, true);
                    } else if (P.mode == 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) mma_bf16(dd[i], ad[i], bd[i], idesc, true);
                    } else if (P.mode == 5) {           // descriptors depend on the loop counter: one uniform add per operand per MMA
                        const uint32_t dl = (uint32_t)(r & 1) * 2u;
#pragma unroll
                        for (int i = 0; i < 8; ++i) mma_bf16(dd[i], ad[i] + dl, bd[i] + dl, idesc, true);
                    } else if (P.mode == 6) {           // as 5, low words only (hi | lo form, like the conv kernels)
                        const uint32_t dl = (uint32_t)(r & 1) * 2u;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint64_t a2 = (ad[i] & 0xFFFFFFFF00000000ull) | (uint64_t)((uint32_t)ad[i] + dl);
                            const uint64_t b2 = (bd[i] & 0xFFFFFFFF00000000ull) | (uint64_t)((uint32_t)bd[i] + dl);
                            mma_bf16(dd[i], a2, b2, idesc, true);
                        }
                    } else if (P.mode == 1 && P.reuse == 1) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { cp_128x256b(at[i], ad[i]); mma_bf16_ts(dd[i], at[i], bd[i], idesc, true); }
                    } else if (P.mode == 1) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            cp_128x256b(at[i], ad[i]);
                            mma_bf16_ts(dd[i], at[i], bd[i], idesc, true);
                            mma_bf16_ts(dd[i], at[i], bd[i] + bstep, idesc, true);
                            mma_bf16_ts(dd[i], at[i], bd[i] + 2 * bstep, idesc, true);
                        }
                    } else if (P.mode == 2) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) mma_bf16_ts(dd[i], at[i], bd[i], idesc, true);
                    } else if (P.mode == 3) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) cp_128x256b(at[i], ad[i]);
                    } else {
                        // software-pipelined: the copy of window i+1 is issued before the MMA of window i
                        cp_128x256b(at[0], ad[0]);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (i + 1 < 8) cp_128x256b(at[i + 1], ad[i + 1]);
                            mma_bf16_ts(dd[i], at[i], bd[i], idesc, true);
                        }
                    }
                }
                __syncwarp();
            }
            if (elect_one()) mma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, pass & 1, 1);
            t1 = clock64();
        }
        if (lane == 0) P.out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // numerical check (block 0): D0 = A·B via SS into columns [0, n), via cp+TS into columns [256, 256+n); dump both
    if (P.dump && blockIdx.x == 0) {
        if (warp == 1) {
            if (elect_one()) {
                for (int k = 0; k < 2; ++k) {
                    const uint64_t ad = umma_desc(a_base + 640 + 64 + k * 32, 16, P.sbo_a, 4);
                    const uint64_t bd = umma_desc(b_base + k * 32, 16, 512, 4);
                    mma_bf16(tb, ad, bd, idesc, k > 0);
                    cp_128x256b(tb + A_COL0 + k * 8, ad);
                    mma_bf16_ts(tb + 256, tb + A_COL0 + k * 8, bd, idesc, k > 0);
                }
                mma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, 0, 2);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        uint32_t r[32];
        for (int half = 0; half < 2; ++half) {
            for (int c = 0; c < P.n; c += 32) {
                tmem_ld_x32(tb + ((uint32_t)(warp * 32) << 16) + half * 256 + c, r);
                tmem_ld_wait();
                for (int j = 0; j < 32 && c + j < P.n; ++j)
                    P.dump[((size_t)half * 128 + warp * 32 + lane) * P.n + c + j] = __uint_as_float(r[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

static double run(const char* name, int mode, int n, int grid, int reuse = 1, bool check = false, int sbo = 640, int layout = 4, int same_b = 0, int zero = 0, int lds = 0) {
    Params P;
    P.mode = mode; P.n = n; P.ne = 8; P.reps = 400; P.reuse = reuse; P.sbo_a = sbo; P.layout = layout; P.same_b = same_b; P.zero = zero; P.lds = lds; P.dump = nullptr;
    cudaMalloc(&P.out, sizeof(long long) * grid);
    if (check) cudaMalloc(&P.dump, sizeof(float) * 2 * 128 * n);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024 + 1024);
    probe<<<grid, 128, 161 * 1024 + 1024>>>(P);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), P.out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    cudaFree(P.out);
    double mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    const int mmas = mode == 3 ? 8 : 8 * reuse;
    double per_seq = mx / P.reps;
    printf("%-40s N=%3d grid %3d: %7.1f cyc per A window  %6.1f cyc/MMA  (tensor floor %5.1f)\n", name, n, grid, per_seq / 8,
           per_seq / mmas, n / 2.0);
    if (check) {
        std::vector<float> d(2 * 128 * n);
        cudaMemcpy(d.data(), P.dump, sizeof(float) * d.size(), cudaMemcpyDeviceToHost);
        cudaFree(P.dump);
        size_t bad = 0; double s = 0;
        for (size_t i = 0; i < (size_t)128 * n; ++i) { if (d[i] != d[i + (size_t)128 * n]) ++bad; s += fabs(d[i]); }
        printf("    check N=%d: cp+TS vs SS mismatches %zu of %d, sum|D| = %.1f\n", n, bad, 128 * n, s);
    }
    return per_seq;
}

int main() {
    for (int n : {64, 96, 192}) run("numerics", 1, n, 1, 1, true);
    for (int grid : std::vector<int>{}) {
        for (int n : {32, 64, 96, 128, 192, 256}) {
            run("SS (A and B from shared memory)", 0, n, grid);
            run("cp smem->TMEM + TS MMA", 1, n, grid);
            run("TS MMA only (A resident in TMEM)", 2, n, grid);
            run("cp only", 3, n, grid);
            run("cp (one ahead) + TS MMA", 4, n, grid);
            if (n <= 128) run("cp + 3 TS MMAs (one A, three B)", 1, n, grid, 3);
        }
    }
    for (int n : {32, 64, 96, 128, 192}) {
        run("SS descriptors loop-invariant", 0, n, 1);
        run("SS 64-bit add per operand per MMA", 5, n, 1);
        run("SS lo-word add per operand per MMA", 6, n, 1);
    }
    return 0;
}
