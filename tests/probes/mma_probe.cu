// Microbenchmark (round 2): how fast does ONE elected thread retire tcgen05.mma (M=128, K=16, bf16) as a function of N and of
// the dependence pattern between consecutive MMAs?  Decides the tile order of conv_igemm4.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../anatomask_b200/csrc/ptx.cuh"
using namespace amb::ptx;

struct Entry { uint32_t d_col, n, a_off, b_off; };
#define MAX_E 64
struct Params { Entry e[MAX_E]; int ne, reps, sbo_a; long long* out; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Params P) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    __shared__ uint64_t s_ad[MAX_E], s_bd[MAX_E];
    __shared__ uint32_t s_id[MAX_E], s_dc[MAX_E];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 112 * 1024;
    if (threadIdx.x < P.ne) {
        const Entry e = P.e[threadIdx.x];
        s_ad[threadIdx.x] = umma_desc(a_base + e.a_off, 16, P.sbo_a, 4);
        s_bd[threadIdx.x] = umma_desc(b_base + e.b_off, 16, 512, 4);
        s_id[threadIdx.x] = umma_idesc_bf16(128, e.n, 0, 0);
        s_dc[threadIdx.x] = e.d_col;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tslot;
    if (warp == 1) {
        long long t0 = 0, t1 = 0;
        for (int pass = 0; pass < 2; ++pass) {          // pass 0 warms up
            t0 = clock64();
            for (int r = 0; r < P.reps; ++r) {
                if (elect_one()) {
#pragma unroll 4
                    for (int i = 0; i < P.ne; ++i) mma_bf16(tb + s_dc[i], s_ad[i], s_bd[i], s_id[i], true);
                }
                __syncwarp();
            }
            if (elect_one()) mma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, pass & 1, 1);
            t1 = clock64();
        }
        if ((threadIdx.x & 31) == 0) P.out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

static double run(const char* name, const std::vector<Entry>& es, int grid, int sbo_a = 1024) {
    Params P;
    P.ne = (int)es.size(); P.reps = 400; P.sbo_a = sbo_a;
    for (int i = 0; i < P.ne; ++i) P.e[i] = es[i];
    cudaMalloc(&P.out, sizeof(long long) * grid);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024 + 1024);
    probe<<<grid, 128, 161 * 1024 + 1024>>>(P);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), P.out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    cudaFree(P.out);
    double mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    double floor_cyc = 0;
    for (auto& en : es) floor_cyc += en.n / 2.0;
    double per_seq = mx / P.reps;
    printf("%-44s grid %3d: %8.1f cyc/seq  %6.1f cyc/MMA  floor %6.1f  eff %.3f\n", name, grid, per_seq, per_seq / P.ne,
           floor_cyc, floor_cyc / per_seq);
    return per_seq;
}

int main() {
    for (int grid : {1, 148}) {
        for (int n : {16, 32, 64, 96, 128, 192, 256}) {
            for (int nacc : {1, 2, 4}) {
                if (nacc * n > 512) continue;
                std::vector<Entry> es;
                for (int i = 0; i < 8; ++i) es.push_back({(uint32_t)((i % nacc) * n), (uint32_t)n, (uint32_t)((i % 6) * 12288 + (i % 2) * 32), 0});
                char nm[96];
                snprintf(nm, sizeof nm, "N=%d accumulators=%d (A from 6 planes)", n, nacc);
                run(nm, es, grid);
            }
        }
        // N-stacked halo-plane pattern, T tiles of NT columns, planes 0..T+1, two k steps
        for (int NT : {64, 32}) for (int T : {4, 6}) {
            if (T * NT > 512) continue;
            for (int order = 0; order < 3; ++order) {
                std::vector<Entry> es;
                std::vector<int> planes;
                for (int p = 0; p < T + 2; ++p) planes.push_back(p);
                if (order == 2) { planes.clear(); for (int p = 0; p < (T + 2 + 1) / 2; ++p) { planes.push_back(p); if (p + (T + 3) / 2 < T + 2) planes.push_back(p + (T + 3) / 2); } }
                auto add = [&](int p, int k) {
                    int lo = p - 2 < 0 ? 0 : p - 2, hi = p > T - 1 ? T - 1 : p;
                    int j0 = 2 - (p - lo);
                    es.push_back({(uint32_t)(lo * NT), (uint32_t)((hi - lo + 1) * NT), (uint32_t)(p * 12288 + 640 + 64 + k * 32),
                                  (uint32_t)(j0 * NT * 64 + k * 32)});
                };
                if (order == 1) { for (int k = 0; k < 2; ++k) for (int p : planes) add(p, k); }
                else { for (int p : planes) for (int k = 0; k < 2; ++k) add(p, k); }
                char nm[96];
                snprintf(nm, sizeof nm, "stacked NT=%d T=%d order=%s", NT, T, order == 0 ? "plane,k" : order == 1 ? "k,plane" : "interleaved-planes,k");
                run(nm, es, grid, 640);
            }
            // today's igemm3 pattern: per tap, k, tile: N = NT, tile t reads plane t+dz
            std::vector<Entry> es;
            for (int dz = 0; dz < 3; ++dz) for (int k = 0; k < 2; ++k) for (int t = 0; t < T; ++t)
                es.push_back({(uint32_t)(t * NT), (uint32_t)NT, (uint32_t)((t + dz) * 12288 + 640 + 64 + k * 32), (uint32_t)(dz * NT * 64 + k * 32)});
            char nm[96];
            snprintf(nm, sizeof nm, "per-tile (igemm3) NT=%d T=%d", NT, T);
            run(nm, es, grid, 640);
        }
    }
    return 0;
}
