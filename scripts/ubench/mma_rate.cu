// micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16, SS mode) vs N, issue style and accumulator rotation
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../dmmfods_b200/csrc/common.cuh"
using namespace dmm;

__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ void umma_u(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc, uint32_t leader) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %7, 0;\n\t"
                 "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ void commit_u(uint64_t* bar, uint32_t leader) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}

// style 0: warp-uniform + predicated asm; style 1: single thread branch (classic)
template <int STYLE, int UNROLL, int N, int nacc>
__global__ void __launch_bounds__(128, 1) k(int iters, int sbo, int astep, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = __shfl_sync(0xffffffffu, holder, 0);
    if (warp == 1) {
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        const uint32_t a0 = (smem_u32(smem) >> 4) | (1u << 16);
        const uint32_t b0 = (smem_u32(smem + 65536) >> 4) | (1u << 16);
        const uint32_t ahi = desc_hi(sbo), bhi = desc_hi(1024);
        const uint32_t leader = lane == 0;
        long long t0 = clock64();
        if (STYLE == 0) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                                        umma_u(tb + (u % nacc) * N, a0 + (u & 3) * 2 + astep * (u & 7), ahi, b0 + (u & 3) * 2, bhi, idesc, 1u, leader);
                }
            }
            commit_u(&bar, leader);
        } else {
            if (STYLE == 2 ? elect_one() : (lane == 0)) {
                for (int i = 0; i < iters; ++i) {
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) {
                                                const uint64_t ad = ((uint64_t)ahi << 32) | (a0 + (u & 3) * 2 + astep * (u & 7));
                        const uint64_t bd = ((uint64_t)bhi << 32) | (b0 + (u & 3) * 2);
                        umma_bf16(tb + (u % nacc) * N, ad, bd, idesc, 1u);
                    }
                }
                umma_commit(&bar);
            }
            __syncwarp();
        }
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int STYLE, int N, int nacc>
void run(int grid, long long* out) {
    const int smem = 98 * 1024, iters = 256, UN = 8;
    cudaFuncSetAttribute(k<STYLE, UN, N, nacc>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cfg = 0; cfg < 2; ++cfg) {
        int sbo = cfg ? 1280 : 1024, astep = cfg ? 8 : 0;
        for (int rep = 0; rep < 2; ++rep) k<STYLE, UN, N, nacc><<<grid, 128, smem>>>(iters, sbo, astep, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%d %3d %d %4d %d %3d : %7.1f %7.1f\n", STYLE, N, nacc, sbo, astep, grid, (double)h[0] / (iters * UN), (double)h[1] / (iters * UN));
    }
}
template <int STYLE> void run_all(int grid, long long* out) {
    run<STYLE, 16, 1>(grid, out); run<STYLE, 16, 4>(grid, out);
    run<STYLE, 32, 1>(grid, out); run<STYLE, 32, 4>(grid, out);
    run<STYLE, 64, 1>(grid, out); run<STYLE, 64, 4>(grid, out);
    run<STYLE, 128, 1>(grid, out); run<STYLE, 128, 4>(grid, out);
    run<STYLE, 256, 1>(grid, out); run<STYLE, 256, 2>(grid, out);
}
int main() {
    long long* out; cudaMalloc(&out, 16);
    printf("style N nacc sbo astep grid : issue_cyc/mma  total_cyc/mma\n");
    run_all<1>(148, out); run_all<2>(148, out);
    return 0;
}
