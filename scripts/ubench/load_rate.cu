// micro-benchmark: how fast does one SM take in a [256 rows x 128 bytes] weight stage (32 KB) that sits in L2?
//   mode 0: 2-D tensor load (box {64 bf16, 256 rows}, SWIZZLE_128B) from a row-major [rows][K] matrix (row pitch K*2 bytes): what
//           igemm2's weight producer does today
//   mode 1: 1-D bulk copy (cp.async.bulk) of ONE contiguous 32 KB block (weights pre-tiled / pre-swizzled in global memory)
//   mode 2: 2-D tensor load from a matrix whose row pitch IS 128 bytes (contiguous rows)
// 4-stage ring, one producer thread, consumer = mbarrier wait only.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o load_rate load_rate.cu ../../dmmfods_b200/csrc/common.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../dmmfods_b200/csrc/common.cuh"
using namespace dmm;

struct P {
    CUtensorMap map;
    const uint8_t* base;
    int mode, iters, nstage, kblocks, ntiles;
    long long* cyc;
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ P p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 6 * 32768);
    uint64_t* empty = full + 8;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < p.iters; ++i) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], 32768);
            const int kb = (i + blockIdx.x * 7) % p.kblocks, nt = (i / p.kblocks + blockIdx.x) % p.ntiles;
            if (p.mode == 1) bulk_load_1d(smem + s * 32768, p.base + ((size_t)nt * p.kblocks + kb) * 32768, 32768, &full[s]);
            else if (p.mode == 0) tma_load_2d(smem + s * 32768, &p.map, &full[s], kb * 64, nt * 256);
            else tma_load_2d(smem + s * 32768, &p.map, &full[s], 0, (nt * p.kblocks + kb) * 256);
            if (++s == p.nstage) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < p.iters; ++i) {
            mbar_wait(&full[s], ph);
            mbar_arrive(&empty[s]);
            if (++s == p.nstage) { s = 0; ph ^= 1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) p.cyc[blockIdx.x] = clock64() - t0;
}

int main() {
    const int kblocks = 16, ntiles = 4;                  // a [1024 rows][1024 K] bf16 weight matrix = 2 MB (L2 resident)
    const size_t bytes = (size_t)kblocks * ntiles * 32768;
    uint8_t* w; cudaMalloc(&w, bytes); cudaMemset(w, 1, bytes);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    const int smem = 6 * 32768 + 1024 + 256;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int nstage : {2, 4, 6}) {
        for (int mode = 0; mode < 3; ++mode) {
            P p; memset(&p, 0, sizeof(p));
            p.base = w; p.mode = mode; p.iters = 2000; p.nstage = nstage; p.kblocks = kblocks; p.ntiles = ntiles; p.cyc = cyc;
            if (mode == 0) {
                uint64_t dims[2] = {(uint64_t)kblocks * 64, (uint64_t)ntiles * 256}, strides[1] = {(uint64_t)kblocks * 64};
                uint32_t box[2] = {64, 256};
                if (make_tmap_bf16(&p.map, w, 2, dims, strides, box, 128)) { printf("tmap: %s\n", dmm_last_error()); return 1; }
            } else if (mode == 2) {
                uint64_t dims[2] = {64, (uint64_t)ntiles * kblocks * 256}, strides[1] = {64};
                uint32_t box[2] = {64, 256};
                if (make_tmap_bf16(&p.map, w, 2, dims, strides, box, 128)) { printf("tmap: %s\n", dmm_last_error()); return 1; }
            }
            float best = 1e9f;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0); k<<<148, 64, smem>>>(p); cudaEventRecord(e1);
                cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i] / 148;
            const char* names[3] = {"2-D tensor load, row pitch 2 KB", "1-D bulk copy of a contiguous 32 KB tile", "2-D tensor load, row pitch 128 B"};
            printf("stages %d  %-44s: %7.3f ms  %6.0f cycles per 32 KB stage per SM = %5.1f B/cycle/SM, %5.0f GB/s aggregate\n", nstage, names[mode], best,
                   avg / p.iters, 32768.0 / (avg / p.iters), 148.0 * p.iters * 32768.0 / best / 1e6);
        }
    }
    return 0;
}
