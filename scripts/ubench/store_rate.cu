// micro-benchmark: how fast can one SM push [128 pixels x 64 channels] bf16 chunks (the igemm epilogue's unit) to global memory?
//   mode 0: staging writes (8 x st.shared.v4 per thread) + TMA tensor store of the swizzled slot (igemm2's epilogue path)
//   mode 1: TMA tensor store only (no staging writes): the pure issue / drain rate of the store engine
//   mode 2: staging writes + ld.shared.v4 + COALESCED st.global.v4 (a warp instruction covers 4 full 128-byte rows)
//   mode 3: per-thread row stores (8 x st.global.v4 per thread, row = pixel) straight from registers
//   mode 4: like mode 0 with TWO staging slots per team (the store of chunk i drains while chunk i+1 is staged)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o store_rate store_rate.cu ../../dmmfods_b200/csrc/common.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../dmmfods_b200/csrc/common.cuh"
using namespace dmm;

static int view_to_tmap(CUtensorMap* out, const dmm_view_t& v, int box_c, int box_w, int box_h, int swizzle) {
    uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.B};
    uint64_t strides[3] = {(uint64_t)v.sw, (uint64_t)v.sh, (uint64_t)v.sb};
    uint32_t box[4] = {(uint32_t)box_c, (uint32_t)box_w, (uint32_t)box_h, 1u};
    return make_tmap_bf16(out, v.ptr, 4, dims, strides, box, swizzle);
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void team_bar(int team) { asm volatile("bar.sync %0, 128;" ::"r"(team + 1) : "memory"); }

struct P {
    CUtensorMap o_map;
    __nv_bfloat16* out;
    long long ld;
    int W, H, B, TW, TH, tiles_x, tiles_y, nchunk;
    long long total_tiles;
    long long* cyc;
};

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ P p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int team = threadIdx.x >> 7, r = threadIdx.x & 127;
    uint8_t* slots = smem + team * 2 * 16384;
    const int px = r % p.TW, py = r / p.TW;
    uint32_t ctr = 0, sl = 0;
    const long long t0 = clock64();
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        long long t = tile;
        const int tx = (int)(t % p.tiles_x); t /= p.tiles_x;
        const int ty = (int)(t % p.tiles_y);
        const int b = (int)(t / p.tiles_y);
        const int x0 = tx * p.TW, y0 = ty * p.TH;
        for (int c = 0; c < p.nchunk; ++c) {
            if (((ctr++) & 1) != (uint32_t)team) continue;
            uint8_t* slot = slots + ((MODE == 4) ? (sl & 1) * 16384 : 0);
            const uint32_t srow = smem_u32(slot) + r * 128;
            uint4 v = make_uint4(r, c, (uint32_t)tile, 0x3f803f80u);
            if (MODE == 0 || MODE == 1 || MODE == 4) {
                if (r == 0) { if (MODE == 4) bulk_wait_read1(); else bulk_wait_read0(); }
                team_bar(team);
                if (MODE != 1) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) sts_v4(srow + ((g ^ (r & 7)) << 4), v);
                }
                fence_proxy_async();
                team_bar(team);
                if (r == 0) { tma_store_4d(&p.o_map, slot, c * 64, x0, y0, b); bulk_commit(); }
                ++sl;
            } else if (MODE == 2) {
                team_bar(team);
#pragma unroll
                for (int g = 0; g < 8; ++g) sts_v4(srow + ((g ^ (r & 7)) << 4), v);
                team_bar(team);
                // thread e handles 16-byte chunk (e & 7) of rows (e >> 3) + 16 i: a warp instruction writes 4 complete 128-byte rows
                const int j = r & 7;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = (r >> 3) + 16 * i;
                    const uint4 w = lds_v4(smem_u32(slot) + row * 128 + ((j ^ (row & 7)) << 4));
                    const int xx = x0 + row % p.TW, yy = y0 + row / p.TW;
                    if (xx < p.W && yy < p.H) {
                        __nv_bfloat16* dst = p.out + (((long long)b * p.H + yy) * p.W + xx) * p.ld + c * 64 + j * 8;
                        *reinterpret_cast<uint4*>(dst) = w;
                    }
                }
            } else {
                const int xx = x0 + px, yy = y0 + py;
                if (xx < p.W && yy < p.H) {
                    __nv_bfloat16* dst = p.out + (((long long)b * p.H + yy) * p.W + xx) * p.ld + c * 64;
#pragma unroll
                    for (int g = 0; g < 8; ++g) reinterpret_cast<uint4*>(dst)[g] = v;
                }
            }
        }
    }
    if ((MODE == 0 || MODE == 1 || MODE == 4) && r == 0) bulk_wait_all();
    __syncthreads();
    if (threadIdx.x == 0) p.cyc[blockIdx.x] = clock64() - t0;
}

template <int MODE>
void run(P p, const char* label) {
    const int smem = 4 * 16384 + 1024;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<148, 256, smem>>>(p);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    long long h[148]; cudaMemcpy(h, p.cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i] / 148;
    const double bytes = (double)p.B * p.H * p.W * p.nchunk * 128.0;
    const double chunks_per_sm = (double)p.total_tiles * p.nchunk / 148.0;
    printf("%-44s ld %5lld chunks/tile %2d : %7.3f ms  %6.0f GB/s  %6.0f cycles per 16 KB chunk per SM (%.1f per 128-byte row)\n", label, p.ld, p.nchunk,
           best, bytes / best / 1e6, avg / chunks_per_sm, avg / chunks_per_sm / 128.0);
}

int main() {
    const int B = 32, H = 160, W = 240, TW = 16, TH = 8;
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    for (int ld : {128, 256, 1024}) {
        __nv_bfloat16* out; cudaMalloc(&out, (size_t)B * H * W * ld * 2);
        P p; memset(&p, 0, sizeof(p));
        dmm_view_t v; v.ptr = out; v.C = ld; v.W = W; v.H = H; v.B = B; v.sw = ld; v.sh = (long long)ld * W; v.sb = (long long)ld * W * H;
        if (view_to_tmap(&p.o_map, v, 64, TW, TH, 128)) { printf("tmap failed: %s\n", dmm_last_error()); return 1; }
        p.out = out; p.ld = ld; p.W = W; p.H = H; p.B = B; p.TW = TW; p.TH = TH;
        p.tiles_x = W / TW; p.tiles_y = H / TH; p.total_tiles = (long long)p.tiles_x * p.tiles_y * B; p.cyc = cyc;
        for (int nchunk : {2, ld / 64}) {
            p.nchunk = nchunk;
            run<0>(p, "0: stage + TMA store (1 slot / team)");
            run<4>(p, "4: stage + TMA store (2 slots / team)");
            run<1>(p, "1: TMA store only");
            run<2>(p, "2: stage + coalesced st.global.v4");
            run<3>(p, "3: per-thread row st.global.v4 x 8");
            if (ld / 64 == 2) break;
        }
        cudaFree(out);
    }
    return 0;
}
