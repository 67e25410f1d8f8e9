// Weight-gradient GEMM for sm_100a:  dw[t][m][n] += sum_pixels X(pix + tap_t)[m] * Y(pix)[n].
//
// The reduction (K) dimension is the pixel index, so with pixel-major activations both operands are
// "MN-major" for tcgen05: a TMA box {64 channels, tile_w, tile_h, 1} lands in shared memory as
// 64 pixel-rows of 128 bytes (SWIZZLE_128B), which is exactly the canonical MN-major SW128 layout
// (8 k-rows x 128 B atoms, SBO = 1024 B between k-groups, LBO = distance between 64-channel chunks).
// One CTA owns (tap, 128-channel m-tile of X, n-tile of Y, a contiguous range of 64-pixel tiles),
// accumulates in TMEM over its pixel range and adds its fp32 partial result to dw with red.global.
// Warp roles as in igemm.cu: warp 0 TMA, warp 1 MMA issue + TMEM owner, warps 2..5 epilogue.
#include "common.cuh"
#include "../../include/dmmfods_b200.h"

namespace dmm {

struct WgradKParams {
    CUtensorMap x_map;
    CUtensorMap y_maps[DMM_MAX_SRC];
    int8_t tap_ysrc[DMM_MAX_TAPS];
    int8_t tap_dy[DMM_MAX_TAPS];
    int8_t tap_dx[DMM_MAX_TAPS];
    int tile_w, tile_h, tiles_x, tiles_y;
    long long total_tiles;
    int splits;
    int M, N;
    int m_tiles, n_tile, n_chunks;   // n_chunks = ceil(n_tile / 64) 64-channel TMA boxes of Y per stage
    int stages;
    uint32_t tmem_cols;
    float* dw;
    long long ldw;
};

constexpr int kWgThreads = 192;
constexpr uint32_t kChunkBytes = 64 * 128;   // 64 pixels x 64 channels bf16

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t stage_bytes = (2 + p.n_chunks) * kChunkBytes;
    uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tmem_full_bar = empty_bar + 8;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tap = blockIdx.z;
    const int mt = blockIdx.y % p.m_tiles;
    const int nt = blockIdx.y / p.m_tiles;
    const int m0 = mt * 128, n0 = nt * p.n_tile;
    const long long tile_lo = p.total_tiles * blockIdx.x / p.splits;
    const long long tile_hi = p.total_tiles * (blockIdx.x + 1) / p.splits;
    const int num_k = (int)(tile_hi - tile_lo);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, p.tmem_cols);
        tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.x_map);
        tma_prefetch_desc(&p.y_maps[p.tap_ysrc[tap]]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (num_k > 0) {
        if (warp == 0) {
            if (lane == 0) {
                const CUtensorMap* ymap = &p.y_maps[p.tap_ysrc[tap]];
                const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
                for (int kb = 0; kb < num_k; ++kb) {
                    long long t = tile_lo + kb;
                    const int tx = (int)(t % p.tiles_x);
                    t /= p.tiles_x;
                    const int ty = (int)(t % p.tiles_y);
                    const int b = (int)(t / p.tiles_y);
                    const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
                    const int s = kb % p.stages;
                    const uint32_t ph = (kb / p.stages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                    tma_load_4d(sa, &p.x_map, &full_bar[s], m0, x0 + dx, y0 + dy, b);
                    tma_load_4d(sa + kChunkBytes, &p.x_map, &full_bar[s], m0 + 64, x0 + dx, y0 + dy, b);
                    for (int c = 0; c < p.n_chunks; ++c)
                        tma_load_4d(sa + (2 + c) * kChunkBytes, ymap, &full_bar[s], n0 + c * 64, x0, y0, b);
                }
            }
        } else if (warp == 1) {
            const uint32_t idesc = make_idesc_bf16(128, p.n_tile, 1, 1);   // both operands MN-major
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (kb / p.stages) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t sb = sa + 2 * kChunkBytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {   // 64 pixels = 4 x UMMA_K(16)
                        const uint64_t ad = make_smem_desc(sa + k * 2048, kChunkBytes, 1024, 2);
                        const uint64_t bd = make_smem_desc(sb + k * 2048, kChunkBytes, 1024, 2);
                        umma_bf16(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                    if (kb == num_k - 1) umma_commit(tmem_full_bar);
                }
                __syncwarp();
            }
        } else {
            const int q = warp & 3;
            const int m = m0 + q * 32 + lane;
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
            float* drow = p.dw + ((long long)tap * p.M + m) * p.ldw + n0;
            for (int ch = 0; ch < p.n_tile / 16; ++ch) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + ch * 16, r);
                tmem_ld_wait();
                if (m < p.M) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + ch * 16 + j < p.N) atomicAdd(drow + ch * 16 + j, __uint_as_float(r[j]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

int view_to_tmap(CUtensorMap* out, const dmm_view_t& v, int box_c, int box_w, int box_h, int swizzle);

}  // namespace dmm

using namespace dmm;

extern "C" int dmm_conv_wgrad(const dmm_wgrad_t* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DMM_CHECK(d != nullptr && d->dw != nullptr, "dmm_conv_wgrad: null descriptor / output");
    DMM_CHECK(d->num_taps >= 1 && d->num_taps <= DMM_MAX_TAPS, "dmm_conv_wgrad: bad num_taps %d", d->num_taps);
    DMM_CHECK(d->num_ysrc >= 1 && d->num_ysrc <= DMM_MAX_SRC, "dmm_conv_wgrad: bad num_ysrc %d", d->num_ysrc);
    DMM_CHECK(d->tile_w == 64 || d->tile_w == 32 || d->tile_w == 16 || d->tile_w == 8, "dmm_conv_wgrad: bad tile_w %d", d->tile_w);
    DMM_CHECK(d->n_tile >= 16 && d->n_tile <= 256 && d->n_tile % 16 == 0, "dmm_conv_wgrad: bad n_tile %d", d->n_tile);
    DMM_CHECK(d->M >= 1 && d->N >= 1 && d->x.ptr != nullptr, "dmm_conv_wgrad: bad operands");
    DMM_CHECK(d->x.C == d->M, "dmm_conv_wgrad: x view has %d channels, M = %d", d->x.C, d->M);
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;

    WgradKParams p;
    memset(&p, 0, sizeof(p));
    p.tile_w = d->tile_w;
    p.tile_h = 64 / d->tile_w;
    int rc = view_to_tmap(&p.x_map, d->x, 64, p.tile_w, p.tile_h, 128);
    if (rc) return rc;
    for (int s = 0; s < d->num_ysrc; ++s) {
        DMM_CHECK(d->y[s].ptr != nullptr && d->y[s].C == d->N, "dmm_conv_wgrad: y view %d has %d channels, N = %d", s, d->y[s].C,
                  d->N);
        rc = view_to_tmap(&p.y_maps[s], d->y[s], 64, p.tile_w, p.tile_h, 128);
        if (rc) return rc;
    }
    for (int t = 0; t < d->num_taps; ++t) {
        DMM_CHECK(d->tap_ysrc[t] >= 0 && d->tap_ysrc[t] < d->num_ysrc, "dmm_conv_wgrad: tap %d bad y source", t);
        p.tap_ysrc[t] = d->tap_ysrc[t];
        p.tap_dy[t] = d->tap_dy[t];
        p.tap_dx[t] = d->tap_dx[t];
    }
    p.tiles_x = ceil_div(d->W, p.tile_w);
    p.tiles_y = ceil_div(d->H, p.tile_h);
    p.total_tiles = (long long)p.tiles_x * p.tiles_y * d->B;
    p.M = d->M;
    p.N = d->N;
    p.m_tiles = ceil_div(d->M, 128);
    p.n_tile = d->n_tile;
    p.n_chunks = ceil_div(d->n_tile, 64);
    const int n_tiles = ceil_div(d->N, d->n_tile);
    const uint32_t stage_bytes = (2 + p.n_chunks) * kChunkBytes;
    int stages = (int)((200u * 1024u) / stage_bytes);
    if (stages > 8) stages = 8;
    p.stages = stages;
    uint32_t cols = 32;
    while ((int)cols < d->n_tile) cols <<= 1;
    p.tmem_cols = cols;
    p.dw = d->dw;
    p.ldw = d->ldw;
    long long splits = d->splits;
    if (splits <= 0) {
        const long long base = (long long)d->num_taps * p.m_tiles * n_tiles;
        splits = (2 * 148 + base - 1) / base;
    }
    if (splits > p.total_tiles) splits = p.total_tiles;
    if (splits < 1) splits = 1;
    p.splits = (int)splits;

    const size_t smem = (size_t)stages * stage_bytes + 256 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        DMM_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set = true;
    }
    DMM_CHECK(smem <= 232448, "dmm_conv_wgrad: %zu bytes of shared memory requested", smem);
    dim3 grid((unsigned)p.splits, (unsigned)(p.m_tiles * n_tiles), (unsigned)d->num_taps);
    wgrad_kernel<<<grid, kWgThreads, smem, stream>>>(p);
    DMM_LAUNCH_CHECK("wgrad_kernel");
    return 0;
}
