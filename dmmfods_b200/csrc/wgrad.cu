// Weight-gradient GEMM for sm_100a:  D[row][col] += sum_pixels A(pix + shift_a)[row] * B(pix + shift_b)[col].
//
// The reduction (K) dimension is the pixel index, so with pixel-major activations both operands are
// "MN-major" for tcgen05: a TMA box {channels, tile_w, tile_h, 1} lands in shared memory as kpx pixel-rows of
// 2*channels bytes - exactly the canonical MN-major swizzled layout (8 k-rows per atom, SBO = 8 rows,
// LBO = distance between channel chunks).
//
// Both operands stream from HBM/L2 with no reuse along K, so the arithmetic intensity of a CTA is
// M*N/(M+N) flop/B: one CTA therefore owns up to FOUR 128-row accumulators x several column groups (all of
// TMEM, 512 columns) at once.  The A side is a list of 64-channel chunks (m-tile = 2 chunks), the B side a list of
// n_tile-channel groups; every (m-tile, group) pair has its own TMEM accumulator.  Each chunk / group carries its
// own source view and pixel shift, so that
//   * a KxK convolution loads the activation tile ONCE and pairs it with KxK shifted copies of the (narrow)
//     output-gradient tile (dense-layer conv2: 9 groups of 32 channels; refine1: 25 groups of 16), or
//   * the roles are swapped (A = shifted output-gradient chunks, B = activation) when the activation is narrow.
// A CTA handles a contiguous range of pixel tiles (split-K over the grid) and adds its fp32 partial result to the
// scratch matrix dw[row*ld + col] with red.global.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue.
#include "common.cuh"
#include "../../include/dmmfods_b200.h"

namespace dmm {

struct WgSlot {
    int8_t map;
    int8_t dy, dx;
    int8_t pad;
    int ch0;     // first channel in the source view
    int out0;    // first dw row (A chunk) / column (B group)
};

struct WgradKParams {
    CUtensorMap a_maps[DMM_MAX_SRC];
    CUtensorMap b_maps[DMM_MAX_SRC];
    int a_C[DMM_MAX_SRC], b_C[DMM_MAX_SRC];
    WgSlot a[DMM_WG_MAX_A];
    WgSlot b[DMM_WG_MAX_B];
    int num_a, num_b, na;          // na = m-tiles = ceil(num_a / 2)
    int n_tile, bw, b_chunks;      // group width, TMA box width (channels), boxes per group
    uint32_t b_layout, b_sbo;      // UMMA layout type / stride byte offset of the B tiles
    int ya, yb, a_step, b_step;
    int kpx, tile_w, tile_h, tiles_x, tiles_y;
    long long total_tiles;
    int splits;
    int stages;
    uint32_t a_chunk_bytes, b_box_bytes, stage_bytes, tmem_cols;
    float* dw;
    long long ld;
};

constexpr int kWgThreads = 192;

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tail = smem + (size_t)p.stages * p.stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tmem_full_bar = empty_bar + 8;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ia = blockIdx.y % p.ya, ib = blockIdx.y / p.ya;
    const int a_off = ia * p.a_step, b_off = ib * p.b_step;
    const long long tile_lo = p.total_tiles * blockIdx.x / p.splits;
    const long long tile_hi = p.total_tiles * (blockIdx.x + 1) / p.splits;
    const int num_k = (int)(tile_hi - tile_lo);
    const uint32_t b_group_bytes = p.b_box_bytes * p.b_chunks;
    const uint32_t a_bytes = p.a_chunk_bytes * 2 * p.na;

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
        tma_prefetch_desc(&p.a_maps[p.a[0].map]);
        tma_prefetch_desc(&p.b_maps[p.b[0].map]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (num_k > 0) {
        if (warp == 0) {
            if (lane == 0) {
                // bytes that actually arrive per stage: chunks that lie completely outside their view are skipped
                uint32_t tx_bytes = 0;
                for (int i = 0; i < p.num_a; ++i)
                    if (p.a[i].ch0 + a_off < p.a_C[p.a[i].map]) tx_bytes += p.a_chunk_bytes;
                for (int g = 0; g < p.num_b; ++g)
                    for (int c = 0; c < p.b_chunks; ++c)
                        if (p.b[g].ch0 + b_off + c * p.bw < p.b_C[p.b[g].map]) tx_bytes += p.b_box_bytes;
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
                    uint8_t* sa = smem + (size_t)s * p.stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                    for (int i = 0; i < p.num_a; ++i) {
                        const WgSlot& c = p.a[i];
                        if (c.ch0 + a_off < p.a_C[c.map])
                            tma_load_4d(sa + i * p.a_chunk_bytes, &p.a_maps[c.map], &full_bar[s], c.ch0 + a_off, x0 + c.dx,
                                        y0 + c.dy, b);
                    }
                    for (int g = 0; g < p.num_b; ++g) {
                        const WgSlot& c = p.b[g];
                        for (int k = 0; k < p.b_chunks; ++k)
                            if (c.ch0 + b_off + k * p.bw < p.b_C[c.map])
                                tma_load_4d(sb + g * b_group_bytes + k * p.b_box_bytes, &p.b_maps[c.map], &full_bar[s],
                                            c.ch0 + b_off + k * p.bw, x0 + c.dx, y0 + c.dy, b);
                    }
                }
            }
        } else if (warp == 1) {
            const uint32_t idesc = make_idesc_bf16(128, p.n_tile, 1, 1);   // both operands MN-major
            const int ksteps = p.kpx / 16;
            const uint32_t b_kstep = 16u * (uint32_t)p.bw * 2u;             // 16 pixel rows of the B tile
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (kb / p.stages) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + (size_t)s * p.stage_bytes);
                    const uint32_t sb = sa + a_bytes;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        for (int mi = 0; mi < p.na; ++mi) {
                            const uint64_t ad = make_smem_desc(sa + mi * 2 * p.a_chunk_bytes + k * 2048, p.a_chunk_bytes, 1024, 2);
                            for (int g = 0; g < p.num_b; ++g) {
                                const uint64_t bd = make_smem_desc(sb + g * b_group_bytes + k * b_kstep, p.b_box_bytes, p.b_sbo,
                                                                   p.b_layout);
                                umma_bf16(tmem_base + (uint32_t)((mi * p.num_b + g) * p.n_tile), ad, bd, idesc, acc);
                            }
                        }
                    }
                    umma_commit(&empty_bar[s]);
                    if (kb == num_k - 1) umma_commit(tmem_full_bar);
                }
                __syncwarp();
            }
        } else {
            const int q = warp & 3;
            const int row = q * 32 + lane;            // accumulator row inside the m-tile
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
            for (int mi = 0; mi < p.na; ++mi) {
                const int ci = mi * 2 + (row >> 6);   // A chunk of this row
                const int r = row & 63;
                bool rvalid = false;
                long long orow = 0;
                if (ci < p.num_a) {
                    const WgSlot& c = p.a[ci];
                    rvalid = (c.ch0 + a_off + r) < p.a_C[c.map];
                    orow = (long long)(c.out0 + a_off + r) * p.ld;
                }
                for (int g = 0; g < p.num_b; ++g) {
                    const WgSlot& c = p.b[g];
                    const int cvalid = p.b_C[c.map] - (c.ch0 + b_off);      // valid columns of this group
                    float* dst = p.dw + orow + c.out0 + b_off;
                    for (int ch = 0; ch < p.n_tile / 16; ++ch) {
                        uint32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((mi * p.num_b + g) * p.n_tile + ch * 16), v);
                        tmem_ld_wait();
                        if (rvalid) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (ch * 16 + j < cvalid) atomicAdd(dst + ch * 16 + j, __uint_as_float(v[j]));
                        }
                    }
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
    DMM_CHECK(d->num_a >= 1 && d->num_a <= DMM_WG_MAX_A, "dmm_conv_wgrad: bad num_a %d", d->num_a);
    DMM_CHECK(d->num_b >= 1 && d->num_b <= DMM_WG_MAX_B, "dmm_conv_wgrad: bad num_b %d", d->num_b);
    DMM_CHECK(d->num_a_src >= 1 && d->num_a_src <= DMM_MAX_SRC && d->num_b_src >= 1 && d->num_b_src <= DMM_MAX_SRC,
              "dmm_conv_wgrad: bad source counts");
    DMM_CHECK(d->kpx == 64 || d->kpx == 32, "dmm_conv_wgrad: kpx must be 32 or 64 (got %d)", d->kpx);
    DMM_CHECK(d->tile_w >= 8 && d->tile_w <= d->kpx && (d->tile_w & (d->tile_w - 1)) == 0, "dmm_conv_wgrad: bad tile_w %d", d->tile_w);
    DMM_CHECK(d->n_tile >= 16 && d->n_tile <= 256 && d->n_tile % 16 == 0, "dmm_conv_wgrad: bad n_tile %d", d->n_tile);
    DMM_CHECK(d->n_tile == 16 || d->n_tile == 32 || d->n_tile >= 64, "dmm_conv_wgrad: n_tile %d (16, 32 or >= 64)", d->n_tile);
    const int na = (d->num_a + 1) / 2;
    DMM_CHECK(na * d->num_b * d->n_tile <= 512, "dmm_conv_wgrad: %d m-tiles x %d groups x %d columns exceed the 512 TMEM columns",
              na, d->num_b, d->n_tile);
    DMM_CHECK(d->ya >= 1 && d->yb >= 1, "dmm_conv_wgrad: bad grid replication");
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;

    WgradKParams p;
    memset(&p, 0, sizeof(p));
    p.kpx = d->kpx;
    p.tile_w = d->tile_w;
    p.tile_h = d->kpx / d->tile_w;
    p.n_tile = d->n_tile;
    p.bw = d->n_tile >= 64 ? 64 : d->n_tile;
    p.b_chunks = (d->n_tile + p.bw - 1) / p.bw;
    const int b_swz = p.bw * 2;                                  // 128 / 64 / 32-byte swizzle = row bytes
    p.b_layout = b_swz == 128 ? 2u : (b_swz == 64 ? 4u : 6u);
    p.b_sbo = 8u * (uint32_t)b_swz;
    for (int s = 0; s < d->num_a_src; ++s) {
        DMM_CHECK(d->a_src[s].ptr != nullptr, "dmm_conv_wgrad: A source %d is null", s);
        int rc = view_to_tmap(&p.a_maps[s], d->a_src[s], 64, p.tile_w, p.tile_h, 128);
        if (rc) return rc;
        p.a_C[s] = d->a_src[s].C;
    }
    for (int s = 0; s < d->num_b_src; ++s) {
        DMM_CHECK(d->b_src[s].ptr != nullptr, "dmm_conv_wgrad: B source %d is null", s);
        int rc = view_to_tmap(&p.b_maps[s], d->b_src[s], p.bw, p.tile_w, p.tile_h, b_swz);
        if (rc) return rc;
        p.b_C[s] = d->b_src[s].C;
    }
    for (int i = 0; i < d->num_a; ++i) {
        DMM_CHECK(d->a[i].src >= 0 && d->a[i].src < d->num_a_src, "dmm_conv_wgrad: A chunk %d bad source", i);
        p.a[i].map = d->a[i].src; p.a[i].dy = d->a[i].dy; p.a[i].dx = d->a[i].dx;
        p.a[i].ch0 = d->a[i].ch0; p.a[i].out0 = d->a[i].out0;
    }
    for (int i = 0; i < d->num_b; ++i) {
        DMM_CHECK(d->b[i].src >= 0 && d->b[i].src < d->num_b_src, "dmm_conv_wgrad: B group %d bad source", i);
        p.b[i].map = d->b[i].src; p.b[i].dy = d->b[i].dy; p.b[i].dx = d->b[i].dx;
        p.b[i].ch0 = d->b[i].ch0; p.b[i].out0 = d->b[i].out0;
    }
    p.num_a = d->num_a; p.num_b = d->num_b; p.na = na;
    p.ya = d->ya; p.yb = d->yb; p.a_step = d->a_step; p.b_step = d->b_step;
    p.tiles_x = ceil_div(d->W, p.tile_w);
    p.tiles_y = ceil_div(d->H, p.tile_h);
    p.total_tiles = (long long)p.tiles_x * p.tiles_y * d->B;
    p.a_chunk_bytes = (uint32_t)d->kpx * 128u;
    p.b_box_bytes = (uint32_t)d->kpx * (uint32_t)p.bw * 2u;
    p.stage_bytes = p.a_chunk_bytes * 2 * na + p.b_box_bytes * p.b_chunks * d->num_b;
    p.stage_bytes = (p.stage_bytes + 1023u) & ~1023u;
    int stages = (int)((200u * 1024u) / p.stage_bytes);
    DMM_CHECK(stages >= 2, "dmm_conv_wgrad: stage of %u bytes does not fit twice in shared memory", p.stage_bytes);
    if (stages > 8) stages = 8;
    p.stages = stages;
    uint32_t cols = 32;
    while ((int)cols < na * d->num_b * d->n_tile) cols <<= 1;
    p.tmem_cols = cols;
    p.dw = d->dw;
    p.ld = d->ld;
    long long splits = d->splits;
    const long long items = (long long)d->ya * d->yb;
    if (splits <= 0) {
        splits = (2 * 148 + items - 1) / items;
        const long long min_k = 8;                    // amortise prologue + atomics epilogue
        if (splits > p.total_tiles / min_k) splits = p.total_tiles / min_k;
    }
    if (splits > p.total_tiles) splits = p.total_tiles;
    if (splits < 1) splits = 1;
    p.splits = (int)splits;

    const size_t smem = (size_t)stages * p.stage_bytes + 256 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        DMM_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set = true;
    }
    DMM_CHECK(smem <= 232448, "dmm_conv_wgrad: %zu bytes of shared memory requested", smem);
    dim3 grid((unsigned)p.splits, (unsigned)items, 1);
    wgrad_kernel<<<grid, kWgThreads, smem, stream>>>(p);
    DMM_LAUNCH_CHECK("wgrad_kernel");
    return 0;
}
