// Implicit-GEMM convolution for sm_100a: TMA (4-D tiled, zero-fill halo) -> 128B/32B-swizzled smem
// -> tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
//
// One CTA = one 128-pixel output tile (tile_h x tile_w patch of one image) x one n-tile.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
// The K loop walks "taps" (source view, dy, dx) x channel blocks of `kwidth`; 1x1 convs, 3x3 / 5x5
// convs, concatenated inputs, the sub-pixel phases of ConvTranspose2d and all data-gradients are
// the same kernel with different tap tables (see include/dmmfods_b200.h: dmm_conv_igemm).
#include "common.cuh"
#include "../../include/dmmfods_b200.h"

#include <stdlib.h>

namespace dmm {

struct IgemmKParams {
    CUtensorMap a_maps[DMM_MAX_SRC];
    CUtensorMap b_map;
    int num_taps;
    int8_t tap_src[DMM_MAX_TAPS];
    int8_t tap_dy[DMM_MAX_TAPS];
    int8_t tap_dx[DMM_MAX_TAPS];
    int src_nblk[DMM_MAX_SRC];   // k-blocks per source
    int src_lastk[DMM_MAX_SRC];  // UMMA k-steps (of 16) in the last block of a source
    int W, H, B;
    int tile_w, tile_h, tiles_x, tiles_y;
    int n_tile;   // UMMA N of this launch (multiple of 16)
    int N;        // valid output channels overall
    int kwidth;   // 64 | 16 (bf16), 32 (fp32 storage)
    int split3;   // TF32: 3xTF32 error-compensated products (A = Ahi + Alo, B = Bhi + Blo; hi*hi + hi*lo + lo*hi)
    int lo_row0;  // split3: first row of the Blo matrix inside the packed weight tensor ([2][n_rows][ktot])
    int stages;
    uint32_t a_bytes, b_bytes;   // per stage
    uint32_t tmem_cols;
    void* out;
    long long ldo;
    int coff;
    int out_sy, out_sx, out_py, out_px, OH, OW;
    double* stats;
    int stats_ld, stats_off;
};

constexpr int kThreads = 192;
constexpr int kMaxSmem = 232448;   // 227 KB opt-in limit per CTA on sm_100

// TF32: fp32 storage (activations, packed weights, output), tcgen05.mma kind::tf32 (K = 8 per instruction, fp32 accumulate) -
// the strict-tolerance arithmetic mode (dmm_igemm_t.dtype == 1); kwidth = 32 fp32 channels = one 128-byte swizzle row.
template <int OUT_MODE, bool TF32>
__global__ void __launch_bounds__(kThreads, 1)
igemm_kernel(const __grid_constant__ IgemmKParams p) {
    pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    // 1024-B alignment (SWIZZLE_128B atoms)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // stage layout: A | (split3: Alo) | B | (split3: Blo)
    const bool split3 = TF32 && p.split3;
    const uint32_t stage_bytes = (split3 ? 2u : 1u) * (p.a_bytes + p.b_bytes);
    const uint32_t off_b = (split3 ? 2u : 1u) * p.a_bytes;
    uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* ready_bar = empty_bar + 8;                       // split3: Alo of the stage has been written (4 warp arrivals)
    uint64_t* tmem_full_bar = ready_bar + 8;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    float* red_smem = reinterpret_cast<float*>(tail + 256);   // [4 warps][2][n_tile]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile coordinates
    int t = blockIdx.x;
    const int tx = t % p.tiles_x;
    t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int b = t / p.tiles_y;
    const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
    const int n0 = blockIdx.y * p.n_tile;

    int num_kb = 0;
    for (int i = 0; i < p.num_taps; ++i) num_kb += p.src_nblk[p.tap_src[i]];

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&ready_bar[s], 4);
        }
        mbar_init(tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, p.tmem_cols);
        tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.a_maps[0]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int kb = 0;
            for (int tp = 0; tp < p.num_taps; ++tp) {
                const int src = p.tap_src[tp];
                const int cx = x0 + p.tap_dx[tp];
                const int cy = y0 + p.tap_dy[tp];
                const int nblk = p.src_nblk[src];
                for (int cb = 0; cb < nblk; ++cb, ++kb) {
                    const int s = kb % p.stages;
                    const uint32_t ph = (kb / p.stages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + (size_t)s * stage_bytes;
                    uint8_t* sb = sa + off_b;
                    mbar_arrive_expect_tx(&full_bar[s], p.a_bytes + (split3 ? 2u : 1u) * p.b_bytes);
                    tma_load_4d(sa, &p.a_maps[src], &full_bar[s], cb * p.kwidth, cx, cy, b);
                    tma_load_2d(sb, &p.b_map, &full_bar[s], kb * p.kwidth, n0);
                    if (split3) tma_load_2d(sb + p.b_bytes, &p.b_map, &full_bar[s], kb * p.kwidth, p.lo_row0 + n0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = TF32 ? make_idesc_tf32(128, p.n_tile) : make_idesc_bf16(128, p.n_tile, 0, 0);
        const uint32_t layout = (TF32 || p.kwidth == 64) ? 2u : 6u;   // SW128 : SW32
        const uint32_t sbo = (TF32 || p.kwidth == 64) ? 1024u : 256u; // 8 rows x row bytes
        const int full_ksteps = TF32 ? p.kwidth / 8 : p.kwidth / 16;
        int kb = 0;
        for (int tp = 0; tp < p.num_taps; ++tp) {
            const int src = p.tap_src[tp];
            const int nblk = p.src_nblk[src];
            for (int cb = 0; cb < nblk; ++cb, ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (kb / p.stages) & 1;
                mbar_wait(split3 ? &ready_bar[s] : &full_bar[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t sb = sa + off_b;
                    const int ksteps = (cb == nblk - 1) ? p.src_lastk[src] : full_ksteps;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t ad = make_smem_desc(sa + k * 32, 16, sbo, layout);
                        const uint64_t bd = make_smem_desc(sb + k * 32, 16, sbo, layout);
                        if (TF32) umma_tf32(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        else umma_bf16(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        if (split3) {
                            // the tensor core reads the upper 19 bits of an fp32 operand (truncation): A, B act as Ahi, Bhi
                            const uint64_t ald = make_smem_desc(sa + p.a_bytes + k * 32, 16, sbo, layout);
                            const uint64_t bld = make_smem_desc(sb + p.b_bytes + k * 32, 16, sbo, layout);
                            umma_tf32(tmem_base, ad, bld, idesc, 1u);
                            umma_tf32(tmem_base, ald, bd, idesc, 1u);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                    if (kb == num_kb - 1) umma_commit(tmem_full_bar);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;              // TMEM lane quadrant this warp may access
        const int m = q * 32 + lane;         // accumulator row = pixel within the tile
        const int py = m / p.tile_w, px = m - py * p.tile_w;
        const int y = y0 + py, x = x0 + px;
        const int oy = y * p.out_sy + p.out_py, ox = x * p.out_sx + p.out_px;
        const bool valid = (y < p.H) && (x < p.W) && (oy < p.OH) && (ox < p.OW);
        const long long pix = ((long long)b * p.OH + oy) * p.OW + ox;
        const bool do_stats = (p.stats != nullptr);

        if (split3) {
            // ---- 3xTF32 split team: Alo = A - trunc_tf32(A) for every landed stage (exact in fp32), written next to A ----
            const int e = (warp - 2) * 32 + lane;          // 0..127: 16-byte chunk e & 7 of rows e >> 3, e >> 3 + 16, ...
            int kb = 0;
            for (int tp = 0; tp < p.num_taps; ++tp) {
                const int nblk = p.src_nblk[p.tap_src[tp]];
                for (int cb = 0; cb < nblk; ++cb, ++kb) {
                    const int s = kb % p.stages;
                    const uint32_t ph = (kb / p.stages) & 1;
                    mbar_wait(&full_bar[s], ph);
                    const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t off = (uint32_t)((e >> 3) + 16 * i) * 128u + (uint32_t)(e & 7) * 16u;   // same (swizzled) position in both tiles
                        const uint4 v = lds_v4(base + off);
                        uint4 lo;
                        lo.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u));
                        lo.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u));
                        lo.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u));
                        lo.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u));
                        sts_v4(base + p.a_bytes + off, lo);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ready_bar[s]);
                }
            }
        }
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int nchunks = p.n_tile / 16;
        for (int ch = 0; ch < nchunks; ++ch) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + ch * 16, r);
            tmem_ld_wait();
            const int nb = n0 + ch * 16;     // first output channel of this chunk
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = valid ? __uint_as_float(r[j]) : 0.f;
            if (OUT_MODE == 0 && TF32) {
                if (valid) {
                    float* orow = reinterpret_cast<float*>(p.out) + pix * p.ldo + p.coff + nb;
                    if (nb + 16 <= p.N && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0)) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) reinterpret_cast<float4*>(orow)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nb + j < p.N) orow[j] = v[j];
                    }
                }
            } else if (OUT_MODE == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = bf16_round(v[j]);
                if (valid) {
                    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + p.coff + nb;
                    if (nb + 16 <= p.N && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0)) {
                        uint4 w0, w1;
                        w0.x = pack_bf16x2(v[0], v[1]);   w0.y = pack_bf16x2(v[2], v[3]);
                        w0.z = pack_bf16x2(v[4], v[5]);   w0.w = pack_bf16x2(v[6], v[7]);
                        w1.x = pack_bf16x2(v[8], v[9]);   w1.y = pack_bf16x2(v[10], v[11]);
                        w1.z = pack_bf16x2(v[12], v[13]); w1.w = pack_bf16x2(v[14], v[15]);
                        reinterpret_cast<uint4*>(orow)[0] = w0;
                        reinterpret_cast<uint4*>(orow)[1] = w1;
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nb + j < p.N) orow[j] = __float2bfloat16_rn(v[j]);
                    }
                }
            } else {
                // fp32 NCHW: out[((b*N + n)*OH + oy)*OW + ox]
                if (valid) {
                    float* o = reinterpret_cast<float*>(p.out);
                    const long long plane = (long long)p.OH * p.OW;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (nb + j < p.N)
                            o[((long long)b * p.N + nb + j) * plane + (long long)oy * p.OW + ox] = v[j];
                }
            }
            if (do_stats) {
                float sq[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) sq[j] = v[j] * v[j];
                const float s1 = warp_colsum16(v, lane);
                const float s2 = warp_colsum16(sq, lane);
                if ((lane & 1) == 0) {
                    const int col = ch * 16 + ((lane >> 1) & 15);
                    red_smem[(q * 2 + 0) * p.n_tile + col] = s1;
                    red_smem[(q * 2 + 1) * p.n_tile + col] = s2;
                }
            }
        }
        if (do_stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only
            const int te = threadIdx.x - 64;                 // 0..127
            const int slot = blockIdx.x % DMM_STATS_SLOTS;
            for (int col = te; col < p.n_tile; col += 128) {
                if (n0 + col < p.N) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        s1 += red_smem[(w * 2 + 0) * p.n_tile + col];
                        s2 += red_smem[(w * 2 + 1) * p.n_tile + col];
                    }
                    double* st = p.stats + (size_t)slot * 2 * p.stats_ld + p.stats_off + n0 + col;
                    atomicAdd(st, (double)s1);
                    atomicAdd(st + p.stats_ld, (double)s2);
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

static uint32_t tmem_cols_for(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

int igemm2_launch(const dmm_igemm_t* d, cudaStream_t stream);

int view_to_tmap_elem(CUtensorMap* out, const dmm_view_t& v, int box_c, int box_w, int box_h, int swizzle, int elem_bytes) {
    uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.B};
    uint64_t strides[3] = {(uint64_t)v.sw, (uint64_t)v.sh, (uint64_t)v.sb};
    // degenerate dims: the driver still wants valid (16-B multiple) strides
    if (v.H == 1 && strides[1] == 0) strides[1] = (uint64_t)v.W * v.sw;
    if (v.B == 1 && strides[2] == 0) strides[2] = strides[1] * (uint64_t)v.H;
    uint32_t box[4] = {(uint32_t)box_c, (uint32_t)box_w, (uint32_t)box_h, 1u};
    return make_tmap_elem(out, elem_bytes, v.ptr, 4, dims, strides, box, swizzle);
}
int view_to_tmap(CUtensorMap* out, const dmm_view_t& v, int box_c, int box_w, int box_h, int swizzle) {
    return view_to_tmap_elem(out, v, box_c, box_w, box_h, swizzle, 2);
}

}  // namespace dmm

using namespace dmm;

extern "C" int dmm_conv_igemm(const dmm_igemm_t* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DMM_CHECK(d != nullptr, "dmm_conv_igemm: null descriptor");
    DMM_CHECK(d->dtype >= 0 && d->dtype <= 2, "dmm_conv_igemm: dtype must be 0 (bf16), 1 (fp32 storage, tf32 MMA) or 2 (fp32 storage, 3xTF32)");
    DMM_CHECK(d->dtype != 0 ? d->kwidth == 32 : (d->kwidth == 64 || d->kwidth == 32 || d->kwidth == 16), "dmm_conv_igemm: kwidth must be 64, 32 or 16 (32 in the fp32 mode), got %d", d->kwidth);
    DMM_CHECK(d->num_src >= 1 && d->num_src <= DMM_MAX_SRC, "dmm_conv_igemm: bad num_src %d", d->num_src);
    DMM_CHECK(d->num_taps >= 1 && d->num_taps <= DMM_MAX_TAPS, "dmm_conv_igemm: bad num_taps %d", d->num_taps);
    DMM_CHECK(d->tile_w == 128 || d->tile_w == 64 || d->tile_w == 32 || d->tile_w == 16 || d->tile_w == 8,
              "dmm_conv_igemm: bad tile_w %d", d->tile_w);
    DMM_CHECK(d->n_tile >= 16 && d->n_tile <= 256 && d->n_tile % 16 == 0, "dmm_conv_igemm: bad n_tile %d", d->n_tile);
    DMM_CHECK(d->N >= 1 && d->out != nullptr && d->weights != nullptr, "dmm_conv_igemm: bad output/weights");
    DMM_CHECK(d->out_mode >= 0 && d->out_mode <= 3, "dmm_conv_igemm: bad out_mode %d", d->out_mode);
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;
    {
        // v2 (persistent, halo patches in shared memory) handles every kwidth-64 launch; DMM_IGEMM_V1=1 keeps v1
        static const bool force_v1 = getenv("DMM_IGEMM_V1") != nullptr && atoi(getenv("DMM_IGEMM_V1")) != 0;
        // (and the kwidth-16 launches whose single tapped source has <= 16 channels: four taps share a 64-wide K block)
        bool packed16 = d->kwidth == 16 || (d->kwidth == 32 && d->dtype == 0);      // several taps share one 64-wide K block
        int tapped = 0;
        for (int s = 0; s < d->num_src && packed16; ++s) {
            bool used = false;
            for (int t = 0; t < d->num_taps; ++t) used = used || d->tap_src[t] == s;
            if (used) { ++tapped; packed16 = d->src[s].C <= d->kwidth; }
        }
        packed16 = packed16 && tapped == 1;
        if (d->dtype == 0 && (d->kwidth == 64 || packed16) && (!force_v1 || d->out_mode >= 2)) return igemm2_launch(d, stream);
    }
    DMM_CHECK(d->out_mode < 2, "dmm_conv_igemm: out_mode 2 / 3 need kwidth 64");
    const bool tf32 = d->dtype == 1 || d->dtype == 2;
    const bool split3 = d->dtype == 2;
    const int esz = tf32 ? 4 : 2;
    const int kstep = tf32 ? 8 : 16;
    if (split3) DMM_CHECK(d->n_rows % d->n_tile == 0, "dmm_conv_igemm: 3xTF32 needs n_rows to be a multiple of n_tile");
    if (tf32) DMM_CHECK(d->kwidth == 32 && !d->pro_enable && d->bnb_sums == nullptr && d->fold_kw == 0,
                        "dmm_conv_igemm: the fp32 / tf32 mode needs kwidth 32 and no fused prologue / BN-backward epilogue");

    IgemmKParams p;
    memset(&p, 0, sizeof(p));
    p.kwidth = d->kwidth;
    p.tile_w = d->tile_w;
    p.tile_h = 128 / d->tile_w;
    const int swz = (tf32 || d->kwidth == 64) ? 128 : 32;
    long long ktot = 0;
    for (int s = 0; s < d->num_src; ++s) {
        const dmm_view_t& v = d->src[s];
        DMM_CHECK(v.ptr != nullptr && v.C >= 1, "dmm_conv_igemm: source %d empty", s);
        p.src_nblk[s] = ceil_div(v.C, d->kwidth);
        const int rem = v.C - (p.src_nblk[s] - 1) * d->kwidth;
        p.src_lastk[s] = ceil_div(rem, kstep);
        int rc = view_to_tmap_elem(&p.a_maps[s], v, d->kwidth, p.tile_w, p.tile_h, swz, esz);
        if (rc) return rc;
    }
    p.num_taps = d->num_taps;
    for (int t = 0; t < d->num_taps; ++t) {
        DMM_CHECK(d->tap_src[t] >= 0 && d->tap_src[t] < d->num_src, "dmm_conv_igemm: tap %d bad source", t);
        p.tap_src[t] = d->tap_src[t];
        p.tap_dy[t] = d->tap_dy[t];
        p.tap_dx[t] = d->tap_dx[t];
        ktot += (long long)p.src_nblk[d->tap_src[t]] * d->kwidth;
    }
    DMM_CHECK(ktot == d->ktot, "dmm_conv_igemm: packed weight K (%lld) != tap table K (%lld)", (long long)d->ktot, ktot);
    DMM_CHECK(d->n_rows >= d->N, "dmm_conv_igemm: weight rows %d < N %d", d->n_rows, d->N);
    {
        uint64_t dims[2] = {(uint64_t)d->ktot, (uint64_t)d->n_rows * (split3 ? 2u : 1u)};      // split3: [Bhi ; Blo] stacked
        uint64_t strides[1] = {(uint64_t)d->ktot};
        uint32_t box[2] = {(uint32_t)d->kwidth, (uint32_t)d->n_tile};
        int rc = make_tmap_elem(&p.b_map, esz, d->weights, 2, dims, strides, box, swz);
        if (rc) return rc;
    }
    p.W = d->W; p.H = d->H; p.B = d->B;
    p.tiles_x = ceil_div(d->W, p.tile_w);
    p.tiles_y = ceil_div(d->H, p.tile_h);
    p.n_tile = d->n_tile;
    p.N = d->N;
    p.a_bytes = 128u * d->kwidth * (uint32_t)esz;
    p.b_bytes = ((uint32_t)d->n_tile * d->kwidth * (uint32_t)esz + 1023u) & ~1023u;
    p.split3 = split3 ? 1 : 0;
    p.lo_row0 = d->n_rows;
    const uint32_t stage_bytes = (split3 ? 2u : 1u) * (p.a_bytes + p.b_bytes);
    // two co-resident CTAs per SM when the stage is small enough, else one with a deeper ring
    const uint32_t budget = (stage_bytes * 3 <= 100 * 1024) ? 100 * 1024 : 200 * 1024;
    int stages = (int)(budget / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    p.stages = stages;
    p.tmem_cols = tmem_cols_for(d->n_tile);
    p.out = d->out;
    p.ldo = d->ldo;
    p.coff = d->coff;
    p.out_sy = d->out_sy > 0 ? d->out_sy : 1;
    p.out_sx = d->out_sx > 0 ? d->out_sx : 1;
    p.out_py = d->out_py; p.out_px = d->out_px;
    p.OH = d->OH > 0 ? d->OH : d->H;
    p.OW = d->OW > 0 ? d->OW : d->W;
    p.stats = d->stats;
    p.stats_ld = d->stats_ld;
    p.stats_off = d->stats_off;

    const size_t smem = (size_t)stages * stage_bytes + 256 + (size_t)8 * d->n_tile * sizeof(float) + 1024;
    dim3 grid((unsigned)((long long)p.tiles_x * p.tiles_y * d->B), (unsigned)ceil_div(d->N, d->n_tile), 1);
    static bool attr_set = false;      // opt in to the full 227 KB of shared memory once (not a stream operation)
    if (!attr_set) {
        DMM_CUDA(cudaFuncSetAttribute(igemm_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        DMM_CUDA(cudaFuncSetAttribute(igemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        DMM_CUDA(cudaFuncSetAttribute(igemm_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        DMM_CUDA(cudaFuncSetAttribute(igemm_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        attr_set = true;
    }
    DMM_CHECK(smem <= (size_t)kMaxSmem, "dmm_conv_igemm: %zu bytes of shared memory requested", smem);
    if (tf32) {
        if (d->out_mode == 0) launch_k(igemm_kernel<0, true>, grid, kThreads, smem, stream, p);
        else launch_k(igemm_kernel<1, true>, grid, kThreads, smem, stream, p);
    } else if (d->out_mode == 0) launch_k(igemm_kernel<0, false>, grid, kThreads, smem, stream, p);
    else launch_k(igemm_kernel<1, false>, grid, kThreads, smem, stream, p);
    DMM_LAUNCH_CHECK("igemm_kernel");
    return 0;
}
