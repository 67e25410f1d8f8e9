// Weight-gradient GEMM for sm_100a:  D[row][col] += sum_pixels A(pix + shift_a)[row] * B(pix + shift_b)[col].
//
// The reduction (K) dimension is the pixel index, so with pixel-major activations both operands are
// "MN-major" for tcgen05: a TMA box {channels, tile_w, tile_h, 1} lands in shared memory as pixel-rows of
// 2*channels bytes - exactly the canonical MN-major swizzled layout (8 k-rows per atom, SBO = 8 rows,
// LBO = distance between channel atoms).
//
// Both operands stream from HBM/L2 with no reuse along K, so the arithmetic intensity of a CTA is
// M*N/(M+N) flop/B: one CTA therefore owns up to FOUR 128-row accumulators x several column groups (all of
// TMEM, 512 columns) at once.  The A side is a list of 64-channel chunks (m-tile = 2 chunks), the B side a list of
// n_tile-channel groups; every (m-tile, group) pair has its own TMEM accumulator.  Each chunk / group carries its
// own source view and pixel shift.
//   * "family" mode (KxK convolution, A = activation, B = the narrow output gradient): all B groups are shifted
//     views of ONE source, so a stage holds a single halo patch of it ((tile_h + kh - 1) x (8 + kw - 1) pixels) and
//     every tap is a shifted shared-memory descriptor into that patch (tile_w = 8: an 8-row k-group is one patch row
//     segment, SBO = patch row pitch).  The kw taps of one kernel row are one pixel apart, i.e. ONE MMA with
//     N = kw * n_tile and LBO = one pixel row covers them (overlapping atoms).
//   * general mode: every group is its own TMA box; consecutive groups are merged into MMAs of N <= 256.
// A CTA handles a contiguous range of pixel tiles (split-K over the grid, one wave) and adds its fp32 partial result
// to the scratch matrix dw[row*ld + col] with TMA reduce-add (cp.reduce.async.bulk.tensor) from a swizzled staging
// buffer - a few dozen bulk operations per CTA instead of tens of thousands of per-lane atomics.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue.
#include "common.cuh"
#include "../../include/dmmfods_b200.h"

#include <stdlib.h>
#include <string.h>

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
    CUtensorMap dw_map;
    int a_C[DMM_MAX_SRC], b_C[DMM_MAX_SRC];
    WgSlot a[DMM_WG_MAX_A];
    WgSlot b[DMM_WG_MAX_B];
    int num_a, num_b, na;          // na = m-tiles = ceil(num_a / 2)
    int n_tile, bw, b_chunks;      // group width, TMA box width (channels), boxes per group
    uint32_t b_layout, b_sbo;      // UMMA layout type / stride between 8-pixel k-groups of the B operand
    uint32_t b_kstep;              // bytes per 16-pixel k-step of the B operand
    uint32_t b_goff[DMM_WG_MAX_B]; // byte offset of group g inside the B part of a stage
    int family;                    // 1: one halo patch of source b[0].map per stage
    int fam_ox, fam_oy;
    uint32_t fam_bytes;
    int num_sg;                    // merged MMAs: super-group i = groups [sg_first, sg_first + sg_count), atom stride sg_lbo
    int sg_first[DMM_WG_MAX_B], sg_count[DMM_WG_MAX_B];
    uint32_t sg_lbo[DMM_WG_MAX_B];
    int ya, yb, a_step, b_step;
    int kpx, tile_w, tile_h, tiles_x, tiles_y;
    FastDiv fd_x, fd_y;
    long long total_tiles;
    int splits;
    int stages;
    uint32_t a_chunk_bytes, b_box_bytes, stage_bytes, tmem_cols;
    long long* prof;
    int pro;                       // BN-ReLU prologue on the A chunks (all from a_src[0], unshifted)
    int pro_kp;                    // a_C[0] rounded up to 64
    int pro_mask;                  // B groups are shifted (K x K): A rows outside the image must be zero (padding of the activation)
    int W, H, tw_log;              // image size and log2(tile_w) for that test
    const float* pro_gamma;
    const float* pro_beta;
    const float* pro_mean;
    const float* pro_invstd;
};

constexpr int kWgThreads = 192;
constexpr int kWgThreadsPro = 320;       // + warps 6..9: extra BN-ReLU prologue warps (they only transform A chunks)

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile(
        "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
            reinterpret_cast<uint64_t>(m)),
        "r"(smem_u32(src)), "r"(c0), "r"(c1)
        : "memory");
}

__global__ void __launch_bounds__(kWgThreadsPro, 1) wgrad_kernel(const __grid_constant__ WgradKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const size_t ring_bytes = (size_t)p.stages * p.stage_bytes;
    uint8_t* tail = smem + (ring_bytes > 32768 ? ring_bytes : 32768);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tmem_full_bar = empty_bar + 8;
    uint64_t* ready_bar = tmem_full_bar + 1;                              // pro: A chunks of the stage transformed (4 warp arrivals)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(ready_bar + 8);
    float* pcoef = reinterpret_cast<float*>(tail + 256);                  // pro: [2][pro_kp] scale / shift

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ia = blockIdx.y % p.ya, ib = blockIdx.y / p.ya;
    const int a_off = ia * p.a_step, b_off = ib * p.b_step;
    const long long tile_lo = p.total_tiles * blockIdx.x / p.splits;
    const long long tile_hi = p.total_tiles * (blockIdx.x + 1) / p.splits;
    const int num_k = (int)(tile_hi - tile_lo);
    const uint32_t a_bytes = p.a_chunk_bytes * 2 * p.na;
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < p.stages; ++s) mbar_init(&ready_bar[s], 8);
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
        tma_prefetch_desc(&p.dw_map);
    }
    // everything above is independent of earlier kernels: wait for them (programmatic dependent launch) only here
    pdl_prologue();
    if (p.pro) {
        for (int c = threadIdx.x; c < p.pro_kp; c += blockDim.x) {
            float sc = 0.f, sh = 0.f;
            if (c < p.a_C[0]) {
                const float mu = p.pro_mean[c];
                sc = (p.pro_gamma ? p.pro_gamma[c] : 1.f) * p.pro_invstd[c];
                sh = (p.pro_beta ? p.pro_beta[c] : 0.f) - mu * sc;
            }
            pcoef[c] = sc;
            pcoef[p.pro_kp + c] = sh;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (num_k > 0) {
        if (warp == 0) {
            // ================= TMA producer =================
            if (lane == 0) {
                // bytes that actually arrive per stage: chunks that lie completely outside their view are skipped
                uint32_t tx_bytes = 0;
                for (int i = 0; i < p.num_a; ++i)
                    if (p.a[i].ch0 + a_off < p.a_C[p.a[i].map]) tx_bytes += p.a_chunk_bytes;
                if (p.family) tx_bytes += p.fam_bytes;
                else
                    for (int g = 0; g < p.num_b; ++g)
                        for (int c = 0; c < p.b_chunks; ++c)
                            if (p.b[g].ch0 + b_off + c * p.bw < p.b_C[p.b[g].map]) tx_bytes += p.b_box_bytes;
                long long w_e = 0;
                const long long t_begin = clock64();
                int s = 0;
                uint32_t ph = 0;
                for (int kb = 0; kb < num_k; ++kb) {
                    uint32_t tx, ty;
                    const uint32_t trow = fast_divmod((uint32_t)(tile_lo + kb), p.fd_x, tx);      // total_tiles < 2^31 (checked by the launcher)
                    const int b = (int)fast_divmod(trow, p.fd_y, ty);
                    const int x0 = (int)tx * p.tile_w, y0 = (int)ty * p.tile_h;
#ifdef DMM_IGEMM_PHASE_PROF
                    const long long c0 = clock64();
#endif
                    mbar_wait(&empty_bar[s], ph ^ 1);
#ifdef DMM_IGEMM_PHASE_PROF
                    w_e += clock64() - c0;
#endif
                    uint8_t* sa = smem + (size_t)s * p.stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                    for (int i = 0; i < p.num_a; ++i) {
                        const WgSlot& c = p.a[i];
                        if (c.ch0 + a_off < p.a_C[c.map])
                            tma_load_4d(sa + i * p.a_chunk_bytes, &p.a_maps[c.map], &full_bar[s], c.ch0 + a_off, x0 + c.dx,
                                        y0 + c.dy, b);
                    }
                    if (p.family) {
                        tma_load_4d(sb, &p.b_maps[p.b[0].map], &full_bar[s], p.b[0].ch0 + b_off, x0 + p.fam_ox, y0 + p.fam_oy, b);
                    } else {
                        for (int g = 0; g < p.num_b; ++g) {
                            const WgSlot& c = p.b[g];
                            for (int k = 0; k < p.b_chunks; ++k)
                                if (c.ch0 + b_off + k * p.bw < p.b_C[c.map])
                                    tma_load_4d(sb + p.b_goff[g] + k * p.b_box_bytes, &p.b_maps[c.map], &full_bar[s],
                                                c.ch0 + b_off + k * p.bw, x0 + c.dx, y0 + c.dy, b);
                        }
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                if (p.prof) {
                    p.prof[cta * 8 + 0] = clock64() - t_begin;
                    p.prof[cta * 8 + 1] = w_e;
                }
            }
        } else if (warp == 1) {
            // ================= MMA issuer =================
            // elect.sync-guarded straight-line issue (see igemm2.cu).  An M=128, K=16 MMA streams its operands from
            // shared memory at 128 B/cycle: (4096 + 32 N) / 128 cycles, so ONE wide MMA over several B groups
            // (N = count * n_tile <= 256, atom stride sg_lbo) is far cheaper than one N = n_tile MMA per group.
            const int ksteps = p.kpx / 16;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t smem_u = smem_u32(smem);
            const uint64_t ahi = (uint64_t)(((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29)) << 32;
            const uint64_t bhi = (uint64_t)(((p.b_sbo >> 4) & 0x3FFFu) | (1u << 14) | ((p.b_layout & 7u) << 29)) << 32;
            const uint32_t a_lbo = ((p.a_chunk_bytes >> 4) & 0x3FFFu) << 16;
            const uint32_t a_mt = (2u * p.a_chunk_bytes) >> 4, b_ks = p.b_kstep >> 4;
            // one elected thread runs the whole issue loop, waits included (no per-stage elect / reconvergence / warp sync); the
            // wait counter costs two clock reads per stage and is compiled in only with the phase profile
            if (elect_one()) {
            long long w_f = 0;
            const long long t_begin = clock64();
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < num_k; ++kb) {
#ifdef DMM_IGEMM_PHASE_PROF
                const long long c0 = clock64();
#endif
                mbar_wait(p.pro ? &ready_bar[s] : &full_bar[s], ph);
#ifdef DMM_IGEMM_PHASE_PROF
                w_f += clock64() - c0;
#endif
                tc_fence_after();
                {
                    const uint32_t sa = ((smem_u + (uint32_t)s * p.stage_bytes) >> 4);
                    const uint32_t sb = sa + (a_bytes >> 4);
                    for (int i = 0; i < p.num_sg; ++i) {
                        const int g0 = p.sg_first[i];
                        const uint32_t idesc = make_idesc_bf16(128, p.sg_count[i] * p.n_tile, 1, 1);
                        const uint32_t b0 = sb + (p.b_goff[g0] >> 4);
                        const uint32_t b_lbo = ((p.sg_lbo[i] >> 4) & 0x3FFFu) << 16;
                        for (int mi = 0; mi < p.na; ++mi) {
                            const uint32_t d = tmem_u + (uint32_t)((mi * p.num_b + g0) * p.n_tile);
                            const uint32_t a0 = sa + mi * a_mt;
                            for (int k = 0; k < ksteps; ++k) {
                                const uint64_t ad = ahi | (uint64_t)(((a0 + k * 128u) & 0x3FFFu) | a_lbo);
                                const uint64_t bd = bhi | (uint64_t)(((b0 + k * b_ks) & 0x3FFFu) | b_lbo);
                                umma_bf16(d, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            }
                        }
                    }
                    umma_commit(&empty_bar[s]);
                    if (kb == num_k - 1) umma_commit(tmem_full_bar);
                }
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            if (p.prof) {
                p.prof[cta * 8 + 2] = clock64() - t_begin;
                p.prof[cta * 8 + 3] = w_f;
            }
            }   // elected thread
        } else {
            // ================= epilogue: TMEM -> swizzled fp32 staging -> TMA reduce-add into dw =================
            const int q = warp & 3;                       // TMEM lane quadrant = rows q*32 .. q*32+31 of the m-tile
            if (p.pro) {
                // the epilogue warps are idle during the main loop: they apply relu(bn(x)) in place to the A chunks of every
                // stage (pixel rows of 128 bytes, swizzled 16-byte chunks; zero-filled out-of-image pixels meet zero B rows)
                const int e = (warp - 2) * 32 + lane;     // 0..255: warps 2..9
                const int j = e & 7;
                int s = 0;
                uint32_t ph = 0;
                for (int kb = 0; kb < num_k; ++kb) {
                    int mx0 = 0, my0 = 0;
                    if (p.pro_mask) {
                        uint32_t tx, ty;
                        fast_divmod(fast_divmod((uint32_t)(tile_lo + kb), p.fd_x, tx), p.fd_y, ty);
                        mx0 = (int)tx * p.tile_w;
                        my0 = (int)ty * p.tile_h;
                    }
                    mbar_wait(&full_bar[s], ph);
                    uint8_t* sa = smem + (size_t)s * p.stage_bytes;
                    for (int i = 0; i < p.num_a; ++i) {
                        const int ch = p.a[i].ch0 + a_off;
                        if (ch >= p.a_C[0]) continue;     // chunk outside the view: nothing was loaded
                        float sc[8], sh[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            sc[t] = pcoef[ch + j * 8 + t];
                            sh[t] = pcoef[p.pro_kp + ch + j * 8 + t];
                        }
                        const uint32_t base = smem_u32(sa + i * p.a_chunk_bytes);
#pragma unroll 4
                        for (int r = e >> 3; r < p.kpx; r += 32) {
                            const uint32_t ptr = base + r * 128 + ((j ^ (r & 7)) << 4);
                            if (p.pro_mask) {
                                const int yy = my0 + p.a[i].dy + (r >> p.tw_log), xx = mx0 + p.a[i].dx + (r & (p.tile_w - 1));
                                if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) {
                                    sts_v4(ptr, make_uint4(0, 0, 0, 0));
                                    continue;
                                }
                            }
                            const uint4 v = lds_v4(ptr);
                            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float lo = fmaxf(fmaf(bf16_lo(w[t]), sc[2 * t], sh[2 * t]), 0.f);
                                const float hi = fmaxf(fmaf(bf16_hi(w[t]), sc[2 * t + 1], sh[2 * t + 1]), 0.f);
                                w[t] = pack_bf16x2(lo, hi);
                            }
                            sts_v4(ptr, make_uint4(w[0], w[1], w[2], w[3]));
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ready_bar[s]);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
            if (warp >= 6) goto done;                     // prologue-only warps have no epilogue work
            const long long t_begin = clock64();
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
            const long long t_mid = clock64();
            // all stages have been consumed (every MMA has completed): the ring is reused as staging, 2 x 4 KB per warp
            uint8_t* slots = smem + q * 8192;
            uint32_t n_ops = 0;
            for (int mi = 0; mi < p.na; ++mi) {
                const int ci = mi * 2 + (q >> 1);         // A chunk of this warp's rows
                if (ci >= p.num_a) continue;
                const WgSlot& ca = p.a[ci];
                if (ca.ch0 + a_off >= p.a_C[ca.map]) continue;                 // chunk outside its view: nothing was loaded
                const int row0 = ca.out0 + a_off + (q & 1) * 32;
                for (int g = 0; g < p.num_b; ++g) {
                    const WgSlot& cb = p.b[g];
                    if (cb.ch0 + b_off >= p.b_C[cb.map]) continue;             // group outside its view
                    const int col0 = cb.out0 + b_off;
                    for (int pc = 0; pc < p.n_tile; pc += 32) {
                        const int w = (p.n_tile - pc) < 32 ? (p.n_tile - pc) : 32;        // 16 or 32 columns
                        uint8_t* slot = slots + (n_ops & 1) * 4096;
                        uint32_t v0[16], v1[16];
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((mi * p.num_b + g) * p.n_tile + pc);
                        tmem_ld16(taddr, v0);
                        if (w > 16) tmem_ld16(taddr + 16, v1);
                        if (lane == 0 && n_ops >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        tmem_ld_wait();
                        __syncwarp();
                        const uint32_t srow = smem_u32(slot) + lane * 128;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            sts_v4(srow + ((j ^ (lane & 7)) << 4), make_uint4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]));
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            sts_v4(srow + (((j + 4) ^ (lane & 7)) << 4),
                                   (w > 16) ? make_uint4(v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]) : make_uint4(0, 0, 0, 0));
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_reduce_add_2d(&p.dw_map, slot, col0 + pc, row0);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        ++n_ops;
                    }
                }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            if (p.prof && threadIdx.x == 64) {
                p.prof[cta * 8 + 4] = clock64() - t_begin;
                p.prof[cta * 8 + 5] = t_mid - t_begin;
            }
        }
    }

done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

int view_to_tmap(CUtensorMap* out, const dmm_view_t& v, int box_c, int box_w, int box_h, int swizzle);

static int wg_env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

}  // namespace dmm

using namespace dmm;

extern "C" int dmm_conv_wgrad(const dmm_wgrad_t* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DMM_CHECK(d != nullptr && d->dw != nullptr, "dmm_conv_wgrad: null descriptor / output");
    DMM_CHECK(d->num_a >= 1 && d->num_a <= DMM_WG_MAX_A, "dmm_conv_wgrad: bad num_a %d", d->num_a);
    DMM_CHECK(d->num_b >= 1 && d->num_b <= DMM_WG_MAX_B, "dmm_conv_wgrad: bad num_b %d", d->num_b);
    DMM_CHECK(d->num_a_src >= 1 && d->num_a_src <= DMM_MAX_SRC && d->num_b_src >= 1 && d->num_b_src <= DMM_MAX_SRC,
              "dmm_conv_wgrad: bad source counts");
    DMM_CHECK(d->kpx == 128 || d->kpx == 64 || d->kpx == 32, "dmm_conv_wgrad: kpx must be 32, 64 or 128 (got %d)", d->kpx);
    DMM_CHECK(d->tile_w >= 8 && d->tile_w <= d->kpx && (d->tile_w & (d->tile_w - 1)) == 0, "dmm_conv_wgrad: bad tile_w %d", d->tile_w);
    DMM_CHECK(d->n_tile >= 16 && d->n_tile <= 256 && d->n_tile % 16 == 0, "dmm_conv_wgrad: bad n_tile %d", d->n_tile);
    DMM_CHECK(d->n_tile == 16 || d->n_tile == 32 || d->n_tile >= 64, "dmm_conv_wgrad: n_tile %d (16, 32 or >= 64)", d->n_tile);
    DMM_CHECK(d->ld % 4 == 0, "dmm_conv_wgrad: ld %lld must be a multiple of 4", (long long)d->ld);
    const int na = (d->num_a + 1) / 2;
    DMM_CHECK(na * d->num_b * d->n_tile <= 512, "dmm_conv_wgrad: %d m-tiles x %d groups x %d columns exceed the 512 TMEM columns",
              na, d->num_b, d->n_tile);
    DMM_CHECK(d->ya >= 1 && d->yb >= 1, "dmm_conv_wgrad: bad grid replication");
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        DMM_CUDA(cudaGetDevice(&dev));
        DMM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }

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

    // ---- family mode: all B groups are shifted views of one source (and nothing is replicated along B) ----
    static const int no_family = wg_env_int("DMM_WGRAD_NO_FAMILY", 0);
    int mindx = 127, maxdx = -128, mindy = 127, maxdy = -128;
    bool family = d->num_b > 1 && d->tile_w == 8 && p.b_chunks == 1 && d->yb == 1 && !no_family;
    for (int g = 0; g < d->num_b; ++g) {
        if (d->b[g].src != d->b[0].src || d->b[g].ch0 != d->b[0].ch0) family = false;
        mindx = d->b[g].dx < mindx ? d->b[g].dx : mindx; maxdx = d->b[g].dx > maxdx ? d->b[g].dx : maxdx;
        mindy = d->b[g].dy < mindy ? d->b[g].dy : mindy; maxdy = d->b[g].dy > maxdy ? d->b[g].dy : maxdy;
    }
    const uint32_t row_bytes = (uint32_t)b_swz;
    p.a_chunk_bytes = (uint32_t)d->kpx * 128u;
    p.b_box_bytes = (uint32_t)d->kpx * row_bytes;
    uint32_t b_part;
    p.family = family ? 1 : 0;
    if (family) {
        const int pw = 8 + (maxdx - mindx), ph = p.tile_h + (maxdy - mindy);
        p.fam_ox = mindx; p.fam_oy = mindy;
        p.fam_bytes = (uint32_t)pw * ph * row_bytes;
        p.b_sbo = (uint32_t)pw * row_bytes;               // next 8-pixel k-group = next patch row
        p.b_kstep = 2u * pw * row_bytes;                  // 16 pixels = two tile rows
        for (int g = 0; g < d->num_b; ++g)
            p.b_goff[g] = (uint32_t)((d->b[g].dy - mindy) * pw + (d->b[g].dx - mindx)) * row_bytes;
        b_part = p.fam_bytes;
        // merge runs of groups one pixel apart in x (same dy) - or one patch row apart in y (same dx) - that are n_tile columns
        // apart in dw: one MMA of N = count * n_tile whose atom stride (LBO) is one pixel / one patch row
        int nsg = 0;
        for (int g = 0; g < d->num_b;) {
            int c = 1;
            while (g + c < d->num_b && (c + 1) * d->n_tile <= 256 && d->b[g + c].dy == d->b[g].dy && d->b[g + c].dx == d->b[g].dx + c &&
                   d->b[g + c].out0 == d->b[g].out0 + c * d->n_tile)
                ++c;
            uint32_t lbo = row_bytes;
            if (c == 1) {
                while (g + c < d->num_b && (c + 1) * d->n_tile <= 256 && d->b[g + c].dx == d->b[g].dx && d->b[g + c].dy == d->b[g].dy + c &&
                       d->b[g + c].out0 == d->b[g].out0 + c * d->n_tile)
                    ++c;
                if (c > 1) lbo = (uint32_t)pw * row_bytes;
            }
            p.sg_first[nsg] = g; p.sg_count[nsg] = c; p.sg_lbo[nsg] = lbo;
            ++nsg;
            g += c;
        }
        p.num_sg = nsg;
    } else {
        p.b_sbo = 8u * row_bytes;
        p.b_kstep = 16u * row_bytes;
        const uint32_t b_group_bytes = p.b_box_bytes * p.b_chunks;
        for (int g = 0; g < d->num_b; ++g) p.b_goff[g] = (uint32_t)g * b_group_bytes;
        b_part = b_group_bytes * d->num_b;
        // consecutive groups sit b_group_bytes apart: merge when that equals n_tile columns of atoms and dw columns follow suit
        const bool mergeable = (d->n_tile % p.bw) == 0;
        const int per = mergeable ? (256 / d->n_tile > 0 ? 256 / d->n_tile : 1) : 1;
        int nsg = 0;
        for (int g = 0; g < d->num_b;) {
            const int left = d->num_b - g;
            const int parts = (left + per - 1) / per;
            const int want = (left + parts - 1) / parts;              // balanced split: 9 groups of 32 -> 5 + 4
            int c = 1;
            while (c < want && d->b[g + c].out0 == d->b[g].out0 + c * d->n_tile) ++c;
            p.sg_first[nsg] = g; p.sg_count[nsg] = c; p.sg_lbo[nsg] = p.b_box_bytes;
            ++nsg;
            g += c;
        }
        p.num_sg = nsg;
    }
    for (int s = 0; s < d->num_a_src; ++s) {
        DMM_CHECK(d->a_src[s].ptr != nullptr, "dmm_conv_wgrad: A source %d is null", s);
        int rc = view_to_tmap(&p.a_maps[s], d->a_src[s], 64, p.tile_w, p.tile_h, 128);
        if (rc) return rc;
        p.a_C[s] = d->a_src[s].C;
    }
    for (int s = 0; s < d->num_b_src; ++s) {
        DMM_CHECK(d->b_src[s].ptr != nullptr, "dmm_conv_wgrad: B source %d is null", s);
        int rc;
        if (family) rc = view_to_tmap(&p.b_maps[s], d->b_src[s], p.bw, 8 + (maxdx - mindx), p.tile_h + (maxdy - mindy), b_swz);
        else rc = view_to_tmap(&p.b_maps[s], d->b_src[s], p.bw, p.tile_w, p.tile_h, b_swz);
        if (rc) return rc;
        p.b_C[s] = d->b_src[s].C;
    }
    if (d->pro_enable) {
        DMM_CHECK(d->pro_mean && d->pro_invstd, "dmm_conv_wgrad: prologue BatchNorm without statistics");
        for (int i = 0; i < d->num_a; ++i)
            DMM_CHECK(d->a[i].src == 0 && d->a[i].dx == 0 && d->a[i].dy == 0, "dmm_conv_wgrad: the prologue needs unshifted A chunks of source 0");
        p.pro = 1;
        p.pro_kp = ceil_div(d->a_src[0].C + (d->ya - 1) * d->a_step, 64) * 64 + 64 * DMM_WG_MAX_A;
        p.pro_gamma = d->pro_gamma; p.pro_beta = d->pro_beta; p.pro_mean = d->pro_mean; p.pro_invstd = d->pro_invstd;
        // shifted B groups pair an out-of-image A row with an in-image B row: such A rows must be the zero padding of the activation
        bool shifted = false;
        for (int g = 0; g < d->num_b; ++g) shifted = shifted || d->b[g].dx != 0 || d->b[g].dy != 0;
        p.pro_mask = shifted ? 1 : 0;
        p.W = d->W; p.H = d->H;
        p.tw_log = 0;
        while ((1 << p.tw_log) < p.tile_w) ++p.tw_log;
        DMM_CHECK((1 << p.tw_log) == p.tile_w, "dmm_conv_wgrad: tile_w must be a power of two");
    }
    p.tiles_x = ceil_div(d->W, p.tile_w);
    p.tiles_y = ceil_div(d->H, p.tile_h);
    p.total_tiles = (long long)p.tiles_x * p.tiles_y * d->B;
    DMM_CHECK(p.total_tiles < (1ll << 31), "dmm_conv_wgrad: %lld tiles", p.total_tiles);
    p.fd_x = make_fastdiv(p.tiles_x); p.fd_y = make_fastdiv(p.tiles_y);
    p.stage_bytes = p.a_chunk_bytes * 2 * na + b_part;
    p.stage_bytes = (p.stage_bytes + 1023u) & ~1023u;
    int stages = (int)((200u * 1024u) / p.stage_bytes);
    DMM_CHECK(stages >= 2, "dmm_conv_wgrad: stage of %u bytes does not fit twice in shared memory", p.stage_bytes);
    if (stages > 8) stages = 8;
    p.stages = stages;
    uint32_t cols = 32;
    while ((int)cols < na * d->num_b * d->n_tile) cols <<= 1;
    p.tmem_cols = cols;
    {
        // dw scratch as an fp32 tensor [rows][ld]; rows = the highest row any replica can touch
        long long rows = 0;
        for (int i = 0; i < d->num_a; ++i) {
            const long long r = (long long)d->a[i].out0 + 64 + (long long)(d->ya - 1) * d->a_step;
            rows = r > rows ? r : rows;
        }
        int rc = make_tmap_f32_2d(&p.dw_map, d->dw, (uint64_t)d->ld, (uint64_t)rows, (uint64_t)d->ld, 32, 32, 128);
        if (rc) return rc;
    }
    long long splits = d->splits;
    const long long items = (long long)d->ya * d->yb;
    if (splits <= 0) {
        splits = num_sms / items;                         // one wave: every CTA is resident at once
        static const int split_div = wg_env_int("DMM_WGRAD_SPLIT_DIV", 1);
        if (split_div > 1) splits = splits / split_div > 0 ? splits / split_div : 1;
        const long long min_k = 8;                        // amortise prologue + reduce epilogue
        if (splits > p.total_tiles / min_k) splits = p.total_tiles / min_k;
        // experiment knob: few-pixel launches on fewer SMs (longer K loops, SMs left to the other stream's kernels)
        static const int small_ctas = wg_env_int("DMM_WGRAD_SMALL_CTAS", 0), small_tiles = wg_env_int("DMM_WGRAD_SMALL_TILES", 1200);
        if (small_ctas > 0 && p.total_tiles <= small_tiles && splits * items > small_ctas) splits = small_ctas / items > 0 ? small_ctas / items : 1;
    }
    if (splits > p.total_tiles) splits = p.total_tiles;
    if (splits < 1) splits = 1;
    p.splits = (int)splits;

    size_t ring = (size_t)stages * p.stage_bytes;
    if (ring < 32768) ring = 32768;
    const size_t smem = ring + 256 + 1024 + (p.pro ? (size_t)2 * p.pro_kp * sizeof(float) : 0);
    static bool attr_set = false;
    if (!attr_set) {
        DMM_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set = true;
    }
    DMM_CHECK(smem <= 232448, "dmm_conv_wgrad: %zu bytes of shared memory requested", smem);
    dim3 grid((unsigned)p.splits, (unsigned)items, 1);
    static const int prof = wg_env_int("DMM_WGRAD_PROF", 0);
    static long long* prof_buf = nullptr;
    const size_t prof_n = 4096 * 8;
    if (prof && (size_t)grid.x * grid.y * 8 <= prof_n) {
        if (!prof_buf) DMM_CUDA(cudaMalloc(&prof_buf, prof_n * sizeof(long long)));
        DMM_CUDA(cudaMemsetAsync(prof_buf, 0, prof_n * sizeof(long long), stream));
        p.prof = prof_buf;
    }
    launch_k(wgrad_kernel, grid, p.pro ? kWgThreadsPro : kWgThreads, smem, stream, p);
    DMM_LAUNCH_CHECK("wgrad_kernel");
    if (p.prof) {
        static long long h[4096 * 8];
        DMM_CUDA(cudaStreamSynchronize(stream));
        DMM_CUDA(cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost));
        double a[8] = {0};
        const unsigned n = grid.x * grid.y;
        for (unsigned i = 0; i < n; ++i)
            for (int j = 0; j < 8; ++j) a[j] += (double)h[i * 8 + j] / n;
        fprintf(stderr,
                "[wgrad] grid %ux%u tiles %lld kpx %d stages %d stage %u B family %d na %d nb %d n_tile %d nsg %d | producer total %.0f "
                "wait empty %.0f | mma total %.0f wait full %.0f | epilogue total %.0f of which waiting for the MMAs %.0f\n",
                grid.x, grid.y, p.total_tiles, p.kpx, p.stages, p.stage_bytes, p.family, p.na, p.num_b, p.n_tile, p.num_sg, a[0], a[1],
                a[2], a[3], a[4], a[5]);
    }
    return 0;
}
