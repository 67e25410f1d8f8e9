// Implicit-GEMM convolution v2 for sm_100a: persistent CTAs, halo patches in shared memory, TMA-store epilogue.
//
//   producer warp : per (source, 64-channel block) ONE TMA load of the halo patch of the CTA's pixel tile
//                   ((TH + kh - 1) x (TW + kw - 1) pixels, each pixel a 128-byte row of the SWIZZLE_128B layout),
//                   per (tap, block) one TMA load of the [n_tile x 64] weight slice (own ring).
//   MMA warp      : every tap is a SHIFTED shared-memory descriptor into the patch (a tap moves the start address by
//                   whole 128-byte pixel rows; the 8-row core-matrix groups are 8 consecutive pixels, the group
//                   stride (SBO) is the patch row pitch), so an A element is fetched from L2 once per tile instead
//                   of once per tap.  One weight stage feeds up to 4 sub-tiles of 128 pixels (msub accumulators).
//                   Accumulators are double-buffered in TMEM: the epilogue of tile i overlaps the MMAs of tile i+1.
//   epilogue warps: tcgen05.ld -> bf16 -> swizzled smem staging -> TMA store (clips partial tiles and the channel
//                   slice); BatchNorm column sums / sums of squares are read back from the staged bf16 tile and kept
//                   in registers across the CTA's tiles.
// Same C-ABI descriptor as v1 (dmm_conv_igemm, include/dmmfods_b200.h); kwidth must be 64.
#include "common.cuh"
#include "../../include/dmmfods_b200.h"

#include <stdlib.h>
#include <string.h>

// per-phase cycle counters of the epilogue (profiles/r02_epilogue_experiments.txt): compiled in only with -DDMM_IGEMM_PHASE_PROF,
// the twelve extra registers make every out_mode-0 instantiation spill
#ifdef DMM_IGEMM_PHASE_PROF
#define DMM_PH(...) __VA_ARGS__
#else
#define DMM_PH(...)
#endif
// wait-time counters of the producer / MMA / epilogue loops: two clock reads per wait are compiled in only with the phase profile
// (a 9-tap convolution paid 20 CS2R per tile inside the MMA issue path)
#ifdef DMM_IGEMM_PHASE_PROF
#define DMM_WAIT_T(ctr, ...) { const long long c0_ = clock64(); __VA_ARGS__; ctr += clock64() - c0_; }
#else
#define DMM_WAIT_T(ctr, ...) { __VA_ARGS__; }
#endif
// "what-if" timing experiments (results are garbage, only the clock counts): -DDMM_IGEMM_WHATIF + env DMM_IGEMM_WHATIF=<mask>
//   1 no A loads  2 no B loads  4 no MMAs  8 no TMA stores  16 no statistics loop  32 no proxy fence in the epilogue
//   64 no conversion / staging writes  128 prologue team does not transform
#ifdef DMM_IGEMM_WHATIF
#define DMM_WI(bit) (p.whatif & (bit))
#else
#define DMM_WI(bit) false
#endif

namespace dmm {

constexpr int kG2Threads = 384;          // warp 0 weight TMA, warp 1 MMA (even tiles), warps 2..9 = two epilogue teams of 4 warps, warp 10 patch
                                         // TMA, warp 11 MMA (odd tiles)
constexpr int kG2ThreadsPro = 512;       // warps 10..13 = BN-ReLU prologue team, warp 14 patch TMA, warp 15 MMA (odd tiles) (PRO instantiations; 128 registers per
                                         // thread: ptxas rounds the budget to 512 threads, so a 448-thread variant gains nothing - measured)
constexpr int kMaxSub = 4;
constexpr int kMaxBStages = 16;          // weight ring slots (resident weights: one slot per (tap group, k-block) of a tile)
constexpr int kG2MaxSmem = 232448;
constexpr int kTailBytes = 1024;         // mbarriers + TMEM address holder behind the rings / staging slots
constexpr uint32_t kStageSlot = 16384;   // one 128-row x 128-byte staging slot
constexpr uint32_t kFoldPitch = 80;      // out_mode 2: fp32 staging row of 16 accumulator columns, padded to 20 words (bank spread)

struct Ig2Params {
    CUtensorMap a_maps[DMM_MAX_SRC];
    CUtensorMap b_map;
    CUtensorMap o_map;
    CUtensorMap x_map;             // bnb: raw input of the BatchNorm whose backward reduce is fused into the epilogue
    int num_src;
    int src_nblk[DMM_MAX_SRC];
    int src_lastk[DMM_MAX_SRC];
    int src_ox[DMM_MAX_SRC], src_oy[DMM_MAX_SRC];
    int src_tap0[DMM_MAX_SRC + 1];
    uint32_t src_sbo[DMM_MAX_SRC];
    uint32_t src_tx[DMM_MAX_SRC];
    uint32_t sub_aoff[DMM_MAX_SRC][kMaxSub];
    uint32_t tap_aoff[DMM_MAX_TAPS];
    int tap_kb0[DMM_MAX_TAPS];
    int sub_x[kMaxSub], sub_y[kMaxSub];
    int msub, sub_w, sub_h;
    int W, H, B, TW, TH, tiles_x, tiles_y, tiles_n;
    FastDiv fd_n, fd_x, fd_y;
    long long total_tiles;
    int n_tile, N;
    int sa, sb;
    uint32_t a_stage, b_stage, b_tap, tmem_cols;   // b_stage = tps * b_tap
    int mma2;                                       // two MMA issuer warps (alternate tiles) / one
    int last_src;                                   // last source with taps
    int a_adv, b_adv;                               // ring stages one tile consumes, modulo the ring depth ...
    uint32_t a_flip, b_flip;                        // ... and the parity of its full wraps (the two MMA warps skip each other's tiles)
    int w_res;                                      // weights RESIDENT: the CTA's whole weight slice is loaded once (sb = stages per tile)
    int tps;                                        // taps per weight stage
    int tpk;                                        // taps per 64-wide K block of the weights: 1, 2 (32-channel sources) or 4 (16-channel sources)
    int tpk_log;                                    // log2(tpk)
    int Wv, Hv;
    int x_step, x_org;             // tile column origin = tx * x_step + x_org (out_mode 2: tiles overlap by fold_kw - 1 columns)
    int nsx;                       // sub-tiles per tile row
    int fold_kw, fold_c;           // out_mode 2: kernel width folded into N, classes (<= 4)
    int tw_log, sw_log, sh_log;    // out_mode 2: log2 of TW, sub_w, sub_h
    long long* prof;     // debug (DMM_IGEMM_PROF=1): per-CTA cycle counters
    int whatif;
    float* out32;
    int OH, OW, out_sy, out_sx, out_py, out_px;
    double* stats;
    int stats_ld, stats_off;
    int pro;                       // BN-ReLU prologue on the A tiles of source 0 (1x1 convolutions)
    int pro_kp;                    // padded channel count of source 0 (multiple of 64)
    int pro_c;                     // valid channels of source 0
    int pro_pw;                    // > 0: K x K prologue - patch width in pixels; out-of-image patch pixels are forced to zero
    dmm_bn_t pro_bn;
    const float* epi_bias;         // out_mode 0: out = act(acc + epi_bias[n]) (a folded eval-mode BatchNorm's shift), or NULL
    int epi_relu;
    int bnb;
    int bnb_np;                    // bnb: channels of the coefficient table in shared memory (tiles_n * n_tile, + 2)
    const float* bnb_gamma;
    const float* bnb_beta;
    const float* bnb_mean;
    const float* bnb_invstd;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
            reinterpret_cast<uint64_t>(m)),
        "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// packed fp32 pairs (sm_100 add.f32x2 / fma.rn.f32x2): the statistics of two neighbouring channels per instruction
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// bf16x2 word -> (fp32 of the low half, fp32 of the high half)
__device__ __forceinline__ uint64_t bf16x2_to_f2(uint32_t u) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(u << 16), "r"(u & 0xffff0000u));
    return d;
}
__device__ __forceinline__ float f2_lo(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
// Statistics of one staged [128 pixels x 64 channels] bf16 chunk, vectorised: thread (j = r & 7, g = r >> 3) owns the 16-byte chunk
// j (channels 8j .. 8j+7) of rows g, g + 16, ..., g + 112 (8 x ld.shared.v4, the swizzled chunk position is the same for all of them),
// and keeps sum / sum of squares of its 8 channels as 8 packed fp32 pairs ACROSS all chunks of the CTA (acc[0..3] sums, acc[4..7] squares).
__device__ __forceinline__ void stats_chunk_v(uint32_t sbase, uint64_t* acc) {
    uint4 w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = lds_v4(sbase + i * 2048);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t u[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t v = bf16x2_to_f2(u[k]);
            acc[k] = f2_add(acc[k], v);
            acc[4 + k] = f2_fma(v, v, acc[4 + k]);
        }
    }
}
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
    return d;
}
// Fused BatchNorm-ReLU backward reduce over one staged chunk, same thread mapping as stats_chunk_v: g = the staged output gradient,
// x = the raw BatchNorm input of the same pixels / channels (second staging tile); dz = g * [scale * x + shift > 0];
// acc[0..3] += dz, acc[4..7] += dz * (x - mean) for the thread's 8 channels.  coef: 8 channels x (scale, shift, -mean) as 12 pairs.
__device__ __forceinline__ void bnb_chunk_v(uint32_t gbase, uint32_t xbase, const uint64_t* coef, uint64_t* acc) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {          // two batches of 4 rows: 8 x ld.shared.v4 in flight
        uint4 gw[4], xw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            gw[i] = lds_v4(gbase + (4 * h + i) * 2048);
            xw[i] = lds_v4(xbase + (4 * h + i) * 2048);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t gu[4] = {gw[i].x, gw[i].y, gw[i].z, gw[i].w};
            const uint32_t xu[4] = {xw[i].x, xw[i].y, xw[i].z, xw[i].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t xv = bf16x2_to_f2(xu[k]);
                const uint64_t t = f2_fma(xv, coef[k], coef[4 + k]);
                const float a0 = f2_lo(t) > 0.f ? __uint_as_float(gu[k] << 16) : 0.f;
                const float a1 = f2_hi(t) > 0.f ? __uint_as_float(gu[k] & 0xffff0000u) : 0.f;
                const uint64_t av = f2_pack(a0, a1);
                acc[k] = f2_add(acc[k], av);
                acc[4 + k] = f2_fma(av, f2_add(xv, coef[8 + k]), acc[4 + k]);
            }
        }
    }
}
__device__ __forceinline__ void epi_bar(int team) { asm volatile("bar.sync %0, 128;" ::"r"(team + 1) : "memory"); }

// upper 32 bits of a K-major SWIZZLE_128B shared-memory descriptor (SBO, descriptor version 1, layout type 2); the
// lower 32 bits are (address >> 4) | (LBO >> 4) << 16.
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29); }
struct TileCoord {
    int b, x0, y0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const Ig2Params& p, long long tile) {
    TileCoord c;
    uint32_t nt, tx, ty;
    uint32_t t = fast_divmod((uint32_t)tile, p.fd_n, nt);      // total_tiles < 2^31 (checked by the launcher)
    t = fast_divmod(t, p.fd_x, tx);
    c.b = (int)fast_divmod(t, p.fd_y, ty);
    c.x0 = (int)tx * p.x_step + p.x_org;
    c.y0 = (int)ty * p.TH;
    c.n0 = (int)nt * p.n_tile;
    return c;
}

// VAR: 0 plain, 1 BN-ReLU prologue on the A tiles (PRO), 2 fused BatchNorm backward reduce in the epilogue (BNB; out_mode 0 only).
// Separate instantiations: the statistics code of BNB (x-tile cursor, coefficient registers) made the plain kernels spill.
template <int NCH, int OUT_MODE, int MSUB, int VAR>
__global__ void __launch_bounds__(VAR == 1 ? kG2ThreadsPro : kG2Threads, 1) igemm2_kernel(const __grid_constant__ Ig2Params p) {
    constexpr bool PRO = VAR == 1;
    // OUT_MODE 3 / 4 (internal): 3 kernel columns folded into N = 3 * FC for FC = 32 / 64 output channels (C-ABI out_mode 3)
    constexpr bool FOLD3 = OUT_MODE == 3 || OUT_MODE == 4;
    constexpr int FC = OUT_MODE == 4 ? 64 : 32;
    // fused BN backward reduce: out_mode 0 without the prologue (the host rejects other combinations); a compile-time false keeps
    // its code - x-tile cursor, second statistics loop - out of the register-starved prologue instantiations
    constexpr bool bnb_on = OUT_MODE == 0 && VAR == 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    uint8_t* b_ring = a_ring + (size_t)p.sa * p.a_stage;
    uint8_t* stg = b_ring + (size_t)p.sb * p.b_stage;
    uint8_t* xstg = stg + ((OUT_MODE == 0 || FOLD3) ? 2 * kStageSlot : (OUT_MODE == 2 ? MSUB * 128 * kFoldPitch : 0));   // bnb: two x tiles per team
    uint8_t* tail = xstg + ((OUT_MODE == 0 && bnb_on) ? 4 * kStageSlot : 0);
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* a_empty = a_full + 8;
    uint64_t* b_full = a_empty + 8;
    uint64_t* b_empty = b_full + kMaxBStages;
    uint64_t* acc_full = b_empty + kMaxBStages;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* x_bar = acc_empty + 2;
    uint64_t* a_ready = x_bar + 2;                                       // pro: A stage transformed (4 warp arrivals); bnb (never
                                                                         // together with pro): a_ready[0..1] are x_bar[2..3]
    uint64_t* turn = a_ready + 8;                                        // hand-over between the two MMA warps (see there)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(turn + 2);
    float* pcoef = reinterpret_cast<float*>(tail + kTailBytes);                 // pro: [2][pro_kp] scale / shift

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.sa; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < p.sb; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < p.sa; ++s) mbar_init(&a_ready[s], 4);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 8);
            mbar_init(&x_bar[s], 1);
        }
        if (bnb_on) {
            mbar_init(&x_bar[2], 1);
            mbar_init(&x_bar[3], 1);
        }
        mbar_init(&turn[0], 1);
        mbar_init(&turn[1], 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, p.tmem_cols);
        tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.b_map);
        if (p.src_tap0[1] > 0) tma_prefetch_desc(&p.a_maps[0]);
        if (OUT_MODE == 0 || FOLD3) tma_prefetch_desc(&p.o_map);
        if (OUT_MODE == 0 && bnb_on) tma_prefetch_desc(&p.x_map);
    }
    // everything above is independent of earlier kernels: wait for them (programmatic dependent launch) only here
    pdl_prologue();
    if (PRO) {
        // per-channel scale / shift of the prologue BatchNorm (batch statistics of the raw input, or running statistics);
        // CTA 0 also records mean / invstd for backward and updates the running statistics (nn.BatchNorm2d semantics)
        const int C = p.pro_kp;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float sc = 0.f, sh = 0.f;
            if (c < p.pro_c) {
                const BnCoef k = bn_coef_fwd(p.pro_bn, c, blockIdx.x == 0);
                sc = k.scale;
                sh = k.shift;
            }
            pcoef[c] = sc;
            pcoef[C + c] = sh;
        }
    }
    if (bnb_on) {
        // fused BatchNorm backward reduce: (scale, shift, -mean) of all N output channels, [3][bnb_np] floats; per-chunk global
        // loads of these cost 2 100 cycles per chunk (phase counters); channels beyond N get zeros (dz = 0)
        const int NP = p.bnb_np;
        for (int c = threadIdx.x; c < NP; c += blockDim.x) {
            float sc = 0.f, sh = 0.f, nm = 0.f;
            if (c < p.N) {
                const float mu = __ldg(p.bnb_mean + c);
                sc = (p.bnb_gamma ? __ldg(p.bnb_gamma + c) : 1.f) * __ldg(p.bnb_invstd + c);
                sh = (p.bnb_beta ? __ldg(p.bnb_beta + c) : 0.f) - mu * sc;
                nm = -mu;
            }
            pcoef[c] = sc;
            pcoef[NP + c] = sh;
            pcoef[2 * NP + c] = nm;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ================= TMA producer =================
        // ================= weight (B) TMA producer; the halo patches (A) have their own producer warp so that the next
        // patch is requested the moment its stage frees up instead of after the current block's weight stages =================
        if (lane == 0) {
            int bst = 0;
            uint32_t bph = 0;
            long long w_a = 0, w_b = 0;
            const long long t_begin = clock64();
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                // resident weights: every tile of this CTA uses the same weight slice (same n0) - it is loaded once, each
                // (tap group, k-block) into its own ring slot, and never released
                if (p.w_res && tile != blockIdx.x) break;
                const TileCoord tc = decode_tile(p, tile);
                for (int s = 0; s < p.num_src; ++s) {
                    const int nblk = p.src_nblk[s];
                    const int t0 = p.src_tap0[s], t1 = p.src_tap0[s + 1];
                    if (t0 == t1) continue;
                    for (int cb = 0; cb < nblk; ++cb) {
                        for (int t = t0; t < t1; t += p.tps) {
                            const int nt = (t1 - t) < p.tps ? (t1 - t) : p.tps;
                            DMM_WAIT_T(w_b, if (!p.w_res) mbar_wait(&b_empty[bst], bph ^ 1));
                            const int nblk_b = (nt + p.tpk - 1) >> p.tpk_log;
                            if (DMM_WI(2)) mbar_arrive(&b_full[bst]);
                            else mbar_arrive_expect_tx(&b_full[bst], (uint32_t)nblk_b * p.b_tap);
                            uint8_t* bdst = b_ring + (size_t)bst * p.b_stage;
                            for (int j = 0; j < nblk_b && !DMM_WI(2); ++j)
                                tma_load_2d(bdst + (size_t)j * p.b_tap, &p.b_map, &b_full[bst], (p.tap_kb0[t + j * p.tpk] + cb) * 64, tc.n0);
                            if (++bst == p.sb) { bst = 0; bph ^= 1; }
                        }
                    }
                }
            }
            if (p.prof) {
                p.prof[blockIdx.x * 16 + 0] = clock64() - t_begin;
                p.prof[blockIdx.x * 16 + 1] = w_a;
                p.prof[blockIdx.x * 16 + 2] = w_b;
            }
        }
    } else if (warp == (PRO ? 14 : 10)) {
        // ================= halo-patch (A) TMA producer =================
        if (lane == 0) {
            int ast = 0;
            uint32_t aph = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode_tile(p, tile);
                for (int s = 0; s < p.num_src; ++s) {
                    if (p.src_tap0[s] == p.src_tap0[s + 1]) continue;
                    const int nblk = p.src_nblk[s];
                    for (int cb = 0; cb < nblk; ++cb) {
                        mbar_wait(&a_empty[ast], aph ^ 1);
                        if (DMM_WI(1)) mbar_arrive(&a_full[ast]);
                        else {
                        mbar_arrive_expect_tx(&a_full[ast], p.src_tx[s]);
                        tma_load_4d(a_ring + (size_t)ast * p.a_stage, &p.a_maps[s], &a_full[ast], cb * 64, tc.x0 + p.src_ox[s],
                                    tc.y0 + p.src_oy[s], tc.b);
                        }
                        if (++ast == p.sa) { ast = 0; aph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == (PRO ? 15 : 11)) {
        // ================= MMA issuers =================
        // TWO issuing warps take the CTA's tiles alternately (warp 1 the even ones = accumulator stage 0, the other the odd ones):
        // the ~2 500 cycles one warp spends per tile on barrier waits, descriptor set-up and commits (measured with the phase
        // counters: more than the tensor time of a 3-tap N = 96 tile) now overlap with the other warp's MMAs.  Ring stages are
        // consumed in tile order whoever waits for them; a commit only tracks the MMAs of the thread that issued it.
        const uint32_t mma_id = warp == 1 ? 0u : 1u;
        const uint32_t nmma = p.mma2 ? 2u : 1u;           // mma2 == 0: warp 1 issues every tile, the second warp idles
        // The issue block is guarded by elect.sync: nvcc then emits straight-line UTCHMMA (a `lane == 0` guard makes it
        // wrap every MMA in an ELECT/BRA.U.ANY loop, ~45 instead of <40 cycles per instruction).  The tensor core needs
        // (4096 + 32 N) / 128 cycles per M=128, K=16 instruction (operands stream from shared memory at 128 B/cycle,
        // measured: scripts/ubench/mma_rate.cu), so the issue stream has to stay below ~40 cycles per MMA.
        const uint32_t idesc = make_idesc_bf16(128, p.n_tile, 0, 0);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t a_ring_u = smem_u32(a_ring), b_ring_u = smem_u32(b_ring);
        const uint64_t bhi = (uint64_t)desc_hi(1024u) << 32;
        // ONE elected thread runs the whole issue loop (waits included): no per-stage elect / reconvergence / warp sync, whose
        // latencies added up to ~2 000 cycles per 9-tap tile next to 2 600 cycles of tensor time (phase counters)
        if (elect_one()) {
        int ast = 0, bst = 0;
        uint32_t aph = 0, bph = 0;
        uint32_t it = mma_id;
        // ring positions after the tile the OTHER warp handles (host: stages per tile modulo the ring depth, parity of the wraps)
        auto skip_tile = [&]() {
            ast += p.a_adv; aph ^= p.a_flip;
            if (ast >= p.sa) { ast -= p.sa; aph ^= 1; }
            bst += p.b_adv; bph ^= p.b_flip;
            if (bst >= p.sb) { bst -= p.sb; bph ^= 1; }
        };
        if (mma_id) skip_tile();
        if (mma_id >= nmma) it = 0x7fffffffu;             // no tiles for this warp
        long long w_a = 0, w_b = 0, w_acc = 0;
        DMM_PH(long long ph_taps = 0, ph_hdr = 0;)      // cycles inside the elected issue block / between a k-block's operand wait and it
        const long long t_begin = clock64();
        for (long long tile = mma_id >= nmma ? p.total_tiles : blockIdx.x + (long long)mma_id * gridDim.x; tile < p.total_tiles;
             tile += (long long)nmma * gridDim.x, it += nmma) {
            const uint32_t as = it & 1, accph = (it >> 1) & 1;
            DMM_WAIT_T(w_acc, mbar_wait(&acc_empty[as], accph ^ 1));
            // A ring barrier is waited for by PARITY: a warp may only start waiting for the stages of its tile once the other warp
            // has completed all operand waits of the tile before (otherwise "use k" and "use k - 2" of a stage are the same parity).
            // turn[w] is arrived on by warp w after the last operand wait of each of its tiles; the two warps alternate strictly.
            if (it > 0 && nmma == 2) mbar_wait(&turn[mma_id ^ 1], ((it - 1) >> 1) & 1);
            tc_fence_after();
            const uint32_t d0 = tmem_u + as * MSUB * p.n_tile;
            uint32_t acc0 = 0;      // 0 only for the first MMA of every accumulator of this tile
            for (int s = 0; s < p.num_src; ++s) {
                const int nblk = p.src_nblk[s];
                const int t0 = p.src_tap0[s], t1 = p.src_tap0[s + 1];
                if (t0 == t1) continue;
                const uint64_t ahi = (uint64_t)desc_hi(p.src_sbo[s]) << 32;
                uint32_t so[MSUB];
#pragma unroll
                for (int i = 0; i < MSUB; ++i) so[i] = p.sub_aoff[s][i] >> 4;
                for (int cb = 0; cb < nblk; ++cb) {
                    DMM_WAIT_T(w_a, mbar_wait(PRO ? &a_ready[ast] : &a_full[ast], aph));
                    DMM_PH(const long long h0 = clock64();)
                    const uint32_t a_lo = ((a_ring_u + (uint32_t)ast * p.a_stage) >> 4) | (1u << 16);
                    const int ksteps = (cb == nblk - 1) ? p.src_lastk[s] : 4;
                    // resident weights, after the CTA's first tile: nothing to wait for, and the (tap group, k-block) slots of one
                    // k-block are consecutive in the ring - all taps are issued from ONE elected block (no per-tap loop overhead)
                    const bool fused = p.w_res && it >= nmma;   // it < nmma: this warp's first tile (it waits for the weight stages once)
                    const int tstep = fused ? (t1 - t0) : p.tps;
                    for (int t = t0; t < t1; t += tstep) {
                        const int nt = (t1 - t) < tstep ? (t1 - t) : tstep;
                        if (!fused) DMM_WAIT_T(w_b, mbar_wait(&b_full[bst], bph));
                        if (s == p.last_src && cb == nblk - 1 && t + nt == t1) mbar_arrive(&turn[mma_id]);      // last operand wait of this tile
                        tc_fence_after();
                        {
                            DMM_PH(const long long h1 = clock64(); if (t == t0) ph_hdr += h1 - h0;)
                            const uint32_t b_lo0 = ((b_ring_u + (uint32_t)bst * p.b_stage) >> 4) | (1u << 16);
                            uint32_t aoff = p.tap_aoff[t] >> 4;
                            const uint32_t tpk_mask = (uint32_t)p.tpk - 1u, tap_sub = 8u >> p.tpk_log, b_tap16 = p.b_tap >> 4;
                            for (int j = 0; j < nt; ++j) {
                                const uint32_t a_tap = a_lo + aoff;
                                // weights of tap j: own 64-wide block, or (16 / 32-channel sources) a quarter / half of a shared block
                                const uint32_t b_lo = b_lo0 + ((uint32_t)j >> p.tpk_log) * b_tap16 + ((uint32_t)j & tpk_mask) * tap_sub;
                                if (j + 1 < nt) aoff = p.tap_aoff[t + j + 1] >> 4;      // prefetch the next tap's offset
                                // fully unrolled per k-step count: keeps the descriptor arithmetic in uniform registers
#define DMM_ISSUE_TAP(KS)                                                                                                       \
    _Pragma("unroll") for (int k = 0; k < KS; ++k) _Pragma("unroll") for (int sub = 0; sub < MSUB; ++sub)                      \
        umma_bf16(d0 + sub * p.n_tile, ahi | (a_tap + so[sub] + 2 * k), bhi | (b_lo + 2 * k), idesc, k == 0 ? acc0 : 1u)
                                if (DMM_WI(4)) {}
                                else if (ksteps == 4) { DMM_ISSUE_TAP(4); }
                                else if (ksteps == 1) { DMM_ISSUE_TAP(1); }
                                else if (ksteps == 2) { DMM_ISSUE_TAP(2); }
                                else { DMM_ISSUE_TAP(3); }
#undef DMM_ISSUE_TAP
                                acc0 = 1;
                            }
                            DMM_PH(ph_taps += clock64() - h1;)
                            if (!p.w_res) umma_commit(&b_empty[bst]);
                            if (t + nt == t1) umma_commit(&a_empty[ast]);
                        }
                        acc0 = 1;
                        bst += fused ? ((nt + p.tpk - 1) >> p.tpk_log) : 1;
                        if (bst >= p.sb) { bst -= p.sb; bph ^= 1; }
                    }
                    if (++ast == p.sa) { ast = 0; aph ^= 1; }
                }
            }
            umma_commit(&acc_full[as]);
            if (nmma == 2) skip_tile();
        }
        if (p.prof && mma_id == 0) {
            p.prof[blockIdx.x * 16 + 4] = clock64() - t_begin;
            p.prof[blockIdx.x * 16 + 5] = w_a;
            p.prof[blockIdx.x * 16 + 6] = w_b;
            p.prof[blockIdx.x * 16 + 7] = w_acc;
        }
        DMM_PH(if (p.prof && mma_id == 0) { p.prof[blockIdx.x * 16 + 1] = ph_taps; p.prof[blockIdx.x * 16 + 3] = ph_hdr; })
        }   // elected thread
    } else {
        // ================= epilogue: two teams of 4 warps, alternating 64-column chunks =================
        const int team = (warp - 2) >> 2;
        if (PRO && team == 2) {
            // ================= prologue team: relu(bn(x)) in place on every A stage =================
            // thread e owns the logical 16-byte chunk (8 channels) e & 7 of rows e >> 3, e >> 3 + 16, ...; the physical chunk of a
            // 128-byte swizzled row r is (chunk ^ (r & 7)).  Zero-filled (out-of-image / beyond-C) elements may become non-zero:
            // out-of-image pixels never reach memory (TMA store clipping, masked statistics) and channels beyond C have zero
            // coefficients.
            const int e = (warp - 10) * 32 + lane;
            const int j = e & 7;
            const int rows = (int)(p.src_tx[0] >> 7);
            int ast = 0;
            uint32_t aph = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nblk = p.src_nblk[0];
                const TileCoord tcp = decode_tile(p, tile);
                const int px_org = tcp.x0 + p.src_ox[0], py_org = tcp.y0 + p.src_oy[0];
                for (int cb = 0; cb < nblk; ++cb) {
                    float sc[8], sh[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        sc[i] = pcoef[cb * 64 + j * 8 + i];
                        sh[i] = pcoef[p.pro_kp + cb * 64 + j * 8 + i];
                    }
                    mbar_wait(&a_full[ast], aph);
                    const uint32_t base = smem_u32(a_ring + (size_t)ast * p.a_stage);
                    if (p.pro_pw > 0) {
                        // K x K: the zero padding applies to the ACTIVATED tensor, so patch pixels outside the image (TMA
                        // zero-filled raw values) must stay zero instead of becoming relu(shift)
                        int pr = 0, pc = e >> 3;
                        while (pc >= p.pro_pw) { pc -= p.pro_pw; ++pr; }
                        for (int r = e >> 3; r < rows; r += 16) {
                            const uint32_t ptr = base + r * 128 + ((j ^ (r & 7)) << 4);
                            const int yy = py_org + pr, xx = px_org + pc;
                            uint4 o = make_uint4(0, 0, 0, 0);
                            if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
                                const uint4 v = lds_v4(ptr);
                                uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float lo = fmaxf(fmaf(bf16_lo(w[i]), sc[2 * i], sh[2 * i]), 0.f);
                                    const float hi = fmaxf(fmaf(bf16_hi(w[i]), sc[2 * i + 1], sh[2 * i + 1]), 0.f);
                                    w[i] = pack_bf16x2(lo, hi);
                                }
                                o = make_uint4(w[0], w[1], w[2], w[3]);
                            }
                            sts_v4(ptr, o);
                            pc += 16;
                            while (pc >= p.pro_pw) { pc -= p.pro_pw; ++pr; }
                        }
                    } else if (!DMM_WI(128)) {
#pragma unroll 4
                    for (int r = e >> 3; r < rows; r += 16) {
                        const uint32_t ptr = base + r * 128 + ((j ^ (r & 7)) << 4);
                        const uint4 v = lds_v4(ptr);
                        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float lo = fmaxf(fmaf(bf16_lo(w[i]), sc[2 * i], sh[2 * i]), 0.f);
                            const float hi = fmaxf(fmaf(bf16_hi(w[i]), sc[2 * i + 1], sh[2 * i + 1]), 0.f);
                            w[i] = pack_bf16x2(lo, hi);
                        }
                        sts_v4(ptr, make_uint4(w[0], w[1], w[2], w[3]));
                    }
                    }
                    fence_proxy_async();                  // generic-proxy writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_ready[ast]);
                    if (++ast == p.sa) { ast = 0; aph ^= 1; }
                }
            }
        } else {
        const int q = warp & 3;              // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;         // accumulator row = pixel within the sub-tile
        const int px = r % p.sub_w, py = r / p.sub_w;
        const bool do_stats = (OUT_MODE == 0 || FOLD3) && (p.stats != nullptr);
        const int cp = r & 31, rq = r >> 5;  // statistics: column pair / row quarter of the staged chunk
        uint8_t* slot = stg + team * kStageSlot;
        if (FOLD3) {
            // compact staging: 4 image rows x (32 - 2) valid pixels = 120 rows; rows 120..127 stay zero for the statistics loop
            if (r >= 120) {
#pragma unroll
                for (int c16 = 0; c16 < 8; ++c16) sts_v4(smem_u32(slot) + r * 128 + c16 * 16, make_uint4(0, 0, 0, 0));
            }
            epi_bar(team);
        }
        uint8_t* srow = slot + r * 128;
        // bnb: the x tile of a chunk is loaded ONE CHUNK AHEAD into the team's other x buffer (a load issued in the chunk itself
        // exposed the whole DRAM latency in front of the statistics loop: 3 600 cycles per chunk, measured).  Thread r == 0 walks a
        // second cursor over the team's chunks; xi = index of the team's current chunk: buffer xi & 1, barrier parity (xi >> 1) & 1.
        uint8_t* xslot = xstg + team * 2 * kStageSlot;
        uint32_t xi = 0;
        const uint32_t slot_u = smem_u32(slot), srow_u = smem_u32(srow), xslot_u = smem_u32(xslot);
        long long xq_tile = blockIdx.x;
        int xq_sub = 0, xq_c = -1;
        uint32_t xq_ctr = 0, xq_n = 0;
        TileCoord xq_tc = decode_tile(p, xq_tile < p.total_tiles ? xq_tile : 0);
        auto x_prefetch_next = [&]() {       // advance the cursor to the team's next chunk and load its x tile (thread r == 0)
            if (xq_tile >= p.total_tiles) return;
            for (;;) {
                ++xq_c;
                if (xq_c >= NCH || xq_tc.n0 + 64 * xq_c >= p.N) { xq_c = 0; ++xq_sub; }
                if (xq_sub >= MSUB) {
                    xq_sub = 0;
                    xq_tile += gridDim.x;
                    if (xq_tile >= p.total_tiles) return;
                    xq_tc = decode_tile(p, xq_tile);
                }
                if (((xq_ctr++) & 1) == (uint32_t)team) break;
            }
            uint64_t* bar = &x_bar[team * 2 + (xq_n & 1)];
            mbar_arrive_expect_tx(bar, kStageSlot);
            tma_load_4d(xslot + (xq_n & 1) * kStageSlot, &p.x_map, bar, xq_tc.n0 + 64 * xq_c, xq_tc.x0 + p.sub_x[xq_sub], xq_tc.y0 + p.sub_y[xq_sub],
                        xq_tc.b);
            ++xq_n;
        };
        if (OUT_MODE == 0 && bnb_on && r == 0) x_prefetch_next();
        // VSTATS (n_tile <= 128, no fused BN backward): vectorised statistics, accumulators = 8 packed fp32 pairs per 64-channel block;
        // otherwise 4 doubles per block (column pair cp of row quarter rq), stored in the same registers
        constexpr bool VS = NCH <= 2;
        const bool vstats = VS;                // NCH <= 2: both statistics kinds are vectorised
        const uint32_t sbase = slot_u + (uint32_t)(r >> 3) * 128u + (uint32_t)(((r & 7) ^ ((r >> 3) & 7)) << 4);
        uint64_t sraw[NCH][VS ? 8 : 4];
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int j = 0; j < (VS ? 8 : 4); ++j) sraw[c][j] = 0ull;
#define sacc_add(c, j, x) sraw[c][j] = (uint64_t)__double_as_longlong(__longlong_as_double((long long)sraw[c][j]) + (double)(x))
#define sacc_get(c, j) __longlong_as_double((long long)sraw[c][j])
        // flush the accumulators of this thread into the global statistics slots (channel block c starts at column n0 + 64 c)
        auto flush_stats = [&](int n0) {
            double* st = p.stats + (size_t)(blockIdx.x % DMM_STATS_SLOTS) * 2 * p.stats_ld + p.stats_off;
            const int lim_l = FOLD3 ? p.fold_c : p.n_tile, lim_g = FOLD3 ? p.fold_c : p.N;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (vstats) {
                    if constexpr (VS) {
                        // the 4 lanes of a warp that share j (lane bits 3, 4 = row groups) are summed first: 16 atomics in 8 lanes
                        float v[16];
#pragma unroll
                        for (int k = 0; k < 8; ++k) { v[2 * k] = f2_lo(sraw[c][k]); v[2 * k + 1] = f2_hi(sraw[c][k]); }
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            v[k] += __shfl_xor_sync(0xffffffffu, v[k], 8);
                            v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
                        }
                        if (lane < 8) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int cl = c * 64 + 8 * lane + e, col = n0 + cl;
                                if (cl < lim_l && col < lim_g) {
                                    atomicAdd(st + col, (double)v[e]);
                                    atomicAdd(st + p.stats_ld + col, bnb_on ? (double)v[8 + e] * (double)p.bnb_invstd[col] : (double)v[8 + e]);
                                }
                            }
                        }
                    }
                } else {
                    const int col = n0 + c * 64 + 2 * cp;
                    if (c * 64 + 2 * cp < lim_l && col < lim_g) {
                        atomicAdd(st + col, sacc_get(c, 0));
                        atomicAdd(st + p.stats_ld + col, bnb_on ? sacc_get(c, 2) * (double)p.bnb_invstd[col] : sacc_get(c, 2));
                        if (col + 1 < lim_g) {
                            atomicAdd(st + col + 1, sacc_get(c, 1));
                            atomicAdd(st + p.stats_ld + col + 1, bnb_on ? sacc_get(c, 3) * (double)p.bnb_invstd[col + 1] : sacc_get(c, 3));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < (VS ? 8 : 4); ++j) sraw[c][j] = 0ull;
            }
        };
        uint32_t chunk_ctr = 0;
        uint32_t it = 0;
        int last_n0 = 0;
        long long w_full = 0;
        DMM_PH(long long ph_store_wait = 0, ph_tmem = 0, ph_bar1 = 0, ph_pack = 0, ph_bar2 = 0, ph_stats = 0;)
        DMM_PH(long long ph_xpre = 0, ph_coef = 0, ph_xwait = 0, ph_loop = 0;)      // BNB variant: x prefetch, coefficient loads, x wait, loop      // phase cycles (thread r == 0)
        const long long t_begin = clock64();
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const TileCoord tc = decode_tile(p, tile);
            const uint32_t as = it & 1, accph = (it >> 1) & 1;
            last_n0 = tc.n0;
            DMM_WAIT_T(w_full, mbar_wait(&acc_full[as], accph));
            tc_fence_after();
            if (OUT_MODE == 2) {
                // kernel columns folded into N: stage the fp32 accumulators of the whole tile, then every output pixel sums its
                // fold_kw horizontal neighbours (column kw*C + n of the pixel kw - fold_kw/2 to its right)
                const uint32_t stg_u = smem_u32(stg);
                asm volatile("bar.sync 3, 256;" ::: "memory");       // the previous tile has been read out
                for (int sub = team; sub < MSUB; sub += 2) {
                    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (as * MSUB + sub) * p.n_tile;
                    uint32_t v[16];
                    tmem_ld16(trow, v);
                    tmem_ld_wait();
                    const uint32_t dst = stg_u + (uint32_t)(sub * 128 + r) * kFoldPitch;
#pragma unroll
                    for (int g = 0; g < 4; ++g) sts_v4(dst + 16 * g, make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
                asm volatile("bar.sync 3, 256;" ::: "memory");
                // thread e owns tile column e % TW (TW, sub_w, sub_h are powers of two: shifts only) and walks down the rows
                const int e = (warp - 2) * 32 + lane;
                const int K = p.fold_kw, C = p.fold_c, TWo = p.TW - (K - 1);
                const int xo = e & (p.TW - 1);
                const int x = tc.x0 + (K >> 1) + xo;
                const long long plane = (long long)p.OH * p.OW;
                if (xo < TWo && x < p.W) {
                    for (int yy = e >> p.tw_log; yy < p.TH; yy += 256 >> p.tw_log) {
                        const int y = tc.y0 + yy;
                        if (y >= p.H) break;
                        const int base = (((yy >> p.sh_log) * p.nsx) << 7) + ((yy & (p.sub_h - 1)) << p.sw_log);
                        float acc[4] = {0.f, 0.f, 0.f, 0.f};
                        for (int kw = 0; kw < K; ++kw) {
                            const int xt = xo + kw;
                            const uint32_t src = stg_u + (uint32_t)(base + ((xt >> p.sw_log) << 7) + (xt & (p.sub_w - 1))) * kFoldPitch +
                                                 (uint32_t)(kw * C) * 4u;
#pragma unroll
                            for (int n = 0; n < 4; ++n)
                                if (n < C) acc[n] += __uint_as_float(lds_u32(src + 4u * n));
                        }
                        float* o = p.out32 + (long long)tc.b * C * plane + (long long)y * p.OW + x;
#pragma unroll
                        for (int n = 0; n < 4; ++n)
                            if (n < C) o[n * plane] = acc[n];
                    }
                }
                continue;
            }
            for (int sub = 0; sub < MSUB; ++sub) {
                const int x = tc.x0 + p.sub_x[sub] + px, y = tc.y0 + p.sub_y[sub] + py;
                const bool valid = (x < p.Wv) && (y < p.Hv);
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (as * MSUB + sub) * p.n_tile;
                if (FOLD3) {
                    // 3 kernel columns folded into N = 3 * FC: lane = pixel of one 32-pixel image row (sub_w == 32), output pixel
                    // px = column block 0 of lane px-1 + block 1 of lane px + block 2 of lane px+1 (warp shuffles); lanes 0 and 31
                    // are the halo.  bf16 result -> compact staging rows (4 x 30) -> TMA store + statistics like out_mode 0.
                    if (((chunk_ctr++) & 1) != (uint32_t)team) continue;
                    float o[FC];
                    {
                        uint32_t v[16];
#pragma unroll
                        for (int g = 0; g < FC / 16; ++g) {
                            tmem_ld16(trow + FC + 16 * g, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[16 * g + i] = __uint_as_float(v[i]);
                        }
#pragma unroll
                        for (int g = 0; g < FC / 16; ++g) {
                            tmem_ld16(trow + 16 * g, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[16 * g + i] += __shfl_up_sync(0xffffffffu, __uint_as_float(v[i]), 1);
                        }
#pragma unroll
                        for (int g = 0; g < FC / 16; ++g) {
                            tmem_ld16(trow + 2 * FC + 16 * g, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[16 * g + i] += __shfl_down_sync(0xffffffffu, __uint_as_float(v[i]), 1);
                        }
                    }
                    if (r == 0) bulk_wait_read0();
                    epi_bar(team);
                    if (lane >= 1 && lane <= 30) {
                        const int rc = q * 30 + lane - 1;                       // compact staging row
                        const uint32_t srow3 = slot_u + rc * 128;
#pragma unroll
                        for (int g = 0; g < FC / 8; ++g) {
                            uint4 w = make_uint4(0, 0, 0, 0);
                            if (valid && x >= 0) {
                                w.x = pack_bf16x2(o[8 * g + 0], o[8 * g + 1]);
                                w.y = pack_bf16x2(o[8 * g + 2], o[8 * g + 3]);
                                w.z = pack_bf16x2(o[8 * g + 4], o[8 * g + 5]);
                                w.w = pack_bf16x2(o[8 * g + 6], o[8 * g + 7]);
                            }
                            sts_v4(srow3 + ((g ^ (rc & 7)) << 4), w);
                        }
                    }
                    fence_proxy_async();
                    epi_bar(team);
                    if (r == 0) {
                        tma_store_4d(&p.o_map, slot, 0, tc.x0 + p.sub_x[sub] + 1, tc.y0 + p.sub_y[sub], tc.b);
                        bulk_commit();
                    }
                    if (do_stats) {
                        if constexpr (VS) stats_chunk_v(sbase, sraw[0]);      // chunks beyond FC hold stale bytes: never flushed
                    }
                } else if (OUT_MODE == 0) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        if (tc.n0 + 64 * c >= p.N) break;
                        if (((chunk_ctr++) & 1) != (uint32_t)team) continue;
                        const int ngrp = min(64, p.n_tile - 64 * c) >> 4;      // 16-column groups in this chunk (1..4)
                        // accumulator columns are read GL groups of 16 at a time: all four at once, or (prologue instantiations,
                        // 128 registers per thread) two and two so that the statistics accumulators need not be spilled
                        constexpr int GL = (PRO || bnb_on) ? 2 : 4;      // (with a maximal shared-memory carve-out the L1 is too small for spills:
                                                                         // 24 spilled registers cost the BNB variant ~2 500 cycles per chunk)
                        uint32_t v[GL][16];
                        DMM_PH(const long long e0 = clock64();)
#pragma unroll
                        for (int g = 0; g < GL; ++g)
                            if (g < ngrp) tmem_ld16(trow + c * 64 + g * 16, v[g]);
                        if (r == 0) bulk_wait_read0();       // the team's previous TMA store has finished reading the slot
                        DMM_PH(const long long e1 = clock64();)
                        tmem_ld_wait();
                        DMM_PH(const long long e2 = clock64();)
                        epi_bar(team);                       // ... and every thread of the team is done with the previous chunk
                        DMM_PH(const long long e3 = clock64(); ph_store_wait += e1 - e0; ph_tmem += e2 - e1; ph_bar1 += e3 - e2;)
                        // fused BN backward reduce: the team's other x buffer is free (the barrier above), load the NEXT chunk's tile
                        DMM_PH(const long long xp0 = clock64();)
                        if (bnb_on && r == 0) x_prefetch_next();
                        DMM_PH(if (bnb_on) ph_xpre += clock64() - xp0;)
#pragma unroll
                        for (int h = 0; h < 4 / GL; ++h) {
                            if (h > 0) {
#pragma unroll
                                for (int g = 0; g < GL; ++g)
                                    if (h * GL + g < ngrp) tmem_ld16(trow + c * 64 + (h * GL + g) * 16, v[g]);
                                tmem_ld_wait();
                            }
                            if (p.epi_bias) {
                                // folded BatchNorm shift (+ ReLU) on the fp32 accumulators; the bias vector is padded to n_rows entries
                                const float4* bp = reinterpret_cast<const float4*>(p.epi_bias + tc.n0 + 64 * c);
                                const float lo = p.epi_relu ? 0.f : -INFINITY;
#pragma unroll
                                for (int g = 0; g < GL; ++g) {
                                    if (h * GL + g < ngrp) {
                                        asm volatile("" ::: "memory");        // keep the 4 bias loads of one group together: 16 hoisted float4 spill
#pragma unroll
                                        for (int q = 0; q < 4; ++q) {
                                            const float4 b4 = __ldg(bp + 4 * (h * GL + g) + q);
                                            v[g][4 * q + 0] = __float_as_uint(fmaxf(__uint_as_float(v[g][4 * q + 0]) + b4.x, lo));
                                            v[g][4 * q + 1] = __float_as_uint(fmaxf(__uint_as_float(v[g][4 * q + 1]) + b4.y, lo));
                                            v[g][4 * q + 2] = __float_as_uint(fmaxf(__uint_as_float(v[g][4 * q + 2]) + b4.z, lo));
                                            v[g][4 * q + 3] = __float_as_uint(fmaxf(__uint_as_float(v[g][4 * q + 3]) + b4.w, lo));
                                        }
                                    }
                                }
                            }
#pragma unroll
                            for (int g = 0; g < GL; ++g) {
                                const int gg = h * GL + g;
                                if (gg < ngrp && !DMM_WI(64)) {
                                    uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
                                    if (valid) {
                                        w0.x = pack_bf16x2(__uint_as_float(v[g][0]), __uint_as_float(v[g][1]));
                                        w0.y = pack_bf16x2(__uint_as_float(v[g][2]), __uint_as_float(v[g][3]));
                                        w0.z = pack_bf16x2(__uint_as_float(v[g][4]), __uint_as_float(v[g][5]));
                                        w0.w = pack_bf16x2(__uint_as_float(v[g][6]), __uint_as_float(v[g][7]));
                                        w1.x = pack_bf16x2(__uint_as_float(v[g][8]), __uint_as_float(v[g][9]));
                                        w1.y = pack_bf16x2(__uint_as_float(v[g][10]), __uint_as_float(v[g][11]));
                                        w1.z = pack_bf16x2(__uint_as_float(v[g][12]), __uint_as_float(v[g][13]));
                                        w1.w = pack_bf16x2(__uint_as_float(v[g][14]), __uint_as_float(v[g][15]));
                                    }
                                    sts_v4(srow_u + (((2 * gg) ^ (r & 7)) << 4), w0);
                                    sts_v4(srow_u + (((2 * gg + 1) ^ (r & 7)) << 4), w1);
                                }
                            }
                        }
                        if (!DMM_WI(32)) fence_proxy_async();
                        DMM_PH(const long long e4 = clock64();)
                        epi_bar(team);
                        DMM_PH(const long long e5 = clock64();)
                        if (r == 0 && !DMM_WI(8)) {
                            tma_store_4d(&p.o_map, slot, tc.n0 + 64 * c, tc.x0 + p.sub_x[sub], tc.y0 + p.sub_y[sub], tc.b);
                            bulk_commit();
                        }
                        DMM_PH(ph_pack += e4 - e3; ph_bar2 += e5 - e4; const long long e6 = clock64();)
                        if (do_stats && !DMM_WI(16)) {
                            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
                            const uint32_t base = slot_u + ((cp & 3) << 2) + rq * 32 * 128;
                            const int j = cp >> 2;
                            if (vstats && !bnb_on) {
                                if constexpr (VS) stats_chunk_v(sbase, sraw[c]);
                            } else if (vstats) {
                                if constexpr (VS) {
                                    // per-channel (scale, shift, -mean) of the thread's 8 channels; channels beyond N get zeros (dz = 0)
                                    DMM_PH(const long long b0 = clock64();)
                                    const int ch0 = tc.n0 + 64 * c + 8 * (r & 7);
                                    uint64_t coef[12];
                                    const uint32_t cb_u = smem_u32(pcoef) + (uint32_t)ch0 * 4u;
#pragma unroll
                                    for (int a = 0; a < 3; ++a) {
                                        const uint4 lo = lds_v4(cb_u + (uint32_t)(a * p.bnb_np) * 4u), hi = lds_v4(cb_u + (uint32_t)(a * p.bnb_np) * 4u + 16u);
                                        coef[4 * a + 0] = ((uint64_t)lo.y << 32) | lo.x;
                                        coef[4 * a + 1] = ((uint64_t)lo.w << 32) | lo.z;
                                        coef[4 * a + 2] = ((uint64_t)hi.y << 32) | hi.x;
                                        coef[4 * a + 3] = ((uint64_t)hi.w << 32) | hi.z;
                                    }
                                    DMM_PH(const long long b1 = clock64();)
                                    mbar_wait(&x_bar[team * 2 + (xi & 1)], (xi >> 1) & 1);
                                    DMM_PH(const long long b2 = clock64();)
                                    bnb_chunk_v(sbase, sbase - slot_u + xslot_u + (xi & 1) * kStageSlot, coef, sraw[c]);
                                    DMM_PH(ph_coef += b1 - b0; ph_xwait += b2 - b1; ph_loop += clock64() - b2;)
                                    ++xi;
                                }
                            } else if (!bnb_on) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    // row = rq*32 + i: (row & 7) == (i & 7), a compile-time pattern after unrolling
                                    const uint32_t u = lds_u32(base + i * 128 + ((j ^ (i & 7)) << 4));
                                    const float a = bf16_lo(u), b = bf16_hi(u);
                                    s1a += a; s1b += b;
                                    s2a = fmaf(a, a, s2a); s2b = fmaf(b, b, s2b);
                                }
                            } else {
                                // fused BatchNorm-ReLU backward reduce: dz = g * [bn(x) > 0]; sums of dz and dz * (x - mean)
                                const int col = tc.n0 + 64 * c + 2 * cp;
                                const float sc0 = pcoef[col], sh0 = pcoef[p.bnb_np + col], mu0 = -pcoef[2 * p.bnb_np + col];
                                const float sc1 = pcoef[col + 1], sh1 = pcoef[p.bnb_np + col + 1], mu1 = -pcoef[2 * p.bnb_np + col + 1];
                                const uint32_t xbase = xslot_u + (xi & 1) * kStageSlot + ((cp & 3) << 2) + rq * 32 * 128;
                                mbar_wait(&x_bar[team * 2 + (xi & 1)], (xi >> 1) & 1);
                                ++xi;
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const int o = i * 128 + ((j ^ (i & 7)) << 4);
                                    const uint32_t u = lds_u32(base + o);
                                    const uint32_t xu = lds_u32(xbase + o);
                                    const float xa = bf16_lo(xu), xb = bf16_hi(xu);
                                    const float a = fmaf(xa, sc0, sh0) > 0.f ? bf16_lo(u) : 0.f;
                                    const float b = fmaf(xb, sc1, sh1) > 0.f ? bf16_hi(u) : 0.f;
                                    s1a += a; s1b += b;
                                    s2a = fmaf(a, xa - mu0, s2a); s2b = fmaf(b, xb - mu1, s2b);
                                }
                            }
                            if (!vstats) {
                                sacc_add(c, 0, s1a); sacc_add(c, 1, s1b);
                                sacc_add(c, 2, s2a); sacc_add(c, 3, s2b);
                            }
                        }
                        DMM_PH(ph_stats += clock64() - e6;)
                    }
                } else {
                    // fp32 NCHW logits: out[((b*N + n)*OH + oy)*OW + ox], N <= 16; the teams alternate sub-tiles
                    if (((chunk_ctr++) & 1) != (uint32_t)team) continue;
                    uint32_t v[16];
                    tmem_ld16(trow, v);
                    tmem_ld_wait();
                    const int oy = y * p.out_sy + p.out_py, ox = x * p.out_sx + p.out_px;
                    if (valid && oy < p.OH && ox < p.OW) {
                        const long long plane = (long long)p.OH * p.OW;
                        float* o = p.out32 + ((long long)tc.b * p.N + tc.n0) * plane + (long long)oy * p.OW + ox;
#pragma unroll
                        for (int jn = 0; jn < 16; ++jn)
                            if (tc.n0 + jn < p.N) o[jn * plane] = __uint_as_float(v[jn]);
                    }
                }
            }
            // all tcgen05.ld of this warp on this accumulator stage have completed: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
            if (do_stats && p.tiles_n > 1) flush_stats(tc.n0);
        }
        if (do_stats && p.tiles_n == 1 && it > 0) flush_stats(last_n0);
#undef sacc_add
#undef sacc_get
        if ((OUT_MODE == 0 || FOLD3) && r == 0) bulk_wait_all();
        if (p.prof && r == 0 && team == 0) {
            p.prof[blockIdx.x * 16 + 8] = clock64() - t_begin;
            p.prof[blockIdx.x * 16 + 9] = w_full;
            DMM_PH(p.prof[blockIdx.x * 16 + 10] = ph_store_wait; p.prof[blockIdx.x * 16 + 11] = ph_tmem; p.prof[blockIdx.x * 16 + 12] = ph_bar1;
                   p.prof[blockIdx.x * 16 + 13] = ph_pack; p.prof[blockIdx.x * 16 + 14] = ph_bar2; p.prof[blockIdx.x * 16 + 15] = ph_stats;
                   if (bnb_on) { p.prof[blockIdx.x * 16 + 10] = ph_xpre; p.prof[blockIdx.x * 16 + 11] = ph_coef; p.prof[blockIdx.x * 16 + 12] = ph_xwait;
                                 p.prof[blockIdx.x * 16 + 14] = ph_loop; })
        }
        }   // epilogue team
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

int view_to_tmap(CUtensorMap* out, const dmm_view_t& v, int box_c, int box_w, int box_h, int swizzle);

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

struct Tiling {
    int msub, nsx, nsy, sub_w, sub_h, TW, TH, sa, sb, tps, w_res;
    uint32_t a_stage;
    long long tiles;
    double cost;
};

static int igemm2_launch_impl(const dmm_igemm_t* d, cudaStream_t stream, bool swapped);

// KxK launches tile the image in sub-tiles of 8 (x) by 16 (y) pixels (an 8-pixel core-matrix group is one patch row segment).
// A 30 x 20 or 60 x 40 image (dense blocks 3 / 4, the deep decoder stages) wastes 60 % / 20 % of every tile column in y; with the
// roles of x and y exchanged - views with swapped extents and strides, taps with swapped offsets, the output / x tensor maps built
// with swapped strides - the same kernel covers it in 8 (y) by 16 (x) sub-tiles: 25 % / 17 % fewer tiles.  Nothing in the kernel
// knows about it.
int igemm2_launch(const dmm_igemm_t* d, cudaStream_t stream) {
    static const int swap_env = env_int("DMM_IGEMM_SWAP_XY", 1);
    bool halo = false;
    for (int t = 0; t < d->num_taps; ++t) halo = halo || d->tap_dx[t] != 0 || d->tap_dy[t] != 0;
    if (swap_env && halo && d->out_mode == 0 && d->W > 0 && d->H > 0) {
        const long long plain = (long long)ceil_div(d->W, 8) * 8 * ceil_div(d->H, 16) * 16;
        const long long swp = (long long)ceil_div(d->H, 8) * 8 * ceil_div(d->W, 16) * 16;
        if (swp * 20 < plain * 19) {
            dmm_igemm_t t = *d;
            t.W = d->H; t.H = d->W;
            for (int i = 0; i < d->num_taps; ++i) { t.tap_dx[i] = d->tap_dy[i]; t.tap_dy[i] = d->tap_dx[i]; }
            for (int s = 0; s < d->num_src; ++s) {
                t.src[s].W = d->src[s].H; t.src[s].H = d->src[s].W;
                t.src[s].sw = d->src[s].sh; t.src[s].sh = d->src[s].sw;
            }
            t.out_sx = d->out_sy; t.out_sy = d->out_sx; t.out_px = d->out_py; t.out_py = d->out_px;
            t.OW = d->OH; t.OH = d->OW;
            return igemm2_launch_impl(&t, stream, true);
        }
    }
    return igemm2_launch_impl(d, stream, false);
}

static int igemm2_launch_impl(const dmm_igemm_t* d, cudaStream_t stream, bool swapped) {
    DMM_CHECK(d->kwidth == 64 || d->kwidth == 32 || d->kwidth == 16, "igemm v2: kwidth must be 64, 32 or 16");
    const int tpk = 64 / d->kwidth;                // kwidth 16 / 32: one source of <= kwidth channels, weights packed [n][tap*kwidth + c]
    DMM_CHECK(d->n_tile % 64 == 0 || d->n_tile >= d->N, "igemm v2: n_tile %d must be a multiple of 64 or cover N=%d", d->n_tile, d->N);
    DMM_CHECK(d->out_mode == 0 || d->out_mode == 3 || (d->N <= 16 && d->n_tile == 16), "igemm v2: fp32 NCHW output needs N <= 16");
    const bool fold = d->out_mode == 2 || d->out_mode == 3;
    if (d->out_mode == 3)
        DMM_CHECK(d->fold_kw == 3 && (d->N == 96 || d->N == 192) && d->n_tile == d->N && d->tile_w == 32 && d->bnb_sums == nullptr && !d->pro_enable,
                  "igemm v2: out_mode 3 is a 3x3 convolution with 32 or 64 output channels whose kernel columns are folded into N = 96 / 192");
    if (fold) {
        DMM_CHECK(d->fold_kw >= 1 && (d->fold_kw & 1) && d->N % d->fold_kw == 0 && d->kwidth == 64 && d->num_src == 1,
                  "igemm v2: out_mode 2 needs an odd fold_kw dividing N, one source, kwidth 64");
        DMM_CHECK(d->out_sy <= 1 && d->out_sx <= 1 && d->out_py == 0 && d->out_px == 0 && (d->OH <= 0 || d->OH == d->H) && (d->OW <= 0 || d->OW == d->W),
                  "igemm v2: out_mode 2: no output stride / phase");
        for (int t = 0; t < d->num_taps; ++t) DMM_CHECK(d->tap_dx[t] == 0, "igemm v2: out_mode 2: taps must be kernel rows (dx == 0)");
        DMM_CHECK(d->tile_w > d->fold_kw - 1, "igemm v2: out_mode 2: tile_w %d too narrow", d->tile_w);
    }
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        DMM_CUDA(cudaGetDevice(&dev));
        DMM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    static const int force_msub = env_int("DMM_IGEMM_MSUB", 0);

    Ig2Params p;
    memset(&p, 0, sizeof(p));
    // ---- taps grouped by source, halo extents ----
    int mindx[DMM_MAX_SRC], maxdx[DMM_MAX_SRC], mindy[DMM_MAX_SRC], maxdy[DMM_MAX_SRC], ntap[DMM_MAX_SRC];
    for (int s = 0; s < DMM_MAX_SRC; ++s) { mindx[s] = mindy[s] = 127; maxdx[s] = maxdy[s] = -128; ntap[s] = 0; }
    for (int t = 0; t < d->num_taps; ++t) {
        const int s = d->tap_src[t];
        DMM_CHECK(s >= 0 && s < d->num_src, "igemm v2: tap %d bad source", t);
        mindx[s] = d->tap_dx[t] < mindx[s] ? d->tap_dx[t] : mindx[s];
        maxdx[s] = d->tap_dx[t] > maxdx[s] ? d->tap_dx[t] : maxdx[s];
        mindy[s] = d->tap_dy[t] < mindy[s] ? d->tap_dy[t] : mindy[s];
        maxdy[s] = d->tap_dy[t] > maxdy[s] ? d->tap_dy[t] : maxdy[s];
        ++ntap[s];
    }
    if (tpk > 1) {
        int tapped = 0;
        for (int s = 0; s < d->num_src; ++s) {
            if (ntap[s] == 0) continue;
            ++tapped;
            DMM_CHECK(d->src[s].C <= d->kwidth, "igemm v2: kwidth %d needs sources of at most %d channels (source %d has %d)", d->kwidth, d->kwidth, s, d->src[s].C);
        }
        DMM_CHECK(tapped == 1, "igemm v2: kwidth 16 / 32 support a single tapped source");
    }
    int hx = 0, hy = 0;   // largest halo over the sources
    long long ktot = 0;
    int nkb_total = 0;    // (source, block) groups per tile
    int nmma_steps = 0;   // sum over groups of taps * ksteps
    for (int s = 0; s < d->num_src; ++s) {
        const dmm_view_t& v = d->src[s];
        DMM_CHECK(v.ptr != nullptr && v.C >= 1, "igemm v2: source %d empty", s);
        p.src_nblk[s] = ceil_div(v.C, 64);
        p.src_lastk[s] = ceil_div(v.C - (p.src_nblk[s] - 1) * 64, 16);
        if (ntap[s] == 0) continue;
        hx = (maxdx[s] - mindx[s]) > hx ? (maxdx[s] - mindx[s]) : hx;
        hy = (maxdy[s] - mindy[s]) > hy ? (maxdy[s] - mindy[s]) : hy;
        p.src_ox[s] = mindx[s];
        p.src_oy[s] = mindy[s];
        nkb_total += p.src_nblk[s];
        nmma_steps += ntap[s] * ((p.src_nblk[s] - 1) * 4 + p.src_lastk[s]);
    }
    // taps sorted by source; k-block base of every tap in the (tap-major) packed weight matrix
    {
        int kb0[DMM_MAX_TAPS];
        int kb = 0;
        for (int t = 0; t < d->num_taps; ++t) {
            kb0[t] = kb;
            kb += p.src_nblk[d->tap_src[t]];
        }
        ktot = (long long)kb * 64;
        if (tpk > 1) {
            // packed: tap t owns K columns [kwidth t, kwidth (t + 1)): block t / tpk (the producer only uses taps that start a block)
            for (int t = 0; t < d->num_taps; ++t) kb0[t] = t / tpk;
            ktot = (long long)d->num_taps * d->kwidth;
        }
        int n = 0;
        for (int s = 0; s < d->num_src; ++s) {
            p.src_tap0[s] = n;
            for (int t = 0; t < d->num_taps; ++t)
                if (d->tap_src[t] == s) {
                    p.tap_kb0[n] = kb0[t];
                    p.tap_aoff[n] = (uint32_t)t;     // original index for now; byte offset filled below
                    ++n;
                }
        }
        for (int s = d->num_src; s <= DMM_MAX_SRC; ++s) p.src_tap0[s] = n;
    }
    DMM_CHECK(ktot == d->ktot, "dmm_conv_igemm: packed weight K (%lld) != tap table K (%lld)", (long long)d->ktot, ktot);
    DMM_CHECK(d->n_rows >= d->N, "dmm_conv_igemm: weight rows %d < N %d", d->n_rows, d->N);

    // ---- tiling ----
    const bool xhalo = hx > 0;
    const bool bnb = d->bnb_sums != nullptr;
    if (bnb) {
        DMM_CHECK(d->out_mode == 0 && d->stats == nullptr, "igemm v2: fused BN backward reduce needs out_mode 0 and no forward statistics");
        DMM_CHECK(!d->pro_enable, "igemm v2: the fused BN backward reduce and the BN-ReLU prologue share barrier slots (never used together)");
        DMM_CHECK(d->out_sy <= 1 && d->out_sx <= 1 && d->out_py == 0 && d->out_px == 0, "igemm v2: fused BN backward reduce: no output stride");
        DMM_CHECK(d->bnb_x && d->bnb_mean && d->bnb_invstd && d->bnb_ldx % 8 == 0, "igemm v2: fused BN backward reduce: missing inputs");
    }
    const bool pro = d->pro_enable != 0;
    int pro_kp = 0;
    if (pro) {
        DMM_CHECK(d->out_mode == 0 && d->num_src == 1 && d->kwidth == 64 && d->out_sy <= 1 && d->out_sx <= 1,
                  "igemm v2: the BN-ReLU prologue needs one source, kwidth 64 and a stride-1 output");
        DMM_CHECK(d->pro_bn.training ? (d->pro_bn.stats != nullptr && d->pro_bn.count > 0) : (d->pro_bn.running_mean && d->pro_bn.running_var),
                  "igemm v2: prologue BatchNorm without statistics");
        pro_kp = ceil_div(d->src[0].C, 64) * 64;
    }
    int max_taps = 1;
    for (int s = 0; s < d->num_src; ++s) max_taps = ntap[s] > max_taps ? ntap[s] : max_taps;
    const int nslot = 1;      // staging slots per epilogue team (2 / 3 measured slower: profiles/r02_epilogue_experiments.txt)
    const int staging = (d->out_mode == 0 ? (2 * nslot + (bnb ? 4 : 0)) * (int)kStageSlot
                                          : (d->out_mode == 3 ? 2 * (int)kStageSlot : (fold ? kMaxSub * 128 * (int)kFoldPitch : 0))) +
                        (pro ? 2 * pro_kp * (int)sizeof(float) : 0) + (bnb ? 3 * (ceil_div(d->N, d->n_tile) * d->n_tile + 8) * (int)sizeof(float) : 0);
    const int avail = kG2MaxSmem - 1024 - kTailBytes - staging;
    const uint32_t b_tap = (uint32_t)d->n_tile * 128u;      // one tap's [n_tile x 64] weight slice
    const int tiles_n = ceil_div(d->N, d->n_tile);
    const int mma_hw = d->n_tile / 2 > 32 + d->n_tile / 4 ? d->n_tile / 2 : 32 + d->n_tile / 4;   // cycles per MMA (measured law)
    Tiling best;
    best.msub = 0;
    for (int m = kMaxSub; m >= 1; m >>= 1) {
        if (2 * m * d->n_tile > 512) continue;
        if (force_msub && m != force_msub && m != 1) continue;
        for (int nsx = 1; nsx <= m; nsx <<= 1) {
            const int nsy = m / nsx;
            Tiling c;
            c.msub = m; c.nsx = nsx; c.nsy = nsy;
            if (xhalo) { c.sub_w = 8; c.sub_h = 16; }
            else {
                c.sub_w = d->tile_w; c.sub_h = 128 / d->tile_w;
                if (c.sub_h == 1) { if (nsy != 1) continue; }       // one-row sub-tiles sit side by side
                else if (nsx != 1) continue;                        // otherwise stacked (patch pitch == sub_w)
            }
            c.TW = c.sub_w * nsx; c.TH = c.sub_h * nsy;
            int pw = c.TW + hx, ph = c.TH + hy;
            if (pw > 256 || ph > 256) continue;
            c.a_stage = ((uint32_t)pw * ph * 128u + 1023u) & ~1023u;
            // shared memory: 2 patch stages first, then weight stages of `tps` taps (the MMA warp pays ~300-400 cycles
            // per stage hand-over, so a stage should carry many MMAs), then more patch stages with what is left
            const int rest = avail - 2 * (int)c.a_stage;
            if (rest < 2 * (int)b_tap) continue;
            static const int min_sb = env_int("DMM_IGEMM_MIN_SB", 2);
            int blk_max = rest / (min_sb * (int)b_tap);                 // weight blocks per stage
            if (blk_max < 1) blk_max = rest / (2 * (int)b_tap);
            if (blk_max * (int)b_tap > 48 * 1024) blk_max = (48 * 1024) / (int)b_tap > 0 ? (48 * 1024) / (int)b_tap : 1;
            int tps_max = blk_max * tpk;
            if (tps_max > max_taps) tps_max = max_taps;
            const int groups = ceil_div(max_taps, tps_max);
            c.tps = ceil_div(ceil_div(max_taps, groups), tpk) * tpk;
            const int b_stage_c = ceil_div(c.tps, tpk) * (int)b_tap;
            c.sb = rest / b_stage_c;
            if (c.sb > 4) c.sb = 4;
            static const int sb1 = env_int("DMM_IGEMM_SB1", 6);      // 1x1: one weight block per stage
            if (max_taps == 1) {
                if (sb1 >= 6) { if (b_stage_c <= 16384 && rest / b_stage_c >= 6) c.sb = 6; }
                else { if (c.sb > sb1) c.sb = sb1; if (c.sb < 2) c.sb = 2; }
            }
            int sa = (avail - c.sb * b_stage_c) / (int)c.a_stage;
            static const int sa_cap1 = env_int("DMM_IGEMM_SA_CAP1", 4);      // 1x1 launches: cap of the A ring (prefetch depth)
            const int sa_cap = max_taps == 1 ? sa_cap1 : 4;
            c.sa = sa > sa_cap ? sa_cap : sa;
            c.tiles = (long long)ceil_div(d->W, fold ? c.TW - (d->fold_kw - 1) : c.TW) * ceil_div(d->H, c.TH) * d->B * tiles_n;
            // crude per-tile time (cycles): L2 -> smem bytes at 32 B/cycle/SM vs MMA time + stage hand-overs
            const double l2 = (double)nkb_total * pw * ph * 128.0 / 32.0;
            double wbytes = 0, waits = 0;
            for (int s = 0; s < d->num_src; ++s) {
                wbytes += (double)ceil_div(ntap[s], tpk) * p.src_nblk[s] * b_tap;
                if (ntap[s]) waits += (double)p.src_nblk[s] * (1 + ceil_div(ntap[s], c.tps));
            }
            const double mma_cyc = (double)m * nmma_steps * mma_hw + 350.0 * waits;
            const double per_tile = (l2 + wbytes / 32.0 > mma_cyc ? l2 + wbytes / 32.0 : mma_cyc) + 1500.0;
            const long long rounds = (c.tiles + num_sms - 1) / num_sms;
            c.cost = (double)rounds * per_tile;
            c.w_res = 0;
            if (force_msub && m == force_msub) c.cost = -1.0;
            if (best.msub == 0 || c.cost < best.cost) best = c;
            // RESIDENT weights: the tile loop of a CTA keeps one weight slice (tiles_n == 1, or a grid that is a multiple of
            // tiles_n so that n0 is the same for all of its tiles); every (tap group of tpk, k-block) gets its own ring slot.
            // Weight rows that are not re-streamed per tile free ring stages, barrier hand-overs and MMA-warp waits (L2-hit weight
            // tiles arrive at 58 B/cycle/SM - scripts/ubench/load_rate.cu - so it is the hand-overs, not the bytes, that cost).
            static const int wres_on = env_int("DMM_IGEMM_WRES", 1);
            if (wres_on && (tiles_n == 1 || num_sms / tiles_n >= 1)) {
                int stages_r = 0;
                for (int s = 0; s < d->num_src; ++s) stages_r += p.src_nblk[s] * ceil_div(ntap[s], tpk);
                const long long bytes_r = (long long)stages_r * b_tap;
                const int grid_r = tiles_n == 1 ? num_sms : (num_sms / tiles_n) * tiles_n;
                if (stages_r <= kMaxBStages && bytes_r + 2 * (long long)c.a_stage <= avail) {
                    Tiling r = c;
                    r.w_res = 1;
                    r.tps = tpk;
                    r.sb = stages_r;
                    int sa_r = (int)((avail - bytes_r) / (long long)c.a_stage);
                    r.sa = sa_r > sa_cap ? sa_cap : sa_r;
                    double waits_r = 0;
                    for (int s = 0; s < d->num_src; ++s)
                        if (ntap[s]) waits_r += (double)p.src_nblk[s];
                    const double mma_r = (double)m * nmma_steps * mma_hw + 350.0 * waits_r;
                    const double per_tile_r = (l2 > mma_r ? l2 : mma_r) + 1500.0;
                    const long long rounds_r = (c.tiles + grid_r - 1) / grid_r;
                    r.cost = (double)rounds_r * per_tile_r + (double)bytes_r / 32.0;
                    if (force_msub && m == force_msub) r.cost = -2.0;
                    if (r.cost < best.cost) best = r;
                }
            }
        }
    }
    DMM_CHECK(best.msub > 0, "igemm v2: no tiling fits in shared memory (n_tile %d, halo %dx%d)", d->n_tile, hx, hy);
    p.msub = best.msub; p.sub_w = best.sub_w; p.sub_h = best.sub_h;
    p.TW = best.TW; p.TH = best.TH;
    p.sa = best.sa; p.sb = best.sb;
    p.w_res = best.w_res;
    p.a_stage = best.a_stage; p.b_tap = b_tap; p.tps = best.tps; p.tpk = tpk;
    p.tpk_log = tpk == 4 ? 2 : (tpk == 2 ? 1 : 0);
    p.b_stage = (uint32_t)ceil_div(best.tps, tpk) * b_tap;
    {
        // ring stages one tile consumes (the MMA warp that skips a tile advances its ring positions by these)
        int a_per = 0, b_per = 0;
        for (int s = 0; s < d->num_src; ++s) {
            if (ntap[s] == 0) continue;
            p.last_src = s;
            a_per += p.src_nblk[s];
            b_per += p.src_nblk[s] * ceil_div(ntap[s], best.tps);
        }
        static const int mma2_env = env_int("DMM_IGEMM_MMA2", 1);
        p.mma2 = mma2_env;
        p.a_adv = a_per % best.sa; p.a_flip = (uint32_t)(a_per / best.sa) & 1u;
        p.b_adv = best.w_res ? 0 : b_per % best.sb; p.b_flip = best.w_res ? 0u : ((uint32_t)(b_per / best.sb) & 1u);
    }
    p.x_step = fold ? p.TW - (d->fold_kw - 1) : p.TW;
    p.x_org = fold ? -(d->fold_kw / 2) : 0;
    p.nsx = best.nsx;
    p.fold_kw = fold ? d->fold_kw : 0;
    p.fold_c = fold ? d->N / d->fold_kw : 0;
    if (fold) {
        DMM_CHECK((d->out_mode == 3 || p.fold_c <= 4) && best.nsx == 1, "igemm v2: out_mode 2 supports at most 4 classes (got %d)", p.fold_c);
        DMM_CHECK(d->out_mode != 3 || (p.sub_w == 32 && p.sub_h == 4), "igemm v2: out_mode 3 needs 32 x 4 sub-tiles");
        auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
        p.tw_log = ilog2(p.TW); p.sw_log = ilog2(p.sub_w); p.sh_log = ilog2(p.sub_h);
        DMM_CHECK((1 << p.tw_log) == p.TW && (1 << p.sw_log) == p.sub_w && (1 << p.sh_log) == p.sub_h && p.TW <= 256,
                  "igemm v2: out_mode 2 needs power-of-two tiles");
    }
    p.tiles_x = ceil_div(d->W, p.x_step);
    p.tiles_y = ceil_div(d->H, p.TH);
    p.tiles_n = tiles_n;
    DMM_CHECK(best.tiles < (1ll << 31), "igemm v2: %lld tiles", best.tiles);
    p.fd_n = make_fastdiv(p.tiles_n); p.fd_x = make_fastdiv(p.tiles_x); p.fd_y = make_fastdiv(p.tiles_y);
    p.total_tiles = best.tiles;
    p.W = d->W; p.H = d->H; p.B = d->B;
    p.n_tile = d->n_tile; p.N = d->N;
    p.num_src = d->num_src;
    for (int i = 0; i < p.msub; ++i) {
        p.sub_x[i] = (i % best.nsx) * p.sub_w;
        p.sub_y[i] = (i / best.nsx) * p.sub_h;
    }
    // ---- per-source patch geometry ----
    for (int s = 0; s < d->num_src; ++s) {
        if (ntap[s] == 0) continue;
        int pw = p.TW + (maxdx[s] - mindx[s]), ph = p.TH + (maxdy[s] - mindy[s]);
        DMM_CHECK((uint32_t)pw * ph * 128u <= p.a_stage, "igemm v2: internal patch size error");
        p.src_tx[s] = (uint32_t)pw * ph * 128u;
        // 8-row groups: 8 consecutive pixels of one patch row.  sub_w == 8: next group = next patch row;
        // otherwise the sub-tile's pixels are contiguous in the patch (pitch == sub_w or a single row).
        if (p.sub_w == 8) p.src_sbo[s] = (uint32_t)pw * 128u;
        else {
            DMM_CHECK(pw == p.sub_w || p.sub_h == 1, "igemm v2: internal sub-tile layout error");
            p.src_sbo[s] = 1024u;
        }
        for (int i = 0; i < p.msub; ++i) p.sub_aoff[s][i] = (uint32_t)(p.sub_y[i] * pw + p.sub_x[i]) * 128u;
        for (int n = p.src_tap0[s]; n < p.src_tap0[s + 1]; ++n) {
            const int t = (int)p.tap_aoff[n];
            p.tap_aoff[n] = (uint32_t)((d->tap_dy[t] - mindy[s]) * pw + (d->tap_dx[t] - mindx[s])) * 128u;
        }
        int rc = view_to_tmap(&p.a_maps[s], d->src[s], 64, pw, ph, 128);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->ktot, (uint64_t)d->n_rows};
        uint64_t strides[1] = {(uint64_t)d->ktot};
        uint32_t box[2] = {64u, (uint32_t)d->n_tile};
        int rc = make_tmap_bf16(&p.b_map, d->weights, 2, dims, strides, box, 128);
        if (rc) return rc;
    }
    const int osy = d->out_sy > 0 ? d->out_sy : 1, osx = d->out_sx > 0 ? d->out_sx : 1;
    const int OH = d->OH > 0 ? d->OH : d->H, OW = d->OW > 0 ? d->OW : d->W;
    DMM_CHECK(d->out_py < OH && d->out_px < OW, "igemm v2: output phase outside the output");
    const int OHv = (OH - d->out_py + osy - 1) / osy, OWv = (OW - d->out_px + osx - 1) / osx;
    p.Wv = d->W < OWv ? d->W : OWv;
    p.Hv = d->H < OHv ? d->H : OHv;
    p.OH = OH; p.OW = OW; p.out_sy = osy; p.out_sx = osx; p.out_py = d->out_py; p.out_px = d->out_px;
    if (d->out_mode == 0 || d->out_mode == 3) {
        DMM_CHECK(d->ldo % 8 == 0 && d->coff % 8 == 0, "igemm v2: output pitch %lld / channel offset %d must be multiples of 8",
                  (long long)d->ldo, d->coff);
        dmm_view_t ov;
        // memory is [B][rows][columns][ldo]; swapped: this launch's x runs over the rows (OH = number of columns)
        ov.ptr = reinterpret_cast<const uint16_t*>(d->out) + (swapped ? ((long long)d->out_px * OH + d->out_py) : ((long long)d->out_py * OW + d->out_px)) * d->ldo + d->coff;
        ov.C = d->out_mode == 3 ? d->N / d->fold_kw : d->N; ov.W = OWv; ov.H = OHv; ov.B = d->B;
        ov.sw = swapped ? (long long)osx * OH * d->ldo : (long long)osx * d->ldo;
        ov.sh = swapped ? (long long)osy * d->ldo : (long long)osy * OW * d->ldo;
        ov.sb = (long long)OH * OW * d->ldo;
        int rc = view_to_tmap(&p.o_map, ov, 64, d->out_mode == 3 ? p.sub_w - 2 : p.sub_w, p.sub_h, 128);
        if (rc) return rc;
    } else {
        p.out32 = reinterpret_cast<float*>(d->out);
    }
    p.stats = (d->out_mode == 0 || d->out_mode == 3) ? d->stats : nullptr;
    p.stats_ld = d->stats_ld;
    p.stats_off = d->stats_off;
    if (pro) {
        p.pro = 1;
        p.pro_kp = pro_kp;
        p.pro_c = d->src[0].C;
        p.pro_pw = (hx > 0 || hy > 0) ? p.TW + (maxdx[0] - mindx[0]) : 0;
        p.pro_bn = d->pro_bn;
    }
    if (d->epi_bias) {
        DMM_CHECK(d->out_mode == 0 && !bnb && (reinterpret_cast<uintptr_t>(d->epi_bias) & 15) == 0,
                  "igemm v2: the bias epilogue needs out_mode 0, no fused BN backward and a 16-byte aligned bias vector");
        p.epi_bias = d->epi_bias;
        p.epi_relu = d->epi_relu;
    }
    if (bnb) {
        dmm_view_t xv;
        xv.ptr = d->bnb_x;
        xv.C = d->N; xv.W = d->W; xv.H = d->H; xv.B = d->B;
        xv.sw = swapped ? (long long)d->H * d->bnb_ldx : d->bnb_ldx;
        xv.sh = swapped ? d->bnb_ldx : (long long)d->W * d->bnb_ldx;
        xv.sb = (long long)d->H * d->W * d->bnb_ldx;
        int rc = view_to_tmap(&p.x_map, xv, 64, p.sub_w, p.sub_h, 128);
        if (rc) return rc;
        p.bnb = 1;
        p.bnb_np = tiles_n * d->n_tile + 8;
        p.bnb_gamma = d->bnb_gamma; p.bnb_beta = d->bnb_beta; p.bnb_mean = d->bnb_mean; p.bnb_invstd = d->bnb_invstd;
        p.stats = d->bnb_sums; p.stats_ld = d->bnb_sums_ld; p.stats_off = d->bnb_sums_off;
    }
    uint32_t cols = 32;
    while ((int)cols < 2 * p.msub * p.n_tile) cols <<= 1;
    p.tmem_cols = cols;

    const size_t smem = (size_t)p.sa * p.a_stage + (size_t)p.sb * p.b_stage + staging + kTailBytes + 1024;
    DMM_CHECK(smem <= (size_t)kG2MaxSmem, "igemm v2: %zu bytes of shared memory requested", smem);
    unsigned grid = (unsigned)(p.total_tiles < num_sms ? p.total_tiles : num_sms);
    if (p.w_res && tiles_n > 1) {
        const unsigned g = (unsigned)((num_sms / tiles_n) * tiles_n);
        grid = p.total_tiles < (long long)g ? (unsigned)p.total_tiles : g;       // total_tiles is a multiple of tiles_n
    }
    const int nch = ceil_div(p.n_tile, 64);
    typedef void (*KernelFn)(const Ig2Params);
    KernelFn fn = nullptr;
#define DMM_IG2_PICK(PROFLAG)                                                                                                              \
    do {                                                                                                                                   \
        if (nch == 1) fn = p.msub == 4 ? igemm2_kernel<1, 0, 4, PROFLAG> : (p.msub == 2 ? igemm2_kernel<1, 0, 2, PROFLAG> : igemm2_kernel<1, 0, 1, PROFLAG>); \
        else if (nch == 2) fn = p.msub == 2 ? igemm2_kernel<2, 0, 2, PROFLAG> : igemm2_kernel<2, 0, 1, PROFLAG>;                           \
        else if (nch == 3) fn = igemm2_kernel<3, 0, 1, PROFLAG>;                                                                           \
        else fn = igemm2_kernel<4, 0, 1, PROFLAG>;                                                                                         \
    } while (0)
    if (d->out_mode == 3 && d->N == 192) fn = igemm2_kernel<1, 4, 1, 0>;
    else if (d->out_mode == 3) fn = p.msub == 2 ? igemm2_kernel<1, 3, 2, 0> : igemm2_kernel<1, 3, 1, 0>;
    else if (d->out_mode == 2) fn = p.msub == 4 ? igemm2_kernel<1, 2, 4, 0> : (p.msub == 2 ? igemm2_kernel<1, 2, 2, 0> : igemm2_kernel<1, 2, 1, 0>);
    else if (d->out_mode == 1) fn = p.msub == 4 ? igemm2_kernel<1, 1, 4, 0> : (p.msub == 2 ? igemm2_kernel<1, 1, 2, 0> : igemm2_kernel<1, 1, 1, 0>);
    else if (pro) DMM_IG2_PICK(1);
    else if (bnb) DMM_IG2_PICK(2);
    else DMM_IG2_PICK(0);
#undef DMM_IG2_PICK
    DMM_CHECK(nch <= 2 || p.msub == 1, "igemm v2: internal msub error");
    DMM_CHECK(nch <= 1 || p.msub <= 2, "igemm v2: internal msub error");
    {
        static KernelFn configured[32];
        static int nconf = 0;
        bool seen = false;
        for (int i = 0; i < nconf; ++i) seen = seen || configured[i] == fn;
        if (!seen) {
            DMM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kG2MaxSmem));
            if (nconf < 32) configured[nconf++] = fn;
        }
    }
    static const int prof = env_int("DMM_IGEMM_PROF", 0);
    static long long* prof_buf = nullptr;
#ifdef DMM_IGEMM_WHATIF
    p.whatif = env_int("DMM_IGEMM_WHATIF", 0);
#endif
    if (prof) {
        if (!prof_buf) DMM_CUDA(cudaMalloc(&prof_buf, 160 * 16 * sizeof(long long)));
        DMM_CUDA(cudaMemsetAsync(prof_buf, 0, 160 * 16 * sizeof(long long), stream));
        p.prof = prof_buf;
    }
    launch_k(fn, grid, pro ? kG2ThreadsPro : kG2Threads, smem, stream, p);
    if (prof) {
        static long long h[160 * 16];
        DMM_CUDA(cudaStreamSynchronize(stream));
        DMM_CUDA(cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost));
        double a[16] = {0};
        for (unsigned i = 0; i < grid; ++i)
            for (int j = 0; j < 16; ++j) a[j] += (double)h[i * 16 + j] / grid;
        fprintf(stderr,
                "[ig2] tiles %lld grid %u msub %d n_tile %d TWxTH %dx%d sa %d sb %d tps %d wres %d a_stage %u | producer total %.0f wait a_empty %.0f b_empty %.0f | "
                "mma total %.0f wait a_full %.0f b_full %.0f acc_empty %.0f (issue blocks %.0f, k-block headers %.0f) | epilogue total %.0f wait acc_full %.0f | team 0 phases: store-read wait %.0f tmem %.0f bar1 %.0f pack %.0f bar2 %.0f stats %.0f (cycles, CTA average)\n",
                p.total_tiles, grid, p.msub, p.n_tile, p.TW, p.TH, p.sa, p.sb, p.tps, p.w_res, p.a_stage, a[0], 0.0, a[2], a[4], a[5], a[6], a[7], a[1], a[3], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15]);
    }
    DMM_LAUNCH_CHECK("igemm2_kernel");
    return 0;
}

}  // namespace dmm
