/* dmmfods_b200 — C-ABI of the B200 (sm_100a) Dense-U-Net hot-path kernels.
 *
 * The reference (p-mc-grath/DMMFODS) has NO FFI / operator layer: its hot path is a Python
 * torch.nn.Module (dmmfods/graphs/models/Dense_U_Net_lidar.py:18-267) calling ATen.  This C-ABI
 * is therefore new; each entry point names the reference operator(s) (file:line) whose arithmetic
 * it replaces.  "tv:" = torchvision/models/densenet.py (third-party dependency of the reference,
 * Dense_U_Net_lidar.py:9).  "helper:" = dmmfods/utils/Dense_U_Net_lidar_helper.py.
 * "Agent:" = dmmfods/agents/Dense_U_Net_lidar_Agent.py.
 *
 * Conventions
 *   - plain pointers + sizes; no allocation, no ownership transfer, no hidden synchronisation:
 *     every call only enqueues work on `stream` (a cudaStream_t passed as void*).
 *   - returns 0 on success, negative on error; dmm_last_error() gives the thread-local message.
 *   - activations are NHWC ("pixel-major") bf16 matrices [P = B*H*W rows, ld elements per row];
 *     a dmm_view_t describes a channel slice of such a buffer (concat-free dense-block buffers:
 *     replaces torch.cat at tv:48, tv:124, Dense_U_Net_lidar.py:228-231,244,258,264).
 *   - batch-norm statistics are accumulated in double: stats[slot][2][ld] (sum, sum of squares),
 *     DMM_STATS_SLOTS slots spread the atomics; consumers add the slots up.  Caller zero-fills.
 */
#ifndef DMMFODS_B200_H
#define DMMFODS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMM_MAX_SRC 4
#define DMM_MAX_TAPS 32
#define DMM_STATS_SLOTS 8

const char* dmm_last_error(void);
int dmm_version(void);
/* 1 if a CUDA device of compute capability 10.x is present and usable. */
int dmm_device_ok(void);

/* A (channel-slice, optionally strided) view of an NHWC bf16 buffer.
 * element (b, y, x, c) lives at ptr[b*sb + y*sh + x*sw + c]; strides in elements. */
typedef struct {
    const void* ptr;
    int32_t C, W, H, B;
    int64_t sw, sh, sb;
} dmm_view_t;

/* ---- batch norm (nn.BatchNorm2d semantics) ---------------------------------------- */
/* Forward-side description of ONE BatchNorm2d (or a channel slice of one; all pointers already
 * offset to the slice's first channel).  training=1: batch statistics are taken from the
 * accumulated column sums stats[slot][2][stats_ld] at stats_off (all slots summed; `count` elements
 * per channel): normalise with mean / BIASED variance, save mean/invstd for backward, update
 * running stats with momentum and the UNBIASED variance over count*rep elements (rep = replication
 * factor of nn.Upsample, else 1).  training=0: running statistics. */
typedef struct {
    const double* stats;
    int32_t stats_ld, stats_off;
    double count, rep;
    const float* gamma;
    const float* beta;
    float* running_mean;     /* nullable in training mode (no update) */
    float* running_var;
    float* save_mean;        /* nullable */
    float* save_invstd;
    float eps, momentum;
    int32_t training;
} dmm_bn_t;

/* Implicit-GEMM convolution on tcgen05 (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue).
 *   out[pix(b, y*out_sy+out_py, x*out_sx+out_px), coff + n] =
 *       sum_t sum_c src[tap_src[t]](b, y + tap_dy[t], x + tap_dx[t], c) * Wp[n, k(t, c)]
 * with zero outside each source view (padding of the ALREADY activated operand).
 * Wp is the packed bf16 weight matrix [n_rows][ktot] (dmm_pack_weights), k-blocks of `kwidth`
 * channels ordered tap-major, each source padded up to a multiple of kwidth.
 * Replaces: nn.Conv2d 1x1 (tv:38,132; Dense_U_Net_lidar.py:111-112,190-191), 3x3 (tv:42; :126-127),
 * 5x5 (:130-131), 7x7/s2 via im2col (:73-74,157-158), the 4 sub-pixel phases of
 * nn.ConvTranspose2d(C,C,3,stride=2,padding=1) (:117-118) and the data-gradients of all of them.
 * out_mode 0: bf16 NHWC rows of pitch ldo.   out_mode 1: fp32 NCHW (B, n_valid, OH, OW).
 * out_mode 2: fp32 NCHW (B, N / fold_kw, H, W) of a fold_kw-wide convolution with few output channels whose KERNEL COLUMNS are
 *   folded into the GEMM's N: the taps are the kernel ROWS only (dx == 0), weight row kw*C + n holds w[n, :, kh, kw]
 *   (C = N / fold_kw classes, N <= 16), and the epilogue forms out[n](y, x) = sum_kw acc[(y, x + kw - fold_kw/2)][kw*C + n]
 *   from tiles that overlap by fold_kw - 1 columns.  K*K -> K MMAs per tile for the 5x5, 64 -> 3 head convolution (:130-131).
 * out_mode 3: the same fold for the 3x3, 128 -> 32 growth convolutions (tv:51-53): fold_kw = 3, N = n_tile = 96 (weight row
 *   kw*32 + n), tile_w = 32; output = bf16 NHWC slice (ldo, coff) + stats of the 32 real output channels like out_mode 0.
 * stats (nullable): double[DMM_STATS_SLOTS][2][stats_ld], column sums / sums of squares of the
 * bf16-rounded outputs are atomically added at [.., stats_off + n]. */
typedef struct {
    dmm_view_t src[DMM_MAX_SRC];
    int32_t num_src;
    int32_t num_taps;
    int8_t tap_src[DMM_MAX_TAPS];
    int8_t tap_dy[DMM_MAX_TAPS];
    int8_t tap_dx[DMM_MAX_TAPS];
    const void* weights;
    int64_t ktot;
    int32_t n_rows;
    int32_t kwidth;          /* 64 (SWIZZLE_128B) or 16 (SWIZZLE_32B) */
    int32_t W, H, B;         /* tile domain = positions (b,y,x) that are computed */
    int32_t tile_w;          /* 128 / 64 / 32 / 16 / 8; tile_h = 128 / tile_w */
    int32_t N;               /* output channels (any value; computed in tiles of <= n_tile) */
    int32_t n_tile;          /* multiple of 16, <= 256 */
    void* out;
    int32_t out_mode;
    int64_t ldo;
    int32_t coff;
    int32_t out_sy, out_sx, out_py, out_px, OH, OW;
    double* stats;
    int32_t stats_ld, stats_off;
    /* Optional (bnb_sums != NULL, out_mode 0, no output stride/phase, stats == NULL): the output g of this launch is the gradient
     * of relu(bn(x)) (a data-gradient launch); the epilogue then also performs dmm_bn_relu_bwd_reduce for the BatchNorm whose
     * input channel n is column n of bnb_x (same pixels): bnb_sums[slot][0][off + n] += sum_p dz,
     * bnb_sums[slot][1][off + n] += invstd[n] * sum_p dz * (x - mean[n]), dz = g (as stored in bf16) * [gamma*invstd*(x-mean)+beta > 0]. */
    const void* bnb_x;
    int64_t bnb_ldx;
    const float* bnb_gamma;
    const float* bnb_beta;
    const float* bnb_mean;
    const float* bnb_invstd;
    double* bnb_sums;
    int32_t bnb_sums_ld, bnb_sums_off;
    /* Optional BN-ReLU PROLOGUE (pro_enable != 0; one source, stride-1 output): src[0] holds the RAW input x and the kernel
     * applies relu(bn(x)) to every A tile / halo patch in shared memory before the MMAs read it (tv:47-50 without materialising
     * the activated tensor).  With K x K taps the patch pixels outside the image stay zero (the convolution pads the ACTIVATED
     * tensor).  pro_bn describes the BatchNorm over the src[0].C channels exactly like dmm_bn_relu_apply (CTA 0 also saves
     * mean / invstd and updates the running statistics). */
    int32_t pro_enable;
    int32_t fold_kw;         /* out_mode 2 / 3: kernel width folded into N (odd); else 0 */
    dmm_bn_t pro_bn;
    /* Arithmetic / storage type: 0 = bf16 storage, tcgen05.mma kind::f16, fp32 accumulate (everything above).
     * 1 = STRICT mode: fp32 storage of sources, packed weights ([n_rows][ktot] float, kwidth 32) and output, tcgen05.mma
     * kind::tf32 with fp32 accumulate; out_mode 0 (fp32 pixel-major rows) or 1; no prologue / fused BN backward / folds.
     * 2 = like 1 with 3xTF32 products: the tensor core truncates fp32 operands to tf32, so A*B is evaluated as
     * Ahi*Bhi + Ahi*Blo + Alo*Bhi with the exact remainders Alo (computed in shared memory) and Blo (second half of the packed
     * weights, [2][n_rows][ktot]); n_rows must be a multiple of n_tile. */
    int32_t dtype;
    /* Optional epilogue of out_mode 0 (inference: an eval-mode BatchNorm FOLDED into this convolution - its scale goes into
     * the packed weight rows (dmm_pack_job_t.rscale), its shift arrives here): out = (epi_relu ? max(., 0) : .)(acc + epi_bias[n]).
     * epi_bias: float[N], 16-byte aligned; NULL = plain store. */
    int32_t epi_relu;
    const float* epi_bias;
} dmm_igemm_t;
int dmm_conv_igemm(const dmm_igemm_t* d, void* stream);

/* Weight-gradient GEMM on tcgen05 (both operands pixel-major = MN-major UMMA descriptors):
 *   dw[row*ld + col] += sum_{b,y,x} A_i(b, y + dy_i, x + dx_i, ch_i + r) * B_g(b, y + dy_g, x + dx_g, ch_g + c)
 * for every A chunk i (64 channels: rows a[i].out0 + r, r < 64) and B group g (n_tile channels: columns
 * b[g].out0 + c); views are zero outside their extent (pixels and channels).  fp32 atomic accumulation into the
 * scratch matrix dw (caller zero-fills); dmm_unpack_wgrad(_batched) scatters it into the parameter layout.
 * Replaces the weight-gradient half of aten::convolution_backward for every Conv2d / ConvTranspose2d of the model
 * (47% of the reference's CPU time).  A KxK convolution is expressed either with A = the activation
 * and one B group per tap (B = output gradient shifted by -tap), or with the roles swapped; several sources are
 * needed for ConvTranspose2d (the four sub-pixel phases of the output gradient).
 * One CTA owns ceil(num_a/2) x num_b accumulators of 128 x n_tile (<= 512 TMEM columns in total); the grid is
 * (splits of the pixel tiles) x (ya x yb replicas, replica (ia, ib) adds ia*a_step / ib*b_step to the channel AND
 * output offsets of the A / B side). */
#define DMM_WG_MAX_A 8
#define DMM_WG_MAX_B 25
typedef struct {
    int8_t src;              /* index into a_src / b_src */
    int8_t dy, dx;
    int8_t pad_;
    int32_t ch0;             /* first channel inside the source view */
    int32_t out0;            /* first dw row (A chunk) / column (B group) */
} dmm_wg_slot_t;
typedef struct {
    dmm_view_t a_src[DMM_MAX_SRC];
    dmm_view_t b_src[DMM_MAX_SRC];
    int32_t num_a_src, num_b_src;
    dmm_wg_slot_t a[DMM_WG_MAX_A];
    dmm_wg_slot_t b[DMM_WG_MAX_B];
    int32_t num_a, num_b;
    int32_t n_tile;          /* 16, 32, or a multiple of 16 in [64, 256] */
    int32_t ya, yb, a_step, b_step;
    int32_t W, H, B;         /* pixel domain */
    int32_t kpx;             /* pixels per pipeline stage: 64 or 32 */
    int32_t tile_w;          /* power of two, 8 <= tile_w <= kpx; tile_h = kpx / tile_w */
    int32_t splits;          /* pixel-range splits; 0 = auto */
    float* dw;
    int64_t ld;
    /* Optional BN-ReLU PROLOGUE on the A side (pro_enable != 0; every A chunk reads a_src[0] unshifted): a_src[0] holds the RAW
     * BatchNorm input and relu(bn(x)) is applied to the A tiles in shared memory (weight gradient of a convolution whose activated
     * input was never materialised, see dmm_igemm_t.pro_*).  pro_bn: save_mean / save_invstd / gamma / beta of the a_src[0].C channels
     * (training = 0 semantics: the forward pass already fixed the statistics; running_mean / running_var are NOT used).
     * When the B groups are shifted (K x K), A rows outside the image are written as zeros (the padding of the activation). */
    int32_t pro_enable;
    int32_t pad_;
    const float* pro_gamma;
    const float* pro_beta;
    const float* pro_mean;
    const float* pro_invstd;
} dmm_wgrad_t;
int dmm_conv_wgrad(const dmm_wgrad_t* d, void* stream);

/* dst[n][t*Kp + c] (bf16, row length T*Kp, Kp = C rounded up to kwidth) <- w[n*sn + c*sc + tap_off[t]]
 * for n < n_valid, c < C; zero elsewhere (rows up to n_rows).  Converts fp32 parameter tensors
 * (Conv2d (Cout,Cin,kh,kw), ConvTranspose2d (Cin,Cout,kh,kw)) into the K-major B operand of
 * dmm_conv_igemm for forward or data-gradient use. */
int dmm_pack_weights(const float* w, void* dst, int32_t n_valid, int32_t n_rows, int32_t C, int32_t kwidth,
                     int32_t T, const int32_t* tap_off, int64_t sn, int64_t sc, void* stream);
/* grad[n*sn + m*sc + tap_off[t]] (= or +=) dw[t*dt + m*dm + n*dn]: dmm_conv_wgrad scratch -> parameter layout
 * (t < T taps, m < M input channels, n < N output channels of the reference weight tensor). */
int dmm_unpack_wgrad(const float* dw, int64_t dt, int64_t dm, int64_t dn, int32_t M, int32_t N, float* grad, int32_t T,
                     const int32_t* tap_off, int64_t sn, int64_t sc, int32_t accumulate, void* stream);

/* Batched forms of the two calls above: the job tables live in DEVICE memory (built once per
 * model), one launch per training step covers every weight tensor of the network.
 */
typedef struct {
    const float* w;
    void* dst;
    int32_t n_valid, n_rows, C, kwidth, T;
    int32_t tap_off[DMM_MAX_TAPS];
    int64_t sn, sc;          /* source element of (n, c, t): w[nidx + (c / cdiv)*sc + (c % cdiv)*sc2 + tap_off[t]] */
    int64_t sc2;
    int32_t cdiv;            /* 0 or 1: plain channel index */
    int32_t ndiv;            /* 0 or 1: nidx = n*sn; else nidx = (n / ndiv)*sn + (n % ndiv)*sn2 (kernel columns folded into n) */
    int64_t sn2;
    const float* rscale;     /* optional per-row factor (float[n_valid]): packed row n = rscale[n] * w row (a folded eval-mode BatchNorm) */
} dmm_pack_job_t;
typedef struct {
    const float* dw;
    float* grad;
    int64_t dt, dm, dn;
    int32_t M, N, T, accumulate;
    int32_t tap_off[DMM_MAX_TAPS];
    int64_t sn, sc;
    int64_t sn2;             /* ndiv > 1: the n-part of the address is (n % ndiv)*sn + (n / ndiv)*sn2 (kernel columns folded into n) */
    int32_t ndiv;
    int32_t pad_;
} dmm_unpack_job_t;
int dmm_pack_weights_batched(const dmm_pack_job_t* jobs_device, int32_t njobs, void* stream);
int dmm_unpack_wgrad_batched(const dmm_unpack_job_t* jobs_device, int32_t njobs, void* stream);
/* Eval-mode BatchNorm folding (inference extras): scale[c] = gamma[c] / sqrt(running_var[c] + eps),
 * shift[c] = beta[c] - running_mean[c] * scale[c] for every job of a device-resident table, one launch. */
typedef struct {
    const float* gamma;
    const float* beta;
    const float* running_mean;
    const float* running_var;
    float* scale;
    float* shift;
    int32_t C;
    float eps;
} dmm_bn_fold_job_t;
int dmm_bn_fold_batched(const dmm_bn_fold_job_t* jobs_device, int32_t njobs, void* stream);
/* Load-balanced forms: `work_device` = int32 pairs (job index, chunk index) in device memory, one thread block per pair
 * handles elements [chunk*chunk_elems, (chunk+1)*chunk_elems) of that job (the tensors of one network span 64 ... 9.4 M
 * elements; 32 blocks per job left the largest job alone on the GPU for 0.9 ms). */
int dmm_pack_weights_work(const dmm_pack_job_t* jobs_device, const int32_t* work_device, int32_t nwork, int32_t chunk_elems,
                          void* stream);
int dmm_unpack_wgrad_work(const dmm_unpack_job_t* jobs_device, const int32_t* work_device, int32_t nwork, int32_t chunk_elems,
                          void* stream);

/* y = relu(bn(x)), bf16 pixel-major [B*H*W, C] (tv:47-50,90; Dense_U_Net_lidar.py:75-76,108-115).
 * pool: 0 none; 1 = then AvgPool2d(2,2) (tv:133 moved in front of the 1x1 conv - they commute);
 *       2 = then MaxPool2d(3, stride 2, padding 1) (Dense_U_Net_lidar.py:77).
 * ystats (nullable): column sums / sums of squares of the stored OUTPUT are added at ystats_off. */
typedef struct {
    const void* x;
    int64_t ldx;
    int32_t B, H, W, C;
    dmm_bn_t bn;
    int32_t pool;
    void* y;
    int64_t ldy;
    double* ystats;
    int32_t ystats_ld, ystats_off;
    void* argmax;            /* pool == 2 only, nullable: uint8 [B*OH*OW, ldarg] position 0..8 of each window's */
    int64_t ldarg;           /* first maximum, consumed by dmm_bn_relu_bwd_* (gmode 2)                           */
} dmm_bn_apply_t;
int dmm_bn_relu_apply(const dmm_bn_apply_t* d, void* stream);

/* Backward of BN-ReLU (aten::threshold_backward + native_batch_norm_backward), two passes:
 *   dz = g' * [bn(x) > 0], g' = gradient of the activated tensor addressed through gmode
 *        (0: same pixel; 1: avg-pool parent, 0.25*g[parent]; 2: max-pool 3x3/s2/p1, the gradient
 *         of each window goes to its FIRST maximum like ATen);
 *   reduce: sums[slot][2][sums_ld] += (sum dz, sum dz*xhat);
 *   apply : dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)); dgamma = sum dz*xhat, dbeta = sum dz.
 * out_mode 0: store bf16; 1: store fp32; 2: accumulate into fp32.  out may be NULL (only dgamma/dbeta). */
typedef struct {
    double* sums;
    int32_t sums_ld, sums_off;
    double count;
    const float* gamma;
    const float* beta;
    const float* save_mean;
    const float* save_invstd;
    float* dgamma;           /* nullable */
    float* dbeta;
} dmm_bn_bwd_t;
typedef struct {
    const void* x;           /* raw (pre-BN) bf16 input of the forward pass */
    int64_t ldx;
    const void* g;
    int64_t ldg;
    int32_t g_is_f32;
    int32_t gmode;
    int32_t B, H, W, C;      /* geometry of x */
    dmm_bn_bwd_t bn;
    void* out;
    int64_t ldo;
    int32_t out_mode;
    void* dz_out;            /* nullable: the reduce pass also stores dz as bf16 [B*H*W, lddz] (then run the apply */
    int64_t lddz;            /* pass with g = dz_out, gmode = 0: the costly pool routing is done only once)       */
    const void* argmax;      /* gmode 2, nullable: window winners recorded by dmm_bn_relu_apply (else recomputed)  */
    int64_t ldarg;
    int32_t out_gw;          /* dmm_bn_relu_bwd_contrib only: > 0 = "planar" slab, channel group g = c / out_gw is its own */
    int32_t pad_;            /* contiguous [rows, out_gw] matrix at out + g * out_plane (elements); ldo is ignored          */
    int64_t out_plane;
    /* dmm_bn_relu_bwd_contrib only, optional: fuse dmm_bn_bwd_finalize into the launch.  fin_k: float[2][C] correction vectors;
     * fin_ctr: uint32 ticket counters (one per 256-channel group of this launch, >= 8 entries), zeroed by the caller before
     * every launch: the LAST block to finish a channel group reads the completed sums and writes dgamma / dbeta / k. */
    float* fin_k;
    uint32_t* fin_ctr;
} dmm_bn_bwd_args_t;
int dmm_bn_relu_bwd_reduce(const dmm_bn_bwd_args_t* d, void* stream);
int dmm_bn_relu_bwd_apply(const dmm_bn_bwd_args_t* d, void* stream);

/* Dense-block gradient flow without read-modify-write (every dense layer contributes to ALL earlier channels).
 * All BatchNorms that consume a dense-block buffer normalise the same raw channels with the same batch statistics, so
 *   dx = A*dz - A*c1 - A*invstd*c2*(x - mean)         (A = gamma*invstd, c1 = mean(dz), c2 = mean(dz*xhat))
 * splits into a per-pixel part A*dz and a per-channel affine correction that can be summed over the consumers:
 *   dmm_bn_relu_bwd_contrib : ONE pass over (x, g): dz = g*[bn(x) > 0]; sums += (sum dz, invstd*sum dz*(x-mean)) like
 *                             dmm_bn_relu_bwd_reduce, and the bf16 slab out = A*dz is stored (args: x, g (bf16), bn, out/ldo).
 *   dmm_bn_bwd_finalize     : dgamma = sum dz*xhat, dbeta = sum dz, k[0][c] = A*c1, k[1][c] = A*invstd*c2 (k: float[2][C]).
 *   dmm_grad_gather         : out = bf16( sum_s src_s - sum_j k1_j - (x - mean) * sum_j k2_j ): the gradient of a channel range
 *                             of the block buffer from the slabs of all its consumers (those with a fully corrected dx pass no k). */
int dmm_bn_relu_bwd_contrib(const dmm_bn_bwd_args_t* d, void* stream);
int dmm_bn_bwd_finalize(const dmm_bn_bwd_t* bn, int32_t C, float* k, void* stream);
#define DMM_GATHER_MAX 56
typedef struct {
    const void* src[DMM_GATHER_MAX];     /* bf16 rows, already offset to the first channel of the range */
    int64_t ld[DMM_GATHER_MAX];
    int64_t plane[DMM_GATHER_MAX];       /* 0: row-major source; else planar (see dmm_bn_bwd_args_t.out_gw): group stride in elements */
    int32_t gw;                          /* channel-group width of the planar sources (ld = gw for them) */
    int32_t pad0_;
    int32_t nsrc;
    int32_t nk;
    const float* k1[DMM_GATHER_MAX];     /* per-consumer correction vectors, offset to the first channel */
    const float* k2[DMM_GATHER_MAX];
    const void* x;                       /* raw block buffer (bf16), offset to the first channel; NULL if nk == 0 */
    int64_t ldx;
    const float* mean;                   /* batch mean of the channels (any consumer's save_mean slice) */
    int64_t rows;
    int32_t C;                           /* multiple of 8 */
    int32_t pad_;
    void* out;                           /* bf16 [rows, ldo] */
    int64_t ldo;
} dmm_grad_gather_t;
int dmm_grad_gather(const dmm_grad_gather_t* d, void* stream);

/* ---- stem / head data movement ------------------------------------------------------------- */
/* im2col for conv0 (7x7, stride 2, padding 3; Dense_U_Net_lidar.py:73-74,157-158): fp32 NCHW
 * (B,C1[+C2],H,W) -> bf16 [B*OH*OW, kpad], k = ci*49 + kh*7 + kw (the flattened Conv2d weight
 * order), zero padded to kpad.  x2 (nullable) supplies the channels after x1's (early fusion cat,
 * :228-229). */
int dmm_im2col_7x7s2(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t B, int32_t H,
                     int32_t W, void* out, int32_t kpad, void* stream);
/* Stem without the im2col matrix: horizontal unfold only.  out[(b, iy, ox)][kw*C + c] = x[c](iy, 2*ox + kw - 3) as bf16 rows of
 * pitch ld (>= 7*C, zero padded), C = C1 + C2, OW = (W - 1) / 2 + 1, ALL H input rows.  conv0 (7x7, stride 2, padding 3,
 * Dense_U_Net_lidar.py:73-74) is then a 7-tap VERTICAL convolution over the even-row / odd-row views of that tensor
 * (dmm_conv_igemm with two strided sources), 2.5x less data than the [P/4, 49*C] im2col matrix. */
int dmm_unfold_w7s2(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t B, int32_t H, int32_t W, void* out,
                    int64_t ld, void* stream);
/* per-plane sums of an fp32 NCHW tensor (B, C, H*W) -> stats (raw network inputs feeding the head BN). */
int dmm_nchw_stats(const float* x, int32_t B, int32_t C, int64_t HW, double* stats, int32_t stats_ld,
                   int32_t stats_off, void* stream);
/* head input: out = relu(bn0(cat(upsample2x(u), x1, x2))) as bf16 [B*H*W, ldo], zero beyond Cu+C1+C2
 * (nn.Upsample :120, torch.cat :264, dec_out_to_heat_maps.norm0/relu0 :124-125).
 * bn_u: the BN slice of the Cu decoder channels (statistics of u, rep = 4); bn_x: the slice of
 * the C1+C2 raw input channels (statistics from dmm_nchw_stats). */
typedef struct {
    const void* u;
    int64_t ldu;
    int32_t Cu;
    const float* x1;
    int32_t C1;
    const float* x2;
    int32_t C2;
    int32_t B, H, W;         /* full (output) resolution; u is (B, H/2, W/2) */
    dmm_bn_t bn_u, bn_x;
    void* out;
    int64_t ldo;
} dmm_head_t;
int dmm_head_input(const dmm_head_t* d, void* stream);
/* backward of the above: reduce over all channels, apply for the Cu decoder channels
 * (du[parent] = sum of dx over the 2x2 children, bf16). */
typedef struct {
    const void* u;
    int64_t ldu;
    int32_t Cu;
    const float* x1;
    int32_t C1;
    const float* x2;
    int32_t C2;
    int32_t B, H, W;
    const void* g;           /* bf16 [B*H*W, ldg] gradient of the head input */
    int64_t ldg;
    dmm_bn_bwd_t bn_u, bn_x;
    void* du;
    int64_t lddu;
} dmm_head_bwd_t;
int dmm_head_input_bwd_reduce(const dmm_head_bwd_t* d, void* stream);
int dmm_head_input_bwd_apply(const dmm_head_bwd_t* d, void* stream);
/* fp32 NCHW (B,C,H,W) -> bf16 NHWC [B*H*W, ldo] (channels >= C zero) : d(logits) for refine1. */
int dmm_nchw_to_nhwc_bf16(const float* x, int32_t B, int32_t C, int32_t H, int32_t W, void* out,
                          int64_t ldo, void* stream);
/* d(logits) (B,C,H,W) fp32 -> bf16 rows [B*H*W, ld], column t*C + n = dlogits[n](y - (kh - K/2), x - (kw - K/2)),
 * t = kh*K + kw, zero outside the image: turns the weight / data gradients of the KxK, N = num_classes head
 * convolution (refine1, Dense_U_Net_lidar.py:130-131) into plain 1x1 GEMMs over K*K*C columns. */
int dmm_dlogits_im2col(const float* dlogits, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K, void* out,
                       int64_t ld, void* stream);
/* Horizontal-only variant: column kw*C + n = dlogits[n](y, x - (kw - K/2)).  With it the data gradient of the KxK head
 * convolution is a K-tap VERTICAL convolution over K*C (<= 16) channels instead of K*K taps over C channels, and the
 * weight gradient needs K instead of K*K shifted views. */
int dmm_dlogits_unfold_w(const float* dlogits, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K, void* out,
                         int64_t ld, void* stream);
/* fp32 rows [P, C] (pitch lds) -> bf16 rows (pitch ldd): slices of the fp32 dense-block gradient buffer. */
int dmm_rows_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t P, int32_t C,
                         void* stream);

/* ---- loss: torch.nn.BCEWithLogitsLoss(reduction='none') + its gradient for a ones cotangent
 * (Agent:54,247,264): loss = max(x,0) - x*t + log1p(exp(-|x|)); grad = sigmoid(x) - t.
 * loss / grad nullable.  class_sums (nullable): double[C] += per-class loss sums (Agent:248). */
int dmm_bce_logits(const float* logits, const float* target, int64_t n, int32_t C, int64_t HW,
                   float* loss, float* grad, double* class_sums, void* stream);

/* ---- integer scatters (bit-exact) --------------------------------------------------------- */
/* helper:493-515 lidar_array_to_image_like_tensor: points (N,3) float32 [x,y,d] in order ->
 * img (1,H,W) float32, -1 background, KxK splat, LAST point wins.  scratch: int32[H*W]. */
int dmm_lidar_splat(const float* points, int32_t n_points, int32_t H, int32_t W, int32_t kernel_size,
                    int32_t* scratch, float* img, void* stream);
/* helper:446-491 pool_lidar_tensor: range transform, MaxPool2d((20,10), stride 10), replicate
 * pad of one bottom row, negatives -> 0.  (1,H,W) -> (1,(H-20)/10+2,(W-10)/10+1). */
int dmm_lidar_pool(const float* img, int32_t H, int32_t W, float* out, void* stream);
/* helper:233-305 create_ground_truth_maps: boxes int32 (N,5) rows [type,x,y,w,h] in dict order ->
 * maps (3,H,W) float32; last box wins per class channel; pedestrian silhouette template.
 * scratch: int32[3*H*W]. */
int dmm_heatmap_boxes(const int32_t* boxes, int32_t n_boxes, int32_t H, int32_t W, int32_t* scratch,
                      float* maps, void* stream);
/* Batched single-pass forms (BASELINE config 4: per-step on-GPU pre-processing of a whole batch in ONE launch each, every
 * output pixel written once, no scratch image): frame b owns rows offsets[b] .. offsets[b+1] of `points` (float32 [.,3]) /
 * `boxes` (int32 [.,5]); offsets: int32[B+1] in DEVICE memory.  img: (B,1,H,W); mode 0 = the raw image of
 * lidar_array_to_image_like_tensor (-1 background), mode 1 = network input: the range transform of pool_lidar_tensor
 * (helper:472-481) applied at full resolution, negatives -> 0.  maps: (B,3,H,W).  Bit-identical to the per-frame calls. */
int dmm_lidar_splat_batched(const float* points, const int32_t* offsets, int32_t B, int32_t H, int32_t W, int32_t kernel_size,
                            int32_t mode, float* img, void* stream);
int dmm_heatmap_boxes_batched(const int32_t* boxes, const int32_t* offsets, int32_t B, int32_t H, int32_t W, float* maps,
                              void* stream);
/* helper:438-444 maxpool_tensor / helper:430-436 avgpool_tensor: (C,H,W) -> (C,H/k,W/k). */
int dmm_pool_kxk(const float* img, int32_t C, int32_t H, int32_t W, int32_t k, int32_t is_max,
                 float* out, void* stream);

/* ---- step metrics (helper:311-401 compute_IoU_whole_img_per_class / compute_accuracy) ---------------
 * pred, gt: fp32 [planes][HW] (planes = samples * classes); counts: int64 [planes][3] += (|pred>=t & gt>=t|,
 * |pred>=t | gt>=t|, |(pred>=t) == (gt>=t)|).  Caller zero-fills counts. */
int dmm_step_metrics(const float* pred, const float* gt, int32_t planes, int64_t HW, float threshold, int64_t* counts,
                     void* stream);

/* ---- optimiser: fused Adam over a flat fp32 parameter/grad buffer ---------- */
int dmm_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                  void* stream);

/* ---- STRICT arithmetic mode (fp32 storage, tf32 tensor-core convolutions: dmm_igemm_t.dtype == 1) - forward pass ----------
 * fp32-row restatements of dmm_bn_relu_apply / dmm_head_input / dmm_im2col_7x7s2 / dmm_pack_weights_work: same descriptors,
 * every activation pointer / pitch refers to float rows (C, ldx, ldy, ldu, ldo in elements; C % 8 == 0).  Used by
 * Engine(precision="tf32") to measure the distance of the bf16 production path from the reference's fp32 arithmetic
 * (Dense_U_Net_lidar_Agent.py:17,54-61 runs fp32 without AMP). */
int dmm_bn_relu_apply_f32(const dmm_bn_apply_t* d, void* stream);
int dmm_head_input_f32(const dmm_head_t* d, void* stream);
int dmm_im2col_7x7s2_f32(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t B, int32_t H, int32_t W, void* out,
                         int32_t kpad, void* stream);
/* split != 0: every job's packed tensor is [2][n_rows][ktot] floats: the weights and their tf32 remainders (3xTF32, dtype 2) */
int dmm_pack_weights_work_f32(const dmm_pack_job_t* jobs_device, const int32_t* work_device, int32_t nwork, int32_t chunk_elems,
                              int32_t split, void* stream);

/* sizeof() of the structs above in declaration order (0 = dmm_view_t, 1 = dmm_igemm_t, 2 = dmm_wgrad_t,
 * 3 = dmm_bn_t, 4 = dmm_bn_apply_t, 5 = dmm_bn_bwd_t, 6 = dmm_bn_bwd_args_t, 7 = dmm_head_t,
 * 8 = dmm_head_bwd_t, 9 = dmm_pack_job_t, 10 = dmm_unpack_job_t, 11 = dmm_grad_gather_t, 12 = dmm_bn_fold_job_t) so a binding can
 * verify its layout. */
int dmm_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif /* DMMFODS_B200_H */
