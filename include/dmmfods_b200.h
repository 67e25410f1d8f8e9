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
 *   - batch-norm statistics are accumulated in double: stats[slot][2][C] (sum, sum of squares).
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
} dmm_igemm_t;
int dmm_conv_igemm(const dmm_igemm_t* d, void* stream);

/* Weight-gradient GEMM on tcgen05 (both operands pixel-major = MN-major UMMA descriptors):
 *   dw[t][m][n] += sum_{b,y,x} X[xsrc[t]](b, y + dy[t], x + dx[t], m) * Y[ysrc[t]](b, y, x, n)
 * fp32 atomic accumulation into dw (caller zero-fills), layout dw[(t*M + m)*ldw + n].
 * Replaces the weight-gradient half of aten::convolution_backward for every Conv2d /
 * ConvTranspose2d above (47% of the reference's CPU time). */
typedef struct {
    dmm_view_t x[DMM_MAX_SRC];
    dmm_view_t y[DMM_MAX_SRC];
    int32_t num_taps;
    int8_t tap_xsrc[DMM_MAX_TAPS];
    int8_t tap_ysrc[DMM_MAX_TAPS];
    int8_t tap_dy[DMM_MAX_TAPS];
    int8_t tap_dx[DMM_MAX_TAPS];
    int32_t W, H, B;         /* pixel domain (of Y; X is read shifted) */
    int32_t tile_w;          /* 64 / 32 / 16 / 8; tile_h = 64 / tile_w */
    int32_t M, N;            /* channels of X (rows of dw) and of Y (columns of dw) */
    int32_t n_tile;          /* multiple of 16, <= 256 */
    int32_t splits;          /* pixel-chunk splits per (tap, m-tile, n-tile); 0 = auto */
    float* dw;
    int64_t ldw;
} dmm_wgrad_t;
int dmm_conv_wgrad(const dmm_wgrad_t* d, void* stream);

/* dst[n][kb*kwidth + c] (bf16, row length ktot) <- w[n*sn + (c0 + c)*sc + off] for the k-block
 * groups listed; zero padding elsewhere.  Converts fp32 parameter tensors (Conv2d (Cout,Cin,kh,kw),
 * ConvTranspose2d (Cin,Cout,kh,kw)) into the K-major operand of dmm_conv_igemm for fprop or dgrad. */
typedef struct {
    int64_t sn, sc, off;     /* element strides/offset into w for (n, c) */
    int32_t c0, cnt;         /* channel range of this group */
    int32_t nblk;            /* k-blocks occupied = ceil(cnt / kwidth) */
} dmm_pack_group_t;
int dmm_pack_weights(const float* w, void* dst, int32_t n, int32_t n_rows_padded, int64_t ktot,
                     int32_t kwidth, const dmm_pack_group_t* groups, int32_t num_groups, void* stream);
/* grad[n*sn + (c0+c)*sc + off] (fp32, parameter layout) <- or += packed dw[(g*M + m)*ldw + n'] ;
 * inverse of the packing for the fp32 weight-gradient produced by dmm_conv_wgrad.
 * rows_are_n: 1 if dw rows index the parameter's "n" (stride sn) and columns its "c"; 0 if swapped. */
int dmm_unpack_wgrad(const float* dw, int64_t ldw, int32_t M, int32_t N, float* grad,
                     const dmm_pack_group_t* groups, int32_t num_groups, int32_t rows_are_n,
                     int32_t accumulate, void* stream);

/* ---- batch-norm pieces (nn.BatchNorm2d semantics) ------------------------------- */
/* column sums / sums of squares of a [P, C] bf16 view into stats[slot][2][stats_ld] at stats_off. */
int dmm_col_stats(const void* x, int64_t ldx, int64_t P, int32_t C, double* stats, int32_t stats_ld,
                  int32_t stats_off, void* stream);
/* per-plane sums of an fp32 NCHW tensor (B, C, H, W) -> stats (raw network inputs, head BN). */
int dmm_nchw_stats(const float* x, int32_t B, int32_t C, int64_t HW, double* stats, int32_t stats_ld,
                   int32_t stats_off, void* stream);
/* From accumulated stats (all slots are summed): mean, invstd (biased var, eps), then for ONE
 * BatchNorm module over channels [0,C) of that statistics row: scale = gamma*invstd,
 * shift = beta - mean*scale, running_mean/var momentum update with the unbiased variance
 * (count*rep/(count*rep-1); rep = replication factor of nn.Upsample), num_batches_tracked += 1.
 * running_* / nbt may be NULL (no update).  training=0: scale/shift from the running stats. */
int dmm_bn_finalize(const double* stats, int32_t stats_ld, int32_t stats_off, int32_t C, double count,
                    double rep, const float* gamma, const float* beta, float eps, float momentum,
                    float* running_mean, float* running_var, int64_t* nbt, int32_t training,
                    float* mean, float* invstd, float* scale, float* shift, void* stream);

/* y[p, c] = relu(scale[c]*x[p, c] + shift[c])  (BN-ReLU prologue, materialised bf16 operand)
 * pool: 0 none; 1 = then AvgPool2d(2,2) (tv:133 moved in front of the 1x1 conv - they commute);
 *       2 = then MaxPool2d(3, stride 2, padding 1) (Dense_U_Net_lidar.py:77).
 * stats (nullable): column stats of the OUTPUT (needed when the output is a raw feature). */
int dmm_bn_relu_apply(const void* x, int64_t ldx, int32_t B, int32_t H, int32_t W, int32_t C,
                      const float* scale, const float* shift, int32_t pool, void* y, int64_t ldy,
                      double* stats, int32_t stats_ld, int32_t stats_off, void* stream);

/* Backward of BN-ReLU (threshold_backward + native_batch_norm_backward), two passes.
 * The incoming gradient g of the ACTIVATED tensor is addressed through `gmode`:
 *   0: g[p, c] same resolution;  1: avg-pool parent: 0.25 * g[parent(p), c];
 *   2: max-pool 3x3/s2/p1 parents (gradient routed to the first maximum of each window).
 * pass 1 (reduce): sums[slot][2][C] += (sum dz, sum dz*xhat), dz = g * [scale*x+shift > 0].
 * pass 2 (apply): dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat));
 *   out_mode 0: store bf16; 1: accumulate into bf16 (out += dx). */
int dmm_bn_relu_bwd_reduce(const void* x, int64_t ldx, const void* g, int64_t ldg, int32_t gmode,
                           int32_t B, int32_t H, int32_t W, int32_t C, const float* mean,
                           const float* invstd, const float* scale, const float* shift, double* sums,
                           int32_t sums_ld, int32_t sums_off, void* stream);
int dmm_bn_bwd_finalize(const double* sums, int32_t sums_ld, int32_t sums_off, int32_t C, double count,
                        const float* gamma, const float* invstd, float* dgamma, float* dbeta,
                        float* c1, float* c2, int32_t accumulate, void* stream);
int dmm_bn_relu_bwd_apply(const void* x, int64_t ldx, const void* g, int64_t ldg, int32_t gmode,
                          int32_t B, int32_t H, int32_t W, int32_t C, const float* mean,
                          const float* invstd, const float* scale, const float* shift,
                          const float* c1, const float* c2, void* out, int64_t ldo, int32_t out_mode,
                          void* stream);

/* ---- stem / head data movement ------------------------------------------------------------- */
/* im2col for conv0 (7x7, stride 2, padding 3): fp32 NCHW (B,Cin,H,W) -> bf16 [B*OH*OW, kpad],
 * k = (kh*7 + kw)*Cin + ci, zero padded to kpad.  c_off/c_cnt select input channels; a second
 * tensor x2 (nullable) supplies channels after x1's (early fusion cat, :228-229). */
int dmm_im2col_7x7s2(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t B, int32_t H,
                     int32_t W, void* out, int32_t kpad, void* stream);
/* head input: a0 = relu(bn0(cat(upsample2x(u), x1, x2))) as bf16 [B*H*W, ldo]
 * (nn.Upsample :120, torch.cat :264, dec_out_to_heat_maps.norm0/relu0 :124-125). */
int dmm_head_input(const void* u, int64_t ldu, int32_t Cu, const float* x1, int32_t C1,
                   const float* x2, int32_t C2, int32_t B, int32_t H, int32_t W, const float* scale,
                   const float* shift, void* out, int64_t ldo, void* stream);
/* backward of the above for the Cu up-sampled channels (sum over the 2x2 children), two passes as
 * for dmm_bn_relu_bwd_*; channels >= Cu (raw inputs) only contribute to the BN sums. */
int dmm_head_input_bwd_reduce(const void* u, int64_t ldu, int32_t Cu, const float* x1, int32_t C1,
                              const float* x2, int32_t C2, const void* g, int64_t ldg, int32_t B,
                              int32_t H, int32_t W, const float* mean, const float* invstd,
                              const float* scale, const float* shift, double* sums, int32_t sums_ld,
                              void* stream);
int dmm_head_input_bwd_apply(const void* u, int64_t ldu, int32_t Cu, const void* g, int64_t ldg,
                             int32_t B, int32_t H, int32_t W, const float* mean, const float* invstd,
                             const float* scale, const float* shift, const float* c1, const float* c2,
                             void* du, int64_t lddu, void* stream);
/* fp32 NCHW (B,C,H,W) -> bf16 NHWC [B*H*W, ldo] (channels >= C zero) : d(logits) for refine1. */
int dmm_nchw_to_nhwc_bf16(const float* x, int32_t B, int32_t C, int32_t H, int32_t W, void* out,
                          int64_t ldo, void* stream);

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
 * maps (3,H,W) float32; last box wins per class channel; pedestrian silhouette template. */
int dmm_heatmap_boxes(const int32_t* boxes, int32_t n_boxes, int32_t H, int32_t W, float* maps,
                      void* stream);
/* helper:438-444 maxpool_tensor / helper:430-436 avgpool_tensor: (C,H,W) -> (C,H/k,W/k). */
int dmm_pool_kxk(const float* img, int32_t C, int32_t H, int32_t W, int32_t k, int32_t is_max,
                 float* out, void* stream);

/* ---- optimiser: fused Adam over a flat fp32 parameter/grad buffer ---------- */
int dmm_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMMFODS_B200_H */
