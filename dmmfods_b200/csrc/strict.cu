// STRICT arithmetic modes of the forward pass (BASELINE north_star: "rel 1e-3 fp32/tf32"): fp32 STORAGE of every activation
// and of the packed weights, tcgen05.mma kind::tf32 with fp32 accumulation for the convolutions (igemm.cu, dmm_igemm_t.dtype 1;
// dtype 2 = 3xTF32: operands split into tf32 head + exact remainder, three MMAs per product - fp32-grade products), fp64
// BatchNorm statistics.  This file holds the HBM-bound forward kernels of that mode: they restate the bf16 kernels of
// elementwise.cu on fp32 rows (same C-ABI descriptors, `_f32` entry points) and favour clarity over the last GB/s - the mode
// exists to measure how far the bf16 production path is from the reference's fp32 arithmetic, not to be the fast path.
//   dmm_bn_relu_apply_f32   y = relu(bn(x)) with optional 2x2 avg-pool / 3x3-s2 max-pool of the activated tensor + output statistics
//   dmm_head_input_f32      nearest x2 upsample + cat(rgb, lidar) + norm0 + ReLU (Dense_U_Net_lidar.py:120-132, :264)
//   dmm_im2col_7x7s2_f32    stem im2col (7x7, stride 2, pad 3) of the fp32 NCHW network inputs
//   dmm_pack_weights_work_f32  parameter layout -> K-major fp32 GEMM operand (kwidth 32)
#include "common.cuh"
#include "../../include/dmmfods_b200.h"

namespace dmm {

constexpr int kSThreads = 256;

__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// thread (tx = 8-channel chunk, ty = row lane); blockIdx.y walks chunk groups, blockIdx.x strides over output rows
template <int POOL>
__global__ void __launch_bounds__(kSThreads) bn_relu_apply_f32_kernel(const dmm_bn_apply_t p, int OH, int OW) {
    __shared__ float cf[2][kSThreads];
    __shared__ double red[2][kSThreads];
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    const bool active = chunk < nchunks;
    const int tid = threadIdx.y * cx + threadIdx.x;
    {
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            const BnCoef k = bn_coef_fwd(p.bn, c, blockIdx.x == 0);
            cf[0][tid] = k.scale;
            cf[1][tid] = k.shift;
        }
    }
    __syncthreads();
    float sc[8], sh[8];
    double s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = active ? cf[0][threadIdx.x * 8 + j] : 0.f;
        sh[j] = active ? cf[1][threadIdx.x * 8 + j] : 0.f;
        s1[j] = s2[j] = 0.0;
    }
    const float* x = reinterpret_cast<const float*>(p.x);
    float* y = reinterpret_cast<float*>(p.y);
    const long long rows = (long long)p.B * OH * OW;
    if (active) {
        for (long long row = (long long)blockIdx.x * ry + threadIdx.y; row < rows; row += (long long)gridDim.x * ry) {
            float o[8];
            if (POOL == 0) {
                float f[8];
                ld8f(x + row * p.ldx + chunk * 8, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
            } else {
                const int ox = (int)(row % OW);
                const long long t = row / OW;
                const int oy = (int)(t % OH);
                const int b = (int)(t / OH);
                if (POOL == 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = 0.f;
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx) {
                            float f[8];
                            ld8f(x + (((long long)b * p.H + (2 * oy + dy)) * p.W + (2 * ox + dx)) * p.ldx + chunk * 8, f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] += fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
                        }
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] *= 0.25f;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = -INFINITY;
                    for (int t9 = 0; t9 < 9; ++t9) {
                        const int iy = 2 * oy + t9 / 3 - 1, ix = 2 * ox + t9 % 3 - 1;
                        if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
                        float f[8];
                        ld8f(x + (((long long)b * p.H + iy) * p.W + ix) * p.ldx + chunk * 8, f);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f));
                    }
                }
            }
            st8f(y + row * p.ldy + chunk * 8, o);
            if (p.ystats) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s1[j] += (double)o[j];
                    s2[j] += (double)o[j] * (double)o[j];
                }
            }
        }
    }
    if (p.ystats) {
        // per-channel block reduction over the ry row lanes, then one double atomic per channel and block
        const int slot = blockIdx.x % DMM_STATS_SLOTS;
        for (int j = 0; j < 8; ++j) {
            __syncthreads();
            red[0][tid] = s1[j];
            red[1][tid] = s2[j];
            __syncthreads();
            if (threadIdx.y == 0 && active) {
                double a = 0.0, b = 0.0;
                for (int r = 0; r < ry; ++r) {
                    a += red[0][r * cx + threadIdx.x];
                    b += red[1][r * cx + threadIdx.x];
                }
                atomicAdd(p.ystats + ((size_t)slot * 2 + 0) * p.ystats_ld + p.ystats_off + chunk * 8 + j, a);
                atomicAdd(p.ystats + ((size_t)slot * 2 + 1) * p.ystats_ld + p.ystats_off + chunk * 8 + j, b);
            }
        }
    }
}

// one thread per (pixel, 8-column chunk of the output row)
__global__ void __launch_bounds__(kSThreads) head_input_f32_kernel(const dmm_head_t p) {
    extern __shared__ float coef[];   // [2][Cpad]
    const int Ct = p.Cu + p.C1 + p.C2;
    const int chunks = (int)(p.ldo >> 3);
    const int Cpad = chunks * 8;
    for (int c = threadIdx.x; c < Cpad; c += blockDim.x) {
        float sc = 0.f, sh = 0.f;
        if (c < Ct) {
            const bool writer = blockIdx.x == 0;
            const BnCoef k = c < p.Cu ? bn_coef_fwd(p.bn_u, c, writer) : bn_coef_fwd(p.bn_x, c - p.Cu, writer);
            sc = k.scale;
            sh = k.shift;
        }
        coef[c] = sc;
        coef[Cpad + c] = sh;
    }
    __syncthreads();
    const int UH = p.H >> 1, UW = p.W >> 1;
    const long long HW = (long long)p.H * p.W;
    const long long total = (long long)p.B * HW * chunks;
    const float* u = reinterpret_cast<const float*>(p.u);
    float* out = reinterpret_cast<float*>(p.out);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long long pix = i / chunks;
        const int xx = (int)(pix % p.W);
        const long long t = pix / p.W;
        const int yy = (int)(t % p.H);
        const int b = (int)(t / p.H);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            float v = 0.f;
            if (c < p.Cu) v = __ldg(u + (((long long)b * UH + (yy >> 1)) * UW + (xx >> 1)) * p.ldu + c);
            else if (c < p.Cu + p.C1) v = __ldg(p.x1 + ((long long)b * p.C1 + (c - p.Cu)) * HW + (long long)yy * p.W + xx);
            else if (c < Ct) v = __ldg(p.x2 + ((long long)b * p.C2 + (c - p.Cu - p.C1)) * HW + (long long)yy * p.W + xx);
            f[j] = c < Ct ? fmaxf(fmaf(v, coef[c], coef[Cpad + c]), 0.f) : 0.f;
        }
        st8f(out + pix * p.ldo + ch * 8, f);
    }
}

// out[(b, oy, ox)][ci*49 + kh*7 + kw] = x[ci](2 oy + kh - 3, 2 ox + kw - 3), zero outside the image and for k >= C*49
__global__ void __launch_bounds__(kSThreads) im2col_7x7s2_f32_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2,
                                                                     int B, int H, int W, int OH, int OW, float* __restrict__ out, int kpad) {
    const int K = (C1 + C2) * 49;
    const long long total = (long long)B * OH * OW * kpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % kpad);
        const long long pix = i / kpad;
        float v = 0.f;
        if (k < K) {
            const int ox = (int)(pix % OW);
            const long long t = pix / OW;
            const int oy = (int)(t % OH);
            const int b = (int)(t / OH);
            const int ci = k / 49, tp = k - ci * 49;
            const int kh = tp / 7, kw = tp - kh * 7;
            const int iy = 2 * oy + kh - 3, ix = 2 * ox + kw - 3;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                v = ci < C1 ? __ldg(x1 + (((long long)b * C1 + ci) * H + iy) * W + ix)
                            : __ldg(x2 + (((long long)b * C2 + (ci - C1)) * H + iy) * W + ix);
        }
        out[i] = v;
    }
}

// split != 0 (3xTF32): the packed tensor is [2][n_rows][ktot]: w itself (the tensor core reads its upper 19 bits = trunc_tf32(w))
// and the exactly representable remainder w - trunc_tf32(w)
__global__ void __launch_bounds__(256) pack_weights_work_f32_kernel(const dmm_pack_job_t* __restrict__ jobs, const int2* __restrict__ work,
                                                                    int chunk_elems, int split) {
    const int2 w = work[blockIdx.x];
    const dmm_pack_job_t& j = jobs[w.x];
    const int Kp = (j.C + j.kwidth - 1) / j.kwidth * j.kwidth;
    const long long ktot = (long long)Kp * j.T;
    const long long total = (long long)j.n_rows * ktot;
    const long long i0 = (long long)w.y * chunk_elems;
    const long long i1 = i0 + chunk_elems < total ? i0 + chunk_elems : total;
    float* dst = reinterpret_cast<float*>(j.dst);
    const int cdiv = j.cdiv > 0 ? j.cdiv : 1;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const int n = (int)(i / ktot);
        const int k = (int)(i - (long long)n * ktot);
        const int t = k / Kp;
        const int c = k - t * Kp;
        float v = 0.f;
        if (n < j.n_valid && c < j.C) {
            const long long nidx = j.ndiv > 1 ? (long long)(n / j.ndiv) * j.sn + (long long)(n % j.ndiv) * j.sn2 : (long long)n * j.sn;
            v = __ldg(j.w + nidx + (long long)(c / cdiv) * j.sc + (long long)(c % cdiv) * j.sc2 + j.tap_off[t]);
        }
        dst[i] = v;
        if (split) dst[total + i] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
}

}  // namespace dmm

using namespace dmm;

extern "C" int dmm_bn_relu_apply_f32(const dmm_bn_apply_t* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DMM_CHECK(d && d->x && d->y, "dmm_bn_relu_apply_f32: null pointer");
    DMM_CHECK(d->C > 0 && d->C % 8 == 0 && d->ldx % 4 == 0 && d->ldy % 4 == 0, "dmm_bn_relu_apply_f32: C=%d must be a positive multiple of 8, pitches multiples of 4", d->C);
    DMM_CHECK(d->pool >= 0 && d->pool <= 2, "dmm_bn_relu_apply_f32: pool=%d", d->pool);
    DMM_CHECK(d->bn.training ? (d->bn.stats != nullptr && d->bn.count > 0) : (d->bn.running_mean && d->bn.running_var), "dmm_bn_relu_apply_f32: BatchNorm without statistics");
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;
    int OH = d->H, OW = d->W;
    if (d->pool == 1) { DMM_CHECK(d->H % 2 == 0 && d->W % 2 == 0, "dmm_bn_relu_apply_f32: avg-pool needs even H, W"); OH = d->H / 2; OW = d->W / 2; }
    if (d->pool == 2) { OH = (d->H - 1) / 2 + 1; OW = (d->W - 1) / 2 + 1; }
    const int chunks = d->C / 8;
    int cx = 1;
    while (cx < chunks && cx < 32) cx <<= 1;
    const int ry = kSThreads / cx;
    const long long rows = (long long)d->B * OH * OW;
    const int gy = (chunks + cx - 1) / cx;
    long long gx = (rows + ry - 1) / ry;
    const long long cap = 148ll * 8 / gy > 0 ? 148ll * 8 / gy : 1;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, (unsigned)gy, 1), block((unsigned)cx, (unsigned)ry, 1);
    if (d->pool == 0) bn_relu_apply_f32_kernel<0><<<grid, block, 0, stream>>>(*d, OH, OW);
    else if (d->pool == 1) bn_relu_apply_f32_kernel<1><<<grid, block, 0, stream>>>(*d, OH, OW);
    else bn_relu_apply_f32_kernel<2><<<grid, block, 0, stream>>>(*d, OH, OW);
    DMM_LAUNCH_CHECK("bn_relu_apply_f32_kernel");
    return 0;
}

extern "C" int dmm_head_input_f32(const dmm_head_t* d, void* stream) {
    DMM_CHECK(d && d->u && d->x1 && d->out && (d->C2 == 0 || d->x2), "dmm_head_input_f32: null pointer");
    DMM_CHECK(d->Cu % 8 == 0 && d->ldo % 8 == 0 && d->ldo >= d->Cu + d->C1 + d->C2 && d->H % 2 == 0 && d->W % 2 == 0,
              "dmm_head_input_f32: Cu / ldo must be multiples of 8, H and W even");
    if (d->B <= 0) return 0;
    const size_t smem = 2 * (size_t)d->ldo * sizeof(float);
    DMM_CHECK(smem <= 48 * 1024, "dmm_head_input_f32: too many channels");
    head_input_f32_kernel<<<148 * 8, kSThreads, smem, (cudaStream_t)stream>>>(*d);
    DMM_LAUNCH_CHECK("head_input_f32_kernel");
    return 0;
}

extern "C" int dmm_im2col_7x7s2_f32(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t B, int32_t H, int32_t W, void* out,
                                    int32_t kpad, void* stream) {
    DMM_CHECK(x1 && out && C1 >= 1 && C2 >= 0 && (C2 == 0 || x2) && kpad >= (C1 + C2) * 49 && kpad % 4 == 0, "dmm_im2col_7x7s2_f32: bad arguments");
    if (B <= 0) return 0;
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    im2col_7x7s2_f32_kernel<<<148 * 16, kSThreads, 0, (cudaStream_t)stream>>>(x1, C1, x2, C2, B, H, W, OH, OW, reinterpret_cast<float*>(out), kpad);
    DMM_LAUNCH_CHECK("im2col_7x7s2_f32_kernel");
    return 0;
}

extern "C" int dmm_pack_weights_work_f32(const dmm_pack_job_t* jobs_device, const int32_t* work_device, int32_t nwork, int32_t chunk_elems,
                                         int32_t split, void* stream) {
    DMM_CHECK(nwork >= 0 && (nwork == 0 || (jobs_device && work_device)) && chunk_elems >= 256, "dmm_pack_weights_work_f32: bad arguments");
    if (nwork == 0) return 0;
    pack_weights_work_f32_kernel<<<(unsigned)nwork, 256, 0, (cudaStream_t)stream>>>(jobs_device, reinterpret_cast<const int2*>(work_device), chunk_elems, split);
    DMM_LAUNCH_CHECK("pack_weights_work_f32_kernel");
    return 0;
}
