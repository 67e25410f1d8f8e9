// HBM-bound kernels of the Dense-U-Net hot path: BN-ReLU (+pool) prologue materialisation with the
// batch-statistics finalize fused in, BN-ReLU backward (two passes), stem im2col, head input
// (nearest x2 upsample + concat + BN-ReLU), sigmoid-BCE loss + gradient, weight packing, Adam.
//
// Layout: activations are pixel-major bf16 matrices [P rows = B*H*W, ld elements per row]; one thread
// moves 8 channels (16 bytes) of one pixel; a block is (cx channel-chunks) x (ry pixels) = 256
// threads so that a warp touches contiguous 16*cx-byte row segments.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/dmmfods_b200.h"

namespace dmm {

constexpr int kEwThreads = 256;
constexpr int kMaxBlocksPerSm = 8;
constexpr int kNumSm = 148;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
    f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
    f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]);
    v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]);
    v.w = pack_bf16x2(f[6], f[7]);
    return v;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ BnCoef bn_coef_bwd(const dmm_bn_bwd_t& bn, int c) {
    BnCoef k;
    k.mean = bn.save_mean[c];
    k.invstd = bn.save_invstd[c];
    const float g = bn.gamma ? bn.gamma[c] : 1.f;
    const float b = bn.beta ? bn.beta[c] : 0.f;
    k.scale = g * k.invstd;
    k.shift = b - k.mean * k.scale;
    return k;
}

struct ColCfg {
    int cx, ry, chunks;
    dim3 grid, block;
};
static int env_int_ew(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// bps: grid cap in blocks per SM (a multiple of the kernel's resident blocks per SM keeps the grid-stride waves full)
static ColCfg col_cfg(int C, long long rows, int bps = kMaxBlocksPerSm) {
    static const int bps_env = env_int_ew("DMM_EW_BPS", 0);
    if (bps_env > 0) bps = bps_env;
    ColCfg k;
    k.chunks = (C + 7) / 8;
    k.cx = 1;
    while (k.cx < k.chunks && k.cx < 32) k.cx <<= 1;
    k.ry = kEwThreads / k.cx;
    const int gy = (k.chunks + k.cx - 1) / k.cx;
    long long gx = (rows + k.ry - 1) / k.ry;
    long long cap = (long long)kNumSm * bps / gy;
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    k.grid = dim3((unsigned)gx, (unsigned)gy, 1);
    k.block = dim3((unsigned)k.cx, (unsigned)k.ry, 1);
    return k;
}

// Block-wide column reduction of per-thread 8-channel partials, then double atomics into
// stats[slot][2][ld] at off + channel.   sm: float[2][ry][cx*8].
__device__ __forceinline__ void block_col_reduce_atomic(float (&a)[8], float (&b)[8], float* sm, int cx, int ry,
                                                        int chunk, int nchunks, double* stats, int ld, int off) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int wcols = cx * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sm[(0 * ry + ty) * wcols + tx * 8 + j] = a[j];
        sm[(1 * ry + ty) * wcols + tx * 8 + j] = b[j];
    }
    __syncthreads();
    const int tid = ty * cx + tx;
    const int slot = blockIdx.x % DMM_STATS_SLOTS;
    for (int i = tid; i < 2 * wcols; i += kEwThreads) {
        const int which = i / wcols, col = i - which * wcols;
        const int ch = (chunk - tx) * 8 + col;   // (chunk - tx) = first chunk of this block
        if (ch < nchunks * 8) {
            float s = 0.f;
            for (int r = 0; r < ry; ++r) s += sm[(which * ry + r) * wcols + col];
            atomicAdd(stats + ((size_t)slot * 2 + which) * ld + off + ch, (double)s);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// y = relu(bn(x)) with optional pooling of the activated tensor.
// ---------------------------------------------------------------------------------------------
template <int POOL>
__global__ void __launch_bounds__(kEwThreads, 3) bn_relu_apply_kernel(const dmm_bn_apply_t p, int OH, int OW) {
    pdl_prologue();
    __shared__ float sm[2 * kEwThreads * 8];
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    const bool active = chunk < nchunks;
    __shared__ float cf[2][kEwThreads];
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            BnCoef k = bn_coef_fwd(p.bn, c, blockIdx.x == 0);
            cf[0][tid] = k.scale;
            cf[1][tid] = k.shift;
        }
    }
    __syncthreads();
    float sc[8], sh[8];
    if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = cf[0][threadIdx.x * 8 + j];
            sh[j] = cf[1][threadIdx.x * 8 + j];
        }
    }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x);
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y);
    const long long rows = (long long)p.B * OH * OW;
    if (active) {
        for (long long row = (long long)blockIdx.x * ry + threadIdx.y; row < rows; row += (long long)gridDim.x * ry) {
            float o[8];
            if (POOL == 0) {
                float f[8];
                unpack8(ldg16(x + row * p.ldx + chunk * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
            } else {
                const unsigned r32 = (unsigned)row;      // host guarantees B*OH*OW < 2^31 for the pooled modes
                const int ox = (int)(r32 % (unsigned)OW);
                const unsigned t = r32 / (unsigned)OW;
                const int oy = (int)(t % (unsigned)OH);
                const int b = (int)(t / (unsigned)OH);
                if (POOL == 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 2; ++dx) {
                            const long long r = ((long long)b * p.H + (2 * oy + dy)) * p.W + (2 * ox + dx);
                            float f[8];
                            unpack8(ldg16(x + r * p.ldx + chunk * 8), f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] += fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
                        }
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] *= 0.25f;
                } else {
                    uint32_t code[8];      // position (dy+1)*3 + (dx+1) of the FIRST maximum (ATen's max_pool2d rule)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        o[j] = -INFINITY;
                        code[j] = 0;
                    }
                    // one window row (3 independent loads) at a time: all nine in flight cost 128 registers = 2 blocks per SM
                    // (ncu r02: 24 % of the warps active, issue-bound at 59 %)
#pragma unroll
                    for (int wy = 0; wy < 3; ++wy) {
                        uint4 win[3];
                        bool ok[3];
                        const int iy = 2 * oy + wy - 1;
#pragma unroll
                        for (int wx = 0; wx < 3; ++wx) {
                            const int ix = 2 * ox + wx - 1;
                            ok[wx] = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                            if (ok[wx]) win[wx] = ldg16(x + (((long long)b * p.H + iy) * p.W + ix) * p.ldx + chunk * 8);
                        }
#pragma unroll
                        for (int wx = 0; wx < 3; ++wx) {
                            if (ok[wx]) {
                                float f[8];
                                unpack8(win[wx], f);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float a = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
                                    if (a > o[j]) {
                                        o[j] = a;
                                        code[j] = (uint32_t)(wy * 3 + wx);
                                    }
                                }
                            }
                        }
                    }
                    if (p.argmax) {
                        uint2 packed;
                        packed.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
                        packed.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
                        *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(p.argmax) + row * p.ldarg + chunk * 8) = packed;
                    }
                }
            }
            const uint4 packed = pack8(o);
            *reinterpret_cast<uint4*>(y + row * p.ldy + chunk * 8) = packed;
            if (p.ystats) {
                float r[8];
                unpack8(packed, r);   // statistics of the stored (bf16-rounded) values
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s1[j] += r[j];
                    s2[j] += r[j] * r[j];
                }
            }
        }
    }
    if (p.ystats) block_col_reduce_atomic(s1, s2, sm, cx, ry, chunk, nchunks, p.ystats, p.ystats_ld, p.ystats_off);
}

// ---------------------------------------------------------------------------------------------
// Lean y = relu(bn(x)) (no pooling, no output statistics): four rows per iteration = four independent 16-byte loads in
// flight per thread, coefficients in registers.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads, 4) bn_apply_fast_kernel(const dmm_bn_apply_t p) {
    pdl_prologue();
    __shared__ float cf[2][kEwThreads];
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            BnCoef k = bn_coef_fwd(p.bn, c, blockIdx.x == 0);
            cf[0][tid] = k.scale;
            cf[1][tid] = k.shift;
        }
    }
    __syncthreads();
    if (chunk >= nchunks) return;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = cf[0][threadIdx.x * 8 + j];
        sh[j] = cf[1][threadIdx.x * 8 + j];
    }
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x) + chunk * 8;
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + chunk * 8;
    const long long rows = (long long)p.B * p.H * p.W;
    const long long step = (long long)gridDim.x * ry;
    for (long long row0 = (long long)blockIdx.x * ry + threadIdx.y; row0 < rows; row0 += 4 * step) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (row0 + u * step < rows) v[u] = ldg16(x + (row0 + u * step) * p.ldx);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (row0 + u * step < rows) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
                *reinterpret_cast<uint4*>(y + (row0 + u * step) * p.ldy) = pack8(f);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// BN-ReLU backward.  dz = g' * [bn(x) > 0] where g' is the gradient of the activated tensor,
// addressed through gmode (0 same pixel, 1 avg-pool parent / 4, 2 max-pool 3x3 s2 p1 argmax).
// ---------------------------------------------------------------------------------------------
template <typename GT>
__device__ __forceinline__ void load_g8(const GT* g, long long idx, float (&f)[8]);
template <>
__device__ __forceinline__ void load_g8<__nv_bfloat16>(const __nv_bfloat16* g, long long idx, float (&f)[8]) {
    unpack8(ldg16(g + idx), f);
}
template <>
__device__ __forceinline__ void load_g8<float>(const float* g, long long idx, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(g + idx));
    const float4 b = __ldg(reinterpret_cast<const float4*>(g + idx + 4));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <int GMODE, typename GT>
__device__ __forceinline__ void bn_relu_dz(const dmm_bn_bwd_args_t& p, const __nv_bfloat16* x, const GT* g,
                                           long long row, int chunk, const float (&sc)[8], const float (&sh)[8],
                                           int OH, int OW, float (&xr)[8], float (&dz)[8]) {
    unpack8(ldg16(x + row * p.ldx + chunk * 8), xr);
    float z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = fmaf(xr[j], sc[j], sh[j]);
    if (GMODE == 0) {
        load_g8<GT>(g, row * p.ldg + chunk * 8, dz);
    } else {
        const unsigned r32 = (unsigned)row;          // host guarantees B*H*W < 2^31 for the pooled modes
        const int xx = (int)(r32 % (unsigned)p.W);
        const unsigned t = r32 / (unsigned)p.W;
        const int yy = (int)(t % (unsigned)p.H);
        const int b = (int)(t / (unsigned)p.H);
        if (GMODE == 1) {
            const int oy = yy >> 1, ox = xx >> 1;
            if (oy < OH && ox < OW) {
                load_g8<GT>(g, (((long long)b * OH + oy) * OW + ox) * p.ldg + chunk * 8, dz);
#pragma unroll
                for (int j = 0; j < 8; ++j) dz[j] *= 0.25f;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) dz[j] = 0.f;
            }
        } else {
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[j] = fmaxf(z[j], 0.f);
                dz[j] = 0.f;
            }
            const int oy0 = yy >> 1, oy1 = (yy + 1) >> 1;   // windows whose rows 2o-1..2o+1 contain yy
            const int ox0 = xx >> 1, ox1 = (xx + 1) >> 1;
            if (p.argmax) {
                // the forward pass recorded the winner of every window: all (up to four) windows are fetched at once
                const int ny = (oy1 != oy0 && oy1 < OH) ? 2 : 1, nx = (ox1 != ox0 && ox1 < OW) ? 2 : 1;
                uint2 cd[4];
                float gg[4][8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int a = q >> 1, bb = q & 1;
                    if (a < ny && bb < nx) {
                        const long long w = ((long long)b * OH + (a ? oy1 : oy0)) * OW + (bb ? ox1 : ox0);
                        cd[q] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(p.argmax) + w * p.ldarg + chunk * 8));
                        load_g8<GT>(g, w * p.ldg + chunk * 8, gg[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int a = q >> 1, bb = q & 1;
                    if (a < ny && bb < nx) {
                        const int oy = a ? oy1 : oy0, ox = bb ? ox1 : ox0;
                        const uint32_t mine = (uint32_t)((yy - 2 * oy + 1) * 3 + (xx - 2 * ox + 1));
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t cj = ((j < 4 ? cd[q].x : cd[q].y) >> (8 * (j & 3))) & 0xffu;
                            if (cj == mine) dz[j] += gg[q][j];
                        }
                    }
                }
            } else
            for (int oy = oy0; oy <= oy1; ++oy) {
                if (oy >= OH) continue;
                for (int ox = ox0; ox <= ox1; ++ox) {
                    if (ox >= OW) continue;
                    if (p.argmax) {     // forward pass recorded the winner of every window: no neighbour scan
                        const long long w = ((long long)b * OH + oy) * OW + ox;
                        const uint2 cd = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(p.argmax) +
                                                                              w * p.ldarg + chunk * 8));
                        const uint32_t mine = (uint32_t)((yy - 2 * oy + 1) * 3 + (xx - 2 * ox + 1));
                        float gg[8];
                        load_g8<GT>(g, w * p.ldg + chunk * 8, gg);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t cj = ((j < 4 ? cd.x : cd.y) >> (8 * (j & 3))) & 0xffu;
                            if (cj == mine) dz[j] += gg[j];
                        }
                        continue;
                    }
                    bool arg[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) arg[j] = true;
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int iy = 2 * oy + dy;
                        if (iy < 0 || iy >= p.H) continue;
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int ix = 2 * ox + dx;
                            if (ix < 0 || ix >= p.W) continue;
                            if (iy == yy && ix == xx) continue;
                            const bool before = (iy < yy) || (iy == yy && ix < xx);
                            float f[8];
                            unpack8(ldg16(x + (((long long)b * p.H + iy) * p.W + ix) * p.ldx + chunk * 8), f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float an = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
                                if (before ? (an >= a[j]) : (an > a[j])) arg[j] = false;   // first maximum wins
                            }
                        }
                    }
                    float gg[8];
                    load_g8<GT>(g, (((long long)b * OH + oy) * OW + ox) * p.ldg + chunk * 8, gg);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (arg[j]) dz[j] += gg[j];
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = z[j] > 0.f ? dz[j] : 0.f;
}

template <int GMODE, typename GT>
__global__ void __launch_bounds__(kEwThreads) bn_relu_bwd_reduce_kernel(const dmm_bn_bwd_args_t p, int OH, int OW) {
    pdl_prologue();
    __shared__ float sm[2 * kEwThreads * 8];
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    const bool active = chunk < nchunks;
    __shared__ float cf[4][kEwThreads];
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            BnCoef k = bn_coef_bwd(p.bn, c);
            cf[0][tid] = k.scale; cf[1][tid] = k.shift; cf[2][tid] = k.mean; cf[3][tid] = k.invstd;
        }
    }
    __syncthreads();
    float sc[8], sh[8], mu[8], is[8];
    if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int t = threadIdx.x * 8 + j;
            sc[j] = cf[0][t]; sh[j] = cf[1][t]; mu[j] = cf[2][t]; is[j] = cf[3][t];
        }
    }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x);
    const GT* g = reinterpret_cast<const GT*>(p.g);
    const long long rows = (long long)p.B * p.H * p.W;
    if (active) {
        for (long long row = (long long)blockIdx.x * ry + threadIdx.y; row < rows; row += (long long)gridDim.x * ry) {
            float xr[8], dz[8];
            bn_relu_dz<GMODE, GT>(p, x, g, row, chunk, sc, sh, OH, OW, xr, dz);
            if (p.dz_out) {   // keep the masked, pool-routed gradient so that the apply pass can run with gmode 0
                const uint4 packed = pack8(dz);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dz_out) + row * p.lddz + chunk * 8) = packed;
                unpack8(packed, dz);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s1[j] += dz[j];
                s2[j] += dz[j] * ((xr[j] - mu[j]) * is[j]);
            }
        }
    }
    block_col_reduce_atomic(s1, s2, sm, cx, ry, chunk, nchunks, p.bn.sums, p.bn.sums_ld, p.bn.sums_off);
}

// ---------------------------------------------------------------------------------------------
// Lean reduce pass behind max_pool2d(3, 2, 1) with recorded window winners (the stem, Dense_U_Net_lidar.py:75-77): a thread owns
// one 8-channel chunk of a 2x2 input QUAD.  The quad (qy, qx) is touched by exactly the windows (qy..qy+1, qx..qx+1), so four
// (gradient, winner-code) loads serve four input pixels (the per-pixel kernel fetches up to four windows per pixel).  Writes
// dz (masked, pool-routed gradient) for the apply pass and accumulates the BatchNorm sums.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads, 2) bn_bwd_maxpool_quad_kernel(const dmm_bn_bwd_args_t p, int OH, int OW) {
    pdl_prologue();
    __shared__ float sm[2 * kEwThreads * 8];
    __shared__ __align__(16) float cf[4][kEwThreads];
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    const bool active = chunk < nchunks;
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            BnCoef k = bn_coef_bwd(p.bn, c);
            cf[0][tid] = k.scale; cf[1][tid] = k.shift; cf[2][tid] = k.mean; cf[3][tid] = k.invstd;
        }
    }
    __syncthreads();
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    if (active) {
        const int t8 = threadIdx.x * 8;
        float sc[8], sh[8], mu[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = cf[0][t8 + j]; sh[j] = cf[1][t8 + j]; mu[j] = cf[2][t8 + j];
        }
        const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x) + chunk * 8;
        const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(p.g) + chunk * 8;
        const uint8_t* am = reinterpret_cast<const uint8_t*>(p.argmax) + chunk * 8;
        __nv_bfloat16* dzo = reinterpret_cast<__nv_bfloat16*>(p.dz_out) + chunk * 8;
        const int QH = (p.H + 1) >> 1, QW = (p.W + 1) >> 1;
        const unsigned quads = (unsigned)p.B * (unsigned)QH * (unsigned)QW;
        for (unsigned q = blockIdx.x * ry + threadIdx.y; q < quads; q += gridDim.x * ry) {
            const int qx = (int)(q % (unsigned)QW);
            const unsigned t = q / (unsigned)QW;
            const int qy = (int)(t % (unsigned)QH);
            const int b = (int)(t / (unsigned)QH);
            // pixel i = 2*a + c of the quad is (2qy + a, 2qx + c); window w = 2*a + c is (qy + a, qx + c)
            bool pok[4], wok[4];
            uint4 xr[4], gr[4];
            uint2 cd[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int a = i >> 1, c = i & 1;
                pok[i] = (2 * qy + a < p.H) && (2 * qx + c < p.W);
                wok[i] = (qy + a < OH) && (qx + c < OW);
                if (pok[i]) xr[i] = ldg16(x + (((long long)b * p.H + 2 * qy + a) * p.W + 2 * qx + c) * p.ldx);
                if (wok[i]) {
                    const long long w = ((long long)b * OH + qy + a) * OW + qx + c;
                    gr[i] = ldg16(g + w * p.ldg);
                    cd[i] = __ldg(reinterpret_cast<const uint2*>(am + w * p.ldarg));
                }
            }
            float gw[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (wok[i]) unpack8(gr[i], gw[i]);
                else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) gw[i][j] = 0.f;
                    cd[i] = make_uint2(0xffffffffu, 0xffffffffu);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!pok[i]) continue;
                const int a = i >> 1, c = i & 1;
                float xa[8], dz[8];
                unpack8(xr[i], xa);
#pragma unroll
                for (int j = 0; j < 8; ++j) dz[j] = 0.f;
                // windows containing this pixel: rows qy (always) and qy+1 (odd pixel rows), same for columns; the winner code
                // of pixel (iy, ix) in window (oy, ox) is (iy - 2 oy + 1) * 3 + (ix - 2 ox + 1)
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int wa = w >> 1, wc = w & 1;
                    if (wa > a || wc > c) continue;               // compile-time after unrolling
                    const uint32_t mine = (uint32_t)((a - 2 * wa + 1) * 3 + (c - 2 * wc + 1));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t cj = ((j < 4 ? cd[w].x : cd[w].y) >> (8 * (j & 3))) & 0xffu;
                        if (cj == mine) dz[j] += gw[w][j];
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) dz[j] = fmaf(xa[j], sc[j], sh[j]) > 0.f ? dz[j] : 0.f;
                const uint4 packed = pack8(dz);
                *reinterpret_cast<uint4*>(dzo + (((long long)b * p.H + 2 * qy + a) * p.W + 2 * qx + c) * p.lddz) = packed;
                unpack8(packed, dz);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s1[j] += dz[j];
                    s2[j] = fmaf(dz[j], xa[j] - mu[j], s2[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s2[j] *= cf[3][threadIdx.x * 8 + j];
    }
    block_col_reduce_atomic(s1, s2, sm, cx, ry, chunk, nchunks, p.bn.sums, p.bn.sums_ld, p.bn.sums_off);
}

template <int GMODE, typename GT>
__global__ void __launch_bounds__(kEwThreads) bn_relu_bwd_apply_kernel(const dmm_bn_bwd_args_t p, int OH, int OW) {
    pdl_prologue();
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    __shared__ float cf[7][kEwThreads];
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            BnCoef k = bn_coef_bwd(p.bn, c);
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int s = 0; s < DMM_STATS_SLOTS; ++s) {
                const double* r = p.bn.sums + (size_t)s * 2 * p.bn.sums_ld + p.bn.sums_off + c;
                a += r[0];
                b += r[p.bn.sums_ld];
            }
            cf[0][tid] = k.scale; cf[1][tid] = k.shift; cf[2][tid] = k.mean; cf[3][tid] = k.invstd;
            cf[4][tid] = (float)(a / p.bn.count);
            cf[5][tid] = (float)(b / p.bn.count);
            cf[6][tid] = (p.bn.gamma ? p.bn.gamma[c] : 1.f) * k.invstd;
            if (blockIdx.x == 0) {
                if (p.bn.dgamma) p.bn.dgamma[c] = (float)b;
                if (p.bn.dbeta) p.bn.dbeta[c] = (float)a;
            }
        }
    }
    __syncthreads();
    if (chunk >= nchunks) return;
    float sc[8], sh[8], mu[8], is[8], c1[8], c2[8], gi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int t = threadIdx.x * 8 + j;
        sc[j] = cf[0][t]; sh[j] = cf[1][t]; mu[j] = cf[2][t]; is[j] = cf[3][t];
        c1[j] = cf[4][t]; c2[j] = cf[5][t]; gi[j] = cf[6][t];
    }
    if (p.out == nullptr) return;
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x);
    const GT* g = reinterpret_cast<const GT*>(p.g);
    const long long rows = (long long)p.B * p.H * p.W;
    for (long long row = (long long)blockIdx.x * ry + threadIdx.y; row < rows; row += (long long)gridDim.x * ry) {
        float xr[8], dz[8], dx[8];
        bn_relu_dz<GMODE, GT>(p, x, g, row, chunk, sc, sh, OH, OW, xr, dz);
#pragma unroll
        for (int j = 0; j < 8; ++j) dx[j] = gi[j] * (dz[j] - c1[j] - (xr[j] - mu[j]) * is[j] * c2[j]);
        const long long o = row * p.ldo + chunk * 8;
        if (p.out_mode == 0) {
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = pack8(dx);
        } else {
            float* of = reinterpret_cast<float*>(p.out) + o;
            float4 a = make_float4(dx[0], dx[1], dx[2], dx[3]);
            float4 b = make_float4(dx[4], dx[5], dx[6], dx[7]);
            if (p.out_mode == 2) {
                const float4 pa = *reinterpret_cast<const float4*>(of);
                const float4 pb = *reinterpret_cast<const float4*>(of + 4);
                a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
                b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
            }
            *reinterpret_cast<float4*>(of) = a;
            *reinterpret_cast<float4*>(of + 4) = b;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Lean same-pixel (gmode 0) variants: per-channel coefficients in registers, four rows per iteration (8 independent
// 16-byte loads in flight per thread), two 256-thread blocks per SM and a grid of exactly one wave - measured 1.4x faster
// than three blocks x two rows with the coefficients re-read from shared memory (scripts/bench_bn.py).
//   reduce: s1 = sum dz, s2 = invstd * sum dz*(x - mean)
//   apply : dx = A*dz + B*x + C with A = gamma*invstd, B = -A*invstd*c2, C = -A*c1 - B*mean
// ---------------------------------------------------------------------------------------------
// raw (still packed) 8-channel gradient word(s) of one row: kept packed while several rows are in flight
template <typename GT>
struct RawG8;
template <>
struct RawG8<__nv_bfloat16> {
    uint4 a;
    __device__ __forceinline__ void load(const __nv_bfloat16* g, long long idx) { a = ldg16(g + idx); }
    __device__ __forceinline__ void unpack(float (&f)[8]) const { unpack8(a, f); }
};
template <>
struct RawG8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* g, long long idx) {
        a = __ldg(reinterpret_cast<const float4*>(g + idx));
        b = __ldg(reinterpret_cast<const float4*>(g + idx + 4));
    }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
constexpr int kFastRows = 4;      // rows in flight per thread of the lean kernels below
constexpr int kFastBps = 2;       // their resident blocks per SM = their grid cap (one full wave, measured best)

template <typename GT>
__global__ void __launch_bounds__(kEwThreads, kFastBps) bn_bwd_reduce_fast_kernel(const dmm_bn_bwd_args_t p) {
    pdl_prologue();
    __shared__ float sm[2 * kEwThreads * 8];
    __shared__ __align__(16) float cf[4][kEwThreads];
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    const bool active = chunk < nchunks;
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            BnCoef k = bn_coef_bwd(p.bn, c);
            cf[0][tid] = k.scale; cf[1][tid] = k.shift; cf[2][tid] = k.mean; cf[3][tid] = k.invstd;
        }
    }
    __syncthreads();
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    if (active) {
        const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x) + chunk * 8;
        const GT* g = reinterpret_cast<const GT*>(p.g) + chunk * 8;
        const long long rows = (long long)p.B * p.H * p.W;
        const long long step = (long long)gridDim.x * ry;
        const int t8 = threadIdx.x * 8;
        float sc[8], sh[8], mu[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = cf[0][t8 + j]; sh[j] = cf[1][t8 + j]; mu[j] = cf[2][t8 + j];
        }
        for (long long row0 = (long long)blockIdx.x * ry + threadIdx.y; row0 < rows; row0 += kFastRows * step) {
            uint4 xr[kFastRows];
            RawG8<GT> gr[kFastRows];
#pragma unroll
            for (int u = 0; u < kFastRows; ++u) {
                const long long row = row0 + u * step;
                if (row < rows) {
                    xr[u] = ldg16(x + row * p.ldx);
                    gr[u].load(g, row * p.ldg);
                }
            }
#pragma unroll
            for (int u = 0; u < kFastRows; ++u) {
                if (row0 + u * step < rows) {
                    float xa[8], ga[8];
                    unpack8(xr[u], xa);
                    gr[u].unpack(ga);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float dz = fmaf(xa[j], sc[j], sh[j]) > 0.f ? ga[j] : 0.f;
                        s1[j] += dz;
                        s2[j] = fmaf(dz, xa[j] - mu[j], s2[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s2[j] *= cf[3][t8 + j];
    }
    block_col_reduce_atomic(s1, s2, sm, cx, ry, chunk, nchunks, p.bn.sums, p.bn.sums_ld, p.bn.sums_off);
}

template <typename GT, int OUT_MODE>
__global__ void __launch_bounds__(kEwThreads, kFastBps) bn_bwd_apply_fast_kernel(const dmm_bn_bwd_args_t p) {
    pdl_prologue();
    __shared__ __align__(16) float cf[5][kEwThreads];
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            BnCoef k = bn_coef_bwd(p.bn, c);
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int s = 0; s < DMM_STATS_SLOTS; ++s) {
                const double* r = p.bn.sums + (size_t)s * 2 * p.bn.sums_ld + p.bn.sums_off + c;
                a += r[0];
                b += r[p.bn.sums_ld];
            }
            const float c1 = (float)(a / p.bn.count), c2 = (float)(b / p.bn.count);
            const float A = (p.bn.gamma ? p.bn.gamma[c] : 1.f) * k.invstd;
            const float Bc = -A * k.invstd * c2;
            cf[0][tid] = k.scale; cf[1][tid] = k.shift; cf[2][tid] = A; cf[3][tid] = Bc; cf[4][tid] = -A * c1 - Bc * k.mean;
            if (blockIdx.x == 0) {
                if (p.bn.dgamma) p.bn.dgamma[c] = (float)b;
                if (p.bn.dbeta) p.bn.dbeta[c] = (float)a;
            }
        }
    }
    __syncthreads();
    if (chunk >= nchunks || p.out == nullptr) return;
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x) + chunk * 8;
    const GT* g = reinterpret_cast<const GT*>(p.g) + chunk * 8;
    const long long rows = (long long)p.B * p.H * p.W;
    const long long step = (long long)gridDim.x * ry;
    const int t8 = threadIdx.x * 8;
    float sc[8], sh[8], cA[8], cB[8], cC[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = cf[0][t8 + j]; sh[j] = cf[1][t8 + j]; cA[j] = cf[2][t8 + j]; cB[j] = cf[3][t8 + j]; cC[j] = cf[4][t8 + j];
    }
    constexpr int R = (OUT_MODE == 2 || sizeof(GT) == 4) ? 2 : kFastRows;      // fp32 streams: fewer rows, same bytes in flight
    for (long long row0 = (long long)blockIdx.x * ry + threadIdx.y; row0 < rows; row0 += R * step) {
        uint4 xr[R];
        RawG8<GT> gr[R];
        float4 pa[R], pb[R];
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const long long row = row0 + u * step;
            if (row < rows) {
                xr[u] = ldg16(x + row * p.ldx);
                gr[u].load(g, row * p.ldg);
                if (OUT_MODE == 2) {
                    const float* of = reinterpret_cast<const float*>(p.out) + row * p.ldo + chunk * 8;
                    pa[u] = *reinterpret_cast<const float4*>(of);
                    pb[u] = *reinterpret_cast<const float4*>(of + 4);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const long long row = row0 + u * step;
            if (row < rows) {
                float xa[8], ga[8], dx[8];
                unpack8(xr[u], xa);
                gr[u].unpack(ga);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float dz = fmaf(xa[j], sc[j], sh[j]) > 0.f ? ga[j] : 0.f;
                    dx[j] = fmaf(cA[j], dz, fmaf(cB[j], xa[j], cC[j]));
                }
                const long long o = row * p.ldo + chunk * 8;
                if (OUT_MODE == 0) {
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = pack8(dx);
                } else {
                    float* of = reinterpret_cast<float*>(p.out) + o;
                    float4 a = make_float4(dx[0], dx[1], dx[2], dx[3]);
                    float4 b = make_float4(dx[4], dx[5], dx[6], dx[7]);
                    if (OUT_MODE == 2) {
                        a.x += pa[u].x; a.y += pa[u].y; a.z += pa[u].z; a.w += pa[u].w;
                        b.x += pb[u].x; b.y += pb[u].y; b.z += pb[u].z; b.w += pb[u].w;
                    }
                    *reinterpret_cast<float4*>(of) = a;
                    *reinterpret_cast<float4*>(of + 4) = b;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Single-pass BN-ReLU backward "contribution" (see include/dmmfods_b200.h): sums like the reduce pass + bf16 slab A*dz.
// ---------------------------------------------------------------------------------------------
template <int ROWS, bool HOIST>
__global__ void __launch_bounds__(kEwThreads, kFastBps) bn_bwd_contrib_kernel(const dmm_bn_bwd_args_t p) {
    pdl_prologue();
    __shared__ float sm[2 * kEwThreads * 8];
    __shared__ __align__(16) float cf[5][kEwThreads];
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    const bool active = chunk < nchunks;
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            BnCoef k = bn_coef_bwd(p.bn, c);
            cf[0][tid] = k.scale; cf[1][tid] = k.shift; cf[2][tid] = k.mean; cf[3][tid] = k.invstd;
            cf[4][tid] = (p.bn.gamma ? p.bn.gamma[c] : 1.f) * k.invstd;
        }
    }
    __syncthreads();
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    if (active) {
        const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x) + chunk * 8;
        const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(p.g) + chunk * 8;
        // slab layout: row-major [rows, ldo], or planar = one contiguous [rows, gw] matrix per channel group (the gathers of
        // the block then read whole cache lines instead of 64-byte pieces of wide rows)
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
        long long ldo = p.ldo;
        if (p.out_gw > 0) {
            const int c = chunk * 8;
            out += (long long)(c / p.out_gw) * p.out_plane + (c % p.out_gw);
            ldo = p.out_gw;
        } else {
            out += chunk * 8;
        }
        const long long rows = (long long)p.B * p.H * p.W;
        const long long step = (long long)gridDim.x * ry;
        const int t8 = threadIdx.x * 8;
        float hsc[8], hsh[8], hmu[8], hA[8];
        if (HOIST) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                hsc[j] = cf[0][t8 + j]; hsh[j] = cf[1][t8 + j]; hmu[j] = cf[2][t8 + j]; hA[j] = cf[4][t8 + j];
            }
        }
        for (long long row0 = (long long)blockIdx.x * ry + threadIdx.y; row0 < rows; row0 += ROWS * step) {
            // ROWS rows in flight; the raw 16-byte words stay packed until their row is processed
            uint4 xr[ROWS], gr[ROWS];
#pragma unroll
            for (int u = 0; u < ROWS; ++u) {
                const long long row = row0 + u * step;
                if (row < rows) {
                    xr[u] = ldg16(x + row * p.ldx);
                    gr[u] = ldg16(g + row * p.ldg);
                }
            }
#pragma unroll
            for (int u = 0; u < ROWS; ++u) {
                const long long row = row0 + u * step;
                if (row < rows) {
                    float xa[8], ga[8], oa[8];
                    unpack8(xr[u], xa);
                    unpack8(gr[u], ga);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float sc = HOIST ? hsc[j] : cf[0][t8 + j], sh = HOIST ? hsh[j] : cf[1][t8 + j];
                        const float mu = HOIST ? hmu[j] : cf[2][t8 + j], A = HOIST ? hA[j] : cf[4][t8 + j];
                        const float dz = fmaf(xa[j], sc, sh) > 0.f ? ga[j] : 0.f;
                        s1[j] += dz;
                        s2[j] = fmaf(dz, xa[j] - mu, s2[j]);
                        oa[j] = A * dz;
                    }
                    *reinterpret_cast<uint4*>(out + row * ldo) = pack8(oa);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s2[j] *= cf[3][threadIdx.x * 8 + j];
    }
    block_col_reduce_atomic(s1, s2, sm, cx, ry, chunk, nchunks, p.bn.sums, p.bn.sums_ld, p.bn.sums_off);
    if (p.fin_k != nullptr) {
        // fused dmm_bn_bwd_finalize: the last block of this channel group (ticket counter) sees every block's atomics
        __shared__ int last;
        __threadfence();
        __syncthreads();
        const int tid = threadIdx.y * cx + threadIdx.x;
        if (tid == 0) last = atomicAdd(p.fin_ctr + blockIdx.y, 1u) == gridDim.x - 1;
        __syncthreads();
        if (last) {
            __threadfence();
            const int c = blockIdx.y * cx * 8 + tid;
            if (tid < cx * 8 && c < p.C) {
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int s = 0; s < DMM_STATS_SLOTS; ++s) {
                    const double* r = p.bn.sums + (size_t)s * 2 * p.bn.sums_ld + p.bn.sums_off + c;
                    a += __ldcg(r);
                    b += __ldcg(r + p.bn.sums_ld);
                }
                const float invstd = p.bn.save_invstd[c];
                const float A = (p.bn.gamma ? p.bn.gamma[c] : 1.f) * invstd;
                if (p.bn.dgamma) p.bn.dgamma[c] = (float)b;
                if (p.bn.dbeta) p.bn.dbeta[c] = (float)a;
                p.fin_k[c] = A * (float)(a / p.bn.count);
                p.fin_k[p.C + c] = A * invstd * (float)(b / p.bn.count);
            }
        }
    }
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const dmm_bn_bwd_t bn, int C, float* __restrict__ k) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int s = 0; s < DMM_STATS_SLOTS; ++s) {
        const double* r = bn.sums + (size_t)s * 2 * bn.sums_ld + bn.sums_off + c;
        a += r[0];
        b += r[bn.sums_ld];
    }
    const float invstd = bn.save_invstd[c];
    const float A = (bn.gamma ? bn.gamma[c] : 1.f) * invstd;
    if (bn.dgamma) bn.dgamma[c] = (float)b;
    if (bn.dbeta) bn.dbeta[c] = (float)a;
    k[c] = A * (float)(a / bn.count);
    k[C + c] = A * invstd * (float)(b / bn.count);
}

// R rows x S sources = 16 independent 16-byte loads in flight per thread (R chosen from the source count at launch)
template <int R, int S>
__global__ void __launch_bounds__(kEwThreads, kFastBps) grad_gather_kernel(const dmm_grad_gather_t p) {
    pdl_prologue();
    __shared__ __align__(16) float ks[3][kEwThreads];       // sum k1, sum k2, mean of the block's channels
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            // eight consumers per batch: sixteen independent loads in flight (a plain loop is one memory round trip per consumer)
            float a = 0.f, b = 0.f;
            for (int j0 = 0; j0 < p.nk; j0 += 8) {
                float va[8], vb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = j0 + u < p.nk;
                    va[u] = ok ? __ldg(p.k1[j0 + u] + c) : 0.f;
                    vb[u] = ok ? __ldg(p.k2[j0 + u] + c) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) { a += va[u]; b += vb[u]; }
            }
            ks[0][tid] = a; ks[1][tid] = b; ks[2][tid] = p.nk ? __ldg(p.mean + c) : 0.f;
        }
    }
    __syncthreads();
    if (chunk >= nchunks) return;
    const int t8 = threadIdx.x * 8;
    const long long step = (long long)gridDim.x * ry;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + chunk * 8;
    const __nv_bfloat16* xin = reinterpret_cast<const __nv_bfloat16*>(p.x) + chunk * 8;
    // element offset of this thread's chunk inside a planar source (group plane + position in the group)
    const int cg = p.gw > 0 ? (chunk * 8) / p.gw : 0, cw = p.gw > 0 ? (chunk * 8) % p.gw : 0;
#define DMM_GSRC(S, ROW) (reinterpret_cast<const __nv_bfloat16*>(p.src[S]) + (p.plane[S] ? (long long)cg * p.plane[S] + cw : (long long)chunk * 8) + (ROW) * p.ld[S])
    for (long long row0 = (long long)blockIdx.x * ry + threadIdx.y; row0 < p.rows; row0 += R * step) {
        float acc[R][8];
        uint4 xv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
            if (p.nk && row0 + r * step < p.rows) xv[r] = ldg16(xin + (row0 + r * step) * p.ldx);
        }
        for (int s0 = 0; s0 < p.nsrc; s0 += S) {
            const int n = p.nsrc - s0;
            uint4 v[R][S];
#pragma unroll
            for (int u = 0; u < S; ++u) {
                if (u < n) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (row0 + r * step < p.rows) v[r][u] = ldg16(DMM_GSRC(s0 + u, row0 + r * step));
                }
            }
#pragma unroll
            for (int u = 0; u < S; ++u) {
                if (u < n) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (row0 + r * step < p.rows) {
                            float f[8];
                            unpack8(v[r][u], f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[r][j] += f[j];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (row0 + r * step < p.rows) {
                if (p.nk) {
                    float xf[8];
                    unpack8(xv[r], xf);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[r][j] -= fmaf(xf[j] - ks[2][t8 + j], ks[1][t8 + j], ks[0][t8 + j]);
                }
                *reinterpret_cast<uint4*>(out + (row0 + r * step) * p.ldo) = pack8(acc[r]);
            }
        }
    }
#undef DMM_GSRC
}

// ---------------------------------------------------------------------------------------------
// Lean avg-pool-parent (gmode 1, bf16 gradient, even H and W) variants: one thread per (pooled pixel, 8-channel chunk),
// the parent gradient is loaded once and the four children are independent 16-byte loads; blockIdx.x walks pooled rows,
// so there is no per-element division.  PASS 0 = reduce, PASS 1 = apply.
// ---------------------------------------------------------------------------------------------
template <int PASS, int OUT_MODE>
__global__ void __launch_bounds__(kEwThreads, 3) bn_bwd_pool_fast_kernel(const dmm_bn_bwd_args_t p, int OH, int OW) {
    pdl_prologue();
    __shared__ float sm[PASS == 0 ? 2 * kEwThreads * 8 : 1];
    __shared__ __align__(16) float cf[5][kEwThreads];
    const int cx = blockDim.x, ry = blockDim.y;
    const int chunk = blockIdx.y * cx + threadIdx.x;
    const int nchunks = p.C >> 3;
    const bool active = chunk < nchunks;
    {
        const int tid = threadIdx.y * cx + threadIdx.x;
        const int c = blockIdx.y * cx * 8 + tid;
        if (tid < cx * 8 && c < p.C) {
            BnCoef k = bn_coef_bwd(p.bn, c);
            if (PASS == 0) {
                cf[0][tid] = k.scale; cf[1][tid] = k.shift; cf[2][tid] = k.mean; cf[3][tid] = k.invstd;
            } else {
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int s = 0; s < DMM_STATS_SLOTS; ++s) {
                    const double* r = p.bn.sums + (size_t)s * 2 * p.bn.sums_ld + p.bn.sums_off + c;
                    a += r[0];
                    b += r[p.bn.sums_ld];
                }
                const float c1 = (float)(a / p.bn.count), c2 = (float)(b / p.bn.count);
                const float A = (p.bn.gamma ? p.bn.gamma[c] : 1.f) * k.invstd;
                const float Bc = -A * k.invstd * c2;
                cf[0][tid] = k.scale; cf[1][tid] = k.shift; cf[2][tid] = A; cf[3][tid] = Bc; cf[4][tid] = -A * c1 - Bc * k.mean;
                if (blockIdx.x == 0) {
                    if (p.bn.dgamma) p.bn.dgamma[c] = (float)b;
                    if (p.bn.dbeta) p.bn.dbeta[c] = (float)a;
                }
            }
        }
    }
    __syncthreads();
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const int t8 = threadIdx.x * 8;
    if (active && (PASS == 0 || p.out != nullptr)) {
        const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x) + chunk * 8;
        const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(p.g) + chunk * 8;
        const int prow_total = p.B * OH;
        for (int pr = blockIdx.x; pr < prow_total; pr += gridDim.x) {
            const int b = pr / OH, oy = pr - b * OH;
            const long long xrow0 = ((long long)b * p.H + 2 * oy) * p.W;
            const long long grow = (long long)pr * OW;
            for (int ox = threadIdx.y; ox < OW; ox += ry) {
                float gv[8], xv[4][8];
                unpack8(ldg16(g + (grow + ox) * p.ldg), gv);
                const long long r00 = xrow0 + 2 * ox;
                const long long rr[4] = {r00, r00 + 1, r00 + p.W, r00 + p.W + 1};
#pragma unroll
                for (int u = 0; u < 4; ++u) unpack8(ldg16(x + rr[u] * p.ldx), xv[u]);
                float4 pa[4], pb[4];
                if (PASS == 1 && OUT_MODE == 2) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float* of = reinterpret_cast<const float*>(p.out) + rr[u] * p.ldo + chunk * 8;
                        pa[u] = *reinterpret_cast<const float4*>(of);
                        pb[u] = *reinterpret_cast<const float4*>(of + 4);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float dx[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float xx = xv[u][j];
                        const float dz = fmaf(xx, cf[0][t8 + j], cf[1][t8 + j]) > 0.f ? 0.25f * gv[j] : 0.f;
                        if (PASS == 0) {
                            s1[j] += dz;
                            s2[j] = fmaf(dz, xx - cf[2][t8 + j], s2[j]);
                        } else {
                            dx[j] = fmaf(cf[2][t8 + j], dz, fmaf(cf[3][t8 + j], xx, cf[4][t8 + j]));
                        }
                    }
                    if (PASS == 1) {
                        const long long o = rr[u] * p.ldo + chunk * 8;
                        if (OUT_MODE == 0) {
                            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = pack8(dx);
                        } else {
                            float* of = reinterpret_cast<float*>(p.out) + o;
                            float4 a = make_float4(dx[0], dx[1], dx[2], dx[3]);
                            float4 bq = make_float4(dx[4], dx[5], dx[6], dx[7]);
                            if (OUT_MODE == 2) {
                                a.x += pa[u].x; a.y += pa[u].y; a.z += pa[u].z; a.w += pa[u].w;
                                bq.x += pb[u].x; bq.y += pb[u].y; bq.z += pb[u].z; bq.w += pb[u].w;
                            }
                            *reinterpret_cast<float4*>(of) = a;
                            *reinterpret_cast<float4*>(of + 4) = bq;
                        }
                    }
                }
            }
        }
        if (PASS == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s2[j] *= cf[3][t8 + j];
        }
    }
    if (PASS == 0) block_col_reduce_atomic(s1, s2, sm, cx, ry, chunk, nchunks, p.bn.sums, p.bn.sums_ld, p.bn.sums_off);
}

// ---------------------------------------------------------------------------------------------
// stem im2col: fp32 NCHW (B,C1[+C2],H,W) -> bf16 [B*OH*OW, kpad], k = ci*49 + kh*7 + kw
// (same k order as the flattened Conv2d weight (Cout, Cin*7*7)), zero padded.
// ---------------------------------------------------------------------------------------------
// One block = one output row segment of kIm2colPx pixels: the 7 input rows x (2*px+5) columns x C channels it needs are
// staged in shared memory with coalesced loads, a k -> patch-offset table replaces the per-element divisions, and every
// thread then emits 16-byte chunks of im2col rows.
constexpr int kIm2colPx = 64;
__global__ void __launch_bounds__(256) im2col_7x7s2_kernel(const float* __restrict__ x1, int C1,
                                                           const float* __restrict__ x2, int C2, int B, int H, int W,
                                                           int OH, int OW, __nv_bfloat16* __restrict__ out, int kpad) {
    pdl_prologue();
    extern __shared__ float im_sm[];                 // [C][7][PW] patch, then int koff[kpad]
    const int C = C1 + C2;
    const int PW = 2 * kIm2colPx + 5;
    float* patch = im_sm;
    int* koff = reinterpret_cast<int*>(im_sm + C * 7 * PW);
    const int K = C * 49;
    const int segs = (OW + kIm2colPx - 1) / kIm2colPx;
    const int seg = blockIdx.x % segs;
    const int row = blockIdx.x / segs;               // b * OH + oy
    const int oy = row % OH, b = row / OH;
    const int ox0 = seg * kIm2colPx;
    const int ix0 = 2 * ox0 - 3, iy0 = 2 * oy - 3;
    for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
        int o = -1;
        if (k < K) {
            const int ci = k / 49, tp = k - ci * 49;
            const int kh = tp / 7, kw = tp - kh * 7;
            o = (ci * 7 + kh) * PW + kw;
        }
        koff[k] = o;
    }
    // one warp per patch row (channel, kernel row): five independent coalesced loads per lane, no divisions in the loop
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int r = warp; r < C * 7; r += (int)(blockDim.x >> 5)) {
            const int ci = r / 7, kh = r - ci * 7;
            const int iy = iy0 + kh;
            const bool yok = iy >= 0 && iy < H;
            const float* src = ci < C1 ? x1 + (((long long)b * C1 + ci) * H + iy) * W : x2 + (((long long)b * C2 + (ci - C1)) * H + iy) * W;
            float v[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const int px = lane + 32 * u, ix = ix0 + px;
                v[u] = (yok && px < PW && ix >= 0 && ix < W) ? __ldg(src + ix) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 5; ++u)
                if (lane + 32 * u < PW) patch[r * PW + lane + 32 * u] = v[u];
        }
    }
    __syncthreads();
    const int chunks = kpad >> 3;
    const int npx = min(kIm2colPx, OW - ox0);
    __nv_bfloat16* orow = out + ((long long)row * OW + ox0) * kpad;
    for (int i = threadIdx.x; i < npx * chunks; i += blockDim.x) {
        const int px = i / chunks, ch = i - px * chunks;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int o = koff[ch * 8 + j];
            f[j] = o >= 0 ? patch[o + 2 * px] : 0.f;
        }
        *reinterpret_cast<uint4*>(orow + (long long)px * kpad + ch * 8) = pack8(f);
    }
}

// ---------------------------------------------------------------------------------------------
// per-plane statistics of fp32 NCHW tensors (raw network inputs feeding the head BN).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_stats_kernel(const float* __restrict__ x, int C, long long HW,
                                                         double* stats, int ld, int off, int splits) {
    pdl_prologue();
    __shared__ double sm[2][8];
    const int plane = blockIdx.x / splits;   // b*C + c
    const int sp = blockIdx.x - plane * splits;
    const int c = plane % C;
    const float* p = x + (long long)plane * HW;
    const long long lo = HW * sp / splits, hi = HW * (sp + 1) / splits;
    float s1 = 0.f, s2 = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float v = __ldg(p + i);
        s1 += v;
        s2 += v * v;
    }
    double d1 = s1, d2 = s2;
    for (int o = 16; o > 0; o >>= 1) {
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sm[0][threadIdx.x >> 5] = d1;
        sm[1][threadIdx.x >> 5] = d2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) {
            a += sm[0][w];
            b += sm[1][w];
        }
        const int slot = blockIdx.x % DMM_STATS_SLOTS;
        atomicAdd(stats + ((size_t)slot * 2 + 0) * ld + off + c, a);
        atomicAdd(stats + ((size_t)slot * 2 + 1) * ld + off + c, b);
    }
}

// ---------------------------------------------------------------------------------------------
// head input: out[p, :] = relu(bn0(cat(upsample2(u)[p], x1[p], x2[p]))), bf16 rows of pitch ldo
// (columns >= Cu+C1+C2 are zero).  One thread per (pixel, 8-channel chunk).
// ---------------------------------------------------------------------------------------------
// One block = one output row (b, y); thread t owns chunk t % chunks for pixels t / chunks, t / chunks + ppb, ... so the
// inner loop has no divisions, 16-byte stores of a row are contiguous, and every up-sampled source row is read by
// exactly two blocks.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) head_input_kernel(const dmm_head_t p) {
    pdl_prologue();
    extern __shared__ float coef[];   // [2][Cpad]
    const int Ct = p.Cu + p.C1 + p.C2;
    const int chunks = (int)(p.ldo >> 3);
    const int Cpad = chunks * 8;
    for (int c = threadIdx.x; c < Cpad; c += blockDim.x) {
        float sc = 0.f, sh = 0.f;
        if (c < Ct) {
            const bool writer = blockIdx.x == 0;
            BnCoef k = c < p.Cu ? bn_coef_fwd(p.bn_u, c, writer) : bn_coef_fwd(p.bn_x, c - p.Cu, writer);
            sc = k.scale;
            sh = k.shift;
        }
        coef[c] = sc;
        coef[Cpad + c] = sh;
    }
    __syncthreads();
    const int cu = p.Cu >> 3;                        // chunks that come from the up-sampled decoder output
    const int Cx = p.C1 + p.C2;
    const int UH = p.H >> 1, UW = p.W >> 1;
    const long long HW = (long long)p.H * p.W;
    float* xs = coef + 2 * Cpad;                     // [Cx][2 rows][W]
    // warps 0..6: thread t owns up-sampled chunk t % cu for pixels t / cu, t / cu + ppb, ... (no divisions in the loop);
    // warp 7 alone writes the chunks of the raw input channels, so that no warp runs both loops (divergence)
    const int nraw = chunks - cu;                    // raw-input chunks per pixel (1 for <= 8 raw channels)
    const bool raw_warp = threadIdx.x >= 224;
    const int ppb = raw_warp ? 32 / (nraw > 0 ? nraw : 1) : 224 / cu;
    const int tl = raw_warp ? (int)threadIdx.x - 224 : (int)threadIdx.x;
    const int nper = raw_warp ? nraw : cu;
    const bool worker = nper > 0 && tl < ppb * nper;
    const int ch = raw_warp ? cu + tl % (nper > 0 ? nper : 1) : tl % cu, px0 = tl / (nper > 0 ? nper : 1);
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = coef[ch * 8 + j]; sh[j] = coef[Cpad + ch * 8 + j]; }
    // persistent over PAIRS of image rows (one row of the up-sampled source): the BatchNorm coefficients are finalised once
    // per block, and one load + one activation of a source chunk feeds its 2x2 output pixels (four 16-byte stores)
    for (int ur = blockIdx.x; ur < p.B * UH; ur += gridDim.x) {
    const int uy = ur % UH, b = ur / UH;
    // the raw network inputs of the two rows (fp32 NCHW planes) are staged in shared memory, already activated: xs[c][r][W]
    if (Cx <= 8 && (p.W & 3) == 0) {
        // one float4 per plane and thread, four planes in flight at once (a scalar loop pays one memory round trip per element)
        const int W4 = p.W >> 2;
        for (int i = threadIdx.x; i < 2 * W4; i += blockDim.x) {
            const int r = i >= W4 ? 1 : 0, x4 = i - r * W4;
            const long long rowoff = (long long)(2 * uy + r) * p.W;
            for (int c0 = 0; c0 < Cx; c0 += 4) {
                float4 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = c0 + k;
                    if (c < Cx) {
                        const float* src = c < p.C1 ? p.x1 + ((long long)b * p.C1 + c) * HW + rowoff
                                                    : p.x2 + ((long long)b * p.C2 + (c - p.C1)) * HW + rowoff;
                        v[k] = __ldg(reinterpret_cast<const float4*>(src) + x4);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = c0 + k;
                    if (c < Cx) {
                        const float sc_ = coef[p.Cu + c], sh_ = coef[Cpad + p.Cu + c];
                        float4 o;
                        o.x = fmaxf(fmaf(v[k].x, sc_, sh_), 0.f); o.y = fmaxf(fmaf(v[k].y, sc_, sh_), 0.f);
                        o.z = fmaxf(fmaf(v[k].z, sc_, sh_), 0.f); o.w = fmaxf(fmaf(v[k].w, sc_, sh_), 0.f);
                        *reinterpret_cast<float4*>(xs + (c * 2 + r) * p.W + 4 * x4) = o;
                    }
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < Cx * 2 * p.W; i += blockDim.x) {
            const int cr = i / p.W, xx = i - cr * p.W;
            const int c = cr >> 1, r = cr & 1;
            const long long off = (long long)(2 * uy + r) * p.W + xx;
            const float v = c < p.C1 ? __ldg(p.x1 + ((long long)b * p.C1 + c) * HW + off)
                                     : __ldg(p.x2 + ((long long)b * p.C2 + (c - p.C1)) * HW + off);
            xs[i] = fmaxf(fmaf(v, coef[p.Cu + c], coef[Cpad + p.Cu + c]), 0.f);
        }
    }
    __syncthreads();
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)b * p.H + 2 * uy) * p.W * p.ldo + ch * 8;
    const long long rstep = (long long)p.W * p.ldo;          // next output row
    if (!worker) {
    } else if (ch < cu) {
        const __nv_bfloat16* urow = reinterpret_cast<const __nv_bfloat16*>(p.u) + ((long long)b * UH + uy) * UW * p.ldu + ch * 8;
#pragma unroll 8
        for (int ux = px0; ux < UW; ux += ppb) {
            float f[8];
            unpack8(ldg16(urow + (long long)ux * p.ldu), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
            const uint4 v = pack8(f);
            __nv_bfloat16* o = orow + (long long)(2 * ux) * p.ldo;
            *reinterpret_cast<uint4*>(o) = v;
            *reinterpret_cast<uint4*>(o + p.ldo) = v;
            *reinterpret_cast<uint4*>(o + rstep) = v;
            *reinterpret_cast<uint4*>(o + rstep + p.ldo) = v;
        }
    } else {
        const int c0 = ch * 8 - p.Cu;                // first raw-input channel of this chunk
        for (int i = px0; i < 2 * p.W; i += ppb) {
            const int r = i >= p.W ? 1 : 0, xx = i - r * p.W;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (c0 + j < Cx) ? xs[((c0 + j) * 2 + r) * p.W + xx] : 0.f;
            *reinterpret_cast<uint4*>(orow + r * rstep + (long long)xx * p.ldo) = pack8(f);
        }
    }
    __syncthreads();                                 // xs is overwritten by the next row pair
    }
}

// Backward of the head input.  PASS 0: sums += (sum dz, sum dz*xhat) over all Cu+C1+C2 channels.
// PASS 1: du[parent, c] = sum over the 2x2 children of dx (c < Cu), dgamma/dbeta written.
// One thread per (UP-SAMPLED-SOURCE pixel, chunk) so that the 4 children are reduced in registers.
template <int PASS>
__global__ void __launch_bounds__(256) head_input_bwd_kernel(const dmm_head_bwd_t p) {
    pdl_prologue();
    extern __shared__ float sm[];   // PASS 0: [2][Cpad] block partial sums ; coefficient tables
    const int Ct = p.Cu + p.C1 + p.C2;
    const int chunks = (Ct + 7) >> 3;
    const int Cpad = chunks * 8;
    float* mean = sm;
    float* istd = sm + Cpad;
    float* scl = sm + 2 * Cpad;
    float* sft = sm + 3 * Cpad;
    float* k1 = sm + 4 * Cpad;   // PASS 0: block sums of dz        PASS 1: c1
    float* k2 = sm + 5 * Cpad;   // PASS 0: block sums of dz*xhat   PASS 1: c2
    for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < Cpad; c += blockDim.x * blockDim.y) {
        float m = 0.f, is = 0.f, sc = 0.f, sh = 0.f, a1 = 0.f, a2 = 0.f;
        if (c < Ct) {
            const dmm_bn_bwd_t& bn = c < p.Cu ? p.bn_u : p.bn_x;
            const int cc = c < p.Cu ? c : c - p.Cu;
            BnCoef k = bn_coef_bwd(bn, cc);
            m = k.mean; is = k.invstd; sc = k.scale; sh = k.shift;
            if (PASS == 1) {
                double a = 0.0, b = 0.0;
                for (int s = 0; s < DMM_STATS_SLOTS; ++s) {
                    const double* r = bn.sums + (size_t)s * 2 * bn.sums_ld + bn.sums_off + cc;
                    a += r[0];
                    b += r[bn.sums_ld];
                }
                a1 = (float)(a / bn.count);
                a2 = (float)(b / bn.count);
                if (blockIdx.x == 0) {
                    if (bn.dgamma) bn.dgamma[cc] = (float)b;
                    if (bn.dbeta) bn.dbeta[cc] = (float)a;
                }
            }
        }
        mean[c] = m; istd[c] = is; scl[c] = sc; sft[c] = sh; k1[c] = a1; k2[c] = a2;
    }
    __syncthreads();
    const __nv_bfloat16* u = reinterpret_cast<const __nv_bfloat16*>(p.u);
    const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(p.g);
    const int UH = p.H >> 1, UW = p.W >> 1;
    const long long HW = (long long)p.H * p.W;
    // block = (chunk lanes, pixel rows): a thread keeps ONE 8-channel chunk of the Cu up-sampled channels and walks over
    // up-sampled-SOURCE pixels, reducing the 2x2 children in registers (the raw input channels are handled by
    // head_raw_bwd_reduce_kernel).  rows < 2^31 (checked on the host): 32-bit index arithmetic.
    const int ch = threadIdx.x;
    const int ry = blockDim.y;
    const unsigned rows = (unsigned)p.B * (unsigned)UH * (unsigned)UW;
    float tot1[8], tot2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tot1[j] = tot2[j] = 0.f;
    float c_sc[8], c_sh[8], c_mu[8], c_is[8], c_k1[8], c_k2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = ch * 8 + j;
        c_sc[j] = scl[c]; c_sh[j] = sft[c]; c_mu[j] = mean[c]; c_is[j] = istd[c]; c_k1[j] = k1[c]; c_k2[j] = k2[c];
    }
    for (unsigned up = blockIdx.x * ry + threadIdx.y; up < rows; up += gridDim.x * ry) {
        const unsigned ux = up % (unsigned)UW;
        const unsigned t = up / (unsigned)UW;
        const unsigned uy = t % (unsigned)UH;
        const unsigned b = t / (unsigned)UH;
        float xu[8];
        unpack8(ldg16(u + (long long)up * p.ldu + ch * 8), xu);
        const long long pix0 = ((long long)b * p.H + 2 * uy) * p.W + 2 * ux;
        float g4[4][8];
        unpack8(ldg16(g + pix0 * p.ldg + ch * 8), g4[0]);
        unpack8(ldg16(g + (pix0 + 1) * p.ldg + ch * 8), g4[1]);
        unpack8(ldg16(g + (pix0 + p.W) * p.ldg + ch * 8), g4[2]);
        unpack8(ldg16(g + (pix0 + p.W + 1) * p.ldg + ch * 8), g4[3]);
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gs = (g4[0][j] + g4[1][j]) + (g4[2][j] + g4[3][j]);      // the 4 children share x, hence the mask
            const float dz = fmaf(xu[j], c_sc[j], c_sh[j]) > 0.f ? gs : 0.f;
            const float xh = (xu[j] - c_mu[j]) * c_is[j];
            if (PASS == 0) {
                tot1[j] += dz;
                tot2[j] = fmaf(dz, xh, tot2[j]);
            } else {
                const float gam = p.bn_u.gamma ? p.bn_u.gamma[ch * 8 + j] : 1.f;
                o[j] = gam * c_is[j] * (dz - 4.f * (c_k1[j] + xh * c_k2[j]));
            }
        }
        if (PASS == 1) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.du) + (long long)up * p.lddu + ch * 8) = pack8(o);
    }
    if (PASS == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&k1[ch * 8 + j], tot1[j]);     // one shared-memory atomic per thread and channel, at the very end
            atomicAdd(&k2[ch * 8 + j], tot2[j]);
        }
        __syncthreads();
        const int slot = blockIdx.x % DMM_STATS_SLOTS;
        for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < p.Cu; c += blockDim.x * blockDim.y) {
            atomicAdd(p.bn_u.sums + ((size_t)slot * 2 + 0) * p.bn_u.sums_ld + p.bn_u.sums_off + c, (double)k1[c]);
            atomicAdd(p.bn_u.sums + ((size_t)slot * 2 + 1) * p.bn_u.sums_ld + p.bn_u.sums_off + c, (double)k2[c]);
        }
    }
}

// (sum dz, sum dz*xhat) of the C1+C2 (<= 8) RAW input channels of the head BatchNorm: one thread per pixel, the
// fp32 planes are read coalesced, the 8 gradient columns [Cu, Cu+8) with one 16-byte load.
__global__ void __launch_bounds__(256) head_raw_bwd_reduce_kernel(const dmm_head_bwd_t p) {
    pdl_prologue();
    __shared__ float red[2][8][8];
    const int Cx = p.C1 + p.C2;
    float sc[8], sh[8], mu[8], is[8], s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = sh[j] = mu[j] = is[j] = s1[j] = s2[j] = 0.f;
        if (j < Cx) {
            BnCoef k = bn_coef_bwd(p.bn_x, j);
            sc[j] = k.scale; sh[j] = k.shift; mu[j] = k.mean; is[j] = k.invstd;
        }
    }
    const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(p.g);
    const long long HW = (long long)p.H * p.W;
    const long long total = (long long)p.B * HW;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        const long long b = pix / HW, r = pix - b * HW;
        float gg[8];
        unpack8(ldg16(g + pix * p.ldg + p.Cu), gg);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < Cx) {
                const float xv = j < p.C1 ? __ldg(p.x1 + (b * p.C1 + j) * HW + r) : __ldg(p.x2 + (b * p.C2 + (j - p.C1)) * HW + r);
                const float dz = fmaf(xv, sc[j], sh[j]) > 0.f ? gg[j] : 0.f;
                s1[j] += dz;
                s2[j] = fmaf(dz, (xv - mu[j]) * is[j], s2[j]);
            }
        }
    }
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float a = s1[j], c = s2[j];
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if ((threadIdx.x & 31) == 0) {
            red[0][j][w] = a;
            red[1][j][w] = c;
        }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        const int which = threadIdx.x >> 3, j = threadIdx.x & 7;
        if (j < Cx) {
            double d = 0.0;
            for (int k = 0; k < 8; ++k) d += red[which][j][k];
            const int slot = blockIdx.x % DMM_STATS_SLOTS;
            atomicAdd(p.bn_x.sums + ((size_t)slot * 2 + which) * p.bn_x.sums_ld + p.bn_x.sums_off + j, d);
        }
    }
}

// dlogits (B, C, H, W) fp32 -> bf16 rows [B*H*W, ld]: column t*C + n = dlogits[n] at pixel (y - (kh - K/2), x - (kw - K/2)),
// t = kh*K + kw (zero outside the image, zero beyond K*K*C).  This "im2col of the output gradient" turns both the
// weight gradient and the data gradient of the N = num_classes KxK convolution (refine1) into plain 1x1 GEMMs.
__global__ void __launch_bounds__(256) dlogits_im2col_kernel(const float* __restrict__ dl, int B, int C, int H, int W, int KH, int K,
                                                             __nv_bfloat16* __restrict__ out, int ld) {
    pdl_prologue();
    const int chunks = ld >> 3;
    const long long total = (long long)B * H * W * chunks;
    const int pad = K / 2, pad_h = KH / 2;
    const int KC = KH * K * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long long pix = i / chunks;
        const int x = (int)(pix % W);
        const long long t2 = pix / W;
        const int y = (int)(t2 % H);
        const int b = (int)(t2 / H);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = ch * 8 + j;
            float v = 0.f;
            if (col < KC) {
                const int t = col / C, n = col - t * C;
                const int kh = t / K, kw = t - kh * K;
                const int yy = y - (kh - pad_h), xx = x - (kw - pad);
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(dl + (((long long)b * C + n) * H + yy) * W + xx);
            }
            f[j] = v;
        }
        *reinterpret_cast<uint4*>(out + pix * ld + ch * 8) = pack8(f);
    }
}

// fp32 NCHW -> bf16 pixel-major rows (channels >= C zero-filled up to ldo)
__global__ void __launch_bounds__(256) nchw_to_rows_kernel(const float* __restrict__ x, int B, int C, long long HW,
                                                           __nv_bfloat16* __restrict__ out, int ldo) {
    pdl_prologue();
    const int chunks = ldo >> 3;
    const long long total = (long long)B * HW * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // pixel-fastest decomposition keeps the fp32 plane reads coalesced
        const long long pin = i % HW;
        const long long t = i / HW;
        const int ch = (int)(t % chunks);
        const int b = (int)(t / chunks);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            f[j] = c < C ? __ldg(x + ((long long)b * C + c) * HW + pin) : 0.f;
        }
        *reinterpret_cast<uint4*>(out + ((long long)b * HW + pin) * ldo + ch * 8) = pack8(f);
    }
}

// fp32 rows -> bf16 rows (gradient slices of the fp32 dense-block gradient buffer)
__global__ void __launch_bounds__(256) rows_f32_to_bf16_kernel(const float* __restrict__ src, long long lds,
                                                               __nv_bfloat16* __restrict__ dst, long long ldd,
                                                               long long P, int C) {
    pdl_prologue();
    const int chunks = C >> 3;
    const long long total = P * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long long r = i / chunks;
        float f[8];
        load_g8<float>(src, r * lds + ch * 8, f);
        *reinterpret_cast<uint4*>(dst + r * ldd + ch * 8) = pack8(f);
    }
}

// ---------------------------------------------------------------------------------------------
// loss
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                         long long n4, int C, long long HW4, float* __restrict__ loss,
                                                         float* __restrict__ grad, double* class_sums) {
    pdl_prologue();
    // vectors of 4 never straddle a (b, c) plane because HW % 4 == 0 (checked on the host)
    __shared__ double sm[8][8];
    float cs[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) cs[c] = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float4 tv = __ldg(reinterpret_cast<const float4*>(t) + i);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        const float ts[4] = {tv.x, tv.y, tv.z, tv.w};
        float l[4], g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float e = expf(-fabsf(xs[j]));
            l[j] = fmaxf(xs[j], 0.f) - xs[j] * ts[j] + log1pf(e);
            const float r = 1.f / (1.f + e);
            g[j] = (xs[j] >= 0.f ? r : e * r) - ts[j];
        }
        if (loss) reinterpret_cast<float4*>(loss)[i] = make_float4(l[0], l[1], l[2], l[3]);
        if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(g[0], g[1], g[2], g[3]);
        if (class_sums) {
            const int c = (int)((i / HW4) % C);
            const float s = (l[0] + l[1]) + (l[2] + l[3]);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k == c) cs[k] += s;
        }
    }
    if (class_sums) {
        const int w = threadIdx.x >> 5;
        for (int c = 0; c < C && c < 8; ++c) {
            double d = cs[c];
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            if ((threadIdx.x & 31) == 0) sm[c][w] = d;
        }
        __syncthreads();
        if (threadIdx.x < C && threadIdx.x < 8) {
            double d = 0;
            for (int k = 0; k < 8; ++k) d += sm[threadIdx.x][k];
            atomicAdd(class_sums + threadIdx.x, d);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// weights: fp32 parameter layout <-> packed bf16 K-major GEMM operand / fp32 wgrad result
// ---------------------------------------------------------------------------------------------
struct PackArgs {
    int T, C, Kp, n_valid, n_rows;
    long long ktot, sn, sc;
    int tap_off[DMM_MAX_TAPS];
};
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst,
                                                           const PackArgs a) {
    pdl_prologue();
    const long long total = (long long)a.n_rows * a.ktot;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / a.ktot);
        const long long k = i - (long long)n * a.ktot;
        const int t = (int)(k / a.Kp);
        const int c = (int)(k - (long long)t * a.Kp);
        float v = 0.f;
        if (n < a.n_valid && c < a.C && t < a.T) v = w[(long long)n * a.sn + (long long)c * a.sc + a.tap_off[t]];
        dst[i] = __float2bfloat16_rn(v);
    }
}
// grad[n*sn + m*sc + tap_off[t]] (=|+=) dw[t*dt + m*dm + n*dn]
__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const float* __restrict__ dw, long long dt, long long dm,
                                                           long long dn, int M, int N, float* __restrict__ grad,
                                                           const PackArgs a, int accumulate) {
    pdl_prologue();
    const long long total = (long long)a.T * M * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const long long r = i / N;
        const int m = (int)(r % M);
        const int t = (int)(r / M);
        const float v = dw[(long long)t * dt + (long long)m * dm + (long long)n * dn];
        float* gp = grad + (long long)n * a.sn + (long long)m * a.sc + a.tap_off[t];
        *gp = accumulate ? *gp + v : v;
    }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                   float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
    pdl_prologue();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i];
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

// batched forms: one launch for all weight tensors of the network (job tables live in device memory).
// `work` (optional): int32 pairs (job, chunk) - one block per `chunk_elems` consecutive elements of a job, so that the 9.4 M
// element ConvTranspose tensors and the 64-element BatchNorm-sized jobs load the SMs evenly (without it: 32 blocks per job).
__device__ __forceinline__ void pack_job_range(const dmm_pack_job_t& j, long long i0, long long i1, long long stride) {
    const int Kp = (j.C + j.kwidth - 1) / j.kwidth * j.kwidth;
    const long long ktot = (long long)Kp * j.T;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(j.dst);
    const int cdiv = j.cdiv > 0 ? j.cdiv : 1;
    for (long long i = i0; i < i1; i += stride) {
        const int n = (int)(i / ktot);
        const int k = (int)(i - (long long)n * ktot);
        const int t = k / Kp;
        const int c = k - t * Kp;
        float v = 0.f;
        if (n < j.n_valid && c < j.C) {
            const long long nidx = j.ndiv > 1 ? (long long)(n / j.ndiv) * j.sn + (long long)(n % j.ndiv) * j.sn2 : (long long)n * j.sn;
            v = __ldg(j.w + nidx + (long long)(c / cdiv) * j.sc + (long long)(c % cdiv) * j.sc2 + j.tap_off[t]);
            if (j.rscale) v *= __ldg(j.rscale + (j.ndiv > 1 ? n % j.ndiv : n));     // folded kernel columns: row kw*Cout + n -> channel n
        }
        dst[i] = __float2bfloat16_rn(v);
    }
}
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const dmm_pack_job_t* __restrict__ jobs) {
    pdl_prologue();
    const dmm_pack_job_t& j = jobs[blockIdx.y];
    const int Kp = (j.C + j.kwidth - 1) / j.kwidth * j.kwidth;
    const long long total = (long long)j.n_rows * Kp * j.T;
    pack_job_range(j, (long long)blockIdx.x * blockDim.x + threadIdx.x, total, (long long)gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) pack_weights_work_kernel(const dmm_pack_job_t* __restrict__ jobs, const int2* __restrict__ work,
                                                                int chunk_elems) {
    pdl_prologue();
    const int2 w = work[blockIdx.x];
    const dmm_pack_job_t& j = jobs[w.x];
    const int Kp = (j.C + j.kwidth - 1) / j.kwidth * j.kwidth;
    const long long total = (long long)j.n_rows * Kp * j.T;
    const long long i0 = (long long)w.y * chunk_elems;
    const long long i1 = i0 + chunk_elems < total ? i0 + chunk_elems : total;
    pack_job_range(j, i0 + threadIdx.x, i1, blockDim.x);
}
__device__ __forceinline__ void unpack_job_range(const dmm_unpack_job_t& j, long long i0, long long i1, long long stride) {
    for (long long i = i0; i < i1; i += stride) {
        const int n = (int)(i % j.N);
        const int r = (int)(i / j.N);
        const int m = r % j.M;
        const int t = r / j.M;
        const float v = j.dw[(long long)t * j.dt + (long long)m * j.dm + (long long)n * j.dn];
        const long long nidx = j.ndiv > 1 ? (long long)(n % j.ndiv) * j.sn + (long long)(n / j.ndiv) * j.sn2 : (long long)n * j.sn;
        float* gp = j.grad + nidx + (long long)m * j.sc + j.tap_off[t];
        *gp = j.accumulate ? *gp + v : v;
    }
}
__global__ void __launch_bounds__(256) unpack_wgrad_batched_kernel(const dmm_unpack_job_t* __restrict__ jobs) {
    pdl_prologue();
    const dmm_unpack_job_t& j = jobs[blockIdx.y];
    unpack_job_range(j, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)j.T * j.M * j.N, (long long)gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) unpack_wgrad_work_kernel(const dmm_unpack_job_t* __restrict__ jobs, const int2* __restrict__ work,
                                                                int chunk_elems) {
    pdl_prologue();
    const int2 w = work[blockIdx.x];
    const dmm_unpack_job_t& j = jobs[w.x];
    const long long total = (long long)j.T * j.M * j.N;
    const long long i0 = (long long)w.y * chunk_elems;
    const long long i1 = i0 + chunk_elems < total ? i0 + chunk_elems : total;
    unpack_job_range(j, i0 + threadIdx.x, i1, blockDim.x);
}

// resident 256-thread blocks per SM of a kernel (cached): a grid-stride kernel runs best as exactly one full wave
template <typename F>
static int resident_bps(F fn) {
    struct Entry { const void* f; int n; };
    static Entry cache[64];
    static int ncache = 0;
    const void* key = reinterpret_cast<const void*>(fn);
    for (int i = 0; i < ncache; ++i)
        if (cache[i].f == key) return cache[i].n;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kEwThreads, 0) != cudaSuccess || n < 1) n = kMaxBlocksPerSm;
    if (ncache < 64) cache[ncache++] = {key, n};
    return n;
}
#define DMM_WAVE_LAUNCH(KERNEL, C_, ROWS_, STREAM, ...)                          \
    do {                                                                         \
        const ColCfg kw_ = col_cfg((C_), (ROWS_), resident_bps(KERNEL));         \
        launch_k(KERNEL, kw_.grid, kw_.block, 0, (STREAM), __VA_ARGS__);               \
    } while (0)

static unsigned flat_grid(long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    const long long cap = (long long)kNumSm * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace dmm

using namespace dmm;

extern "C" int dmm_bn_relu_apply(const dmm_bn_apply_t* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DMM_CHECK(d && d->x && d->y, "dmm_bn_relu_apply: null pointer");
    DMM_CHECK(d->C > 0 && d->C % 8 == 0, "dmm_bn_relu_apply: C=%d must be a positive multiple of 8", d->C);
    DMM_CHECK(d->ldx % 8 == 0 && d->ldy % 8 == 0, "dmm_bn_relu_apply: row pitches must be multiples of 8");
    DMM_CHECK(d->pool >= 0 && d->pool <= 2, "dmm_bn_relu_apply: pool=%d", d->pool);
    DMM_CHECK(d->bn.training ? (d->bn.stats != nullptr && d->bn.count > 0) : (d->bn.running_mean && d->bn.running_var),
              "dmm_bn_relu_apply: missing statistics");
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;
    int OH = d->H, OW = d->W;
    if (d->pool == 1) { OH = d->H / 2; OW = d->W / 2; }
    if (d->pool == 2) { OH = (d->H - 1) / 2 + 1; OW = (d->W - 1) / 2 + 1; }
    if (OH <= 0 || OW <= 0) return 0;
    DMM_CHECK(d->pool == 0 || (long long)d->B * d->H * d->W < (1ll << 31), "dmm_bn_relu_apply: too many pixels for a pooled mode");
    const long long orows = (long long)d->B * OH * OW;
    if (d->pool == 0 && d->ystats == nullptr) DMM_WAVE_LAUNCH(bn_apply_fast_kernel, d->C, orows, stream, *d);
    else if (d->pool == 0) DMM_WAVE_LAUNCH(bn_relu_apply_kernel<0>, d->C, orows, stream, *d, OH, OW);
    else if (d->pool == 1) DMM_WAVE_LAUNCH(bn_relu_apply_kernel<1>, d->C, orows, stream, *d, OH, OW);
    else DMM_WAVE_LAUNCH(bn_relu_apply_kernel<2>, d->C, orows, stream, *d, OH, OW);
    DMM_LAUNCH_CHECK("bn_relu_apply_kernel");
    return 0;
}

template <int PASS>
static int launch_bn_bwd(const dmm_bn_bwd_args_t* d, cudaStream_t stream) {
    DMM_CHECK(d && d->x && d->g, "dmm_bn_relu_bwd: null pointer");
    DMM_CHECK(d->C > 0 && d->C % 8 == 0, "dmm_bn_relu_bwd: C=%d must be a positive multiple of 8", d->C);
    DMM_CHECK(d->ldx % 8 == 0 && d->ldg % 8 == 0 && d->ldo % 8 == 0, "dmm_bn_relu_bwd: row pitches must be multiples of 8");
    DMM_CHECK(d->gmode >= 0 && d->gmode <= 2, "dmm_bn_relu_bwd: gmode=%d", d->gmode);
    DMM_CHECK(d->bn.sums && d->bn.save_mean && d->bn.save_invstd && d->bn.count > 0, "dmm_bn_relu_bwd: missing BN state");
    DMM_CHECK(d->out_mode >= 0 && d->out_mode <= 2, "dmm_bn_relu_bwd: out_mode=%d", d->out_mode);
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;
    DMM_CHECK(d->gmode == 0 || (long long)d->B * d->H * d->W < (1ll << 31), "dmm_bn_relu_bwd: too many pixels for a pooled gradient mode");
    int OH = d->H, OW = d->W;
    if (d->gmode == 1) { OH = d->H / 2; OW = d->W / 2; }
    if (d->gmode == 2) { OH = (d->H - 1) / 2 + 1; OW = (d->W - 1) / 2 + 1; }
    const long long irows = (long long)d->B * d->H * d->W;
    if (d->gmode == 0 && d->dz_out == nullptr) {      // lean same-pixel kernels
        const ColCfg k = col_cfg(d->C, irows, kFastBps);
        if (PASS == 0) {
            if (d->g_is_f32) launch_k(bn_bwd_reduce_fast_kernel<float>, k.grid, k.block, 0, stream, *d);
            else launch_k(bn_bwd_reduce_fast_kernel<__nv_bfloat16>, k.grid, k.block, 0, stream, *d);
        } else {
#define DMM_APPLY_FAST(GT)                                                                              \
    do {                                                                                                \
        if (d->out_mode == 0) launch_k(bn_bwd_apply_fast_kernel<GT, 0>, k.grid, k.block, 0, stream, *d);      \
        else if (d->out_mode == 1) launch_k(bn_bwd_apply_fast_kernel<GT, 1>, k.grid, k.block, 0, stream, *d); \
        else launch_k(bn_bwd_apply_fast_kernel<GT, 2>, k.grid, k.block, 0, stream, *d);                       \
    } while (0)
            if (d->g_is_f32) DMM_APPLY_FAST(float);
            else DMM_APPLY_FAST(__nv_bfloat16);
#undef DMM_APPLY_FAST
        }
        DMM_LAUNCH_CHECK("bn_bwd fast kernel");
        return 0;
    }
    if (d->gmode == 1 && !d->g_is_f32 && d->dz_out == nullptr && d->H % 2 == 0 && d->W % 2 == 0) {      // lean pooled kernels
        ColCfg kp = col_cfg(d->C, (long long)d->B * OH * OW);
        unsigned gx = (unsigned)(d->B * OH);
        const unsigned cap = (unsigned)(kNumSm * 3) / kp.grid.y;      // launch bounds (256, 3): one full wave
        if (gx > cap) gx = cap > 0 ? cap : 1;
        kp.grid.x = gx;
        if (PASS == 0) launch_k(bn_bwd_pool_fast_kernel<0, 0>, kp.grid, kp.block, 0, stream, *d, OH, OW);
        else if (d->out_mode == 0) launch_k(bn_bwd_pool_fast_kernel<1, 0>, kp.grid, kp.block, 0, stream, *d, OH, OW);
        else if (d->out_mode == 1) launch_k(bn_bwd_pool_fast_kernel<1, 1>, kp.grid, kp.block, 0, stream, *d, OH, OW);
        else launch_k(bn_bwd_pool_fast_kernel<1, 2>, kp.grid, kp.block, 0, stream, *d, OH, OW);
        DMM_LAUNCH_CHECK("bn_bwd pooled kernel");
        return 0;
    }
    if (PASS == 0 && d->gmode == 2 && !d->g_is_f32 && d->argmax && d->dz_out) {      // lean max-pool reduce (the stem)
        const long long quads = (long long)d->B * ((d->H + 1) / 2) * ((d->W + 1) / 2);
        DMM_WAVE_LAUNCH(bn_bwd_maxpool_quad_kernel, d->C, quads, stream, *d, OH, OW);
        DMM_LAUNCH_CHECK("bn_bwd_maxpool_quad_kernel");
        return 0;
    }
#define DMM_BWD_LAUNCH(GM, GT)                                                                      \
    do {                                                                                            \
        if (PASS == 0) DMM_WAVE_LAUNCH((bn_relu_bwd_reduce_kernel<GM, GT>), d->C, irows, stream, *d, OH, OW); \
        else DMM_WAVE_LAUNCH((bn_relu_bwd_apply_kernel<GM, GT>), d->C, irows, stream, *d, OH, OW);           \
    } while (0)
    if (d->g_is_f32) {
        if (d->gmode == 0) DMM_BWD_LAUNCH(0, float);
        else if (d->gmode == 1) DMM_BWD_LAUNCH(1, float);
        else DMM_BWD_LAUNCH(2, float);
    } else {
        if (d->gmode == 0) DMM_BWD_LAUNCH(0, __nv_bfloat16);
        else if (d->gmode == 1) DMM_BWD_LAUNCH(1, __nv_bfloat16);
        else DMM_BWD_LAUNCH(2, __nv_bfloat16);
    }
#undef DMM_BWD_LAUNCH
    DMM_LAUNCH_CHECK("bn_relu_bwd kernel");
    return 0;
}

extern "C" int dmm_bn_relu_bwd_reduce(const dmm_bn_bwd_args_t* d, void* stream) {
    return launch_bn_bwd<0>(d, (cudaStream_t)stream);
}
extern "C" int dmm_bn_relu_bwd_apply(const dmm_bn_bwd_args_t* d, void* stream) {
    return launch_bn_bwd<1>(d, (cudaStream_t)stream);
}

extern "C" int dmm_bn_relu_bwd_contrib(const dmm_bn_bwd_args_t* d, void* stream) {
    DMM_CHECK(d && d->x && d->g && d->out, "dmm_bn_relu_bwd_contrib: null pointer");
    DMM_CHECK(d->C > 0 && d->C % 8 == 0, "dmm_bn_relu_bwd_contrib: C=%d must be a positive multiple of 8", d->C);
    DMM_CHECK(d->ldx % 8 == 0 && d->ldg % 8 == 0 && d->ldo % 8 == 0, "dmm_bn_relu_bwd_contrib: row pitches must be multiples of 8");
    DMM_CHECK(d->out_gw == 0 || (d->out_gw % 8 == 0 && d->out_plane % 8 == 0 && d->out_plane > 0), "dmm_bn_relu_bwd_contrib: bad planar slab");
    DMM_CHECK(d->gmode == 0 && !d->g_is_f32, "dmm_bn_relu_bwd_contrib: same-pixel bf16 gradients only");
    DMM_CHECK(d->bn.sums && d->bn.save_mean && d->bn.save_invstd, "dmm_bn_relu_bwd_contrib: missing BN state");
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;
    ColCfg k = col_cfg(d->C, (long long)d->B * d->H * d->W, kFastBps);
    launch_k(bn_bwd_contrib_kernel<kFastRows, true>, k.grid, k.block, 0, (cudaStream_t)stream, *d);
    DMM_LAUNCH_CHECK("bn_bwd_contrib_kernel");
    return 0;
}

extern "C" int dmm_bn_bwd_finalize(const dmm_bn_bwd_t* bn, int32_t C, float* k, void* stream) {
    DMM_CHECK(bn && k && bn->sums && bn->save_invstd && bn->count > 0, "dmm_bn_bwd_finalize: missing inputs");
    if (C <= 0) return 0;
    launch_k(bn_bwd_finalize_kernel, (unsigned)((C + 255) / 256), 256, 0, (cudaStream_t)stream, *bn, C, k);
    DMM_LAUNCH_CHECK("bn_bwd_finalize_kernel");
    return 0;
}

extern "C" int dmm_grad_gather(const dmm_grad_gather_t* d, void* stream) {
    DMM_CHECK(d && d->out && d->nsrc >= 1 && d->nsrc <= DMM_GATHER_MAX && d->nk >= 0 && d->nk <= DMM_GATHER_MAX,
              "dmm_grad_gather: bad descriptor");
    DMM_CHECK(d->C > 0 && d->C % 8 == 0 && d->ldo % 8 == 0, "dmm_grad_gather: C / ldo must be multiples of 8");
    DMM_CHECK(d->nk == 0 || (d->x && d->mean && d->ldx % 8 == 0), "dmm_grad_gather: corrections need x and mean");
    for (int s = 0; s < d->nsrc; ++s) {
        DMM_CHECK(d->src[s] && d->ld[s] % 8 == 0, "dmm_grad_gather: source %d", s);
        DMM_CHECK(d->plane[s] == 0 || (d->gw > 0 && d->gw % 8 == 0 && d->plane[s] % 8 == 0), "dmm_grad_gather: planar source %d", s);
    }
    if (d->rows <= 0) return 0;
    static const int gather_bps = env_int_ew("DMM_GATHER_BPS", kFastBps);
    ColCfg k = col_cfg(d->C, d->rows, gather_bps);
    launch_k(grad_gather_kernel<2, 8>, k.grid, k.block, 0, (cudaStream_t)stream, *d);      // (4 rows x 4 sources measured slower)
    DMM_LAUNCH_CHECK("grad_gather_kernel");
    return 0;
}

extern "C" int dmm_im2col_7x7s2(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t B, int32_t H,
                                int32_t W, void* out, int32_t kpad, void* stream) {
    DMM_CHECK(x1 && out && C1 > 0 && C2 >= 0 && (C2 == 0 || x2), "dmm_im2col_7x7s2: bad inputs");
    DMM_CHECK(kpad % 8 == 0 && kpad >= (C1 + C2) * 49, "dmm_im2col_7x7s2: kpad=%d too small / not a multiple of 8", kpad);
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
    const int segs = (OW + kIm2colPx - 1) / kIm2colPx;
    const size_t smem = (size_t)(C1 + C2) * 7 * (2 * kIm2colPx + 5) * sizeof(float) + (size_t)kpad * sizeof(int);
    DMM_CHECK(smem <= 48 * 1024, "dmm_im2col_7x7s2: %d input channels need %zu bytes of shared memory", C1 + C2, smem);
    launch_k(im2col_7x7s2_kernel, (unsigned)((long long)B * OH * segs), 256, smem, (cudaStream_t)stream, 
        x1, C1, x2, C2, B, H, W, OH, OW, reinterpret_cast<__nv_bfloat16*>(out), kpad);
    DMM_LAUNCH_CHECK("im2col_7x7s2_kernel");
    return 0;
}

extern "C" int dmm_nchw_stats(const float* x, int32_t B, int32_t C, int64_t HW, double* stats, int32_t stats_ld,
                              int32_t stats_off, void* stream) {
    DMM_CHECK(x && stats && C > 0, "dmm_nchw_stats: bad inputs");
    if (B <= 0 || HW <= 0) return 0;
    int splits = (int)((HW + 65535) / 65536);
    if (splits < 1) splits = 1;
    launch_k(nchw_stats_kernel, (unsigned)(B * C * splits), 256, 0, (cudaStream_t)stream, x, C, HW, stats, stats_ld, stats_off, splits);
    DMM_LAUNCH_CHECK("nchw_stats_kernel");
    return 0;
}

extern "C" int dmm_head_input(const dmm_head_t* d, void* stream) {
    DMM_CHECK(d && d->u && d->x1 && d->out, "dmm_head_input: null pointer");
    DMM_CHECK(d->Cu % 8 == 0 && d->ldu % 8 == 0 && d->ldo % 8 == 0, "dmm_head_input: Cu / pitches must be multiples of 8");
    DMM_CHECK(d->ldo >= d->Cu + d->C1 + d->C2, "dmm_head_input: ldo too small");
    DMM_CHECK(d->Cu >= 8 && d->Cu <= 2048, "dmm_head_input: Cu=%d out of range", d->Cu);
    DMM_CHECK(d->H % 2 == 0 && d->W % 2 == 0, "dmm_head_input: H and W must be even (nn.Upsample x2 of the decoder output)");
    DMM_CHECK(d->C2 == 0 || d->x2, "dmm_head_input: x2 missing");
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;
    const int chunks = (int)(d->ldo / 8);
    DMM_CHECK(chunks <= 256, "dmm_head_input: ldo %lld too large", (long long)d->ldo);
    DMM_CHECK(d->Cu >= 8 && d->Cu <= 224 * 8 && chunks - d->Cu / 8 <= 32 && chunks * 8 >= d->Cu + d->C1 + d->C2,
              "dmm_head_input: unsupported channel split (Cu=%d, raw=%d, ldo=%lld)", d->Cu, d->C1 + d->C2, (long long)d->ldo);
    const size_t smem = ((size_t)2 * chunks * 8 + (size_t)(d->C1 + d->C2) * 2 * d->W) * sizeof(float);
    DMM_CHECK(smem <= 96 * 1024, "dmm_head_input: two rows of %d raw channels x %d pixels do not fit in shared memory", d->C1 + d->C2, d->W);
    // MINB = 3 (85 registers, no spills; default) or 4 (64 registers, 120 bytes of spills): DMM_HEAD_MINB
    static const int minb = env_int_ew("DMM_HEAD_MINB", 3);      // measured r02: 1.85 -> 1.71 ms
    void (*kern)(const dmm_head_t) = minb == 3 ? head_input_kernel<3> : head_input_kernel<4>;
    static bool head_attr = false;
    if (!head_attr) {
        DMM_CUDA(cudaFuncSetAttribute(head_input_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        DMM_CUDA(cudaFuncSetAttribute(head_input_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        head_attr = true;
    }
    int head_bps = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&head_bps, kern, 256, smem) != cudaSuccess || head_bps < 1) head_bps = minb;
    static const int head_bps_env = env_int_ew("DMM_HEAD_BPS", 0);
    const long long hrows = (long long)d->B * (d->H / 2), hcap = (long long)kNumSm * (head_bps_env > 0 ? head_bps_env : head_bps);
    launch_k(kern, (unsigned)(hrows < hcap ? hrows : hcap), 256, smem, (cudaStream_t)stream, *d);
    DMM_LAUNCH_CHECK("head_input_kernel");
    return 0;
}

template <int PASS>
static int launch_head_bwd(const dmm_head_bwd_t* d, cudaStream_t stream) {
    DMM_CHECK(d && d->u && d->x1 && d->g, "dmm_head_input_bwd: null pointer");
    DMM_CHECK(d->Cu % 8 == 0 && d->ldu % 8 == 0 && d->ldg % 8 == 0, "dmm_head_input_bwd: Cu / pitches must be multiples of 8");
    DMM_CHECK(d->H % 2 == 0 && d->W % 2 == 0, "dmm_head_input_bwd: H and W must be even");
    DMM_CHECK(PASS == 0 || (d->du && d->lddu % 8 == 0), "dmm_head_input_bwd_apply: du missing");
    if (d->B <= 0 || d->H <= 0 || d->W <= 0) return 0;
    const int Ct = d->Cu + d->C1 + d->C2;
    const int chunks = (Ct + 7) / 8;
    DMM_CHECK(d->ldg >= chunks * 8, "dmm_head_input_bwd: ldg too small");
    const int nch = d->Cu / 8;
    DMM_CHECK(nch >= 1 && nch <= 128, "dmm_head_input_bwd: %d channel chunks (supported: 1..128)", nch);
    DMM_CHECK(d->C1 + d->C2 <= 8, "dmm_head_input_bwd: at most 8 raw input channels are supported (got %d)", d->C1 + d->C2);
    const int ry = 256 / nch > 0 ? 256 / nch : 1;
    const long long rows = (long long)d->B * (d->H / 2) * (d->W / 2);
    DMM_CHECK(rows < (1ll << 31), "dmm_head_input_bwd: too many pixels");
    long long gx = (rows + ry - 1) / ry;
    if (gx > 148 * 8) gx = 148 * 8;
    const size_t smem = (size_t)6 * chunks * 8 * sizeof(float);
    launch_k(head_input_bwd_kernel<PASS>, dim3((unsigned)gx), dim3((unsigned)nch, (unsigned)ry), smem, stream, *d);
    if (PASS == 0)
        launch_k(head_raw_bwd_reduce_kernel, flat_grid((long long)d->B * d->H * d->W, 256), 256, 0, stream, *d);
    DMM_LAUNCH_CHECK("head_input_bwd_kernel");
    return 0;
}
extern "C" int dmm_head_input_bwd_reduce(const dmm_head_bwd_t* d, void* stream) {
    return launch_head_bwd<0>(d, (cudaStream_t)stream);
}
extern "C" int dmm_head_input_bwd_apply(const dmm_head_bwd_t* d, void* stream) {
    return launch_head_bwd<1>(d, (cudaStream_t)stream);
}

extern "C" int dmm_nchw_to_nhwc_bf16(const float* x, int32_t B, int32_t C, int32_t H, int32_t W, void* out,
                                     int64_t ldo, void* stream) {
    DMM_CHECK(x && out && C > 0 && ldo % 8 == 0 && ldo >= C, "dmm_nchw_to_nhwc_bf16: bad arguments");
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const long long total = (long long)B * H * W * (ldo / 8);
    launch_k(nchw_to_rows_kernel, flat_grid(total, 256), 256, 0, (cudaStream_t)stream, 
        x, B, C, (long long)H * W, reinterpret_cast<__nv_bfloat16*>(out), (int)ldo);
    DMM_LAUNCH_CHECK("nchw_to_rows_kernel");
    return 0;
}

extern "C" int dmm_rows_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t P, int32_t C,
                                    void* stream) {
    DMM_CHECK(src && dst && C > 0 && C % 8 == 0 && lds % 4 == 0 && ldd % 8 == 0, "dmm_rows_f32_to_bf16: bad arguments");
    DMM_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "dmm_rows_f32_to_bf16: pointers must be 16-byte aligned");
    if (P <= 0) return 0;
    launch_k(rows_f32_to_bf16_kernel, flat_grid(P * (C / 8), 256), 256, 0, (cudaStream_t)stream, 
        src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, P, C);
    DMM_LAUNCH_CHECK("rows_f32_to_bf16_kernel");
    return 0;
}

extern "C" int dmm_bce_logits(const float* logits, const float* target, int64_t n, int32_t C, int64_t HW, float* loss,
                              float* grad, double* class_sums, void* stream) {
    DMM_CHECK(logits && target, "dmm_bce_logits: null input");
    DMM_CHECK(n % 4 == 0 && HW % 4 == 0, "dmm_bce_logits: element counts must be multiples of 4 (n=%lld HW=%lld)", (long long)n,
              (long long)HW);
    DMM_CHECK(class_sums == nullptr || (C >= 1 && C <= 8), "dmm_bce_logits: per-class sums support at most 8 classes");
    if (n <= 0) return 0;
    launch_k(bce_logits_kernel, flat_grid(n / 4, 256), 256, 0, (cudaStream_t)stream, logits, target, n / 4, C, HW / 4, loss, grad,
                                                                              class_sums);
    DMM_LAUNCH_CHECK("bce_logits_kernel");
    return 0;
}

static int fill_pack_args(PackArgs& a, int32_t n_valid, int32_t n_rows, int32_t C, int32_t kwidth, int32_t T,
                          const int32_t* tap_off, int64_t sn, int64_t sc) {
    DMM_CHECK(T >= 1 && T <= DMM_MAX_TAPS && tap_off, "weights: bad tap count %d", T);
    DMM_CHECK(kwidth > 0 && C > 0 && n_valid > 0 && n_rows >= n_valid, "weights: bad sizes");
    a.T = T;
    a.C = C;
    a.Kp = (C + kwidth - 1) / kwidth * kwidth;
    a.n_valid = n_valid;
    a.n_rows = n_rows;
    a.ktot = (long long)a.Kp * T;
    a.sn = sn;
    a.sc = sc;
    for (int t = 0; t < T; ++t) a.tap_off[t] = tap_off[t];
    return 0;
}

extern "C" int dmm_pack_weights(const float* w, void* dst, int32_t n_valid, int32_t n_rows, int32_t C, int32_t kwidth,
                                int32_t T, const int32_t* tap_off, int64_t sn, int64_t sc, void* stream) {
    DMM_CHECK(w && dst, "dmm_pack_weights: null pointer");
    PackArgs a;
    int rc = fill_pack_args(a, n_valid, n_rows, C, kwidth, T, tap_off, sn, sc);
    if (rc) return rc;
    launch_k(pack_weights_kernel, flat_grid((long long)n_rows * a.ktot, 256), 256, 0, (cudaStream_t)stream, 
        w, reinterpret_cast<__nv_bfloat16*>(dst), a);
    DMM_LAUNCH_CHECK("pack_weights_kernel");
    return 0;
}

extern "C" int dmm_unpack_wgrad(const float* dw, int64_t dt, int64_t dm, int64_t dn, int32_t M, int32_t N, float* grad,
                                int32_t T, const int32_t* tap_off, int64_t sn, int64_t sc, int32_t accumulate, void* stream) {
    DMM_CHECK(dw && grad, "dmm_unpack_wgrad: null pointer");
    PackArgs a;
    int rc = fill_pack_args(a, N, N, M, 1, T, tap_off, sn, sc);
    if (rc) return rc;
    launch_k(unpack_wgrad_kernel, flat_grid((long long)T * M * N, 256), 256, 0, (cudaStream_t)stream, dw, dt, dm, dn, M, N, grad, a,
                                                                                                accumulate);
    DMM_LAUNCH_CHECK("unpack_wgrad_kernel");
    return 0;
}

extern "C" int dmm_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream) {
    DMM_CHECK(param && grad && exp_avg && exp_avg_sq && step >= 1, "dmm_adam_flat: bad arguments");
    if (n <= 0) return 0;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2 = 1.f - powf(beta2, (float)step);
    launch_k(adam_kernel, flat_grid(n, 256), 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                    weight_decay, bc1, sqrtf(bc2));
    DMM_LAUNCH_CHECK("adam_kernel");
    return 0;
}

extern "C" int dmm_pack_weights_batched(const dmm_pack_job_t* jobs_device, int32_t njobs, void* stream) {
    DMM_CHECK(njobs >= 0 && (njobs == 0 || jobs_device), "dmm_pack_weights_batched: bad arguments");
    if (njobs == 0) return 0;
    launch_k(pack_weights_batched_kernel, dim3(32, (unsigned)njobs, 1), 256, 0, (cudaStream_t)stream, jobs_device);
    DMM_LAUNCH_CHECK("pack_weights_batched_kernel");
    return 0;
}

extern "C" int dmm_unpack_wgrad_batched(const dmm_unpack_job_t* jobs_device, int32_t njobs, void* stream) {
    DMM_CHECK(njobs >= 0 && (njobs == 0 || jobs_device), "dmm_unpack_wgrad_batched: bad arguments");
    if (njobs == 0) return 0;
    launch_k(unpack_wgrad_batched_kernel, dim3(32, (unsigned)njobs, 1), 256, 0, (cudaStream_t)stream, jobs_device);
    DMM_LAUNCH_CHECK("unpack_wgrad_batched_kernel");
    return 0;
}

__global__ void __launch_bounds__(128) bn_fold_kernel(const dmm_bn_fold_job_t* __restrict__ jobs) {
    const dmm_bn_fold_job_t& j = jobs[blockIdx.x];
    for (int c = threadIdx.x; c < j.C; c += blockDim.x) {
        const float sc = (j.gamma ? j.gamma[c] : 1.f) / sqrtf(j.running_var[c] + j.eps);
        j.scale[c] = sc;
        j.shift[c] = (j.beta ? j.beta[c] : 0.f) - j.running_mean[c] * sc;
    }
}

extern "C" int dmm_bn_fold_batched(const dmm_bn_fold_job_t* jobs_device, int32_t njobs, void* stream) {
    DMM_CHECK(njobs >= 0 && (njobs == 0 || jobs_device), "dmm_bn_fold_batched: bad arguments");
    if (njobs == 0) return 0;
    bn_fold_kernel<<<(unsigned)njobs, 128, 0, (cudaStream_t)stream>>>(jobs_device);
    DMM_LAUNCH_CHECK("bn_fold_kernel");
    return 0;
}

extern "C" int dmm_pack_weights_work(const dmm_pack_job_t* jobs_device, const int32_t* work_device, int32_t nwork, int32_t chunk_elems,
                                     void* stream) {
    DMM_CHECK(nwork >= 0 && (nwork == 0 || (jobs_device && work_device)) && chunk_elems >= 256, "dmm_pack_weights_work: bad arguments");
    if (nwork == 0) return 0;
    launch_k(pack_weights_work_kernel, dim3((unsigned)nwork, 1, 1), 256, 0, (cudaStream_t)stream, jobs_device,
             reinterpret_cast<const int2*>(work_device), chunk_elems);
    DMM_LAUNCH_CHECK("pack_weights_work_kernel");
    return 0;
}

extern "C" int dmm_unpack_wgrad_work(const dmm_unpack_job_t* jobs_device, const int32_t* work_device, int32_t nwork, int32_t chunk_elems,
                                      void* stream) {
    DMM_CHECK(nwork >= 0 && (nwork == 0 || (jobs_device && work_device)) && chunk_elems >= 256, "dmm_unpack_wgrad_work: bad arguments");
    if (nwork == 0) return 0;
    launch_k(unpack_wgrad_work_kernel, dim3((unsigned)nwork, 1, 1), 256, 0, (cudaStream_t)stream, jobs_device,
             reinterpret_cast<const int2*>(work_device), chunk_elems);
    DMM_LAUNCH_CHECK("unpack_wgrad_work_kernel");
    return 0;
}

namespace dmm {
// stem: horizontal unfold of the 7x7 / stride-2 convolution.  out[(b, iy, ox)][kw*C + c] = x[c](iy, 2*ox + kw - 3); one thread
// per (image row, output column); the 7 shifted reads of neighbouring threads overlap (L1), the row is written as 16-byte words
// One block per image row (persistent): the C input planes of the row are staged in shared memory with coalesced float4 loads,
// then thread (pixel, 8-column chunk) gathers its 8 values from shared memory and writes one 16-byte word; consecutive threads
// write consecutive words.  CT: compile-time channel count (0 = generic) keeps the column -> (kw, c) split division-free.
template <int CT>
__global__ void __launch_bounds__(256) unfold_w7s2_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2,
                                                          int B, int H, int W, int OW, __nv_bfloat16* __restrict__ out, int ld,
                                                          int K, int stride, int flip) {
    // out[(b, iy, ox)][kw*C + c] = x[c](iy, stride*ox + kw' - K/2), kw' = flip ? K-1-kw : kw   (K <= 7)
    pdl_prologue();
    extern __shared__ float urow[];                       // [C][W + 8]: 3 zero columns left, >= 3 right (padding of the conv)
    const int C = CT > 0 ? CT : C1 + C2;
    const int chunks = ld >> 3;
    const int pitch = W + 8;
    const long long HW = (long long)H * W;
    for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
        const int b = row / H, iy = row - b * H;
        if ((W & 3) == 0 && C <= 4) {
            // float4 loads, all planes of the row in flight at once; margins zeroed separately
            for (int x4 = threadIdx.x; x4 < (W >> 2); x4 += blockDim.x) {
                float4 v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < C)
                        v[c] = __ldg(reinterpret_cast<const float4*>(c < C1 ? x1 + ((long long)b * C1 + c) * HW + (long long)iy * W
                                                                            : x2 + ((long long)b * C2 + (c - C1)) * HW + (long long)iy * W) + x4);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < C) {
                        float* d = urow + c * pitch + 3 + 4 * x4;
                        d[0] = v[c].x; d[1] = v[c].y; d[2] = v[c].z; d[3] = v[c].w;
                    }
            }
            for (int i = threadIdx.x; i < C * 8; i += blockDim.x) {
                const int c = i >> 3, m = i & 7;
                urow[c * pitch + (m < 3 ? m : W + m)] = 0.f;          // columns 0..2 and W+3..W+7
            }
        } else {
            for (int i = threadIdx.x; i < C * pitch; i += blockDim.x) {
                const int c = i / pitch, xx = i - c * pitch - 3;
                float v = 0.f;
                if (xx >= 0 && xx < W)
                    v = c < C1 ? __ldg(x1 + ((long long)b * C1 + c) * HW + (long long)iy * W + xx)
                               : __ldg(x2 + ((long long)b * C2 + (c - C1)) * HW + (long long)iy * W + xx);
                urow[i] = v;
            }
        }
        __syncthreads();
        __nv_bfloat16* orow = out + (long long)row * OW * ld;
        if (CT > 0 && chunks <= 3) {
            // one thread per output pixel, its (at most three) 16-byte chunks unrolled: column -> (kernel column, channel) is
            // then a compile-time split (the generic loop below spent ~100 instructions per chunk on it: ncu r02, 77 % issue-bound)
            const int ncol = K * CT, koff = 3 - (K >> 1);
            for (int ox = threadIdx.x; ox < OW; ox += blockDim.x) {
                const float* u = urow + stride * ox + koff;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    if (ch < chunks) {
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int col = ch * 8 + j, kw = col / (CT > 0 ? CT : 1), c = col - kw * (CT > 0 ? CT : 1);
                            const int kk = flip ? K - 1 - kw : kw;
                            f[j] = col < ncol ? u[c * pitch + kk] : 0.f;
                        }
                        *reinterpret_cast<uint4*>(orow + ((long long)ox * chunks + ch) * 8) = pack8(f);
                    }
                }
            }
        } else
        for (int i = threadIdx.x; i < OW * chunks; i += blockDim.x) {
            const int ox = i / chunks, ch = i - ox * chunks;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = ch * 8 + j;
                float v = 0.f;
                if (col < K * C) {
                    const int kw = col / C, c = col - kw * C;
                    const int kk = flip ? K - 1 - kw : kw;
                    v = urow[c * pitch + stride * ox + kk + (3 - (K >> 1))];      // 3-column left margin
                }
                f[j] = v;
            }
            *reinterpret_cast<uint4*>(orow + (long long)i * 8) = pack8(f);
        }
        __syncthreads();
    }
}
}  // namespace dmm

extern "C" int dmm_unfold_w7s2(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t B, int32_t H, int32_t W, void* out,
                               int64_t ld, void* stream) {
    DMM_CHECK(x1 && out && C1 > 0 && C2 >= 0 && (C2 == 0 || x2), "dmm_unfold_w7s2: bad inputs");
    DMM_CHECK(ld % 8 == 0 && ld >= 7 * (C1 + C2), "dmm_unfold_w7s2: ld=%lld too small / not a multiple of 8", (long long)ld);
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const int OW = (W + 6 - 7) / 2 + 1;
    const int Cc = C1 + C2;
    const size_t usmem = (size_t)Cc * (W + 8) * sizeof(float);
    DMM_CHECK(usmem <= 48 * 1024, "dmm_unfold_w7s2: a row of %d channels x %d pixels does not fit in shared memory", Cc, W);
    const long long urows = (long long)B * H, ucap = (long long)kNumSm * 8;
    const unsigned ugrid = (unsigned)(urows < ucap ? urows : ucap);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (Cc == 1) launch_k(unfold_w7s2_kernel<1>, ugrid, 256, usmem, (cudaStream_t)stream, x1, C1, x2, C2, B, H, W, OW, o, (int)ld, 7, 2, 0);
    else if (Cc == 3) launch_k(unfold_w7s2_kernel<3>, ugrid, 256, usmem, (cudaStream_t)stream, x1, C1, x2, C2, B, H, W, OW, o, (int)ld, 7, 2, 0);
    else if (Cc == 4) launch_k(unfold_w7s2_kernel<4>, ugrid, 256, usmem, (cudaStream_t)stream, x1, C1, x2, C2, B, H, W, OW, o, (int)ld, 7, 2, 0);
    else launch_k(unfold_w7s2_kernel<0>, ugrid, 256, usmem, (cudaStream_t)stream, x1, C1, x2, C2, B, H, W, OW, o, (int)ld, 7, 2, 0);
    DMM_LAUNCH_CHECK("unfold_w7s2_kernel");
    return 0;
}

extern "C" int dmm_dlogits_im2col(const float* dlogits, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K, void* out,
                                  int64_t ld, void* stream) {
    DMM_CHECK(dlogits && out && C > 0 && K >= 1 && (K & 1) && ld % 8 == 0 && ld >= (int64_t)K * K * C,
              "dmm_dlogits_im2col: bad arguments");
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const long long total = (long long)B * H * W * (ld / 8);
    launch_k(dlogits_im2col_kernel, flat_grid(total, 256), 256, 0, (cudaStream_t)stream, dlogits, B, C, H, W, K, K,
                                                                                  reinterpret_cast<__nv_bfloat16*>(out), (int)ld);
    DMM_LAUNCH_CHECK("dlogits_im2col_kernel");
    return 0;
}

namespace dmm {
// one thread per pixel: K*C (<= 16) gathered values -> one 32-byte row; neighbouring threads read neighbouring x (coalesced,
// the K shifted reads of a plane row hit L1)
__global__ void __launch_bounds__(256) dlogits_unfold_w16_kernel(const float* __restrict__ dl, int B, int C, int H, int W, int K,
                                                                 __nv_bfloat16* __restrict__ out) {
    pdl_prologue();
    const long long total = (long long)B * H * W;
    const long long HW = (long long)H * W;
    const int pad = K / 2;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(pix % W);
        const long long t = pix / W;                     // b * H + y
        const long long b = t / H;
        const float* row = dl + b * C * HW + (t - b * H) * W;      // class 0, this image row
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = 0.f;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
            if (kw < K) {
                const int xx = x - (kw - pad);
                if (xx >= 0 && xx < W) {
#pragma unroll
                    for (int n = 0; n < 3; ++n)
                        if (n < C && kw * C + n < 16) f[kw * C + n] = __ldg(row + n * HW + xx);
                }
            }
        }
        float lo[8], hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { lo[j] = f[j]; hi[j] = f[8 + j]; }
        uint4* o = reinterpret_cast<uint4*>(out + pix * 16);
        o[0] = pack8(lo);
        o[1] = pack8(hi);
    }
}
}  // namespace dmm

extern "C" int dmm_dlogits_unfold_w(const float* dlogits, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K, void* out,
                                    int64_t ld, void* stream) {
    DMM_CHECK(dlogits && out && C > 0 && K >= 1 && (K & 1) && ld % 8 == 0 && ld >= (int64_t)K * C,
              "dmm_dlogits_unfold_w: bad arguments");
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    if (K <= 7 && (size_t)C * (W + 8) * sizeof(float) <= 48 * 1024) {
        // row-staged gather (same kernel as the stem unfold, stride 1, columns in flipped order: x - (kw - K/2))
        const size_t usmem = (size_t)C * (W + 8) * sizeof(float);
        const long long urows = (long long)B * H, ucap = (long long)kNumSm * 8;
        const unsigned ugrid = (unsigned)(urows < ucap ? urows : ucap);
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
        const float* nul = nullptr;
        if (C == 3) launch_k(unfold_w7s2_kernel<3>, ugrid, 256, usmem, (cudaStream_t)stream, dlogits, C, nul, 0, B, H, W, W, o, (int)ld, K, 1, 1);
        else launch_k(unfold_w7s2_kernel<0>, ugrid, 256, usmem, (cudaStream_t)stream, dlogits, C, nul, 0, B, H, W, W, o, (int)ld, K, 1, 1);
        DMM_LAUNCH_CHECK("unfold kernel (dlogits)");
        return 0;
    }
    if (ld == 16 && K <= 5 && C <= 3) {
        launch_k(dlogits_unfold_w16_kernel, flat_grid((long long)B * H * W, 256), 256, 0, (cudaStream_t)stream, 
            dlogits, B, C, H, W, K, reinterpret_cast<__nv_bfloat16*>(out));
        DMM_LAUNCH_CHECK("dlogits_unfold_w16_kernel");
        return 0;
    }
    const long long total = (long long)B * H * W * (ld / 8);
    launch_k(dlogits_im2col_kernel, flat_grid(total, 256), 256, 0, (cudaStream_t)stream, dlogits, B, C, H, W, 1, K,
                                                                                  reinterpret_cast<__nv_bfloat16*>(out), (int)ld);
    DMM_LAUNCH_CHECK("dlogits_unfold_w");
    return 0;
}
