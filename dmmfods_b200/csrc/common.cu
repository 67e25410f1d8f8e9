#include <stdlib.h>
// Host-side plumbing shared by all kernels: thread-local error string, CUDA error mapping and
// TMA tensor-map encoding (cuTensorMapEncodeTiled resolved at run time through
// cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda and can be
// dlopen'ed on a CPU-only box for symbol checks).
#include "common.cuh"
#include "../../include/dmmfods_b200.h"

#include <stdarg.h>
#include <string.h>

namespace dmm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return -2;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        (void)cudaGetLastError();
    }
    return fn;
}

// L2 promotion of the bf16 tensor maps (DMM_TMA_L2_PROMO = 0 / 64 / 128 / 256, default 256): with 256 B a 128-byte box row of a
// narrow channel slice of a wide block buffer pulls a second, unused 128 bytes from DRAM (ncu: 2x reads on 64-channel slices).
static CUtensorMapL2promotion l2_promotion() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DMM_TMA_L2_PROMO");
        v = (e && *e) ? atoi(e) : 256;
    }
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                  : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
}

int make_tmap_elem(CUtensorMap* out, int elem_bytes, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_elems, const uint32_t* box, int swizzle_bytes) {
    DMM_CHECK(elem_bytes == 2 || elem_bytes == 4, "tensor map element size %d", elem_bytes);
    EncodeTiledFn enc = get_encode();
    DMM_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    DMM_CHECK(rank >= 2 && rank <= 5, "tensor map rank %d", rank);
    DMM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p is not 16-byte aligned", base);
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        DMM_CHECK(dims[i] >= 1 && box[i] >= 1 && box[i] <= 256, "tensor map dim %d: size %llu box %u", i,
                  (unsigned long long)dims[i], box[i]);
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_elems[i] * (unsigned long long)elem_bytes;
        DMM_CHECK(gstr[i] % 16 == 0 && gstr[i] > 0, "tensor map stride %d = %llu bytes is not a positive multiple of 16", i,
                  (unsigned long long)gstr[i]);
    }
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
    else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
    else if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
    DMM_CHECK((int)(box[0] * elem_bytes) <= (swizzle_bytes ? swizzle_bytes : 512), "tensor map inner box %u elements exceeds the swizzle span",
              box[0]);
    CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2_promotion(),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DMM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu.. box %u,%u..)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return 0;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_elems, const uint32_t* box, int swizzle_bytes) {
    return make_tmap_elem(out, 2, base, rank, dims, strides_elems, box, swizzle_bytes);
}

int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_elems, uint32_t box_cols,
                     uint32_t box_rows, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode();
    DMM_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    DMM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "fp32 tensor map base %p is not 16-byte aligned", base);
    DMM_CHECK((row_stride_elems * 4ull) % 16 == 0 && row_stride_elems > 0, "fp32 tensor map row stride %llu elements",
              (unsigned long long)row_stride_elems);
    DMM_CHECK(box_cols >= 1 && box_cols <= 256 && box_rows >= 1 && box_rows <= 256, "fp32 tensor map box %u x %u", box_cols, box_rows);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {row_stride_elems * 4ull};
    cuuint32_t bx[2] = {box_cols, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DMM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d (%llu x %llu, box %u x %u)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, box_cols, box_rows);
    return 0;
}

}  // namespace dmm

extern "C" const char* dmm_last_error(void) { return dmm::g_err; }

extern "C" int dmm_version(void) { return 100; }

extern "C" int dmm_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return 0;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    return prop.major == 10 ? 1 : 0;
}

/* sizeof() of every struct that crosses the C-ABI, so that a binding can verify its own layout. */
extern "C" int dmm_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(dmm_view_t);
        case 1: return (int)sizeof(dmm_igemm_t);
        case 2: return (int)sizeof(dmm_wgrad_t);
        case 3: return (int)sizeof(dmm_bn_t);
        case 4: return (int)sizeof(dmm_bn_apply_t);
        case 5: return (int)sizeof(dmm_bn_bwd_t);
        case 6: return (int)sizeof(dmm_bn_bwd_args_t);
        case 7: return (int)sizeof(dmm_head_t);
        case 8: return (int)sizeof(dmm_head_bwd_t);
        case 9: return (int)sizeof(dmm_pack_job_t);
        case 10: return (int)sizeof(dmm_unpack_job_t);
        case 11: return (int)sizeof(dmm_grad_gather_t);
        case 12: return (int)sizeof(dmm_bn_fold_job_t);
        default: return -1;
    }
}

namespace dmm {
bool pdl_enabled() {
    static const bool on = getenv("DMM_PDL") && atoi(getenv("DMM_PDL")) != 0;      // off by default: measured 0.7 % slower (85.0 vs 84.4 ms)
    return on;
}
}  // namespace dmm
