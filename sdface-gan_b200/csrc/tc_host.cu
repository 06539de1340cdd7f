// Host side of the tensor-core path: TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#include <cudaTypedefs.h>

#include "tc_common.cuh"

namespace sdfg {

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = []() -> PFN_cuTensorMapEncodeTiled_v12000 {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }();
    return fn;
}

int make_tensor_map_16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                       uint32_t box_cols, uint32_t fmt) {
    PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
    SDFG_REQUIRE(fn, SDFG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    SDFG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, SDFG_ERR_INVALID,
                 "tensor map: base and row pitch must be 16-byte aligned (ld = %llu)", (unsigned long long)ld);
    SDFG_REQUIRE(box_cols * 2 <= 128 && box_rows <= 256, SDFG_ERR_INVALID, "tensor map: box too large");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, fmt == tc::FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDFG_REQUIRE(r == CUDA_SUCCESS, SDFG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r,
                 (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    return SDFG_OK;
}

// channels-last fp16 activation [B, H, W, C] as a 4-D tensor (C, W, H, B); box = 64 channels x box_w x box_h pixels of one sample,
// 128B swizzle: a loaded box is a K-major [box_w * box_h pixels x 64 channels] operand tile; coordinates outside the image
// (negative or past the edge) are zero-filled -- the zero padding of a convolution
int make_tensor_map_16_4d(CUtensorMap* out, const void* base, uint32_t B, uint32_t H, uint32_t W, uint32_t C, uint32_t box_w, uint32_t box_h) {
    PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
    SDFG_REQUIRE(fn, SDFG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    SDFG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && C % 8 == 0, SDFG_ERR_INVALID, "tensor map (4d): base must be 16-byte aligned, channels a multiple of 8");
    SDFG_REQUIRE(box_w <= 256 && box_h <= 256 && C >= 64, SDFG_ERR_INVALID, "tensor map (4d): box too large or fewer than 64 channels");
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, box_w, box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDFG_REQUIRE(r == CUDA_SUCCESS, SDFG_ERR_CUDA, "cuTensorMapEncodeTiled (4d) failed (%d) B=%u H=%u W=%u C=%u box %ux%u", (int)r, B, H, W, C, box_w, box_h);
    return SDFG_OK;
}

}  // namespace sdfg
