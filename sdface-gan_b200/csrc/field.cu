// C-ABI entry points of the style-modulated SIREN field (include/sdfg.h): argument checks + precision dispatch.
#include "field.cuh"

using namespace sdfg;

extern "C" uint64_t sdfg_field_workspace_bytes(const sdfg_field_params* p, uint64_t N, int save_for_backward, int precision) {
    if (!p) return 0;
    if (precision == SDFG_PRECISION_FP32) return field_workspace_bytes_f32(p, N, save_for_backward);
    if (precision == SDFG_PRECISION_TC16) return field_workspace_bytes_tc(p, N, save_for_backward);
    return 0;
}

extern "C" uint64_t sdfg_field_backward_scratch_bytes(const sdfg_field_params* p, uint64_t N, int precision) {
    if (!p) return 0;
    if (precision == SDFG_PRECISION_TC16) return field_backward_scratch_bytes_tc(p, N);
    return 2ull * N * p->width * sizeof(float);
}

extern "C" int sdfg_field_forward(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N,
                                  float* out_sdf, float* out_rgb, float* out_feat, void* workspace, int save_for_backward,
                                  int precision, void* stream) {
    if (int e = field_check_params(p, N)) return e;
    if (N == 0) return SDFG_OK;
    SDFG_REQUIRE(x_in && workspace, SDFG_ERR_INVALID, "field_forward: null pointer");
    SDFG_REQUIRE(out_sdf || out_rgb || out_feat, SDFG_ERR_INVALID, "field_forward: no output requested");
    if (precision == SDFG_PRECISION_FP32)
        return field_forward_f32(p, x_in, view_feat, N, out_sdf, out_rgb, out_feat, workspace, save_for_backward, (cudaStream_t)stream);
    if (precision == SDFG_PRECISION_TC16)
        return field_forward_tc(p, x_in, view_feat, N, out_sdf, out_rgb, out_feat, nullptr, workspace, save_for_backward, (cudaStream_t)stream);
    return set_error(SDFG_ERR_UNSUPPORTED, "field_forward: unknown precision %d", precision);
}

extern "C" int sdfg_field_forward_h(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N,
                                    float* out_sdf, float* out_rgb, uint16_t* out_feat16, void* workspace, void* stream) {
    if (int e = field_check_params(p, N)) return e;
    if (N == 0) return SDFG_OK;
    SDFG_REQUIRE(x_in && workspace && out_feat16, SDFG_ERR_INVALID, "field_forward_h: null pointer");
    return field_forward_tc(p, x_in, view_feat, N, out_sdf, out_rgb, nullptr, out_feat16, workspace, 0, (cudaStream_t)stream);
}

extern "C" int sdfg_field_backward(const sdfg_field_params* p, const sdfg_field_grads* g, const float* x_in, const float* view_feat,
                                   uint64_t N, const float* d_sdf, const float* d_rgb, const float* d_feat, const float* out_feat,
                                   const void* workspace, void* scratch, float* d_x_in, int precision, void* stream) {
    if (int e = field_check_params(p, N)) return e;
    if (N == 0) return SDFG_OK;
    SDFG_REQUIRE(x_in && workspace && scratch, SDFG_ERR_INVALID, "field_backward: null pointer");
    if (precision == SDFG_PRECISION_FP32)
        return field_backward_f32(p, g, x_in, view_feat, N, d_sdf, d_rgb, d_feat, out_feat, workspace, scratch, d_x_in, (cudaStream_t)stream);
    if (precision == SDFG_PRECISION_TC16)
        return field_backward_tc(p, g, x_in, view_feat, N, d_sdf, d_rgb, d_feat, workspace, scratch, d_x_in, (cudaStream_t)stream);
    return set_error(SDFG_ERR_UNSUPPORTED, "field_backward: unknown precision %d", precision);
}

extern "C" int sdfg_field_backward_phase(const sdfg_field_params* p, const sdfg_field_grads* g, const float* x_in, const float* view_feat,
                                         uint64_t N, const float* d_sdf, const float* d_rgb, const float* d_feat, const float* out_feat,
                                         const void* workspace, void* scratch, float* d_x_in, int precision, int phases, void* stream) {
    SDFG_REQUIRE((phases & SDFG_BWD_BOTH) && !(phases & ~SDFG_BWD_BOTH), SDFG_ERR_INVALID, "field_backward_phase: phases must be CHAIN, WGRAD or both");
    if (precision != SDFG_PRECISION_TC16) {      // the fp32 path interleaves both halves: all of it runs with the CHAIN phase
        if (!(phases & SDFG_BWD_CHAIN)) return field_check_params(p, N);
        return sdfg_field_backward(p, g, x_in, view_feat, N, d_sdf, d_rgb, d_feat, out_feat, workspace, scratch, d_x_in, precision, stream);
    }
    if (int e = field_check_params(p, N)) return e;
    if (N == 0) return SDFG_OK;
    SDFG_REQUIRE(x_in && workspace && scratch, SDFG_ERR_INVALID, "field_backward: null pointer");
    return field_backward_tc(p, g, x_in, view_feat, N, d_sdf, d_rgb, d_feat, workspace, scratch, d_x_in, (cudaStream_t)stream, phases);
}

extern "C" int sdfg_field_eikonal(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N, const float* d_sdf,
                                  const void* workspace, void* scratch, const float* dy_dx, uint32_t D, uint32_t C, float scale, float* d_pts,
                                  int precision, void* stream) {
    SDFG_REQUIRE(precision == SDFG_PRECISION_TC16, SDFG_ERR_UNSUPPORTED, "field_eikonal: tensor-core path only (use sdfg_field_backward + sdfg_grid_encode_backward)");
    SDFG_REQUIRE(D == 3 && C == 2, SDFG_ERR_UNSUPPORTED, "field_eikonal: 3-D points and 2 features per level (got D = %u, C = %u)", D, C);
    if (int e = field_check_params(p, N)) return e;
    if (N == 0) return SDFG_OK;
    SDFG_REQUIRE(x_in && workspace && scratch && d_sdf && dy_dx && d_pts, SDFG_ERR_INVALID, "field_eikonal: null pointer");
    const EikFuse eik = {dy_dx, d_pts, scale};
    return field_backward_tc(p, nullptr, x_in, view_feat, N, d_sdf, nullptr, nullptr, workspace, scratch, nullptr, (cudaStream_t)stream,
                             SDFG_BWD_BOTH, &eik);
}
