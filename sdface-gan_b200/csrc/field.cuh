// Internal (C++) interface between the C-ABI field entry points (field.cu) and the two implementations.
#pragma once
#include "common.cuh"

namespace sdfg {

int field_check_params(const sdfg_field_params* p, uint64_t N);

// fp32 SIMT path (field_f32.cu)
uint64_t field_workspace_bytes_f32(const sdfg_field_params* p, uint64_t N, int save);
int field_forward_f32(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N, float* out_sdf,
                      float* out_rgb, float* out_feat, void* workspace, int save, cudaStream_t st);
int field_backward_f32(const sdfg_field_params* p, const sdfg_field_grads* g, const float* x_in, const float* view_feat,
                       uint64_t N, const float* d_sdf, const float* d_rgb, const float* d_feat, const float* out_feat,
                       const void* workspace, void* scratch, float* d_x_in, cudaStream_t st);

// tcgen05 path (field_tc.cu)
uint64_t field_workspace_bytes_tc(const sdfg_field_params* p, uint64_t N, int save);
int field_forward_tc(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N, float* out_sdf, float* out_rgb,
                     float* out_feat, uint16_t* out_feat16, void* workspace, int save, cudaStream_t st);
uint64_t field_backward_scratch_bytes_tc(const sdfg_field_params* p, uint64_t N);
// eikonal pass: contract d_x_in with the hash encoder's dy_dx inside the chain (tc_bchain3.cuh) instead of writing it
struct EikFuse {
    const float* dy_dx;     // [in_dim / 2 levels * 6, N] component-major (sdfg_grid_encode_forward, D = 3, C = 2)
    float* d_pts;           // [N, 3], pre-zeroed
    float scale;
};
int field_backward_tc(const sdfg_field_params* p, const sdfg_field_grads* g, const float* x_in, const float* view_feat, uint64_t N,
                      const float* d_sdf, const float* d_rgb, const float* d_feat, const void* workspace, void* scratch, float* d_x_in,
                      cudaStream_t st, int phases = SDFG_BWD_BOTH, const EikFuse* eik = nullptr);

}  // namespace sdfg
