// Frustum -> box resampling of the SDF volume (ref align_volume, sdf_utils.py:164-184): the reference builds a [H,W,D,3] sampling
// grid with torch.meshgrid, scales x/y by linspace(far/near, 1, D) along depth, resamples with F.grid_sample (trilinear,
// align_corners=True, border padding) through two permuting copies and overwrites the out-of-frustum cells with 1.  Here: one
// gather kernel, thread per output cell (z fastest = contiguous in [B,H,W,D,C]), no grid tensor, no permutes.
// HBM-bound: 4*C bytes out per cell, ~4*C bytes in (the 8 corners of neighbouring cells share sectors).
#include "common.cuh"

namespace sdfg {

// torch.linspace(a, b, n)[i] as torch computes it on CPU/CUDA: step = (b - a) / (n - 1); first half a + i*step, second half b - (n-1-i)*step
__device__ __forceinline__ float linspace_at(float a, float b, uint32_t n, uint32_t i) {
    if (n == 1) return a;
    const float step = (b - a) / (float)(n - 1);
    return i < n / 2 ? a + step * (float)i : b - step * (float)(n - 1 - i);
}

__global__ void __launch_bounds__(256) align_volume_kernel(const float* __restrict__ vol, float* __restrict__ out, uint32_t B, uint32_t H,
                                                            uint32_t W, uint32_t D, uint32_t C, float ratio) {
    const uint64_t cell = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)B * H * W * D;
    if (cell >= total) return;
    const uint32_t z = (uint32_t)(cell % D);
    const uint32_t x = (uint32_t)((cell / D) % W);
    const uint32_t y = (uint32_t)((cell / ((uint64_t)D * W)) % H);
    const uint32_t b = (uint32_t)(cell / ((uint64_t)D * W * H));
    const float k = linspace_at(ratio, 1.f, D, z);
    const float gx = linspace_at(-1.f, 1.f, W, x) * k, gy = linspace_at(-1.f, 1.f, H, y) * k, gz = linspace_at(-1.f, 1.f, D, z);
    float* o = out + cell * C;
    if (gx < -1.f || gx > 1.f || gy < -1.f || gy > 1.f || gz < -1.f || gz > 1.f) {
        for (uint32_t c = 0; c < C; c++) o[c] = 1.f;
        return;
    }
    // grid_sample, align_corners = True: pixel = (g + 1) / 2 * (size - 1); border padding clamps the coordinate
    auto unnorm = [](float g, uint32_t n) { return fminf(fmaxf((g + 1.f) * 0.5f * (float)(n - 1), 0.f), (float)(n - 1)); };
    const float fx = unnorm(gx, W), fy = unnorm(gy, H), fz = unnorm(gz, D);
    const float x0f = floorf(fx), y0f = floorf(fy), z0f = floorf(fz);
    const float tx = fx - x0f, ty = fy - y0f, tz = fz - z0f;
    const uint32_t x0 = (uint32_t)x0f, y0 = (uint32_t)y0f, z0 = (uint32_t)z0f;
    const uint32_t x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1), z1 = min(z0 + 1, D - 1);
    const float* vb = vol + (uint64_t)b * H * W * D * C;
    auto at = [&](uint32_t yy, uint32_t xx, uint32_t zz, uint32_t c) { return __ldg(vb + (((uint64_t)yy * W + xx) * D + zz) * C + c); };
    for (uint32_t c = 0; c < C; c++) {
        // weights in torch's order: (1-tx)(1-ty)(1-tz) ... over the corners (x: W axis, y: H axis, z: D axis)
        const float v = at(y0, x0, z0, c) * ((1.f - tx) * (1.f - ty) * (1.f - tz)) + at(y0, x1, z0, c) * (tx * (1.f - ty) * (1.f - tz)) +
                        at(y1, x0, z0, c) * ((1.f - tx) * ty * (1.f - tz)) + at(y1, x1, z0, c) * (tx * ty * (1.f - tz)) +
                        at(y0, x0, z1, c) * ((1.f - tx) * (1.f - ty) * tz) + at(y0, x1, z1, c) * (tx * (1.f - ty) * tz) +
                        at(y1, x0, z1, c) * ((1.f - tx) * ty * tz) + at(y1, x1, z1, c) * (tx * ty * tz);
        o[c] = v;
    }
}

}  // namespace sdfg

extern "C" int sdfg_align_volume(const float* volume, float* out, uint32_t B, uint32_t H, uint32_t W, uint32_t D, uint32_t C, float near_,
                                 float far_, void* stream) {
    using namespace sdfg;
    const uint64_t total = (uint64_t)B * H * W * D;
    if (total == 0 || C == 0) return SDFG_OK;
    SDFG_REQUIRE(volume && out && volume != out, SDFG_ERR_INVALID, "align_volume: null or aliased pointer");
    SDFG_REQUIRE(near_ > 0.f && far_ > 0.f, SDFG_ERR_INVALID, "align_volume: near and far must be positive");
    align_volume_kernel<<<(unsigned)ceil_div<uint64_t>(total, 256), 256, 0, (cudaStream_t)stream>>>(volume, out, B, H, W, D, C, far_ / near_);
    return check_launch("align_volume_kernel");
}
