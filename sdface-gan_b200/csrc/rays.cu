// Ray generation, depth sampling and sample-point construction in ONE launch.
//
// Replaces ~15 elementwise torch launches and ~300 MB of temporaries of the reference
// (ref = /root/reference/im2scene/sdf/models/sdf_model.py):
//   get_rays        :207-222 (+ pixel-centre buffers i,j :167-171)
//   viewdir norm    :367
//   z_vals          :324, offset jitter :326-331,340, stratified jitter :332-340 (and mlp_init_pass :389-396)
//   pts, normalized :343, :348-351
// One thread per SAMPLE so every store is coalesced; the per-ray quantities (direction, rotation) are a handful of
// FMAs and are simply recomputed per sample.  Products and sums are kept un-contracted (__fmul_rn/__fadd_rn) where the
// reference evaluates them as separate torch ops, so sample positions agree with torch to the last bit or two.
#include "common.cuh"

namespace sdfg {

__global__ void __launch_bounds__(256) sample_rays_kernel(const float* __restrict__ c2w, const float* __restrict__ focal,
                                                          const float* __restrict__ near, const float* __restrict__ far,
                                                          const float* __restrict__ t_vals, const float* __restrict__ t_rand,
                                                          int jitter_mode, int static_viewdirs, int z_normalize, uint32_t B,
                                                          uint32_t R, uint32_t S, float* __restrict__ z_vals,
                                                          float* __restrict__ pts, float* __restrict__ npts,
                                                          float* __restrict__ viewdirs, float* __restrict__ rays_d) {
    const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)B * R * R * S;
    if (n >= total) return;
    const uint32_t k = (uint32_t)(n % S);
    const size_t ray = n / S;
    const uint32_t x = (uint32_t)(ray % R);            // column -> i = x + 0.5
    const uint32_t y = (uint32_t)((ray / R) % R);      // row    -> j = y + 0.5
    const uint32_t b = (uint32_t)(ray / ((size_t)R * R));

    const float f = __ldg(focal + b), nr = __ldg(near + b), fr = __ldg(far + b);
    const float half = (float)R * 0.5f;
    float dir[3];
    dir[0] = __fdiv_rn(((float)x + 0.5f) - half, f);
    dir[1] = -__fdiv_rn(((float)y + 0.5f) - half, f);
    dir[2] = -1.0f;
    const float* m = c2w + (size_t)b * 12;
    float rd[3], ro[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        // torch.sum(dirs[..., None, :] * c2w[:, :3, :3], -1): three rounded products, summed left to right
        rd[a] = __fadd_rn(__fadd_rn(__fmul_rn(dir[0], __ldg(m + a * 4 + 0)), __fmul_rn(dir[1], __ldg(m + a * 4 + 1))),
                          __fmul_rn(dir[2], __ldg(m + a * 4 + 2)));
        ro[a] = __ldg(m + a * 4 + 3);
    }

    // depth of this sample
    const float t = __ldg(t_vals + k);
    float z = __fadd_rn(__fmul_rn(nr, 1.f - t), __fmul_rn(fr, t));
    if (jitter_mode != 0) {
        float lower, upper, u;
        if (jitter_mode == 1) {   // one offset per ray; the interval of the last sample ends at `far`
            lower = z;
            if (k + 1 < S) { const float t1 = __ldg(t_vals + k + 1); upper = __fadd_rn(__fmul_rn(nr, 1.f - t1), __fmul_rn(fr, t1)); }
            else upper = fr;
            u = __ldg(t_rand + ray);
        } else {                  // stratified: between the mid-points of neighbouring samples
            float zp = z, zn = z;
            if (k > 0) { const float t0 = __ldg(t_vals + k - 1); zp = __fadd_rn(__fmul_rn(nr, 1.f - t0), __fmul_rn(fr, t0)); }
            if (k + 1 < S) { const float t1 = __ldg(t_vals + k + 1); zn = __fadd_rn(__fmul_rn(nr, 1.f - t1), __fmul_rn(fr, t1)); }
            lower = k > 0 ? 0.5f * __fadd_rn(z, zp) : z;
            upper = k + 1 < S ? 0.5f * __fadd_rn(zn, z) : z;
            u = __ldg(t_rand + n);
        }
        z = __fadd_rn(lower, __fmul_rn(upper - lower, u));
    }
    if (z_vals) z_vals[n] = z;

    const float span = fr - nr;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float p = __fadd_rn(ro[a], __fmul_rn(rd[a], z));
        if (pts) pts[n * 3 + a] = p;
        if (npts) npts[n * 3 + a] = z_normalize ? __fdiv_rn(p * 2.f, span) : p;
    }

    if (k == 0) {
        if (rays_d) { rays_d[ray * 3 + 0] = rd[0]; rays_d[ray * 3 + 1] = rd[1]; rays_d[ray * 3 + 2] = rd[2]; }
        if (viewdirs) {
            const float* v = static_viewdirs ? dir : rd;
            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
            viewdirs[ray * 3 + 0] = __fdiv_rn(v[0], nrm);
            viewdirs[ray * 3 + 1] = __fdiv_rn(v[1], nrm);
            viewdirs[ray * 3 + 2] = __fdiv_rn(v[2], nrm);
        }
    }
}

}  // namespace sdfg

extern "C" int sdfg_sample_rays(const float* c2w, const float* focal, const float* near, const float* far, const float* t_vals,
                                const float* t_rand, int jitter_mode, int static_viewdirs, int z_normalize, uint32_t B,
                                uint32_t R, uint32_t S, float* z_vals, float* pts, float* npts, float* viewdirs, float* rays_d,
                                void* stream) {
    using namespace sdfg;
    SDFG_REQUIRE(c2w && focal && near && far && t_vals, SDFG_ERR_INVALID, "sample_rays: null camera pointer");
    SDFG_REQUIRE(jitter_mode >= 0 && jitter_mode <= 2, SDFG_ERR_INVALID, "sample_rays: jitter_mode must be 0, 1 or 2");
    SDFG_REQUIRE(jitter_mode == 0 || t_rand, SDFG_ERR_INVALID, "sample_rays: jitter requested without t_rand");
    SDFG_REQUIRE(R >= 1 && S >= 1, SDFG_ERR_INVALID, "sample_rays: R and S must be positive");
    const size_t total = (size_t)B * R * R * S;
    if (total == 0) return SDFG_OK;
    SDFG_REQUIRE(total < ((size_t)1 << 40), SDFG_ERR_UNSUPPORTED, "sample_rays: too many samples");
    sample_rays_kernel<<<(unsigned)ceil_div<size_t>(total, 256), 256, 0, (cudaStream_t)stream>>>(
        c2w, focal, near, far, t_vals, t_rand, jitter_mode, static_viewdirs, z_normalize, B, R, S, z_vals, pts, npts, viewdirs,
        rays_d);
    return check_launch("sample_rays_kernel");
}
