// SDF -> density -> alpha -> front-to-back compositing along each ray, forward and backward, one warp per ray.
//
// Behavioural contract (ref = /root/reference/im2scene/sdf/models/sdf_model.py):
//   dists                 :238-241   (z[k+1]-z[k]) * |d|, last = 1e10 * |d|
//   sdf_activation        :231-234   sigma = sigmoid(-sdf/beta)/beta
//   alpha                 :262 (sdf) / :267 (density branch: softplus(raw + noise))
//   visibility, weights   :269-272   T = exclusive cumprod(1 - alpha + 1e-10), w = alpha*T
//   force_background      :279-280   w[S-1] = 1 - sum_{k<S-1} w[k]
//   rgb / feature / xyz / mask maps :282-296
// The reference runs this as ~20 torch launches with six [N,1] temporaries and reads the [N,256] features twice; here
// the scan is a warp shuffle scan (lane = sample, up to 4 samples per lane) and the feature reduction streams each
// sample's 1 KB row exactly once with the lanes across channels.  HBM-bound: algorithmic bytes per sample are
// 4*(1 + 3 + F + 1 [+3]) in, plus 4*(3 + F [+4]) per ray out.
#include "common.cuh"

namespace sdfg {

constexpr int kMaxPerLane = 8;   // samples per lane -> S <= 256 (sdf_mesh.py renders 128- and 256-sample frusta)
constexpr int kMaxF4 = 4;        // float4 per lane over the channels -> F <= 512

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // torch threshold = 20

struct RayState {
    float alpha[kMaxPerLane], w[kMaxPerLane], T[kMaxPerLane], dist[kMaxPerLane], x[kMaxPerLane];
};

// alpha / transmittance / weights of the lane's samples [lane*KS, lane*KS+KS)
template <int KS>
__device__ __forceinline__ void ray_weights(const float* __restrict__ sdf, const float* __restrict__ z,
                                            const float* __restrict__ noise, float dnorm, float beta, uint32_t S,
                                            int with_sdf, int force_background, int lane, RayState& st) {
    const float inv_beta = 1.f / beta;
    float u_prod = 1.f;
#pragma unroll
    for (int i = 0; i < KS; i++) {
        const uint32_t s = lane * KS + i;
        float a = 0.f, dist = 0.f, xin = 0.f;
        if (s < S) {
            const float z0 = __ldg(z + s);
            dist = (s + 1 < S ? __ldg(z + s + 1) - z0 : 1e10f) * dnorm;
            xin = __ldg(sdf + s);
            float sigma;
            if (with_sdf) sigma = sigmoidf_(-xin * inv_beta) * inv_beta;
            else sigma = softplusf_(xin + (noise ? __ldg(noise + s) : 0.f));
            a = 1.f - expf(-sigma * dist);
        }
        st.alpha[i] = a;
        st.dist[i] = dist;
        st.x[i] = xin;
        st.T[i] = u_prod;                                  // local exclusive product
        u_prod *= s < S ? (1.f - a + 1e-10f) : 1.f;
    }
    // exclusive product scan of the per-lane products
    float incl = u_prod;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= v;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    float wsum = 0.f;
#pragma unroll
    for (int i = 0; i < KS; i++) {
        st.T[i] *= excl;
        st.w[i] = st.alpha[i] * st.T[i];
        const uint32_t s = lane * KS + i;
        if (s + 1 < S) wsum += st.w[i];
    }
    if (force_background) {
        wsum = warp_sum(wsum);
#pragma unroll
        for (int i = 0; i < KS; i++)
            if ((uint32_t)(lane * KS + i) == S - 1) st.w[i] = 1.f - wsum;
    }
}

// FEAT16: `feat` points at fp16 features (the tensor-core field's inference output), 8 bytes per lane and sample instead of 16
template <int KS, bool FEAT16 = false>
__global__ void __launch_bounds__(256) composite_forward_kernel(
    const float* __restrict__ sdf, const float* __restrict__ rgb, const float* __restrict__ feat, const float* __restrict__ z_vals,
    const float* __restrict__ rays_d, const float* __restrict__ pts, const float* __restrict__ noise,
    const float* __restrict__ sigmoid_beta, uint64_t NR, uint32_t S, uint32_t F, int with_sdf, int force_background,
    float* __restrict__ rgb_map, float* __restrict__ feat_map, float* __restrict__ xyz_map, float* __restrict__ mask,
    float* __restrict__ weights) {
    const int lane = threadIdx.x & 31;
    const uint64_t ray = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= NR) return;
    const float beta = with_sdf ? __ldg(sigmoid_beta) : 1.f;
    const float dx = __ldg(rays_d + ray * 3), dy = __ldg(rays_d + ray * 3 + 1), dz = __ldg(rays_d + ray * 3 + 2);
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const size_t base = (size_t)ray * S;
    RayState st;
    ray_weights<KS>(sdf + base, z_vals + base, noise ? noise + base : nullptr, dnorm, beta, S, with_sdf, force_background, lane, st);

    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // rgb, xyz
#pragma unroll
    for (int i = 0; i < KS; i++) {
        const uint32_t s = lane * KS + i;
        if (s < S) {
            const float w = st.w[i];
            if (weights) weights[base + s] = w;
            if (rgb) {                                          // NULL: sdf-only query (xyz / mask from the weights alone)
                const float* c = rgb + (base + s) * 3;
                acc[0] = fmaf(w, sigmoidf_(__ldg(c)), acc[0]);
                acc[1] = fmaf(w, sigmoidf_(__ldg(c + 1)), acc[1]);
                acc[2] = fmaf(w, sigmoidf_(__ldg(c + 2)), acc[2]);
            }
            if (xyz_map) {
                const float* p = pts + (base + s) * 3;
                acc[3] = fmaf(w, __ldg(p), acc[3]);
                acc[4] = fmaf(w, __ldg(p + 1), acc[4]);
                acc[5] = fmaf(w, __ldg(p + 2), acc[5]);
            }
            if (mask && s == S - 1) mask[ray] = w;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
        if (rgb_map) {
            rgb_map[ray * 3 + 0] = -1.f + 2.f * acc[0];
            rgb_map[ray * 3 + 1] = -1.f + 2.f * acc[1];
            rgb_map[ray * 3 + 2] = -1.f + 2.f * acc[2];
        }
        if (xyz_map) { xyz_map[ray * 3] = acc[3]; xyz_map[ray * 3 + 1] = acc[4]; xyz_map[ray * 3 + 2] = acc[5]; }
    }
    if (FEAT16 && feat_map && F == 256) {
        // the tensor-core field's fp16 features at the model width: a lane owns 8 consecutive channels = one 16-byte load per sample
        // (a warp reads the 512-byte row in one instruction), 4 samples per trip
        float fa8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const uint4* frow = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(feat) + base * F) + lane;
#pragma unroll 4
        for (uint32_t s = 0; s < S; s++) {
            float w = 0.f;
#pragma unroll
            for (int i = 0; i < KS; i++) {
                const float wi = __shfl_sync(0xffffffffu, st.w[i], (int)(s / KS));
                if ((int)(s % KS) == i) w = wi;
            }
            const uint4 h = __ldg(frow + (size_t)s * 32);
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[k]));
                fa8[2 * k] = fmaf(w, a.x, fa8[2 * k]); fa8[2 * k + 1] = fmaf(w, a.y, fa8[2 * k + 1]);
            }
        }
        float4* orow = reinterpret_cast<float4*>(feat_map + (size_t)ray * F) + 2 * lane;
        orow[0] = make_float4(fa8[0], fa8[1], fa8[2], fa8[3]);
        orow[1] = make_float4(fa8[4], fa8[5], fa8[6], fa8[7]);
    } else if (feat_map) {
        // lanes across channels, samples streamed in order; w broadcast from the lane that owns the sample
        const uint32_t F4 = F >> 2;
        float4 fa[kMaxF4];
#pragma unroll
        for (int j = 0; j < kMaxF4; j++) fa[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* frow = reinterpret_cast<const float4*>(feat + base * F);
        const uint2* frow16 = reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(feat) + base * F);
        // 4 samples per trip: 8 independent loads per lane in flight (one sample at a time left the stream latency-bound at ~60 % of HBM)
#pragma unroll 4
        for (uint32_t s = 0; s < S; s++) {
            float w = 0.f;
#pragma unroll
            for (int i = 0; i < KS; i++) {
                const float wi = __shfl_sync(0xffffffffu, st.w[i], (int)(s / KS));
                if ((int)(s % KS) == i) w = wi;
            }
#pragma unroll
            for (int j = 0; j < kMaxF4; j++) {
                const uint32_t f4 = lane + 32 * j;
                if (f4 < F4) {
                    float4 v;
                    if (FEAT16) {
                        const uint2 h = __ldg(frow16 + (size_t)s * F4 + f4);
                        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
                        v = make_float4(a.x, a.y, b.x, b.y);
                    } else {
                        v = ldg_stream4(frow + (size_t)s * F4 + f4);
                    }
                    fa[j].x = fmaf(w, v.x, fa[j].x); fa[j].y = fmaf(w, v.y, fa[j].y);
                    fa[j].z = fmaf(w, v.z, fa[j].z); fa[j].w = fmaf(w, v.w, fa[j].w);
                }
            }
        }
        float4* orow = reinterpret_cast<float4*>(feat_map + (size_t)ray * F);
#pragma unroll
        for (int j = 0; j < kMaxF4; j++) {
            const uint32_t f4 = lane + 32 * j;
            if (f4 < F4) orow[f4] = fa[j];
        }
    }
}

// reverse inclusive scan of affine maps f_k(x) = c_k + m_k x under composition (f_lane o f_lane+1 o ...)
__device__ __forceinline__ void affine_suffix_scan(float& c, float& m, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float c2 = __shfl_down_sync(0xffffffffu, c, o);
        const float m2 = __shfl_down_sync(0xffffffffu, m, o);
        if (lane + o < 32) { c = fmaf(m, c2, c); m *= m2; }
    }
}

template <int KS>
__global__ void __launch_bounds__(256) composite_backward_kernel(
    const float* __restrict__ sdf, const float* __restrict__ rgb, const float* __restrict__ feat, const float* __restrict__ z_vals,
    const float* __restrict__ rays_d, const float* __restrict__ pts, const float* __restrict__ noise,
    const float* __restrict__ sigmoid_beta, uint64_t NR, uint32_t S, uint32_t F, int with_sdf, int force_background,
    const float* __restrict__ d_rgb_map, const float* __restrict__ d_feat_map, const float* __restrict__ d_xyz_map,
    const float* __restrict__ d_mask, float* __restrict__ d_sdf, float* __restrict__ d_rgb, float* __restrict__ d_feat,
    float* __restrict__ d_pts, float* __restrict__ d_sigmoid_beta) {
    __shared__ float beta_partial[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t ray = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    float dbeta = 0.f;
    if (ray < NR) {
        const float beta = with_sdf ? __ldg(sigmoid_beta) : 1.f;
        const float inv_beta = 1.f / beta;
        const float dx = __ldg(rays_d + ray * 3), dy = __ldg(rays_d + ray * 3 + 1), dz = __ldg(rays_d + ray * 3 + 2);
        const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
        const size_t base = (size_t)ray * S;
        RayState st;
        ray_weights<KS>(sdf + base, z_vals + base, noise ? noise + base : nullptr, dnorm, beta, S, with_sdf, force_background, lane, st);

        float g[KS];   // dL/dw of the lane's samples
        float grm[3] = {0.f, 0.f, 0.f}, gxm[3] = {0.f, 0.f, 0.f};
        if (d_rgb_map) { grm[0] = __ldg(d_rgb_map + ray * 3); grm[1] = __ldg(d_rgb_map + ray * 3 + 1); grm[2] = __ldg(d_rgb_map + ray * 3 + 2); }
        if (d_xyz_map) { gxm[0] = __ldg(d_xyz_map + ray * 3); gxm[1] = __ldg(d_xyz_map + ray * 3 + 1); gxm[2] = __ldg(d_xyz_map + ray * 3 + 2); }
#pragma unroll
        for (int i = 0; i < KS; i++) {
            const uint32_t s = lane * KS + i;
            g[i] = 0.f;
            if (s < S) {
                const float w = st.w[i];
                const float* c = rgb + (base + s) * 3;
                float gi = 0.f;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float sg = sigmoidf_(__ldg(c + k));
                    gi = fmaf(2.f * grm[k], sg, gi);
                    if (d_rgb) d_rgb[(base + s) * 3 + k] = 2.f * grm[k] * w * sg * (1.f - sg);
                }
                if (d_xyz_map) {
                    const float* p = pts + (base + s) * 3;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        gi = fmaf(gxm[k], __ldg(p + k), gi);
                        if (d_pts) d_pts[(base + s) * 3 + k] = gxm[k] * w;
                    }
                } else if (d_pts) {
                    d_pts[(base + s) * 3] = 0.f; d_pts[(base + s) * 3 + 1] = 0.f; d_pts[(base + s) * 3 + 2] = 0.f;
                }
                if (d_mask && s == S - 1) gi += __ldg(d_mask + ray);
                g[i] = gi;
            }
        }
        if (feat && (d_feat_map || d_feat)) {
            const uint32_t F4 = F >> 2;
            float4 gm[kMaxF4];
#pragma unroll
            for (int j = 0; j < kMaxF4; j++) {
                const uint32_t f4 = lane + 32 * j;
                gm[j] = (d_feat_map && f4 < F4) ? __ldg(reinterpret_cast<const float4*>(d_feat_map + (size_t)ray * F) + f4)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float4* frow = reinterpret_cast<const float4*>(feat + base * F);
            float4* drow = d_feat ? reinterpret_cast<float4*>(d_feat + base * F) : nullptr;
            for (uint32_t s = 0; s < S; s++) {
                float w = 0.f;
#pragma unroll
                for (int i = 0; i < KS; i++) {
                    const float wi = __shfl_sync(0xffffffffu, st.w[i], (int)(s / KS));
                    if ((int)(s % KS) == i) w = wi;
                }
                float dot = 0.f;
#pragma unroll
                for (int j = 0; j < kMaxF4; j++) {
                    const uint32_t f4 = lane + 32 * j;
                    if (f4 < F4) {
                        const float4 v = ldg_stream4(frow + (size_t)s * F4 + f4);
                        dot = fmaf(gm[j].x, v.x, dot); dot = fmaf(gm[j].y, v.y, dot);
                        dot = fmaf(gm[j].z, v.z, dot); dot = fmaf(gm[j].w, v.w, dot);
                        if (drow) drow[(size_t)s * F4 + f4] = make_float4(w * gm[j].x, w * gm[j].y, w * gm[j].z, w * gm[j].w);
                    }
                }
                dot = warp_sum(dot);
#pragma unroll
                for (int i = 0; i < KS; i++)
                    if ((uint32_t)(lane * KS + i) == s) g[i] += dot;
            }
        }
        if (force_background) {
            // w[S-1] was replaced by 1 - sum_{k<S-1} w[k]: its gradient flows (negated) into every earlier weight
            float glast = 0.f;
#pragma unroll
            for (int i = 0; i < KS; i++)
                if ((uint32_t)(lane * KS + i) == S - 1) glast = g[i];
            glast = warp_sum(glast);
#pragma unroll
            for (int i = 0; i < KS; i++) {
                const uint32_t s = lane * KS + i;
                if (s + 1 < S) g[i] -= glast;
                else g[i] = 0.f;
            }
        }
        // Q_s = sum_{k>s} g_k alpha_k prod_{s<j<k} u_j  via a suffix scan of affine maps f_k(x) = g_k alpha_k + u_k x
        float lc = 0.f, lm = 1.f;          // composition of the lane's own maps, first sample outermost
        float cs[KS], ms[KS];
#pragma unroll
        for (int i = KS - 1; i >= 0; i--) {
            const uint32_t s = lane * KS + i;
            const float u = s < S ? (1.f - st.alpha[i] + 1e-10f) : 1.f;
            const float a = s < S ? g[i] * st.alpha[i] : 0.f;
            cs[i] = lc; ms[i] = lm;        // composition of the lane's maps strictly after sample i
            lc = fmaf(u, lc, a);
            lm = u * lm;
        }
        float sc = lc, sm = lm;
        affine_suffix_scan(sc, sm, lane);
        float nc = __shfl_down_sync(0xffffffffu, sc, 1);   // constant term of everything after this lane
        if (lane == 31) nc = 0.f;
#pragma unroll
        for (int i = 0; i < KS; i++) {
            const uint32_t s = lane * KS + i;
            if (s < S) {
                const float Q = fmaf(ms[i], nc, cs[i]);
                const float dalpha = st.T[i] * (g[i] - Q);
                const float dxs = dalpha * (1.f - st.alpha[i]) * st.dist[i];    // d/d(sigma)
                float dsdf;
                if (with_sdf) {
                    const float sg = sigmoidf_(-st.x[i] * inv_beta);
                    const float dsg = sg * (1.f - sg);
                    dsdf = -dxs * dsg * inv_beta * inv_beta;
                    dbeta += dxs * (dsg * st.x[i] * inv_beta * inv_beta * inv_beta - sg * inv_beta * inv_beta);
                } else {
                    dsdf = dxs * sigmoidf_(st.x[i] + (noise ? __ldg(noise + base + s) : 0.f));
                }
                d_sdf[base + s] = dsdf;
            }
        }
    }
    if (d_sigmoid_beta && with_sdf) {
        dbeta = warp_sum(dbeta);
        if (lane == 0) beta_partial[warp] = dbeta;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += beta_partial[k];
            red_add_f32(d_sigmoid_beta, t);
        }
    }
}

}  // namespace sdfg

using namespace sdfg;

static int check_composite(uint32_t S, uint32_t F, const void* feat) {
    SDFG_REQUIRE(S >= 1 && S <= 32 * kMaxPerLane, SDFG_ERR_UNSUPPORTED, "composite: samples per ray must be in 1..%d (got %u)", 32 * kMaxPerLane, S);
    SDFG_REQUIRE(!feat || (F % 4 == 0 && F <= 128 * kMaxF4), SDFG_ERR_UNSUPPORTED, "composite: feature width must be a multiple of 4 and <= %d (got %u)", 128 * kMaxF4, F);
    return SDFG_OK;
}

extern "C" int sdfg_composite_forward(const float* sdf, const float* rgb, const float* feat, const float* z_vals,
                                      const float* rays_d, const float* pts, const float* noise, const float* sigmoid_beta,
                                      uint64_t NR, uint32_t S, uint32_t F, int with_sdf, int force_background, float* rgb_map,
                                      float* feat_map, float* xyz_map, float* mask, float* weights, void* stream) {
    if (int e = check_composite(S, F, feat)) return e;
    if (NR == 0) return SDFG_OK;
    SDFG_REQUIRE(sdf && z_vals && rays_d && (rgb_map || xyz_map || mask || weights), SDFG_ERR_INVALID, "composite_forward: null pointer");
    SDFG_REQUIRE((rgb != nullptr) == (rgb_map != nullptr), SDFG_ERR_INVALID, "composite_forward: rgb and rgb_map go together (both NULL = sdf-only query)");
    SDFG_REQUIRE(!with_sdf || sigmoid_beta, SDFG_ERR_INVALID, "composite_forward: sdf mode needs sigmoid_beta");
    SDFG_REQUIRE(!xyz_map || pts, SDFG_ERR_INVALID, "composite_forward: xyz_map needs pts");
    SDFG_REQUIRE(!feat_map || feat, SDFG_ERR_INVALID, "composite_forward: feat_map needs feat");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)ceil_div<uint64_t>(NR, 8);
#define LAUNCH(KS)                                                                                                          \
    composite_forward_kernel<KS><<<blocks, 256, 0, st>>>(sdf, rgb, feat_map ? feat : nullptr, z_vals, rays_d, pts, noise,    \
                                                         sigmoid_beta, NR, S, F, with_sdf, force_background, rgb_map,        \
                                                         feat_map, xyz_map, mask, weights)
    if (S <= 32) LAUNCH(1);
    else if (S <= 64) LAUNCH(2);
    else if (S <= 128) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    return check_launch("composite_forward_kernel");
}

extern "C" int sdfg_composite_forward_h(const float* sdf, const float* rgb, const uint16_t* feat16, const float* z_vals,
                                        const float* rays_d, const float* pts, const float* noise, const float* sigmoid_beta,
                                        uint64_t NR, uint32_t S, uint32_t F, int with_sdf, int force_background, float* rgb_map,
                                        float* feat_map, float* xyz_map, float* mask, float* weights, void* stream) {
    if (int e = check_composite(S, F, feat16)) return e;
    if (NR == 0) return SDFG_OK;
    SDFG_REQUIRE(sdf && rgb && z_vals && rays_d && rgb_map && feat16 && feat_map, SDFG_ERR_INVALID, "composite_forward_h: null pointer");
    SDFG_REQUIRE(!with_sdf || sigmoid_beta, SDFG_ERR_INVALID, "composite_forward_h: sdf mode needs sigmoid_beta");
    SDFG_REQUIRE(!xyz_map || pts, SDFG_ERR_INVALID, "composite_forward_h: xyz_map needs pts");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)ceil_div<uint64_t>(NR, 8);
    const float* feat = reinterpret_cast<const float*>(feat16);
#define LAUNCH(KS)                                                                                                          \
    composite_forward_kernel<KS, true><<<blocks, 256, 0, st>>>(sdf, rgb, feat, z_vals, rays_d, pts, noise, sigmoid_beta, NR, \
                                                               S, F, with_sdf, force_background, rgb_map, feat_map, xyz_map, \
                                                               mask, weights)
    if (S <= 32) LAUNCH(1);
    else if (S <= 64) LAUNCH(2);
    else if (S <= 128) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    return check_launch("composite_forward_kernel<h>");
}

extern "C" int sdfg_composite_backward(const float* sdf, const float* rgb, const float* feat, const float* z_vals,
                                       const float* rays_d, const float* pts, const float* noise, const float* sigmoid_beta,
                                       uint64_t NR, uint32_t S, uint32_t F, int with_sdf, int force_background,
                                       const float* d_rgb_map, const float* d_feat_map, const float* d_xyz_map,
                                       const float* d_mask, float* d_sdf, float* d_rgb, float* d_feat, float* d_pts,
                                       float* d_sigmoid_beta, void* stream) {
    if (int e = check_composite(S, F, feat)) return e;
    if (NR == 0) return SDFG_OK;
    SDFG_REQUIRE(sdf && rgb && z_vals && rays_d && d_sdf, SDFG_ERR_INVALID, "composite_backward: null pointer");
    SDFG_REQUIRE(!with_sdf || sigmoid_beta, SDFG_ERR_INVALID, "composite_backward: sdf mode needs sigmoid_beta");
    SDFG_REQUIRE(!d_xyz_map || pts, SDFG_ERR_INVALID, "composite_backward: d_xyz_map needs pts");
    SDFG_REQUIRE(!(d_feat_map || d_feat) || feat, SDFG_ERR_INVALID, "composite_backward: feature gradients need feat");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)ceil_div<uint64_t>(NR, 8);
#define LAUNCH(KS)                                                                                                          \
    composite_backward_kernel<KS><<<blocks, 256, 0, st>>>(sdf, rgb, feat, z_vals, rays_d, pts, noise, sigmoid_beta, NR, S, F, \
                                                          with_sdf, force_background, d_rgb_map, d_feat_map, d_xyz_map,      \
                                                          d_mask, d_sdf, d_rgb, d_feat, d_pts, d_sigmoid_beta)
    if (S <= 32) LAUNCH(1);
    else if (S <= 64) LAUNCH(2);
    else if (S <= 128) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    return check_launch("composite_backward_kernel");
}
