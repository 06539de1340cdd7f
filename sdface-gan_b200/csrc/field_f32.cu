// Style-modulated SIREN field, fp32 SIMT path (SDFG_PRECISION_FP32): forward and first-order backward.
//
// Behavioural contract (ref = /root/reference/im2scene/sdf/models/sdf_model.py):
//   LinearLayer.forward      :38-41       std_init * (x W^T + b) + bias_init   (heads and input_linear use 1 / 0)
//   FiLMSiren.forward        :61-69       sin(gamma * (x W^T + b) + beta), gamma/beta per image
//   NGPSIRENGenerator.forward:1566-1592   input_linear -> 3 FiLM -> sdf head; cat(h, SH) -> views FiLM -> rgb head
//   SirenGenerator.forward   :121-139     8 FiLM (first K = 3) -> sdf head; cat(h, dirs) -> views FiLM -> rgb head
// The reference runs every layer as fp32 cuBLAS SGEMM + 3 elementwise launches and materialises cat(h, views) [N,272]
// and raw [N,260]; here every layer is ONE tiled fp32 GEMM with the FiLM/sin epilogue fused, the view feature is a
// second K-segment read per RAY (never expanded over the samples), and the heads are streaming dot products.
// This path exists for <= 1e-3 parity with the fp32 reference; the throughput path is field_tc.cu (tcgen05).
#include <algorithm>

#include "common.cuh"
#include "field.cuh"

namespace sdfg {

constexpr int BM = 128, BN = 128, BK = 8, PAD = 4;

// One GEMM operand: element(i, r) with i the output index (row of C for A, column of C for B) and r the reduction index.
//   rcontig:  address = p + (i / div) * ld + r        (rows of samples x features, reduced over features)
//   !rcontig: address = p + (r / div) * ld + i        (reduced over the leading index)
// `div` broadcasts one stored row over `div` consecutive leading indices (the per-ray view feature over its samples).
struct Operand {
    const float* p;
    int64_t ld;
    uint32_t div;
};

struct Segment {
    Operand a, b;
    uint32_t K;     // reduction length of this segment
};

enum Epi { EPI_LINEAR = 0, EPI_FILM = 1, EPI_STORE = 2, EPI_RED = 3 };

struct GemmParams {
    Segment seg[2];
    int nseg;
    uint32_t M, N;          // C is [M, N]
    float* c;               // output [M, ldc]
    int64_t ldc;
    const float* bias;      // [N]                       (LINEAR, FILM)
    const float* gamma;     // + image * gstride + n     (FILM)
    const float* beta;
    int64_t gstride;
    uint32_t rows_per_image;
    float* pre;             // [M, ldc] or NULL          (FILM: saved pre-activation)
    int accumulate;         // STORE: c += acc
    uint32_t k_split;       // RED: reduction rows handled per blockIdx.z
};

template <bool RC>
__device__ __forceinline__ float4 load_operand(const Operand& o, uint32_t i, uint32_t r, uint32_t I, uint32_t R, bool vec_ok) {
    // returns 4 consecutive elements along the contiguous index (r if RC, else i), zero-filled out of range
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t lead = RC ? i : r, minor = RC ? r : i;
    const uint32_t lead_lim = RC ? I : R, minor_lim = RC ? R : I;
    if (lead >= lead_lim || minor >= minor_lim) return v;
    const float* q = o.p + (int64_t)(lead / o.div) * o.ld + minor;
    if (vec_ok && minor + 3 < minor_lim) return __ldg(reinterpret_cast<const float4*>(q));
    v.x = __ldg(q);
    if (minor + 1 < minor_lim) v.y = __ldg(q + 1);
    if (minor + 2 < minor_lim) v.z = __ldg(q + 2);
    if (minor + 3 < minor_lim) v.w = __ldg(q + 3);
    return v;
}

__host__ __device__ inline bool operand_vec_ok(const Operand& o) {
    return (reinterpret_cast<uintptr_t>(o.p) & 15) == 0 && (o.ld & 3) == 0;
}

template <bool A_RC, bool B_RC, int EPI>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const __grid_constant__ GemmParams P) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const uint32_t m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    for (int s = 0; s < P.nseg; s++) {
        const Segment& sg = P.seg[s];
        uint32_t k_begin = 0, k_end = sg.K;
        if (EPI == EPI_RED) {
            k_begin = blockIdx.z * P.k_split;
            k_end = min(sg.K, k_begin + P.k_split);
            if (k_begin >= k_end) continue;
        }
        const bool a_vec = operand_vec_ok(sg.a), b_vec = operand_vec_ok(sg.b);
        // tile loaders: RC -> thread (i = tid/2, r = (tid%2)*4 .. +3); !RC -> thread (r = tid/32, i = (tid%32)*4 .. +3)
        const uint32_t a_i = A_RC ? tid >> 1 : (tid & 31) * 4, a_r = A_RC ? (tid & 1) * 4 : tid >> 5;
        const uint32_t b_i = B_RC ? tid >> 1 : (tid & 31) * 4, b_r = B_RC ? (tid & 1) * 4 : tid >> 5;
        auto fetch_a = [&](uint32_t k0) { return load_operand<A_RC>(sg.a, m0 + a_i, k0 + a_r, P.M, k_end, a_vec); };
        auto fetch_b = [&](uint32_t k0) { return load_operand<B_RC>(sg.b, n0 + b_i, k0 + b_r, P.N, k_end, b_vec); };
        auto stash = [&](int buf, const float4& va, const float4& vb) {
            if (A_RC) { As[buf][a_r][a_i] = va.x; As[buf][a_r + 1][a_i] = va.y; As[buf][a_r + 2][a_i] = va.z; As[buf][a_r + 3][a_i] = va.w; }
            else *reinterpret_cast<float4*>(&As[buf][a_r][a_i]) = va;
            if (B_RC) { Bs[buf][b_r][b_i] = vb.x; Bs[buf][b_r + 1][b_i] = vb.y; Bs[buf][b_r + 2][b_i] = vb.z; Bs[buf][b_r + 3][b_i] = vb.w; }
            else *reinterpret_cast<float4*>(&Bs[buf][b_r][b_i]) = vb;
        };
        float4 va = fetch_a(k_begin), vb = fetch_b(k_begin);
        __syncthreads();                   // previous segment's readers are done with buffer 0
        stash(0, va, vb);
        __syncthreads();
        int buf = 0;
        for (uint32_t k0 = k_begin; k0 < k_end; k0 += BK) {
            const bool more = k0 + BK < k_end;
            if (more) { va = fetch_a(k0 + BK); vb = fetch_b(k0 + BK); }
#pragma unroll
            for (int k = 0; k < BK; k++) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (more) {
                stash(buf ^ 1, va, vb);
                __syncthreads();
                buf ^= 1;
            }
        }
    }

    // epilogue: thread owns rows {ty*4+i, 64+ty*4+i} x cols {tx*4+j, 64+tx*4+j}
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= P.M) continue;
        const uint32_t img = (EPI == EPI_FILM) ? m / P.rows_per_image : 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t n = n0 + h * 64 + tx * 4;
            if (n >= P.N) continue;
            float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
            float* crow = P.c + (int64_t)m * P.ldc + n;
            const int lim = min(4u, P.N - n);
            if (EPI == EPI_LINEAR || EPI == EPI_FILM) {
                float* prow = (EPI == EPI_FILM && P.pre) ? P.pre + (int64_t)m * P.ldc + n : nullptr;
                for (int j = 0; j < lim; j++) {
                    float t = v[j] + __ldg(P.bias + n + j);
                    if (EPI == EPI_FILM) {
                        if (prow) prow[j] = t;
                        const float g = __ldg(P.gamma + (int64_t)img * P.gstride + n + j);
                        const float bt = __ldg(P.beta + (int64_t)img * P.gstride + n + j);
                        t = sinf(fmaf(g, t, bt));
                    }
                    v[j] = t;
                }
            }
            if (EPI == EPI_RED) {
                for (int j = 0; j < lim; j++) red_add_f32(crow + j, v[j]);
            } else {
                const bool vec = lim == 4 && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
                if (EPI == EPI_STORE && P.accumulate) {
                    for (int j = 0; j < lim; j++) crow[j] += v[j];
                } else if (vec) {
                    *reinterpret_cast<float4*>(crow) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    for (int j = 0; j < lim; j++) crow[j] = v[j];
                }
            }
        }
    }
}

template <bool A_RC, bool B_RC, int EPI>
static int launch_gemm(const GemmParams& P, cudaStream_t st, const char* what) {
    dim3 grid(ceil_div<uint32_t>(P.M, BM), ceil_div<uint32_t>(P.N, BN), 1);
    GemmParams Q = P;
    if (EPI == EPI_RED) {
        // split the (long) reduction so that the grid fills the machine about twice
        const uint32_t tiles = grid.x * grid.y;
        const uint32_t K = P.seg[0].K;
        uint32_t splits = max(1u, min(ceil_div<uint32_t>(2 * sm_count(), tiles), ceil_div<uint32_t>(K, 1024)));
        uint32_t per = ceil_div<uint32_t>(ceil_div<uint32_t>(K, splits), BK) * BK;
        Q.k_split = per;
        grid.z = ceil_div<uint32_t>(K, per);
    }
    ProfScope prof(what, st);
    gemm_f32_kernel<A_RC, B_RC, EPI><<<grid, 256, 0, st>>>(Q);
    return check_launch(what);
}

// ---------------------------------------------------------------------------------------------------------------
// heads: out[m, c] = h[m, :] . w[c, :] + b[c], c < NOUT (1 = sdf, 3 = rgb); one warp per row
template <int NOUT>
__global__ void __launch_bounds__(256) head_forward_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                                           const float* __restrict__ b, float* __restrict__ out, uint64_t M,
                                                           uint32_t W) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t m = warp0; m < M; m += nwarps) {
        float acc[NOUT];
#pragma unroll
        for (int c = 0; c < NOUT; c++) acc[c] = 0.f;
        for (uint32_t k = lane * 4; k < W; k += 128) {
            const float4 x = ldg_stream4(reinterpret_cast<const float4*>(h + m * W + k));
#pragma unroll
            for (int c = 0; c < NOUT; c++) {
                const float4 ww = __ldg(reinterpret_cast<const float4*>(w + (size_t)c * W + k));
                acc[c] = fmaf(x.x, ww.x, acc[c]); acc[c] = fmaf(x.y, ww.y, acc[c]);
                acc[c] = fmaf(x.z, ww.z, acc[c]); acc[c] = fmaf(x.w, ww.w, acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < NOUT; c++) acc[c] = warp_sum(acc[c]);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < NOUT; c++) out[m * NOUT + c] = acc[c] + __ldg(b + c);
        }
    }
}

// heads backward: dh[m, k] (=|+=) sum_c dout[m,c] w[c,k] (+ extra[m,k]);  dw[c,k] += sum_m dout[m,c] h[m,k];  db[c] += sum_m dout[m,c]
// block = W threads (one per column k), 128 rows per block
template <int NOUT>
__global__ void head_backward_kernel(const float* __restrict__ dout, const float* __restrict__ h, const float* __restrict__ w,
                                     const float* __restrict__ extra, float* __restrict__ dh, int accumulate,
                                     float* __restrict__ dw, float* __restrict__ db, uint64_t M, uint32_t W, uint32_t rows_per_block) {
    const uint32_t k = threadIdx.x;
    const uint64_t m_begin = (uint64_t)blockIdx.x * rows_per_block;
    const uint64_t m_end = min(M, m_begin + rows_per_block);
    float wk[NOUT], gw[NOUT], gb[NOUT];
#pragma unroll
    for (int c = 0; c < NOUT; c++) { wk[c] = __ldg(w + (size_t)c * W + k); gw[c] = 0.f; gb[c] = 0.f; }
    for (uint64_t m = m_begin; m < m_end; m++) {
        float d[NOUT];
        float v = extra ? ldg_stream1(extra + m * W + k) : 0.f;
#pragma unroll
        for (int c = 0; c < NOUT; c++) { d[c] = __ldg(dout + m * NOUT + c); v = fmaf(d[c], wk[c], v); }
        if (accumulate) dh[m * W + k] += v;
        else dh[m * W + k] = v;
        if (dw) {
            const float hv = ldg_stream1(h + m * W + k);
#pragma unroll
            for (int c = 0; c < NOUT; c++) { gw[c] = fmaf(d[c], hv, gw[c]); gb[c] += d[c]; }
        }
    }
    if (dw) {
#pragma unroll
        for (int c = 0; c < NOUT; c++) red_add_f32(dw + (size_t)c * W + k, gw[c]);
        if (db && k == 0) {
#pragma unroll
            for (int c = 0; c < NOUT; c++) red_add_f32(db + c, gb[c]);
        }
    }
}

// FiLM backward (elementwise + column reductions), in place on g:
//   z = gamma*pre + beta; dz = g*cos(z); g <- dpre = dz*gamma; dbeta[img,k] += sum dz; dgamma[img,k] += sum dz*pre; dbias[k] += sum dpre
// film == 0: plain linear layer -- g untouched, dbias[k] += sum g.
// block = W threads (one per column), rows_per_block rows which never straddle an image when rows_per_image % rows_per_block == 0
__global__ void film_backward_kernel(float* __restrict__ g, const float* __restrict__ pre, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, int64_t gstride, uint32_t rows_per_image, int film,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, uint64_t M,
                                     uint32_t W, uint32_t rows_per_block) {
    const uint32_t k = threadIdx.x;
    const uint64_t m_begin = (uint64_t)blockIdx.x * rows_per_block;
    const uint64_t m_end = min(M, m_begin + rows_per_block);
    float sb = 0.f, sg = 0.f, sbias = 0.f;
    uint64_t img = m_begin / rows_per_image;
    float gm = film ? __ldg(gamma + img * gstride + k) : 1.f, bt = film ? __ldg(beta + img * gstride + k) : 0.f;
    for (uint64_t m = m_begin; m < m_end; m++) {
        if (film) {
            const uint64_t im = m / rows_per_image;
            if (im != img) {       // flush the finished image's partial sums
                if (dbeta) red_add_f32(dbeta + img * gstride + k, sb);
                if (dgamma) red_add_f32(dgamma + img * gstride + k, sg);
                sb = 0.f; sg = 0.f;
                img = im;
                gm = __ldg(gamma + img * gstride + k);
                bt = __ldg(beta + img * gstride + k);
            }
            const float p = ldg_stream1(pre + m * W + k);
            const float dz = g[m * W + k] * cosf(fmaf(gm, p, bt));
            sb += dz;
            sg = fmaf(dz, p, sg);
            const float dp = dz * gm;
            g[m * W + k] = dp;
            sbias += dp;
        } else {
            sbias += g[m * W + k];
        }
    }
    if (film) {
        if (dbeta) red_add_f32(dbeta + img * gstride + k, sb);
        if (dgamma) red_add_f32(dgamma + img * gstride + k, sg);
    }
    if (dbias) red_add_f32(dbias + k, sbias);
}

// ---------------------------------------------------------------------------------------------------------------
// host orchestration

struct Slots {          // [N, W] fp32 activation slots inside the caller's workspace
    float* base;
    uint64_t stride;    // N * W
    float* at(int i) const { return base + (uint64_t)i * stride; }
};

// slot map when save_for_backward: 0 = h0 (input_linear output, if any); 1+2i = pre_i; 2+2i = h_i (i < n_film);
// 1+2n = pre_views; 2+2n = h_views.  Without saving: two ping-pong slots + one for h_views.
static int n_slots(const sdfg_field_params* p, int save) { return save ? 3 + 2 * (int)p->n_film : 3; }

static int check_params(const sdfg_field_params* p, uint64_t N) {
    SDFG_REQUIRE(p, SDFG_ERR_INVALID, "field: null params");
    SDFG_REQUIRE(p->width >= 4 && p->width % 4 == 0 && p->width <= 1024, SDFG_ERR_UNSUPPORTED, "field: width must be a multiple of 4 in 4..1024 (got %u)", p->width);
    SDFG_REQUIRE(p->n_film >= 1 && p->n_film + 1 <= SDFG_MAX_FILM, SDFG_ERR_UNSUPPORTED, "field: n_film must be in 1..%d (got %u)", SDFG_MAX_FILM - 1, p->n_film);
    SDFG_REQUIRE(p->in_dim >= 1 && p->view_dim >= 1, SDFG_ERR_INVALID, "field: in_dim / view_dim must be positive");
    SDFG_REQUIRE(p->samples_per_ray >= 1 && p->samples_per_image >= 1, SDFG_ERR_INVALID, "field: samples_per_ray / samples_per_image must be positive");
    SDFG_REQUIRE(N % p->samples_per_ray == 0, SDFG_ERR_INVALID, "field: N (%llu) is not a multiple of samples_per_ray (%u)", (unsigned long long)N, p->samples_per_ray);
    SDFG_REQUIRE(N < (1ull << 32), SDFG_ERR_UNSUPPORTED, "field: at most 2^32-1 samples per call");
    SDFG_REQUIRE(!p->has_input_linear || (p->input_w && p->input_b), SDFG_ERR_INVALID, "field: input_linear weights missing");
    for (uint32_t l = 0; l <= p->n_film; l++)
        SDFG_REQUIRE(p->film_w[l] && p->film_b[l], SDFG_ERR_INVALID, "field: FiLM layer %u weights missing", l);
    SDFG_REQUIRE(p->gamma && p->beta && p->sigma_w && p->sigma_b, SDFG_ERR_INVALID, "field: gamma/beta/sigma head missing");
    return SDFG_OK;
}

static int film_gemm(const sdfg_field_params* p, uint32_t layer, const float* a, uint32_t K, const float* view_feat,
                     float* out, float* pre, uint64_t N, cudaStream_t st) {
    const uint32_t W = p->width;
    const bool views = layer == p->n_film;
    const int64_t ldw = views ? (int64_t)W + p->view_dim : (int64_t)K;
    GemmParams g = {};
    g.seg[0] = {{a, (int64_t)K, 1}, {p->film_w[layer], ldw, 1}, K};
    g.nseg = 1;
    if (views) {
        g.seg[1] = {{view_feat, (int64_t)p->view_dim, p->samples_per_ray}, {p->film_w[layer] + W, ldw, 1}, p->view_dim};
        g.nseg = 2;
    }
    g.M = (uint32_t)N; g.N = W; g.c = out; g.ldc = W; g.bias = p->film_b[layer];
    g.gamma = p->gamma + (size_t)layer * W; g.beta = p->beta + (size_t)layer * W;
    g.gstride = (int64_t)(p->n_film + 1) * W;
    g.rows_per_image = p->samples_per_image;
    g.pre = pre;
    return launch_gemm<true, true, EPI_FILM>(g, st, "gemm_f32_kernel<film>");
}

int field_forward_f32(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N, float* out_sdf,
                           float* out_rgb, float* out_feat, void* workspace, int save, cudaStream_t st) {
    const uint32_t W = p->width;
    Slots S{(float*)workspace, N * W};
    const int nf = (int)p->n_film;
    const float* cur = x_in;
    uint32_t K = p->in_dim;
    const unsigned head_blocks = (unsigned)std::min<uint64_t>(ceil_div<uint64_t>(N, 8), 1u << 20);
    if (p->has_input_linear) {
        GemmParams g = {};
        g.seg[0] = {{x_in, (int64_t)p->in_dim, 1}, {p->input_w, (int64_t)p->in_dim, 1}, p->in_dim};
        g.nseg = 1; g.M = (uint32_t)N; g.N = W; g.c = S.at(0); g.ldc = W; g.bias = p->input_b;
        if (int e = launch_gemm<true, true, EPI_LINEAR>(g, st, "gemm_f32_kernel<linear>")) return e;
        cur = S.at(0);
        K = W;
    }
    for (int i = 0; i < nf; i++) {
        // saving: pre_i -> slot 1+2i, h_i -> slot 2+2i; otherwise ping-pong between slots 1 and 0 (slot 0 holds h0 first)
        float* out = save ? S.at(2 + 2 * i) : S.at(cur == S.at(1) ? 0 : 1);
        float* pre = save ? S.at(1 + 2 * i) : nullptr;
        if (int e = film_gemm(p, i, cur, K, nullptr, out, pre, N, st)) return e;
        cur = out;
        K = W;
    }
    if (out_sdf) {
        head_forward_kernel<1><<<head_blocks, 256, 0, st>>>(cur, p->sigma_w, p->sigma_b, out_sdf, N, W);
        if (int e = check_launch("head_forward_kernel<1>")) return e;
    }
    if (out_rgb || out_feat) {
        SDFG_REQUIRE(view_feat, SDFG_ERR_INVALID, "field_forward: view_feat is required for the rgb / feature outputs");
        SDFG_REQUIRE(!out_rgb || (p->rgb_w && p->rgb_b), SDFG_ERR_INVALID, "field_forward: rgb head missing");
        float* hv = out_feat ? out_feat : S.at(save ? 2 + 2 * nf : 2);
        float* pre = save ? S.at(1 + 2 * nf) : nullptr;
        if (int e = film_gemm(p, nf, cur, W, view_feat, hv, pre, N, st)) return e;
        if (out_rgb) {
            head_forward_kernel<3><<<head_blocks, 256, 0, st>>>(hv, p->rgb_w, p->rgb_b, out_rgb, N, W);
            if (int e = check_launch("head_forward_kernel<3>")) return e;
        }
    }
    return SDFG_OK;
}

int field_backward_f32(const sdfg_field_params* p, const sdfg_field_grads* g, const float* x_in, const float* view_feat,
                            uint64_t N, const float* d_sdf, const float* d_rgb, const float* d_feat, const float* out_feat,
                            const void* workspace, void* scratch, float* d_x_in, cudaStream_t st) {
    const uint32_t W = p->width;
    SDFG_REQUIRE(W <= 1024, SDFG_ERR_UNSUPPORTED, "field_backward: width > 1024");
    Slots S{(float*)workspace, N * W};
    float* G[2] = {(float*)scratch, (float*)scratch + N * W};
    const int nf = (int)p->n_film;
    const uint32_t RPB = 256;     // rows per block of the column-reduction kernels
    const unsigned rblocks = (unsigned)ceil_div<uint64_t>(N, RPB);
    const int64_t gstride = (int64_t)(nf + 1) * W;
    auto h_of = [&](int i) -> const float* { return i < 0 ? (p->has_input_linear ? S.at(0) : x_in) : S.at(2 + 2 * i); };
    const float* h_last = h_of(nf - 1);
    int cur = 0;
    bool have = false;            // G[cur] holds d(h_last)
    if (d_rgb || d_feat) {
        SDFG_REQUIRE(view_feat, SDFG_ERR_INVALID, "field_backward: view_feat is required");
        const float* hv = out_feat ? out_feat : S.at(2 + 2 * nf);
        // d(h_views) = d_rgb W_rgb + d_feat
        if (d_rgb) {
            head_backward_kernel<3><<<rblocks, W, 0, st>>>(d_rgb, hv, p->rgb_w, d_feat, G[0], 0, g ? g->rgb_w : nullptr,
                                                           g ? g->rgb_b : nullptr, N, W, RPB);
            if (int e = check_launch("head_backward_kernel<3>")) return e;
        } else {
            if (cudaMemcpyAsync(G[0], d_feat, N * W * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                return set_error(SDFG_ERR_CUDA, "field_backward: copy of d_feat failed");
        }
        film_backward_kernel<<<rblocks, W, 0, st>>>(G[0], S.at(1 + 2 * nf), p->gamma + (size_t)nf * W, p->beta + (size_t)nf * W, gstride,
                                                    p->samples_per_image, 1, g ? g->gamma + (size_t)nf * W : nullptr,
                                                    g ? g->beta + (size_t)nf * W : nullptr, g ? g->film_b[nf] : nullptr, N, W, RPB);
        if (int e = check_launch("film_backward_kernel")) return e;
        const int64_t ldw = (int64_t)W + p->view_dim;
        if (g && g->film_w[nf]) {
            GemmParams q = {};
            q.seg[0] = {{G[0], (int64_t)W, 1}, {h_last, (int64_t)W, 1}, (uint32_t)N};
            q.nseg = 1; q.M = W; q.N = W; q.c = g->film_w[nf]; q.ldc = ldw;
            if (int e = launch_gemm<false, false, EPI_RED>(q, st, "gemm_f32_kernel<wgrad>")) return e;
            q.seg[0] = {{G[0], (int64_t)W, 1}, {view_feat, (int64_t)p->view_dim, p->samples_per_ray}, (uint32_t)N};
            q.N = p->view_dim; q.c = g->film_w[nf] + W;
            if (int e = launch_gemm<false, false, EPI_RED>(q, st, "gemm_f32_kernel<wgrad-view>")) return e;
        }
        GemmParams q = {};
        q.seg[0] = {{G[0], (int64_t)W, 1}, {p->film_w[nf], ldw, 1}, W};
        q.nseg = 1; q.M = (uint32_t)N; q.N = W; q.c = G[1]; q.ldc = W;
        if (int e = launch_gemm<true, false, EPI_STORE>(q, st, "gemm_f32_kernel<dgrad>")) return e;
        cur = 1;
        have = true;
    }
    if (d_sdf) {
        head_backward_kernel<1><<<rblocks, W, 0, st>>>(d_sdf, h_last, p->sigma_w, nullptr, G[cur], have ? 1 : 0,
                                                       g ? g->sigma_w : nullptr, g ? g->sigma_b : nullptr, N, W, RPB);
        if (int e = check_launch("head_backward_kernel<1>")) return e;
        have = true;
    }
    SDFG_REQUIRE(have, SDFG_ERR_INVALID, "field_backward: no output gradient given");
    for (int i = nf - 1; i >= 0; i--) {
        film_backward_kernel<<<rblocks, W, 0, st>>>(G[cur], S.at(1 + 2 * i), p->gamma + (size_t)i * W, p->beta + (size_t)i * W, gstride,
                                                    p->samples_per_image, 1, g ? g->gamma + (size_t)i * W : nullptr,
                                                    g ? g->beta + (size_t)i * W : nullptr, g ? g->film_b[i] : nullptr, N, W, RPB);
        if (int e = check_launch("film_backward_kernel")) return e;
        const float* hin = h_of(i - 1);
        const uint32_t K = (i == 0 && !p->has_input_linear) ? p->in_dim : W;
        if (g && g->film_w[i]) {
            GemmParams q = {};
            q.seg[0] = {{G[cur], (int64_t)W, 1}, {hin, (int64_t)K, 1}, (uint32_t)N};
            q.nseg = 1; q.M = W; q.N = K; q.c = g->film_w[i]; q.ldc = K;
            if (int e = launch_gemm<false, false, EPI_RED>(q, st, "gemm_f32_kernel<wgrad>")) return e;
        }
        const bool need_dx = i > 0 || p->has_input_linear || d_x_in;
        if (need_dx) {
            float* dst = (i == 0 && !p->has_input_linear) ? d_x_in : G[cur ^ 1];
            GemmParams q = {};
            q.seg[0] = {{G[cur], (int64_t)W, 1}, {p->film_w[i], (int64_t)K, 1}, W};
            q.nseg = 1; q.M = (uint32_t)N; q.N = K; q.c = dst; q.ldc = K;
            if (int e = launch_gemm<true, false, EPI_STORE>(q, st, "gemm_f32_kernel<dgrad>")) return e;
            cur ^= 1;
        }
    }
    if (p->has_input_linear) {
        if (g && g->input_b) {
            film_backward_kernel<<<rblocks, W, 0, st>>>(G[cur], nullptr, nullptr, nullptr, 0, p->samples_per_image, 0, nullptr, nullptr,
                                                        g->input_b, N, W, RPB);
            if (int e = check_launch("film_backward_kernel<bias>")) return e;
        }
        if (g && g->input_w) {
            GemmParams q = {};
            q.seg[0] = {{G[cur], (int64_t)W, 1}, {x_in, (int64_t)p->in_dim, 1}, (uint32_t)N};
            q.nseg = 1; q.M = W; q.N = p->in_dim; q.c = g->input_w; q.ldc = p->in_dim;
            if (int e = launch_gemm<false, false, EPI_RED>(q, st, "gemm_f32_kernel<wgrad-in>")) return e;
        }
        if (d_x_in) {
            GemmParams q = {};
            q.seg[0] = {{G[cur], (int64_t)W, 1}, {p->input_w, (int64_t)p->in_dim, 1}, W};
            q.nseg = 1; q.M = (uint32_t)N; q.N = p->in_dim; q.c = d_x_in; q.ldc = p->in_dim;
            if (int e = launch_gemm<true, false, EPI_STORE>(q, st, "gemm_f32_kernel<dgrad-in>")) return e;
        }
    }
    return SDFG_OK;
}

uint64_t field_workspace_bytes_f32(const sdfg_field_params* p, uint64_t N, int save) {
    return (uint64_t)n_slots(p, save) * N * p->width * sizeof(float);
}

int field_check_params(const sdfg_field_params* p, uint64_t N) { return check_params(p, N); }

}  // namespace sdfg
