// Style-modulated SIREN field on the tcgen05 tensor cores (SDFG_PRECISION_TC16): host orchestration + small prep kernels.
//
// Same behavioural contract as field_f32.cu (ref sdf_model.py:38-41, :61-69, :121-139, :1566-1592).  Activations and weights are
// fp16 (values are bounded: |sin| <= 1, SIREN weights << 1; fp16's 11-bit mantissa keeps the gamma ~ 30 amplification of SIREN
// inside the north star's 2e-2 relative band, which bf16 measured at 3-4.5% does not), gradients are bf16 (range), accumulation
// is fp32 in TMEM.
// Data layout in HBM (all inside the caller's workspace):
//   Wb_l    fp16 [W, Kp_l]      weights of every layer, re-cast from the fp32 masters each call (K padded to a multiple of 8)
//   X0      fp16 [N, Kp_in]     encoder features
//   A_l     fp16 [N, Kp_l]      input of FiLM layer l; the last trunk output is written with pitch Kp_views and the per-ray view
//                               feature is expanded into its tail columns, so the views layer is one K = W + V contraction
//   HV      fp16 [N, W]         output of the views layer (kept for the rgb-head weight gradient)
// Algorithmic HBM bytes per sample and layer: 2*K in + 2*W out (fp16) -- 1 KB for a 256x256 layer, against 2 * 131072 flop.
#include <algorithm>

#include "field.cuh"
#include "tc_layer.cuh"

namespace sdfg {

using tc::LayerParams;

__host__ __device__ inline uint32_t round_up(uint32_t a, uint32_t b) { return (a + b - 1) / b * b; }

typedef uint16_t h16;   // storage of a 16-bit float (fp16 for activations / weights, bf16 for gradients)

// fp32 [rows, cols] (pitch ld_in) -> 16-bit [rows, cols_p] (pitch ld_out), zero padded; `div` broadcasts input rows (view feature per ray)
__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ in, int64_t ld_in, uint32_t div, h16* __restrict__ out,
                                                        int64_t ld_out, uint64_t rows, uint32_t cols, uint32_t cols_p, uint32_t fmt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols_p) return;
    const uint64_t r = i / cols_p;
    const uint32_t c = (uint32_t)(i % cols_p);
    const float v = c < cols ? __ldg(in + (r / div) * ld_in + c) : 0.f;
    out[r * ld_out + c] = fmt == tc::FMT_BF16 ? __bfloat16_as_ushort(__float2bfloat16(v)) : __half_as_ushort(__float2half_rn(v));
}

static int cast_pad(const float* in, int64_t ld_in, uint32_t div, h16* out, int64_t ld_out, uint64_t rows, uint32_t cols,
                    uint32_t cols_p, cudaStream_t st, uint32_t fmt = tc::FMT_F16) {
    if (rows == 0) return SDFG_OK;
    const uint64_t total = rows * cols_p;
    cast_pad_kernel<<<(unsigned)ceil_div<uint64_t>(total, 256), 256, 0, st>>>(in, ld_in, div, out, ld_out, rows, cols, cols_p, fmt);
    return check_launch("cast_pad_kernel");
}

template <int MODE>
static int launch_layer(const void* a, uint64_t a_rows, uint32_t K, int64_t lda, const void* b, uint64_t b_rows, int64_t ldb, LayerParams P,
                        cudaStream_t st, const char* what) {
    SDFG_REQUIRE(P.N_out % 32 == 0 && P.N_out >= 32 && P.N_out <= 256, SDFG_ERR_UNSUPPORTED, "tc layer: N_out must be a multiple of 32 in 32..256 (got %u)", P.N_out);
    SDFG_REQUIRE(K % 8 == 0 && K >= 8 && K <= tc::MAX_KCH * tc::KCH, SDFG_ERR_UNSUPPORTED, "tc layer: K must be a multiple of 8 in 8..%u (got %u)", tc::MAX_KCH * tc::KCH, K);
    CUtensorMap tmA, tmB;
    if (int e = make_tensor_map_16(&tmA, a, a_rows, K, (uint64_t)lda, tc::TILE_M, tc::KCH, P.ab_fmt)) return e;
    if (int e = make_tensor_map_16(&tmB, b, b_rows, K, (uint64_t)ldb, P.N_out, tc::KCH, P.ab_fmt)) return e;
    P.K = K;
    P.n_tiles = (uint32_t)ceil_div<uint64_t>(P.M_total, tc::TILE_M);
    const uint32_t ctas = std::min<uint32_t>((uint32_t)sm_count(), P.n_tiles);
    P.tiles_per_cta = ceil_div<uint32_t>(P.n_tiles, ctas);
    const uint32_t grid = ceil_div<uint32_t>(P.n_tiles, P.tiles_per_cta);
    const uint32_t smem = tc::layer_smem_bytes(K, P.N_out);
    static thread_local uint32_t configured[3] = {0, 0, 0};
    if (configured[MODE] < smem) {
        if (cudaFuncSetAttribute(tc::tc_layer_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return set_error(SDFG_ERR_CUDA, "%s: cannot opt in to %u bytes of shared memory", what, smem);
        configured[MODE] = smem;
    }
    ProfScope prof(what, st);
    tc::tc_layer_kernel<MODE><<<grid, tc::LAYER_THREADS, smem, st>>>(tmA, tmB, P);
    return check_launch(what);
}

// ---------------------------------------------------------------------------------------------------------------
// workspace carving

struct TcLayout {
    uint32_t W, Kp_in, Kp_v, n_film, n_layers;   // n_layers = FiLM layers incl. views
    uint64_t N;
    uint64_t off_w[SDFG_MAX_FILM + 1];           // fp16 weights: [0] = input_linear, [1 + l] = FiLM layer l
    uint64_t off_x0, off_a[SDFG_MAX_FILM + 1], off_hv, total;
    int save;
};

static uint64_t align256(uint64_t x) { return (x + 255) & ~uint64_t(255); }

static TcLayout tc_layout(const sdfg_field_params* p, uint64_t N, int save) {
    TcLayout L = {};
    L.W = p->width; L.N = N; L.n_film = p->n_film; L.n_layers = p->n_film + 1; L.save = save;
    L.Kp_in = round_up(p->in_dim, 8);
    L.Kp_v = round_up(p->width + p->view_dim, 8);
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) { const uint64_t o = off; off = align256(off + bytes); return o; };
    L.off_w[0] = take(p->has_input_linear ? (uint64_t)L.W * L.Kp_in * 2 : 0);
    for (uint32_t l = 0; l < L.n_layers; l++) {
        const uint32_t K = l == L.n_film ? L.Kp_v : ((l == 0 && !p->has_input_linear) ? L.Kp_in : L.W);
        L.off_w[1 + l] = take((uint64_t)L.W * K * 2);
    }
    L.off_x0 = take(N * L.Kp_in * 2);
    // A_l: input of FiLM layer l.  Without input_linear A_0 is X0 itself.  Not saving: two ping-pong trunk buffers.
    for (uint32_t l = 0; l < L.n_layers; l++) {
        if (l == 0 && !p->has_input_linear) { L.off_a[0] = L.off_x0; continue; }
        if (l == L.n_film) { L.off_a[l] = take(N * L.Kp_v * 2); continue; }
        if (!save && l >= 2 + (p->has_input_linear ? 0u : 1u)) { L.off_a[l] = L.off_a[l - 2]; continue; }
        L.off_a[l] = take(N * L.W * 2);
    }
    L.off_hv = take(save ? N * L.W * 2 : 0);
    L.total = off;
    return L;
}

uint64_t field_workspace_bytes_tc(const sdfg_field_params* p, uint64_t N, int save) { return tc_layout(p, N, save).total; }

static int check_tc(const sdfg_field_params* p, uint64_t N) {
    SDFG_REQUIRE(p->width == 256, SDFG_ERR_UNSUPPORTED, "tc field: width must be 256 (got %u)", p->width);
    SDFG_REQUIRE(p->samples_per_image % tc::TILE_M == 0, SDFG_ERR_UNSUPPORTED,
                 "tc field: samples_per_image (%u) must be a multiple of %u so a tile never straddles two images", p->samples_per_image, tc::TILE_M);
    SDFG_REQUIRE(round_up(p->width + p->view_dim, 8) <= tc::MAX_KCH * tc::KCH && round_up(p->in_dim, 8) <= tc::MAX_KCH * tc::KCH, SDFG_ERR_UNSUPPORTED,
                 "tc field: in_dim / view_dim too large");
    (void)N;
    return SDFG_OK;
}

int field_forward_tc(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N, float* out_sdf, float* out_rgb,
                     float* out_feat, void* workspace, int save, cudaStream_t st) {
    if (int e = check_tc(p, N)) return e;
    const TcLayout L = tc_layout(p, N, save);
    uint8_t* ws = (uint8_t*)workspace;
    auto Wb = [&](uint32_t i) { return (h16*)(ws + L.off_w[i]); };
    auto A = [&](uint32_t l) { return (h16*)(ws + L.off_a[l]); };
    const uint32_t W = L.W, nf = L.n_film;
    const int64_t gstride = (int64_t)(nf + 1) * W;
    const bool want_views = out_rgb || out_feat;
    // 1. weights -> fp16 (padded K)
    if (p->has_input_linear)
        if (int e = cast_pad(p->input_w, p->in_dim, 1, Wb(0), L.Kp_in, W, p->in_dim, L.Kp_in, st)) return e;
    for (uint32_t l = 0; l <= nf; l++) {
        if (l == nf && !want_views) break;
        const uint32_t K = l == nf ? W + p->view_dim : ((l == 0 && !p->has_input_linear) ? p->in_dim : W);
        if (int e = cast_pad(p->film_w[l], K, 1, Wb(1 + l), round_up(K, 8), W, K, round_up(K, 8), st)) return e;
    }
    // 2. encoder features -> fp16
    h16* X0 = (h16*)(ws + L.off_x0);
    if (int e = cast_pad(x_in, p->in_dim, 1, X0, L.Kp_in, N, p->in_dim, L.Kp_in, st)) return e;
    // 3. input_linear
    if (p->has_input_linear) {
        LayerParams P = {};
        P.M_total = (uint32_t)N; P.N_out = W; P.rows_per_image = p->samples_per_image; P.act = 0; P.bias = p->input_b;
        P.out16 = A(0); P.ld_out = W;
        if (int e = launch_layer<tc::MODE_F>(X0, N, L.Kp_in, L.Kp_in, Wb(0), W, L.Kp_in, P, st, "tc_layer_kernel<F,gemm,linear>")) return e;
    }
    // 4. trunk
    for (uint32_t l = 0; l < nf; l++) {
        const uint32_t K = (l == 0 && !p->has_input_linear) ? L.Kp_in : W;
        LayerParams P = {};
        P.M_total = (uint32_t)N; P.N_out = W; P.rows_per_image = p->samples_per_image; P.act = 1; P.bias = p->film_b[l];
        P.gamma = p->gamma + (size_t)l * W; P.beta = p->beta + (size_t)l * W; P.gstride = gstride;
        const bool last = l + 1 == nf;
        if (!last || want_views || save) { P.out16 = A(l + 1); P.ld_out = last ? L.Kp_v : W; }
        if (last && out_sdf) { P.nh = 1; P.head_w = p->sigma_w; P.head_b = p->sigma_b; P.out_head = out_sdf; }
        if (int e = launch_layer<tc::MODE_F>(A(l), N, K, K, Wb(1 + l), W, K, P, st, "tc_layer_kernel<F,gemm,film>")) return e;
    }
    if (!want_views) return SDFG_OK;
    SDFG_REQUIRE(view_feat, SDFG_ERR_INVALID, "field_forward: view_feat is required for the rgb / feature outputs");
    SDFG_REQUIRE(!out_rgb || (p->rgb_w && p->rgb_b), SDFG_ERR_INVALID, "field_forward: rgb head missing");
    // 5. per-ray view feature -> tail columns of the views input
    if (int e = cast_pad(view_feat, p->view_dim, p->samples_per_ray, A(nf) + W, L.Kp_v, N, p->view_dim, L.Kp_v - W, st)) return e;
    // 6. views layer + rgb head
    {
        LayerParams P = {};
        P.M_total = (uint32_t)N; P.N_out = W; P.rows_per_image = p->samples_per_image; P.act = 1; P.bias = p->film_b[nf];
        P.gamma = p->gamma + (size_t)nf * W; P.beta = p->beta + (size_t)nf * W; P.gstride = gstride;
        if (save) { P.out16 = (h16*)(ws + L.off_hv); P.ld_out = W; }
        if (out_feat) { P.out_f32 = out_feat; P.ld_out_f32 = W; }
        if (out_rgb) { P.nh = 3; P.head_w = p->rgb_w; P.head_b = p->rgb_b; P.out_head = out_rgb; }
        if (int e = launch_layer<tc::MODE_F>(A(nf), N, L.Kp_v, L.Kp_v, Wb(1 + nf), W, L.Kp_v, P, st, "tc_layer_kernel<F,gemm,views>")) return e;
    }
    return SDFG_OK;
}

// probe for the parity tests: out[M,N] = f16(x)[M,K] * f16(w)[N,K]^T through the MODE_F pipeline (linear epilogue, zero bias)
int tc_linear_probe(const float* x, const float* w, float* out, uint32_t M, uint32_t K, uint32_t N, void* workspace, cudaStream_t st) {
    const uint32_t Kp = round_up(K, 8);
    h16* xb = (h16*)workspace;
    h16* wb = xb + align256((uint64_t)M * Kp * 2) / 2;
    float* zero = (float*)(wb + align256((uint64_t)N * Kp * 2) / 2);
    if (cudaMemsetAsync(zero, 0, N * sizeof(float), st) != cudaSuccess) return set_error(SDFG_ERR_CUDA, "tc_linear_probe: memset failed");
    if (int e = cast_pad(x, K, 1, xb, Kp, M, K, Kp, st)) return e;
    if (int e = cast_pad(w, K, 1, wb, Kp, N, K, Kp, st)) return e;
    LayerParams P = {};
    P.M_total = M; P.N_out = N; P.rows_per_image = std::max(M, 1u); P.act = 0; P.bias = zero; P.out_f32 = out; P.ld_out_f32 = N;
    return launch_layer<tc::MODE_F>(xb, M, Kp, Kp, wb, N, Kp, P, st, "tc_layer_kernel<F,gemm,probe>");
}

}  // namespace sdfg

extern "C" uint64_t sdfg_tc_linear_probe_workspace_bytes(uint32_t M, uint32_t K, uint32_t N) {
    const uint64_t Kp = sdfg::round_up(K, 8);
    return sdfg::align256((uint64_t)M * Kp * 2) + sdfg::align256((uint64_t)N * Kp * 2) + sdfg::align256((uint64_t)N * 4) + 256;
}

extern "C" int sdfg_tc_linear_probe(const float* x, const float* w, float* out, uint32_t M, uint32_t K, uint32_t N, void* workspace,
                                    void* stream) {
    using namespace sdfg;
    if (M == 0) return SDFG_OK;
    SDFG_REQUIRE(x && w && out && workspace, SDFG_ERR_INVALID, "tc_linear_probe: null pointer");
    return tc_linear_probe(x, w, out, M, K, N, workspace, (cudaStream_t)stream);
}
