// StyleGAN2 decoder (ref sdf_model.py:614-1056 + sdf_op.py) -- forward kernels behind the C ABI (include/sdfg.h):
//   sdfg_nhwc16                 fp32 [B, H*W, C] (the renderer's feature map, channels last) -> fp16 [B, H, W, C]
//   sdfg_modconv_fold           per-sample weights of a ModulatedConv2d: scale * W * style, demodulated           (:655-669)
//   sdfg_conv_forward           3 x 3 convolution / plain GEMM on tcgen05 (tc_conv.cuh) + noise + bias + leaky ReLU (:790-818)
//   sdfg_upconv_gather          transposed-convolution taps -> blur -> noise + bias + leaky ReLU                 (:671-684, Blur :522-538)
//   sdfg_to_rgb                 1 x 1 modulated convolution to 3 channels + bias + up-sampled skip               (:821-843, Upsample :480-499)
// Activations are channels-last fp16; every kernel takes a stream; nothing allocates.
#include <algorithm>
#include <memory>
#include <vector>

#include "tc_conv.cuh"

namespace sdfg {

int make_tensor_map_16_4d(CUtensorMap* out, const void* base, uint32_t B, uint32_t H, uint32_t W, uint32_t C, uint32_t box_w, uint32_t box_h);

static int optin_smem_conv(const void* fn, uint32_t smem) {
    struct Done { int dev; uint32_t smem; };
    static thread_local std::vector<Done> done;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(SDFG_ERR_CUDA, "conv: cudaGetDevice failed");
    for (const Done& d : done)
        if (d.dev == dev && d.smem >= smem) return SDFG_OK;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return set_error(SDFG_ERR_CUDA, "conv: cannot opt in to %u bytes of shared memory", smem);
    done.push_back({dev, smem});
    return SDFG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nhwc16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, uint64_t n4) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = ldg_stream4(reinterpret_cast<const float4*>(in) + i);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(tc::pack_f16_sat(v.x, v.y), tc::pack_f16_sat(v.z, v.w));
}

// demod[b, o] = rsqrt(sum_{i, tap} (scale * W[o, i, tap] * s[b, i])^2 + 1e-8)                      block = (o, b), threads over i * taps
__global__ void __launch_bounds__(256) demod_kernel(const float* __restrict__ W, const float* __restrict__ style, float scale, uint32_t Cin,
                                                     uint32_t Cout, uint32_t taps, float* __restrict__ demod) {
    __shared__ float red[8];
    const uint32_t o = blockIdx.x, b = blockIdx.y;
    const float* Wo = W + (size_t)o * Cin * taps;
    float acc = 0.f;
    for (uint32_t k = threadIdx.x; k < Cin * taps; k += blockDim.x) {
        const float w = scale * __ldg(Wo + k) * __ldg(style + (size_t)b * Cin + k / taps);
        acc = fmaf(w, w, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += red[i];
        demod[(size_t)b * Cout + o] = rsqrtf(t + 1e-8f);
    }
}

// Wf[b][tap][o][i] = fp16(scale * W[o, i, tap] * s[b, i] * demod[b, o])          (W is [Cout, Cin, taps] row-major: the reference's
// weight[0, o, i, a, b'] with tap = a * k + b')            grid (Cout, B), threads over i; one 16-bit store per (tap, i)
__global__ void __launch_bounds__(256) modconv_fold_kernel(const float* __restrict__ W, const float* __restrict__ style, const float* __restrict__ demod,
                                                            float scale, uint32_t Cin, uint32_t Cout, uint32_t taps, uint16_t* __restrict__ out) {
    const uint32_t o = blockIdx.x, b = blockIdx.y;
    const float d = demod ? __ldg(demod + (size_t)b * Cout + o) : 1.f;
    for (uint32_t i = threadIdx.x; i < Cin; i += blockDim.x) {
        const float s = scale * d * __ldg(style + (size_t)b * Cin + i);
        for (uint32_t t = 0; t < taps; t++)
            out[(((size_t)b * taps + t) * Cout + o) * Cin + i] = __half_as_ushort(__float2half_rn(s * __ldg(W + ((size_t)o * Cin + i) * taps + t)));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Up-sampling StyledConv, second half.  Y[b, (y, x), tap * C + o] holds the nine tap products of the transposed convolution
// (conv_transpose2d, stride 2: T[2y + a, 2x + b'] += Y[(y, x), (a, b')], T is (2H + 1)^2).  out = lrelu(blur(T) + noise + bias) * sqrt(2)
// with blur = upfirdn2d(T, outer([1,3,3,1]) / 16, pad (1, 1)):  out[Y, X] = sum_{p, q < 4} k[p] k[q] T[Y + p - 1, X + q - 1].
// Block: 16 x 16 output pixels x 64 channels.  The 19 x 19 patch of T is assembled in shared memory as fp16 (<= 4 reads of Y per
// element, 16-byte loads: a thread owns 8 channels), then every thread blurs its pixels from shared memory and writes 16 bytes.
// HBM: Y is read ~1.4x (patch halo), the output written once.
constexpr int UG_T = 16, UG_P = UG_T + 3;
__global__ void __launch_bounds__(256) upconv_gather_kernel(const uint16_t* __restrict__ Y, uint32_t B, uint32_t H, uint32_t W, uint32_t C,
                                                             const float* __restrict__ bias, const float* __restrict__ noise,
                                                             const float* __restrict__ noise_w, uint16_t* __restrict__ out) {
    __shared__ uint4 T[UG_P * UG_P][8];                                // [position][8 channel groups of 8 fp16]
    const uint32_t cg = threadIdx.x & 7, slot = threadIdx.x >> 3;       // 8 threads per position / pixel, 32 positions per pass
    const uint32_t Ho = 2 * H, Wo = 2 * W;
    const uint32_t c0 = blockIdx.y * 64 + cg * 8;
    const uint32_t tiles_x = (Wo + UG_T - 1) / UG_T, tiles_y = (Ho + UG_T - 1) / UG_T;
    const uint32_t b = blockIdx.x / (tiles_x * tiles_y), t = blockIdx.x % (tiles_x * tiles_y);
    const int Y0 = (int)(t / tiles_x) * UG_T, X0 = (int)(t % tiles_x) * UG_T;
    const size_t ldy = (size_t)9 * C;
    const uint16_t* Yb = Y + (size_t)b * H * W * ldy + c0;
    // Positions of the patch by parity class: tile origins are multiples of 16, so patch row tr is T row Y0 + tr - 1 -- ODD for even tr.
    // An odd T row receives tap a = 1 only, an even one taps a = 0 and 2 (likewise for columns): 1, 2, 2 or 4 reads of Y per position.
    // One loop per class keeps the bodies branch-free, so the loads of several positions are in flight together.
    auto ld8 = [&](int y, int x, int tap, float (&acc)[8]) {
        const bool ok = y >= 0 && x >= 0 && y < (int)H && x < (int)W;
        const uint4 v = ok ? __ldg(reinterpret_cast<const uint4*>(Yb + ((size_t)y * W + x) * ldy + (size_t)tap * C)) : make_uint4(0, 0, 0, 0);
        const uint32_t hw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float2 f = tc::unpack_f16(hw[k]);
            acc[2 * k] += f.x; acc[2 * k + 1] += f.y;
        }
    };
    auto put = [&](int tr, int tcn, const float (&acc)[8]) {
        T[tr * UG_P + tcn][cg] = make_uint4(tc::pack_f16_sat(acc[0], acc[1]), tc::pack_f16_sat(acc[2], acc[3]), tc::pack_f16_sat(acc[4], acc[5]), tc::pack_f16_sat(acc[6], acc[7]));
    };
    constexpr int NE = (UG_P + 1) / 2, NO = UG_P / 2;                  // even / odd patch indices: 10 / 9
    // (even tr, even tc): T row and column odd -> tap (1, 1)
#pragma unroll 2
    for (int e = slot; e < NE * NE; e += 32) {
        const int tr = 2 * (e / NE), tcn = 2 * (e % NE);
        const int r = Y0 + tr - 1, c = X0 + tcn - 1;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        ld8((r - 1) >> 1, (c - 1) >> 1, 4, acc);
        put(tr, tcn, acc);
    }
    // (even tr, odd tc): row odd (a = 1), column even (b' = 0, 2)
#pragma unroll 2
    for (int e = slot; e < NE * NO; e += 32) {
        const int tr = 2 * (e / NO), tcn = 2 * (e % NO) + 1;
        const int r = Y0 + tr - 1, c = X0 + tcn - 1;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        ld8((r - 1) >> 1, c >> 1, 3, acc);
        ld8((r - 1) >> 1, (c >> 1) - 1, 5, acc);
        put(tr, tcn, acc);
    }
    // (odd tr, even tc): row even (a = 0, 2), column odd (b' = 1)
#pragma unroll 2
    for (int e = slot; e < NO * NE; e += 32) {
        const int tr = 2 * (e / NE) + 1, tcn = 2 * (e % NE);
        const int r = Y0 + tr - 1, c = X0 + tcn - 1;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        ld8(r >> 1, (c - 1) >> 1, 1, acc);
        ld8((r >> 1) - 1, (c - 1) >> 1, 7, acc);
        put(tr, tcn, acc);
    }
    // (odd tr, odd tc): both even -> four taps
#pragma unroll 2
    for (int e = slot; e < NO * NO; e += 32) {
        const int tr = 2 * (e / NO) + 1, tcn = 2 * (e % NO) + 1;
        const int r = Y0 + tr - 1, c = X0 + tcn - 1;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        ld8(r >> 1, c >> 1, 0, acc);
        ld8(r >> 1, (c >> 1) - 1, 2, acc);
        ld8((r >> 1) - 1, c >> 1, 6, acc);
        ld8((r >> 1) - 1, (c >> 1) - 1, 8, acc);
        put(tr, tcn, acc);
    }
    __syncthreads();
    const float nw = (noise && noise_w) ? __ldg(noise_w) : 0.f;
    float bs[8];
#pragma unroll
    for (int k = 0; k < 8; k++) bs[k] = bias ? __ldg(bias + c0 + k) : 0.f;
    const float k4[4] = {0.25f, 0.75f, 0.75f, 0.25f};                  // [1,3,3,1] / 4 per axis: make_kernel (outer / 64) * upsample_factor^2 (Blur :522-531)
    for (int e = slot; e < UG_T * UG_T; e += 32) {
        const int oy = e / UG_T, ox = e % UG_T;
        const int Yo = Y0 + oy, Xo = X0 + ox;
        if (Yo >= (int)Ho || Xo >= (int)Wo) continue;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int p = 0; p < 4; p++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 v = T[(oy + p) * UG_P + ox + q][cg];
                const uint32_t hw[4] = {v.x, v.y, v.z, v.w};
                const float kk = k4[p] * k4[q];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float2 f = tc::unpack_f16(hw[k]);
                    acc[2 * k] = fmaf(kk, f.x, acc[2 * k]); acc[2 * k + 1] = fmaf(kk, f.y, acc[2 * k + 1]);
                }
            }
        }
        const size_t pix = ((size_t)b * Ho + Yo) * Wo + Xo;
        const float nz = nw != 0.f ? nw * __ldg(noise + pix) : 0.f;
        uint32_t h[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float v0 = acc[2 * k] + bs[2 * k] + nz, v1 = acc[2 * k + 1] + bs[2 * k + 1] + nz;
            v0 = (v0 > 0.f ? v0 : 0.2f * v0) * 1.4142135623730951f;
            v1 = (v1 > 0.f ? v1 : 0.2f * v1) * 1.4142135623730951f;
            h[k] = tc::pack_f16_sat(v0, v1);
        }
        *reinterpret_cast<uint4*>(out + pix * C + c0) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}

// wrgb16[b * 8 + o, i] = fp16(scale * style[b, i] * W[o, i]) for o < 3, zero rows 3..7: the B operand of the ToRGB GEMM (8 rows per CTA)
__global__ void __launch_bounds__(256) rgb_weight16_kernel(const float* __restrict__ W, const float* __restrict__ style, float scale, uint32_t C,
                                                            uint32_t B, __half* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const uint32_t b = i / C, k = i % C;
    const float sc = scale * __ldg(style + i);
#pragma unroll
    for (uint32_t o = 0; o < 8; o++) out[((size_t)b * 8 + o) * C + k] = __float2half_rn(o < 3 ? sc * __ldg(W + o * C + k) : 0.f);
}

}  // namespace sdfg

using namespace sdfg;

extern "C" int sdfg_nhwc16(const float* in, uint16_t* out, uint64_t n_elems, void* stream) {
    if (n_elems == 0) return SDFG_OK;
    SDFG_REQUIRE(in && out && n_elems % 4 == 0, SDFG_ERR_INVALID, "nhwc16: null pointer or element count not a multiple of 4");
    nhwc16_kernel<<<(unsigned)ceil_div<uint64_t>(n_elems / 4, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n_elems / 4);
    return check_launch("nhwc16_kernel");
}

extern "C" int sdfg_modconv_fold(const float* weight, const float* style, float scale, uint32_t B, uint32_t Cin, uint32_t Cout, uint32_t taps,
                                 int demodulate, float* demod_scratch, uint16_t* out, void* stream) {
    if (B == 0) return SDFG_OK;
    SDFG_REQUIRE(weight && style && out && (!demodulate || demod_scratch), SDFG_ERR_INVALID, "modconv_fold: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (demodulate) {
        demod_kernel<<<dim3(Cout, B), 256, 0, st>>>(weight, style, scale, Cin, Cout, taps, demod_scratch);
        if (int e = check_launch("demod_kernel")) return e;
    }
    modconv_fold_kernel<<<dim3(Cout, B), 256, 0, st>>>(weight, style, demodulate ? demod_scratch : nullptr, scale, Cin, Cout, taps, out);
    return check_launch("modconv_fold_kernel");
}

// pixel tiling, unit split over the CTA pairs, tensor maps and launch of tc_conv_kernel; P carries B, H, W, Cin, taps, ncols, NT, epi and the epilogue pointers
static int launch_conv(tc::ConvParams& P, const uint16_t* x, const uint16_t* wf, uint64_t w_rows, const char* name, cudaStream_t stream) {
    P.bw = std::min(P.W, 128u); P.bh = 128 / P.bw;
    P.tiles_x = P.W / P.bw; P.tiles_y = ceil_div<uint32_t>(P.H, P.bh);
    P.pairs_per_sample = ceil_div<uint32_t>(P.tiles_x * P.tiles_y, 2);
    P.n_nt = P.ncols / P.NT;
    P.n_units = P.B * P.pairs_per_sample * P.n_nt;
    const uint32_t pairs = std::max(1u, std::min<uint32_t>((uint32_t)sm_count() / 2, P.n_units));
    P.units_per_pair = ceil_div<uint32_t>(P.n_units, pairs);
    const uint32_t grid = 2 * ceil_div<uint32_t>(P.n_units, P.units_per_pair);
    CUtensorMap tmA, tmB;
    if (int e = make_tensor_map_16_4d(&tmA, x, P.B, P.H, P.W, P.Cin, P.bw, P.bh)) return e;
    if (int e = make_tensor_map_16(&tmB, wf, w_rows, P.Cin, P.Cin, P.NT / 2, 64, tc::FMT_F16)) return e;
    const uint32_t smem = tc::conv_smem_bytes(P.NT);
    if (int e = optin_smem_conv((const void*)tc::tc_conv_kernel, smem)) return e;
    ProfScope prof(name, stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc::CV_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, tc::tc_conv_kernel, tmA, tmB, P) != cudaSuccess) { (void)check_launch(name); return SDFG_ERR_CUDA; }
    return check_launch(name);
}

extern "C" int sdfg_conv_forward(const uint16_t* x, const uint16_t* wf, uint32_t B, uint32_t H, uint32_t W, uint32_t Cin, uint32_t Cout, uint32_t taps,
                                 int gemm_mode, const float* bias, const float* noise, const float* noise_w, uint16_t* out, void* stream) {
    if (B == 0) return SDFG_OK;
    SDFG_REQUIRE(x && wf && out, SDFG_ERR_INVALID, "conv_forward: null pointer");
    SDFG_REQUIRE(Cin % 64 == 0 && Cout % 128 == 0, SDFG_ERR_UNSUPPORTED, "conv_forward: Cin must be a multiple of 64 and Cout of 128 (got %u, %u)", Cin, Cout);
    SDFG_REQUIRE(taps == 9 || taps == 1, SDFG_ERR_UNSUPPORTED, "conv_forward: 3 x 3 (taps = 9) or 1 x 1 (taps = 1) only");
    SDFG_REQUIRE(W >= 8 && (W & (W - 1)) == 0, SDFG_ERR_UNSUPPORTED, "conv_forward: width must be a power of two >= 8 (got %u)", W);
    tc::ConvParams P = {};
    P.B = B; P.H = H; P.W = W; P.Cin = Cin;
    // gemm_mode: the nine tap matrices are NOT summed over shifted inputs but laid side by side as 9 * Cout output columns
    P.taps = gemm_mode ? 1 : taps;
    P.ncols = gemm_mode ? taps * Cout : Cout;
    P.wrows_per_sample = taps * Cout;
    P.NT = (P.ncols % 256 == 0) ? 256 : 128;
    P.epi = gemm_mode ? tc::EPI_RAW : tc::EPI_ACT;
    P.bias = bias; P.noise = noise; P.noise_w = noise_w; P.out = out; P.ld_out = P.ncols;
    SDFG_REQUIRE(gemm_mode || P.ncols <= 2304, SDFG_ERR_UNSUPPORTED, "conv_forward: too many output channels");      // bias table in shared memory
    return launch_conv(P, x, wf, (uint64_t)B * taps * Cout, "tc_conv_kernel<gemm>", (cudaStream_t)stream);
}

extern "C" int sdfg_upconv_gather(const uint16_t* y, uint32_t B, uint32_t H, uint32_t W, uint32_t C, const float* bias, const float* noise,
                                  const float* noise_w, uint16_t* out, void* stream) {
    if (B == 0) return SDFG_OK;
    SDFG_REQUIRE(y && out && C % 64 == 0, SDFG_ERR_INVALID, "upconv_gather: null pointer or channels not a multiple of 64");
    const uint32_t tiles = ceil_div<uint32_t>(2 * W, UG_T) * ceil_div<uint32_t>(2 * H, UG_T);
    upconv_gather_kernel<<<dim3(B * tiles, C / 64), 256, 0, (cudaStream_t)stream>>>(y, B, H, W, C, bias, noise, noise_w, out);
    return check_launch("upconv_gather_kernel");
}

extern "C" int sdfg_to_rgb(const uint16_t* x, const float* weight, const float* style, float scale, const float* bias, const float* skip,
                           uint32_t B, uint32_t H, uint32_t W, uint32_t C, float* wrgb_scratch, float* out_nhwc, float* out_nchw, void* stream) {
    if (B == 0) return SDFG_OK;
    SDFG_REQUIRE(x && weight && style && bias && wrgb_scratch && (out_nhwc || out_nchw), SDFG_ERR_INVALID, "to_rgb: null pointer");
    SDFG_REQUIRE(C % 64 == 0, SDFG_ERR_UNSUPPORTED, "to_rgb: channels must be a multiple of 64 (got %u)", C);
    SDFG_REQUIRE(W >= 8 && (W & (W - 1)) == 0, SDFG_ERR_UNSUPPORTED, "to_rgb: width must be a power of two >= 8 (got %u)", W);
    cudaStream_t st = (cudaStream_t)stream;
    // the scratch ([B, C, 4] floats = 16 B per sample and channel) holds the fp16 weight rows [B * 8, C]
    rgb_weight16_kernel<<<ceil_div<uint32_t>(B * C, 256), 256, 0, st>>>(weight, style, scale, C, B, reinterpret_cast<__half*>(wrgb_scratch));
    if (int e = check_launch("rgb_weight16_kernel")) return e;
    tc::ConvParams P = {};
    P.B = B; P.H = H; P.W = W; P.Cin = C;
    P.taps = 1; P.ncols = 16; P.NT = 16; P.wrows_per_sample = 8;        // CTA 1's weight rows (8..15 of the tile) are the next sample's / out of range: columns never read
    P.epi = tc::EPI_RGB;
    P.bias = bias; P.skip = skip; P.out_nhwc = out_nhwc; P.out_nchw = out_nchw;
    return launch_conv(P, x, reinterpret_cast<const uint16_t*>(wrgb_scratch), (uint64_t)B * 8, "tc_conv_kernel<rgb>", st);
}
