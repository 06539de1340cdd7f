// StyleGAN2 decoder (ref sdf_model.py:614-1056 + sdf_op.py) -- forward kernels behind the C ABI (include/sdfg.h):
//   sdfg_nhwc16                 fp32 [B, H*W, C] (the renderer's feature map, channels last) -> fp16 [B, H, W, C]
//   sdfg_modconv_fold           per-sample weights of a ModulatedConv2d: scale * W * style, demodulated           (:655-669)
//   sdfg_conv_forward           3 x 3 / 1 x 1 convolution on tcgen05 (tc_conv.cuh) + noise + bias + leaky ReLU           (:790-818)
//   sdfg_upconv_forward         transposed convolution (four parity classes on tcgen05) -> blur -> noise + bias + leaky ReLU (:671-684, Blur :522-538)
//   sdfg_to_rgb                 1 x 1 modulated convolution to 3 channels (tcgen05, N = 16) + bias + up-sampled skip     (:821-843, Upsample :480-499)
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

// demod[b, o] = rsqrt(sum_{i, tap} (scale * W[o, i, tap] * s[b, i])^2 + 1e-8) = rsqrt(scale^2 * sum_i s[b, i]^2 * W2[o, i] + 1e-8) with
// W2[o, i] = sum_tap W[o, i, tap]^2: a block owns output channel o, squares its weight row ONCE into shared memory and serves every
// sample from it (a block per (o, b) re-read the row B times: 0.3 ms per configs[2] pass).      grid Cout, dynamic smem Cin floats
__global__ void __launch_bounds__(256) demod_kernel(const float* __restrict__ W, const float* __restrict__ style, float scale, uint32_t B, uint32_t Cin,
                                                     uint32_t Cout, uint32_t taps, float* __restrict__ demod) {
    extern __shared__ float w2[];
    const uint32_t o = blockIdx.x;
    const float* Wo = W + (size_t)o * Cin * taps;
    for (uint32_t i = threadIdx.x; i < Cin; i += blockDim.x) {
        float a = 0.f;
        for (uint32_t t = 0; t < taps; t++) { const float w = __ldg(Wo + i * taps + t); a = fmaf(w, w, a); }
        w2[i] = a;
    }
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t b = warp; b < B; b += blockDim.x >> 5) {
        float acc = 0.f;
        for (uint32_t i = lane; i < Cin; i += 32) { const float sv = __ldg(style + (size_t)b * Cin + i); acc = fmaf(sv * sv, w2[i], acc); }
        acc = warp_sum(acc);
        if (lane == 0) demod[(size_t)b * Cout + o] = rsqrtf(fmaf(scale * scale, acc, 1e-8f));
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
// Up-sampling StyledConv, second half: out = lrelu(blur(T) + noise + bias) * sqrt(2), T [B, 2H + 1, 2W + 1, C] fp16 = the transposed
// convolution written by tc_conv_kernel, blur = upfirdn2d(T, outer([1,3,3,1]) / 16 * 4, pad (1, 1)):
//     out[Y, X] = sum_{p, q < 4} k[p] k[q] T[Y + p - 1, X + q - 1]           (ref Blur :522-531 after conv_transpose2d :671-684)
// Block: 16 x 16 output pixels x 64 channels.  The 19 x 19 patch of T goes to shared memory with 16-byte loads (a thread owns 8
// channels); the blur is separable: a thread owns a column of 8 output pixels, forms the horizontal 4-tap sums of the 11 patch rows it
// needs once and combines them vertically -- 8.5 instead of 16 taps per output.  HBM: T is read ~1.4x (halo, mostly L2 hits), the
// output written once.
constexpr int UG_T = 16, UG_P = UG_T + 3;
__global__ void __launch_bounds__(256, 4) upconv_blur_kernel(const uint16_t* __restrict__ T, uint32_t B, uint32_t H, uint32_t W, uint32_t C,
                                                           const float* __restrict__ bias, const float* __restrict__ noise,
                                                           const float* __restrict__ noise_w, uint16_t* __restrict__ out) {
    __shared__ uint4 P[UG_P * UG_P][8];                                // [position][8 channel groups of 8 fp16]
    const uint32_t cg = threadIdx.x & 7, slot = threadIdx.x >> 3;       // 8 threads per position / pixel, 32 positions per pass
    const uint32_t Ho = 2 * H, Wo = 2 * W, Ht = Ho + 1, Wt = Wo + 1;
    const uint32_t c0 = blockIdx.y * 64 + cg * 8;
    const uint32_t tiles_x = (Wo + UG_T - 1) / UG_T, tiles_y = (Ho + UG_T - 1) / UG_T;
    const uint32_t b = blockIdx.x / (tiles_x * tiles_y), t = blockIdx.x % (tiles_x * tiles_y);
    const int Y0 = (int)(t / tiles_x) * UG_T, X0 = (int)(t % tiles_x) * UG_T;
    const uint16_t* Tb = T + (size_t)b * Ht * Wt * C + c0;
    // global -> shared without a register round trip: all 12 copies of a thread are in flight together (with ld + st the loop ran 3-4
    // loads deep and the load phase, not the arithmetic, set the block time); out-of-image positions are zero-filled (src-size 0)
#pragma unroll
    for (int it = 0; it < (UG_P * UG_P + 31) / 32; it++) {
        const int e = slot + it * 32;
        if (e < UG_P * UG_P) {
            const int r = Y0 + e / UG_P - 1, c = X0 + e % UG_P - 1;
            const bool ok = r >= 0 && c >= 0 && r < (int)Ht && c < (int)Wt;
            const uint16_t* src = ok ? Tb + ((size_t)r * Wt + c) * C : Tb;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(&P[e][cg])), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float nw = (noise && noise_w) ? __ldg(noise_w) : 0.f;
    uint64_t bs2[4];                                                    // bias of this thread's 8 channels as packed fp32 pairs
#pragma unroll
    for (int k = 0; k < 4; k++) bs2[k] = bias ? tc::pk2(__ldg(bias + c0 + 2 * k), __ldg(bias + c0 + 2 * k + 1)) : 0ull;
    // [1,3,3,1] / 4 per axis: make_kernel (outer / 64) * upsample_factor^2.  The horizontal pass runs in packed halves (weights 0.25 /
    // 0.75 are exact, a 4-term sum of fp16 inputs: one more fp16 rounding, the size of the one T already carries), the vertical pass and
    // the epilogue in packed fp32 pairs: ~90 instructions per pixel and 8 channels instead of ~170 for the plain fp32 form, which
    // left the kernel issue-bound at a third of the HBM rate (ncu r02fdec).
    const __half2 kq = __floats2half2_rn(0.25f, 0.25f), kt = __floats2half2_rn(0.75f, 0.75f);
    const uint64_t kq2 = tc::pk2(0.25f, 0.25f), kt2 = tc::pk2(0.75f, 0.75f);
    const uint64_t a_pos = tc::pk2(1.4142135623730951f, 1.4142135623730951f), a_neg = tc::pk2(0.2f * 1.4142135623730951f, 0.2f * 1.4142135623730951f);
    const int ox = slot & 15, oy0 = (slot >> 4) * 8;                    // this thread: output column ox, rows oy0 .. oy0 + 7
    const int Xo = X0 + ox;
    uint64_t hs[4][4];                                                  // sliding window of horizontal sums (patch rows oy + 0 .. 3), 4 fp32 pairs each
    auto hsum = [&](int pr, uint64_t (&h)[4]) {
        const uint4 v0 = P[pr * UG_P + ox][cg], v1 = P[pr * UG_P + ox + 1][cg], v2 = P[pr * UG_P + ox + 2][cg], v3 = P[pr * UG_P + ox + 3][cg];
        const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w}, w2[4] = {v2.x, v2.y, v2.z, v2.w}, w3[4] = {v3.x, v3.y, v3.z, v3.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            __half2 a = __hmul2(kq, *reinterpret_cast<const __half2*>(&w0[k]));
            a = __hfma2(kt, *reinterpret_cast<const __half2*>(&w1[k]), a);
            a = __hfma2(kt, *reinterpret_cast<const __half2*>(&w2[k]), a);
            a = __hfma2(kq, *reinterpret_cast<const __half2*>(&w3[k]), a);
            const float2 fa = __half22float2(a);
            h[k] = tc::pk2(fa.x, fa.y);
        }
    };
    hsum(oy0, hs[0]); hsum(oy0 + 1, hs[1]); hsum(oy0 + 2, hs[2]);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        hsum(oy0 + j + 3, hs[(j + 3) & 3]);
        const int Yo = Y0 + oy0 + j;
        if (Yo >= (int)Ho || Xo >= (int)Wo) continue;
        const size_t pix = ((size_t)b * Ho + Yo) * Wo + Xo;
        const float nz = nw != 0.f ? nw * __ldg(noise + pix) : 0.f;
        const uint64_t nz2 = tc::pk2(nz, nz);
        uint32_t h[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t v = tc::add2(bs2[k], nz2);
            v = tc::fma2(kq2, hs[j & 3][k], v);
            v = tc::fma2(kt2, hs[(j + 1) & 3][k], v);
            v = tc::fma2(kt2, hs[(j + 2) & 3][k], v);
            v = tc::fma2(kq2, hs[(j + 3) & 3][k], v);
            float p0, p1, n0, n1;
            tc::upk2(tc::mul2(v, a_pos), p0, p1);                       // leaky_relu(v, 0.2) * sqrt(2) = max(v * sqrt 2, v * 0.2 sqrt 2)
            tc::upk2(tc::mul2(v, a_neg), n0, n1);
            h[k] = tc::pack_f16_sat(fmaxf(p0, n0), fmaxf(p1, n1));
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
        demod_kernel<<<Cout, 256, Cin * sizeof(float), st>>>(weight, style, scale, B, Cin, Cout, taps, demod_scratch);
        if (int e = check_launch("demod_kernel")) return e;
    }
    modconv_fold_kernel<<<dim3(Cout, B), 256, 0, st>>>(weight, style, demodulate ? demod_scratch : nullptr, scale, Cin, Cout, taps, out);
    return check_launch("modconv_fold_kernel");
}

// unit split over the CTA pairs, tensor maps and launch of tc_conv_kernel; P carries the shapes, the tap classes, the pixel tile (bw x bh;
// tiles_x/y default to covering H x W), NT, epi and the epilogue pointers
static int launch_conv(tc::ConvParams& P, const uint16_t* x, const uint16_t* wf, uint64_t w_rows, const char* name, cudaStream_t stream) {
    if (P.tiles_x == 0) { P.tiles_x = P.W / P.bw; P.tiles_y = ceil_div<uint32_t>(P.H, P.bh); }
    P.pairs_per_sample = ceil_div<uint32_t>(P.tiles_x * P.tiles_y, 2);
    P.n_nt = P.ncols / P.NT;
    P.n_units = P.B * P.pairs_per_sample * P.n_cls * P.n_nt;
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

// tap table of a 3 x 3 (stride 1, zero padding 1) or 1 x 1 convolution
static void plain_conv_class(tc::ConvClass& K, uint32_t taps) {
    K.n_taps = taps; K.oy = 0; K.ox = 0;
    for (uint32_t t = 0; t < taps; t++) {
        K.dy[t] = taps == 9 ? (int8_t)((int)(t / 3) - 1) : 0;
        K.dx[t] = taps == 9 ? (int8_t)((int)(t % 3) - 1) : 0;
        K.wtap[t] = (uint8_t)t;
    }
}

extern "C" int sdfg_conv_forward(const uint16_t* x, const uint16_t* wf, uint32_t B, uint32_t H, uint32_t W, uint32_t Cin, uint32_t Cout, uint32_t taps,
                                 const float* bias, const float* noise, const float* noise_w, uint16_t* out, void* stream) {
    if (B == 0) return SDFG_OK;
    SDFG_REQUIRE(x && wf && out, SDFG_ERR_INVALID, "conv_forward: null pointer");
    SDFG_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, SDFG_ERR_UNSUPPORTED, "conv_forward: Cin and Cout must be multiples of 64 (got %u, %u)", Cin, Cout);
    SDFG_REQUIRE(taps == 9 || taps == 1, SDFG_ERR_UNSUPPORTED, "conv_forward: 3 x 3 (taps = 9) or 1 x 1 (taps = 1) only");
    SDFG_REQUIRE(W >= 8 && (W & (W - 1)) == 0, SDFG_ERR_UNSUPPORTED, "conv_forward: width must be a power of two >= 8 (got %u)", W);
    SDFG_REQUIRE(Cout <= 2304, SDFG_ERR_UNSUPPORTED, "conv_forward: too many output channels");      // bias table in shared memory
    tc::ConvParams P = {};
    P.B = B; P.H = H; P.W = W; P.Cin = Cin;
    P.n_cls = 1; plain_conv_class(P.cls[0], taps);
    P.y_end = H; P.x_end = W; P.sy = P.sx = 1; P.out_h = H; P.out_w = W;
    P.ncols = Cout; P.wrows_per_sample = taps * Cout;
    P.NT = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : 64;
    P.epi = tc::EPI_ACT;
    P.bias = bias; P.noise = noise; P.noise_w = noise_w; P.out = out; P.ld_out = Cout;
    P.bw = std::min(W, 128u); P.bh = 128 / P.bw;
    return launch_conv(P, x, wf, (uint64_t)B * taps * Cout, "tc_conv_kernel<conv>", (cudaStream_t)stream);
}

extern "C" int sdfg_upconv_forward(const uint16_t* x, const uint16_t* wf, uint32_t B, uint32_t H, uint32_t W, uint32_t Cin, uint32_t Cout,
                                   uint16_t* t_scratch, const float* bias, const float* noise, const float* noise_w, uint16_t* out, void* stream) {
    if (B == 0) return SDFG_OK;
    SDFG_REQUIRE(x && wf && t_scratch && out, SDFG_ERR_INVALID, "upconv_forward: null pointer");
    SDFG_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, SDFG_ERR_UNSUPPORTED, "upconv_forward: Cin and Cout must be multiples of 64 (got %u, %u)", Cin, Cout);
    SDFG_REQUIRE(W >= 8 && (W & (W - 1)) == 0, SDFG_ERR_UNSUPPORTED, "upconv_forward: width must be a power of two >= 8 (got %u)", W);
    cudaStream_t st = (cudaStream_t)stream;
    tc::ConvParams P = {};
    P.B = B; P.H = H; P.W = W; P.Cin = Cin;
    P.sy = P.sx = 2; P.out_h = 2 * H + 1; P.out_w = 2 * W + 1;
    P.ncols = Cout; P.wrows_per_sample = 9 * Cout;
    P.NT = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : 64;
    P.epi = tc::EPI_RAW;
    P.out = t_scratch; P.ld_out = Cout;
    // class (cy, cx): T[2y' + cy, 2x' + cx] = sum over taps a = cy (mod 2), b' = cx (mod 2) of in[y' - (a - cy) / 2, x' - (b' - cx) / 2] * W[a, b']
    auto fill = [&](tc::ConvClass& K, int cy, int cx) {
        K.n_taps = 0; K.oy = cy; K.ox = cx;
        for (int a = cy; a < 3; a += 2)
            for (int bb = cx; bb < 3; bb += 2) {
                K.dy[K.n_taps] = (int8_t)(-(a - cy) / 2); K.dx[K.n_taps] = (int8_t)(-(bb - cx) / 2);
                K.wtap[K.n_taps] = (uint8_t)(a * 3 + bb);
                K.n_taps++;
            }
    };
    // interior: y' < H, x' < W, all four classes
    P.n_cls = 4;
    fill(P.cls[0], 0, 0); fill(P.cls[1], 0, 1); fill(P.cls[2], 1, 0); fill(P.cls[3], 1, 1);
    P.y_org = P.x_org = 0; P.y_end = H; P.x_end = W;
    P.bw = std::min(W, 128u); P.bh = 128 / P.bw; P.tiles_x = W / P.bw; P.tiles_y = ceil_div<uint32_t>(H, P.bh);
    if (int e = launch_conv(P, x, wf, (uint64_t)B * 9 * Cout, "tc_conv_kernel<upconv>", st)) return e;
    // bottom strip: y' = H (output row 2H), x' < W, the classes with cy = 0
    P.n_cls = 2;
    fill(P.cls[0], 0, 0); fill(P.cls[1], 0, 1);
    P.y_org = H; P.y_end = H + 1; P.x_org = 0; P.x_end = W;
    P.tiles_x = W / P.bw; P.tiles_y = 1;
    if (int e = launch_conv(P, x, wf, (uint64_t)B * 9 * Cout, "tc_conv_kernel<upconv edge>", st)) return e;
    // right strip: x' = W (output column 2W), y' <= H, the classes with cx = 0 (the odd row 2H + 1 of class (1, 0) falls outside T)
    fill(P.cls[0], 0, 0); fill(P.cls[1], 1, 0);
    P.x_org = W; P.x_end = W + 1; P.y_org = 0; P.y_end = H + 1;
    P.bw = 8; P.bh = 16; P.tiles_x = 1; P.tiles_y = ceil_div<uint32_t>(H + 1, 16);
    if (int e = launch_conv(P, x, wf, (uint64_t)B * 9 * Cout, "tc_conv_kernel<upconv edge>", st)) return e;
    const uint32_t tiles = ceil_div<uint32_t>(2 * W, UG_T) * ceil_div<uint32_t>(2 * H, UG_T);
    upconv_blur_kernel<<<dim3(B * tiles, Cout / 64), 256, 0, st>>>(t_scratch, B, H, W, Cout, bias, noise, noise_w, out);
    return check_launch("upconv_blur_kernel");
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
    P.n_cls = 1; plain_conv_class(P.cls[0], 1);
    P.y_end = H; P.x_end = W; P.sy = P.sx = 1; P.out_h = H; P.out_w = W;
    P.bw = std::min(W, 128u); P.bh = 128 / P.bw;
    P.ncols = 16; P.NT = 16; P.wrows_per_sample = 8;        // CTA 1's weight rows (8..15 of the tile) are the next sample's / out of range: columns never read
    P.epi = tc::EPI_RGB;
    P.bias = bias; P.skip = skip; P.out_nhwc = out_nhwc; P.out_nchw = out_nchw;
    return launch_conv(P, x, reinterpret_cast<const uint16_t*>(wrgb_scratch), (uint64_t)B * 8, "tc_conv_kernel<rgb>", st);
}
