// StyleGAN2 decoder convolutions on the tensor cores: one persistent, warp-specialised tcgen05 kernel on CTA pairs.
//
//   3 x 3 convolution (stride 1, zero padding 1) as an implicit GEMM over channels-last fp16 activations [B, H, W, Cin]:
//       out[b, y, x, o] = sum_{tap = (a, b')} sum_i  in[b, y + a - 1, x + b' - 1, i] * Wf[b][tap][o][i]
//   The A operand of tap (a, b') is the SAME 4-D TMA box shifted by (a - 1, b' - 1): out-of-image rows / columns are zero-filled by
//   the TMA unit, which is the convolution's zero padding; no im2col buffer exists.  Wf are per-SAMPLE weights with the style
//   modulation and the demodulation folded in (ref ModulatedConv2d.forward sdf_model.py:655-704 builds the same per-sample
//   weights and runs a grouped convolution), [B * taps * Cout, Cin] fp16, streamed as [NT/2 x 64] boxes per CTA.
//   The transposed convolution of the up-sampling layers (conv_transpose2d, stride 2: T[2y + a, 2x + b'] += in[y, x] * W[a, b'], T is
//   (2H + 1)^2) runs as its four output-parity CLASSES: T[2y' + cy, 2x' + cx] is an ordinary convolution of the low-resolution input with
//   the taps a = cy (mod 2), b' = cx (mod 2) -- 4, 2, 2 and 1 taps, shifts 0 / -1 -- i.e. the same implicit GEMM with a per-class tap
//   table and a stride-2 store.  The same 9 * Cin * Cout MACs per input pixel as the tap-by-tap product, but no 9 * Cout-wide
//   intermediate: T is written once (fp16) and the blur kernel (conv.cu) reads it once.  Row 2H / column 2W of T (y' = H, x' = W) are
//   two thin extra launches of the same kernel over a one-tile-wide pixel grid.
//   epilogue (EPI_ACT):  v = acc + noise_w * noise[b, y, x] + bias[o];  v = leaky_relu(v, 0.2) * sqrt(2)  -> fp16   (NoiseInjection
//   sdf_model.py:783-790 + FusedLeakyReLU sdf_op.py:83-117);  EPI_RAW: fp16 store of the accumulator.
//   EPI_RGB (ToRGB, sdf_model.py:887-909): the 1 x 1 modulated convolution to 3 channels as a GEMM with NT = 16 (8 weight rows per CTA,
//   3 of them real) -- the tensor pipe idles, the point is that the activation streams through TMA at the HBM rate instead of
//   through 100 SIMT instructions per pixel and 16 channels; the epilogue adds the bias and the up-sampled skip image
//   (upfirdn2d [1,3,3,1], up 2, pad (2,1): sdf_model.py:624-641) and writes fp32 NHWC (next skip) and / or NCHW (the image).
//
//   warp 0      TMA producer (both CTAs): own 128-pixel A box + own half (NT/2 rows) of the B box per K chunk, ring of NSTG stages
//   warp 1      MMA issuer (leader CTA): tcgen05.mma.cta_group::2, M = 256 (2 x 128 pixels), N = NT, K = 16 per instruction
//   warps 2-5 / 6-9   two epilogue warpgroups, one per TMEM accumulator stage: the epilogue of unit u overlaps the MMAs of unit u + 1
// Per K chunk a CTA stages 16 KB (A) + NT/2 * 128 B (B half): 32 KB per 512 clk of MMA at NT = 256 = 64 B/clk/SM -- the pair
// halves the weight stream; a single-CTA tile would need 96 B/clk/SM, 2.2x the chip's L2 bandwidth at the nominal MMA rate.
#pragma once
#include "tc_common.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t CV_THREADS = 320;
constexpr uint32_t CV_A_BYTES = 128 * 128;                  // [128 pixels x 64 channels] fp16
constexpr uint32_t CV_NSTG = 5;
enum ConvEpi : uint32_t { EPI_RAW = 0, EPI_ACT = 1, EPI_RGB = 2 };

struct ConvClass {
    uint32_t n_taps;            // K steps of this class = n_taps * Cin / 64
    int32_t oy, ox;             // output pixel = (sy * y + oy, sx * x + ox)
    int8_t dy[9], dx[9];        // input shift of each tap
    uint8_t wtap[9];            // which of the sample's tap matrices [tap][Cout][Cin] it multiplies with
};

struct ConvParams {
    uint32_t B, H, W, Cin;      // input activation [B, H, W, Cin]
    uint32_t n_cls;             // 1, or the output-parity classes of a transposed convolution
    ConvClass cls[4];
    uint32_t y_org, x_org;      // origin of the pixel grid the tiles cover (0 except for the edge strips of a transposed convolution)
    uint32_t y_end, x_end;      // pixels at or beyond are not stored
    uint32_t sy, sx;            // output stride (1; 2 for the transposed convolution)
    uint32_t out_h, out_w;      // output image
    uint32_t ncols;             // output columns (Cout)
    uint32_t NT;                // N tile: 16 (ToRGB), 64, 128 or 256 (divides ncols)
    uint32_t bw, bh;            // pixel tile = bw x bh = 128 pixels of one sample
    uint32_t tiles_x, tiles_y;  // pixel tiles per sample
    uint32_t pairs_per_sample;  // ceil(tiles / 2): a unit = 2 adjacent pixel tiles (one per CTA) x one N tile
    uint32_t n_nt;              // ncols / NT
    uint32_t n_units, units_per_pair;
    uint32_t epi;
    uint32_t wrows_per_sample;  // rows of Wf per sample: tap matrices * ncols
    const float* bias;          // [ncols]   (EPI_ACT)
    const float* noise;         // [B, H, W] or NULL
    const float* noise_w;       // device scalar or NULL
    uint16_t* out;              // [B, H, W, ld_out]
    int64_t ld_out;
    const float* skip;          // EPI_RGB: [B, H/2, W/2, 3] fp32 or NULL
    float* out_nhwc;            // EPI_RGB: [B, H, W, 3] or NULL
    float* out_nchw;            // EPI_RGB: [B, 3, H, W] or NULL
};

struct ConvSmem {
    uint64_t full[CV_NSTG], empty[CV_NSTG];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
    alignas(16) float bias[2304];
};

__host__ __device__ inline uint32_t conv_smem_bytes(uint32_t NT) { return 1024 + CV_NSTG * (CV_A_BYTES + NT / 2 * 128) + (uint32_t)sizeof(ConvSmem); }

// 4-D tile load (c0 = channel, c1 = x, c2 = y, c3 = sample) into this CTA's shared memory, transaction bytes counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_4d_2cta(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

__global__ void __launch_bounds__(CV_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t b_bytes = P.NT / 2 * 128, stage_bytes = CV_A_BYTES + b_bytes;
    ConvSmem& S = *reinterpret_cast<ConvSmem*>(smem + CV_NSTG * stage_bytes);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t u_begin = (blockIdx.x / 2) * P.units_per_pair;
    const uint32_t u_end = min(P.n_units, u_begin + P.units_per_pair);
    const uint32_t n_kc = P.Cin / 64;
    const uint32_t tiles = P.tiles_x * P.tiles_y;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < CV_NSTG; i++) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 1); }
        for (uint32_t i = 0; i < 2; i++) { mbar_init(&S.tmem_full[i], 1); mbar_init(&S.tmem_empty[i], 8); }      // 4 epilogue warps of each CTA
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    if (warp == 1) tmem_alloc_2cta(&S.tmem_base, 512);
    if (P.epi == EPI_ACT)
        for (uint32_t i = threadIdx.x; i < P.ncols; i += blockDim.x) S.bias[i] = P.bias ? __ldg(P.bias + i) : 0.f;
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;

    // unit -> (sample, pixel tile of THIS CTA, class, N tile); the N tile is the fastest index, then the class: consecutive units of a
    // pair share their A boxes (L2 hits)
    auto decode = [&](uint32_t u, uint32_t& b, uint32_t& tile, uint32_t& cl, uint32_t& nt) {
        nt = u % P.n_nt;
        uint32_t pp = u / P.n_nt;
        cl = pp % P.n_cls;
        pp /= P.n_cls;
        b = pp / P.pairs_per_sample;
        tile = (pp % P.pairs_per_sample) * 2 + rank;
    };

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t u = u_begin; u < u_end; u++) {
                uint32_t b, tile, cl, nt;
                decode(u, b, tile, cl, nt);
                const ConvClass& K = P.cls[cl];
                // a pixel tile beyond the sample's last one (odd tile count): coordinates past the image -> the box is zero-filled
                const int32_t x0 = (int32_t)(P.x_org + (tile % P.tiles_x) * P.bw), y0 = tile < tiles ? (int32_t)(P.y_org + (tile / P.tiles_x) * P.bh) : (int32_t)(P.H + 8);
                const uint32_t n_k = K.n_taps * n_kc;
                for (uint32_t k = 0; k < n_k; k++) {
                    const uint32_t tap = k / n_kc, kc = k % n_kc;
                    const int32_t dy = K.dy[tap], dx = K.dx[tap];
                    mbar_wait(&S.empty[stage], phase ^ 1);
                    if (leader) mbar_arrive_expect_tx(&S.full[stage], 2 * stage_bytes);
                    uint8_t* dst = smem + stage * stage_bytes;
                    tma_load_4d_2cta(dst, &tmA, &S.full[stage], (int32_t)(kc * 64), x0 + dx, y0 + dy, (int32_t)b);
                    const int32_t wrow = (int32_t)(b * P.wrows_per_sample + K.wtap[tap] * P.ncols + nt * P.NT + rank * (P.NT / 2));
                    tma_load_2d_2cta(dst + CV_A_BYTES, &tmB, &S.full[stage], (int32_t)(kc * 64), wrow);
                    if (++stage == CV_NSTG) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (leader CTA)
        if (lane == 0 && leader) {
            const uint32_t idesc = idesc_f16(256, P.NT, FMT_F16, FMT_F16, 0, 0);
            uint32_t stage = 0, phase = 0, local = 0;
            for (uint32_t u = u_begin; u < u_end; u++, local++) {
                const uint32_t acc = local & 1, acc_phase = (local >> 1) & 1;
                mbar_wait(&S.tmem_empty[acc], acc_phase ^ 1);          // both CTAs' epilogues have drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * 256;
                const uint32_t n_k = P.cls[(u / P.n_nt) % P.n_cls].n_taps * n_kc;
                for (uint32_t k = 0; k < n_k; k++) {
                    mbar_wait(&S.full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * stage_bytes), b_addr = a_addr + CV_A_BYTES;
                    for (uint32_t s = 0; s < 4; s++)
                        umma_f16_2cta(tmem_d, smem_desc_sw128(a_addr + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), idesc, (k | s) != 0);
                    umma_commit_2cta(&S.empty[stage], 3);
                    if (++stage == CV_NSTG) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(&S.tmem_full[acc], 3);
            }
        }
    } else {
        // ===================================================== epilogue warpgroups
        const uint32_t wg = (warp - 2) >> 2;                           // accumulator stage owned
        const uint32_t q = warp & 3;                                   // TMEM lane quarter this warp may access
        const uint32_t r = q * 32 + lane;                              // pixel of the tile = TMEM lane
        const float nw = (P.epi == EPI_ACT && P.noise && P.noise_w) ? __ldg(P.noise_w) : 0.f;
        uint32_t local = wg;
        for (uint32_t u = u_begin + wg; u < u_end; u += 2, local += 2) {
            uint32_t b, tile, cl, nt;
            decode(u, b, tile, cl, nt);
            const uint32_t gx = P.x_org + (tile % P.tiles_x) * P.bw + r % P.bw, gy = P.y_org + (tile / P.tiles_x) * P.bh + r / P.bw;
            const uint32_t px = P.sx * gx + (uint32_t)P.cls[cl].ox, py = P.sy * gy + (uint32_t)P.cls[cl].oy;       // output pixel
            const bool valid = tile < tiles && gx < P.x_end && gy < P.y_end && px < P.out_w && py < P.out_h;
            const uint64_t pix = ((uint64_t)b * P.out_h + py) * P.out_w + px;
            const float nz = (valid && nw != 0.f) ? nw * __ldg(P.noise + pix) : 0.f;
            mbar_wait(&S.tmem_full[wg], (local >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((q * 32) << 16) + wg * 256;
            if (P.epi == EPI_RGB) {
                uint32_t raw[16];
                tmem_ld16(taddr, raw);
                tmem_ld_wait16(raw);
                if (valid) {
                    float v[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) v[c] = __uint_as_float(raw[c]) + __ldg(P.bias + c);
                    if (P.skip) {
                        const uint32_t Hs = P.out_h / 2, Ws = P.out_w / 2;
                        const float k4[4] = {0.25f, 0.75f, 0.75f, 0.25f};      // [1,3,3,1] / 8 * 2 per axis
                        // U[2y, 2x] = skip[y, x]: of the 4 x 4 taps the two per axis with Y + p even contribute
#pragma unroll
                        for (int pi = 0; pi < 2; pi++) {
                            const int p = (int)(py & 1) + 2 * pi, rr = ((int)py + p - 2) >> 1;
                            if (rr < 0 || rr >= (int)Hs) continue;
#pragma unroll
                            for (int qi = 0; qi < 2; qi++) {
                                const int qq = (int)(px & 1) + 2 * qi, cc = ((int)px + qq - 2) >> 1;
                                if (cc < 0 || cc >= (int)Ws) continue;
                                const float* sp = P.skip + (((size_t)b * Hs + rr) * Ws + cc) * 3;
                                const float kw = k4[p] * k4[qq];
#pragma unroll
                                for (int c = 0; c < 3; c++) v[c] = fmaf(kw, __ldg(sp + c), v[c]);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        if (P.out_nhwc) P.out_nhwc[pix * 3 + c] = v[c];
                        if (P.out_nchw) P.out_nchw[(((size_t)b * 3 + c) * P.out_h + py) * P.out_w + px] = v[c];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (leader) mbar_arrive(&S.tmem_empty[wg]); else mbar_arrive_remote(&S.tmem_empty[wg], 0); }
                continue;
            }
            uint16_t* orow = P.out + pix * P.ld_out + nt * P.NT;
            for (uint32_t c = 0; c < P.NT; c += 32) {
                uint32_t raw[32];
                tmem_ld32(taddr + c, raw);
                tmem_ld_wait();
                uint32_t h[16];
                if (P.epi == EPI_ACT) {
                    const float* bs = S.bias + nt * P.NT + c;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float v0 = __uint_as_float(raw[i]) + nz + bs[i], v1 = __uint_as_float(raw[i + 1]) + nz + bs[i + 1];
                        v0 = (v0 > 0.f ? v0 : 0.2f * v0) * 1.4142135623730951f;
                        v1 = (v1 > 0.f ? v1 : 0.2f * v1) * 1.4142135623730951f;
                        h[i / 2] = pack_f16_sat(v0, v1);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) h[i / 2] = pack_f16_sat(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
                }
                if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(orow + c);
#pragma unroll
                    for (int j = 0; j < 4; j++) dst[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (leader) mbar_arrive(&S.tmem_empty[wg]); else mbar_arrive_remote(&S.tmem_empty[wg], 0); }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2cta(tmem_base, 512);
}

}  // namespace tc
}  // namespace sdfg
