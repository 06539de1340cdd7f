// One layer of the field on the 5th-generation tensor cores: a persistent, warp-specialised tcgen05 GEMM
//     D[128 samples, N_out] = A[128 samples, K] * B[N_out, K]^T        (fp16 or bf16 operands, fp32 accumulation in TMEM)
// with the layer's elementwise work fused into the TMEM -> register epilogue.  Three epilogues share the pipeline:
//   MODE_F  forward:    h = sin(gamma*(acc + bias) + beta) (or acc + bias), optional sdf / rgb head dot products
//   MODE_R  backward 1: recompute z from the saved input, dz = dh * cos(z)           (dh from memory and / or a rank-r term)
//           gradients are fp16 carrying a power-of-two loss scale s (P.gscale): fp32 sources are multiplied by s on the way
//           in, fp32 results by 1/s on the way out, 16-bit stores saturate
//   MODE_D  backward 2: dh_in = dz * (gamma o W)  (+ rank-1 term d_sdf * w_sigma)    B = per-image (gamma o W)^T
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer: B (the layer's weight matrix, <= 160 KB) is loaded ONCE per CTA (per image in MODE_D) and stays
//               resident in shared memory; A tiles stream through a 3-stage ring of 128x64 bf16 boxes (128B swizzle)
//   warp 1      TMEM allocator + MMA issuer: one thread issues tcgen05.mma (M=128, N=N_out, K=16) per 32-byte K slice,
//               tcgen05.commit releases ring stages and publishes finished accumulators
//   warps 2-5 / 6-9   two epilogue warpgroups, each owning one of the two 256-column TMEM accumulator stages, so the epilogue
//               of tile i overlaps the MMAs of tile i+1.  A thread owns one sample row (TMEM lane) and walks the columns in
//               chunks of 32 (tcgen05.ld 32x32b.x32), so head dot products need no cross-thread reduction.
#pragma once
#include "tc_common.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t TILE_M = 128;
constexpr uint32_t KCH = 64;                       // bf16 elements per 128-byte swizzled row
constexpr uint32_t A_STAGE_BYTES = TILE_M * 128;   // 16 KB
constexpr uint32_t NSTAGES = 3;
constexpr uint32_t MAX_KCH = 5;                    // K <= 320
constexpr uint32_t ACC_COLS = 256;                 // TMEM columns per accumulator stage
constexpr uint32_t LAYER_THREADS = 320;

enum Mode { MODE_F = 0, MODE_R = 1, MODE_D = 2 };

struct LayerParams {
    uint32_t M_total, K, N_out, rows_per_image, n_tiles, tiles_per_cta;
    uint32_t ab_fmt, out_fmt;       // operand / 16-bit output element type: tc::FMT_F16 (activations, weights) or tc::FMT_BF16 (gradients)
    uint32_t b_rows_per_image;      // MODE_D: image b's B matrix starts at row b * b_rows_per_image of the B tensor map (0 = shared B)
    int act;                        // MODE_F: 1 = FiLM + sin, 0 = linear
    const float* bias;              // [N_out]
    const float* gamma;             // + img * gstride + n   (already offset to the layer)
    const float* beta;
    int64_t gstride;
    uint16_t* out16;                // F: h (fp16), R: dz (bf16), D: dh_in (bf16)     [M_total, ld_out]   (NULL ok)
    int64_t ld_out;
    uint16_t* out16b;               // F: second copy of h in bf16 (operand of the weight-gradient contraction)   (NULL ok)
    int64_t ld_out_b;
    float* out_f32;                 // F: fp32 copy of h, D: fp32 result             [M_total, ld_out_f32]   (NULL ok)
    int64_t ld_out_f32;
    int nh;                         // MODE_F heads: out_head[row*nh + c] = sum_n h[row,n] * head_w[c*N_out + n] + head_b[c]
    const float* head_w;
    const float* head_b;
    float* out_head;
    const float* gscale;            // MODE_R / MODE_D: device pointer to {s, 1/s}: the power-of-two loss scale every fp16 gradient carries (NULL = 1)
    const uint16_t* dh16;           // MODE_R: upstream gradient sources (each optional); dh16 is fp16 and already carries the scale
    int64_t ld_dh;
    const float* dh_f32;
    int64_t ld_dh_f32;
    int rank;                       // MODE_R / MODE_D: + sum_r rank_s[row*rank + r] * rank_v[r*N_out + n]
    const float* rank_s;
    const float* rank_v;
};

struct LayerSmem {                  // small, fixed part (after the big operand buffers)
    uint64_t b_full, b_free;
    uint64_t a_full[NSTAGES], a_empty[NSTAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    alignas(16) float gam[2][256];              // per epilogue warpgroup: gamma and gamma*bias + beta of the current image
    float cst[2][256];
    float vecs[3][256];             // head weights (MODE_F) or rank vectors (MODE_R / MODE_D)
};

__host__ __device__ inline uint32_t layer_smem_bytes(uint32_t K, uint32_t N_out) {
    const uint32_t nk = (K + KCH - 1) / KCH;
    return 1024 /* alignment slack */ + nk * N_out * 128 + NSTAGES * A_STAGE_BYTES + (uint32_t)sizeof(LayerSmem);
}

template <int MODE>
__global__ void __launch_bounds__(LAYER_THREADS, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ LayerParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t nk = (P.K + KCH - 1) / KCH;
    const uint32_t b_chunk_bytes = P.N_out * 128;
    uint8_t* smB = smem;
    uint8_t* smA = smem + nk * b_chunk_bytes;
    LayerSmem& S = *reinterpret_cast<LayerSmem*>(smA + NSTAGES * A_STAGE_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t_begin = blockIdx.x * P.tiles_per_cta;
    const uint32_t t_end = min(P.n_tiles, t_begin + P.tiles_per_cta);

    if (threadIdx.x == 0) {
        mbar_init(&S.b_full, 1);
        mbar_init(&S.b_free, 1);
        for (uint32_t i = 0; i < NSTAGES; i++) { mbar_init(&S.a_full[i], 1); mbar_init(&S.a_empty[i], 1); }
        for (uint32_t i = 0; i < 2; i++) { mbar_init(&S.tmem_full[i], 1); mbar_init(&S.tmem_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    if (warp == 1) tmem_alloc(&S.tmem_base, 512);
    // per-column vectors shared by both epilogue warpgroups
    {
        const int nvec = MODE == MODE_F ? P.nh : P.rank;
        const float* src = MODE == MODE_F ? P.head_w : P.rank_v;
        for (uint32_t i = threadIdx.x; i < (uint32_t)nvec * P.N_out; i += blockDim.x) S.vecs[i / P.N_out][i % P.N_out] = __ldg(src + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0 && t_begin < t_end) {
            uint32_t stage = 0, phase = 0, bgen = 0;
            int cur_img = -1;
            for (uint32_t t = t_begin; t < t_end; t++) {
                const int img = P.b_rows_per_image ? (int)(((uint64_t)t * TILE_M) / P.rows_per_image) : 0;
                if (img != cur_img) {
                    mbar_wait(&S.b_free, (bgen & 1) ^ 1);          // every MMA that read the previous B has completed
                    mbar_arrive_expect_tx(&S.b_full, nk * b_chunk_bytes);
                    for (uint32_t kc = 0; kc < nk; kc++)
                        tma_load_2d(smB + kc * b_chunk_bytes, &tmB, &S.b_full, (int32_t)(kc * KCH), (int32_t)(img * P.b_rows_per_image));
                    bgen++;
                    cur_img = img;
                }
                for (uint32_t kc = 0; kc < nk; kc++) {
                    mbar_wait(&S.a_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&S.a_full[stage], A_STAGE_BYTES);
                    tma_load_2d(smA + stage * A_STAGE_BYTES, &tmA, &S.a_full[stage], (int32_t)(kc * KCH), (int32_t)(t * TILE_M));
                    if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0 && t_begin < t_end) {
            const uint32_t idesc = idesc_f16(TILE_M, P.N_out, P.ab_fmt, P.ab_fmt, 0, 0);
            uint32_t stage = 0, phase = 0, bgen = 0, local = 0;
            int cur_img = -1;
            for (uint32_t t = t_begin; t < t_end; t++, local++) {
                const int img = P.b_rows_per_image ? (int)(((uint64_t)t * TILE_M) / P.rows_per_image) : 0;
                if (img != cur_img) {
                    mbar_wait(&S.b_full, bgen & 1);
                    bgen++;
                    cur_img = img;
                }
                const uint32_t acc = local & 1, acc_phase = (local >> 1) & 1;
                mbar_wait(&S.tmem_empty[acc], acc_phase ^ 1);      // the epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
                for (uint32_t kc = 0; kc < nk; kc++) {
                    mbar_wait(&S.a_full[stage], phase);
                    tc_fence_after();
                    const uint32_t ksteps = min(4u, (P.K - kc * KCH + 15) / 16);
                    const uint32_t a_addr = smem_u32(smA + stage * A_STAGE_BYTES);
                    const uint32_t b_addr = smem_u32(smB + kc * b_chunk_bytes);
                    for (uint32_t s = 0; s < ksteps; s++)
                        umma_bf16(tmem_d, smem_desc_sw128(a_addr + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), idesc,
                                  (kc | s) != 0);
                    umma_commit(&S.a_empty[stage]);                // ring stage reusable once these MMAs have read it
                    if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&S.tmem_full[acc]);                    // accumulator complete -> epilogue
                if (P.b_rows_per_image && t + 1 < t_end && (int)(((uint64_t)(t + 1) * TILE_M) / P.rows_per_image) != img)
                    umma_commit(&S.b_free);                        // next tile belongs to another image: B may be overwritten
            }
        }
    } else {
        // ===================================================== epilogue warpgroups
        const uint32_t wg = (warp - 2) >> 2;                       // 0 / 1 = accumulator stage owned
        const uint32_t q = warp & 3;                               // TMEM lane quarter this warp may access
        const uint32_t wg_tid = ((warp - 2) & 3) * 32 + lane;
        const uint32_t row_in_tile = q * 32 + lane;
        float* gam = S.gam[wg];
        float* cst = S.cst[wg];
        int my_img = -1;
        uint32_t local = wg;
        for (uint32_t t = t_begin + wg; t < t_end; t += 2, local += 2) {
            const uint32_t acc_phase = (local >> 1) & 1;
            if (MODE != MODE_D) {
                const int img = (int)(((uint64_t)t * TILE_M) / P.rows_per_image);
                if (img != my_img) {                               // refresh the per-image FiLM constants of this warpgroup
                    named_bar_sync(1 + wg, 128);
                    for (uint32_t n = wg_tid; n < P.N_out; n += 128) {
                        const float b = __ldg(P.bias + n);
                        if (MODE == MODE_F && !P.act) { gam[n] = 1.f; cst[n] = b; }
                        else {
                            const float g = __ldg(P.gamma + (int64_t)img * P.gstride + n);
                            gam[n] = g;
                            cst[n] = fmaf(g, b, __ldg(P.beta + (int64_t)img * P.gstride + n));
                        }
                    }
                    named_bar_sync(1 + wg, 128);
                    my_img = img;
                }
            }
            const uint64_t row = (uint64_t)t * TILE_M + row_in_tile;
            const bool valid = row < P.M_total;
            float rs[3] = {0.f, 0.f, 0.f};
            const float gs = (MODE != MODE_F && P.gscale) ? __ldg(P.gscale) : 1.f;
            const float gs_inv = (MODE != MODE_F && P.gscale) ? __ldg(P.gscale + 1) : 1.f;
            if (MODE != MODE_F && valid)
                for (int r = 0; r < P.rank; r++) rs[r] = gs * __ldg(P.rank_s + row * P.rank + r);
            float hacc[3] = {0.f, 0.f, 0.f};

            mbar_wait(&S.tmem_full[wg], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((q * 32) << 16) + wg * ACC_COLS;
            for (uint32_t c = 0; c < P.N_out; c += 32) {
                uint32_t raw[32];
                tmem_ld32(taddr + c, raw);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; i++) v[i] = __uint_as_float(raw[i]);
                if (MODE == MODE_F) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 g4 = *reinterpret_cast<const float4*>(gam + c + i);
                        const float4 c4 = *reinterpret_cast<const float4*>(cst + c + i);
                        v[i] = fmaf(v[i], g4.x, c4.x); v[i + 1] = fmaf(v[i + 1], g4.y, c4.y);
                        v[i + 2] = fmaf(v[i + 2], g4.z, c4.z); v[i + 3] = fmaf(v[i + 3], g4.w, c4.w);
                    }
                    if (P.act) {
#pragma unroll
                        for (int i = 0; i < 32; i++) v[i] = __sinf(v[i]);
                    }
                    for (int hd = 0; hd < P.nh; hd++) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(&S.vecs[hd][c + i]);
                            hacc[hd] = fmaf(v[i], w4.x, hacc[hd]); hacc[hd] = fmaf(v[i + 1], w4.y, hacc[hd]);
                            hacc[hd] = fmaf(v[i + 2], w4.z, hacc[hd]); hacc[hd] = fmaf(v[i + 3], w4.w, hacc[hd]);
                        }
                    }
                } else if (MODE == MODE_R) {
                    float dh[32];
#pragma unroll
                    for (int i = 0; i < 32; i++) dh[i] = 0.f;
                    if (P.dh16 && valid) {
                        const uint4* src = reinterpret_cast<const uint4*>(P.dh16 + row * P.ld_dh + c);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint4 u = __ldg(src + j);
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const float2 f = unpack_f16(w[k]);
                                dh[j * 8 + 2 * k] = f.x;
                                dh[j * 8 + 2 * k + 1] = f.y;
                            }
                        }
                    }
                    if (P.dh_f32 && valid) {
                        const float4* src = reinterpret_cast<const float4*>(P.dh_f32 + row * P.ld_dh_f32 + c);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float4 u = ldg_stream4(src + j);
                            dh[j * 4] = fmaf(gs, u.x, dh[j * 4]); dh[j * 4 + 1] = fmaf(gs, u.y, dh[j * 4 + 1]);
                            dh[j * 4 + 2] = fmaf(gs, u.z, dh[j * 4 + 2]); dh[j * 4 + 3] = fmaf(gs, u.w, dh[j * 4 + 3]);
                        }
                    }
                    for (int r = 0; r < P.rank; r++) {
#pragma unroll
                        for (int i = 0; i < 32; i++) dh[i] = fmaf(rs[r], S.vecs[r][c + i], dh[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 g4 = *reinterpret_cast<const float4*>(gam + c + i);
                        const float4 c4 = *reinterpret_cast<const float4*>(cst + c + i);
                        v[i] = dh[i] * __cosf(fmaf(v[i], g4.x, c4.x));
                        v[i + 1] = dh[i + 1] * __cosf(fmaf(v[i + 1], g4.y, c4.y));
                        v[i + 2] = dh[i + 2] * __cosf(fmaf(v[i + 2], g4.z, c4.z));
                        v[i + 3] = dh[i + 3] * __cosf(fmaf(v[i + 3], g4.w, c4.w));
                    }
                } else {   // MODE_D
                    for (int r = 0; r < P.rank; r++) {
#pragma unroll
                        for (int i = 0; i < 32; i++) v[i] = fmaf(rs[r], S.vecs[r][c + i], v[i]);
                    }
                }
                if (valid) {
                    if (P.out16) {
                        uint4* dst = reinterpret_cast<uint4*>(P.out16 + row * P.ld_out + c);
                        const uint32_t f = P.out_fmt;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (MODE == MODE_F)
                                dst[j] = make_uint4(pack16(v[j * 8], v[j * 8 + 1], f), pack16(v[j * 8 + 2], v[j * 8 + 3], f),
                                                    pack16(v[j * 8 + 4], v[j * 8 + 5], f), pack16(v[j * 8 + 6], v[j * 8 + 7], f));
                            else
                                dst[j] = make_uint4(pack_f16_sat(v[j * 8], v[j * 8 + 1]), pack_f16_sat(v[j * 8 + 2], v[j * 8 + 3]),
                                                    pack_f16_sat(v[j * 8 + 4], v[j * 8 + 5]), pack_f16_sat(v[j * 8 + 6], v[j * 8 + 7]));
                        }
                    }
                    if (MODE == MODE_F && P.out16b) {
                        uint4* dst = reinterpret_cast<uint4*>(P.out16b + row * P.ld_out_b + c);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            dst[j] = make_uint4(pack_bf16(v[j * 8], v[j * 8 + 1]), pack_bf16(v[j * 8 + 2], v[j * 8 + 3]),
                                                pack_bf16(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16(v[j * 8 + 6], v[j * 8 + 7]));
                    }
                    if (P.out_f32) {
                        float4* dst = reinterpret_cast<float4*>(P.out_f32 + row * P.ld_out_f32 + c);
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            dst[j] = make_float4(gs_inv * v[j * 4], gs_inv * v[j * 4 + 1], gs_inv * v[j * 4 + 2], gs_inv * v[j * 4 + 3]);
                    }
                }
            }
            // all TMEM reads of this tile are complete (wait::ld above): hand the accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.tmem_empty[wg]);
            if (MODE == MODE_F && P.nh && valid) {
                for (int hd = 0; hd < P.nh; hd++) P.out_head[row * P.nh + hd] = hacc[hd] + __ldg(P.head_b + hd);
            }
        }
    }
    // teardown: everything issued has been consumed (the epilogue waited on the last tmem_full)
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace sdfg
