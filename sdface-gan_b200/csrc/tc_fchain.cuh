// The whole field forward as ONE persistent tcgen05 kernel on CTA pairs, with the FiLM modulation folded into the weights.
//
// Per image b and layer l the host-side prep kernel (fold_weights_kernel, field_tc.cu) writes fp16 matrices
//     WF[b][l] = [ gamma_b o W_l | small chunk ]          gamma_b o W_l: K = 256 main columns (absent for the first layer)
// whose last 64 columns ("small chunk") mirror the column layout of the per-tile SMALL operand tile:
//     SMALL (A side)   [ x part (hash features / raw points) | view part (SH of the ray direction) | 1 1 0 ... ]
//     small chunk (B)  [ gamma o W_l[:, x cols]              | gamma o W_l[:, view cols]            | c_hi c_lo 0 ... ]
// with c = gamma_b o bias_l + beta_b split into two fp16 values (c_hi + c_lo carries 22 mantissa bits).  One K = 16 step against
// the constant-one columns therefore adds the whole FiLM offset inside the tensor core, and the accumulator a layer's epilogue
// reads IS the sine's argument u = gamma (W h + b) + beta  (ref FiLMSiren.forward sdf_model.py:61-69; input_linear :38-41 is the
// same with gamma = 1, beta = 0).  What is gone compared with an epilogue-side FiLM: the per-layer table of 512 constants, its
// 512-thread barrier, two LDS.128 and four FFMA per four elements -- the epilogue was issue-bound on exactly those.
//
//   loader warp   x_in / view_feat (fp32, HBM) -> fp16 -> SMALL (128B-swizzled K-major operand tile, written by hand)
//   TMA producer  streams this CTA's half of every [256 x 64] weight chunk of the tile's image through a ring (L2 hits: 0.6 MB per image)
//   MMA thread    (leader CTA of the pair) tcgen05.mma.cta_group::2, M = 256: layer i accumulates into TMEM accumulator (i & 1);
//                 the small chunk goes first, the K = 256 part follows 64-column chunk by chunk as the epilogue of layer i-1
//                 produces it (act_ready[kc]) -> MMAs of layer i overlap the epilogue of layer i-1
//   epilogue      16 warps (4 per TMEM lane quarter, 16 columns of every 64-column chunk each): tcgen05.ld (prefetched one chunk
//                 ahead) -> [sign(cos u) bit masks for the backward] -> sin -> head dot products -> fp16 -> st.shared into ACT in place
//   storer        (SAVE) TMA-stores finished ACT chunks (saved activations / fp16 features)
// The first FiLM layer (K = in_dim <= 32, |W| ~ 1/3, gamma ~ 30) is where 16-bit operand rounding hurts: with split_x its x part runs
// as hi + lo pairs -- x = x_hi + x_lo (a second small operand tile), W = W_hi + W_lo (a second weight chunk), three products
// W_hi x_hi + W_hi x_lo + W_lo x_hi -- i.e. ~fp32 operands for 4 extra K-steps per tile (tests/test_operand_precision.py).
// The sine: MUFU.SIN runs at 16 / clk / SM -- 2048 clk per layer and SM, exactly the layer's MMA time -- so some of the 8
// element pairs of every piece (POLY_PAIRS below) can take the FMA pipe instead: r = u - k pi with k = rint(u / pi) (the same fma against 1.5 * 2^23 that
// yields the sign bit), an odd degree-7 minimax polynomial on [-pi/2, pi/2] (max error 1e-6, the level of sin.approx) in packed
// fma.rn.f32x2, and the sign (-1)^k xor-ed in.
// CTA pairs (CG = 2): each CTA stages its own 128 rows of A and HALF of every weight chunk, halving the L2 -> SM weight stream
// (0.6 MB per tile: at 2 ms per pass a single-CTA chain would need the chip's whole L2 bandwidth for it).
#pragma once
#include <type_traits>

#include "tc_chain.cuh"

namespace sdfg {
namespace tc {

// How many of the 8 element pairs per thread and piece take the polynomial.  Measured (B = 32, scripts/gpu_poly.sh), pairs 0 / 1 / 2 / 3:
// inference chain 1.59 / 1.55 / 1.55 / 1.57 ms (thumbnail), 1.60 / 1.61 / 1.63 / 1.67 ms (with features); training forward 2.23 / 2.28 /
// 2.33 / 2.36 ms -- there the epilogue is issue-bound on recording the derivative planes and MUFU.SIN costs fewer issue slots than the
// polynomial (the range reduction fma is needed for the sign bit either way).
#ifndef SDFG_POLY_PAIRS
#define SDFG_POLY_PAIRS 1
#endif
#ifndef SDFG_POLY_PAIRS_TRAIN
#define SDFG_POLY_PAIRS_TRAIN 0
#endif

constexpr uint32_t FC_MAX_LAYERS = SDFG_MAX_FILM + 1;
__host__ __device__ constexpr uint32_t fc_w_bytes(int cg) { return 32768u / (uint32_t)cg; }
__host__ __device__ constexpr uint32_t fc_w_stages(int cg) { return cg == 2 ? 6u : 3u; }
constexpr uint32_t FC_MAX_W_STAGES = 7;

// layer kinds: the epilogue body is compiled once per kind with the layer's properties as constants (FK_GENERIC reads them at run time)
enum FKind : int {
    FK_GENERIC = 0,
    FK_FILM = 1,        // sin, no head, output feeds the next layer
    FK_FILM_SDF = 2,    // sin + one head row (sdf), output feeds the next layer
    FK_VIEWS_FIN = 3,   // sin + three head rows (rgb), last layer, output kept in ACT for the TMA store (features / saved activation)
    FK_SDF_FIN = 4      // sin + one head row, last layer, output not kept (sdf-only query)
};

struct FLayer {
    uint32_t kind;              // FKind
    uint32_t n_main;            // 4: K = 256 part (A = ACT, chunks 0..3 of the layer's matrix); 0: small chunk only
    uint32_t small_mask;        // K-steps (of 16 columns) of SMALL / of the small chunk this layer multiplies
    uint32_t use_x, use_v;      // the small chunk reads the x / view part (loader hand-shake)
    uint32_t split_x;           // x part as hi + lo pairs: a second weight chunk (W_lo) follows the small chunk, SMALL2 holds x_lo
    uint32_t act;               // 1: sin, 0: linear
    uint32_t to_act;            // write the fp16 output into ACT (input of the next layer and / or source of the TMA store)
    uint32_t store;             // SAVE: TMA-store the output with tensor map st[layer]
    uint32_t nh;                // head rows: out_head[row*nh + c] = sum_n h[row,n] * head_w[c*256 + n] + head_b[c]
    const float* head_w;
    const float* head_b;
    float* out_head;
    float* out_f32;             // optional fp32 copy of the output in HBM
    int64_t ld_out_f32;
    uint8_t* sgn;               // COS: sign(cos u) bit masks, [tiles][CH_SGN_TILE_BYTES] (NULL = not wanted)
};

struct FChainParams {
    uint32_t M_total, rows_per_image, rows_per_ray, n_units, units_per_cta, n_layers;   // unit = CG adjacent 128-row tiles
    uint32_t in_dim, view_dim, x_nk, v_nk;     // v_nk = 0: no view part
    uint32_t split_x, pad0;                    // the loader also fills SMALL2 with x_lo = fp16(x - fp16(x))
    const float* x_in;          // [M, in_dim] fp32
    const float* view_feat;     // [M / rows_per_ray, view_dim] fp32
    uint16_t* x16;              // SAVE: fp16 copy of x, [M, kp_x] zero padded (NULL ok)
    uint32_t kp_x, kp_v;
    uint16_t* v16;              // SAVE: view part expanded per sample, kp_v columns (NULL ok)
    int64_t ld_v16;
    FLayer layer[FC_MAX_LAYERS];
};

struct alignas(64) FChainMaps {
    CUtensorMap w[FC_MAX_LAYERS];      // WF of layer l: [B*256, n_main*64 + 64] fp16, box (256 / CG) x 64
    CUtensorMap st[FC_MAX_LAYERS];     // output store of layer l: [M, 256] fp16, box 128 x 64
};

struct FChainSmem {
    uint64_t w_full[FC_MAX_W_STAGES], w_empty[FC_MAX_W_STAGES];
    uint64_t act_ready[4];             // MMA side (on the leader CTA: every epilogue warp of the pair arrives)
    uint64_t act_ready_st[4], fin_ready[4], st_done[4];   // storer side (local)
    uint64_t acc_full[2], acc_empty[2];
    uint64_t x_full, x_free, v_full, v_free;
    uint32_t tmem_base;
    uint32_t pad[3];
    alignas(16) float heads[4][256];   // row 0: first head layer (sdf), rows 1..3: second head layer (rgb)
    float hbias[4];
    float hx[3][CH_TILE_M][3];         // head partial sums of column sub-blocks 1..3
};

__host__ __device__ inline uint32_t fchain_smem_bytes(int cg) {
    return 1024 + CH_ACT_BYTES + 2 * CH_CHUNK_BYTES + fc_w_stages(cg) * fc_w_bytes(cg) + 2 * CH_SGN_TILE_BYTES + (uint32_t)sizeof(FChainSmem);
}

template <bool SAVE, bool COS, int CG>
__global__ void __launch_bounds__(CH_THREADS, 1)
tc_fchain_fwd_kernel(const __grid_constant__ FChainMaps maps, const __grid_constant__ FChainParams P) {
    constexpr bool PAIR = CG == 2;
    constexpr int POLY_PAIRS = COS ? SDFG_POLY_PAIRS_TRAIN : SDFG_POLY_PAIRS;      // sine on the FMA pipe for that many of the 8 pairs
    constexpr uint32_t W_BYTES = fc_w_bytes(CG), NW = fc_w_stages(CG), W_ROWS = 256 / CG;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smACT = smem;
    uint8_t* smSMALL = smACT + CH_ACT_BYTES;
    uint8_t* smSMALL2 = smSMALL + CH_CHUNK_BYTES;                       // x_lo (split_x)
    uint8_t* smRING = smSMALL2 + CH_CHUNK_BYTES;
    uint8_t* smSGN = smRING + NW * W_BYTES;                             // two sign-mask tiles (double-buffered by layer)
    FChainSmem& S = *reinterpret_cast<FChainSmem*>(smSGN + 2 * CH_SGN_TILE_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const uint32_t u_begin = (blockIdx.x / CG) * P.units_per_cta;
    const uint32_t u_end = min(P.n_units, u_begin + P.units_per_cta);
    const uint32_t nL = P.n_layers;
    const uint32_t ones_step = P.x_nk + P.v_nk;                        // K-step of SMALL holding the constant-one columns

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < NW; i++) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (uint32_t i = 0; i < 4; i++) {
            mbar_init(&S.act_ready[i], CH_EPI_WARPS * CG); mbar_init(&S.act_ready_st[i], CH_EPI_WARPS);
            mbar_init(&S.fin_ready[i], CH_EPI_WARPS); mbar_init(&S.st_done[i], 1);
        }
        for (uint32_t i = 0; i < 2; i++) { mbar_init(&S.acc_full[i], 1); mbar_init(&S.acc_empty[i], CH_EPI_WARPS * CG); }
        mbar_init(&S.x_full, CG); mbar_init(&S.x_free, 1); mbar_init(&S.v_full, CG); mbar_init(&S.v_free, 1);
        fence_barrier_init();
    }
    if (warp == CH_WARP_TMA && lane == 0)
        for (uint32_t i = 0; i < nL; i++) {
            tma_prefetch_desc(&maps.w[i]);
            if (SAVE && P.layer[i].store) tma_prefetch_desc(&maps.st[i]);
        }
    if (warp == CH_WARP_MMA) { if (PAIR) tmem_alloc_2cta(&S.tmem_base, 512); else tmem_alloc(&S.tmem_base, 512); }
    {
        // head vectors
        uint32_t hrow = 0;
        for (uint32_t i = 0; i < nL; i++) {
            const uint32_t nh = P.layer[i].nh;
            for (uint32_t k = threadIdx.x; k < nh * 256 && hrow + nh <= 4; k += blockDim.x) S.heads[hrow + k / 256][k % 256] = __ldg(P.layer[i].head_w + k);
            if (threadIdx.x < nh && hrow + nh <= 4) S.hbias[hrow + threadIdx.x] = __ldg(P.layer[i].head_b + threadIdx.x);
            hrow += nh;
        }
        // SMALL: zero once (padding columns are never written again), then the two constant-one columns of every row
        for (uint32_t i = threadIdx.x; i < 2 * CH_CHUNK_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(smSMALL)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        for (uint32_t r = threadIdx.x; r < CH_TILE_M; r += blockDim.x)
            *reinterpret_cast<uint32_t*>(smSMALL + sw128(r, 2 * ones_step)) = 0x3C003C00u;      // fp16 {1, 1}
        fence_proxy_async();
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;

    if (warp == CH_WARP_TMA) {
        // ===================================================== weight producer (both CTAs): own half of every chunk, in MMA issue order
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint64_t keep = l2_policy_evict_last();              // re-read by every tile of the image; activations stream past
            for (uint32_t u = u_begin; u < u_end; u++) {
                const int32_t img = (int32_t)(((u * CG + rank) * CH_TILE_M) / P.rows_per_image);
                const int32_t row = img * 256 + (int32_t)(rank * W_ROWS);
                for (uint32_t i = 0; i < nL; i++) {
                    const uint32_t n_main = P.layer[i].n_main, n_small = 1 + P.layer[i].split_x;
                    for (uint32_t k = 0; k < n_small + n_main; k++) {   // the small chunk(s) (column blocks n_main, n_main + 1) first, then the main chunks
                        const int32_t c0 = (int32_t)((k < n_small ? n_main + k : k - n_small) * 64);
                        mbar_wait(&S.w_empty[stage], phase ^ 1);
                        if (leader) mbar_arrive_expect_tx(&S.w_full[stage], CG * W_BYTES);
                        if (PAIR) tma_load_2d_2cta_hint(smRING + stage * W_BYTES, &maps.w[i], &S.w_full[stage], c0, row, keep);
                        else tma_load_2d_hint(smRING + stage * W_BYTES, &maps.w[i], &S.w_full[stage], c0, row, keep);
                        if (++stage == NW) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == CH_WARP_MMA) {
        // ===================================================== MMA issuer (leader CTA only)
        if (lane == 0 && leader) {
            const uint32_t idesc = idesc_f16(CH_TILE_M * CG, 256, FMT_F16, FMT_F16, 0, 0);
            const uint32_t a_small = smem_u32(smSMALL), a_act = smem_u32(smACT);
            uint32_t stage = 0, phase = 0, n = 0, actgen = 0, it = 0;
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t accum) {
                if (PAIR) umma_f16_2cta(d, da, db, idesc, accum); else umma_bf16(d, da, db, idesc, accum);
            };
            auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_2cta(bar, 3); else umma_commit(bar); };
            for (uint32_t u = u_begin; u < u_end; u++, it++)
                for (uint32_t i = 0; i < nL; i++, n++) {
                    const uint32_t n_main = P.layer[i].n_main, mask = P.layer[i].small_mask;
                    const uint32_t acc = n & 1, use = n >> 1;
                    mbar_wait(&S.acc_empty[acc], (use & 1) ^ 1);          // the epilogues have drained this accumulator
                    if (P.layer[i].use_x) mbar_wait(&S.x_full, it & 1);
                    if (P.layer[i].use_v) mbar_wait(&S.v_full, it & 1);
                    mbar_wait(&S.w_full[stage], phase);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * 256;
                    uint32_t accumulate = 0;
                    {
                        const uint32_t b_addr = smem_u32(smRING + stage * W_BYTES);
                        for (uint32_t s = 0; s < 4; s++)
                            if (mask & (1u << s)) {
                                mma(tmem_d, smem_desc_sw128(a_small + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), accumulate);
                                accumulate = 1;
                            }
                        if (P.layer[i].split_x) {                        // + W_hi x_lo (same chunk), then + W_lo x_hi (next chunk)
                            const uint32_t a_lo = smem_u32(smSMALL2);
                            for (uint32_t s = 0; s < P.x_nk; s++)
                                mma(tmem_d, smem_desc_sw128(a_lo + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), 1);
                            commit(&S.w_empty[stage]);
                            if (++stage == NW) { stage = 0; phase ^= 1; }
                            mbar_wait(&S.w_full[stage], phase);
                            tc_fence_after();
                            const uint32_t b_lo = smem_u32(smRING + stage * W_BYTES);
                            for (uint32_t s = 0; s < P.x_nk; s++)
                                mma(tmem_d, smem_desc_sw128(a_small + s * 32, 16, 1024), smem_desc_sw128(b_lo + s * 32, 16, 1024), 1);
                        }
                        if (P.layer[i].use_x) commit(&S.x_free);
                        if (P.layer[i].use_v) commit(&S.v_free);
                        commit(&S.w_empty[stage]);
                        if (++stage == NW) { stage = 0; phase ^= 1; }
                    }
                    if (n_main) {
                        for (uint32_t kc = 0; kc < 4; kc++) {
                            mbar_wait(&S.act_ready[kc], actgen & 1);      // chunk kc of the previous layer's output is in ACT (both CTAs)
                            mbar_wait(&S.w_full[stage], phase);
                            tc_fence_after();
                            const uint32_t a_addr = a_act + kc * CH_CHUNK_BYTES;
                            const uint32_t b_addr = smem_u32(smRING + stage * W_BYTES);
                            for (uint32_t s = 0; s < 4; s++)
                                mma(tmem_d, smem_desc_sw128(a_addr + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), 1);
                            commit(&S.w_empty[stage]);
                            if (++stage == NW) { stage = 0; phase ^= 1; }
                        }
                        actgen++;
                    }
                    commit(&S.acc_full[acc]);
                }
        }
    } else if (warp == CH_WARP_LOAD) {
        // ===================================================== loader: x / view parts of the tile -> SMALL (+ fp16 copies in HBM)
        const uint32_t xu = 2 * P.x_nk, vu = 2 * P.v_nk;               // 16-byte units per row
        const bool x_fast = P.in_dim % 8 == 0 && (reinterpret_cast<uintptr_t>(P.x_in) & 15) == 0;
        const bool v_fast = vu && P.view_dim % 8 == 0 && (reinterpret_cast<uintptr_t>(P.view_feat) & 15) == 0;
        auto fill = [&](const float* src, uint32_t src_ld, uint32_t src_div, uint32_t n_valid, bool fast, uint32_t nu, uint32_t u_off,
                        uint32_t row0, uint16_t* copy, uint64_t copy_ld, uint32_t copy_cols, bool lo) {
            const uint32_t total = CH_TILE_M * nu;
            for (uint32_t base = 0; base < total; base += 128) {
                float v[4][8];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t i = min(base + j * 32 + lane, total - 1);
                    const uint32_t r = i / nu, u = i % nu;
                    const uint32_t row = min(row0 + r, P.M_total - 1);
                    const float* sp = src + (uint64_t)(row / src_div) * src_ld;
                    if (fast) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(sp + u * 8)), b = __ldg(reinterpret_cast<const float4*>(sp + u * 8) + 1);
                        v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w; v[j][4] = b.x; v[j][5] = b.y; v[j][6] = b.z; v[j][7] = b.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; k++) v[j][k] = (u * 8 + k < n_valid) ? __ldg(sp + u * 8 + k) : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t i = base + j * 32 + lane;
                    if (i < total) {
                        const uint32_t r = i / nu, u = i % nu;
                        const uint32_t row = row0 + r;
                        const uint4 h = pack8(v[j], FMT_F16);
                        *reinterpret_cast<uint4*>(smSMALL + sw128(r, u_off + u)) = h;
                        if (lo) {                                       // residual of the fp16 rounding, itself as fp16
                            const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
                            float d[8];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const float2 f = unpack_f16(hw[k]);
                                d[2 * k] = v[j][2 * k] - f.x;
                                d[2 * k + 1] = v[j][2 * k + 1] - f.y;
                            }
                            *reinterpret_cast<uint4*>(smSMALL2 + sw128(r, u_off + u)) = pack8(d, FMT_F16);
                        }
                        if (SAVE && copy && row < P.M_total && u * 8 < copy_cols) *reinterpret_cast<uint4*>(copy + (uint64_t)row * copy_ld + u * 8) = h;
                    }
                }
            }
        };
        auto arrive_mma = [&](uint64_t* bar) { if (PAIR && !leader) mbar_arrive_remote(bar, 0); else mbar_arrive(bar); };
        uint32_t it = 0;
        for (uint32_t u = u_begin; u < u_end; u++, it++) {
            const uint32_t row0 = (u * CG + rank) * CH_TILE_M;
            mbar_wait(&S.x_free, (it & 1) ^ 1);
            fill(P.x_in, P.in_dim, 1, P.in_dim, x_fast, xu, 0, row0, P.x16, P.kp_x, P.kp_x, P.split_x != 0);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) arrive_mma(&S.x_full);
            if (vu) {
                mbar_wait(&S.v_free, (it & 1) ^ 1);
                fill(P.view_feat, P.view_dim, P.rows_per_ray, P.view_dim, v_fast, vu, xu, row0, P.v16, (uint64_t)P.ld_v16, P.kp_v, false);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) arrive_mma(&S.v_full);
            }
        }
    } else if (warp == CH_WARP_STORE) {
        // ===================================================== storer (SAVE): finished ACT chunks -> saved activations in HBM
        if (SAVE && lane == 0) {
            uint32_t actgen = 0, fingen = 0, nn = 0;
            // One bulk group per chunk (the sign-mask tile of a layer rides with its last chunk); one group may still be reading
            // shared memory while the next chunk's store is issued -- a chunk is released for overwriting one iteration behind.
            uint64_t* pend = nullptr;
            const uint64_t stream = l2_policy_evict_first();
            for (uint32_t u = u_begin; u < u_end; u++) {
                const uint32_t t = u * CG + rank;
                for (uint32_t i = 0; i < nL; i++, nn++) {
                    const bool st = P.layer[i].store != 0, fin = i + 1 == nL;
                    if (!P.layer[i].to_act) continue;
                    for (uint32_t c = 0; c < 4; c++) {
                        if (fin) mbar_wait(&S.fin_ready[c], fingen & 1);
                        else mbar_wait(&S.act_ready_st[c], actgen & 1);
                        if (st) tma_store_2d_hint(&maps.st[i], smACT + c * CH_CHUNK_BYTES, (int32_t)(c * 64), (int32_t)(t * CH_TILE_M), stream);
                        if (COS && c == 3 && P.layer[i].sgn)              // every warp has written its bits of all 4 chunks
                            bulk_store(P.layer[i].sgn + (size_t)t * CH_SGN_TILE_BYTES, smSGN + (nn & 1) * CH_SGN_TILE_BYTES, CH_SGN_TILE_BYTES);
                        tma_store_commit();
                        if (pend) { tma_store_wait_read_pending<1>(); mbar_arrive(pend); }
                        pend = &S.st_done[c];
                    }
                    if (fin) fingen++; else actgen++;
                }
            }
            tma_store_wait_read();
            if (pend) mbar_arrive(pend);
            tma_store_wait_all();
        }
    } else {
        // ===================================================== epilogue: 16 warps, 4 per TMEM lane quarter
        const uint32_t q = warp & 3;                                   // TMEM lane quarter this warp may access
        const uint32_t sb = warp >> 2;                                 // 16-column sub-block of every 64-column chunk
        const uint32_t r = q * 32 + lane;                              // row of the tile = TMEM lane
        const uint32_t act_row = smem_u32(smACT) + r * 128;
        const uint32_t u0 = ((2 * sb) ^ (r & 7)) << 4, u1 = ((2 * sb + 1) ^ (r & 7)) << 4;
        auto arrive_mma = [&](uint64_t* bar) { if (PAIR && !leader) mbar_arrive_remote(bar, 0); else mbar_arrive(bar); };
        const uint64_t inv_pi2 = pk2(0.31830988618379067f, 0.31830988618379067f), magic2 = pk2(12582912.f, 12582912.f);
        const uint64_t nmagic2 = pk2(-12582912.f, -12582912.f), npi2 = pk2(-3.14159265358979f, -3.14159265358979f);
        const uint64_t c7 = pk2(-0.0001849218097049743f, -0.0001849218097049743f), c5 = pk2(0.008312365971505642f, 0.008312365971505642f);
        const uint64_t c3 = pk2(-0.16665680706501007f, -0.16665680706501007f);
        uint32_t n = 0, stgen = 0;
        for (uint32_t u = u_begin; u < u_end; u++) {
            const uint32_t t = u * CG + rank;
            const uint64_t row = (uint64_t)t * CH_TILE_M + r;
            const bool valid = row < P.M_total;
            uint32_t hrow = 0;
            for (uint32_t i = 0; i < nL; i++, n++) {
                // The layer body is instantiated per layer KIND (set by the host): what a layer does -- activation, head rows, final
                // layer, fp32 copy -- is then a compile-time constant inside the piece loop.  With runtime flags the loop carried ~25
                // branch / predicate / reconvergence instructions per 16-element piece and predicated-off stores (ncu source page,
                // profiles/r02a: 13.5 issued instructions per element against ~5 of arithmetic).
                auto body = [&](auto kind_c) {
                    constexpr int KIND = decltype(kind_c)::value;
                    constexpr bool GEN = KIND == FK_GENERIC;
                    const uint32_t L_act = GEN ? P.layer[i].act : 1u;
                    const uint32_t L_nh = GEN ? P.layer[i].nh : (KIND == FK_FILM ? 0u : (KIND == FK_FILM_SDF || KIND == FK_SDF_FIN) ? 1u : 3u);
                    const uint32_t L_to_act = GEN ? P.layer[i].to_act : (KIND == FK_SDF_FIN ? 0u : 1u);
                    const bool fin = GEN ? (i + 1 == nL) : (KIND == FK_VIEWS_FIN || KIND == FK_SDF_FIN);
                    float* const o32_row = (GEN && P.layer[i].out_f32 && valid) ? P.layer[i].out_f32 + row * P.layer[i].ld_out_f32 : nullptr;
                    const bool do_sgn = COS && (GEN ? P.layer[i].sgn != nullptr : true);
                    const uint32_t acc = n & 1, use = n >> 1;
                    const uint32_t heads_s = smem_u32(&S.heads[hrow][0]);
                    uint64_t hacc2[3] = {0ull, 0ull, 0ull};                 // packed (even, odd) partial head sums
                    mbar_wait(&S.acc_full[acc], use & 1);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((q * 32) << 16) + acc * 256 + sb * 16;
                    uint32_t raw[2][16];
                    tmem_ld16_issue(taddr, raw[0]);
#pragma unroll
                    for (uint32_t c = 0; c < 4; c++) {
                        const uint32_t col = c * 64 + sb * 16;
                        tmem_ld_wait16(raw[c & 1]);
                        if (c < 3) tmem_ld16_issue(taddr + (c + 1) * 64, raw[(c + 1) & 1]);
                        float v[16];
                        uint32_t sgn_m = 0;                                 // sign(cos) mask of this piece, in bits 16..31 once all 16 are in
#pragma unroll
                        for (int k = 0; k < 16; k++) v[k] = __uint_as_float(raw[c & 1][k]);
                        if (L_act) {
                            // t = u / pi + 1.5 * 2^23: the low mantissa bits hold k = rint(u / pi); parity(k) = sign of cos(u) = sign flip of the
                            // range-reduced sine.  Needed for all pairs when the masks are recorded, else for the polynomial pairs only.
                            uint64_t t2[8];
#pragma unroll
                            for (int j = 0; j < 8; j++)
                                if (do_sgn || j >= 8 - POLY_PAIRS) t2[j] = fma2(pk2(v[2 * j], v[2 * j + 1]), inv_pi2, magic2);
                            if (do_sgn) {
                                // A funnel shift per element moves the parity bit into the mask: even elements first, then odd ones, so that
                                // bit j = element 2j and bit 8 + j = element 2j + 1 (the order the backward chain's packed-half sign flip wants).
#pragma unroll
                                for (int j = 0; j < 8; j++) sgn_m = __funnelshift_r(sgn_m, (uint32_t)t2[j], 1);
#pragma unroll
                                for (int j = 0; j < 8; j++) sgn_m = __funnelshift_r(sgn_m, (uint32_t)(t2[j] >> 32), 1);
                            }
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                if (j >= 8 - POLY_PAIRS) {
                                    const uint64_t x2 = pk2(v[2 * j], v[2 * j + 1]);
                                    const uint64_t rr = fma2(add2(t2[j], nmagic2), npi2, x2);          // r = u - k pi in [-pi/2, pi/2]
                                    const uint64_t r2 = mul2(rr, rr);
                                    uint64_t p = fma2(c7, r2, c5);
                                    p = fma2(p, r2, c3);
                                    const uint64_t y = fma2(rr, mul2(p, r2), rr);                      // r + r^3 (c3 + c5 r^2 + c7 r^4)
                                    float ylo, yhi;
                                    upk2(y, ylo, yhi);
                                    v[2 * j] = __uint_as_float(__float_as_uint(ylo) ^ ((uint32_t)t2[j] << 31));
                                    v[2 * j + 1] = __uint_as_float(__float_as_uint(yhi) ^ ((uint32_t)(t2[j] >> 32) << 31));
                                } else {
                                    v[2 * j] = __sinf(v[2 * j]);
                                    v[2 * j + 1] = __sinf(v[2 * j + 1]);
                                }
                            }
                        }
                        if (L_nh) {
#pragma unroll
                            for (int hd = 0; hd < 3; hd++) {
                                if ((uint32_t)hd < L_nh) {
#pragma unroll
                                    for (int k = 0; k < 16; k += 4) {
                                        const float4 w4 = lds128(heads_s + (hd * 256 + col + k) * 4);
                                        hacc2[hd] = fma2(pk2(v[k], v[k + 1]), pk2(w4.x, w4.y), hacc2[hd]);
                                        hacc2[hd] = fma2(pk2(v[k + 2], v[k + 3]), pk2(w4.z, w4.w), hacc2[hd]);
                                    }
                                }
                            }
                        }
                        if (L_to_act) {
                            uint32_t hp[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) hp[j] = pack_f16(v[2 * j], v[2 * j + 1]);
                            const uint4 h0 = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                            const uint4 h1 = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                            if (do_sgn) {
                                // rounding bits: the fp16 conversion rounded the magnitude UP exactly where the first dropped mantissa bit (bit 12 of
                                // the fp32 word) is set (ties and fp16 subnormals aside); pair j contributes bit j (low half) and bit 16 + j (high half)
                                uint32_t rm = 0;
#if SDFG_RBIT
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const uint32_t t = __byte_perm(__float_as_uint(v[2 * j]), __float_as_uint(v[2 * j + 1]), 0x5511);   // bit 12 -> bits 4, 20
                                    rm |= (j >= 4 ? t << (j - 4) : t >> (4 - j)) & (0x00010001u << j);
                                }
#endif
                                const uint32_t plane = (sgn_m >> 16) | __byte_perm(rm, 0, 0x2044);      // [sign16 | rounding16]
                                asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_u32(smSGN) + (((n & 1) * 16 + c * 4 + sb) * 128 + r) * 4), "r"(plane) : "memory");
                            }
                            if (SAVE) mbar_wait(&S.st_done[c], (stgen & 1) ^ 1);   // the previous contents of the chunk have been stored
                            const uint32_t chunk = act_row + c * CH_CHUNK_BYTES;
                            sts128(chunk + u0, h0);
                            sts128(chunk + u1, h1);
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) {
                                if (fin) mbar_arrive(&S.fin_ready[c]);
                                else {
                                    arrive_mma(&S.act_ready[c]);
                                    if (SAVE) mbar_arrive(&S.act_ready_st[c]);
                                }
                            }
                        }
                        if (GEN && o32_row) {
                            float4* dst = reinterpret_cast<float4*>(o32_row + col);
#pragma unroll
                            for (int j = 0; j < 4; j++) dst[j] = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                        }
                    }
                    if (L_to_act) stgen++;
                    // every TMEM read of this layer has completed (wait::ld): hand the accumulator back to the MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_mma(&S.acc_empty[acc]);
                    if (L_nh) {                                             // combine the four sub-blocks' partial dot products
                        float hacc[3];
#pragma unroll
                        for (int hd = 0; hd < 3; hd++) { float lo, hi; upk2(hacc2[hd], lo, hi); hacc[hd] = lo + hi; }
                        if (sb != 0) {
#pragma unroll
                            for (int hd = 0; hd < 3; hd++) S.hx[sb - 1][r][hd] = hacc[hd];
                        }
                        named_bar_sync(2 + q, 128);                       // only the four warps that share these rows (one per column sub-block)
                        if (sb == 0 && valid) {
                            float* oh = P.layer[i].out_head;
#pragma unroll
                            for (int hd = 0; hd < 3; hd++)
                                if ((uint32_t)hd < L_nh) oh[row * L_nh + hd] = hacc[hd] + S.hx[0][r][hd] + S.hx[1][r][hd] + S.hx[2][r][hd] + S.hbias[hrow + hd];
                        }
                        named_bar_sync(2 + q, 128);                       // hx is free again before the next head layer writes it
                        hrow += L_nh;
                    }
                };
                switch (P.layer[i].kind) {
                    case FK_FILM: body(std::integral_constant<int, FK_FILM>{}); break;
                    case FK_FILM_SDF: body(std::integral_constant<int, FK_FILM_SDF>{}); break;
                    case FK_VIEWS_FIN: body(std::integral_constant<int, FK_VIEWS_FIN>{}); break;
                    case FK_SDF_FIN: body(std::integral_constant<int, FK_SDF_FIN>{}); break;
                    default: body(std::integral_constant<int, FK_GENERIC>{}); break;
                }
            }
        }
    }
    // teardown: the epilogues consumed the last accumulator, so every MMA and TMA load issued has completed
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == CH_WARP_MMA) { if (PAIR) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace sdfg
