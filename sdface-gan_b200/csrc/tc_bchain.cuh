// The backward pass through the FiLM-SIREN layers as ONE persistent tcgen05 kernel: for a 128-sample tile the gradient flows from
// the top layer to the input without leaving the SM (ref: autograd of FiLMSiren.forward sdf_model.py:61-69 chained through
// NGPSIRENGenerator.forward :1566-1592; the same kernel runs the eikonal chain d sdf / d x of get_eikonal_term :224-229).
//
// Per layer l (top -> bottom), two GEMMs share the tile:
//   R_l   u  = A_l W_l^T                 recompute of the pre-activation from the SAVED layer input A_l (fp16, HBM -> TMA ring)
//   epi_R du = dh * cos(gamma u + c)     dh: fp16 gradient tile G in shared memory (+ rank-r head terms / fp32 d_feat for the
//                                        top layer); du overwrites G in place, chunk by chunk           [MUFU.COS bound]
//   D_l   dh' = du (gamma o W_l)         A = G (K-major as written by the epilogue), B = per-image (gamma o W_l)^T chunks
//   epi_D G = fp16(dh' [+ d_sdf w_sigma])  -> input gradient of the layer below
// and a final N = in_dim GEMM gives d x_in = dh_0 W_in (fp32 out).  R_{l-1} is issued right behind D_l, so its MMAs overlap epi_D.
// Gradients are fp16 with the power-of-two loss scale of field_tc.cu (gscale = {s, 1/s}); stores saturate.
// With STORE the du tiles (and dh_0) are TMA-stored to HBM for the sample-axis weight-gradient kernels (tc_wgrad.cuh) that run
// afterwards; G chunks are only overwritten once the store has read them (st_done).
// CG = 2: two CTAs of a cluster run one tcgen05.mma.cta_group::2 (M = 256) per K-step: each stages its own 128 rows of A / G
// and HALF of every weight chunk, so the L2 -> SM weight stream (the measured bound of the CG = 1 kernel: ~1.3 MB per tile,
// 6.9 TB/s over the chip) is halved.  The leader CTA's MMA thread issues for the pair; the peer's epilogue warps arrive on
// the leader's barriers through the cluster (mapa + mbarrier.arrive.shared::cluster), commits are multicast to both CTAs.
// Algorithmic HBM traffic per sample and layer: 512 B (A_l) in, 512 B (du) out -- the per-layer kernels moved 2.5 KB.
#pragma once
#include "tc_chain.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t BC_MAX_LAYERS = SDFG_MAX_FILM;            // FiLM layers incl. views
// Two operand rings.  A ring: 4 x 16 KB = one whole saved-activation tile [128 x 256] -- the HBM loads of the NEXT layer's recompute
// are in flight while this layer's epilogue runs.  W ring: weight chunks ([256 / CG rows] x 64) of R_l, D_l, R_{l-1}, ... in issue
// order (L2 hits); a CTA pair (CG = 2) stages half of every chunk per CTA.
constexpr uint32_t BC_A_STAGES = 4;
__host__ __device__ constexpr uint32_t bc_w_bytes(int cg) { return 32768u / (uint32_t)cg; }
__host__ __device__ constexpr uint32_t bc_w_stages(int cg) { return cg == 2 ? 5u : 2u; }
constexpr uint32_t BC_MAX_W_STAGES = 5;
constexpr uint32_t BC_G_BYTES = 4 * CH_CHUNK_BYTES;          // gradient tile [128 x 256] fp16

struct BLayer {
    uint32_t nk, last_ksteps;   // recompute GEMM: K chunks of 64 and K-steps (of 16) in the last chunk
    uint32_t film;              // row of gamma / beta
    uint32_t r_src_g;           // epi_R: dh comes from G
    uint32_t r_rank, r_vec0;    // epi_R: + sum_r gs * r_rank_s[row*r_rank + r] * vecs[r_vec0 + r][col]
    uint32_t do_D;              // run D (the layer below needs its gradient)
    uint32_t d_rank, d_vec0;    // epi_D: + gs * d_rank_s[row] * vecs[d_vec0][col]
    uint32_t pad;
    const float* r_rank_s;
    const float* r_dfeat;       // fp32 [M, 256] added to dh (times gs), or NULL
    const float* d_rank_s;
    const float* bias;
};

struct BChainParams {
    uint32_t M_total, rows_per_image, n_units, units_per_cta, n_layers;   // unit = CG adjacent 128-row tiles (one per CTA of the pair)
    uint32_t has_in, in_dim;    // final stage: d_x_in[M, in_dim] = gs_inv * dh_0 W_in
    uint32_t pad;
    float* d_x_in;
    const float* gamma;         // + img * gstride + film * 256 + n
    const float* beta;
    int64_t gstride;
    const float* gscale;        // {s, 1/s}
    const float* vecs[4];       // rank vectors [256] each: 0 = w_sigma, 1..3 = w_rgb rows
    unsigned long long* dbg;    // debugging: event log of CTA 0 (see tc_chain.cuh)
    BLayer layer[BC_MAX_LAYERS];   // index 0 = TOP layer
};

struct alignas(64) BChainMaps {
    CUtensorMap a[BC_MAX_LAYERS];      // saved layer input A_l      [M, K_l]      box 128 x 64
    CUtensorMap w[BC_MAX_LAYERS];      // W_l fp16                    [256, K_l]    box 256 x 64
    CUtensorMap wgt[BC_MAX_LAYERS];    // (gamma o W_l)^T per image   [B*256, 256]  box 256 x 64
    CUtensorMap dz[BC_MAX_LAYERS];     // du store                    [M, 256]      box 128 x 64
    CUtensorMap wgt_in;                // W_in^T                      [in_dim, 256] box in_dim x 64
    CUtensorMap dh0;                   // dh_0 store                  [M, 256]      box 128 x 64
};

struct BChainSmem {
    uint64_t a_full[BC_A_STAGES], a_empty;      // one buffer = one tile: per-chunk arrival, released once per layer
    uint64_t w_full[BC_MAX_W_STAGES], w_empty[BC_MAX_W_STAGES];
    uint64_t dz_ready[4], dh0_ready[4];          // MMA side (on the leader CTA: every epilogue warp of the pair arrives)
    uint64_t dz_ready_st[4], dh0_ready_st[4];    // storer side (local)
    uint64_t st_done[4];
    uint64_t accR_full, accR_empty, accD_full, accD_empty;
    uint32_t tmem_base;
    uint32_t pad[3];
    alignas(16) float gam[2][256];
    float cst[2][256];
    float vecs[4][256];
};

__host__ __device__ inline uint32_t bchain_smem_bytes(int cg) {
    return 1024 + BC_G_BYTES + BC_A_STAGES * CH_CHUNK_BYTES + bc_w_stages(cg) * bc_w_bytes(cg) + (uint32_t)sizeof(BChainSmem);
}

__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

template <bool STORE, int CG>
__global__ void __launch_bounds__(CH_THREADS, 1)
tc_chain_bwd_kernel(const __grid_constant__ BChainMaps maps, const __grid_constant__ BChainParams P) {
    constexpr bool PAIR = CG == 2;
    constexpr uint32_t W_BYTES = bc_w_bytes(CG), NW = bc_w_stages(CG), NA = BC_A_STAGES;
    constexpr uint32_t W_ROWS = 256 / CG;                              // weight rows (output neurons) staged per CTA
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smG = smem;
    uint8_t* smA = smG + BC_G_BYTES;
    uint8_t* smW = smA + NA * CH_CHUNK_BYTES;
    BChainSmem& S = *reinterpret_cast<BChainSmem*>(smW + NW * W_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;               // 0 = leader (issues the MMAs of the pair)
    const bool leader = rank == 0;
    const uint32_t u_begin = (blockIdx.x / CG) * P.units_per_cta;
    const uint32_t u_end = min(P.n_units, u_begin + P.units_per_cta);
    const uint32_t nL = P.n_layers;
    uint32_t dbg_n = 0;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < NA; i++) mbar_init(&S.a_full[i], 1);
        mbar_init(&S.a_empty, 1);
        for (uint32_t i = 0; i < NW; i++) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (uint32_t i = 0; i < 4; i++) {
            mbar_init(&S.dz_ready[i], CH_EPI_WARPS * CG); mbar_init(&S.dh0_ready[i], CH_EPI_WARPS * CG);
            mbar_init(&S.dz_ready_st[i], CH_EPI_WARPS); mbar_init(&S.dh0_ready_st[i], CH_EPI_WARPS);
            mbar_init(&S.st_done[i], 1);
        }
        mbar_init(&S.accR_full, 1); mbar_init(&S.accR_empty, CH_EPI_WARPS * CG);
        mbar_init(&S.accD_full, 1); mbar_init(&S.accD_empty, CH_EPI_WARPS * CG);
        fence_barrier_init();
    }
    if (warp == CH_WARP_TMA && lane == 0) {
        for (uint32_t i = 0; i < nL; i++) {
            tma_prefetch_desc(&maps.a[i]); tma_prefetch_desc(&maps.w[i]);
            if (P.layer[i].do_D) tma_prefetch_desc(&maps.wgt[i]);
            if (STORE) tma_prefetch_desc(&maps.dz[i]);
        }
        if (P.has_in) tma_prefetch_desc(&maps.wgt_in);
    }
    if (warp == CH_WARP_MMA) { if (PAIR) tmem_alloc_2cta(&S.tmem_base, 512); else tmem_alloc(&S.tmem_base, 512); }
    for (uint32_t i = threadIdx.x; i < 4 * 256; i += blockDim.x) S.vecs[i >> 8][i & 255] = P.vecs[i >> 8] ? __ldg(P.vecs[i >> 8] + (i & 255)) : 0.f;
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();              // the peer's barriers exist before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const uint32_t in_rows = P.in_dim / CG, in_box_bytes = in_rows * 128;

    if (warp == CH_WARP_TMA) {
        // ===================================================== weight producer (both CTAs): own half of every chunk, in MMA issue order
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            auto put = [&](const CUtensorMap* m, uint32_t bytes, int32_t c0, int32_t c1) {
                mbar_wait(&S.w_empty[stage], phase ^ 1);
                if (leader) mbar_arrive_expect_tx(&S.w_full[stage], CG * bytes);
                if (PAIR) tma_load_2d_2cta(smW + stage * W_BYTES, m, &S.w_full[stage], c0, c1);
                else tma_load_2d(smW + stage * W_BYTES, m, &S.w_full[stage], c0, c1);
                if (++stage == NW) { stage = 0; phase ^= 1; }
            };
            for (uint32_t u = u_begin; u < u_end; u++) {
                const uint32_t t = u * CG + rank;
                const int32_t img = (int32_t)((t * CH_TILE_M) / P.rows_per_image);
                for (uint32_t i = 0; i < nL; i++) {
                    const uint32_t nk = P.layer[i].nk;
                    for (uint32_t kc = 0; kc < nk; kc++) {
                        // a 5th K chunk (view tail) does not fit the A buffer: its activation half travels through this ring too
                        if (kc >= NA) put(&maps.a[i], CH_CHUNK_BYTES, (int32_t)(kc * 64), (int32_t)(t * CH_TILE_M));
                        put(&maps.w[i], W_BYTES, (int32_t)(kc * 64), (int32_t)(rank * W_ROWS));
                    }
                    if (P.layer[i].do_D)
                        for (uint32_t kc = 0; kc < 4; kc++) put(&maps.wgt[i], W_BYTES, (int32_t)(kc * 64), img * 256 + (int32_t)(rank * W_ROWS));
                }
                if (P.has_in)
                    for (uint32_t kc = 0; kc < 4; kc++) put(&maps.wgt_in, in_box_bytes, (int32_t)(kc * 64), (int32_t)(rank * in_rows));
            }
        }
    } else if (warp == CH_WARP_LOAD) {
        // ===================================================== activation producer (both CTAs): saved layer inputs of the own tile
        if (lane == 0) {
            uint32_t agen = 0;
            for (uint32_t u = u_begin; u < u_end; u++) {
                const int32_t row0 = (int32_t)((u * CG + rank) * CH_TILE_M);
                for (uint32_t i = 0; i < nL; i++, agen++) {
                    const uint32_t nk = min(P.layer[i].nk, NA);
                    mbar_wait(&S.a_empty, (agen & 1) ^ 1);            // every MMA that read the previous tile has completed
                    for (uint32_t kc = 0; kc < nk; kc++) {
                        if (leader) mbar_arrive_expect_tx(&S.a_full[kc], CG * CH_CHUNK_BYTES);
                        if (PAIR) tma_load_2d_2cta(smA + kc * CH_CHUNK_BYTES, &maps.a[i], &S.a_full[kc], (int32_t)(kc * 64), row0);
                        else tma_load_2d(smA + kc * CH_CHUNK_BYTES, &maps.a[i], &S.a_full[kc], (int32_t)(kc * 64), row0);
                    }
                }
            }
        }
    } else if (warp == CH_WARP_MMA) {
        // ===================================================== MMA issuer (leader CTA only)
        if (lane == 0 && leader) {
            const uint32_t idesc = idesc_f16(CH_TILE_M * CG, 256, FMT_F16, FMT_F16, 0, 0);
            const uint32_t idesc_in = idesc_f16(CH_TILE_M * CG, P.in_dim, FMT_F16, FMT_F16, 0, 0);
            const uint32_t g_addr = smem_u32(smG);
            uint32_t stage = 0, phase = 0, nR = 0, nD = 0, dzgen = 0, it = 0;
            auto next = [&]() { if (++stage == NW) { stage = 0; phase ^= 1; } };
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t accum) {
                if (PAIR) umma_f16_2cta(d, da, db, id, accum); else umma_bf16(d, da, db, id, accum);
            };
            auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_2cta(bar, 3); else umma_commit(bar); };
            auto wait_x = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };
            for (uint32_t u = u_begin; u < u_end; u++, it++) {
                for (uint32_t i = 0; i < nL; i++) {
                    const uint32_t nk = P.layer[i].nk, lks = P.layer[i].last_ksteps;
                    // ---- R_l -> accumulator 0
                    wait_x(&S.accR_empty, (nR & 1) ^ 1);
                    tc_fence_after();
                    CH_DBG(0, 50 + i);
                    uint32_t accumulate = 0;
                    for (uint32_t kc = 0; kc < nk; kc++) {
                        uint32_t a_addr;
                        uint64_t* tail_bar = nullptr;
                        if (kc < NA) {
                            mbar_wait(&S.a_full[kc], nR & 1);
                            a_addr = smem_u32(smA + kc * CH_CHUNK_BYTES);
                        } else {                                          // view tail: the activation chunk sits in the W ring
                            mbar_wait(&S.w_full[stage], phase);
                            a_addr = smem_u32(smW + stage * W_BYTES);
                            tail_bar = &S.w_empty[stage];
                            next();
                        }
                        CH_DBG(0, 3000 + i * 16 + kc);
                        mbar_wait(&S.w_full[stage], phase);
                        tc_fence_after();
                        CH_DBG(0, 4000 + i * 16 + kc);
                        const uint32_t b_addr = smem_u32(smW + stage * W_BYTES);
                        const uint32_t ks = kc + 1 == nk ? lks : 4;
                        for (uint32_t s = 0; s < ks; s++, accumulate = 1)
                            mma(tmem_base, smem_desc_sw128(a_addr + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), idesc, accumulate);
                        if (tail_bar) commit(tail_bar);
                        commit(&S.w_empty[stage]);
                        next();
                    }
                    commit(&S.a_empty);
                    commit(&S.accR_full);
                    CH_DBG(0, 100 + i);
                    nR++;
                    // ---- D_l -> accumulator 1, chunk by chunk behind the epilogue(s)
                    if (P.layer[i].do_D) {
                        wait_x(&S.accD_empty, (nD & 1) ^ 1);
                        tc_fence_after();
                        for (uint32_t kc = 0; kc < 4; kc++) {
                            wait_x(&S.dz_ready[kc], dzgen & 1);
                            CH_DBG(0, 1000 + i * 16 + kc);
                            mbar_wait(&S.w_full[stage], phase);
                            tc_fence_after();
                            CH_DBG(0, 2000 + i * 16 + kc);
                            const uint32_t b_addr = smem_u32(smW + stage * W_BYTES), a_addr = g_addr + kc * CH_CHUNK_BYTES;
                            for (uint32_t s = 0; s < 4; s++)
                                mma(tmem_base + 256, smem_desc_sw128(a_addr + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), idesc, (kc | s) != 0);
                            commit(&S.w_empty[stage]);
                            CH_DBG(0, 200 + i * 16 + kc);
                            next();
                        }
                        commit(&S.accD_full);
                        nD++;
                    } else {
                        for (uint32_t kc = 0; kc < 4; kc++) wait_x(&S.dz_ready[kc], dzgen & 1);   // keep the phase parity in step
                    }
                    dzgen++;
                }
                if (P.has_in) {
                    wait_x(&S.accD_empty, (nD & 1) ^ 1);
                    tc_fence_after();
                    for (uint32_t kc = 0; kc < 4; kc++) {
                        wait_x(&S.dh0_ready[kc], it & 1);
                        mbar_wait(&S.w_full[stage], phase);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(smW + stage * W_BYTES), a_addr = g_addr + kc * CH_CHUNK_BYTES;
                        for (uint32_t s = 0; s < 4; s++)
                            mma(tmem_base + 256, smem_desc_sw128(a_addr + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), idesc_in, (kc | s) != 0);
                        commit(&S.w_empty[stage]);
                        next();
                    }
                    commit(&S.accD_full);
                    nD++;
                }
            }
        }
    } else if (warp == CH_WARP_STORE) {
        // ===================================================== storer (STORE): du tiles and dh_0 -> HBM for the weight-gradient kernels
        if (STORE && lane == 0) {
            uint32_t dzgen = 0, it = 0;
            for (uint32_t u = u_begin; u < u_end; u++, it++) {
                const int32_t row0 = (int32_t)((u * CG + rank) * CH_TILE_M);
                for (uint32_t i = 0; i < nL; i++, dzgen++)
                    for (uint32_t c = 0; c < 4; c++) {
                        mbar_wait(&S.dz_ready_st[c], dzgen & 1);
                        tma_store_2d(&maps.dz[i], smG + c * CH_CHUNK_BYTES, (int32_t)(c * 64), row0);
                        tma_store_commit();
                        tma_store_wait_read();
                        mbar_arrive(&S.st_done[c]);
                    }
                if (P.has_in)
                    for (uint32_t c = 0; c < 4; c++) {
                        mbar_wait(&S.dh0_ready_st[c], it & 1);
                        tma_store_2d(&maps.dh0, smG + c * CH_CHUNK_BYTES, (int32_t)(c * 64), row0);
                        tma_store_commit();
                        tma_store_wait_read();
                        mbar_arrive(&S.st_done[c]);
                    }
            }
            tma_store_wait_all();
        }
    } else if (warp < CH_EPI_WARPS) {
        // ===================================================== epilogue: 16 warps, 4 per TMEM lane quarter, 16 columns of every chunk each
        const uint32_t q = warp & 3, sb = warp >> 2;
        const uint32_t etid = threadIdx.x;
        const uint32_t r = q * 32 + lane;
        const uint32_t g_row = smem_u32(smG) + r * 128;
        const uint32_t u0 = ((2 * sb) ^ (r & 7)) << 4, u1 = ((2 * sb + 1) ^ (r & 7)) << 4;
        const float gs = __ldg(P.gscale), gs_inv = __ldg(P.gscale + 1);
        const uint32_t lane_base = (q * 32) << 16;
        // arrive on a barrier the leader's MMA thread waits on
        auto arrive_mma = [&](uint64_t* bar) { if (PAIR && !leader) mbar_arrive_remote(bar, 0); else mbar_arrive(bar); };
        uint32_t nR = 0, nD = 0, stgen = 0, n = 0;
        bool stored_dh0 = false;                                        // G holds a dh_0 tile the storer is (or was) reading
        for (uint32_t u = u_begin; u < u_end; u++) {
            const uint32_t t = u * CG + rank;
            const uint64_t row = (uint64_t)t * CH_TILE_M + r;
            const uint32_t img = (t * CH_TILE_M) / P.rows_per_image;
            for (uint32_t i = 0; i < nL; i++, n++) {
                const uint32_t L_film = P.layer[i].film, r_src_g = P.layer[i].r_src_g, r_rank = P.layer[i].r_rank, r_vec0 = P.layer[i].r_vec0;
                const uint32_t do_D = P.layer[i].do_D, d_rank = P.layer[i].d_rank, d_vec0 = P.layer[i].d_vec0;
                const float* const dfeat = P.layer[i].r_dfeat;
                const uint32_t tb = n & 1;
                const uint32_t gam_s = smem_u32(&S.gam[tb][0]), cst_s = smem_u32(&S.cst[tb][0]);
                {   // FiLM constants of this layer: threads 0..255 gamma, 256..511 gamma*bias + beta
                    const uint32_t col = etid & 255;
                    const float gm = __ldg(P.gamma + (int64_t)img * P.gstride + L_film * 256 + col);
                    if (etid < 256) sts32(gam_s + col * 4, gm);
                    else sts32(cst_s + col * 4, fmaf(gm, __ldg(P.layer[i].bias + col), __ldg(P.beta + (int64_t)img * P.gstride + L_film * 256 + col)));
                    named_bar_sync(1, CH_EPI_THREADS);
                }
                float rs[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 3; k++)
                    if ((uint32_t)k < r_rank) rs[k] = gs * __ldg(P.layer[i].r_rank_s + row * r_rank + k);
                const float ds = d_rank ? gs * __ldg(P.layer[i].d_rank_s + row) : 0.f;
                const uint32_t rvec_s = smem_u32(&S.vecs[r_vec0][0]), dvec_s = smem_u32(&S.vecs[d_vec0][0]);

                // ---------------- epi_R: du = dh * cos(gamma u + c), in place in G
                if (threadIdx.x == 0) CH_DBG(1, 300 + i);
                mbar_wait(&S.accR_full, nR & 1);
                tc_fence_after();
                if (threadIdx.x == 0) CH_DBG(1, 400 + i);
                {
                    const uint32_t taddr = tmem_base + lane_base + sb * 16;
                    uint32_t raw[2][16];
                    tmem_ld16_issue(taddr, raw[0]);
#pragma unroll
                    for (uint32_t c = 0; c < 4; c++) {
                        const uint32_t col = c * 64 + sb * 16;
                        const uint32_t chunk = g_row + c * CH_CHUNK_BYTES;
                        float dh[16];
                        if (r_src_g) {
                            const uint4 a = lds128u(chunk + u0), b = lds128u(chunk + u1);
                            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int k = 0; k < 8; k++) { const float2 f = unpack_f16(w[k]); dh[2 * k] = f.x; dh[2 * k + 1] = f.y; }
                        } else {
#pragma unroll
                            for (int k = 0; k < 16; k++) dh[k] = 0.f;
                        }
                        if (dfeat) {
                            const float4* src = reinterpret_cast<const float4*>(dfeat + row * 256 + col);
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const float4 f = ldg_stream4(src + j);
                                dh[4 * j] = fmaf(gs, f.x, dh[4 * j]); dh[4 * j + 1] = fmaf(gs, f.y, dh[4 * j + 1]);
                                dh[4 * j + 2] = fmaf(gs, f.z, dh[4 * j + 2]); dh[4 * j + 3] = fmaf(gs, f.w, dh[4 * j + 3]);
                            }
                        }
                        if (r_rank) {
#pragma unroll
                            for (int rr = 0; rr < 3; rr++) {
                                if ((uint32_t)rr < r_rank) {
#pragma unroll
                                    for (int k = 0; k < 16; k += 4) {
                                        const float4 w4 = lds128(rvec_s + (rr * 256 + col + k) * 4);
                                        dh[k] = fmaf(rs[rr], w4.x, dh[k]); dh[k + 1] = fmaf(rs[rr], w4.y, dh[k + 1]);
                                        dh[k + 2] = fmaf(rs[rr], w4.z, dh[k + 2]); dh[k + 3] = fmaf(rs[rr], w4.w, dh[k + 3]);
                                    }
                                }
                            }
                        }
                        tmem_ld_wait16(raw[c & 1]);
                        if (c < 3) tmem_ld16_issue(taddr + (c + 1) * 64, raw[(c + 1) & 1]);
                        float v[16];
#pragma unroll
                        for (int k = 0; k < 16; k += 4) {
                            const float4 g4 = lds128(gam_s + (col + k) * 4);
                            const float4 c4 = lds128(cst_s + (col + k) * 4);
                            v[k] = dh[k] * __cosf(fmaf(__uint_as_float(raw[c & 1][k]), g4.x, c4.x));
                            v[k + 1] = dh[k + 1] * __cosf(fmaf(__uint_as_float(raw[c & 1][k + 1]), g4.y, c4.y));
                            v[k + 2] = dh[k + 2] * __cosf(fmaf(__uint_as_float(raw[c & 1][k + 2]), g4.z, c4.z));
                            v[k + 3] = dh[k + 3] * __cosf(fmaf(__uint_as_float(raw[c & 1][k + 3]), g4.w, c4.w));
                        }
                        const uint4 h0 = make_uint4(pack_f16_sat(v[0], v[1]), pack_f16_sat(v[2], v[3]), pack_f16_sat(v[4], v[5]), pack_f16_sat(v[6], v[7]));
                        const uint4 h1 = make_uint4(pack_f16_sat(v[8], v[9]), pack_f16_sat(v[10], v[11]), pack_f16_sat(v[12], v[13]), pack_f16_sat(v[14], v[15]));
                        if (STORE && stored_dh0 && i == 0) mbar_wait(&S.st_done[c], stgen & 1);   // the previous tile's dh_0 store has read the chunk
                        sts128(chunk + u0, h0);
                        sts128(chunk + u1, h1);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            arrive_mma(&S.dz_ready[c]);
                            if (STORE) mbar_arrive(&S.dz_ready_st[c]);
                        }
                        if (threadIdx.x == 0) CH_DBG(1, 500 + i * 16 + c);
                        if (threadIdx.x == 0 && P.dbg && blockIdx.x == 1 && dbg_n < 1023) { P.dbg[2 * 2048 + 2 * dbg_n] = 500 + i * 16 + c; P.dbg[2 * 2048 + 2 * dbg_n + 1] = clock64(); dbg_n++; }
                    }
                    if (STORE && stored_dh0 && i == 0) { stgen++; stored_dh0 = false; }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_mma(&S.accR_empty);
                nR++;

                // ---------------- epi_D: G = fp16(dh' + rank-1 term)
                if (do_D) {
                    if (threadIdx.x == 0) CH_DBG(1, 600 + i);
                    mbar_wait(&S.accD_full, nD & 1);
                    tc_fence_after();
                    if (threadIdx.x == 0) CH_DBG(1, 700 + i);
                    const uint32_t taddr = tmem_base + lane_base + 256 + sb * 16;
                    const bool to_in = P.has_in && i + 1 == nL;
                    uint32_t raw[2][16];
                    tmem_ld16_issue(taddr, raw[0]);
#pragma unroll
                    for (uint32_t c = 0; c < 4; c++) {
                        const uint32_t col = c * 64 + sb * 16;
                        const uint32_t chunk = g_row + c * CH_CHUNK_BYTES;
                        tmem_ld_wait16(raw[c & 1]);
                        if (c < 3) tmem_ld16_issue(taddr + (c + 1) * 64, raw[(c + 1) & 1]);
                        float v[16];
#pragma unroll
                        for (int k = 0; k < 16; k++) v[k] = __uint_as_float(raw[c & 1][k]);
                        if (d_rank) {
#pragma unroll
                            for (int k = 0; k < 16; k += 4) {
                                const float4 w4 = lds128(dvec_s + (col + k) * 4);
                                v[k] = fmaf(ds, w4.x, v[k]); v[k + 1] = fmaf(ds, w4.y, v[k + 1]);
                                v[k + 2] = fmaf(ds, w4.z, v[k + 2]); v[k + 3] = fmaf(ds, w4.w, v[k + 3]);
                            }
                        }
                        const uint4 h0 = make_uint4(pack_f16_sat(v[0], v[1]), pack_f16_sat(v[2], v[3]), pack_f16_sat(v[4], v[5]), pack_f16_sat(v[6], v[7]));
                        const uint4 h1 = make_uint4(pack_f16_sat(v[8], v[9]), pack_f16_sat(v[10], v[11]), pack_f16_sat(v[12], v[13]), pack_f16_sat(v[14], v[15]));
                        if (STORE) mbar_wait(&S.st_done[c], stgen & 1);     // the du store of this layer has read the chunk
                        sts128(chunk + u0, h0);
                        sts128(chunk + u1, h1);
                        if (to_in) {
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) {
                                arrive_mma(&S.dh0_ready[c]);
                                if (STORE) mbar_arrive(&S.dh0_ready_st[c]);
                            }
                        }
                    }
                    if (STORE) stgen++;
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_mma(&S.accD_empty);
                    if (threadIdx.x == 0) CH_DBG(1, 800 + i);
                    nD++;
                } else if (STORE) {
                    // no D: the du store of this layer still owns G until the storer is done with it
                    for (uint32_t c = 0; c < 4; c++) mbar_wait(&S.st_done[c], stgen & 1);
                    stgen++;
                }
            }
            // ---------------- input stage: d_x_in = gs_inv * acc
            if (P.has_in) {
                mbar_wait(&S.accD_full, nD & 1);
                tc_fence_after();
                if (sb * 16 < P.in_dim && P.d_x_in) {                   // warp-uniform: tcgen05.ld is a whole-warp instruction
                    uint32_t raw[16];
                    tmem_ld16(tmem_base + lane_base + 256 + sb * 16, raw);
                    tmem_ld_wait();
                    if (row < P.M_total) {
                        float4* dst = reinterpret_cast<float4*>(P.d_x_in + row * P.in_dim + sb * 16);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            dst[j] = make_float4(gs_inv * __uint_as_float(raw[4 * j]), gs_inv * __uint_as_float(raw[4 * j + 1]),
                                                 gs_inv * __uint_as_float(raw[4 * j + 2]), gs_inv * __uint_as_float(raw[4 * j + 3]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_mma(&S.accD_empty);
                nD++;
                if (STORE) stored_dh0 = true;
            }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();              // nobody may exit while the pair can still touch its shared / tensor memory
    if (warp == CH_WARP_MMA) { if (PAIR) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace sdfg
