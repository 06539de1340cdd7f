// Backward chain: the forward chain saved every FiLM layer's output s_l = sin(gamma u_l + c) (fp16) and the sign of its derivative
// cos(gamma u_l + c) as one bit per element (tc_chain.cuh), so the gradient of a 128-sample tile flows through the layers with ONE
// GEMM per layer and a light epilogue (cos = +-sqrt(1 - s^2)):
//
//   du_top = (head rank terms + d_feat) * c_top                                   (first epilogue, no GEMM)
//   for l = top .. bottom:   D_l: dh = du_l (gamma o W_l)   ->   epilogue: du_{l-1} = (dh [+ d_sdf w_sigma]) * c_{l-1}  -> G (fp16)
//   input stage:             d x_in = dh_0 W_in  (fp32 out)
//
// (ref: autograd of FiLMSiren.forward sdf_model.py:61-69 through NGPSIRENGenerator.forward :1566-1592; without stores and
// started from d_sdf = 1 this is the eikonal chain of get_eikonal_term :224-229).  dh never leaves TMEM in fp16: the epilogue
// reads the fp32 accumulator, multiplies by the cos tile (64 KB shared-memory buffer, TMA-prefetched chunk by chunk as the
// previous layer's epilogue releases it) and writes the next GEMM's A operand in place into G; D_{l-1}'s MMAs trail the
// epilogue at 64-column chunk granularity into the other accumulator.  Nothing is recomputed: no second GEMM per layer, no
// W_l stream, no FiLM constant tables.
// CG = 2: two CTAs of a cluster run one tcgen05.mma.cta_group::2 (M = 256) per K-step: each stages its own 128 rows of G and HALF
// of every weight chunk, so the L2 -> SM weight stream (~1.3 MB per tile, the measured bound of a single-CTA chain) is halved.
// The leader CTA's MMA thread issues for the pair, commits are multicast to both CTAs, the peer's epilogue warps arrive on the
// leader's barriers through the cluster (mapa + mbarrier.arrive.shared::cluster); cos tiles are per-CTA data on local barriers.
// Gradients are fp16 with the power-of-two loss scale of field_tc.cu (gscale = {s, 1/s}); stores saturate.
// With STORE the du tiles (and dh_0) are TMA-stored for the weight-gradient kernels (tc_wgrad.cuh).
// Algorithmic HBM traffic per sample and layer: 512 B (c_l) in, 512 B (du_l) out.
#pragma once
#include "tc_chain.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t BC_MAX_LAYERS = SDFG_MAX_FILM;            // FiLM layers incl. views
// weight ring: chunks ([256 / CG rows] x 64 fp16) of (gamma o W_l)^T in issue order (L2 hits); a CTA pair stages half of every chunk per CTA
__host__ __device__ constexpr uint32_t bc_w_bytes(int cg) { return 32768u / (uint32_t)cg; }
__host__ __device__ constexpr uint32_t bc_w_stages(int cg) { return cg == 2 ? 4u : 2u; }
constexpr uint32_t BC_MAX_W_STAGES = 4;
constexpr uint32_t BC_G_BYTES = 4 * CH_CHUNK_BYTES;          // gradient tile [128 x 256] fp16

__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

struct B2Layer {
    uint32_t do_D;              // run D (the layer below, or the input stage, needs the gradient)
    uint32_t d_rank, d_vec0;    // epilogue of D_l: + gs * d_rank_s[row] * vecs[d_vec0][col]   (d_sdf w_sigma behind the view layer)
    uint32_t pad;
    const float* d_rank_s;
    const uint8_t* sgn;         // sign(cos) masks of this layer, [tiles][CH_SGN_TILE_BYTES] (tc_chain.cuh)
};

struct B2ChainParams {
    uint32_t M_total, rows_per_image, n_units, units_per_cta, n_layers;
    uint32_t has_in, in_dim;            // input stage: d x_in = (last gradient tile) * wgt_in  -- dh_0 W_in, or, when the host collapsed
    uint32_t in_per_image;              // input_linear into the first FiLM layer, du_0 (gamma_b o W_0 W_in) with per-image weights
    uint32_t top_rank, top_vec0;        // first epilogue: dh_top = sum_r gs * top_rank_s[row*top_rank + r] * vecs[top_vec0 + r] + gs * top_dfeat
    const float* top_rank_s;
    const float* top_dfeat;             // fp32 [M, 256] or NULL
    float* d_x_in;
    // eikonal pass: instead of writing d_x_in [M, in_dim], contract it with the hash encoder's dy_dx ([in_dim / 2 levels][3][2][M], component
    // major) in the input-stage epilogue: eik_out[row, d] += eik_scale * sum_j d_x_in[row, j] dy_dx[level(j), d, c(j)][row]   (tc_bchain3 only)
    const float* eik_dydx;
    float* eik_out;                     // [M, 3], pre-zeroed (each row receives in_dim / 16 partial sums)
    float eik_scale;
    const float* gscale;                // {s, 1/s}
    const float* vecs[4];               // 0 = w_sigma, 1..3 = w_rgb rows
    unsigned long long* dbg;
    B2Layer layer[BC_MAX_LAYERS];       // index 0 = TOP layer
};

struct alignas(64) B2ChainMaps {
    CUtensorMap c[BC_MAX_LAYERS];       // sin tile of layer l (its saved output)  [M, 256]  box 128 x 64
    CUtensorMap wgt[BC_MAX_LAYERS];     // (gamma o W_l)^T per image    [B*256, 256]  box (256/CG) x 64
    CUtensorMap dz[BC_MAX_LAYERS];      // du store                     [M, 256]      box 128 x 64
    CUtensorMap wgt_in;                 // W_in^T [in_dim, 256] or per image [B*in_dim, 256]; box (in_dim/CG) x 64
    CUtensorMap dh0;                    // dh_0 store                   [M, 256]      box 128 x 64
};

struct B2ChainSmem {
    uint64_t c_full[4], c_empty[4];
    uint64_t w_full[BC_MAX_W_STAGES], w_empty[BC_MAX_W_STAGES];
    uint64_t g_ready[4], g_ready_st[4], st_done[4];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
    alignas(16) float vecs[4][256];
};

__host__ __device__ inline uint32_t bchain2_smem_bytes(int cg) {
    return 1024 + 2 * BC_G_BYTES + bc_w_stages(cg) * bc_w_bytes(cg) + 2 * CH_SGN_TILE_BYTES + (uint32_t)sizeof(B2ChainSmem);
}

template <bool STORE, int CG>
__global__ void __launch_bounds__(CH_THREADS, 1)
tc_chain_bwd2_kernel(const __grid_constant__ B2ChainMaps maps, const __grid_constant__ B2ChainParams P) {
    constexpr bool PAIR = CG == 2;
    constexpr uint32_t W_BYTES = bc_w_bytes(CG), NW = bc_w_stages(CG);
    constexpr uint32_t W_ROWS = 256 / CG;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smG = smem;                                               // gradient tile = A operand of the D GEMMs
    uint8_t* smC = smG + BC_G_BYTES;                                   // cos tile of the layer being entered
    uint8_t* smW = smC + BC_G_BYTES;
    uint8_t* smSGN = smW + NW * W_BYTES;                               // two sign-mask tiles (double-buffered by layer)
    B2ChainSmem& S = *reinterpret_cast<B2ChainSmem*>(smSGN + 2 * CH_SGN_TILE_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const uint32_t u_begin = (blockIdx.x / CG) * P.units_per_cta;
    const uint32_t u_end = min(P.n_units, u_begin + P.units_per_cta);
    const uint32_t nL = P.n_layers;
    uint32_t nD = 0;                                                   // D GEMMs per unit (layers with do_D)
    for (uint32_t i = 0; i < nL; i++) nD += P.layer[i].do_D ? 1u : 0u;
    uint32_t dbg_n = 0;
    (void)dbg_n;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < 4; i++) {
            mbar_init(&S.c_full[i], 1); mbar_init(&S.c_empty[i], CH_EPI_WARPS);
            mbar_init(&S.g_ready[i], CH_EPI_WARPS * CG); mbar_init(&S.g_ready_st[i], CH_EPI_WARPS); mbar_init(&S.st_done[i], 1);
        }
        for (uint32_t i = 0; i < NW; i++) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (uint32_t i = 0; i < 2; i++) { mbar_init(&S.acc_full[i], 1); mbar_init(&S.acc_empty[i], CH_EPI_WARPS * CG); }
        fence_barrier_init();
    }
    if (warp == CH_WARP_TMA && lane == 0) {
        for (uint32_t i = 0; i < nL; i++) {
            tma_prefetch_desc(&maps.c[i]);
            if (P.layer[i].do_D) tma_prefetch_desc(&maps.wgt[i]);
            if (STORE) tma_prefetch_desc(&maps.dz[i]);
        }
        if (P.has_in) tma_prefetch_desc(&maps.wgt_in);
    }
    if (warp == CH_WARP_MMA) { if (PAIR) tmem_alloc_2cta(&S.tmem_base, 512); else tmem_alloc(&S.tmem_base, 512); }
    for (uint32_t i = threadIdx.x; i < 4 * 256; i += blockDim.x) S.vecs[i >> 8][i & 255] = P.vecs[i >> 8] ? __ldg(P.vecs[i >> 8] + (i & 255)) : 0.f;
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const uint32_t in_rows = P.in_dim / CG, in_box_bytes = in_rows * 128;

    if (warp == CH_WARP_TMA) {
        // ===================================================== weight producer (both CTAs): own half of every chunk, in MMA issue order
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            // every CTA re-reads the per-image weights (a few MB in total) for each of its tiles: keep them in L2 while the
            // saved-activation reads and the du stores (GBs, touched once) stream past
            const uint64_t keep = l2_policy_evict_last();
            auto put = [&](const CUtensorMap* m, uint32_t bytes, int32_t c0, int32_t c1) {
                mbar_wait(&S.w_empty[stage], phase ^ 1);
                if (leader) mbar_arrive_expect_tx(&S.w_full[stage], CG * bytes);
                if (PAIR) tma_load_2d_2cta_hint(smW + stage * W_BYTES, m, &S.w_full[stage], c0, c1, keep);
                else tma_load_2d_hint(smW + stage * W_BYTES, m, &S.w_full[stage], c0, c1, keep);
                if (++stage == NW) { stage = 0; phase ^= 1; }
            };
            for (uint32_t u = u_begin; u < u_end; u++) {
                const uint32_t t = u * CG + rank;
                const int32_t img = (int32_t)((t * CH_TILE_M) / P.rows_per_image);
                for (uint32_t i = 0; i < nL; i++)
                    if (P.layer[i].do_D)
                        for (uint32_t kc = 0; kc < 4; kc++) put(&maps.wgt[i], W_BYTES, (int32_t)(kc * 64), img * 256 + (int32_t)(rank * W_ROWS));
                if (P.has_in)
                    for (uint32_t kc = 0; kc < 4; kc++)
                        put(&maps.wgt_in, in_box_bytes, (int32_t)(kc * 64), (int32_t)((P.in_per_image ? img * (int32_t)P.in_dim : 0) + rank * in_rows));
            }
        }
    } else if (warp == CH_WARP_LOAD) {
        // ===================================================== cos-tile producer (per CTA, local barriers): layer after layer, chunk by chunk
        if (lane == 0) {
            uint32_t cgen = 0;
            const uint64_t stream = l2_policy_evict_first();
            for (uint32_t u = u_begin; u < u_end; u++) {
                const int32_t row0 = (int32_t)((u * CG + rank) * CH_TILE_M);
                for (uint32_t i = 0; i < nL; i++, cgen++)
                    for (uint32_t kc = 0; kc < 4; kc++) {
                        mbar_wait(&S.c_empty[kc], (cgen & 1) ^ 1);
                        mbar_arrive_expect_tx(&S.c_full[kc], CH_CHUNK_BYTES + (kc == 0 ? CH_SGN_TILE_BYTES : 0u));
                        tma_load_2d_hint(smC + kc * CH_CHUNK_BYTES, &maps.c[i], &S.c_full[kc], (int32_t)(kc * 64), row0, stream);
                        // the layer's sign masks arrive with chunk 0 (double-buffered: the previous layer's may still be in use)
                        if (kc == 0) bulk_load(smSGN + (cgen & 1) * CH_SGN_TILE_BYTES, P.layer[i].sgn + (size_t)(u * CG + rank) * CH_SGN_TILE_BYTES, CH_SGN_TILE_BYTES, &S.c_full[0]);
                    }
            }
        }
    } else if (warp == CH_WARP_MMA) {
        // ===================================================== MMA issuer (leader CTA only)
        if (lane == 0 && leader) {
            const uint32_t idesc = idesc_f16(CH_TILE_M * CG, 256, FMT_F16, FMT_F16, 0, 0);
            const uint32_t idesc_in = idesc_f16(CH_TILE_M * CG, P.in_dim, FMT_F16, FMT_F16, 0, 0);
            const uint32_t g_addr = smem_u32(smG);
            uint32_t stage = 0, phase = 0, gev = 0, ng = 0;             // gev: G-write events seen, ng: GEMMs issued
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t accum) {
                if (PAIR) umma_f16_2cta(d, da, db, id, accum); else umma_bf16(d, da, db, id, accum);
            };
            auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_2cta(bar, 3); else umma_commit(bar); };
            auto gemm = [&](uint32_t id) {                              // one GEMM over the 4 chunks of the current G event
                const uint32_t acc = ng & 1;
                mbar_wait(&S.acc_empty[acc], ((ng >> 1) & 1) ^ 1);
                tc_fence_after();
                for (uint32_t kc = 0; kc < 4; kc++) {
                    mbar_wait(&S.g_ready[kc], gev & 1);
                    mbar_wait(&S.w_full[stage], phase);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(smW + stage * W_BYTES), a_addr = g_addr + kc * CH_CHUNK_BYTES;
                    for (uint32_t s = 0; s < 4; s++)
                        mma(tmem_base + acc * 256, smem_desc_sw128(a_addr + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), id, (kc | s) != 0);
                    commit(&S.w_empty[stage]);
                    if (++stage == NW) { stage = 0; phase ^= 1; }
                }
                commit(&S.acc_full[acc]);
                ng++;
                gev++;
            };
            for (uint32_t u = u_begin; u < u_end; u++) {
                for (uint32_t i = 0; i < nL; i++) {
                    if (P.layer[i].do_D) gemm(idesc);
                    else if (!P.has_in) {                               // du of the bottom layer: written for the storer only (with an input
                                                                        // stage it is the A operand of that GEMM, below)
                        for (uint32_t kc = 0; kc < 4; kc++) mbar_wait(&S.g_ready[kc], gev & 1);
                        gev++;
                    }
                }
                if (P.has_in) gemm(idesc_in);
            }
        }
    } else if (warp == CH_WARP_STORE) {
        // ===================================================== storer (STORE): every G event -> HBM (du_l for the weight gradients, dh_0)
        if (STORE && lane == 0) {
            uint32_t gev = 0;
            uint64_t* pend = nullptr;                                   // one store group may still be reading G while the next is issued
            const uint64_t stream = l2_policy_evict_first();
            for (uint32_t u = u_begin; u < u_end; u++) {
                const int32_t row0 = (int32_t)((u * CG + rank) * CH_TILE_M);
                const uint32_t n_ev = 1 + nD;                          // top, then one per D epilogue
                for (uint32_t e = 0; e < n_ev; e++, gev++) {
                    const CUtensorMap* m = e < nL ? &maps.dz[e] : &maps.dh0;
                    for (uint32_t c = 0; c < 4; c++) {
                        mbar_wait(&S.g_ready_st[c], gev & 1);
                        tma_store_2d_hint(m, smG + c * CH_CHUNK_BYTES, (int32_t)(c * 64), row0, stream);
                        tma_store_commit();
                        if (pend) { tma_store_wait_read_pending<1>(); mbar_arrive(pend); }
                        pend = &S.st_done[c];
                    }
                }
            }
            tma_store_wait_read();
            if (pend) mbar_arrive(pend);
            tma_store_wait_all();
        }
    } else if (warp < CH_EPI_WARPS) {
        // ===================================================== epilogue: 16 warps, 4 per TMEM lane quarter, 16 columns of every chunk each
        const uint32_t q = warp & 3, sb = warp >> 2;
        const uint32_t r = q * 32 + lane;
        const uint32_t g_row = smem_u32(smG) + r * 128, c_row = smem_u32(smC) + r * 128;
        const uint32_t u0 = ((2 * sb) ^ (r & 7)) << 4, u1 = ((2 * sb + 1) ^ (r & 7)) << 4;
        const float gs = __ldg(P.gscale), gs_inv = __ldg(P.gscale + 1);
        const uint32_t lane_base = (q * 32) << 16;
        auto arrive_mma = [&](uint64_t* bar) { if (PAIR && !leader) mbar_arrive_remote(bar, 0); else mbar_arrive(bar); };
        uint32_t gev = 0, cgen = 0, ng = 0;
        // write one 16-column piece of a G event: v (fp32) [* cos tile chunk] -> fp16 (saturating) -> G, publish
        auto emit = [&](const float (&v)[16], uint32_t c, bool mul_cos) {
            const uint32_t chunk = g_row + c * CH_CHUNK_BYTES;
            uint32_t hw[8];
            if (mul_cos) {
                mbar_wait(&S.c_full[c], cgen & 1);
                const uint4 a = lds128u(c_row + c * CH_CHUNK_BYTES + u0), b = lds128u(c_row + c * CH_CHUNK_BYTES + u1);
                uint32_t msk;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(msk) : "r"(smem_u32(smSGN) + (((cgen & 1) * 16 + c * 4 + sb) * 128 + r) * 4));
                const uint32_t cw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                // cos = (-1)^bit * sqrt(1 - sin^2): the saved activation is the sine of the same argument.  1 - s^2 as one packed-half fma
                // (s is an fp16 value in [-1, 1], so the result is >= 0 and its rounding is below the error s already carries); the
                // sign flips are applied to the packed fp16 products: mask bit j = element 2j, bit 8 + j = element 2j + 1.
                const uint32_t m2 = __byte_perm(msk, 0, 0x4140);       // sign bits: 0..7 stay, 8..15 -> 16..23
                const uint32_t r2 = __byte_perm(msk, 0, 0x4342);       // rounding bits (high half of the plane word), same arrangement
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const __half2 s2 = *reinterpret_cast<const __half2*>(&cw[k]);
                    // cos^2 = 1 - v^2 with v = the true sine.  The stored fp16 value s is off by up to half an ulp (2^-12 where |s| >= 1/2)
                    // and the rounding bit says to which side: |v| ~ |s| -+ 2^-13, so 1 - v^2 ~ (1 - s^2) +- 2^-12 |s| (+ when |s| was rounded
                    // up).  It only matters where |s| -> 1, so |s| is replaced by a constant (0.8 measured best: rms error of the cosine
                    // 1.72e-3 -> 0.94e-3); where |s| is small the +-2e-4 is below the fp16 rounding of 1 - s^2.
                    const __half2 h1 = __hfma2(__hneg2(s2), s2, __float2half2_rn(1.f));
#if SDFG_RBIT
                    const uint32_t dvb = ((r2 << (15 - k)) & 0x80008000u) ^ 0x8A668A66u;      // +-0.8 * 2^-12 as an fp16 pair
                    // a stored +-1 with the bit clear makes x negative: the |x| operand modifier of the square root reads it as the set bit's value
                    const float2 x = __half22float2(__hadd2(h1, *reinterpret_cast<const __half2*>(&dvb)));
#else
                    const float2 x = __half22float2(h1);
#endif
                    float c0, c1;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(fabsf(x.x)));
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(fabsf(x.y)));
                    hw[k] = pack_f16_sat(v[2 * k] * c0, v[2 * k + 1] * c1) ^ ((m2 << (15 - k)) & 0x80008000u);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.c_empty[c]);              // this warp is done with the cos chunk
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) hw[k] = pack_f16_sat(v[2 * k], v[2 * k + 1]);
            }
            const uint4 h0 = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            const uint4 h1 = make_uint4(hw[4], hw[5], hw[6], hw[7]);
            if (STORE && gev > 0) mbar_wait(&S.st_done[c], (gev - 1) & 1);   // the store of the previous event has read the chunk
            sts128(chunk + u0, h0);
            sts128(chunk + u1, h1);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                arrive_mma(&S.g_ready[c]);
                if (STORE) mbar_arrive(&S.g_ready_st[c]);
            }
        };
        // Per-row head gradients (d_rgb / d_sdf scalars) are fetched ONE UNIT AHEAD: as dependent loads at the start of a unit they
        // exposed a full HBM round trip (~4000 clk, a fifth of the eikonal pass) before the first piece could be computed.
        float rs_n[3] = {0.f, 0.f, 0.f}, ds_n = 0.f;
        uint32_t i_dr = nL;                                            // the D layer whose epilogue adds a rank-1 term (at most one)
        for (uint32_t i = 0; i < nL; i++)
            if (P.layer[i].do_D && P.layer[i].d_rank) { i_dr = i; break; }
        auto fetch_unit = [&](uint32_t u) {
            const uint64_t row_n = (uint64_t)(u * CG + rank) * CH_TILE_M + r;
#pragma unroll
            for (int k = 0; k < 3; k++)
                if ((uint32_t)k < P.top_rank) rs_n[k] = ldg_early(P.top_rank_s + row_n * P.top_rank + k);
            if (i_dr < nL) ds_n = ldg_early(P.layer[i_dr].d_rank_s + row_n);
        };
        if (u_begin < u_end) fetch_unit(u_begin);
        for (uint32_t u = u_begin; u < u_end; u++) {
            const uint32_t t = u * CG + rank;
            const uint64_t row = (uint64_t)t * CH_TILE_M + r;
            // ---------------- top: du_top = (rank terms + d_feat) * c_top
            const float rs[3] = {gs * rs_n[0], gs * rs_n[1], gs * rs_n[2]};
            const float ds_u = gs * ds_n;
            if (u + 1 < u_end) fetch_unit(u + 1);
            {
                const uint32_t rvec_s = smem_u32(&S.vecs[P.top_vec0][0]);
#pragma unroll 1
                for (uint32_t c = 0; c < 4; c++) {
                    const uint32_t col = c * 64 + sb * 16;
                    float dh[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) dh[k] = 0.f;
                    if (P.top_dfeat) {
                        const float4* src = reinterpret_cast<const float4*>(P.top_dfeat + row * 256 + col);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float4 f = ldg_stream4(src + j);
                            dh[4 * j] = gs * f.x; dh[4 * j + 1] = gs * f.y; dh[4 * j + 2] = gs * f.z; dh[4 * j + 3] = gs * f.w;
                        }
                    }
#pragma unroll
                    for (int rr = 0; rr < 3; rr++) {
                        if ((uint32_t)rr < P.top_rank) {
#pragma unroll
                            for (int k = 0; k < 16; k += 4) {
                                const float4 w4 = lds128(rvec_s + (rr * 256 + col + k) * 4);
                                dh[k] = fmaf(rs[rr], w4.x, dh[k]); dh[k + 1] = fmaf(rs[rr], w4.y, dh[k + 1]);
                                dh[k + 2] = fmaf(rs[rr], w4.z, dh[k + 2]); dh[k + 3] = fmaf(rs[rr], w4.w, dh[k + 3]);
                            }
                        }
                    }
                    emit(dh, c, true);
                }
                gev++; cgen++;
            }
            // ---------------- per D GEMM: dh (fp32, TMEM) [+ rank-1] [* cos of the layer below] -> next G event
            for (uint32_t i = 0; i < nL; i++) {
                if (!P.layer[i].do_D) continue;
                const bool last = i + 1 == nL;                          // dh_0: no layer below inside the chain
                const uint32_t d_rank = P.layer[i].d_rank;
                const float ds = d_rank ? (i == i_dr ? ds_u : gs * __ldg(P.layer[i].d_rank_s + row)) : 0.f;
                const uint32_t dvec_s = smem_u32(&S.vecs[P.layer[i].d_vec0][0]);
                const uint32_t acc = ng & 1;
                mbar_wait(&S.acc_full[acc], (ng >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_base + acc * 256 + sb * 16;
                uint32_t raw[2][16];
                tmem_ld16_issue(taddr, raw[0]);
#pragma unroll
                for (uint32_t c = 0; c < 4; c++) {
                    const uint32_t col = c * 64 + sb * 16;
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
                    emit(v, c, !last);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_mma(&S.acc_empty[acc]);
                ng++;
                gev++;
                if (!last) cgen++;
            }
            // ---------------- input stage: d_x_in = gs_inv * acc
            if (P.has_in) {
                const uint32_t acc = ng & 1;
                mbar_wait(&S.acc_full[acc], (ng >> 1) & 1);
                tc_fence_after();
                if (sb * 16 < P.in_dim && P.d_x_in) {                   // warp-uniform: tcgen05.ld is a whole-warp instruction
                    uint32_t raw[16];
                    tmem_ld16(tmem_base + lane_base + acc * 256 + sb * 16, raw);
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
                if (lane == 0) arrive_mma(&S.acc_empty[acc]);
                ng++;
            }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == CH_WARP_MMA) { if (PAIR) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace sdfg
