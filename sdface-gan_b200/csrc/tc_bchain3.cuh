// Backward / eikonal chain with TWO units in flight per CTA pair (ping-pong).  Same math, parameters and tensor maps as
// tc_bchain2.cuh; what changes is the schedule.  In tc_bchain2 a tile's layers form one dependency chain
//     epilogue(l) -> GEMM(l-1) -> epilogue(l-1) -> ...
// whose links overlap only at 64-column chunk granularity: after a layer's last chunk the epilogue warps wait ~1500 clk for the
// chunk's MMAs, the commit and the first TMEM load, and the tensor pipe idles while the epilogue runs.  Here every CTA owns two
// gradient tiles A and B (2 x 64 KB) with ONE 256-column TMEM accumulator each, and the roles walk the two chains interleaved:
//     epilogue:  E(A,e) E(B,e) E(A,e+1) E(B,e+1) ...        MMA:  G(A,e) G(B,e) G(A,e+1) ...
// While the 16 epilogue warps work on B's event e the tensor pipe runs A's GEMM e, and vice versa: the hand-over latency and the
// GEMM hide behind the other tile's epilogue.  The price is shared memory: the sin tiles stream through a 3 x 16 KB ring instead
// of a 64 KB buffer and the weights through 2 x 16 KB stages (CTA pairs only: each CTA stages half of every weight chunk); both
// are prefetched by dedicated producer threads, and a weight stall of up to (epilogue - GEMM) ~ 1000 clk per layer is free.
//   shared memory: G_A, G_B 128 KB | sin ring 48 KB | weight ring 32 KB | derivative planes (sign + rounding bit) 2 x 8 KB | barriers 0.4 KB
//   (the head vectors are read from global memory: L1 hits after the first tile)
// Measured on B200 (N = 3.1 M): eikonal pass 1.95 -> 1.83 ms, backward with stores 3.00 -> 3.11 ms -- far from the ~1.7x the
// latency picture promised: with the GEMM fully concurrent the epilogue's pieces take longer (~1750 instead of ~1100 clk), so a layer
// still costs about epilogue + GEMM.  It is not the shared-memory port (scripts/ubench/umma_rate.cu: the MMAs keep their nominal
// 128 clk per step next to 70-100 B/clk of ld/st.shared traffic) and not the sin stream (the ring is full 88 % of the time); the
// epilogue's own phases -- SFU-bound math a third, TMEM/mbarrier round trips the rest -- simply stretch when the tensor pipe is
// busy.  The host uses this kernel for the pass without stores only (field_tc.cu).
#pragma once
#include "tc_bchain2.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t B3_NC = 3;                                   // sin-chunk ring slots
constexpr uint32_t B3_NW = 2;                                   // weight ring stages
constexpr uint32_t B3_W_BYTES = 128 * 128;                      // half of a [256 x 64] fp16 weight chunk (the CTA's 128 neurons)

struct B3ChainSmem {
    uint64_t c_full[B3_NC], c_empty[B3_NC];
    uint64_t w_full[B3_NW], w_empty[B3_NW];
    uint64_t g_ready[2][4], g_ready_st[2][4], st_done[2][4];   // [tile slot][chunk]
    uint64_t acc_full[2], acc_empty[2];                         // [tile slot]: one accumulator per tile
    uint32_t tmem_base;
    uint32_t pad[3];
};

__host__ __device__ inline uint32_t bchain3_smem_bytes() {
    return 1024 + 2 * BC_G_BYTES + B3_NC * CH_CHUNK_BYTES + B3_NW * B3_W_BYTES + 2 * CH_SGN_TILE_BYTES + (uint32_t)sizeof(B3ChainSmem);
}

// Launch: clusters of 2 CTAs (cta_group::2), grid = 2 * groups; P.n_units = pairs of tiles, P.units_per_cta units per cluster.
template <bool STORE>
__global__ void __launch_bounds__(CH_THREADS, 1)
tc_chain_bwd3_kernel(const __grid_constant__ B2ChainMaps maps, const __grid_constant__ B2ChainParams P) {
    constexpr uint32_t CG = 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smG = smem;                                               // G_A, G_B: gradient tiles = A operands of the D GEMMs
    uint8_t* smC = smG + 2 * BC_G_BYTES;                               // sin-chunk ring
    uint8_t* smW = smC + B3_NC * CH_CHUNK_BYTES;                       // weight ring
    uint8_t* smSGN = smW + B3_NW * B3_W_BYTES;                         // two sign planes (alternating by cos event)
    B3ChainSmem& S = *reinterpret_cast<B3ChainSmem*>(smSGN + 2 * CH_SGN_TILE_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t u_begin = (blockIdx.x / CG) * P.units_per_cta;
    const uint32_t u_end = min(P.n_units, u_begin + P.units_per_cta);
    const uint32_t n_mine = u_end > u_begin ? u_end - u_begin : 0u;
    const uint32_t n_pairs = (n_mine + 1) / 2;                         // pair p = units (u_begin + 2p, u_begin + 2p + 1); the last may lack B
    const uint32_t nL = P.n_layers;
    uint32_t nD = 0;                                                   // D GEMMs per unit (layers with do_D; only the bottom layer can lack one)
    for (uint32_t i = 0; i < nL; i++) nD += P.layer[i].do_D ? 1u : 0u;
    const uint32_t n_ev = 1 + nD;                                      // G events per unit: top, then one per D epilogue
    const uint32_t n_gemm = nD + (P.has_in ? 1u : 0u);                 // GEMMs per unit = accumulator reads per unit; GEMM g consumes event g

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < B3_NC; i++) { mbar_init(&S.c_full[i], 1); mbar_init(&S.c_empty[i], CH_EPI_WARPS); }
        for (uint32_t i = 0; i < B3_NW; i++) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (uint32_t s = 0; s < 2; s++) {
            for (uint32_t i = 0; i < 4; i++) {
                mbar_init(&S.g_ready[s][i], CH_EPI_WARPS * CG); mbar_init(&S.g_ready_st[s][i], CH_EPI_WARPS); mbar_init(&S.st_done[s][i], 1);
            }
            mbar_init(&S.acc_full[s], 1); mbar_init(&S.acc_empty[s], CH_EPI_WARPS * CG);
        }
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
    if (warp == CH_WARP_MMA) tmem_alloc_2cta(&S.tmem_base, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const uint32_t in_rows = P.in_dim / CG, in_box_bytes = in_rows * 128;
    // g-th D GEMM -> chain layer index (layers without do_D are skipped; only the bottom layer can be one)
    auto d_layer = [&](uint32_t g) -> uint32_t {
        uint32_t n = 0;
        for (uint32_t i = 0; i < nL; i++)
            if (P.layer[i].do_D) { if (n == g) return i; n++; }
        return nL - 1;
    };

    if (warp == CH_WARP_TMA) {
        // ===================================================== weight producer (both CTAs): own half of every chunk, in MMA issue order
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint64_t keep = l2_policy_evict_last();             // per-image weights are re-read for every tile: keep them in L2
            auto put = [&](const CUtensorMap* m, uint32_t bytes, int32_t c0, int32_t c1) {
                mbar_wait(&S.w_empty[stage], phase ^ 1);
                if (leader) mbar_arrive_expect_tx(&S.w_full[stage], CG * bytes);
                tma_load_2d_2cta_hint(smW + stage * B3_W_BYTES, m, &S.w_full[stage], c0, c1, keep);
                if (++stage == B3_NW) { stage = 0; phase ^= 1; }
            };
            for (uint32_t p = 0; p < n_pairs; p++) {
                const uint32_t n_act = (2 * p + 1 < n_mine) ? 2u : 1u;
                for (uint32_t g = 0; g < n_gemm; g++) {
                    const uint32_t li = g < nD ? d_layer(g) : 0u;
                    for (uint32_t s = 0; s < n_act; s++) {
                        const uint32_t t = (u_begin + 2 * p + s) * CG + rank;
                        const int32_t img = (int32_t)((t * CH_TILE_M) / P.rows_per_image);
                        if (g < nD) {
                            for (uint32_t kc = 0; kc < 4; kc++) put(&maps.wgt[li], B3_W_BYTES, (int32_t)(kc * 64), img * 256 + (int32_t)(rank * 128));
                        } else {
                            for (uint32_t kc = 0; kc < 4; kc++)
                                put(&maps.wgt_in, in_box_bytes, (int32_t)(kc * 64), (int32_t)((P.in_per_image ? img * (int32_t)P.in_dim : 0) + rank * in_rows));
                        }
                    }
                }
            }
        }
    } else if (warp == CH_WARP_LOAD) {
        // ===================================================== sin-chunk producer (per CTA, local barriers), in the epilogue's order
        // (an L2 prefetch cursor running 16 chunks ahead of these loads was tried: no gain for the eikonal pass, +14 % time with stores)
        if (lane == 0) {
            uint32_t slot = 0, phase = 0, cev = 0;
            const uint64_t stream = l2_policy_evict_first();
            for (uint32_t p = 0; p < n_pairs; p++) {
                const uint32_t n_act = (2 * p + 1 < n_mine) ? 2u : 1u;
                for (uint32_t e = 0; e < nL; e++)                       // the nL events that multiply by a cos tile
                    for (uint32_t s = 0; s < n_act; s++, cev++) {
                        const uint32_t tile = (u_begin + 2 * p + s) * CG + rank;
                        for (uint32_t kc = 0; kc < 4; kc++) {
                            mbar_wait(&S.c_empty[slot], phase ^ 1);
                            mbar_arrive_expect_tx(&S.c_full[slot], CH_CHUNK_BYTES + (kc == 0 ? CH_SGN_TILE_BYTES : 0u));
                            tma_load_2d_hint(smC + slot * CH_CHUNK_BYTES, &maps.c[e], &S.c_full[slot], (int32_t)(kc * 64), (int32_t)(tile * CH_TILE_M), stream);
                            // the event's sign plane arrives with its chunk 0.  Two planes suffice: before ring item 4n is loaded the
                            // item three places earlier -- (event n-1, chunk 1) -- has been released, i.e. event n-2 is finished.
                            if (kc == 0) bulk_load(smSGN + (cev & 1) * CH_SGN_TILE_BYTES, P.layer[e].sgn + (size_t)tile * CH_SGN_TILE_BYTES, CH_SGN_TILE_BYTES, &S.c_full[slot]);
                            if (++slot == B3_NC) { slot = 0; phase ^= 1; }
                        }
                    }
            }
        }
    } else if (warp == CH_WARP_MMA) {
        // ===================================================== MMA issuer (leader CTA only): G(A,g) G(B,g) G(A,g+1) ...
        if (lane == 0 && leader) {
            const uint32_t idesc = idesc_f16(CH_TILE_M * CG, 256, FMT_F16, FMT_F16, 0, 0);
            const uint32_t idesc_in = idesc_f16(CH_TILE_M * CG, P.in_dim, FMT_F16, FMT_F16, 0, 0);
            uint32_t stage = 0, phase = 0;
            for (uint32_t p = 0; p < n_pairs; p++) {
                const uint32_t n_act = (2 * p + 1 < n_mine) ? 2u : 1u;
                for (uint32_t g = 0; g < n_gemm; g++)
                    for (uint32_t s = 0; s < n_act; s++) {
                        const uint32_t id = g < nD ? idesc : idesc_in;
                        const uint32_t n_acc = p * n_gemm + g;          // this slot's GEMM ordinal = its accumulator generation
                        const uint32_t n_g = p * n_ev + g;              // this slot's G event ordinal
                        mbar_wait(&S.acc_empty[s], (n_acc & 1) ^ 1);    // the epilogue has read the previous result out of this accumulator
                        tc_fence_after();
                        const uint32_t g_addr = smem_u32(smG + s * BC_G_BYTES);
                        for (uint32_t kc = 0; kc < 4; kc++) {
                            mbar_wait(&S.g_ready[s][kc], n_g & 1);
                            mbar_wait(&S.w_full[stage], phase);
                            tc_fence_after();
                            const uint32_t b_addr = smem_u32(smW + stage * B3_W_BYTES), a_addr = g_addr + kc * CH_CHUNK_BYTES;
                            for (uint32_t k = 0; k < 4; k++)
                                umma_f16_2cta(tmem_base + s * 256, smem_desc_sw128(a_addr + k * 32, 16, 1024), smem_desc_sw128(b_addr + k * 32, 16, 1024), id, (kc | k) != 0);
                            umma_commit_2cta(&S.w_empty[stage], 3);
                            if (++stage == B3_NW) { stage = 0; phase ^= 1; }
                        }
                        umma_commit_2cta(&S.acc_full[s], 3);
                    }
            }
        }
    } else if (warp == CH_WARP_STORE) {
        // ===================================================== storer (STORE): every G event -> HBM (du_l for the weight gradients, dh_0)
        if (STORE && lane == 0) {
            uint64_t* pend = nullptr;                                   // one store group may still be reading G while the next is issued
            const uint64_t stream = l2_policy_evict_first();
            for (uint32_t p = 0; p < n_pairs; p++) {
                const uint32_t n_act = (2 * p + 1 < n_mine) ? 2u : 1u;
                for (uint32_t e = 0; e < n_ev; e++)
                    for (uint32_t s = 0; s < n_act; s++) {
                        const int32_t row0 = (int32_t)(((u_begin + 2 * p + s) * CG + rank) * CH_TILE_M);
                        const CUtensorMap* m = e < nL ? &maps.dz[e] : &maps.dh0;
                        const uint32_t n_g = p * n_ev + e;
                        for (uint32_t c = 0; c < 4; c++) {
                            mbar_wait(&S.g_ready_st[s][c], n_g & 1);
                            tma_store_2d_hint(m, smG + s * BC_G_BYTES + c * CH_CHUNK_BYTES, (int32_t)(c * 64), row0, stream);
                            tma_store_commit();
                            if (pend) { tma_store_wait_read_pending<1>(); mbar_arrive(pend); }
                            pend = &S.st_done[s][c];
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
        const uint32_t u0 = ((2 * sb) ^ (r & 7)) << 4, u1 = ((2 * sb + 1) ^ (r & 7)) << 4;
        const float gs = __ldg(P.gscale), gs_inv = __ldg(P.gscale + 1);
        const uint32_t lane_base = (q * 32) << 16;
        auto arrive_mma = [&](uint64_t* bar) { if (!leader) mbar_arrive_remote(bar, 0); else mbar_arrive(bar); };
        uint32_t cslot = 0, cphase = 0, cev = 0;                        // sin ring consumer position; cos events seen (sign plane = cev & 1)
        // write one 16-column piece of G event n_g of tile slot s: v (fp32) [* cos piece] -> fp16 (saturating) -> G_s, publish
        auto emit = [&](const float (&v)[16], uint32_t c, bool mul_cos, uint32_t s, uint32_t n_g) {
            const uint32_t chunk = smem_u32(smG) + s * BC_G_BYTES + r * 128 + c * CH_CHUNK_BYTES;
            uint32_t hw[8];
            if (mul_cos) {
                mbar_wait(&S.c_full[cslot], cphase);
                const uint32_t c_row = smem_u32(smC) + cslot * CH_CHUNK_BYTES + r * 128;
                const uint4 a = lds128u(c_row + u0), b = lds128u(c_row + u1);
                uint32_t msk;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(msk) : "r"(smem_u32(smSGN) + (((cev & 1) * 16 + c * 4 + sb) * 128 + r) * 4));
                const uint32_t cw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                // cos = (-1)^bit * sqrt(1 - sin^2), see tc_bchain2.cuh: packed-half 1 - s^2, sign flips on the packed fp16 products
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
                if (lane == 0) mbar_arrive(&S.c_empty[cslot]);          // this warp is done with the sin chunk
                if (++cslot == B3_NC) { cslot = 0; cphase ^= 1; }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) hw[k] = pack_f16_sat(v[2 * k], v[2 * k + 1]);
            }
            const uint4 h0 = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            const uint4 h1 = make_uint4(hw[4], hw[5], hw[6], hw[7]);
            if (STORE && n_g > 0) mbar_wait(&S.st_done[s][c], (n_g - 1) & 1);   // the store of this slot's previous event has read the chunk
            sts128(chunk + u0, h0);
            sts128(chunk + u1, h1);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                arrive_mma(&S.g_ready[s][c]);
                if (STORE) mbar_arrive(&S.g_ready_st[s][c]);
            }
        };
        // Per-row head gradients (d_rgb / d_sdf scalars) of both tiles are fetched ONE PAIR AHEAD: as dependent loads at the start of
        // a unit they exposed a full HBM round trip before the first piece could be computed (see tc_bchain2.cuh).
        float rA[3] = {0.f, 0.f, 0.f}, rB[3] = {0.f, 0.f, 0.f}, dA = 0.f, dB = 0.f;
        uint32_t i_dr = nL;                                            // the D layer whose epilogue adds a rank-1 term (at most one)
        for (uint32_t i = 0; i < nL; i++)
            if (P.layer[i].do_D && P.layer[i].d_rank) { i_dr = i; break; }
        auto fetch_pair = [&](uint32_t p) {
#pragma unroll
            for (uint32_t s = 0; s < 2; s++) {
                if (2 * p + s >= n_mine) break;
                const uint64_t row_n = (uint64_t)((u_begin + 2 * p + s) * CG + rank) * CH_TILE_M + r;
#pragma unroll
                for (int k = 0; k < 3; k++)
                    if ((uint32_t)k < P.top_rank) (s ? rB[k] : rA[k]) = ldg_early(P.top_rank_s + row_n * P.top_rank + k);
                if (i_dr < nL) (s ? dB : dA) = ldg_early(P.layer[i_dr].d_rank_s + row_n);
            }
        };
        if (n_pairs) fetch_pair(0);
        for (uint32_t p = 0; p < n_pairs; p++) {
            const uint32_t n_act = (2 * p + 1 < n_mine) ? 2u : 1u;
            const float rsA[3] = {gs * rA[0], gs * rA[1], gs * rA[2]}, rsB[3] = {gs * rB[0], gs * rB[1], gs * rB[2]};
            const float dsA = gs * dA, dsB = gs * dB;
            if (p + 1 < n_pairs) fetch_pair(p + 1);
            if (P.eik_out) {
                // the input-stage epilogue of this pair will read 48 dy_dx lines per thread: pull the tiles' 3 * in_dim component rows
                // (512 bytes each) into L2 now, a whole chain ahead
                const uint32_t tid = warp * 32 + lane;
                for (uint32_t s = 0; s < n_act; s++) {
                    const uint64_t row0 = (uint64_t)((u_begin + 2 * p + s) * CG + rank) * CH_TILE_M;
                    for (uint32_t li = tid; li < P.in_dim * 3 * 4; li += CH_EPI_WARPS * 32) {
                        const uint64_t col = row0 + (li & 3) * 32;
                        if (col < P.M_total)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.eik_dydx + (size_t)(li >> 2) * P.M_total + col));
                    }
                }
            }
            for (uint32_t e = 0; e < n_ev; e++)
                for (uint32_t s = 0; s < n_act; s++) {
                    const uint32_t t = (u_begin + 2 * p + s) * CG + rank;
                    const uint64_t row = (uint64_t)t * CH_TILE_M + r;
                    const uint32_t n_g = p * n_ev + e;
                    if (e == 0) {
                        // ---------------- top: du_top = (rank terms + d_feat) * c_top
                        const float rs[3] = {s ? rsB[0] : rsA[0], s ? rsB[1] : rsA[1], s ? rsB[2] : rsA[2]};
                        // head vectors straight from global memory (L1 hits after the first tile): the two gradient tiles, the sin ring and the
                        // derivative planes leave no room for a 4 KB table in shared memory
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
                                        const float4 w4 = __ldg(reinterpret_cast<const float4*>(P.vecs[P.top_vec0 + rr] + col + k));
                                        dh[k] = fmaf(rs[rr], w4.x, dh[k]); dh[k + 1] = fmaf(rs[rr], w4.y, dh[k + 1]);
                                        dh[k + 2] = fmaf(rs[rr], w4.z, dh[k + 2]); dh[k + 3] = fmaf(rs[rr], w4.w, dh[k + 3]);
                                    }
                                }
                            }
                            emit(dh, c, true, s, n_g);
                        }
                        cev++;
                    } else {
                        // ---------------- D epilogue e: dh (fp32, TMEM) [+ rank-1] [* cos of the layer below] -> G event e
                        const uint32_t i = d_layer(e - 1);
                        const bool last = i + 1 == nL;                  // dh_0: no layer below inside the chain
                        const uint32_t d_rank = P.layer[i].d_rank;
                        const float ds = d_rank ? (i == i_dr ? (s ? dsB : dsA) : gs * __ldg(P.layer[i].d_rank_s + row)) : 0.f;
                        const float* dvec = P.vecs[P.layer[i].d_vec0];
                        const uint32_t n_acc = p * n_gemm + (e - 1);
                        mbar_wait(&S.acc_full[s], n_acc & 1);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + lane_base + s * 256 + sb * 16;
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
                                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(dvec + col + k));
                                    v[k] = fmaf(ds, w4.x, v[k]); v[k + 1] = fmaf(ds, w4.y, v[k + 1]);
                                    v[k + 2] = fmaf(ds, w4.z, v[k + 2]); v[k + 3] = fmaf(ds, w4.w, v[k + 3]);
                                }
                            }
                            emit(v, c, !last, s, n_g);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_mma(&S.acc_empty[s]);
                        if (!last) cev++;
                    }
                }
            // ---------------- input stage: d_x_in = gs_inv * acc
            if (P.has_in) {
                for (uint32_t s = 0; s < n_act; s++) {
                    const uint32_t t = (u_begin + 2 * p + s) * CG + rank;
                    const uint64_t row = (uint64_t)t * CH_TILE_M + r;
                    const uint32_t n_acc = p * n_gemm + (n_gemm - 1);
                    mbar_wait(&S.acc_full[s], n_acc & 1);
                    tc_fence_after();
                    if (P.eik_out) {
                        // eikonal pass: d sdf / d point without the [M, in_dim] round trip.  The four warps of a row quarter split the
                        // 32 columns (the host admits in_dim = 32 only): warp sb takes 8 of them = 4 levels x 2 features
                        const uint32_t c0 = (sb & 1) * 16, half = (sb >> 1) * 8;
                        uint32_t raw[16];
                        tmem_ld16(tmem_base + lane_base + s * 256 + c0, raw);
                        tmem_ld_wait();
                        if (row < P.M_total) {
                            // per level the 6 components (d, c) of dy_dx are rows of a component-major matrix: lanes = consecutive
                            // samples, every load a full 128-byte line
                            const float* q = P.eik_dydx + (size_t)((c0 + half) / 2) * 6 * P.M_total + row;
                            float dv[24];
#pragma unroll
                            for (int i = 0; i < 24; i++) dv[i] = __ldg(q + (size_t)i * P.M_total);
                            float e[3] = {0.f, 0.f, 0.f};
#pragma unroll
                            for (int l = 0; l < 4; l++) {
                                const float a0 = __uint_as_float(half ? raw[8 + 2 * l] : raw[2 * l]), a1 = __uint_as_float(half ? raw[9 + 2 * l] : raw[2 * l + 1]);
#pragma unroll
                                for (int d = 0; d < 3; d++) e[d] = fmaf(a1, dv[l * 6 + d * 2 + 1], fmaf(a0, dv[l * 6 + d * 2], e[d]));
                            }
                            const float k = gs_inv * P.eik_scale;
#pragma unroll
                            for (int d = 0; d < 3; d++) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(P.eik_out + row * 3 + d), "f"(e[d] * k) : "memory");
                        }
                    } else if (sb * 16 < P.in_dim && P.d_x_in) {        // warp-uniform: tcgen05.ld is a whole-warp instruction
                        uint32_t raw[16];
                        tmem_ld16(tmem_base + lane_base + s * 256 + sb * 16, raw);
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
                    if (lane == 0) arrive_mma(&S.acc_empty[s]);
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == CH_WARP_MMA) tmem_dealloc_2cta(tmem_base, 512);
}

}  // namespace tc
}  // namespace sdfg
