// Eikonal / input-gradient chain with the GEMMs' A operand in TENSOR MEMORY.  Same math, parameters and tensor maps as
// tc_bchain2.cuh without stores (STORE = false), CTA pairs only.
//
// Why: without stores the fp16 gradient tile G is needed by nobody but the next GEMM, so it does not have to exist in shared memory:
//   * the epilogue converts the fp32 accumulator IN PLACE: it reads its 16 columns of a 64-column chunk (tcgen05.ld), multiplies
//     by cos, packs to fp16 and writes the 8 packed words back over the first 8 of those 16 columns (tcgen05.st) -- exactly the
//     layout tcgen05.mma wants for an A operand in TMEM (lane = row, 32-bit column = 2 consecutive K elements, low half first);
//   * GEMM n reads A from buffer (n+1)&1 -- the accumulator of GEMM n-1, converted -- and accumulates into buffer n&1, whose old
//     contents (the A operand of GEMM n-1) are dead because the tensor pipe executes MMAs in issue order.  Two 256-column
//     buffers = the whole TMEM, no accumulator-empty barrier is needed;
//   * per layer-tile the shared-memory port loses 64 KB of MMA reads and 64 KB of st.shared (of ~456 KB), the fence.proxy.async
//     per piece disappears (tcgen05.wait::st + tcgen05.fence instead), and 64 KB of shared memory are free: the sin ring is two
//     layers deep.
// Status: bit-compatible with tc_bchain2/3 (tests/test_gpu_tc.py runs it as a variant) and measured EQUAL to them (eikonal pass
// 1.81-1.89 ms vs 1.80-1.83 ms for tc_bchain3): neither the shared-memory port nor the sin stream is what bounds these chains
// (scripts/ubench/umma_rate.cu; the producer finds its ring full 88 % of the time).  Kept as the proven recipe for an A operand in
// tensor memory (opt-in: SDFG_TC_TS=1); the default for this pass is tc_bchain3.cuh.
//   shared memory: sin ring 8 x 16 KB | weight ring 5 x 16 KB | sign planes 3 x 4 KB | barriers + head vectors
#pragma once
#include "tc_bchain2.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t B4_NC = 8;                                   // sin-chunk ring slots (two layers deep: covers the loaded HBM latency)
constexpr uint32_t B4_NP = 3;                                   // sign planes

struct B4ChainSmem {
    uint64_t c_full[B4_NC], c_empty[B4_NC];
    uint64_t w_full[BC_MAX_W_STAGES], w_empty[BC_MAX_W_STAGES];
    uint64_t g_ready[4];
    uint64_t acc_full[2];
    uint32_t tmem_base;
    uint32_t pad[3];
    alignas(16) float vecs[4][256];
};

__host__ __device__ inline uint32_t bchain4_smem_bytes() {
    return 1024 + B4_NC * CH_CHUNK_BYTES + bc_w_stages(2) * bc_w_bytes(2) + B4_NP * CH_SGN_TILE_BYTES + (uint32_t)sizeof(B4ChainSmem);
}

__global__ void __launch_bounds__(CH_THREADS, 1)
tc_chain_bwd4_kernel(const __grid_constant__ B2ChainMaps maps, const __grid_constant__ B2ChainParams P) {
    constexpr uint32_t CG = 2;
    constexpr uint32_t W_BYTES = bc_w_bytes(CG), NW = bc_w_stages(CG);
    constexpr uint32_t W_ROWS = 256 / CG;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smC = smem;                                               // sin-chunk ring
    uint8_t* smW = smC + B4_NC * CH_CHUNK_BYTES;
    uint8_t* smSGN = smW + NW * W_BYTES;                               // sign planes, one per event in flight
    B4ChainSmem& S = *reinterpret_cast<B4ChainSmem*>(smSGN + B4_NP * CH_SGN_TILE_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t u_begin = (blockIdx.x / CG) * P.units_per_cta;
    const uint32_t u_end = min(P.n_units, u_begin + P.units_per_cta);
    const uint32_t nL = P.n_layers;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < B4_NC; i++) { mbar_init(&S.c_full[i], 1); mbar_init(&S.c_empty[i], CH_EPI_WARPS); }
        for (uint32_t i = 0; i < 4; i++) mbar_init(&S.g_ready[i], CH_EPI_WARPS * CG);
        for (uint32_t i = 0; i < NW; i++) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (uint32_t i = 0; i < 2; i++) mbar_init(&S.acc_full[i], 1);
        fence_barrier_init();
    }
    if (warp == CH_WARP_TMA && lane == 0) {
        for (uint32_t i = 0; i < nL; i++) {
            tma_prefetch_desc(&maps.c[i]);
            if (P.layer[i].do_D) tma_prefetch_desc(&maps.wgt[i]);
        }
        if (P.has_in) tma_prefetch_desc(&maps.wgt_in);
    }
    if (warp == CH_WARP_MMA) tmem_alloc_2cta(&S.tmem_base, 512);
    for (uint32_t i = threadIdx.x; i < 4 * 256; i += blockDim.x) S.vecs[i >> 8][i & 255] = P.vecs[i >> 8] ? __ldg(P.vecs[i >> 8] + (i & 255)) : 0.f;
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const uint32_t in_rows = P.in_dim / CG, in_box_bytes = in_rows * 128;

    if (warp == CH_WARP_TMA) {
        // ===================================================== weight producer (both CTAs): own half of every chunk, in MMA issue order
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint64_t keep = l2_policy_evict_last();
            auto put = [&](const CUtensorMap* m, uint32_t bytes, int32_t c0, int32_t c1) {
                mbar_wait(&S.w_empty[stage], phase ^ 1);
                if (leader) mbar_arrive_expect_tx(&S.w_full[stage], CG * bytes);
                tma_load_2d_2cta_hint(smW + stage * W_BYTES, m, &S.w_full[stage], c0, c1, keep);
                if (++stage == NW) { stage = 0; phase ^= 1; }
            };
            for (uint32_t u = u_begin; u < u_end; u++) {
                const uint32_t t = u * CG + rank;
                const int32_t img = (int32_t)((t * CH_TILE_M) / P.rows_per_image);
                for (uint32_t i = 0; i < nL; i++)
                    if (P.layer[i].do_D)
                        for (uint32_t kc = 0; kc < 4; kc++) put(&maps.wgt[i], W_BYTES, (int32_t)(kc * 64), img * 256 + (int32_t)(rank * W_ROWS));
                if (P.has_in)
                    for (uint32_t kc = 0; kc < 4; kc++) put(&maps.wgt_in, in_box_bytes, (int32_t)(kc * 64), (int32_t)(rank * in_rows));
            }
        }
    } else if (warp == CH_WARP_LOAD) {
        // ===================================================== sin-chunk producer (per CTA, local barriers): layer after layer, chunk by chunk
        // The ring is two layers deep: with a single 4-chunk buffer the epilogue waited for these loads a quarter of the time at
        // B = 32 (a chunk is needed ~2 us after its slot frees, the loaded HBM latency is of that order).
        if (lane == 0) {
            uint32_t slot = 0, phase = 0, cev = 0;
            const uint64_t stream = l2_policy_evict_first();
#ifdef SDFG_CHAIN_DEBUG
            unsigned long long w_empty_clk = 0, n_wait = 0; const long long tp0 = clock64();
#endif
            for (uint32_t u = u_begin; u < u_end; u++) {
                const uint32_t tile = u * CG + rank;
                for (uint32_t i = 0; i < nL; i++, cev++)
                    for (uint32_t kc = 0; kc < 4; kc++) {
#ifdef SDFG_CHAIN_DEBUG
                        const long long tw = clock64();
                        if (!mbar_try_wait(&S.c_empty[slot], phase ^ 1)) n_wait++;
#endif
                        mbar_wait(&S.c_empty[slot], phase ^ 1);
#ifdef SDFG_CHAIN_DEBUG
                        w_empty_clk += clock64() - tw;
#endif
                        mbar_arrive_expect_tx(&S.c_full[slot], CH_CHUNK_BYTES + (kc == 0 ? CH_SGN_TILE_BYTES : 0u));
                        tma_load_2d_hint(smC + slot * CH_CHUNK_BYTES, &maps.c[i], &S.c_full[slot], (int32_t)(kc * 64), (int32_t)(tile * CH_TILE_M), stream);
                        // the event's sign plane arrives with its chunk 0.  Three planes: plane n % 3 is overwritten with event n + 3's
                        // chunk 0 (ring item 4n + 12), which waits for item 4n + 4 -- (event n + 1, chunk 0) -- to be released.
                        if (kc == 0) bulk_load(smSGN + (cev % B4_NP) * CH_SGN_TILE_BYTES, P.layer[i].sgn + (size_t)tile * CH_SGN_TILE_BYTES, CH_SGN_TILE_BYTES, &S.c_full[slot]);
                        if (++slot == B4_NC) { slot = 0; phase ^= 1; }
                    }
            }
#ifdef SDFG_CHAIN_DEBUG
            if (P.dbg && blockIdx.x == 0) { P.dbg[32] = 100; P.dbg[33] = w_empty_clk; P.dbg[34] = 101; P.dbg[35] = clock64() - tp0; P.dbg[36] = 102; P.dbg[37] = n_wait; }
#endif
        }
    } else if (warp == CH_WARP_MMA) {
        // ===================================================== MMA issuer (leader CTA only)
        if (lane == 0 && leader) {
            const uint32_t idesc = idesc_f16(CH_TILE_M * CG, 256, FMT_F16, FMT_F16, 0, 0);
            const uint32_t idesc_in = idesc_f16(CH_TILE_M * CG, P.in_dim, FMT_F16, FMT_F16, 0, 0);
            uint32_t stage = 0, phase = 0, gev = 0, ng = 0;             // gev: A-operand events seen, ng: GEMMs issued
            auto gemm = [&](uint32_t id) {                              // GEMM ng: A = buffer (ng+1)&1 (converted in place), D = buffer ng&1
                const uint32_t d_addr = tmem_base + (ng & 1) * 256, a_addr = tmem_base + ((ng & 1) ^ 1) * 256;
                for (uint32_t kc = 0; kc < 4; kc++) {
                    mbar_wait(&S.g_ready[kc], gev & 1);
                    mbar_wait(&S.w_full[stage], phase);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(smW + stage * W_BYTES);
                    for (uint32_t k = 0; k < 4; k++)
                        umma_f16_2cta_ts(d_addr, a_addr + kc * 64 + k * 16, smem_desc_sw128(b_addr + k * 32, 16, 1024), id, (kc | k) != 0);
                    umma_commit_2cta(&S.w_empty[stage], 3);
                    if (++stage == NW) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(&S.acc_full[ng & 1], 3);
                ng++;
                gev++;
            };
            for (uint32_t u = u_begin; u < u_end; u++) {
                for (uint32_t i = 0; i < nL; i++) {
                    if (P.layer[i].do_D) gemm(idesc);
                    else {                                              // bottom layer without D: its A event is consumed by nobody
                        for (uint32_t kc = 0; kc < 4; kc++) mbar_wait(&S.g_ready[kc], gev & 1);
                        gev++;
                    }
                }
                if (P.has_in) gemm(idesc_in);
            }
        }
    } else if (warp < CH_EPI_WARPS) {
        // ===================================================== epilogue: 16 warps, 4 per TMEM lane quarter, 16 columns of every chunk each
        const uint32_t q = warp & 3, sb = warp >> 2;
        const uint32_t r = q * 32 + lane;
        const uint32_t u0 = ((2 * sb) ^ (r & 7)) << 4, u1 = ((2 * sb + 1) ^ (r & 7)) << 4;
        const float gs = __ldg(P.gscale), gs_inv = __ldg(P.gscale + 1);
        const uint32_t lane_base = (q * 32) << 16;
        auto arrive_mma = [&](uint64_t* bar) { if (!leader) mbar_arrive_remote(bar, 0); else mbar_arrive(bar); };
        uint32_t cslot = 0, cphase = 0, cev = 0, ng = 0;                // sin ring consumer position; cos events seen
#ifdef SDFG_CHAIN_DEBUG
        uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}; uint32_t tph = (uint32_t)clock();
#define PH4(k) do { const uint32_t now_ = (uint32_t)clock(); ph[k] += now_ - tph; tph = now_; } while (0)
#else
#define PH4(k) do { } while (0)
#endif
        // one 16-column piece of an A-operand event: v (fp32) [* cos piece] -> fp16 (saturating) -> 8 packed columns of TMEM buffer `buf`
        // next_ld != 0: TMEM address of the caller's NEXT accumulator piece, loaded into nxt.  It is issued only after this piece's
        // barrier wait and shared-memory loads: a tcgen05.ld takes ~450 clk while the next layer's MMAs run, and anything that orders
        // behind it (mbarrier.try_wait does) would otherwise expose that latency in every piece; here it hides behind the math.
        auto emit = [&](const float (&v)[16], uint32_t c, bool mul_cos, uint32_t buf, uint32_t next_ld, uint32_t (&nxt)[16]) {
            uint32_t hw[8];
            if (!mul_cos && next_ld) tmem_ld16_issue(next_ld, nxt);
            if (mul_cos) {
                PH4(7);
                mbar_wait(&S.c_full[cslot], cphase);
                PH4(1);
                const uint32_t c_row = smem_u32(smC) + cslot * CH_CHUNK_BYTES + r * 128;
                const uint4 a = lds128u(c_row + u0), b = lds128u(c_row + u1);
                uint32_t msk;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(msk) : "r"(smem_u32(smSGN) + (((cev % B4_NP) * 16 + c * 4 + sb) * 128 + r) * 2));
                if (next_ld) tmem_ld16_issue(next_ld, nxt);
                const uint32_t cw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                const uint32_t m2 = __byte_perm(msk, 0, 0x4140);       // cos = (-1)^bit * sqrt(1 - sin^2), see tc_bchain2.cuh
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const __half2 s2 = *reinterpret_cast<const __half2*>(&cw[k]);
                    const float2 x = __half22float2(__hfma2(__hneg2(s2), s2, __float2half2_rn(1.f)));
                    float c0, c1;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(x.x));
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(x.y));
                    hw[k] = pack_f16_sat(v[2 * k] * c0, v[2 * k + 1] * c1) ^ ((m2 << (15 - k)) & 0x80008000u);
                }
                PH4(2);
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.c_empty[cslot]);          // this warp is done with the sin chunk
                if (++cslot == B4_NC) { cslot = 0; cphase ^= 1; }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) hw[k] = pack_f16_sat(v[2 * k], v[2 * k + 1]);
            }
            tmem_st8(tmem_base + lane_base + buf * 256 + c * 64 + sb * 16, hw);
            tmem_st_wait();
            tc_fence_before();
            PH4(3);
            __syncwarp();
            if (lane == 0) arrive_mma(&S.g_ready[c]);
            PH4(5);
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
            // ---------------- top: du_top = (rank terms + d_feat) * c_top  -> A operand of GEMM ng, i.e. buffer (ng+1)&1
            const float rs[3] = {gs * rs_n[0], gs * rs_n[1], gs * rs_n[2]};
            const float ds_u = gs * ds_n;
            if (u + 1 < u_end) fetch_unit(u + 1);
            {
                const uint32_t rvec_s = smem_u32(&S.vecs[P.top_vec0][0]);
                tc_fence_after();                                       // orders the stores below after the previous unit's accumulator reads
                uint32_t dummy[16];
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
                    emit(dh, c, true, (ng & 1) ^ 1, 0u, dummy);
                }
                cev++;
            }
            // ---------------- per D GEMM: dh (fp32, TMEM) [+ rank-1] [* cos of the layer below] -> converted in place
            for (uint32_t i = 0; i < nL; i++) {
                if (!P.layer[i].do_D) continue;
                const bool last = i + 1 == nL;                          // dh_0: no layer below inside the chain
                const uint32_t d_rank = P.layer[i].d_rank;
                const float ds = d_rank ? (i == i_dr ? ds_u : gs * __ldg(P.layer[i].d_rank_s + row)) : 0.f;
                const uint32_t dvec_s = smem_u32(&S.vecs[P.layer[i].d_vec0][0]);
                const uint32_t acc = ng & 1;
                PH4(4);
                mbar_wait(&S.acc_full[acc], (ng >> 1) & 1);
                tc_fence_after();
                PH4(6);
                const uint32_t taddr = tmem_base + lane_base + acc * 256 + sb * 16;
                uint32_t raw[2][16];
                tmem_ld16_issue(taddr, raw[0]);
#pragma unroll
                for (uint32_t c = 0; c < 4; c++) {
                    const uint32_t col = c * 64 + sb * 16;
                    tmem_ld_wait16(raw[c & 1]);
                    PH4(0);
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
                    emit(v, c, !last, acc, c < 3 ? taddr + (c + 1) * 64 : 0u, raw[(c + 1) & 1]);   // in place: becomes the A operand of GEMM ng + 1
                }
                ng++;
                if (!last) cev++;
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
                ng++;
                PH4(4);
            }
        }
#ifdef SDFG_CHAIN_DEBUG
        if ((threadIdx.x == 0 || threadIdx.x == 9 * 32) && P.dbg && blockIdx.x == 0)
            for (int k = 0; k < 8; k++) { P.dbg[(threadIdx.x ? 1 : 0) * 16 + 2 * k] = k; P.dbg[(threadIdx.x ? 1 : 0) * 16 + 2 * k + 1] = ph[k]; }
#endif
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == CH_WARP_MMA) tmem_dealloc_2cta(tmem_base, 512);
}

}  // namespace tc
}  // namespace sdfg
