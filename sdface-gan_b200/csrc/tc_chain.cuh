// The whole field forward as ONE persistent tcgen05 kernel: every layer of a 128-sample tile runs back to back on the SM that
// owns the tile, activations never leave the chip.
//
//   layer 0         K <= 32 "x part" (hash features / raw points)            A = SMALL tile,  B = one streamed chunk
//   layers 1..n-1   K = 256 hidden activations                               A = ACT tile,    B = weight chunks streamed from L2
//   last layer      K = 256 + "view part" (SH of the view direction, <= 16)  A = ACT + SMALL, B = 4 + 1 streamed chunks
//
// Data flow per tile (ref NGPSIRENGenerator.forward sdf_model.py:1566-1592, SirenGenerator.forward :121-139):
//   loader warp   x_in / view_feat (fp32, HBM) -> fp16 -> SMALL (128B-swizzled K-major operand tile, written by hand)
//   MMA thread    layer i accumulates into TMEM accumulator (i & 1); the K = 256 part is issued 64-column chunk by chunk as soon
//                 as the epilogue of layer i-1 has produced that chunk (act_ready[kc]) -> the MMAs of layer i overlap the epilogue
//                 of layer i-1 at chunk granularity, with only 2 x 256 TMEM columns
//   epilogue      16 warps (4 per TMEM lane quarter = 4 per SM sub-partition, 16 columns of every 64-column chunk each):
//                 tcgen05.ld (prefetched one chunk ahead) -> FiLM + sin.approx (+ sdf / rgb head dot products) -> fp16 ->
//                 st.shared into ACT *in place* (every MMA that read the old contents has completed: acc_full) -> act_ready[kc]
//   TMA producer  streams [256 x 64] fp16 weight chunks (32 KB) of the K = 256 layers through a 3-stage ring, in layer order,
//                 tile after tile (weights live in L2: 0.5 MB per network)
//   storer        (SAVE) TMA-stores each finished ACT chunk to the layer's saved-activation matrix in HBM; the epilogue waits for
//                 the store to have read the chunk (st_done[kc]) before it overwrites it one layer later
// The epilogue is the critical path: per layer and SM it has to push 32 768 sin.approx through the SFUs (16 / clk) and read
// 128 KB of TMEM, each ~2 000 clk -- the same as the layer's 16 MMAs (128 clk each).  Role warps sit at the highest warp ids
// because the issue arbiter favours them.
// Algorithmic HBM traffic per sample (inference): in_dim*4 B in, 4 B (sdf) + 12 B (rgb) + 1 KB (features, if wanted) out --
// against 1 KB per sample PER LAYER for the per-layer kernels (tc_layer.cuh), which stay as the fallback for odd shapes.
#pragma once
#include "tc_common.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t CH_TILE_M = 128;
constexpr uint32_t CH_CHUNK_BYTES = CH_TILE_M * 128;        // [128 samples x 64 fp16] = 16 KB
constexpr uint32_t CH_ACT_BYTES = 4 * CH_CHUNK_BYTES;       // K = 256
constexpr uint32_t CH_SGN_TILE_BYTES = 4096;                // sign(cos) bits of one tile and layer: [4 chunks][4 sub-blocks][128 rows] x 16 bit
constexpr uint32_t CH_AUX_BYTES = 32768;                    // inference: resident small weights; training: 2 sign-mask tiles
constexpr uint32_t CH_W_STAGE_BYTES = 256 * 128;            // one streamed weight chunk
constexpr uint32_t CH_W_STAGES = 3;
constexpr uint32_t CH_MAX_LAYERS = SDFG_MAX_FILM + 1;
constexpr uint32_t CH_MAX_MAPS = SDFG_MAX_FILM;
constexpr uint32_t CH_EPI_WARPS = 16;
constexpr uint32_t CH_EPI_THREADS = CH_EPI_WARPS * 32;
constexpr uint32_t CH_WARP_TMA = 16, CH_WARP_MMA = 17, CH_WARP_LOAD = 18, CH_WARP_STORE = 19;
constexpr uint32_t CH_THREADS = 640;

struct ChainLayer {
    uint32_t has_main;          // K = 256 part: A = ACT, B streamed through the ring with tensor map `tm`
    uint32_t tm;
    uint32_t small_k0, small_nk;   // K-steps [small_k0, small_k0 + small_nk) of SMALL; its weights are one extra streamed chunk (tensor map `tm_small`, column c0_small)
    uint32_t tm_small, c0_small;
    uint32_t act;               // 1: FiLM + sin, 0: linear
    uint32_t film;              // row of gamma / beta
    uint32_t to_act;            // write the fp16 output into ACT (input of the next layer and / or source of the TMA store)
    uint32_t store;             // SAVE: TMA-store the output with tensor map stores.m[layer]
    uint32_t nh;                // head rows: out_head[row*nh + c] = sum_n h[row,n] * head_w[c*256 + n] + head_b[c]
    uint32_t pad;
    const float* bias;          // [256]
    const float* head_w;
    const float* head_b;
    float* out_head;
    float* out_f32;             // optional fp32 copy of the output in HBM
    int64_t ld_out_f32;
    uint8_t* sgn;               // COS: sign(cos(gamma u + c)) bit masks, [tiles][CH_SGN_TILE_BYTES] -- with |cos| = sqrt(1 - sin^2) from the saved
                                // activation this is the derivative the backward chain multiplies with (NULL = not wanted)
};

struct ChainParams {
    uint32_t M_total, rows_per_image, rows_per_ray, n_tiles, tiles_per_cta, n_layers;
    uint32_t in_dim, view_dim, x_nk, v_nk;     // v_nk = 0: no view part
    const float* x_in;          // [M, in_dim] fp32
    const float* view_feat;     // [M / rows_per_ray, view_dim] fp32
    const float* w_x;           // inference: fp32 [256, in_dim] (pitch ld_wx), layer 0's weights -> resident small-weight tile
    int64_t ld_wx;
    const float* w_v;           // inference: fp32 [256, view_dim] (pitch ld_wv), the view columns of the last layer's weights
    int64_t ld_wv;
    const float* gamma;         // + img * gstride + film * 256 + n
    const float* beta;
    int64_t gstride;
    uint16_t* x16;              // SAVE: fp16 copy of x, [M, kp_x] zero padded (NULL ok)
    uint32_t kp_x, kp_v;
    uint16_t* v16;              // SAVE: view part expanded per sample, kp_v columns (NULL ok)
    int64_t ld_v16;
    unsigned long long* dbg;    // debugging: per-role (tag, clock) event log of CTA 0, 4 x 2048 entries (NULL = off)
    ChainLayer layer[CH_MAX_LAYERS];
};

struct alignas(64) ChainMaps { CUtensorMap m[CH_MAX_MAPS + 1]; };      // K = 256 layers (+ layer 0's small weight matrix)
struct alignas(64) ChainStoreMaps { CUtensorMap m[CH_MAX_LAYERS]; };

struct ChainSmem {
    uint64_t w_full[CH_W_STAGES], w_empty[CH_W_STAGES];
    uint64_t act_ready[4], fin_ready[4], st_done[4];   // fin_ready: chunks of the LAST layer's output (consumed by the storer only)
    uint64_t acc_full[2], acc_empty[2];
    uint64_t x_full, x_free, v_full, v_free;
    uint32_t tmem_base;
    uint32_t pad[3];
    alignas(16) float gam[2][256];      // double-buffered per-layer FiLM constants: gamma, gamma*bias + beta
    float cst[2][256];
    float heads[4][256];                // row 0: first head layer (sdf), rows 1..3: second head layer (rgb)
    float hbias[4];                     // their biases (a dependent global load per head layer and tile otherwise)
    float hx[3][CH_TILE_M][3];          // head partial sums of column sub-blocks 1..3
    float stg_g[CH_EPI_THREADS];        // cp.async staging of the NEXT layer's raw FiLM inputs: gamma (one slot per epilogue thread),
    float stg_b[256], stg_be[256];      // bias and beta (threads 256..511)
};

__host__ __device__ inline uint32_t chain_smem_bytes() {
    return 1024 + CH_ACT_BYTES + CH_CHUNK_BYTES + CH_W_STAGES * CH_W_STAGE_BYTES + CH_AUX_BYTES + (uint32_t)sizeof(ChainSmem);
}

// byte offset of 16-byte unit u of row r inside a 128B-swizzled tile (what TMA SWIZZLE_128B / the UMMA descriptor expect)
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t u) { return r * 128 + ((u ^ (r & 7)) << 4); }

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// 8 consecutive fp32 (bounds-checked against n_valid) -> registers; fast path when all 8 are in range and 16-byte aligned
__device__ __forceinline__ void load8(const float* src, uint32_t k0, uint32_t n_valid, float (&v)[8]) {
    if (k0 + 8 <= n_valid && ((reinterpret_cast<uintptr_t>(src + k0) & 15) == 0)) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + k0)), b = __ldg(reinterpret_cast<const float4*>(src + k0) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (k0 + i < n_valid) ? __ldg(src + k0 + i) : 0.f;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8], uint32_t fmt) {
    return make_uint4(pack16(v[0], v[1], fmt), pack16(v[2], v[3], fmt), pack16(v[4], v[5], fmt), pack16(v[6], v[7], fmt));
}

// event log for scripts/gpu_dbg.sh; costs instructions in every chunk of the epilogue, so only with -DSDFG_CHAIN_DEBUG
#ifdef SDFG_CHAIN_DEBUG
#define CH_DBG(role, tag)                                                                                  \
    do {                                                                                                    \
        if (P.dbg && blockIdx.x == 0 && dbg_n < 1023) {                                                     \
            P.dbg[(role) * 2048 + 2 * dbg_n] = (tag);                                                       \
            P.dbg[(role) * 2048 + 2 * dbg_n + 1] = clock64();                                               \
            dbg_n++;                                                                                        \
        }                                                                                                   \
    } while (0)
#else
#define CH_DBG(role, tag) do { } while (0)
#endif

// SAVE: a storer thread TMA-stores finished activation chunks (training: every layer; inference: the fp16 features).
// COS (training only): FiLM layers also record sign(cos(gamma u + c)) as one bit per element (4 KB per tile and layer, bulk-stored
// with the activation tile); the backward chain rebuilds cos = +-sqrt(1 - sin^2) from the saved activation.  The aux region
// holds the two mask tiles and the small weights are streamed through the ring instead of being resident.
template <bool SAVE, bool COS>
__global__ void __launch_bounds__(CH_THREADS, 1)
tc_chain_fwd_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainStoreMaps stores, const __grid_constant__ ChainParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smACT = smem;
    uint8_t* smSMALL = smACT + CH_ACT_BYTES;
    uint8_t* smRING = smSMALL + CH_CHUNK_BYTES;
    uint8_t* smAUX = smRING + CH_W_STAGES * CH_W_STAGE_BYTES;   // inference: the RESIDENT small weights (layer 0 + view columns);
    uint8_t* smWSMALL = smAUX;                                  // training (COS): two sign-mask tiles, the small weights are streamed
    uint8_t* smSGN = smAUX;
    constexpr bool RESIDENT = !COS;
    ChainSmem& S = *reinterpret_cast<ChainSmem*>(smAUX + CH_AUX_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t_begin = blockIdx.x * P.tiles_per_cta;
    const uint32_t t_end = min(P.n_tiles, t_begin + P.tiles_per_cta);
    const uint32_t nL = P.n_layers;
    uint32_t dbg_n = 0;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < CH_W_STAGES; i++) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (uint32_t i = 0; i < 4; i++) { mbar_init(&S.act_ready[i], CH_EPI_WARPS); mbar_init(&S.fin_ready[i], CH_EPI_WARPS); mbar_init(&S.st_done[i], 1); }
        for (uint32_t i = 0; i < 2; i++) { mbar_init(&S.acc_full[i], 1); mbar_init(&S.acc_empty[i], CH_EPI_WARPS); }
        mbar_init(&S.x_full, 1); mbar_init(&S.x_free, 1); mbar_init(&S.v_full, 1); mbar_init(&S.v_free, 1);
        fence_barrier_init();
    }
    if (warp == CH_WARP_TMA && lane == 0)
        for (uint32_t i = 0; i < nL; i++) {
            if (P.layer[i].has_main) tma_prefetch_desc(&maps.m[P.layer[i].tm]);
            if (P.layer[i].small_nk) tma_prefetch_desc(&maps.m[P.layer[i].tm_small]);
            if (SAVE && P.layer[i].store) tma_prefetch_desc(&stores.m[i]);
        }
    if (warp == CH_WARP_MMA) tmem_alloc(&S.tmem_base, 512);
    // inference: resident small weights, K-steps [0, x_nk) = layer 0, [x_nk, x_nk + v_nk) = view columns of the last layer
    if (RESIDENT) {
        const uint32_t units = 2 * (P.x_nk + P.v_nk);
        for (uint32_t i = threadIdx.x; i < 256 * 8; i += blockDim.x) {
            const uint32_t j = i >> 3, u = i & 7;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (u < 2 * P.x_nk) load8(P.w_x + (int64_t)j * P.ld_wx, u * 8, P.in_dim, v);
            else if (u < units) load8(P.w_v + (int64_t)j * P.ld_wv, (u - 2 * P.x_nk) * 8, P.view_dim, v);
            *reinterpret_cast<uint4*>(smWSMALL + sw128(j, u)) = pack8(v, FMT_F16);
        }
    }
    // head vectors
    {
        uint32_t hrow = 0;
        for (uint32_t i = 0; i < nL; i++) {
            const uint32_t nh = P.layer[i].nh;
            for (uint32_t k = threadIdx.x; k < nh * 256 && hrow + nh <= 4; k += blockDim.x) S.heads[hrow + k / 256][k % 256] = __ldg(P.layer[i].head_w + k);
            if (threadIdx.x < nh && hrow + nh <= 4) S.hbias[hrow + threadIdx.x] = __ldg(P.layer[i].head_b + threadIdx.x);
            hrow += nh;
        }
        // zero the activation-side small tile once (its padding columns are never written again)
        for (uint32_t i = threadIdx.x; i < CH_CHUNK_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(smSMALL)[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;

    if (warp == CH_WARP_TMA) {
        // ===================================================== TMA producer: weight chunks of the K = 256 layers
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = t_begin; t < t_end; t++)
                for (uint32_t i = 0; i < nL; i++) {
                    // issue order of the MMA thread: layer 0's small chunk first, a view layer's small chunk after its 4 main chunks
                    const uint32_t n_main = P.layer[i].has_main ? 4u : 0u, n_chunks = n_main + ((P.layer[i].small_nk && !RESIDENT) ? 1u : 0u);
                    for (uint32_t k = 0; k < n_chunks; k++) {
                        const bool small = !RESIDENT && P.layer[i].small_nk && (i == 0 ? k == 0 : k == n_main);
                        const uint32_t kc = (i == 0 && P.layer[i].small_nk) ? k - 1 : k;
                        mbar_wait(&S.w_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&S.w_full[stage], CH_W_STAGE_BYTES);
                        if (small) tma_load_2d(smRING + stage * CH_W_STAGE_BYTES, &maps.m[P.layer[i].tm_small], &S.w_full[stage], (int32_t)P.layer[i].c0_small, 0);
                        else tma_load_2d(smRING + stage * CH_W_STAGE_BYTES, &maps.m[P.layer[i].tm], &S.w_full[stage], (int32_t)(kc * 64), 0);
                        CH_DBG(3, i * 16 + k);
                        if (++stage == CH_W_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
        }
    } else if (warp == CH_WARP_MMA) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = idesc_f16(CH_TILE_M, 256, FMT_F16, FMT_F16, 0, 0);
            const uint32_t a_small = smem_u32(smSMALL), a_act = smem_u32(smACT);
            uint32_t stage = 0, phase = 0, n = 0, actgen = 0, it = 0;
            for (uint32_t t = t_begin; t < t_end; t++, it++)
                for (uint32_t i = 0; i < nL; i++, n++) {
                    const uint32_t has_main = P.layer[i].has_main, sk0 = P.layer[i].small_k0, snk = P.layer[i].small_nk;
                    const uint32_t acc = n & 1, use = n >> 1;
                    mbar_wait(&S.acc_empty[acc], (use & 1) ^ 1);          // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * 256;
                    uint32_t accumulate = 0;
                    if (snk && i == 0) {                                  // x part
                        mbar_wait(&S.x_full, it & 1);
                        if (!RESIDENT) mbar_wait(&S.w_full[stage], phase);
                        tc_fence_after();
                        const uint32_t b_addr = RESIDENT ? smem_u32(smWSMALL) + sk0 * 32 : smem_u32(smRING + stage * CH_W_STAGE_BYTES);
                        for (uint32_t s = 0; s < snk; s++, accumulate = 1)
                            umma_bf16(tmem_d, smem_desc_sw128(a_small + (sk0 + s) * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), idesc, accumulate);
                        umma_commit(&S.x_free);
                        if (!RESIDENT) {
                            umma_commit(&S.w_empty[stage]);
                            if (++stage == CH_W_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                    if (has_main) {
                        for (uint32_t kc = 0; kc < 4; kc++) {
                            mbar_wait(&S.act_ready[kc], actgen & 1);      // chunk kc of the previous layer's output is in ACT
                            CH_DBG(0, 100 + i * 16 + kc);
                            mbar_wait(&S.w_full[stage], phase);
                            tc_fence_after();
                            const uint32_t a_addr = a_act + kc * CH_CHUNK_BYTES;
                            const uint32_t b_addr = smem_u32(smRING + stage * CH_W_STAGE_BYTES);
                            for (uint32_t s = 0; s < 4; s++, accumulate = 1)
                                umma_bf16(tmem_d, smem_desc_sw128(a_addr + s * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), idesc, accumulate);
                            umma_commit(&S.w_empty[stage]);
                            CH_DBG(0, 200 + i * 16 + kc);
                            if (++stage == CH_W_STAGES) { stage = 0; phase ^= 1; }
                        }
                        actgen++;
                    }
                    if (snk && i != 0) {                                  // view part
                        mbar_wait(&S.v_full, it & 1);
                        if (!RESIDENT) mbar_wait(&S.w_full[stage], phase);
                        tc_fence_after();
                        const uint32_t b_addr = RESIDENT ? smem_u32(smWSMALL) + sk0 * 32 : smem_u32(smRING + stage * CH_W_STAGE_BYTES);
                        for (uint32_t s = 0; s < snk; s++, accumulate = 1)
                            umma_bf16(tmem_d, smem_desc_sw128(a_small + (sk0 + s) * 32, 16, 1024), smem_desc_sw128(b_addr + s * 32, 16, 1024), idesc, accumulate);
                        umma_commit(&S.v_free);
                        if (!RESIDENT) {
                            umma_commit(&S.w_empty[stage]);
                            if (++stage == CH_W_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                    umma_commit(&S.acc_full[acc]);
                }
        }
    } else if (warp == CH_WARP_LOAD) {
        // ===================================================== loader: x / view parts of the tile -> SMALL (+ fp16 copies in HBM)
        const uint32_t xu = 2 * P.x_nk, vu = 2 * P.v_nk;               // 16-byte units per row
        // fast path: every unit is 8 in-range, 16-byte aligned floats -> the loads of 4 units are issued before any is used
        const bool x_fast = P.in_dim % 8 == 0 && (reinterpret_cast<uintptr_t>(P.x_in) & 15) == 0;
        const bool v_fast = vu && P.view_dim % 8 == 0 && (reinterpret_cast<uintptr_t>(P.view_feat) & 15) == 0;
        auto fill = [&](const float* src, uint32_t src_ld, uint32_t src_div, uint32_t n_valid, bool fast, uint32_t nu, uint32_t u_off,
                        uint32_t row0, uint16_t* copy, uint64_t copy_ld, uint32_t copy_cols) {
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
                        if (SAVE && copy && row < P.M_total && u * 8 < copy_cols) *reinterpret_cast<uint4*>(copy + (uint64_t)row * copy_ld + u * 8) = h;
                    }
                }
            }
        };
        uint32_t it = 0;
        for (uint32_t t = t_begin; t < t_end; t++, it++) {
            const uint32_t row0 = t * CH_TILE_M;
            mbar_wait(&S.x_free, (it & 1) ^ 1);
            if (lane == 0) CH_DBG(2, 1);
            fill(P.x_in, P.in_dim, 1, P.in_dim, x_fast, xu, 0, row0, P.x16, P.kp_x, P.kp_x);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&S.x_full); CH_DBG(2, 2); }
            if (vu) {
                mbar_wait(&S.v_free, (it & 1) ^ 1);
                fill(P.view_feat, P.view_dim, P.rows_per_ray, P.view_dim, v_fast, vu, xu, row0, P.v16, (uint64_t)P.ld_v16, P.kp_v);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&S.v_full); CH_DBG(2, 3); }
            }
        }
    } else if (warp == CH_WARP_STORE) {
        // ===================================================== storer (SAVE): finished ACT chunks -> saved activations in HBM
        if (SAVE && lane == 0) {
            uint32_t actgen = 0, fingen = 0, nn = 0;
            // One bulk group per chunk (the sign-mask tile of a layer rides with its last chunk); one group may still be reading
            // shared memory while the next chunk's store is issued -- a chunk is released for overwriting one iteration behind.
            uint64_t* pend = nullptr;
            for (uint32_t t = t_begin; t < t_end; t++)
                for (uint32_t i = 0; i < nL; i++, nn++) {
                    const bool st = P.layer[i].store != 0, fin = i + 1 == nL;
                    if (!P.layer[i].to_act) continue;
                    for (uint32_t c = 0; c < 4; c++) {
                        if (fin) mbar_wait(&S.fin_ready[c], fingen & 1);
                        else mbar_wait(&S.act_ready[c], actgen & 1);
                        if (st) tma_store_2d(&stores.m[i], smACT + c * CH_CHUNK_BYTES, (int32_t)(c * 64), (int32_t)(t * CH_TILE_M));
                        if (COS && c == 3 && P.layer[i].sgn)              // every warp has written its bits of all 4 chunks
                            bulk_store(P.layer[i].sgn + (size_t)t * CH_SGN_TILE_BYTES, smSGN + (nn & 1) * CH_SGN_TILE_BYTES, CH_SGN_TILE_BYTES);
                        tma_store_commit();
                        if (pend) { tma_store_wait_read_pending<1>(); mbar_arrive(pend); }
                        pend = &S.st_done[c];
                    }
                    if (fin) fingen++; else actgen++;
                }
            tma_store_wait_read();
            if (pend) mbar_arrive(pend);
            tma_store_wait_all();
        }
    } else {
        // ===================================================== epilogue: 16 warps, 4 per TMEM lane quarter
        const uint32_t q = warp & 3;                                   // TMEM lane quarter this warp may access
        const uint32_t sb = warp >> 2;                                 // 16-column sub-block of every 64-column chunk
        const uint32_t etid = threadIdx.x;                             // 0..511
        const uint32_t r = q * 32 + lane;                              // row of the tile = TMEM lane
        const uint32_t act_row = smem_u32(smACT) + r * 128;
        const uint32_t u0 = ((2 * sb) ^ (r & 7)) << 4, u1 = ((2 * sb + 1) ^ (r & 7)) << 4;
        uint32_t n = 0, stgen = 0;
        // FiLM constant of (layer, image) this thread publishes: threads 0..255 gamma (1 for a linear layer), 256..511 gamma*bias + beta.
        // Its raw inputs are fetched one layer AHEAD with cp.async into per-thread staging slots: no register waits on the L2
        // latency, which therefore hides behind the chunk loop instead of sitting between two layers.
        const uint32_t fcol = etid & 255;
        const uint32_t stg_g_s = smem_u32(&S.stg_g[etid]), stg_b_s = smem_u32(&S.stg_b[fcol]), stg_be_s = smem_u32(&S.stg_be[fcol]);
        auto cp_async4 = [](uint32_t dst, const float* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); };
        auto film_fetch = [&](uint32_t i, uint32_t img) {
            if (P.layer[i].act) {
                const int64_t off = (int64_t)img * P.gstride + P.layer[i].film * 256 + fcol;
                cp_async4(stg_g_s, P.gamma + off);
                if (etid >= 256) cp_async4(stg_be_s, P.beta + off);
            }
            if (etid >= 256) cp_async4(stg_b_s, P.layer[i].bias + fcol);
        };
        auto film_value = [&](uint32_t act) -> float {
            asm volatile("cp.async.wait_all;" ::: "memory");
            const float gm = act ? S.stg_g[etid] : 1.f;
            if (etid < 256) return gm;
            const float b = S.stg_b[fcol];
            return act ? fmaf(gm, b, S.stg_be[fcol]) : b;
        };
        if (t_begin < t_end) film_fetch(0, (t_begin * CH_TILE_M) / P.rows_per_image);
#ifdef SDFG_CHAIN_DEBUG
        uint32_t ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; uint32_t tph = (uint32_t)clock();
#define PHF(k) do { const uint32_t now_ = (uint32_t)clock(); ph[k] += now_ - tph; tph = now_; } while (0)
#else
#define PHF(k) do { } while (0)
#endif
        for (uint32_t t = t_begin; t < t_end; t++) {
            const uint64_t row = (uint64_t)t * CH_TILE_M + r;
            const bool valid = row < P.M_total;
            const uint32_t img = (t * CH_TILE_M) / P.rows_per_image;
            uint32_t hrow = 0;
            for (uint32_t i = 0; i < nL; i++, n++) {
                // layer description -> registers (constant-bank reads with a dynamic index are slow inside the chunk loop)
                const uint32_t L_act = P.layer[i].act, L_nh = P.layer[i].nh, L_to_act = P.layer[i].to_act;
                // fp32 copy of this layer's output (legacy feature output): row pointer once per layer, NULL when not wanted
                float* const o32_row = (P.layer[i].out_f32 && valid) ? P.layer[i].out_f32 + row * P.layer[i].ld_out_f32 : nullptr;
                const bool do_sgn = COS && P.layer[i].sgn != nullptr;
                const uint32_t acc = n & 1, use = n >> 1, tb = n & 1;
                const uint32_t gam_s = smem_u32(&S.gam[tb][0]), cst_s = smem_u32(&S.cst[tb][0]);
                PHF(9);
                {   // publish this layer's FiLM constants (prefetched), then fetch the next layer's
                    // (tables of all layers resident per image, without this per-layer barrier, measured 2-5 % SLOWER in training:
                    // the time reappears as waiting for the accumulator, and the warps drift apart)
                    sts32((etid < 256 ? gam_s : cst_s) + fcol * 4, film_value(L_act));
                    named_bar_sync(1, CH_EPI_THREADS);
                    if (i + 1 < nL) film_fetch(i + 1, img);
                    else if (t + 1 < t_end) film_fetch(0, ((t + 1) * CH_TILE_M) / P.rows_per_image);
                }
                PHF(0);
                const uint32_t heads_s = smem_u32(&S.heads[hrow][0]);
                float hacc[3] = {0.f, 0.f, 0.f};
                if (threadIdx.x == 0) CH_DBG(1, 300 + i * 16);
                mbar_wait(&S.acc_full[acc], use & 1);
                tc_fence_after();
                PHF(1);
                if (threadIdx.x == 0) CH_DBG(1, 400 + i * 16);
                const uint32_t taddr = tmem_base + ((q * 32) << 16) + acc * 256 + sb * 16;
                uint32_t raw[2][16];
                tmem_ld16_issue(taddr, raw[0]);
#pragma unroll
                for (uint32_t c = 0; c < 4; c++) {
                    const uint32_t col = c * 64 + sb * 16;
                    tmem_ld_wait16(raw[c & 1]);
                    PHF(2);
                    if (c < 3) tmem_ld16_issue(taddr + (c + 1) * 64, raw[(c + 1) & 1]);
                    float v[16];
#pragma unroll
                    for (int k = 0; k < 16; k += 4) {
                        const float4 g4 = lds128(gam_s + (col + k) * 4);
                        const float4 c4 = lds128(cst_s + (col + k) * 4);
                        v[k] = fmaf(__uint_as_float(raw[c & 1][k]), g4.x, c4.x);
                        v[k + 1] = fmaf(__uint_as_float(raw[c & 1][k + 1]), g4.y, c4.y);
                        v[k + 2] = fmaf(__uint_as_float(raw[c & 1][k + 2]), g4.z, c4.z);
                        v[k + 3] = fmaf(__uint_as_float(raw[c & 1][k + 3]), g4.w, c4.w);
                    }
                    if (L_act) {
                        if (do_sgn) {
                            // sign of cos(u) = parity of rint(u / pi): one fma against the 1.5 * 2^23 magic constant puts that integer
                            // into the low mantissa bits.  Together with |cos| = sqrt(1 - sin^2) from the saved activation this is the
                            // whole derivative -- 16 bits per thread and chunk instead of a second fp16 tile and a second SFU op.
                            // A funnel shift per element moves that bit into the mask: even elements first, then odd ones, so that
                            // bit j = element 2j and bit 8 + j = element 2j + 1 (the order the backward chain's packed-half sign flip wants).
                            uint32_t m = 0;
#pragma unroll
                            for (int k = 0; k < 16; k++) {
                                const int e = k < 8 ? 2 * k : 2 * (k - 8) + 1;
                                m = __funnelshift_r(m, __float_as_uint(fmaf(v[e], 0.31830988618379067f, 12582912.f)), 1);
                            }
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(smem_u32(smSGN) + (((n & 1) * 16 + c * 4 + sb) * 128 + r) * 2), "h"((uint16_t)(m >> 16)) : "memory");
                        }
#pragma unroll
                        for (int k = 0; k < 16; k++) v[k] = __sinf(v[k]);
                    }
                    if (L_nh) {
#pragma unroll
                        for (int hd = 0; hd < 3; hd++) {
                            if ((uint32_t)hd < L_nh) {
#pragma unroll
                                for (int k = 0; k < 16; k += 4) {
                                    const float4 w4 = lds128(heads_s + (hd * 256 + col + k) * 4);
                                    hacc[hd] = fmaf(v[k], w4.x, hacc[hd]); hacc[hd] = fmaf(v[k + 1], w4.y, hacc[hd]);
                                    hacc[hd] = fmaf(v[k + 2], w4.z, hacc[hd]); hacc[hd] = fmaf(v[k + 3], w4.w, hacc[hd]);
                                }
                            }
                        }
                    }
                    PHF(3);
                    if (L_to_act) {
                        const uint4 h0 = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
                        const uint4 h1 = make_uint4(pack_f16(v[8], v[9]), pack_f16(v[10], v[11]), pack_f16(v[12], v[13]), pack_f16(v[14], v[15]));
                        if (SAVE) mbar_wait(&S.st_done[c], (stgen & 1) ^ 1);   // the previous contents of the chunk have been stored
                        PHF(4);
                        const uint32_t chunk = act_row + c * CH_CHUNK_BYTES;
                        sts128(chunk + u0, h0);
                        sts128(chunk + u1, h1);
                        fence_proxy_async();
                        PHF(5);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(i + 1 == nL ? &S.fin_ready[c] : &S.act_ready[c]);
                        PHF(6);
                        if (threadIdx.x == 0) CH_DBG(1, 500 + i * 16 + c);
                        if (lane == 0 && warp != 0) CH_DBG(4 + warp, 500 + i * 16 + c);
                    }
                    if (o32_row) {
                        float4* dst = reinterpret_cast<float4*>(o32_row + col);
#pragma unroll
                        for (int j = 0; j < 4; j++) dst[j] = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                    }
                }
                PHF(7);
                if (L_to_act) stgen++;
                // every TMEM read of this layer has completed (wait::ld): hand the accumulator back to the MMA thread
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.acc_empty[acc]);
                if (L_nh) {                                             // combine the four sub-blocks' partial dot products
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
                    hrow += L_nh;
                }
            }
        }
#ifdef SDFG_CHAIN_DEBUG
        PHF(8);
        if (threadIdx.x == 0 && P.dbg && blockIdx.x == 0)
            for (int k = 0; k < 10; k++) { P.dbg[4 * 2048 + 2 * k] = 1000 + k; P.dbg[4 * 2048 + 2 * k + 1] = ph[k] + 1; }      // role 4 = 4 + warp 0: unused
#endif
    }
    // teardown: the epilogue consumed the last accumulator, so every MMA and TMA load issued has completed
    tc_fence_before();
    __syncthreads();
    if (warp == CH_WARP_MMA) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace sdfg
