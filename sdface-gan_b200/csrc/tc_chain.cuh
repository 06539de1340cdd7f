// The whole field forward as ONE persistent tcgen05 kernel: every layer of a 128-sample tile runs back to back on the SM that
// owns the tile, activations never leave the chip.
//
//   layer 0         K <= 32 "x part" (hash features / raw points)            A = SMALL tile,  B = WSMALL (resident)
//   layers 1..n-1   K = 256 hidden activations                               A = ACT tile,    B = weight chunks streamed from L2
//   last layer      K = 256 + "view part" (SH of the view direction, <= 16)  A = ACT + SMALL, B = streamed + WSMALL
//
// Data flow per tile (ref NGPSIRENGenerator.forward sdf_model.py:1566-1592, SirenGenerator.forward :121-139):
//   loader warp   x_in / view_feat (fp32, HBM) -> fp16 -> SMALL (128B-swizzled K-major operand tile, written by hand)
//   MMA thread    layer i accumulates into TMEM accumulator (i & 1); the K = 256 part is issued 64-column chunk by chunk as soon
//                 as the epilogue of layer i-1 has produced that chunk (act_ready[kc]) -> the MMAs of layer i overlap the epilogue
//                 of layer i-1 at chunk granularity, with only 2 x 256 TMEM columns
//   epilogue      8 warps (2 per TMEM lane quarter): tcgen05.ld -> FiLM + sin.approx (+ sdf / rgb head dot products) -> fp16 ->
//                 st.shared into ACT *in place* (every MMA that read the old contents has completed: acc_full) -> act_ready[kc]
//   TMA producer  streams [256 x 64] fp16 weight chunks (32 KB) of the K = 256 layers through a 3-stage ring, in layer order,
//                 tile after tile (weights live in L2: 0.5 MB per network)
// Algorithmic HBM traffic per sample (inference): in_dim*4 B in, 4 B (sdf) + 12 B (rgb) + 1 KB (features, if wanted) out --
// against 1 KB per sample PER LAYER for the per-layer kernels (tc_layer.cuh), which stay as the fallback for odd shapes.
// When the forward is saved for backward the epilogue also stores each layer's output to HBM (fp16 + bf16 copies, same
// workspace layout as the per-layer path, so the backward kernels are unchanged).
#pragma once
#include "tc_common.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t CH_TILE_M = 128;
constexpr uint32_t CH_CHUNK_BYTES = CH_TILE_M * 128;        // [128 samples x 64 fp16] = 16 KB
constexpr uint32_t CH_ACT_BYTES = 4 * CH_CHUNK_BYTES;       // K = 256
constexpr uint32_t CH_WSMALL_BYTES = 256 * 128;             // [256 neurons x 64 fp16] = 32 KB (4 K-steps of 16)
constexpr uint32_t CH_W_STAGE_BYTES = 256 * 128;            // one streamed weight chunk
constexpr uint32_t CH_W_STAGES = 3;
constexpr uint32_t CH_MAX_LAYERS = SDFG_MAX_FILM + 1;
constexpr uint32_t CH_MAX_MAPS = SDFG_MAX_FILM;
constexpr uint32_t CH_THREADS = 384;                        // warp 0 TMA, 1 MMA, 2 loader, 3 spare, 4..11 epilogue
constexpr uint32_t CH_EPI_WARP0 = 4;
constexpr uint32_t CH_EPI_THREADS = 256;

struct ChainLayer {
    uint32_t has_main;          // K = 256 part: A = ACT, B streamed through the ring with tensor map `tm`
    uint32_t tm;
    uint32_t small_k0, small_nk;   // K-steps [small_k0, small_k0 + small_nk) of SMALL / WSMALL (0 = none)
    uint32_t act;               // 1: FiLM + sin, 0: linear
    uint32_t film;              // row of gamma / beta
    uint32_t to_act;            // write the fp16 output into ACT (input of the next layer)
    uint32_t nh;                // head rows: out_head[row*nh + c] = sum_n h[row,n] * head_w[c*256 + n] + head_b[c]
    const float* bias;          // [256]
    const float* head_w;
    const float* head_b;
    float* out_head;
    uint16_t* out16;            // optional copies of the output in HBM: fp16, bf16, fp32
    int64_t ld_out;
    uint16_t* out16b;
    int64_t ld_out_b;
    float* out_f32;
    int64_t ld_out_f32;
};

struct ChainParams {
    uint32_t M_total, rows_per_image, rows_per_ray, n_tiles, tiles_per_cta, n_layers;
    uint32_t in_dim, view_dim, x_nk, v_nk;     // v_nk = 0: no view part
    const float* x_in;          // [M, in_dim] fp32
    const float* view_feat;     // [M / rows_per_ray, view_dim] fp32
    const float* w_x;           // fp32 [256, in_dim] (pitch ld_wx): layer 0's weights
    int64_t ld_wx;
    const float* w_v;           // fp32 [256, view_dim] (pitch ld_wv): the view columns of the last layer's weights
    int64_t ld_wv;
    const float* gamma;         // + img * gstride + film * 256 + n
    const float* beta;
    int64_t gstride;
    uint16_t* x16;              // save: fp16 / bf16 copies of x, [M, kp_x] zero padded (NULL ok)
    uint16_t* x16b;
    uint32_t kp_x, kp_v;
    uint16_t* v16;              // save: view part expanded per sample, kp_v columns, fp16 / bf16 (NULL ok)
    int64_t ld_v16;
    uint16_t* v16b;
    int64_t ld_v16b;
    unsigned long long* dbg;    // debugging: per-role (tag, clock) event log of CTA 0, 4 x 2048 entries (NULL = off)
    ChainLayer layer[CH_MAX_LAYERS];
};

struct alignas(64) ChainMaps { CUtensorMap m[CH_MAX_MAPS]; };

struct ChainSmem {
    uint64_t w_full[CH_W_STAGES], w_empty[CH_W_STAGES];
    uint64_t act_ready[4];
    uint64_t acc_full[2], acc_empty[2];
    uint64_t x_full, x_free, v_full, v_free;
    uint32_t tmem_base;
    uint32_t pad[3];
    alignas(16) float gam[2][256];      // double-buffered per-layer FiLM constants: gamma, gamma*bias + beta
    float cst[2][256];
    float heads[4][256];                // row 0: first head layer (sdf), rows 1..3: second head layer (rgb)
    float hx[CH_TILE_M][4];             // head partial sums of the second column group
};

__host__ __device__ inline uint32_t chain_smem_bytes() {
    return 1024 + CH_ACT_BYTES + CH_CHUNK_BYTES + CH_WSMALL_BYTES + CH_W_STAGES * CH_W_STAGE_BYTES + (uint32_t)sizeof(ChainSmem);
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

// 8 consecutive fp32 (bounds-checked against n_valid) -> 8 fp16 / bf16 packed in a uint4
__device__ __forceinline__ void load8(const float* src, uint32_t k0, uint32_t n_valid, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (k0 + i < n_valid) ? __ldg(src + k0 + i) : 0.f;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8], uint32_t fmt) {
    return make_uint4(pack16(v[0], v[1], fmt), pack16(v[2], v[3], fmt), pack16(v[4], v[5], fmt), pack16(v[6], v[7], fmt));
}

#define CH_DBG(role, tag)                                                                                  \
    do {                                                                                                    \
        if (P.dbg && blockIdx.x == 0 && dbg_n < 1023) {                                                     \
            P.dbg[(role) * 2048 + 2 * dbg_n] = (tag);                                                       \
            P.dbg[(role) * 2048 + 2 * dbg_n + 1] = clock64();                                               \
            dbg_n++;                                                                                        \
        }                                                                                                   \
    } while (0)

__global__ void __launch_bounds__(CH_THREADS, 1)
tc_chain_fwd_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smACT = smem;
    uint8_t* smSMALL = smACT + CH_ACT_BYTES;
    uint8_t* smWSMALL = smSMALL + CH_CHUNK_BYTES;
    uint8_t* smRING = smWSMALL + CH_WSMALL_BYTES;
    ChainSmem& S = *reinterpret_cast<ChainSmem*>(smRING + CH_W_STAGES * CH_W_STAGE_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t_begin = blockIdx.x * P.tiles_per_cta;
    const uint32_t t_end = min(P.n_tiles, t_begin + P.tiles_per_cta);
    const uint32_t nL = P.n_layers;
    uint32_t dbg_n = 0;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < CH_W_STAGES; i++) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (uint32_t i = 0; i < 4; i++) mbar_init(&S.act_ready[i], CH_EPI_THREADS / 32);
        for (uint32_t i = 0; i < 2; i++) { mbar_init(&S.acc_full[i], 1); mbar_init(&S.acc_empty[i], CH_EPI_THREADS / 32); }
        mbar_init(&S.x_full, 1); mbar_init(&S.x_free, 1); mbar_init(&S.v_full, 1); mbar_init(&S.v_free, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0)
        for (uint32_t i = 0; i < nL; i++)
            if (P.layer[i].has_main) tma_prefetch_desc(&maps.m[P.layer[i].tm]);
    if (warp == 1) tmem_alloc(&S.tmem_base, 512);
    // resident small weights: K-steps [0, x_nk) = layer 0, [x_nk, x_nk + v_nk) = view columns of the last layer; head vectors
    {
        const uint32_t units = 2 * (P.x_nk + P.v_nk);
        for (uint32_t i = threadIdx.x; i < 256 * 8; i += blockDim.x) {
            const uint32_t j = i >> 3, u = i & 7;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (u < 2 * P.x_nk) load8(P.w_x + (int64_t)j * P.ld_wx, u * 8, P.in_dim, v);
            else if (u < units) load8(P.w_v + (int64_t)j * P.ld_wv, (u - 2 * P.x_nk) * 8, P.view_dim, v);
            *reinterpret_cast<uint4*>(smWSMALL + sw128(j, u)) = pack8(v, FMT_F16);
        }
        uint32_t hrow = 0;
        for (uint32_t i = 0; i < nL; i++) {
            const ChainLayer& Ly = P.layer[i];
            for (uint32_t k = threadIdx.x; k < Ly.nh * 256 && hrow + Ly.nh <= 4; k += blockDim.x) S.heads[hrow + k / 256][k % 256] = __ldg(Ly.head_w + k);
            hrow += Ly.nh;
        }
        // zero the activation-side small tile once (its padding columns are never written again)
        for (uint32_t i = threadIdx.x; i < CH_CHUNK_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(smSMALL)[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;

    if (warp == 0) {
        // ===================================================== TMA producer: weight chunks of the K = 256 layers
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = t_begin; t < t_end; t++)
                for (uint32_t i = 0; i < nL; i++) {
                    if (!P.layer[i].has_main) continue;
                    const CUtensorMap* tm = &maps.m[P.layer[i].tm];
                    for (uint32_t kc = 0; kc < 4; kc++) {
                        mbar_wait(&S.w_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&S.w_full[stage], CH_W_STAGE_BYTES);
                        tma_load_2d(smRING + stage * CH_W_STAGE_BYTES, tm, &S.w_full[stage], (int32_t)(kc * 64), 0);
                        CH_DBG(3, i * 16 + kc);
                        if (++stage == CH_W_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = idesc_f16(CH_TILE_M, 256, FMT_F16, FMT_F16, 0, 0);
            const uint32_t a_small = smem_u32(smSMALL), b_small = smem_u32(smWSMALL), a_act = smem_u32(smACT);
            uint32_t stage = 0, phase = 0, n = 0, actgen = 0, it = 0;
            for (uint32_t t = t_begin; t < t_end; t++, it++)
                for (uint32_t i = 0; i < nL; i++, n++) {
                    const ChainLayer& Ly = P.layer[i];
                    const uint32_t acc = n & 1, use = n >> 1;
                    mbar_wait(&S.acc_empty[acc], (use & 1) ^ 1);          // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * 256;
                    uint32_t accumulate = 0;
                    if (Ly.small_nk && i == 0) {                          // x part
                        mbar_wait(&S.x_full, it & 1);
                        tc_fence_after();
                        for (uint32_t s = Ly.small_k0; s < Ly.small_k0 + Ly.small_nk; s++, accumulate = 1)
                            umma_bf16(tmem_d, smem_desc_sw128(a_small + s * 32, 16, 1024), smem_desc_sw128(b_small + s * 32, 16, 1024), idesc, accumulate);
                        umma_commit(&S.x_free);
                    }
                    if (Ly.has_main) {
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
                    if (Ly.small_nk && i != 0) {                          // view part
                        mbar_wait(&S.v_full, it & 1);
                        tc_fence_after();
                        for (uint32_t s = Ly.small_k0; s < Ly.small_k0 + Ly.small_nk; s++, accumulate = 1)
                            umma_bf16(tmem_d, smem_desc_sw128(a_small + s * 32, 16, 1024), smem_desc_sw128(b_small + s * 32, 16, 1024), idesc, accumulate);
                        umma_commit(&S.v_free);
                    }
                    umma_commit(&S.acc_full[acc]);
                }
        }
    } else if (warp == 2) {
        // ===================================================== loader: x / view parts of the tile -> SMALL (+ 16-bit copies in HBM)
        const uint32_t xu = 2 * P.x_nk, vu = 2 * P.v_nk;               // 16-byte units per row
        uint32_t it = 0;
        for (uint32_t t = t_begin; t < t_end; t++, it++) {
            const uint64_t row0 = (uint64_t)t * CH_TILE_M;
            mbar_wait(&S.x_free, (it & 1) ^ 1);
            if (lane == 0) CH_DBG(2, 1);
            for (uint32_t i = lane; i < CH_TILE_M * xu; i += 32) {
                const uint32_t r = i / xu, u = i % xu;
                const uint64_t row = row0 + r;
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (row < P.M_total) load8(P.x_in + row * P.in_dim, u * 8, P.in_dim, v);
                const uint4 h = pack8(v, FMT_F16);
                *reinterpret_cast<uint4*>(smSMALL + sw128(r, u)) = h;
                if (row < P.M_total && u * 8 < P.kp_x) {
                    if (P.x16) *reinterpret_cast<uint4*>(P.x16 + row * P.kp_x + u * 8) = h;
                    if (P.x16b) *reinterpret_cast<uint4*>(P.x16b + row * P.kp_x + u * 8) = pack8(v, FMT_BF16);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&S.x_full); CH_DBG(2, 2); }
            if (vu) {
                mbar_wait(&S.v_free, (it & 1) ^ 1);
                for (uint32_t i = lane; i < CH_TILE_M * vu; i += 32) {
                    const uint32_t r = i / vu, u = i % vu;
                    const uint64_t row = row0 + r;
                    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (row < P.M_total) load8(P.view_feat + (row / P.rows_per_ray) * P.view_dim, u * 8, P.view_dim, v);
                    const uint4 h = pack8(v, FMT_F16);
                    *reinterpret_cast<uint4*>(smSMALL + sw128(r, xu + u)) = h;
                    if (row < P.M_total && u * 8 < P.kp_v) {
                        if (P.v16) *reinterpret_cast<uint4*>(P.v16 + row * P.ld_v16 + u * 8) = h;
                        if (P.v16b) *reinterpret_cast<uint4*>(P.v16b + row * P.ld_v16b + u * 8) = pack8(v, FMT_BF16);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&S.v_full); CH_DBG(2, 3); }
            }
        }
    } else if (warp >= CH_EPI_WARP0) {
        // ===================================================== epilogue: 8 warps, 2 per TMEM lane quarter
        const uint32_t q = warp & 3;                                   // TMEM lane quarter this warp may access
        const uint32_t g = (warp - CH_EPI_WARP0) >> 2;                 // column group: 32-column half of every 64-column chunk
        const uint32_t etid = (warp - CH_EPI_WARP0) * 32 + lane;       // 0..255 = the column this thread prepares constants for
        const uint32_t r = q * 32 + lane;                              // row of the tile = TMEM lane
        uint32_t n = 0;
        for (uint32_t t = t_begin; t < t_end; t++) {
            const uint64_t row = (uint64_t)t * CH_TILE_M + r;
            const bool valid = row < P.M_total;
            const uint32_t img = (uint32_t)(((uint64_t)t * CH_TILE_M) / P.rows_per_image);
            uint32_t hrow = 0;
            for (uint32_t i = 0; i < nL; i++, n++) {
                // layer description -> registers (constant-bank reads with a dynamic index are slow inside the chunk loop)
                const uint32_t L_act = P.layer[i].act, L_nh = P.layer[i].nh, L_to_act = P.layer[i].to_act, L_film = P.layer[i].film;
                uint16_t* const o16 = P.layer[i].out16;
                uint16_t* const o16b = P.layer[i].out16b;
                float* const o32 = P.layer[i].out_f32;
                const int64_t ld16 = P.layer[i].ld_out, ld16b = P.layer[i].ld_out_b, ld32 = P.layer[i].ld_out_f32;
                const uint32_t acc = n & 1, use = n >> 1, tb = n & 1;
                const uint32_t gam_s = smem_u32(&S.gam[tb][0]), cst_s = smem_u32(&S.cst[tb][0]);
                {   // per-layer FiLM constants for column `etid` (overlaps the MMAs of this layer)
                    const float b = __ldg(P.layer[i].bias + etid);
                    float gm = 1.f, cs = b;
                    if (L_act) {
                        gm = __ldg(P.gamma + (int64_t)img * P.gstride + L_film * 256 + etid);
                        cs = fmaf(gm, b, __ldg(P.beta + (int64_t)img * P.gstride + L_film * 256 + etid));
                    }
                    sts32(gam_s + etid * 4, gm);
                    sts32(cst_s + etid * 4, cs);
                    named_bar_sync(1, CH_EPI_THREADS);
                }
                const uint32_t heads_s = smem_u32(&S.heads[hrow][0]);
                float hacc[3] = {0.f, 0.f, 0.f};
                if (etid == 0) CH_DBG(1, 300 + i * 16);
                mbar_wait(&S.acc_full[acc], use & 1);
                tc_fence_after();
                if (etid == 0) CH_DBG(1, 400 + i * 16);
                const uint32_t taddr = tmem_base + ((q * 32) << 16) + acc * 256;
                const uint32_t act_row = smem_u32(smACT) + r * 128;
#pragma unroll 1
                for (uint32_t c = 0; c < 4; c++) {
                    const uint32_t col = c * 64 + g * 32;
                    uint32_t raw[32];
                    tmem_ld32(taddr + col, raw);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int k = 0; k < 32; k += 4) {
                        const float4 g4 = lds128(gam_s + (col + k) * 4);
                        const float4 c4 = lds128(cst_s + (col + k) * 4);
                        v[k] = fmaf(__uint_as_float(raw[k]), g4.x, c4.x);
                        v[k + 1] = fmaf(__uint_as_float(raw[k + 1]), g4.y, c4.y);
                        v[k + 2] = fmaf(__uint_as_float(raw[k + 2]), g4.z, c4.z);
                        v[k + 3] = fmaf(__uint_as_float(raw[k + 3]), g4.w, c4.w);
                    }
                    if (L_act) {
#pragma unroll
                        for (int k = 0; k < 32; k++) v[k] = __sinf(v[k]);
                    }
                    if (L_nh) {
#pragma unroll
                        for (int hd = 0; hd < 3; hd++) {
                            if ((uint32_t)hd < L_nh) {
#pragma unroll
                                for (int k = 0; k < 32; k += 4) {
                                    const float4 w4 = lds128(heads_s + (hd * 256 + col + k) * 4);
                                    hacc[hd] = fmaf(v[k], w4.x, hacc[hd]); hacc[hd] = fmaf(v[k + 1], w4.y, hacc[hd]);
                                    hacc[hd] = fmaf(v[k + 2], w4.z, hacc[hd]); hacc[hd] = fmaf(v[k + 3], w4.w, hacc[hd]);
                                }
                            }
                        }
                    }
                    uint4 h16[4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        h16[j] = make_uint4(pack_f16(v[j * 8], v[j * 8 + 1]), pack_f16(v[j * 8 + 2], v[j * 8 + 3]),
                                            pack_f16(v[j * 8 + 4], v[j * 8 + 5]), pack_f16(v[j * 8 + 6], v[j * 8 + 7]));
                    if (L_to_act) {
                        const uint32_t chunk = act_row + c * CH_CHUNK_BYTES;
#pragma unroll
                        for (int j = 0; j < 4; j++) sts128(chunk + (((g * 4 + j) ^ (r & 7)) << 4), h16[j]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&S.act_ready[c]);
                        if (etid == 0) CH_DBG(1, 500 + i * 16 + c);
                    }
                    if (valid) {
                        if (o16) {
                            uint4* dst = reinterpret_cast<uint4*>(o16 + row * ld16 + col);
#pragma unroll
                            for (int j = 0; j < 4; j++) dst[j] = h16[j];
                        }
                        if (o16b) {
                            uint4* dst = reinterpret_cast<uint4*>(o16b + row * ld16b + col);
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                dst[j] = make_uint4(pack_bf16(v[j * 8], v[j * 8 + 1]), pack_bf16(v[j * 8 + 2], v[j * 8 + 3]),
                                                    pack_bf16(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16(v[j * 8 + 6], v[j * 8 + 7]));
                        }
                        if (o32) {
                            float4* dst = reinterpret_cast<float4*>(o32 + row * ld32 + col);
#pragma unroll
                            for (int j = 0; j < 8; j++) dst[j] = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                        }
                    }
                }
                // every TMEM read of this layer has completed (wait::ld): hand the accumulator back to the MMA thread
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.acc_empty[acc]);
                if (L_nh) {                                             // combine the two column groups' partial dot products
                    if (g == 1) {
#pragma unroll
                        for (int hd = 0; hd < 3; hd++) S.hx[r][hd] = hacc[hd];
                    }
                    named_bar_sync(2, CH_EPI_THREADS);
                    if (g == 0 && valid) {
                        float* oh = P.layer[i].out_head;
                        const float* hb = P.layer[i].head_b;
#pragma unroll
                        for (int hd = 0; hd < 3; hd++)
                            if ((uint32_t)hd < L_nh) oh[row * L_nh + hd] = hacc[hd] + S.hx[r][hd] + __ldg(hb + hd);
                    }
                    hrow += L_nh;
                }
            }
        }
    }
    // teardown: the epilogue consumed the last accumulator, so every MMA and TMA load issued has completed
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace sdfg
