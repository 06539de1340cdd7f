// Weight-gradient contraction of one layer on the tensor cores, reduced over the SAMPLE axis:
//     G_b[j, k] = sum_{n in image b} dz[n, j] * x[n, k]            j < 256 (output neurons), k < Kx (layer inputs)
//     G_b[j, ones] = sum_{n in image b} dz[n, j]                    (an extra all-ones B operand: the bias / beta gradient)
// Both operands are read exactly as they sit in HBM ([samples, features] row-major): with the sample axis as the MMA's K
// dimension they are MN-major, which tcgen05 consumes directly from the 128B-swizzled TMA boxes (no transposes anywhere).
// dz (gradient, loss-scaled) and x (activation) are both fp16.
//
// Work split: the 256 output neurons do not fit one CTA's TMEM together with the ones column (2 x 272 > 512 columns), so CTAs
// work in PAIRS on the same range of samples, CTA h of a pair owning neurons [128h, 128h+128) -- the partner's reads of x hit L2.
// Per-image results are flushed from TMEM with vector red.global.add into G[b] when the image changes and at the end; a small
// finishing kernel turns G into dW, db, dgamma, dbeta (field_tc.cu).  This contraction is HBM-bound by construction
// (1 KB of operands per 131 kflop per sample): the ring only has to keep loads in flight.
#pragma once
#include "tc_common.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t WG_ROWS = 64;                    // samples per pipeline stage
constexpr uint32_t WG_BOX_BYTES = WG_ROWS * 128;    // one [64 samples x 64 features] box = 8 KB
constexpr uint32_t WG_MAX_STAGES = 8;                // ring depth = as many stages as fit in ~200 KB (3 .. 8, WgradParams::n_ring)
constexpr uint32_t WG_MAX_XBOX = 5;                 // Kx <= 320
constexpr uint32_t WG_THREADS = 192;

struct WgradParams {
    uint32_t n_stage_total;     // N / 64
    uint32_t stages_per_pair;
    uint32_t rows_per_image;    // multiple of 128
    uint32_t Kx;                // layer input width (valid columns of x)
    uint32_t n_xbox;            // ceil(Kx / 64)
    uint32_t n_main;            // MMA N of the first x block: min(round_up(Kx,64), 256)
    uint32_t n_extra;           // 0 or 64: second x block for Kx > 256
    uint32_t ones_col;          // n_main + n_extra: TMEM / G column of the ones accumulator
    uint32_t ldg;               // pitch of G rows (floats) >= ones_col + 16
    uint32_t n_ring;            // pipeline stages resident in shared memory (wgrad_ring_depth)
    uint32_t x_fmt;             // FMT_F16 / FMT_BF16 of x and dz (tcgen05.mma .kind::f16 rejects mixed 16-bit operand types)
    float* G;                   // [B, 256, ldg] fp32, pre-zeroed
};

struct WgradSmem {
    uint64_t full[WG_MAX_STAGES], empty[WG_MAX_STAGES];
    uint64_t acc_full, acc_empty;
    uint32_t tmem_base;
    uint32_t pad;
};

__host__ __device__ inline uint32_t wgrad_stage_bytes(uint32_t n_xbox) { return (2 + n_xbox) * WG_BOX_BYTES; }
// A stage holds 64 samples: 16 KB of gradient tile + 8 KB per 64 input columns.  HBM needs ~50 KB in flight per SM: three stages of the
// K = 32 input-stage contraction (24 KB each) were not enough (4.9 TB/s against 6.5 for the K = 256 layers), so the ring takes what fits.
__host__ __device__ inline uint32_t wgrad_ring_depth(uint32_t n_xbox) {
    const uint32_t fit = (200u * 1024u) / wgrad_stage_bytes(n_xbox);
    return fit < 3 ? 3 : (fit > WG_MAX_STAGES ? WG_MAX_STAGES : fit);
}
__host__ __device__ inline uint32_t wgrad_smem_bytes(uint32_t n_xbox) {
    return 1024 + wgrad_ring_depth(n_xbox) * wgrad_stage_bytes(n_xbox) + 2048 /* ones tile */ + (uint32_t)sizeof(WgradSmem);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ WgradParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t stage_bytes = wgrad_stage_bytes(P.n_xbox);
    uint8_t* ones = smem + P.n_ring * stage_bytes;                  // 16 rows x 128 B of 1.0 (layout-agnostic: all equal)
    WgradSmem& S = *reinterpret_cast<WgradSmem*>(ones + 2048);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t pair = blockIdx.x >> 1, half = blockIdx.x & 1;
    const uint32_t s_begin = pair * P.stages_per_pair;
    const uint32_t s_end = min(P.n_stage_total, s_begin + P.stages_per_pair);
    const uint32_t stages_per_image = P.rows_per_image / WG_ROWS;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < P.n_ring; i++) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 1); }
        mbar_init(&S.acc_full, 1);
        mbar_init(&S.acc_empty, 4);
        fence_barrier_init();
    }
    {   // ones tile, written with generic stores -> make it visible to the async (tensor core) proxy
        const uint16_t one = P.x_fmt == FMT_BF16 ? 0x3F80 : 0x3C00;
        for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<uint16_t*>(ones)[i] = one;
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmDZ); tma_prefetch_desc(&tmX); }
    if (warp == 1) tmem_alloc(&S.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint64_t stream = l2_policy_evict_first();            // du columns are read exactly once (x is shared with the partner CTA)
            for (uint32_t s = s_begin; s < s_end; s++) {
                mbar_wait(&S.empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&S.full[stage], stage_bytes);
                uint8_t* base = smem + stage * stage_bytes;
                const int32_t row = (int32_t)(s * WG_ROWS);
                tma_load_2d_hint(base, &tmDZ, &S.full[stage], (int32_t)(half * 128), row, stream);
                tma_load_2d_hint(base + WG_BOX_BYTES, &tmDZ, &S.full[stage], (int32_t)(half * 128 + 64), row, stream);
                for (uint32_t b = 0; b < P.n_xbox; b++)
                    tma_load_2d(base + (2 + b) * WG_BOX_BYTES, &tmX, &S.full[stage], (int32_t)(b * 64), row);
                if (++stage == P.n_ring) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0 && s_begin < s_end) {
            const uint32_t id_main = idesc_f16(128, P.n_main, P.x_fmt, P.x_fmt, 1, 1);
            const uint32_t id_extra = idesc_f16(128, 64, P.x_fmt, P.x_fmt, 1, 1);
            const uint32_t id_ones = idesc_f16(128, 16, P.x_fmt, P.x_fmt, 1, 0);
            const uint64_t d_ones = smem_desc_sw128(smem_u32(ones), 16, 1024);
            uint32_t stage = 0, phase = 0, flushes = 0;
            bool fresh = true;                                       // next MMA starts a new accumulation
            for (uint32_t s = s_begin; s < s_end; s++) {
                mbar_wait(&S.full[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
                const uint32_t x_addr = a_addr + 2 * WG_BOX_BYTES;
                for (uint32_t k = 0; k < WG_ROWS / 16; k++) {
                    const uint32_t acc = (fresh && k == 0) ? 0u : 1u;
                    const uint64_t da = smem_desc_sw128(a_addr + k * 2048, WG_BOX_BYTES, 1024);
                    umma_bf16(tmem_base, da, smem_desc_sw128(x_addr + k * 2048, WG_BOX_BYTES, 1024), id_main, acc);
                    if (P.n_extra)
                        umma_bf16(tmem_base + P.n_main, da, smem_desc_sw128(x_addr + 4 * WG_BOX_BYTES + k * 2048, WG_BOX_BYTES, 1024), id_extra, acc);
                    umma_bf16(tmem_base + P.ones_col, da, d_ones, id_ones, acc);
                }
                fresh = false;
                umma_commit(&S.empty[stage]);
                if (++stage == P.n_ring) { stage = 0; phase ^= 1; }
                const bool last = s + 1 == s_end;
                if (last || (s + 1) / stages_per_image != s / stages_per_image) {
                    umma_commit(&S.acc_full);                        // image finished: epilogue flushes the accumulators
                    if (!last) {
                        mbar_wait(&S.acc_empty, flushes & 1);
                        tc_fence_after();
                    }
                    flushes++;
                    fresh = true;
                }
            }
        }
    } else {
        // ===================================================== flush warps (TMEM -> red.global.add)
        const uint32_t q = warp & 3;
        const uint32_t j = half * 128 + q * 32 + lane;               // output neuron = TMEM lane
        uint32_t flushes = 0;
        for (uint32_t s = s_begin; s < s_end; s++) {
            const bool last = s + 1 == s_end;
            if (!(last || (s + 1) / stages_per_image != s / stages_per_image)) continue;
            const uint32_t img = s / stages_per_image;
            mbar_wait(&S.acc_full, flushes & 1);
            tc_fence_after();
            float* grow = P.G + ((size_t)img * 256 + j) * P.ldg;
            const uint32_t taddr = tmem_base + ((q * 32) << 16);
            const uint32_t ncols = P.ones_col + 16;
            for (uint32_t c = 0; c < ncols; c += 16) {
                uint32_t raw[16];
                tmem_ld16(taddr + c, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    red_add_v4(grow + c + i, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]), __uint_as_float(raw[i + 2]),
                               __uint_as_float(raw[i + 3]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.acc_empty);
            flushes++;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc
}  // namespace sdfg
