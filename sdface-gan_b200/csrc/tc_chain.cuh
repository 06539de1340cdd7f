// Shared definitions of the fused chain kernels (forward tc_fchain.cuh, backward / eikonal tc_bchain2.cuh, tc_bchain3.cuh): tile
// geometry, warp roles, and the hand-written accesses to 128B-swizzled operand tiles.
//   tile      128 samples (rows = TMEM lanes) x 256 features; operand tiles are K-major, 64 fp16 (128 B) per row and chunk,
//             16-byte unit u of row r stored at u ^ (r & 7) -- what TMA SWIZZLE_128B writes and the UMMA descriptors expect
//   threads   640 = 16 epilogue warps (4 per TMEM lane quarter = 4 per SM sub-partition, 16 columns of every 64-column chunk each)
//             + 4 role warps (TMA producer, MMA issuer, loader, storer) at the highest warp ids (issue priority)
//   derivative plane  sign(cos u) + rounding bit of the saved sine, 2 bits per element: [4 chunks][4 sub-blocks][128 rows] x u32 = 8 KB
#pragma once
#include "tc_common.cuh"

namespace sdfg {
namespace tc {

constexpr uint32_t CH_TILE_M = 128;
constexpr uint32_t CH_CHUNK_BYTES = CH_TILE_M * 128;        // [128 samples x 64 fp16] = 16 KB
constexpr uint32_t CH_ACT_BYTES = 4 * CH_CHUNK_BYTES;       // K = 256
// derivative planes of one tile and layer: [4 chunks][4 sub-blocks][128 rows] x 32 bit.  Per thread and 16-column piece: low half =
// sign(cos u) of its 16 elements, high half = the ROUNDING bit of the saved fp16 sine (1: round-to-nearest went away from zero,
// |stored| > |sin u|) -- it halves the interval the backward's cos = sqrt(1 - sin^2) has to guess in where |sin| -> 1.
// Bit j = element 2j, bit 8 + j = element 2j + 1 in both halves.
constexpr uint32_t CH_SGN_TILE_BYTES = 8192;
constexpr uint32_t CH_AUX_BYTES = 32768;                    // inference: resident small weights; training: 2 derivative-plane tiles
constexpr uint32_t CH_W_STAGE_BYTES = 256 * 128;            // one streamed weight chunk
constexpr uint32_t CH_W_STAGES = 3;
constexpr uint32_t CH_MAX_LAYERS = SDFG_MAX_FILM + 1;
constexpr uint32_t CH_MAX_MAPS = SDFG_MAX_FILM;
constexpr uint32_t CH_EPI_WARPS = 16;
constexpr uint32_t CH_EPI_THREADS = CH_EPI_WARPS * 32;
constexpr uint32_t CH_WARP_TMA = 16, CH_WARP_MMA = 17, CH_WARP_LOAD = 18, CH_WARP_STORE = 19;
constexpr uint32_t CH_THREADS = 640;

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

}  // namespace tc
}  // namespace sdfg
