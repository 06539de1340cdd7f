// Shared device/host helpers for the sm_100a kernels of the SDF-generator hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/sdfg.h"

namespace sdfg {

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> SDFG_ERR_CUDA; also bumps the launch counter
int sm_count();

// Optional per-kernel timing for bench.py's roofline line: when enabled (sdfg_prof_enable) every launch site wrapped in a
// ProfScope whose tag contains the configured substring is bracketed by CUDA events on its own stream.
struct ProfScope {
    ProfScope(const char* tag, cudaStream_t st);
    ~ProfScope();
    void* rec;
    cudaStream_t st;
};

#define SDFG_REQUIRE(cond, code, ...)                        \
    do {                                                     \
        if (!(cond)) return ::sdfg::set_error(code, __VA_ARGS__); \
    } while (0)

template <typename T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming (read-once) loads / stores that do not pollute L1
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_stream2(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// vectorised no-return float reduction into global memory (sm_90+): one L2 atomic transaction for 2 / 4 floats
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// the same with an L2 eviction-priority policy (createpolicy): the table gradient should outlive traffic streaming past it
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void red_add_v2_hint(float* addr, float a, float b, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(addr), "f"(a), "f"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_v4_hint(float* addr, float a, float b, float c, float d, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

}  // namespace sdfg
