// Microbenchmark: issue rate of tcgen05.mma.kind::f16 for the chain kernels' instruction shape (M = 128, N = 256, K = 16, fp16, fp32
// accumulate, cta_group::1), operands in 128B-swizzled shared memory (SS) or A in tensor memory (TS); one CTA per SM on every SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I sdface-gan_b200/csrc -I include -o umma_rate umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace sdfg::tc;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a),
                 "l"(desc_b), "r"(idesc), "r"(accumulate)
                 : "memory");
}

template <bool TS>
__global__ void __launch_bounds__(512, 1) k(unsigned long long* out, int iters, int n_cols, int traffic) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tmem_base_s, dummy_s;
    __shared__ uint64_t bar;
    const uint32_t warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < (64 + 128) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0h everywhere
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_f16(128, n_cols, FMT_F16, FMT_F16, 0, 0);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            // one 128 x n_cols x 256 layer: 4 chunks x 4 k-steps, A chunk = 16 KB, B chunk = 32 KB (as in the chains)
            for (uint32_t kc = 0; kc < 4; kc++)
                for (uint32_t ks = 0; ks < 4; ks++) {
                    if (TS) umma_ts(tb + (it & 1) * 256, tb + ((it & 1) ^ 1) * 256 + kc * 64 + ks * 16, smem_desc_sw128(b_addr + kc * 32768 + ks * 32, 16, 1024), idesc, (kc | ks) != 0);
                    else umma_bf16(tb + (it & 1) * 256, smem_desc_sw128(a_addr + kc * 16384 + ks * 32, 16, 1024), smem_desc_sw128(b_addr + kc * 32768 + ks * 32, 16, 1024), idesc, (kc | ks) != 0);
                }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = (unsigned long long)(t1 - t0);
    } else if (traffic && warp >= 1) {
        // epilogue-like shared-memory traffic next to the MMAs: every thread streams 16-byte loads and stores over a 32 KB window
        const uint32_t base = smem_u32(smem) + (threadIdx.x & 511) * 16;
        uint32_t acc = 0;
        const long long t0 = clock64();
        long long n = 0;
        while (clock64() - t0 < (long long)iters * 16 * 128) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t a, b, c, d;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + j * 8192));
                acc ^= a ^ b ^ c ^ d;
                if (traffic == 2) asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + 32768 + j * 8192), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
            }
            n += 4;
        }
        if (acc == 0x1234567u) out[1000] = acc;
        if (threadIdx.x == 32) out[256 + blockIdx.x] = (unsigned long long)n;      // 16-byte accesses per thread (x 480 threads)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 1024 * 8);
    const int iters = 400, smem = (64 + 128) * 1024 + 1024;
    cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {148}) for (int n : {256}) for (int ts = 0; ts < 2; ts++) for (int traffic = 0; traffic < 3; traffic++) {
        cudaMemset(d, 0, 1024 * 8);
        if (ts) k<true><<<grid, 512, smem>>>(d, iters, n, traffic); else k<false><<<grid, 512, smem>>>(d, iters, n, traffic);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        unsigned long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
        const double per_instr = (double)mx / (iters * 16), flop = 2.0 * 128 * n * 16;
        unsigned long long acc16 = 0; cudaMemcpy(&acc16, d + 256, 8, cudaMemcpyDeviceToHost);
        const double lsu_bytes_per_clk = (double)acc16 * 480 * 16 * (traffic == 2 ? 2 : 1) / (double)mx;
        printf("grid %3d  N=%3d  %s, co-running %s: %.1f clk per MMA (K=16) -> %.0f flop/clk/SM; LSU shared traffic %.0f B/clk/SM\n", grid, n, ts ? "A in TMEM" : "A in smem",
               traffic == 0 ? "nothing" : traffic == 1 ? "ld.shared" : "ld+st.shared", per_instr, flop / per_instr, traffic ? lsu_bytes_per_clk : 0.0);
    }
    return 0;
}
