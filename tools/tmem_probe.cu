// Microbenchmark: tcgen05.ld (TMEM -> registers) cost on sm_100a, per warp and with several warps on one TMEM lane quarter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_probe tools/tmem_probe.cu && /tmp/tmem_probe
// One CTA per SM allocates all 512 TMEM columns; `nwarps` warps (warp w reads lane quarter w % 4) each sweep the 512 columns
// `iters` times with .32x32b.xN loads, either waiting after every load (dep = 1) or after a whole sweep (dep = 0).
// Prints clocks per sweep = per 64 KB read by one lane quarter (so 4 x that figure is one whole 128 x 512 accumulator).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int N>
__global__ void __launch_bounds__(512, 1) probe(int nwarps, int iters, int dep, long long* out, uint32_t* sink) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t = 0;
    if (warp < nwarps) {
        uint32_t r[N];
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int c = 0; c < 512; c += N) {
                tmem_ld<N>(base + c, r);
                if (dep) {
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < N; ++j) acc ^= r[j];
                }
            }
            if (!dep) {
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < N; ++j) acc ^= r[j];
            }
        }
        t = clock64() - t0;
    }
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    if (blockIdx.x == 0 && lane == 0 && warp < nwarps) out[warp] = t / iters;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512) : "memory");
}

int main() {
    long long* out;
    uint32_t* sink;
    cudaMalloc(&out, 16 * sizeof(long long));
    cudaMalloc(&sink, 512 * sizeof(uint32_t));
    long long h[16];
    const int iters = 200;
    printf("clocks per sweep of one lane quarter over 512 columns (64 KB); max over the participating warps\n");
    printf("%-6s %-7s %-5s %10s %12s\n", "shape", "warps", "dep", "clk/sweep", "B/clk/warp");
    for (int shape = 16; shape <= 32; shape *= 2)
        for (int dep = 0; dep <= 1; ++dep)
            for (int nw : {1, 4, 8, 12, 16}) {
                cudaMemset(out, 0, 16 * sizeof(long long));
                if (shape == 16) probe<16><<<148, 512>>>(nw, iters, dep, out, sink);
                else probe<32><<<148, 512>>>(nw, iters, dep, out, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
                printf("x%-5d %-7d %-5d %10lld %12.1f\n", shape, nw, dep, mx, 65536.0 / mx);
            }
    return 0;
}
