// CTF point-spread application (a-8, train_particles.py:298-302: grouped F.conv2d of y_hat with one (n-1)x(n-1)
// real-space kernel per image, padding (n-1)//2) as a banded-Toeplitz GEMM on the tensor cores.
//
//   mu[b][i][j] = sum_{v,u} ctf[b][v][u] ypad[b][i+v][j+u]          m = n - 1 taps per side, ypad = y_hat padded by c = (m-1)/2
//
// Per image this is  mu^T[j][i] = sum_k A[j][k] B[i][k]  with k = (v, w), w in [0, Wp):
//   B[i][(v,w)] = ypad[b][i+v][w]        rows of the padded fp16 image: a plain 2-D TMA box {64 w, 128 rows} at row v
//   A[j][(v,w)] = ctf[b][v][w - j]       (0 outside [0, m)): every row j is filter row v shifted by j - a Toeplitz
//                                        window the generator warps copy out of a zero-padded fp16 filter held in smem.
// The band structure doubles the MAC count of the direct form (half of A is zeros), which the tensor pipe does not
// notice: 100 images of 128 x 128 take ~0.1 ms where the CUDA-core kernel needed 14 ms.  The adjoint (gradient w.r.t.
// y_hat) is the same correlation with the filter flipped in both axes, so one kernel serves both passes.
//
// Smem filter layout (halves): Z[256 + v*FP + x] = ctf[v][x], x < m, every other element 0, FP >= m + max(Wp - m, n - 1):
// index (w - j) in [-(n-1), Wp-1] relative to a row start lands in that row's values, its trailing zeros or the
// previous row's trailing zeros.  A thread's 32-half run starts at an arbitrary half: it loads the 17 aligned words
// covering it and funnel-shifts by 16 bits when the start is odd.
#pragma once
#include <cuda_fp16.h>

#include "tc_gemm.cuh"

namespace tvae {

struct CtfGeom {
    int B, n, m;        // images, image side, filter side (n - 1)
    int Hp, Wp;         // padded image rows (n + m - 1) and row pitch in halves (multiple of 64)
    int CP;             // global fp16 filter row pitch (halves, multiple of 8, >= m + 1)
    int FP;             // smem filter row pitch (halves, multiple of 8)
    int WQ;             // 64-wide w chunks per filter row (Wp / 64)
};

struct CtfApplyParams {
    CUtensorMap tmB;          // ypad fp16 [B*Hp][Wp], boxes {64, 128 rows}
    int num_stages, num_tiles;
    CtfGeom g;
    const __half* ctf16;      // [B][m][CP] (already flipped for the adjoint), column m.. zero
    float* out;               // (B, n, n) fp32
    const float* acc_scale;   // device scalar multiplied into the result (undoes the input's fp16 scale) or null
};

struct CtfApply : PolicyBase {
    static constexpr const char* kName = "ctf_apply";
    using Params = CtfApplyParams;
    static constexpr int kBN = 128;
    static constexpr bool kAGen = true;
    static constexpr int kProdWarps = 8;
    struct GenState { int b; };
    __device__ static void prefetch_descs(const Params& p) { tma_prefetch_desc(&p.tmB); }
    // extra smem: the zero-padded filter, (256 + m*FP + 256) halves
    __device__ static int filter_halves(const CtfGeom& g) { return 256 + g.m * g.FP + 256; }
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        uint4* z = reinterpret_cast<uint4*>(extra);
        const int n16 = filter_halves(p.g) / 8;
        for (int i = tid; i < n16; i += nthreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // tile = one image (n <= 128: all output rows and columns in one 128 x 128 accumulator)
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        ti.m0 = 0;
        ti.n0 = 0;
        ti.a0 = tile;                       // image
        ti.kc_begin = 0;
        ti.kc_end = p.g.m * p.g.WQ;
    }
    __device__ static constexpr uint32_t tx_bytes() { return kBN * 128; }
    __device__ static void issue_tma(const Params& p, const TileInfo& ti, int kc, uint32_t, uint32_t sb, uint32_t bar) {
        const int v = kc / p.g.WQ, wq = kc - v * p.g.WQ;
        tma_load_2d(sb, &p.tmB, bar, wq * kBKh, ti.a0 * p.g.Hp + v);
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.b = -1; }
    // load this image's filter values into the padded smem layout (the zeros around them are never overwritten)
    __device__ static void gen_tile_begin(const Params& p, const TileInfo& ti, GenState& s, uint8_t* extra, int ptid) {
        if (s.b == ti.a0) return;
        const CtfGeom& g = p.g;
        named_bar_sync(1, kProdWarps * 32);                 // every generator thread is done with the previous image
        __half* z = reinterpret_cast<__half*>(extra) + 256;
        const __half* src = p.ctf16 + (long long)ti.a0 * g.m * g.CP;
        const int per_row = g.CP / 8;                        // uint4 per filter row (covers >= m + 1 halves, tail zero)
        for (int idx = ptid; idx < g.m * per_row; idx += kProdWarps * 32) {
            const int v = idx / per_row, q = idx - v * per_row;
            *reinterpret_cast<uint4*>(z + v * g.FP + q * 8) = __ldg(reinterpret_cast<const uint4*>(src + (long long)v * g.CP) + q);
        }
        named_bar_sync(1, kProdWarps * 32);
        s.b = ti.a0;
    }
    // 256 generator threads: thread = (row j, 32-half run of the 64-wide chunk)
    __device__ static void gen_chunk(const Params& p, const TileInfo&, GenState&, int kc, uint8_t* a_stage, uint8_t* extra, int ptid) {
        const CtfGeom& g = p.g;
        const int j = ptid & (kBM - 1), half = ptid >> 7;
        const int v = kc / g.WQ, wq = kc - v * g.WQ;
        const int a = 256 + v * g.FP + wq * kBKh + half * 32 - j;      // half index of the run's first element (>= 0)
        const uint32_t* zw = reinterpret_cast<const uint32_t*>(extra) + (a >> 1);
        uint32_t w[17];
#pragma unroll
        for (int i = 0; i < 17; ++i) w[i] = zw[i];
        const uint32_t sh = (a & 1) ? 16u : 0u;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint4 q;
            q.x = __funnelshift_r(w[4 * ch], w[4 * ch + 1], sh);
            q.y = __funnelshift_r(w[4 * ch + 1], w[4 * ch + 2], sh);
            q.z = __funnelshift_r(w[4 * ch + 2], w[4 * ch + 3], sh);
            q.w = __funnelshift_r(w[4 * ch + 3], w[4 * ch + 4], sh);
            *reinterpret_cast<uint4*>(a_stage + sw128_offset(j, half * 4 + ch)) = q;
        }
    }
    // accumulator row = output column j, accumulator column = output row i: lanes write consecutive floats
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState&, uint32_t taddr, int row, uint8_t*) {
        const CtfGeom& g = p.g;
        const float sc = p.acc_scale ? __ldg(p.acc_scale) : 1.f;
        float* dst = p.out + (long long)ti.a0 * g.n * g.n + row;
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            if (row >= g.n) continue;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int i = c * 32 + q;
                if (i < g.n) dst[(long long)i * g.n] = __uint_as_float(rr[q]) * sc;
            }
        }
    }
};

// ypad[b][r][w] = fp16(in[b][r - c][w - c] * scale) (0 outside the image), r < Hp, w < Wp.  scale: device scalar or null.
__global__ void __launch_bounds__(256) ctf_pad_input_kernel(const float* __restrict__ in, __half* __restrict__ ypad, CtfGeom g,
                                                            const float* __restrict__ scale) {
    const long long total = (long long)g.B * g.Hp * g.Wp;
    const int c = (g.m - 1) / 2;
    const float sc = scale ? __ldg(scale) : 1.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int w = static_cast<int>(idx % g.Wp);
        const long long br = idx / g.Wp;
        const int r = static_cast<int>(br % g.Hp);
        const long long b = br / g.Hp;
        const int y = r - c, x = w - c;
        float v = 0.f;
        if (y >= 0 && y < g.n && x >= 0 && x < g.n) v = in[(b * g.n + y) * g.n + x] * sc;
        ypad[idx] = __float2half_rn(v);
    }
}

// ctf16[b][v][x] = fp16(ctf[b][v][x]) and flip16[b][v][x] = fp16(ctf[b][m-1-v][m-1-x]); columns x >= m are zero.
__global__ void __launch_bounds__(256) ctf_to_half_kernel(const float* __restrict__ ctf, __half* __restrict__ ctf16,
                                                          __half* __restrict__ flip16, CtfGeom g) {
    const long long total = (long long)g.B * g.m * g.CP;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = static_cast<int>(idx % g.CP);
        const long long bv = idx / g.CP;
        const int v = static_cast<int>(bv % g.m);
        const long long b = bv / g.m;
        float a = 0.f, f = 0.f;
        if (x < g.m) {
            const float* cb = ctf + b * g.m * g.m;
            a = cb[v * g.m + x];
            f = cb[(g.m - 1 - v) * g.m + (g.m - 1 - x)];
        }
        ctf16[idx] = __float2half_rn(a);
        flip16[idx] = __float2half_rn(f);
    }
}

}  // namespace tvae
