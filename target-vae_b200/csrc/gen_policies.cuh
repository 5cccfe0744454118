// Generator layer-1 GEMM policies (a-4 coordinate transform + a-5 random Fourier features + a-6 first
// Linear; models.py:53-58,95-117 and train_mnist.py:222,234-239).
//
// The (B*n^2) x 1024 Fourier feature matrix is never written to HBM: generator warps evaluate
//   x' = (x - dx_b) R(theta_b),  feat[m][f] = cos(x'_0 wx_f + x'_1 wy_f + b_f)      (wx, wy = W / sigma)
// per tile directly into the swizzled A operand, forward and again in the weight-gradient GEMM; the
// coordinate gradient GEMM applies -sin(.) and the projection onto (wx, wy) in its epilogue.
//
// All three run kind::f16: features, weights and the stored activations / gradients are fp16 (11-bit significand,
// the precision class of TF32); gradients carry a power-of-two scale that the epilogues divide out (acc_scale).
//
//   GenL1Fwd   : h1 = LeakyReLU(feat W1^T + b1 + zb[b])                 A generated (K-major), B = W1 (TMA)
//   GenL1Wgrad : dW1[j][f] = sum_m dpre[m][j] feat[m][f]                A = feat^T generated (MN-major), B = dpre (TMA)
//   GenL1Dgrad : dx'[m] = sum_f (-sin(phase) * (dpre W1)[m][f]) (wx_f, wy_f)   A = dpre (TMA), B = W1^T (TMA)
#pragma once
#include <cuda_fp16.h>
#include "linear_policies.cuh"

namespace tvae {

struct CoordXform {
    const float* x;         // base coords (N,2), or explicit per-row coords (M,2) when theta == nullptr
    const float* theta;     // (B) or nullptr
    const float* dx;        // (B,2)
    int N;                  // pixels per image
    long long M;            // B*N rows
};

__device__ __forceinline__ void transformed_coord(const CoordXform& c, long long m, float& x0, float& x1) {
    if (m >= c.M) { x0 = 0.f; x1 = 0.f; return; }
    if (c.theta == nullptr) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(c.x) + m);
        x0 = v.x; x1 = v.y;
        return;
    }
    const int b = static_cast<int>(m / c.N);
    const int px = static_cast<int>(m - (long long)b * c.N);
    const float2 v = __ldg(reinterpret_cast<const float2*>(c.x) + px);
    const float2 d = __ldg(reinterpret_cast<const float2*>(c.dx) + b);
    float sn, cs;
    sincosf(__ldg(c.theta + b), &sn, &cs);
    const float t0 = v.x - d.x, t1 = v.y - d.y;
    // x' = (x - dx) [[cos, sin], [-sin, cos]]   (train_mnist.py:234-239)
    x0 = t0 * cs - t1 * sn;
    x1 = t0 * sn + t1 * cs;
}

// feature table in shared memory: float4 {wx, wy, b, 0} per feature
__device__ __forceinline__ void load_fourier_table(float4* tab, const float* wf_scaled, const float* bf, int E, int tid, int nthreads) {
    for (int f = tid; f < E; f += nthreads)
        tab[f] = make_float4(__ldg(wf_scaled + 2 * f), __ldg(wf_scaled + 2 * f + 1), __ldg(bf + f), 0.f);
}

__device__ __forceinline__ float fourier_phase(const float4& w, float x0, float x1) {
    return fmaf(x1, w.y, fmaf(x0, w.x, w.z));
}

// ------------------------------------------------------------------------------------------------
struct GenL1FwdParams {
    CUtensorMap tmB;          // W1 fp16 [H][E]
    int num_stages, num_tiles, tiles_n, k_chunks;
    CoordXform cx;
    const float* wf_scaled;   // (E,2) = embed_latent.weight / sigma
    const float* bf;          // (E)
    int E, H;
    const float* bias;        // (H)
    const float* zb;          // (B,H) latent_linear(z)
    __half* h1;               // fp16 [M][H]
    int act;                  // kActTanh or LeakyReLU
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <int BN, bool TANH = false>
struct GenL1Fwd : PolicyBase {
    static constexpr const char* kName = "gen_l1_fwd";
    using Params = GenL1FwdParams;
    static constexpr int kBN = BN;
    static constexpr bool kF16 = true;
    static constexpr bool kAGen = true;
    static constexpr int kProdWarps = 8;
    struct GenState { float x0, x1; };
    __device__ static void prefetch_descs(const Params& p) { tma_prefetch_desc(&p.tmB); }
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        load_fourier_table(reinterpret_cast<float4*>(extra), p.wf_scaled, p.bf, p.E, tid, nthreads);
    }
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        const int mt = tile / p.tiles_n, nt = tile - mt * p.tiles_n;
        ti.m0 = mt * kBM;
        ti.n0 = nt * BN;
        ti.kc_begin = 0;
        ti.kc_end = p.k_chunks;
    }
    __device__ static constexpr uint32_t tx_bytes() { return BN * 128; }
    __device__ static void issue_tma(const Params& p, const TileInfo& ti, int kc, uint32_t, uint32_t sb, uint32_t bar) {
        tma_kmajor_h(sb, &p.tmB, bar, kc, ti.n0);
    }
    __device__ static void gen_tile_begin(const Params& p, const TileInfo& ti, GenState& s, uint8_t*, int ptid) {
        transformed_coord(p.cx, (long long)ti.m0 + (ptid & (kBM - 1)), s.x0, s.x1);
    }
    // 256 generator threads: thread = one row, 32 of the chunk's 64 features (4 swizzled 16-byte stores of 8 halves)
    __device__ static void gen_chunk(const Params& p, const TileInfo&, GenState& s, int kc, uint8_t* a_stage, uint8_t* extra, int ptid) {
        const float4* tab = reinterpret_cast<const float4*>(extra);
        const int row = ptid & (kBM - 1), half = ptid >> 7;
        const int f0 = kc * kBKh + half * 32;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float e[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = f0 + ch * 8 + q;
                e[q] = f < p.E ? __cosf(fourier_phase(tab[f], s.x0, s.x1)) : 0.f;
            }
            *reinterpret_cast<uint4*>(a_stage + sw128_offset(row, half * 4 + ch)) =
                make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
        }
    }
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState&, uint32_t taddr, int row, uint8_t*) {
        const long long m = (long long)ti.m0 + row;
        const bool ok = m < p.cx.M;
        const float* zb = (ok && p.zb) ? p.zb + (m / p.cx.N) * p.H : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            const int n0 = ti.n0 + c * 32;
            if (!ok || n0 >= p.H) continue;
            __half* dst = p.h1 + m * p.H + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                float t[8];
#pragma unroll
                for (int q = 0; q < 8; q += 4) {
                    const float4 bz = zb ? __ldg(reinterpret_cast<const float4*>(zb + n0 + j + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j + q));
                    t[q] = __uint_as_float(rr[j + q]) + bb.x + bz.x;
                    t[q + 1] = __uint_as_float(rr[j + q + 1]) + bb.y + bz.y;
                    t[q + 2] = __uint_as_float(rr[j + q + 2]) + bb.z + bz.z;
                    t[q + 3] = __uint_as_float(rr[j + q + 3]) + bb.w + bz.w;
                }
                act_vec<TANH>(t);
                *reinterpret_cast<uint4*>(dst + j) =
                    make_uint4(pack_half2(t[0], t[1]), pack_half2(t[2], t[3]), pack_half2(t[4], t[5]), pack_half2(t[6], t[7]));
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
struct GenL1WgradParams {
    CUtensorMap tmQ;          // dpre fp16 (scaled) [M][H], MN-major boxes
    int num_stages, num_tiles, tiles_m, tiles_n, splits, chunks_total, chunks_per_split;
    CoordXform cx;
    const float* wf_scaled; const float* bf;
    int E, H;
    float* dW1;               // [H][E], zero-filled by the caller
    const float* acc_scale;   // device scalar: 1 / (scale dpre was stored with)
};

template <int BN>
struct GenL1Wgrad : PolicyBase {
    static constexpr const char* kName = "gen_l1_wgrad";
    using Params = GenL1WgradParams;
    static constexpr int kBN = BN;
    static constexpr bool kF16 = true;
    static constexpr bool kAMajorMN = true;
    static constexpr bool kBMajorMN = true;
    static constexpr bool kAGen = true;
    static constexpr int kProdWarps = 8;
    __device__ static void prefetch_descs(const Params& p) { tma_prefetch_desc(&p.tmQ); }
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        load_fourier_table(reinterpret_cast<float4*>(extra), p.wf_scaled, p.bf, p.E, tid, nthreads);
    }
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        const int per_split = p.tiles_m * p.tiles_n;
        const int sp = tile / per_split;
        const int rem = tile - sp * per_split;
        const int mt = rem / p.tiles_n, nt = rem - mt * p.tiles_n;
        ti.m0 = mt * kBM;   // feature f
        ti.n0 = nt * BN;    // hidden j
        ti.kc_begin = sp * p.chunks_per_split;
        ti.kc_end = min(ti.kc_begin + p.chunks_per_split, p.chunks_total);
    }
    __device__ static constexpr uint32_t tx_bytes() { return BN * 128; }
    __device__ static void issue_tma(const Params& p, const TileInfo& ti, int kc, uint32_t, uint32_t sb, uint32_t bar) {
        tma_mnmajor_h(sb, &p.tmQ, bar, ti.n0, kc * kBKh, BN / 32);
    }
    // A = feat^T, MN-major: four 32-feature blocks [64 pixel rows][64 B] in the 64 B swizzle.  256 generator threads:
    // thread = (pixel row of the chunk, feature block), 32 features = 4 x 16-byte stores; the 8 lanes of a store phase
    // write 8 consecutive rows of one granule column = 8 distinct 16-byte slots (conflict-free).
    __device__ static void gen_chunk(const Params& p, const TileInfo& ti, GenState&, int kc, uint8_t* a_stage, uint8_t* extra, int ptid) {
        const float4* tab = reinterpret_cast<const float4*>(extra);
        const int rrow = ptid & 63, fb = ptid >> 6;
        float x0, x1;
        const long long m = (long long)kc * kBKh + rrow;
        transformed_coord(p.cx, m, x0, x1);
        const bool ok = m < p.cx.M;
        const int f0 = ti.m0 + fb * 32;
        uint8_t* blk = a_stage + fb * (kBKh * 64);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float e[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = f0 + ch * 8 + q;
                e[q] = (ok && f < p.E) ? __cosf(fourier_phase(tab[f], x0, x1)) : 0.f;
            }
            *reinterpret_cast<uint4*>(blk + rrow * 64 + ((ch ^ ((rrow >> 1) & 3)) << 4)) =
                make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
        }
    }
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState&, uint32_t taddr, int row, uint8_t*) {
        const int f = ti.m0 + row;
        const bool empty = ti.kc_begin >= ti.kc_end;
        const float acc_scale = __ldg(p.acc_scale);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            if (f >= p.E || empty) continue;
            const int j0 = ti.n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j0 + j < p.H) atomicAdd(p.dW1 + (long long)(j0 + j) * p.E + f, __uint_as_float(rr[j]) * acc_scale);
        }
    }
};

// ------------------------------------------------------------------------------------------------
struct GenL1DgradParams {
    CUtensorMap tmA, tmB;     // dpre fp16 (scaled) [M][H]; W1^T fp16 [E][H]
    int num_stages, num_tiles, tiles_n, k_chunks;
    CoordXform cx;
    const float* wf_scaled; const float* bf;
    int E, H;
    float* dxp;               // [M][2] gradient w.r.t. transformed coords, zero-filled by the caller
    const float* acc_scale;   // device scalar: 1 / (scale dpre was stored with)
};

template <int BN>
struct GenL1Dgrad : PolicyBase {
    static constexpr const char* kName = "gen_l1_dgrad";
    using Params = GenL1DgradParams;
    static constexpr int kBN = BN;
    static constexpr bool kF16 = true;
    __device__ static void prefetch_descs(const Params& p) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        load_fourier_table(reinterpret_cast<float4*>(extra), p.wf_scaled, p.bf, p.E, tid, nthreads);
    }
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        const int mt = tile / p.tiles_n, nt = tile - mt * p.tiles_n;
        ti.m0 = mt * kBM;
        ti.n0 = nt * BN;
        ti.kc_begin = 0;
        ti.kc_end = p.k_chunks;
    }
    __device__ static constexpr uint32_t tx_bytes() { return kAStageBytes + BN * 128; }
    __device__ static void issue_tma(const Params& p, const TileInfo& ti, int kc, uint32_t sa, uint32_t sb, uint32_t bar) {
        tma_kmajor_h(sa, &p.tmA, bar, kc, ti.m0);
        tma_kmajor_h(sb, &p.tmB, bar, kc, ti.n0);
    }
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState&, uint32_t taddr, int row, uint8_t* extra) {
        const float4* tab = reinterpret_cast<const float4*>(extra);
        const float acc_scale = __ldg(p.acc_scale);
        const long long m = (long long)ti.m0 + row;
        float x0, x1;
        transformed_coord(p.cx, m, x0, x1);
        float g0 = 0.f, g1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            const int f0 = ti.n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int f = f0 + j;
                if (f < p.E) {
                    const float4 w = tab[f];
                    const float dph = -__sinf(fourier_phase(w, x0, x1)) * __uint_as_float(rr[j]);
                    g0 = fmaf(dph, w.x, g0);
                    g1 = fmaf(dph, w.y, g1);
                }
            }
        }
        if (m < p.cx.M) {
            atomicAdd(p.dxp + 2 * m, g0 * acc_scale);
            atomicAdd(p.dxp + 2 * m + 1, g1 * acc_scale);
        }
    }
};

}  // namespace tvae
