// Group-convolution geometry shared by the conv1 policies (conv_f16_policies.cuh), and the Conv2Heads policy.
//
// conv1 is an implicit GEMM  X1[(b,pos), (r,o)] = sum_kk im2col[(b,pos), kk] * bank[(r,o), kk]
// with kk = (c*k + v)*k + u and im2col[(b,pos=(i,j)), kk] = y[b, c, i+v-p, j+u-p] (zero outside).
#pragma once
#include "linear_policies.cuh"

namespace tvae {

struct ConvGeom {
    int B, C, n, k, p, G, O, d, P;   // d = H' = W', P = d*d
    int K;                           // C*k*k
    int kpad;                        // dbank row pitch (multiple of 32, > K: column K collects the bias gradient)
};

// (channel, filter row, filter column) of reduction index kk
struct Im2colCursor {
    int c, v, u;
};
__device__ __forceinline__ Im2colCursor im2col_cursor(int kk, int k) {
    Im2colCursor cur;
    const int kk2 = k * k;
    cur.c = kk / kk2;
    const int rem = kk - cur.c * kk2;
    cur.v = rem / k;
    cur.u = rem - cur.v * k;
    return cur;
}
// Conv2Heads: h = LeakyReLU(x1 W2^T + b2) (Conv3d 1x1x1, models.py:347,356) with the attention / theta / z
// heads (models.py:358,390,392) and the "+p_r", "+offset_r" adds (models.py:382,394-399) fused in the epilogue.
// O <= BN so one accumulator tile holds a full channel vector per row.
constexpr int kMaxNH = 36;
struct Conv2HeadsParams {
    CUtensorMap tmA, tmB;     // x1 fp16 [R][O]; W2 fp16 [O][O]
    int num_stages, num_tiles, k_chunks;
    long long R;              // B*G*P rows
    int O, NH, G, P;
    const float* b2;          // (O)
    const float* wh;          // [NH][O]
    const float* bh;          // [NH]
    const float* head_add;    // [NH][G]
    __half* h;                // fp16 [R][O]
    float* heads;             // (B,NH,G,P)
};

template <int BN, int NHMAX>
struct Conv2Heads : PolicyBase {
    static constexpr const char* kName = "conv2_heads";
    using Params = Conv2HeadsParams;
    static constexpr int kBN = BN;
    static constexpr bool kF16 = true;        // x1 and W2 are fp16 operands (64 k-elements per stage)
    static constexpr int kEpiGroups = 2;      // K = O is 2 stages: the epilogue is the critical path
    __device__ static void prefetch_descs(const Params& p) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        float* s = reinterpret_cast<float*>(extra);   // [NH][O] head weights then [O] b2
        for (int i = tid; i < p.NH * p.O; i += nthreads) s[i] = __ldg(p.wh + i);
        for (int i = tid; i < p.O; i += nthreads) s[p.NH * p.O + i] = __ldg(p.b2 + i);
    }
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        ti.m0 = tile * kBM;
        ti.n0 = 0;
        ti.kc_begin = 0;
        ti.kc_end = p.k_chunks;
    }
    __device__ static constexpr uint32_t tx_bytes() { return kAStageBytes + BN * 128; }
    __device__ static void issue_tma(const Params& p, const TileInfo& ti, int kc, uint32_t sa, uint32_t sb, uint32_t bar) {
        tma_kmajor_h(sa, &p.tmA, bar, kc, ti.m0);
        tma_kmajor_h(sb, &p.tmB, bar, kc, 0);
    }
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState&, uint32_t taddr, int row, uint8_t* extra) {
        const float* s_wh = reinterpret_cast<const float*>(extra);
        const float* s_b2 = s_wh + p.NH * p.O;
        const long long m = (long long)ti.m0 + row;
        const bool ok = m < p.R;
        float t[NHMAX];
#pragma unroll
        for (int j = 0; j < NHMAX; ++j) t[j] = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            const int o0 = c * 32;
            if (!ok || o0 >= p.O) continue;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = lrelu(__uint_as_float(rr[j]) + s_b2[o0 + j]);
            __half* dst = p.h + m * p.O + o0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 q;
                __half2 hv;
                hv = __floats2half2_rn(v[j], v[j + 1]);     q.x = *reinterpret_cast<uint32_t*>(&hv);
                hv = __floats2half2_rn(v[j + 2], v[j + 3]); q.y = *reinterpret_cast<uint32_t*>(&hv);
                hv = __floats2half2_rn(v[j + 4], v[j + 5]); q.z = *reinterpret_cast<uint32_t*>(&hv);
                hv = __floats2half2_rn(v[j + 6], v[j + 7]); q.w = *reinterpret_cast<uint32_t*>(&hv);
                *reinterpret_cast<uint4*>(dst + j) = q;
            }
#pragma unroll
            for (int hh = 0; hh < NHMAX; ++hh) {
                if (hh < p.NH) {
                    const float4* w4 = reinterpret_cast<const float4*>(s_wh + hh * p.O + o0);
                    float acc = t[hh];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 w = w4[j];
                        acc = fmaf(v[4 * j], w.x, acc);
                        acc = fmaf(v[4 * j + 1], w.y, acc);
                        acc = fmaf(v[4 * j + 2], w.z, acc);
                        acc = fmaf(v[4 * j + 3], w.w, acc);
                    }
                    t[hh] = acc;
                }
            }
        }
        if (ok) {
            const long long br = m / p.P;            // b*G + r
            const int pos = static_cast<int>(m - br * p.P);
            const int b = static_cast<int>(br / p.G), r = static_cast<int>(br - (long long)b * p.G);
#pragma unroll
            for (int hh = 0; hh < NHMAX; ++hh) {
                if (hh < p.NH)
                    p.heads[(((long long)b * p.NH + hh) * p.G + r) * p.P + pos] = t[hh] + __ldg(p.bh + hh) + __ldg(p.head_add + hh * p.G + r);
            }
        }
    }
};

}  // namespace tvae
