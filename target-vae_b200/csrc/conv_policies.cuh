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
    int act;                  // kActTanh or LeakyReLU
};

template <int BN, int NHMAX, bool TANH = false>
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
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) + s_b2[o0 + j];
            act_vec<TANH>(v);
            __half* dst = p.h + m * p.O + o0;
#pragma unroll
            for (int j = 0; p.h != nullptr && j < 32; j += 8) {      // h == NULL: inference only, the hidden map is not kept
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

// Conv2HeadsTC: the same stage with the heads ALSO on the tensor core (O == 128, NH <= 32).
//
// The epilogue above spends 128 x NH FMAs per row on the heads (NH = 19 for z = 8) and writes h with per-row stores; it
// runs at 35 % (cfg2) .. 15 % (cfg5) of the HBM rate this streaming stage is bound by.  Here an epilogue thread only does
// bias + LeakyReLU + fp16 packing of its row into 128 B-swizzled staging tiles in shared memory.  The tile h_hi = fp16(h) is
//   (1) handed to the TMA store unit (coalesced, asynchronous write of h), and
//   (2) the A operand of a second, tiny tensor-core product  heads[128 x NH] = h_tile[128 x 128] . Wh^T[NH x 128],
//       issued by one thread of the epilogue group into TMEM columns the accumulator stages leave free and read back
//       with tcgen05.ld.
// Precision of (2): the attention logits feed a softmax over G*H'*W' cells and the parameter gradients are kink-limited
// (tests/helpers.py), so the heads keep the fp32-level accuracy of the CUDA-core version through a two-term split of both
// operands: h = h_hi + h_lo, Wh = W_hi + W_lo (fp16 each, 22 significand bits together) and
//       heads = h_hi.W_hi + h_hi.W_lo + h_lo.W_hi       (fp32 accumulation; the lo.lo term is below fp32 rounding)
// as  D[:, 0:2N) = h_hi . [W_hi; W_lo]^T  and  D[:, 0:N) += h_lo . W_hi^T, summed by the reader.
// W2 (the B operand of the main product) is identical for every tile and stays resident in shared memory
// (kBResidentChunks); the accumulator stage is released as soon as it has been read (kEpiSelfRelease).
struct Conv2HeadsTCParams {
    CUtensorMap tmA, tmB;     // x1 fp16 [R][O]; W2 fp16 [O][O]
    CUtensorMap tmH;          // h fp16 [R][O] store view, boxes {64, 128 rows}
    int num_stages, num_tiles, k_chunks;
    long long R;
    int O, NH, NHpad, G, P;
    int store_h;              // 0: tmH is not set up and h is not written (inference only)
    int act;                  // kActTanh or LeakyReLU
    const float* b2;
    const float* wh;          // [NH][O] fp32 (split into fp16 hi / lo operand tiles in setup)
    const float* bh;
    const float* head_add;
    float* heads;
};

template <bool TANH>
struct Conv2HeadsTCT : PolicyBase {
    static constexpr const char* kName = "conv2_heads";
    using Params = Conv2HeadsTCParams;
    static constexpr int kBN = 128;
    static constexpr bool kF16 = true;
    static constexpr int kEpiGroups = 2;
    static constexpr int kMaxAccStages = 2;          // TMEM columns [256, 384): 64 heads columns per epilogue group
    static constexpr bool kEpiSelfRelease = true;
    static constexpr int kBResidentChunks = 2;       // W2 = 2 K chunks of [128][64] fp16
    // extra smem (1024-byte aligned): Wh operand tiles 2 chunks x [64 rows: hi 0..NHpad-1, lo NHpad..2 NHpad-1][128 B],
    // staging per group: h_hi 2 x 16 KB, h_lo 2 x 16 KB; then b2 and the two heads barriers
    static constexpr int kWhChunkBytes = 64 * 128;
    static constexpr int kWhBytes = 2 * kWhChunkBytes;
    static constexpr int kStageOff = kWhBytes;
    static constexpr int kBlockBytes = kBM * 128;
    static constexpr int kGroupStageBytes = 4 * kBlockBytes;
    static constexpr int kB2Off = kStageOff + 2 * kGroupStageBytes;
    static constexpr int kBarOff = kB2Off + 128 * 4;
    static constexpr int kExtraBytes = kBarOff + 16;
    struct EpiState {
        uint32_t tempty, tmem_free;   // set by the kernel (kEpiSelfRelease)
        uint32_t hphase;
        int grp, tiles;
    };
    __device__ static void prefetch_descs(const Params& p) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        if (p.store_h) tma_prefetch_desc(&p.tmH);
    }
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        // Wh as K-major fp16 operand tiles: chunk kc = o / 64, rows [0, NHpad) = hi part, [NHpad, 2 NHpad) = lo part
        for (int i = tid; i < 2 * 32 * 64; i += nthreads) {
            const int kc = i / (32 * 64), rem = i - kc * (32 * 64);
            const int hh = rem / 64, e = rem - hh * 64;
            if (hh >= p.NHpad) continue;
            const float w = (hh < p.NH) ? __ldg(p.wh + hh * p.O + kc * 64 + e) : 0.f;
            const __half hi = __float2half_rn(w);
            const __half lo = __float2half_rn(w - __half2float(hi));
            uint8_t* tile = extra + kc * kWhChunkBytes;
            *reinterpret_cast<__half*>(tile + sw128_offset(hh, e >> 3) + (e & 7) * 2) = hi;
            *reinterpret_cast<__half*>(tile + sw128_offset(p.NHpad + hh, e >> 3) + (e & 7) * 2) = lo;
        }
        float* s_b2 = reinterpret_cast<float*>(extra + kB2Off);
        for (int i = tid; i < 128; i += nthreads) s_b2[i] = __ldg(p.b2 + i);
        if (tid == 0) {
            mbar_init(smem_u32(extra + kBarOff), 1);
            mbar_init(smem_u32(extra + kBarOff + 8), 1);
            fence_barrier_init();
        }
        fence_proxy_async_smem();      // the Wh tiles are read by the tensor core (async proxy)
    }
    __device__ static void epi_init(const Params&, EpiState& st, uint8_t*, int slot) {
        st.grp = slot >> 7;
        st.tiles = 0;
        st.hphase = 0;
    }
    __device__ static void epi_finish(const Params&, EpiState&, uint8_t*, int slot) {
        if ((slot & 127) == 0) tma_store_wait<0>();
    }
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        ti.m0 = tile * kBM;
        ti.n0 = 0;
        ti.kc_begin = 0;
        ti.kc_end = p.k_chunks;
    }
    __device__ static void issue_tma_a(const Params& p, const TileInfo& ti, int kc, uint32_t sa, uint32_t bar) {
        tma_kmajor_h(sa, &p.tmA, bar, kc, ti.m0);
    }
    __device__ static void issue_tma_b(const Params& p, const TileInfo&, int kc, uint32_t sb, uint32_t bar) {
        tma_kmajor_h(sb, &p.tmB, bar, kc, 0);
    }
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState& st, uint32_t taddr, int row, uint8_t* extra) {
        const float* s_b2 = reinterpret_cast<const float*>(extra + kB2Off);
        uint8_t* buf = extra + kStageOff + st.grp * kGroupStageBytes;      // h_hi blocks 0,1 then h_lo blocks 0,1
        const uint32_t hbar = smem_u32(extra + kBarOff + 8 * st.grp);
        const int bar_id = 2 + st.grp;
        // the previous tile's TMA store has read the staging tiles (its heads product completed before we got here)
        if (st.tiles > 0) {
            if (row == 0) tma_store_wait_read<0>();
            named_bar_sync(bar_id, kEpiWarps * 32);
        }
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
            uint32_t rr[2][32];
            tmem_ld_32x32(taddr + blk * 64, rr[0]);
            tmem_ld_32x32(taddr + blk * 64 + 32, rr[1]);
            tmem_ld_wait();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float v[8];
#pragma unroll
                    for (int q = 0; q < 8; q += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(s_b2 + blk * 64 + hf * 32 + j + q);
                        v[q] = __uint_as_float(rr[hf][j + q]) + bb.x;
                        v[q + 1] = __uint_as_float(rr[hf][j + q + 1]) + bb.y;
                        v[q + 2] = __uint_as_float(rr[hf][j + q + 2]) + bb.z;
                        v[q + 3] = __uint_as_float(rr[hf][j + q + 3]) + bb.w;
                    }
                    act_vec<TANH>(v);
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const __half2 hv = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
                        const float2 back = __half22float2(hv);
                        const __half2 lv = __floats2half2_rn(v[2 * q] - back.x, v[2 * q + 1] - back.y);
                        hi[q] = *reinterpret_cast<const uint32_t*>(&hv);
                        lo[q] = *reinterpret_cast<const uint32_t*>(&lv);
                    }
                    const uint32_t off = blk * kBlockBytes + sw128_offset(row, hf * 4 + (j >> 3));
                    *reinterpret_cast<uint4*>(buf + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(buf + 2 * kBlockBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
        // accumulator stage fully read: hand it back to the MMA warp now
        tc_fence_before();
        mbar_arrive(st.tempty);
        // staging tiles -> async proxy; one thread stores h and issues the heads product
        fence_proxy_async_smem();
        named_bar_sync(bar_id, kEpiWarps * 32);
        const uint32_t d_tmem = st.tmem_free + st.grp * 64;                  // this group's heads accumulator (lane 0)
        if (row == 0) {
            if (p.store_h) {      // 0: inference only (get_latent) - h feeds the heads product from shared memory and is not kept
                tma_store_2d(&p.tmH, smem_u32(buf), 0, ti.m0);
                tma_store_2d(&p.tmH, smem_u32(buf + kBlockBytes), 64, ti.m0);
                tma_store_commit();
            }
            tc_fence_after();
            const uint32_t idesc2 = make_idesc_f16(kBM, 2 * p.NHpad, false, false, 0, 0);
            const uint32_t idesc1 = make_idesc_f16(kBM, p.NHpad, false, false, 0, 0);
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
                for (int ks = 0; ks < kKSteps; ++ks) {      // D[:, 0:2N) (+)= h_hi . [W_hi; W_lo]^T
                    const uint64_t adesc = make_smem_desc(smem_u32(buf + kc * kBlockBytes) + ks * 32, 16, 1024, kLayoutSw128);
                    const uint64_t bdesc = make_smem_desc(smem_u32(extra + kc * kWhChunkBytes) + ks * 32, 16, 1024, kLayoutSw128);
                    umma_f16(d_tmem, adesc, bdesc, idesc2, (kc | ks) ? 1u : 0u);
                }
            }
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
                for (int ks = 0; ks < kKSteps; ++ks) {      // D[:, 0:N) += h_lo . W_hi^T
                    const uint64_t adesc = make_smem_desc(smem_u32(buf + (2 + kc) * kBlockBytes) + ks * 32, 16, 1024, kLayoutSw128);
                    const uint64_t bdesc = make_smem_desc(smem_u32(extra + kc * kWhChunkBytes) + ks * 32, 16, 1024, kLayoutSw128);
                    umma_f16(d_tmem, adesc, bdesc, idesc1, 1u);
                }
            }
            umma_commit(hbar);
        }
        mbar_wait(hbar, st.hphase);
        st.hphase ^= 1;
        tc_fence_after();
        const uint32_t haddr = d_tmem + (static_cast<uint32_t>((row >> 5) * 32) << 16);   // this warp's lane quarter
        const long long m = (long long)ti.m0 + row;
        const bool ok = m < p.R;
        const long long br = m / p.P;            // b*G + r
        const int pos = static_cast<int>(m - br * p.P);
        const int b = static_cast<int>(br / p.G), r = static_cast<int>(br - (long long)b * p.G);
        float* dst = p.heads + ((long long)b * p.NH * p.G + r) * p.P + pos;
        const long long hstride = (long long)p.G * p.P;
#pragma unroll 1
        for (int c = 0; c * 16 < p.NHpad; ++c) {
            uint32_t t0[16], t1[16];
            tmem_ld_32x16(haddr + c * 16, t0);                 // h_hi.W_hi + h_lo.W_hi
            tmem_ld_32x16(haddr + p.NHpad + c * 16, t1);       // h_hi.W_lo
            tmem_ld_wait();
            if (ok) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int hh = c * 16 + j;
                    if (hh < p.NH)
                        dst[hh * hstride] = (__uint_as_float(t0[j]) + __uint_as_float(t1[j])) + __ldg(p.bh + hh) + __ldg(p.head_add + hh * p.G + r);
                }
            }
        }
        tc_fence_before();       // orders these TMEM reads before the next tile's heads MMA (issued after a named barrier)
        ++st.tiles;
    }
};
using Conv2HeadsTC = Conv2HeadsTCT<false>;

}  // namespace tvae
