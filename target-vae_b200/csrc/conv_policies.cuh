// Group-convolution GEMM policies (a-2, models.py:202-225 and its weight gradient).
//
// The convolution is an implicit GEMM  X1[(b,pos), (r,o)] = sum_kk im2col[(b,pos), kk] * bank[(r,o), kk]
// with kk = (c*k + v)*k + u and im2col[(b,pos=(i,j)), kk] = y[b, c, i+v-p, j+u-p] (zero outside).
// C = 1 or 3 is far below the 16-byte TMA im2col granule, so generator warps expand the window
// from an image slab held in shared memory straight into the swizzled A tile; the im2col matrix
// never exists in HBM.  Row tiles never straddle images.
//
//   Conv1Fwd   : A = im2col (generated, K-major), B = rotated filter bank (TMA), epilogue
//                +bias[o], LeakyReLU, tf32 round, store X1 as [(b*G + r)*P + pos][O].
//   Conv1Wgrad : dbank[(r,o), kk] = sum_{b,pos} dX1[(b,r,pos), o] * im2col[(b,pos), kk]
//                A = im2col^T (generated, MN-major; kk on the accumulator rows), B = dX1 (TMA,
//                MN-major), split over (b,pos) with fp32 atomics.  Column kk == K of the generated
//                operand is 1 so that dbank[:, K] accumulates the bias gradient.
#pragma once
#include "linear_policies.cuh"

namespace tvae {

struct ConvGeom {
    int B, C, n, k, p, G, O, d, P;   // d = H' = W', P = d*d
    int K;                           // C*k*k
    int kpad;                        // bank row pitch (multiple of 32, > K when the ones column is needed)
};

// Fills `cnt` (multiple of 4) consecutive kk values of one im2col row into 16-byte chunks.
// slab: [C][slab_rows][n] floats holding image rows [row0, row0 + slab_rows).
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
__device__ __forceinline__ float im2col_fetch(const float* slab, const ConvGeom& g, int slab_rows, int row0, int i, int j,
                                              const Im2colCursor& cur, bool valid) {
    const int iy = i + cur.v - g.p - row0;   // slab-relative row
    const int ix = j + cur.u - g.p;
    const bool ok = valid && cur.c < g.C && iy >= 0 && iy < slab_rows && ix >= 0 && ix < g.n;
    return ok ? slab[(cur.c * slab_rows + iy) * g.n + ix] : 0.f;
}
__device__ __forceinline__ void im2col_advance(Im2colCursor& cur, int k) {
    if (++cur.u == k) {
        cur.u = 0;
        if (++cur.v == k) { cur.v = 0; ++cur.c; }
    }
}

// ------------------------------------------------------------------------------------------------
struct Conv1FwdParams {
    CUtensorMap tmB;          // bank [G*O][kpad]
    int num_stages, num_tiles, tiles_n, tiles_per_image, k_chunks;
    ConvGeom g;
    const float* y;           // (B,C,n,n)
    const float* bias;        // (O)
    float* x1;                // [(b*G + r)*P + pos][O]
    int slab_rows_max;
    int act;                  // 1: LeakyReLU + tf32 rounding (encoder path), 0: raw conv + bias (GroupConv.forward)
};

template <int BN>
struct Conv1Fwd : PolicyBase {
    static constexpr const char* kName = "conv1_fwd";
    using Params = Conv1FwdParams;
    static constexpr int kBN = BN;
    static constexpr bool kAGen = true;
    static constexpr int kProdWarps = 8;
    struct GenState {
        int b, row0, rows;   // slab currently resident
        int i, j;            // output cell of this thread's row
        bool valid;
    };
    __device__ static void prefetch_descs(const Params& p) { tma_prefetch_desc(&p.tmB); }
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        const int mt = tile / p.tiles_n, nt = tile - mt * p.tiles_n;
        const int b = mt / p.tiles_per_image, pt = mt - b * p.tiles_per_image;
        ti.m0 = pt * kBM;
        ti.n0 = nt * BN;
        ti.a0 = b;
        ti.kc_begin = 0;
        ti.kc_end = p.k_chunks;
    }
    __device__ static constexpr uint32_t tx_bytes() { return BN * 128; }
    __device__ static void issue_tma(const Params& p, const TileInfo& ti, int kc, uint32_t, uint32_t sb, uint32_t bar) {
        tma_kmajor(sb, &p.tmB, bar, kc, ti.n0);
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.b = -1; s.row0 = 0; s.rows = 0; }
    __device__ static void gen_tile_begin(const Params& p, const TileInfo& ti, GenState& s, uint8_t* extra, int ptid) {
        const ConvGeom& g = p.g;
        const int i0 = ti.m0 / g.d;
        const int last = min(ti.m0 + kBM - 1, g.P - 1);
        const int i1 = last / g.d;
        const int row0 = max(0, i0 - g.p);
        const int row1 = min(g.n, i1 - g.p + g.k);
        const int rows = max(0, row1 - row0);
        if (s.b != ti.a0 || s.row0 != row0 || s.rows != rows) {   // uniform across the generator warps
            float* slab = reinterpret_cast<float*>(extra);
            named_bar_sync(1, kProdWarps * 32);                   // previous tile's gathers are done
            const float* img = p.y + (long long)ti.a0 * g.C * g.n * g.n;
            const int per_c = rows * g.n;
            for (int idx = ptid; idx < g.C * per_c; idx += kProdWarps * 32) {
                const int c = idx / per_c, rem = idx - c * per_c;
                slab[idx] = to_tf32(__ldg(img + (long long)c * g.n * g.n + (long long)row0 * g.n + rem));
            }
            named_bar_sync(1, kProdWarps * 32);
            s.b = ti.a0; s.row0 = row0; s.rows = rows;
        }
        const int pos = ti.m0 + (ptid & (kBM - 1));
        s.valid = pos < g.P;
        s.i = pos / g.d;
        s.j = pos - s.i * g.d;
    }
    __device__ static void gen_chunk(const Params& p, const TileInfo&, GenState& s, int kc, uint8_t* a_stage, uint8_t* extra, int ptid) {
        const ConvGeom& g = p.g;
        const float* slab = reinterpret_cast<const float*>(extra);
        const int row = ptid & (kBM - 1), half = ptid >> 7;
        Im2colCursor cur = im2col_cursor(kc * kBK + half * 16, g.k);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float4 v;
            v.x = im2col_fetch(slab, g, s.rows, s.row0, s.i, s.j, cur, s.valid); im2col_advance(cur, g.k);
            v.y = im2col_fetch(slab, g, s.rows, s.row0, s.i, s.j, cur, s.valid); im2col_advance(cur, g.k);
            v.z = im2col_fetch(slab, g, s.rows, s.row0, s.i, s.j, cur, s.valid); im2col_advance(cur, g.k);
            v.w = im2col_fetch(slab, g, s.rows, s.row0, s.i, s.j, cur, s.valid); im2col_advance(cur, g.k);
            *reinterpret_cast<float4*>(a_stage + sw128_offset(row, half * 4 + ch)) = v;
        }
    }
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState&, uint32_t taddr, int row, uint8_t*) {
        const ConvGeom& g = p.g;
        const int pos = ti.m0 + row;
        const bool ok = pos < g.P;
        const int N = g.G * g.O;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            const int np = ti.n0 + c * 32;
            if (!ok || np >= N) continue;
            const int r = np / g.O, o0 = np - r * g.O;
            float* dst = p.x1 + (((long long)ti.a0 * g.G + r) * g.P + pos) * g.O + o0;
            const float* bs = p.bias + o0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 t;
                t.x = __uint_as_float(rr[j]) + (p.bias ? __ldg(bs + j) : 0.f);
                t.y = __uint_as_float(rr[j + 1]) + (p.bias ? __ldg(bs + j + 1) : 0.f);
                t.z = __uint_as_float(rr[j + 2]) + (p.bias ? __ldg(bs + j + 2) : 0.f);
                t.w = __uint_as_float(rr[j + 3]) + (p.bias ? __ldg(bs + j + 3) : 0.f);
                if (p.act) {
                    t.x = to_tf32(lrelu(t.x)); t.y = to_tf32(lrelu(t.y)); t.z = to_tf32(lrelu(t.z)); t.w = to_tf32(lrelu(t.w));
                }
                *reinterpret_cast<float4*>(dst + j) = t;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
struct Conv1WgradParams {
    CUtensorMap tmQ;          // dX1 [(B*G*P)][O], MN-major boxes {32, 32}
    int num_stages, num_tiles, tiles_m, tiles_n, splits, chunks_total, chunks_per_split, chunks_per_image;
    ConvGeom g;
    const float* y;
    float* dbank;             // [G*O][kpad], zero-filled by the caller
};

template <int BN>
struct Conv1Wgrad : PolicyBase {
    static constexpr const char* kName = "conv1_wgrad";
    using Params = Conv1WgradParams;
    static constexpr int kBN = BN;
    static constexpr bool kAMajorMN = true;
    static constexpr bool kBMajorMN = true;
    static constexpr bool kAGen = true;
    static constexpr int kProdWarps = 8;
    struct GenState { int b; };
    __device__ static void prefetch_descs(const Params& p) { tma_prefetch_desc(&p.tmQ); }
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        const int per_split = p.tiles_m * p.tiles_n;
        const int sp = tile / per_split;
        const int rem = tile - sp * per_split;
        const int mt = rem / p.tiles_n, nt = rem - mt * p.tiles_n;
        ti.m0 = mt * kBM;   // kk
        ti.n0 = nt * BN;    // (r,o)
        ti.kc_begin = sp * p.chunks_per_split;
        ti.kc_end = min(ti.kc_begin + p.chunks_per_split, p.chunks_total);
    }
    __device__ static constexpr uint32_t tx_bytes() { return BN * 128; }
    __device__ static void issue_tma(const Params& p, const TileInfo& ti, int kc, uint32_t, uint32_t sb, uint32_t bar) {
        const ConvGeom& g = p.g;
        const int b = kc / p.chunks_per_image, pc = kc - b * p.chunks_per_image;
        for (int cb = 0; cb < BN / 32; ++cb) {
            const int np = ti.n0 + cb * 32;
            const int r = np / g.O, o0 = np - r * g.O;
            // rows past the end of this (b,r) segment are multiplied by generated zeros; rows past the
            // end of the tensor (and r >= G) are zero-filled by TMA.
            const int row = (r < g.G) ? ((b * g.G + r) * g.P + pc * kBK) : 0x3fffffff;
            tma_load_2d(sb + cb * (kBK * 128), &p.tmQ, bar, o0, row);
        }
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.b = -1; }
    __device__ static void gen_chunk(const Params& p, const TileInfo& ti, GenState& s, int kc, uint8_t* a_stage, uint8_t* extra, int ptid) {
        const ConvGeom& g = p.g;
        float* slab = reinterpret_cast<float*>(extra);   // whole image [C][n][n]
        const int b = kc / p.chunks_per_image, pc = kc - b * p.chunks_per_image;
        if (s.b != b) {
            named_bar_sync(1, kProdWarps * 32);
            const float* img = p.y + (long long)b * g.C * g.n * g.n;
            for (int idx = ptid; idx < g.C * g.n * g.n; idx += kProdWarps * 32) slab[idx] = to_tf32(__ldg(img + idx));
            named_bar_sync(1, kProdWarps * 32);
            s.b = b;
        }
        const int rrow = ptid & 31, cb = (ptid >> 5) & 3, half = ptid >> 7;
        const int pos = pc * kBK + rrow;
        const bool valid = pos < g.P;
        const int i = pos / g.d, j = pos - i * g.d;
        int kk = ti.m0 + cb * 32 + half * 16;
        Im2colCursor cur = im2col_cursor(kk, g.k);
        uint8_t* blk = a_stage + cb * (kBK * 128);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float e[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                e[q] = im2col_fetch(slab, g, g.n, 0, i, j, cur, valid);
                if (kk == g.K && valid) e[q] = 1.f;   // ones column -> bias gradient
                im2col_advance(cur, g.k);
                ++kk;
            }
            *reinterpret_cast<float4*>(blk + sw128b32_offset(rrow, half * 4 + ch)) = make_float4(e[0], e[1], e[2], e[3]);
        }
    }
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState&, uint32_t taddr, int row, uint8_t*) {
        const ConvGeom& g = p.g;
        const int kk = ti.m0 + row;
        const int N = g.G * g.O;
        const bool empty = ti.kc_begin >= ti.kc_end;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            if (kk >= g.kpad || empty) continue;
            const int np0 = ti.n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int np = np0 + j;
                if (np < N) atomicAdd(p.dbank + (long long)np * g.kpad + kk, __uint_as_float(rr[j]));
            }
        }
    }
};

}  // namespace tvae

namespace tvae {

// ------------------------------------------------------------------------------------------------
// Conv2Heads: h = LeakyReLU(x1 W2^T + b2) (Conv3d 1x1x1, models.py:347,356) with the attention / theta / z
// heads (models.py:358,390,392) and the "+p_r", "+offset_r" adds (models.py:382,394-399) fused in the epilogue.
// O <= BN so one accumulator tile holds a full channel vector per row.
constexpr int kMaxNH = 36;
struct Conv2HeadsParams {
    CUtensorMap tmA, tmB;     // x1 [R][O]; W2 [O][O]
    int num_stages, num_tiles, k_chunks;
    long long R;              // B*G*P rows
    int O, NH, G, P;
    const float* b2;          // (O)
    const float* wh;          // [NH][O]
    const float* bh;          // [NH]
    const float* head_add;    // [NH][G]
    float* h;                 // [R][O]
    float* heads;             // (B,NH,G,P)
};

template <int BN, int NHMAX>
struct Conv2Heads : PolicyBase {
    static constexpr const char* kName = "conv2_heads";
    using Params = Conv2HeadsParams;
    static constexpr int kBN = BN;
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
        tma_kmajor(sa, &p.tmA, bar, kc, ti.m0);
        tma_kmajor(sb, &p.tmB, bar, kc, 0);
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
            float* dst = p.h + m * p.O + o0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
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
