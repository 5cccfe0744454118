// CTA-pair group-convolution policies for tc_gemm2_kernel (a-2, models.py:202-225, and its weight gradient).
//
// Both GEMMs consume the im2col matrix  im2col[(b,pos=(i,j)), kk=(c,v,u)] = ypad[b, c, i+v, j+u]  (ypad = image
// zero-padded by p on every side), which is never materialised: generator warps copy it out of a zero-padded
// image slab held in shared memory.  Padding lives in the slab, so the inner loop is 4 x LDS.32 at
// lane-consecutive addresses + one swizzled STS.128 per 16-byte granule, with the (c,v,u) -> slab offset of
// every granule read from a small table (no per-element index arithmetic, no bounds checks).
//
//   Conv1Fwd2   : X1[(b,pos), (r,o)] = lrelu(im2col . bank^T + bias)       A = im2col tile (K-major), B = bank (TMA)
//   Conv1Wgrad2 : dbank[(r,o), kk]  += sum_(b,pos) dX1[(b,r,pos), o] im2col[(b,pos), kk]
//                 accumulator rows = kk (A = im2col^T generated MN-major), columns = (r,o) (B = dX1, TMA, MN-major),
//                 reduction over (b,pos) split across CTA pairs, fp32 atomics into dbank.
#pragma once
#include "conv_policies.cuh"
#include "tc_gemm2.cuh"

namespace tvae {

struct SlabGeom {
    int Wp;        // padded image width  n + 2p
    int pitch;     // slab row pitch in floats, = d (mod 32) so that lane-consecutive cells never bank-conflict
    int rows_max;  // slab rows allocated per channel
};

// Fill slab[c][ip - r_lo][x] (ip, x in padded coordinates) for rows [r_lo, r_lo + rows) of channels [c_lo, c_lo + nc).
__device__ __forceinline__ void fill_slab(float* slab, const SlabGeom& sg, const ConvGeom& g, const float* img, int c_lo, int nc,
                                          int r_lo, int rows, int tid, int nthreads) {
    const int per_c = rows * sg.pitch;
    for (int idx = tid; idx < nc * per_c; idx += nthreads) {
        const int c = idx / per_c, rem = idx - c * per_c;
        const int rr = rem / sg.pitch, x = rem - rr * sg.pitch;
        const int iy = r_lo + rr - g.p, ix = x - g.p;
        float v = 0.f;
        if (iy >= 0 && iy < g.n && ix >= 0 && ix < g.n) v = to_tf32(__ldg(img + ((long long)(c_lo + c) * g.n + iy) * g.n + ix));
        slab[c * sg.rows_max * sg.pitch + rr * sg.pitch + x] = v;
    }
}

// ------------------------------------------------------------------------------------------------
struct Conv1Fwd2Params {
    CUtensorMap tmB;          // bank [G*O][kpad], boxes {32 k, 128 rows}
    int num_stages, num_tiles, n_passes, tiles_per_image, m_tiles, k_chunks;
    ConvGeom g;
    SlabGeom sg;
    const float* y;           // (B,C,n,n)
    const float* bias;        // (O) or null
    float* x1;                // [(b*G + r)*P + pos][O]
    int act;
    int gran;                 // 1: offset table per 4-float granule (k % 4 == 0), 0: per element
    int tab_entries;
    int skip;                 // 1: skip K chunks that only meet zero padding (needs k*k % 32 == 0)
    int chunks_per_channel;   // k*k / 32 when skip, else k_chunks
};

struct Conv1Fwd2 : PolicyBase {
    static constexpr const char* kName = "conv1_fwd";
    using Params = Conv1Fwd2Params;
    struct GenState {
        int b, r_lo, rows;   // slab currently resident
        int base;            // slab offset of this thread's output cell
    };
    __device__ static void prefetch_descs(const Params& p) { tma_prefetch_desc(&p.tmB); }
    // extra smem: [tab_entries] int offsets, then the slab
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        int* tab = reinterpret_cast<int*>(extra);
        const ConvGeom& g = p.g;
        const int step = p.gran ? 4 : 1;
        for (int e = tid; e < p.tab_entries; e += nthreads) {
            const int kk = e * step;
            int off = -1;
            if (kk < g.K) {
                const Im2colCursor cur = im2col_cursor(kk, g.k);
                off = (cur.c * p.sg.rows_max + cur.v) * p.sg.pitch + cur.u;
            }
            tab[e] = off;
        }
    }
    // Zero-padding skip: filter row v only meets image rows for output rows i with 0 <= i + v - p < n.  For the
    // output rows of BOTH CTAs' tiles the live v range is [v_lo, v_hi); K chunks outside it multiply zeros and
    // are never generated, loaded or issued (cfg2/cfg3: ~25 % of the dense count).  Needs chunk-aligned channels.
    __device__ static void tile_info(const Params& p, int tile, uint32_t rank, PairTile& ti) {
        const ConvGeom& g = p.g;
        const int mp = tile / p.n_passes, np = tile - mp * p.n_passes;
        const int N = g.G * g.O;
        ti.n0 = np * (kAcc * kAccN);
        ti.n_acc = (N - ti.n0 > kAccN) ? 2 : 1;
        const int mt = 2 * mp + static_cast<int>(rank);
        ti.m_tile = mt < p.m_tiles ? mt : -1;
        ti.a0 = mt / p.tiles_per_image;                               // image
        ti.a1 = (mt - ti.a0 * p.tiles_per_image) * kBM;               // first position
        if (p.skip) {
            int v_lo = g.k, v_hi = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int t = 2 * mp + h;
                if (t < p.m_tiles) {
                    const int pos0 = (t % p.tiles_per_image) * kBM;
                    const int i_first = pos0 / g.d, i_last = min(pos0 + kBM - 1, g.P - 1) / g.d;
                    v_lo = min(v_lo, max(0, g.p - i_last));
                    v_hi = max(v_hi, min(g.k, g.p - i_first + g.n));
                }
            }
            const int lo = (v_lo * g.k) / kBK, hi = (v_hi * g.k + kBK - 1) / kBK;   // chunks within one channel
            ti.a2 = lo;
            ti.a3 = hi > lo ? hi - lo : 0;
            ti.kc_begin = 0;
            ti.kc_end = g.C * ti.a3;
        } else {
            ti.a2 = 0;
            ti.a3 = p.k_chunks;
            ti.kc_begin = 0;
            ti.kc_end = p.k_chunks;
        }
    }
    __device__ static int chunk(const Params& p, const PairTile& ti, int q) {
        const int c = q / ti.a3;
        return c * p.chunks_per_channel + ti.a2 + (q - c * ti.a3);
    }
    __device__ static void issue_tma(const Params& p, const PairTile&, int kc, int n_row0, uint32_t sb, uint32_t bar) {
        tma_load_2d_pair(sb, &p.tmB, bar, kc * kBK, n_row0);
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.b = -1; s.r_lo = 0; s.rows = 0; s.base = 0; }
    __device__ static void gen_tile_begin(const Params& p, const PairTile& ti, GenState& s, uint8_t* extra, int ptid) {
        if (ti.m_tile < 0) return;
        const ConvGeom& g = p.g;
        float* slab = reinterpret_cast<float*>(extra + p.tab_entries * 4);
        const int i0 = ti.a1 / g.d;
        const int last = min(ti.a1 + kBM - 1, g.P - 1);
        const int i1 = last / g.d;
        const int r_lo = i0, rows = i1 - i0 + g.k;                   // padded rows [i0, i1 + k)
        if (s.b != ti.a0 || s.r_lo != r_lo || s.rows != rows) {       // uniform across the generator warps
            named_bar_sync(1, kGenWarps * 32);                        // previous tile's gathers are done
            fill_slab(slab, p.sg, g, p.y + (long long)ti.a0 * g.C * g.n * g.n, 0, g.C, r_lo, rows, ptid, kGenWarps * 32);
            named_bar_sync(1, kGenWarps * 32);
            s.b = ti.a0; s.r_lo = r_lo; s.rows = rows;
        }
        const int pos = min(ti.a1 + (ptid & (kBM - 1)), g.P - 1);     // rows past the image end are discarded by the epilogue
        const int i = pos / g.d, j = pos - i * g.d;
        s.base = (i - r_lo) * p.sg.pitch + j;
    }
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, int kc, uint8_t* a_stage, uint8_t* extra, int ptid) {
        const int* tab = reinterpret_cast<const int*>(extra);
        const float* slab = reinterpret_cast<const float*>(extra + p.tab_entries * 4) + s.base;
        const int row = ptid & (kBM - 1), half = ptid >> 7;
        const bool live = ti.m_tile >= 0;
        if (p.gran) {
            const int4 offs = *reinterpret_cast<const int4*>(tab + (kc * kBK + half * 16) / 4);
            const int o[4] = {offs.x, offs.y, offs.z, offs.w};
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && o[ch] >= 0) {
                    const float* q = slab + o[ch];
                    v = make_float4(q[0], q[1], q[2], q[3]);
                }
                *reinterpret_cast<float4*>(a_stage + sw128_offset(row, half * 4 + ch)) = v;
            }
        } else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const int4 o = *reinterpret_cast<const int4*>(tab + kc * kBK + half * 16 + ch * 4);
                float4 v;
                v.x = (live && o.x >= 0) ? slab[o.x] : 0.f;
                v.y = (live && o.y >= 0) ? slab[o.y] : 0.f;
                v.z = (live && o.z >= 0) ? slab[o.z] : 0.f;
                v.w = (live && o.w >= 0) ? slab[o.w] : 0.f;
                *reinterpret_cast<float4*>(a_stage + sw128_offset(row, half * 4 + ch)) = v;
            }
        }
    }
    __device__ static void epilogue(const Params& p, const PairTile& ti, int n0, uint32_t taddr, int row, bool has_work, uint8_t*) {
        const ConvGeom& g = p.g;
        const int pos = ti.a1 + row;
        const bool ok = has_work && ti.m_tile >= 0 && pos < g.P;
        const int N = g.G * g.O;
#pragma unroll 1
        for (int c = 0; c < kAccN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            const int np = n0 + c * 32;
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
struct Conv1Wgrad2Params {
    CUtensorMap tmQ;          // dX1 [(B*G*P)][O], MN-major boxes {32 o, 32 rows}
    int num_stages, num_tiles, m_pairs, n_passes, m_tiles, splits, chunks_total, chunks_per_split, chunks_per_image;
    ConvGeom g;
    SlabGeom sg;
    const float* y;
    float* dbank;             // [G*O][kpad], zero-filled by the caller
    int ones_col;             // 1: accumulator row kk == K is fed with ones (conv1 bias gradient)
    int skip;                 // 1: skip position chunks that only meet zero padding
};

struct Conv1Wgrad2 : PolicyBase {
    static constexpr const char* kName = "conv1_wgrad";
    using Params = Conv1Wgrad2Params;
    static constexpr bool kAMajorMN = true;
    static constexpr bool kBMajorMN = true;
    struct GenState {
        int b, m_tile;       // image / kk-tile whose slab + table are resident
        int c_lo, r_lo;
    };
    __device__ static void prefetch_descs(const Params& p) { tma_prefetch_desc(&p.tmQ); }
    // Zero-padding skip: accumulator rows kk of the pair cover filter rows [va, vb]; only output rows i with
    // 0 <= i + v - p < n for some such v contribute, i.e. a contiguous range of position chunks per image.
    // The reduction runs over the compact index q = b * cnt + (pc - lo) and is split evenly in q.
    __device__ static void tile_info(const Params& p, int tile, uint32_t rank, PairTile& ti) {
        const ConvGeom& g = p.g;
        const int per_split = p.m_pairs * p.n_passes;
        const int sp = tile / per_split;
        const int rem = tile - sp * per_split;
        const int mp = rem / p.n_passes, np = rem - mp * p.n_passes;
        const int N = g.G * g.O;
        ti.n0 = np * (kAcc * kAccN);
        ti.n_acc = (N - ti.n0 > kAccN) ? 2 : 1;
        const int mt = 2 * mp + static_cast<int>(rank);
        ti.m_tile = mt < p.m_tiles ? mt : -1;
        ti.a0 = mt * kBM;                                             // first kk of this CTA's accumulator rows
        int lo = 0, cnt = p.chunks_per_image;
        const int kk0 = 2 * mp * kBM, kk1 = kk0 + 2 * kBM - 1;        // kk range of the pair
        if (p.skip && !(p.ones_col && kk1 >= g.K)) {
            const Im2colCursor c0 = im2col_cursor(kk0, g.k), c1 = im2col_cursor(min(kk1, g.K - 1), g.k);
            if (c0.c == c1.c) {
                const int i_lo = max(0, g.p - c1.v), i_hi = min(g.d - 1, g.p - c0.v + g.n - 1);
                if (i_hi >= i_lo) {
                    lo = (i_lo * g.d) / kBK;
                    cnt = ((i_hi + 1) * g.d + kBK - 1) / kBK - lo;
                } else {
                    cnt = 0;
                }
            }
        }
        ti.a1 = lo;
        ti.a2 = cnt;
        const int total = g.B * cnt;
        const int cps = (total + p.splits - 1) / p.splits;
        ti.kc_begin = min(sp * cps, total);
        ti.kc_end = min(ti.kc_begin + cps, total);
    }
    __device__ static int chunk(const Params& p, const PairTile& ti, int q) {
        const int b = q / ti.a2;
        return b * p.chunks_per_image + ti.a1 + (q - b * ti.a2);
    }
    // B-half of one accumulator: 128 (r,o) columns starting at n_col0, reduction rows = 32 positions of chunk kc
    __device__ static void issue_tma(const Params& p, const PairTile&, int kc, int n_col0, uint32_t sb, uint32_t bar) {
        const ConvGeom& g = p.g;
        const int b = kc / p.chunks_per_image, pc = kc - b * p.chunks_per_image;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const int np = n_col0 + cb * 32;
            const int r = np / g.O, o0 = np - r * g.O;
            // rows past the end of this (b,r) segment are multiplied by generated zeros; rows past the end of
            // the tensor (and r >= G) are zero-filled by TMA.
            const int row = (r < g.G) ? ((b * g.G + r) * g.P + pc * kBK) : 0x3fffffff;
            tma_load_2d_pair(sb + cb * (kBK * 128), &p.tmQ, bar, o0, row);
        }
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.b = -1; s.m_tile = -2; s.c_lo = 0; s.r_lo = 0; }
    // extra smem: [128] int offsets of this CTA's kk rows, then the slab
    __device__ static void gen_tile_begin(const Params& p, const PairTile& ti, GenState& s, uint8_t* extra, int ptid) {
        if (ti.m_tile < 0 || s.m_tile == ti.m_tile) return;
        const ConvGeom& g = p.g;
        int* tab = reinterpret_cast<int*>(extra);
        named_bar_sync(1, kGenWarps * 32);
        // channel / row window touched by kk in [a0, a0 + 128)
        const int kk_hi = min(ti.a0 + kBM, g.K) - 1;
        const Im2colCursor lo = im2col_cursor(min(ti.a0, g.K - 1), g.k), hi = im2col_cursor(kk_hi, g.k);
        const int v_lo = (lo.c == hi.c) ? lo.v : 0;
        s.c_lo = lo.c;
        s.r_lo = v_lo;
        if (ptid < kBM) {
            const int kk = ti.a0 + ptid;
            int off = -1;                       // zero row
            if (kk < g.K) {
                const Im2colCursor cur = im2col_cursor(kk, g.k);
                off = ((cur.c - lo.c) * p.sg.rows_max + (cur.v - v_lo)) * p.sg.pitch + cur.u;
            } else if (kk == g.K && p.ones_col) {
                off = -2;                       // ones row
            }
            tab[ptid] = off;
        }
        s.m_tile = ti.m_tile;
        s.b = -1;                               // slab must be refilled for the new window
        named_bar_sync(1, kGenWarps * 32);
    }
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, int kc, uint8_t* a_stage, uint8_t* extra, int ptid) {
        const ConvGeom& g = p.g;
        const int* tab = reinterpret_cast<const int*>(extra);
        float* slab = reinterpret_cast<float*>(extra + kBM * 4);
        // lane -> (position row, chunk parity): every aligned group of 8 lanes covers 4 rows x 2 adjacent 16-byte
        // chunks = 8 distinct slots of the 32-byte-atom swizzle (conflict-free STS.128), while the LDS stay within
        // 20 consecutive words.  warp -> (32-wide kk block, half of the 32 position rows).
        const int lane = ptid & 31, w = ptid >> 5;
        const int cb = w & 3;
        const int rrow = (lane & 3) | (((lane >> 3) & 3) << 2) | ((w >> 2) << 4);
        const int par = (lane >> 2) & 1;
        uint8_t* blk = a_stage + cb * (kBK * 128);
        if (ti.m_tile < 0) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
                *reinterpret_cast<float4*>(blk + sw128b32_offset(rrow, 2 * ch + par)) = make_float4(0.f, 0.f, 0.f, 0.f);
            return;
        }
        const int b = kc / p.chunks_per_image, pc = kc - b * p.chunks_per_image;
        if (s.b != b) {
            named_bar_sync(1, kGenWarps * 32);
            const int kk_hi = min(ti.a0 + kBM, g.K) - 1;
            const Im2colCursor lo = im2col_cursor(min(ti.a0, g.K - 1), g.k), hi = im2col_cursor(kk_hi, g.k);
            const int nc = hi.c - lo.c + 1;
            const int rows = (nc == 1 ? hi.v - lo.v : g.k - 1) + g.d;      // padded rows [r_lo, r_lo + rows)
            fill_slab(slab, p.sg, g, p.y + (long long)b * g.C * g.n * g.n, lo.c, nc, s.r_lo, rows, ptid, kGenWarps * 32);
            named_bar_sync(1, kGenWarps * 32);
            s.b = b;
        }
        const int pos = pc * kBK + rrow;
        const bool valid = pos < g.P;
        const int i = pos / g.d, j = pos - i * g.d;
        const float* src = slab + i * p.sg.pitch + j;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const int chunk = 2 * ch + par;
            const int4 o = *reinterpret_cast<const int4*>(tab + cb * 32 + chunk * 4);
            float4 v;
            v.x = !valid ? 0.f : (o.x >= 0 ? src[o.x] : (o.x == -2 ? 1.f : 0.f));
            v.y = !valid ? 0.f : (o.y >= 0 ? src[o.y] : (o.y == -2 ? 1.f : 0.f));
            v.z = !valid ? 0.f : (o.z >= 0 ? src[o.z] : (o.z == -2 ? 1.f : 0.f));
            v.w = !valid ? 0.f : (o.w >= 0 ? src[o.w] : (o.w == -2 ? 1.f : 0.f));
            *reinterpret_cast<float4*>(blk + sw128b32_offset(rrow, chunk)) = v;
        }
    }
    __device__ static void epilogue(const Params& p, const PairTile& ti, int n0, uint32_t taddr, int row, bool has_work, uint8_t*) {
        const ConvGeom& g = p.g;
        const int kk = ti.a0 + row;
        const int N = g.G * g.O;
        const bool ok = has_work && ti.m_tile >= 0 && kk < g.kpad;
#pragma unroll 1
        for (int c = 0; c < kAccN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            if (!ok) continue;
            const int np0 = n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int np = np0 + j;
                if (np < N) atomicAdd(p.dbank + (long long)np * g.kpad + kk, __uint_as_float(rr[j]));
            }
        }
    }
};

}  // namespace tvae
