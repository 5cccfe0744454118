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
// One warp per slab row, lanes stride the row: no per-element index arithmetic, all loads of a row independent.
__device__ __forceinline__ void fill_slab(float* slab, const SlabGeom& sg, const ConvGeom& g, const float* img, int c_lo, int nc,
                                          int r_lo, int rows, int tid, int nthreads) {
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
    for (int rw = warp; rw < nc * rows; rw += nwarps) {
        const int c = rw / rows, rr = rw - c * rows;
        const int iy = r_lo + rr - g.p;
        const bool row_ok = iy >= 0 && iy < g.n;
        const float* src = img + ((long long)(c_lo + c) * g.n + (row_ok ? iy : 0)) * g.n - g.p;
        float* dst = slab + (c * sg.rows_max + rr) * sg.pitch;
        for (int x = lane; x < sg.pitch; x += 32) {
            const int ix = x - g.p;
            float v = 0.f;
            if (row_ok && ix >= 0 && ix < g.n) v = to_tf32(__ldg(src + x));
            dst[x] = v;
        }
    }
}

// 3-D TMA load for the pair kernel (coordinates innermost first)
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ------------------------------------------------------------------------------------------------
struct Conv1Fwd2Params {
    CUtensorMap tmB;          // bank [G*O][kpad], boxes {32 k, 128 rows}
    int num_stages, num_tiles, n_passes, tiles_per_image, m_tiles, m_pairs, k_chunks;
    int pairs;                // CTA pairs launched (tile -> (m-pair, pass) mapping keeps both passes of an m-pair on one pair)
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
    // walks the live K chunks of a tile: `rel` of `a3` chunks per channel, then jumps to the next channel
    struct ChunkWalk {
        int kc, rel;
        __device__ void begin(const PairTile& ti) { kc = ti.a2; rel = 0; }
        __device__ void next(const Params& p, const PairTile& ti) {
            ++kc;
            if (++rel == ti.a3) { rel = 0; kc += p.chunks_per_channel - ti.a3; }
        }
    };
    struct TmaState { ChunkWalk w; int n_row0; };
    struct GenState {
        int b, r_lo, rows;   // slab currently resident
        int base;            // slab offset of this thread's output cell
        ChunkWalk w;
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
    // Tile order: the kernel hands pair `q` the tiles q, q + pairs, q + 2 pairs, ...; iteration `it` of a pair is
    // pass (it % n_passes) of m-pair q + (it / n_passes) * pairs, so both N passes of an m-pair run back to back on
    // the same pair and reuse its image slab.
    // Zero-padding skip: filter row v only meets image rows for output rows i with 0 <= i + v - p < n.  For the
    // output rows of BOTH CTAs' tiles the live v range is [v_lo, v_hi); K chunks outside it multiply zeros and
    // are never generated, loaded or issued (cfg2/cfg3: ~25 % of the dense count).  Needs chunk-aligned channels.
    __device__ static void tile_info(const Params& p, int tile, uint32_t rank, PairTile& ti) {
        const ConvGeom& g = p.g;
        const int it = tile / p.pairs, q = tile - it * p.pairs;
        const int sup = it / p.n_passes, np = it - sup * p.n_passes;
        const int mp = q + sup * p.pairs;
        const int N = g.G * g.O;
        ti.n0 = np * (kAcc * kAccN);
        ti.n_acc = (N - ti.n0 > kAccN) ? 2 : 1;
        const int mt = 2 * mp + static_cast<int>(rank);
        ti.m_tile = mt < p.m_tiles ? mt : -1;
        ti.a0 = mt / p.tiles_per_image;                               // image
        ti.a1 = (mt - ti.a0 * p.tiles_per_image) * kBM;               // first position
        ti.kc_begin = 0;
        if (mp >= p.m_pairs) { ti.a2 = 0; ti.a3 = 1; ti.kc_end = 0; return; }
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
            ti.a3 = hi > lo ? hi - lo : 1;
            ti.kc_end = hi > lo ? g.C * (hi - lo) : 0;
        } else {
            ti.a2 = 0;
            ti.a3 = p.k_chunks;
            ti.kc_end = p.k_chunks;
        }
    }
    __device__ static void tma_tile_begin(const Params&, const PairTile& ti, uint32_t rank, TmaState& s) {
        s.w.begin(ti);
        s.n_row0 = ti.n0 + static_cast<int>(rank) * 128;
    }
    __device__ static void tma_chunk(const Params& p, const PairTile& ti, TmaState& s, uint32_t sb, uint32_t bar) {
        for (int a = 0; a < ti.n_acc; ++a) tma_load_2d_pair(sb + a * kBHalfBytes, &p.tmB, bar, s.w.kc * kBK, s.n_row0 + a * kAccN);
        s.w.next(p, ti);
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.b = -1; s.r_lo = 0; s.rows = 0; s.base = 0; }
    __device__ static void gen_tile_begin(const Params& p, const PairTile& ti, GenState& s, uint8_t* extra, int ptid) {
        s.w.begin(ti);
        if (ti.m_tile < 0 || ti.kc_end <= ti.kc_begin) return;
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
    __device__ static void gen_prepare(const Params&, const PairTile&, GenState&, uint8_t*, int) {}
    __device__ static void gen_advance(const Params& p, const PairTile& ti, GenState& s) { s.w.next(p, ti); }
    // one group (128 threads): thread = one A row, all 32 k of the chunk (8 swizzled 16-byte stores)
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, uint8_t* a_stage, uint8_t* extra, int gtid) {
        const int* tab = reinterpret_cast<const int*>(extra);
        const float* slab = reinterpret_cast<const float*>(extra + p.tab_entries * 4) + s.base;
        const int row = gtid;
        const bool live = ti.m_tile >= 0;
        const int kc = s.w.kc;
        if (p.gran) {
            const int4 oa = *reinterpret_cast<const int4*>(tab + kc * 8);
            const int4 ob = *reinterpret_cast<const int4*>(tab + kc * 8 + 4);
            const int o[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
            float4 v[8];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                v[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && o[ch] >= 0) {
                    const float* q = slab + o[ch];
                    v[ch] = make_float4(q[0], q[1], q[2], q[3]);
                }
            }
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<float4*>(a_stage + sw128_offset(row, ch)) = v[ch];
        } else {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const int4 o = *reinterpret_cast<const int4*>(tab + kc * kBK + ch * 4);
                float4 v;
                v.x = (live && o.x >= 0) ? slab[o.x] : 0.f;
                v.y = (live && o.y >= 0) ? slab[o.y] : 0.f;
                v.z = (live && o.z >= 0) ? slab[o.z] : 0.f;
                v.w = (live && o.w >= 0) ? slab[o.w] : 0.f;
                *reinterpret_cast<float4*>(a_stage + sw128_offset(row, ch)) = v;
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
    CUtensorMap tmQ;          // dX1 [(B*G*P)][O] as a 3-D map {32 o, rows, O/32 o-blocks}, MN-major boxes {32, 32, nb}
    int num_stages, num_tiles, m_pairs, n_passes, m_tiles, splits, chunks_total, chunks_per_split, chunks_per_image;
    int nb;                   // o-blocks (of 32 columns) per TMA box: 4, 2 or 1
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
    // walks the live position chunks of the tile's reduction range: image b, chunk lo + rel of that image
    struct ChunkWalk {
        int b, rel;
        __device__ void begin(const PairTile& ti) {
            const int cnt = ti.a2 > 0 ? ti.a2 : 1;
            b = ti.kc_begin / cnt;
            rel = ti.kc_begin - b * cnt;
        }
        // returns true when the walk moved on to the next image
        __device__ bool next(const PairTile& ti) {
            if (++rel == ti.a2) { rel = 0; ++b; return true; }
            return false;
        }
    };
    struct TmaState {
        ChunkWalk w;
        int row0;             // first dX1 row of the current chunk for rotation 0: b*G*P + (lo + rel)*32
        int rterm[kAcc][4];   // r*P per box (or a far out-of-bounds row when r >= G: TMA zero-fills)
        int oblk[kAcc][4];    // first o-block of the box
    };
    struct GenState {
        int b, m_tile;        // image / kk-tile whose slab + table are resident
        int c_lo, r_lo;
        ChunkWalk w;
        int j[2], off[2];     // per position row of this thread (rrow, rrow + 16): column j and slab offset i*pitch + j
        int pos0;             // position of row 0 of the current chunk
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
    // B-half of one accumulator: 128 (r,o) columns starting at n0 + a*256 + rank*128, reduction rows = the 32
    // positions of the chunk; 4 / nb boxes of nb o-blocks each.  Rows past the end of a (b,r) segment are multiplied
    // by generated zeros; rows past the end of the tensor (and r >= G) are zero-filled by TMA.
    __device__ static void tma_tile_begin(const Params& p, const PairTile& ti, uint32_t rank, TmaState& s) {
        const ConvGeom& g = p.g;
        s.w.begin(ti);
        s.row0 = s.w.b * g.G * g.P + (ti.a1 + s.w.rel) * kBK;
#pragma unroll
        for (int a = 0; a < kAcc; ++a) {
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int np = ti.n0 + a * kAccN + static_cast<int>(rank) * 128 + x * 32 * p.nb;
                const int r = np / g.O, o0 = np - r * g.O;
                s.rterm[a][x] = (r < g.G) ? r * g.P : 0x30000000;
                s.oblk[a][x] = o0 >> 5;
            }
        }
    }
    __device__ static void tma_chunk(const Params& p, const PairTile& ti, TmaState& s, uint32_t sb, uint32_t bar) {
        const int nbx = 4 / p.nb;
#pragma unroll
        for (int a = 0; a < kAcc; ++a) {
            if (a < ti.n_acc) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    if (x < nbx)
                        tma_load_3d_pair(sb + a * kBHalfBytes + x * p.nb * (kBK * 128), &p.tmQ, bar, 0, s.row0 + s.rterm[a][x], s.oblk[a][x]);
            }
        }
        s.row0 += kBK;
        if (s.w.next(ti)) s.row0 = s.w.b * p.g.G * p.g.P + ti.a1 * kBK;
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.b = -1; s.m_tile = -2; s.c_lo = 0; s.r_lo = 0; }
    // lane -> (position row, chunk parity): every aligned group of 8 lanes covers 4 rows x 2 adjacent 16-byte
    // chunks = 8 distinct slots of the 32-byte-atom swizzle (conflict-free STS.128), while the LDS stay within
    // 20 consecutive words.  warp of the group -> 32-wide kk block; each thread does position rows rrow and rrow + 16.
    __device__ static int lane_row(int gtid) { const int lane = gtid & 31; return (lane & 3) | (((lane >> 3) & 3) << 2); }
    __device__ static void seek_rows(const Params& p, const PairTile& ti, GenState& s, int gtid) {
        const ConvGeom& g = p.g;
        s.pos0 = (ti.a1 + s.w.rel) * kBK;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int pos = s.pos0 + lane_row(gtid) + 16 * hh;
            const int i = pos / g.d;
            s.j[hh] = pos - i * g.d;
            s.off[hh] = i * p.sg.pitch + s.j[hh];
        }
    }
    // extra smem: [128] int offsets of this CTA's kk rows, then the slab
    __device__ static void gen_tile_begin(const Params& p, const PairTile& ti, GenState& s, uint8_t* extra, int ptid) {
        s.w.begin(ti);
        if (ti.m_tile < 0 || ti.kc_end <= ti.kc_begin) return;
        seek_rows(p, ti, s, ptid & 127);
        if (s.m_tile == ti.m_tile) return;
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
    // all 8 generator warps, every chunk: refill the slab when the walk reaches a new image
    __device__ static void gen_prepare(const Params& p, const PairTile& ti, GenState& s, uint8_t* extra, int ptid) {
        if (ti.m_tile < 0 || s.b == s.w.b) return;
        const ConvGeom& g = p.g;
        float* slab = reinterpret_cast<float*>(extra + kBM * 4);
        named_bar_sync(1, kGenWarps * 32);
        const int kk_hi = min(ti.a0 + kBM, g.K) - 1;
        const Im2colCursor lo = im2col_cursor(min(ti.a0, g.K - 1), g.k), hi = im2col_cursor(kk_hi, g.k);
        const int nc = hi.c - lo.c + 1;
        const int rows = (nc == 1 ? hi.v - lo.v : g.k - 1) + g.d;      // padded rows [r_lo, r_lo + rows)
        fill_slab(slab, p.sg, g, p.y + (long long)s.w.b * g.C * g.n * g.n, lo.c, nc, s.r_lo, rows, ptid, kGenWarps * 32);
        named_bar_sync(1, kGenWarps * 32);
        s.b = s.w.b;
    }
    __device__ static void gen_advance(const Params& p, const PairTile& ti, GenState& s) {
        if (ti.m_tile < 0) return;
        if (s.w.next(ti)) {
            seek_rows(p, ti, s, (threadIdx.x & 127));
            return;
        }
        s.pos0 += kBK;
        const int d = p.g.d, wrap = p.sg.pitch - d;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            s.j[hh] += kBK; s.off[hh] += kBK;
            while (s.j[hh] >= d) { s.j[hh] -= d; s.off[hh] += wrap; }
        }
    }
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, uint8_t* a_stage, uint8_t* extra, int gtid) {
        const int* tab = reinterpret_cast<const int*>(extra);
        const float* slab = reinterpret_cast<const float*>(extra + kBM * 4);
        const int lane = gtid & 31, cb = gtid >> 5;
        const int rrow = lane_row(gtid);
        const int par = (lane >> 2) & 1;
        uint8_t* blk = a_stage + cb * (kBK * 128);
        if (ti.m_tile < 0) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    *reinterpret_cast<float4*>(blk + sw128b32_offset(rrow + 16 * hh, 2 * ch + par)) = make_float4(0.f, 0.f, 0.f, 0.f);
            return;
        }
        int4 o[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) o[ch] = *reinterpret_cast<const int4*>(tab + cb * 32 + (2 * ch + par) * 4);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const bool valid = s.pos0 + rrow + 16 * hh < p.g.P;
            const float* src = slab + s.off[hh];
            float4 v[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                v[ch].x = !valid ? 0.f : (o[ch].x >= 0 ? src[o[ch].x] : (o[ch].x == -2 ? 1.f : 0.f));
                v[ch].y = !valid ? 0.f : (o[ch].y >= 0 ? src[o[ch].y] : (o[ch].y == -2 ? 1.f : 0.f));
                v[ch].z = !valid ? 0.f : (o[ch].z >= 0 ? src[o[ch].z] : (o[ch].z == -2 ? 1.f : 0.f));
                v[ch].w = !valid ? 0.f : (o[ch].w >= 0 ? src[o[ch].w] : (o[ch].w == -2 ? 1.f : 0.f));
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
                *reinterpret_cast<float4*>(blk + sw128b32_offset(rrow + 16 * hh, 2 * ch + par)) = v[ch];
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
