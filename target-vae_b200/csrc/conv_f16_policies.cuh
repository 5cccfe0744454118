// FP16-operand CTA-pair group-convolution policies for tc_gemm2_kernel (a-2, models.py:202-225, and its weight
// gradient) as implicit GEMMs with fp16 operands:
//
//   Conv1FwdH   : X1[(b,pos), (r,o)] = lrelu(im2col . bank^T + bias)     A = im2col tile (K-major, generated), B = bank (TMA)
//   Conv1WgradH : dbank[(r,o), kk]  += sum_(b,pos) dX1[(b,r,pos), o] im2col[(b,pos), kk]
//                 A = im2col^T (MN-major, generated), B = dX1 (fp16 with a power-of-two scale, MN-major, TMA),
//                 fp32 atomics into dbank
//
// Why FP16: an fp16 value carries the same 11-bit significand a TF32 operand does, so for values inside fp16's
// normal range (6.1e-5 .. 65504; images and filter taps are) the products are the ones kind::tf32 would form, while
// kind::f16 issues at twice the TF32 rate and every operand byte (smem, L2, TMA) is halved.  A 128-byte stage row now
// holds 64 reduction elements (kBK16) and one tcgen05.mma consumes 16 of them, so the pipeline, stage sizes and
// barrier protocol of tc_gemm2_kernel are unchanged - a tile simply needs half as many stages.
//
// The im2col operand is gathered from a zero-padded image slab kept in shared memory as fp16, TWICE: copy 0 holds
// the slab, copy 1 the same halves shifted down by one element.  A run of 4 consecutive taps (k % 4 == 0) of any
// position then starts on a 4-byte boundary in one of the two copies, so a 16-byte granule of the operand (8 taps)
// costs 4 LDS.32 + 1 swizzled STS.128 with no packing arithmetic.  Copy 1 starts 16 banks after copy 0: the 32
// lanes of a warp (32 consecutive positions) read 16 consecutive words of each copy - conflict-free.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "conv2_policies.cuh"

namespace tvae {

constexpr int kBK16 = 64;   // 16-bit reduction elements per stage row

struct Slab16Geom {
    int Wp;          // padded image width n + 2p
    int pitch;       // slab row pitch in halves (even)
    int rows_max;    // slab rows allocated per channel
    int copy_words;  // 32-bit words from copy 0 to copy 1 (= 16 mod 32)
};

// slab[(c*rows_max + rr)*pitch + x] = fp16(image[c_lo + c][r_lo + rr - p][x - p]) (zero outside), written to both copies.
// (A latency-oriented variant - 8-byte loads, eight in flight per warp, padding columns zeroed once - was measured and did not
// shorten the kernels: the time the generator warps spend around a refill is the wait for the OTHER generator group at
// the refill barrier, i.e. for the MMA issuer to free a stage, not the refill itself.)
template <bool kBf16 = false>
__device__ __forceinline__ void fill_slab16(uint32_t* slabw, const Slab16Geom& sg, const ConvGeom& g, const float* img, int c_lo,
                                            int nc, int r_lo, int rows, int tid, int nthreads) {
    unsigned short* c0 = reinterpret_cast<unsigned short*>(slabw);
    unsigned short* c1 = reinterpret_cast<unsigned short*>(slabw + sg.copy_words);
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
    for (int rw = warp; rw < nc * rows; rw += nwarps) {
        const int c = rw / rows, rr = rw - c * rows;
        const int iy = r_lo + rr - g.p;
        const bool row_ok = iy >= 0 && iy < g.n;
        const float* src = img + ((long long)(c_lo + c) * g.n + (row_ok ? iy : 0)) * g.n - g.p;
        const int h0 = (c * sg.rows_max + rr) * sg.pitch;
        for (int x = lane; x < sg.pitch; x += 32) {
            const int ix = x - g.p;
            float v = 0.f;
            if (row_ok && ix >= 0 && ix < g.n) v = __ldg(src + x);
            const unsigned short hv = kBf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
            c0[h0 + x] = hv;
            if (h0 + x > 0) c1[h0 + x - 1] = hv;
        }
    }
}

// word pointer from which the 8 halves starting at slab half-index `a` can be read 4-byte aligned
__device__ __forceinline__ const uint32_t* slab16_words(const uint32_t* slabw, const Slab16Geom& sg, int a) {
    const int par = a & 1;
    return slabw + (par ? sg.copy_words : 0) + ((a - par) >> 1);
}

// conv1 forward epilogue: bias (+ LeakyReLU), then
//   * fp16 x1 (the encoder's activation) with O % 64 == 0: every 64-column block [128 positions][64 o] of the tile is a
//     contiguous 16 KB piece of x1; the 128 epilogue threads write their rows into a swizzled staging buffer (two
//     buffers, alternating) and one thread hands it to the TMA store unit - fully coalesced, asynchronous, and rows past
//     the end of the image are clipped by the tensor map;
//   * otherwise (fp32 output of GroupConv.forward alone, or O % 64 != 0) direct per-row stores.

// NV (16 or 32) accumulator columns of one row: + add[0..NV) -> activation -> fp16 -> NV / 8 swizzled 16-byte chunks, starting at
// chunk `chunk0` of the row's 128 bytes, of a staging buffer.  One warp per scheduler runs this, so it is written for
// latency: the additive row of the NEXT 8 columns is loaded before the current 8 are stored (the compiler does not move a
// shared-memory load above a shared-memory store), and LeakyReLU-or-identity is max(v, slope * v) without a branch.
// Post-processing hook of the staged epilogue: load(col) fetches what apply() needs for the 8 columns starting at col - it is
// called one 8-column group AHEAD (before the previous group's shared-memory store, which the compiler will not move a
// load across), apply(w, col, v) sees the group's fp32 values after the activation.
// packed(col0, j, q) sees the same group as 8 packed halves (what is stored), j = its static offset inside the 32-column piece.
struct NoPost {
    struct W {};
    __device__ W load(int) const { return W(); }
    __device__ void apply(const W&, int, const float (&)[8]) const {}
    __device__ void packed(int, int, const uint4&) const {}
};
// One-bit LeakyReLU derivative mask of the stored activation (bit set = derivative 1, i.e. the stored half is not negative):
// word (column / 64, row) of a [H / 64][M] array - the layout LinearNT's input-gradient epilogue reads (aux_bits).  `dst` points
// at the row's word of the accumulator's first 64-column block (null: row past M or no mask wanted); lo / acc are two
// registers of the caller.
struct MaskBitsPost {
    unsigned long long* dst;
    long long M;
    uint32_t* lo; uint32_t* acc;
    struct W {};
    __device__ W load(int) const { return W(); }
    __device__ void apply(const W&, int, const float (&)[8]) const {}
    __device__ void packed(int col0, int j, const uint4& q) const {
        *acc |= half8_sign_bits(q) << j;
        if (j == 24) {
            if (col0 & 32) {
                if (dst) dst[(long long)(col0 >> 6) * M] = ~((static_cast<unsigned long long>(*acc) << 32) | *lo);
            } else {
                *lo = *acc;
            }
            *acc = 0u;
        }
    }
};
template <bool TANH, int NV, class Post = NoPost>
__device__ __forceinline__ void epi_piece_store(const uint32_t (&rr)[NV], const float* add, bool act, uint8_t* buf, int row, int chunk0,
                                                Post post = Post(), int col0 = 0) {
    const float slope = act ? kLreluSlope : 1.f;
    float4 b0 = *reinterpret_cast<const float4*>(add), b1 = *reinterpret_cast<const float4*>(add + 4);
    typename Post::W pw = post.load(col0);
#pragma unroll
    for (int j = 0; j < NV; j += 8) {
        float4 n0 = b0, n1 = b1;
        typename Post::W pn = pw;
        if (j + 8 < NV) {
            n0 = *reinterpret_cast<const float4*>(add + j + 8);
            n1 = *reinterpret_cast<const float4*>(add + j + 12);
            pn = post.load(col0 + j + 8);
        }
        float v[8];
        v[0] = __uint_as_float(rr[j]) + b0.x;     v[1] = __uint_as_float(rr[j + 1]) + b0.y;
        v[2] = __uint_as_float(rr[j + 2]) + b0.z; v[3] = __uint_as_float(rr[j + 3]) + b0.w;
        v[4] = __uint_as_float(rr[j + 4]) + b1.x; v[5] = __uint_as_float(rr[j + 5]) + b1.y;
        v[6] = __uint_as_float(rr[j + 6]) + b1.z; v[7] = __uint_as_float(rr[j + 7]) + b1.w;
        if (TANH) {
            if (act) act_vec<true>(v);
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], slope * v[q]);
        }
        post.apply(pw, col0 + j, v);
        uint4 q4;
        __half2 hv;
        hv = __floats2half2_rn(v[0], v[1]); q4.x = *reinterpret_cast<uint32_t*>(&hv);
        hv = __floats2half2_rn(v[2], v[3]); q4.y = *reinterpret_cast<uint32_t*>(&hv);
        hv = __floats2half2_rn(v[4], v[5]); q4.z = *reinterpret_cast<uint32_t*>(&hv);
        hv = __floats2half2_rn(v[6], v[7]); q4.w = *reinterpret_cast<uint32_t*>(&hv);
        *reinterpret_cast<uint4*>(buf + sw128_offset(row, chunk0 + (j >> 3))) = q4;
        post.packed(col0, j, q4);
        b0 = n0; b1 = n1;
        pw = pn;
    }
}

// State of an epilogue that leaves through staged TMA stores (tc_gemm2's store issuer, warp 3)
struct StagedEpiState {
    uint32_t blocks;                    // 64-column blocks written so far by this CTA: buffer blocks % NBUF, use blocks / NBUF
    uint32_t staged_bar, freed_bar;     // mbarrier arrays of the kernel
    int sel, b_lo, b_hi;                // policy scratch
    float proj[4];                      // policy scratch (fused output projection of the row)
};

// `nblk` 64-column blocks of one accumulator: value + add -> act -> fp16 -> swizzled staging buffer (ring of NBUF); the store
// issuer warp hands each finished block to the TMA store unit.  The TMEM loads are software-pipelined in 32-column pieces
// (the next piece is in flight while this one is converted).  Per block: wait until the store that last used the buffer
// has read it (freed[b]), write, publish to the async proxy, one arrive per warp on staged[b].  No named barrier: a warp
// never waits for the other epilogue warps, only for a free buffer.   add_ptr(blk) -> 64 floats (shared memory).
template <bool TANH, int NBUF, int PROBE_SLOT, class AddPtr, class Post = NoPost>
__device__ __forceinline__ void staged_store_epilogue(uint32_t taddr, int nblk, StagedEpiState& st, uint8_t* stage0, int row, bool act,
                                                      AddPtr add_ptr, Post post = Post()) {
    if (nblk <= 0) return;
    uint32_t ra[32], rb[32];
#ifdef TVAE_PROBE
    long long pr[6] = {0, 0, 0, 0, 0, 0}, pt = clock64(), pn;
#define TVAE_EPI_LAP(i) pn = clock64(); pr[i] += pn - pt; pt = pn
#else
#define TVAE_EPI_LAP(i)
#endif
    tmem_ld_32x32(taddr, ra);
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
        const uint32_t b = st.blocks % NBUF, use = st.blocks / NBUF;
        uint8_t* buf = stage0 + b * kStoreBlockBytes;
        const float* add = add_ptr(blk);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + blk * 64 + 32, rb);
        TVAE_EPI_LAP(0);
        if (use >= 1) mbar_wait(st.freed_bar + 8 * b, (use - 1) & 1);
        TVAE_EPI_LAP(3);
        epi_piece_store<TANH, 32>(ra, add, act, buf, row, 0, post, blk * 64);
        TVAE_EPI_LAP(1);
        tmem_ld_wait();
        if (blk + 1 < nblk) tmem_ld_32x32(taddr + (blk + 1) * 64, ra);
        TVAE_EPI_LAP(0);
        epi_piece_store<TANH, 32>(rb, add + 32, act, buf, row, 4, post, blk * 64 + 32);
        TVAE_EPI_LAP(1);
        fence_proxy_async_smem();
        __syncwarp();
        if ((row & 31) == 0) mbar_arrive(st.staged_bar + 8 * b);
        TVAE_EPI_LAP(2);
        ++st.blocks;
    }
#ifdef TVAE_PROBE
    if (row == 0 && cluster_ctarank() == 0)
        for (int i = 0; i < 6; ++i) atomicAdd(&g_pair_probe[PROBE_SLOT][8 + i], (unsigned long long)pr[i]);
#endif
#undef TVAE_EPI_LAP
}

// Weight-gradient epilogue: `npieces` 32-column pieces of one accumulator, value * scale -> fp32 staging tile [32 columns][128
// rows] (ring of NBUF 16 KB buffers) -> the store issuer adds the tile to global memory with ONE TMA reduce
// (cp.reduce.async.bulk.tensor .add) instead of 4096 red.global.add.f32 issued by the SM: the atomics epilogue held the
// accumulator 38 k clocks per tile at cfg2 (clock64 probe), 20 % of the conv1 weight-gradient kernel and 40 % at cfg1.
// Rows / columns past the matrix are clipped by the tensor map.  Same buffer protocol as staged_store_epilogue.
template <int NBUF>
__device__ __forceinline__ void staged_reduce_epilogue(uint32_t taddr, int npieces, StagedEpiState& st, uint8_t* stage0, int row, float scale) {
    if (npieces <= 0) return;
    uint32_t ra[32], rb[32];
    tmem_ld_32x32(taddr, ra);
    auto piece = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int c) {
        const uint32_t b = st.blocks % NBUF, use = st.blocks / NBUF;
        float* buf = reinterpret_cast<float*>(stage0 + b * kStoreBlockBytes);
        tmem_ld_wait();
        if (c + 1 < npieces) tmem_ld_32x32(taddr + (c + 1) * 32, nxt);
        if (use >= 1) mbar_wait(st.freed_bar + 8 * b, (use - 1) & 1);
#pragma unroll
        for (int j = 0; j < 32; ++j) buf[j * kBM + row] = __uint_as_float(cur[j]) * scale;     // lanes = consecutive words: conflict-free
        fence_proxy_async_smem();
        __syncwarp();
        if ((row & 31) == 0) mbar_arrive(st.staged_bar + 8 * b);
        ++st.blocks;
    };
#pragma unroll 1
    for (int c = 0; c < npieces; c += 2) {
        piece(ra, rb, c);
        if (c + 1 < npieces) piece(rb, ra, c + 1);
    }
}

// 64-column blocks of the accumulator starting at column n0 that leave through staged TMA stores (uniform over the CTA)
template <class Prm>
__device__ __forceinline__ int conv1_store_blocks(const Prm& p, const PairTile& ti, int n0, bool has_work) {
    if (!p.tma_store || !has_work || ti.m_tile < 0) return 0;
    return min(kAccN / 64, (p.g.G * p.g.O - n0) / 64);
}

template <bool TANH, class Prm>
__device__ __forceinline__ void conv1_fwd_epilogue(const Prm& p, const PairTile& ti, StagedEpiState& st, int n0, uint32_t taddr, int row,
                                                   bool has_work, uint8_t* extra) {
    const ConvGeom& g = p.g;
    const int pos = ti.a1 + row;
    const int N = g.G * g.O;
    const float* s_bias = reinterpret_cast<const float*>(extra + p.bias_off);
    if (p.tma_store) {
        const int o_first = n0 % g.O;                               // O % 64 == 0: a block lies inside one rotation's O columns
        staged_store_epilogue<TANH, 2, 1>(taddr, conv1_store_blocks(p, ti, n0, has_work), st, extra + p.stage_off, row, p.act != 0,
                                          [&](int blk) { int o = o_first + blk * 64; while (o >= g.O) o -= g.O; return s_bias + o; });
        return;
    }
    const bool ok = has_work && ti.m_tile >= 0 && pos < g.P;
#pragma unroll 1
    for (int c = 0; c < kAccN / 32; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(taddr + c * 32, rr);
        tmem_ld_wait();
        const int np = n0 + c * 32;
        if (!ok || np >= N) continue;
        const int r = np / g.O, o0 = np - r * g.O;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) + s_bias[o0 + j];
        if (p.act) act_vec<TANH>(v);
        const long long off = (((long long)ti.a0 * g.G + r) * g.P + pos) * g.O + o0;
        if (p.x1h) {
            __half* dst = p.x1h + off;
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
        } else {
            float* dst = p.x1 + off;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
struct Conv1FwdHParams {
    CUtensorMap tmB;          // bank fp16 [G*O][kpad16], boxes {64 k, 128 rows}
    CUtensorMap tmX;          // x1 fp16 as [B*G][P][O] for the epilogue's TMA stores (tma_store == 1)
    int num_stages, num_tiles, n_passes, tiles_per_image, m_tiles, m_pairs, k_chunks;
    int pairs;                // CTA pairs launched (both passes of an m-pair stay on one pair)
    int tma_store;            // 1: fp16 output through staged TMA stores
    int bias_off, stage_off;  // byte offsets inside the policy's extra smem: [O] bias floats, 2 staging buffers
    ConvGeom g;
    Slab16Geom sg;
    const float* y;           // (B,C,n,n)
    const float* bias;        // (O) or null
    float* x1;                // fp32 [(b*G + r)*P + pos][O] or null
    __half* x1h;              // fp16 output in the same layout or null
    int act;
    int quad;                 // 1: offset table per 4-tap quad (k % 4 == 0), 0: per tap
    int tab_entries;
    int skip;                 // 1: skip K chunks that only meet zero padding (needs k*k % 64 == 0)
    int chunks_per_channel;   // k*k / 64 when skip, else k_chunks
    int per_channel;          // 1 (C > 1 with chunk-aligned channels): the slab holds ONE channel and is refilled when the
                              // chunk walk reaches the next channel; the offset table is channel-relative
};

template <bool TANH>
struct Conv1FwdHT : PolicyBase {
    static constexpr const char* kName = "conv1_fwd";
    static constexpr int kProbeSlot = 1;
    using Params = Conv1FwdHParams;
    static constexpr bool kF16 = true;
    struct ChunkWalk {
        int kc, rel, c;      // K chunk, chunk index inside the channel's live range, channel
        __device__ void begin(const PairTile& ti) { kc = ti.a2; rel = 0; c = 0; }
        __device__ void next(const Params& p, const PairTile& ti) {
            ++kc;
            if (++rel == ti.a3) { rel = 0; kc += p.chunks_per_channel - ti.a3; ++c; }
        }
    };
    struct TmaState { ChunkWalk w; int n_row0; };
    struct GenState {
        int b, r_lo, rows, c;     // slab currently resident (c: channel, per-channel mode only)
        int t_r_lo, t_rows;       // row window the current tile needs
        int base;                 // slab half-index of this thread's output cell
        ChunkWalk w;
    };
    using EpiState = StagedEpiState;
    static constexpr int kStoreBufs = 2;      // staging buffers of the x1 stores (tc_gemm2's store issuer)
    __device__ static void epi_init(const Params&, EpiState& st, uint8_t*, int) { st.blocks = 0; st.sel = 0; }
    __device__ static int store_off(const Params& p) { return p.stage_off; }
    __device__ static int store_blocks(const Params& p, const PairTile& ti, int n0, bool has_work) { return conv1_store_blocks(p, ti, n0, has_work); }
    __device__ static void store_issue(const Params& p, const PairTile& ti, int n0, int blk, uint32_t src) {
        const int np = n0 + blk * 64;
        const int r = np / p.g.O;
        tma_store_3d(&p.tmX, src, np - r * p.g.O, ti.a1, ti.a0 * p.g.G + r);
    }
    __device__ static void prefetch_descs(const Params& p) {
        tma_prefetch_desc(&p.tmB);
        if (p.tma_store) tma_prefetch_desc(&p.tmX);
    }
    // extra smem: [tab_entries] int offsets (quad mode: 32-bit word offsets; tap mode: half offsets), then the two slab copies
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        int* tab = reinterpret_cast<int*>(extra);
        const ConvGeom& g = p.g;
        const int step = p.quad ? 4 : 1;
        for (int e = tid; e < p.tab_entries; e += nthreads) {
            const int kk = e * step;
            int off = -1;
            if (kk < g.K) {
                const Im2colCursor cur = im2col_cursor(kk, g.k);
                off = (cur.c * p.sg.rows_max + cur.v) * p.sg.pitch + cur.u;
                if (p.quad) off >>= 1;
            }
            tab[e] = off;
        }
        float* s_bias = reinterpret_cast<float*>(extra + p.bias_off);
        for (int o = tid; o < g.O; o += nthreads) s_bias[o] = p.bias ? __ldg(p.bias + o) : 0.f;
    }
    // Tile order: the kernel hands pair `q` the tiles q, q + pairs, q + 2 pairs, ...; iteration `it` of a pair is
    // pass (it % n_passes) of m-pair q + (it / n_passes) * pairs, so both N passes of an m-pair run back to back on
    // the same pair and reuse its image slab.
    // Zero-padding skip: filter row v only meets image rows for output rows i with 0 <= i + v - p < n.  For the
    // output rows of BOTH CTAs' tiles the live v range is [v_lo, v_hi); 64-tap K chunks outside it multiply zeros
    // and are never generated, loaded or issued (cfg2/cfg3: ~25 % of the dense count).  Needs chunk-aligned channels.
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
            const int lo = (v_lo * g.k) / kBK16, hi = (v_hi * g.k + kBK16 - 1) / kBK16;   // chunks within one channel
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
        for (int a = 0; a < ti.n_acc; ++a) tma_load_2d_pair(sb + a * kBHalfBytes, &p.tmB, bar, s.w.kc * kBK16, s.n_row0 + a * kAccN);
        s.w.next(p, ti);
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) {
        s.b = -1; s.r_lo = 0; s.rows = 0; s.c = -1; s.t_r_lo = 0; s.t_rows = 0; s.base = 0;
    }
    __device__ static void gen_tile_begin(const Params& p, const PairTile& ti, GenState& s, uint8_t* extra, int ptid) {
        s.w.begin(ti);
        if (ti.m_tile < 0 || ti.kc_end <= ti.kc_begin) return;
        const ConvGeom& g = p.g;
        uint32_t* slabw = reinterpret_cast<uint32_t*>(extra + p.tab_entries * 4);
        const int i0 = ti.a1 / g.d;
        const int last = min(ti.a1 + kBM - 1, g.P - 1);
        const int i1 = last / g.d;
        const int r_lo = i0, rows = i1 - i0 + g.k;                   // padded rows [i0, i1 + k)
        s.t_r_lo = r_lo; s.t_rows = rows;
        if (p.per_channel) {
            // the slab is (re)filled channel by channel in gen_prepare
        } else if (s.b != ti.a0 || s.r_lo != r_lo || s.rows != rows) {       // uniform across the generator warps
            named_bar_sync(1, kGenWarps * 32);                        // previous tile's gathers are done
            fill_slab16(slabw, p.sg, g, p.y + (long long)ti.a0 * g.C * g.n * g.n, 0, g.C, r_lo, rows, ptid, kGenWarps * 32);
            named_bar_sync(1, kGenWarps * 32);
            s.b = ti.a0; s.r_lo = r_lo; s.rows = rows;
        }
        const int pos = min(ti.a1 + (ptid & (kBM - 1)), g.P - 1);     // rows past the image end are discarded by the epilogue
        const int i = pos / g.d, j = pos - i * g.d;
        s.base = (i - r_lo) * p.sg.pitch + j;
    }
    // all 8 generator warps, every chunk (per-channel mode): refill the one-channel slab when the walk reaches a
    // chunk of another channel / image / row window.  Both groups walk every chunk, so the condition is uniform.
    __device__ static void gen_prepare(const Params& p, const PairTile& ti, GenState& s, uint8_t* extra, int ptid) {
        if (!p.per_channel || ti.m_tile < 0) return;
        if (s.b == ti.a0 && s.r_lo == s.t_r_lo && s.rows == s.t_rows && s.c == s.w.c) return;
        const ConvGeom& g = p.g;
        uint32_t* slabw = reinterpret_cast<uint32_t*>(extra + p.tab_entries * 4);
        named_bar_sync(1, kGenWarps * 32);                            // gathers from the previous channel are done
        fill_slab16(slabw, p.sg, g, p.y + (long long)ti.a0 * g.C * g.n * g.n, s.w.c, 1, s.t_r_lo, s.t_rows, ptid, kGenWarps * 32);
        named_bar_sync(1, kGenWarps * 32);
        s.b = ti.a0; s.r_lo = s.t_r_lo; s.rows = s.t_rows; s.c = s.w.c;
    }
    __device__ static void gen_advance(const Params& p, const PairTile& ti, GenState& s) { s.w.next(p, ti); }
    // one group (128 threads): thread = one A row, all 64 taps of the chunk (8 swizzled 16-byte stores)
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, uint8_t* a_stage, uint8_t* extra, int gtid) {
        const int* tab = reinterpret_cast<const int*>(extra);
        const uint32_t* slabw = reinterpret_cast<const uint32_t*>(extra + p.tab_entries * 4);
        const int row = gtid;
        const bool live = ti.m_tile >= 0;
        const int kc = p.per_channel ? ti.a2 + s.w.rel : s.w.kc;      // table row: channel-relative in per-channel mode
        if (p.quad) {
            const uint32_t* src = slab16_words(slabw, p.sg, s.base);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int4 oa = *reinterpret_cast<const int4*>(tab + kc * 16 + hf * 8);
                const int4 ob = *reinterpret_cast<const int4*>(tab + kc * 16 + hf * 8 + 4);
                const int o[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
                uint32_t w[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    w[2 * q] = 0u; w[2 * q + 1] = 0u;
                    if (live && o[q] >= 0) { w[2 * q] = src[o[q]]; w[2 * q + 1] = src[o[q] + 1]; }
                }
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    *reinterpret_cast<uint4*>(a_stage + sw128_offset(row, hf * 4 + ch)) = make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
            }
        } else {
            const unsigned short* s0 = reinterpret_cast<const unsigned short*>(slabw) + s.base;
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int o0 = tab[kc * kBK16 + ch * 8 + 2 * e], o1 = tab[kc * kBK16 + ch * 8 + 2 * e + 1];
                    const uint32_t lo = (live && o0 >= 0) ? s0[o0] : 0u;
                    const uint32_t hi = (live && o1 >= 0) ? s0[o1] : 0u;
                    w[e] = lo | (hi << 16);
                }
                *reinterpret_cast<uint4*>(a_stage + sw128_offset(row, ch)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    __device__ static void epilogue(const Params& p, const PairTile& ti, EpiState& st, int n0, uint32_t taddr, int row, bool has_work,
                                    uint8_t* extra) {
        conv1_fwd_epilogue<TANH>(p, ti, st, n0, taddr, row, has_work, extra);
    }
};
using Conv1FwdH = Conv1FwdHT<false>;

// ------------------------------------------------------------------------------------------------
struct Conv1WgradHParams {
    CUtensorMap tmQ;          // dX1 fp16 (scaled) [(B*G*P)][O] as a 3-D map {32 o, rows, O/32 o-blocks}, MN-major boxes {32, 64, nb}
    int num_stages, num_tiles, m_pairs, n_passes, m_tiles, splits, chunks_total, chunks_per_split, chunks_per_image;
    int nb;                   // o-blocks (of 32 columns) per TMA box: 4, 2 or 1
    ConvGeom g;
    Slab16Geom sg;
    const float* y;
    float* dbank;             // [G*O][kpad] fp32, zero-filled by the caller
    const float* acc_scale;   // device scalar: 1 / (scale dX1 was stored with)
    CUtensorMap tmD;          // dbank as [G*O][K] fp32 (pitch kpad): target of the epilogue's TMA reduce-adds (tma_reduce == 1)
    int tma_reduce, stage_off;   // stage_off: byte offset of the two fp32 staging tiles in the extra smem
    int quad;                 // 1: offset table per 4-tap quad, 0: per tap
    int skip;                 // 1: skip position chunks that only meet zero padding
};

// B128 = true (O % 64 == 0): dX1 is staged in 64-column blocks with the 128 B swizzle (tmQ from make_tmap_3d_mn128_h, p.nb
// counts 64-column blocks per box); false: 32-column blocks / 64 B swizzle.
// NB = staging tiles of the TMA-reduce epilogue: 2, or 1 where the second would cost a pipeline stage (cfg4 / cfg5 slabs).
template <bool B128, int NB = 2>
struct Conv1WgradHT : PolicyBase {
    static constexpr const char* kName = "conv1_wgrad";
    static constexpr int kProbeSlot = 2;
    static constexpr bool kBSw128MN = B128;
    static constexpr int kBlkCols = B128 ? 64 : 32;            // columns per staged block
    static constexpr int kBoxes = 128 / kBlkCols;              // blocks per 128-column accumulator half
    using Params = Conv1WgradHParams;
    static constexpr bool kF16 = true;
    // both operands fp16 (the tensor core rejects mixed fp16 x bf16 operands: illegal instruction).  dX1 is stored
    // multiplied by a power-of-two scale that keeps it inside fp16's range; the epilogue divides it out (acc_scale).
    static constexpr bool kAMajorMN = true;
    static constexpr bool kBMajorMN = true;
    struct ChunkWalk {
        int b, rel;
        __device__ void begin(const PairTile& ti) {
            const int cnt = ti.a2 > 0 ? ti.a2 : 1;
            b = ti.kc_begin / cnt;
            rel = ti.kc_begin - b * cnt;
        }
        __device__ bool next(const PairTile& ti) {
            if (++rel == ti.a2) { rel = 0; ++b; return true; }
            return false;
        }
    };
    struct TmaState {
        ChunkWalk w;
        int row0;             // first dX1 row of the current chunk for rotation 0: b*G*P + (lo + rel)*64
        int rterm[kAcc][4];   // r*P per box (or a far out-of-bounds row when r >= G: TMA zero-fills)
        int oblk[kAcc][4];    // first o-block (of kBlkCols columns) of the box
    };
    struct GenState {
        int b, m_tile;        // image / kk-tile whose slab + table are resident
        int c_lo, r_lo;
        ChunkWalk w;
        int j, off;           // this thread's position row of the current chunk: column j, slab half-index i*pitch + j
        int pos0;             // position of row 0 of the current chunk
    };
    __device__ static void prefetch_descs(const Params& p) { tma_prefetch_desc(&p.tmQ); }
    // Zero-padding skip: accumulator rows kk of the pair cover filter rows [va, vb]; only output rows i with
    // 0 <= i + v - p < n for some such v contribute, i.e. a contiguous range of 64-position chunks per image.
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
        const int kk0 = 2 * mp * kBM, kk1 = min(kk0 + 2 * kBM, g.K) - 1;   // kk range of the pair
        if (p.skip) {
            const Im2colCursor c0 = im2col_cursor(kk0, g.k), c1 = im2col_cursor(kk1, g.k);
            if (c0.c == c1.c) {
                const int i_lo = max(0, g.p - c1.v), i_hi = min(g.d - 1, g.p - c0.v + g.n - 1);
                if (i_hi >= i_lo) {
                    lo = (i_lo * g.d) / kBK16;
                    cnt = ((i_hi + 1) * g.d + kBK16 - 1) / kBK16 - lo;
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
    __device__ static void tma_tile_begin(const Params& p, const PairTile& ti, uint32_t rank, TmaState& s) {
        const ConvGeom& g = p.g;
        s.w.begin(ti);
        s.row0 = s.w.b * g.G * g.P + (ti.a1 + s.w.rel) * kBK16;
#pragma unroll
        for (int a = 0; a < kAcc; ++a) {
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int np = ti.n0 + a * kAccN + static_cast<int>(rank) * 128 + x * kBlkCols * p.nb;
                const int r = np / g.O, o0 = np - r * g.O;
                s.rterm[a][x] = (r < g.G) ? r * g.P : 0x30000000;
                s.oblk[a][x] = o0 / kBlkCols;
            }
        }
    }
    __device__ static void tma_chunk(const Params& p, const PairTile& ti, TmaState& s, uint32_t sb, uint32_t bar) {
        const int nbx = kBoxes / p.nb;
#pragma unroll
        for (int a = 0; a < kAcc; ++a) {
            if (a < ti.n_acc) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    if (x < nbx)
                        tma_load_3d_pair(sb + a * kBHalfBytes + x * p.nb * (kBK16 * kBlkCols * 2), &p.tmQ, bar, 0, s.row0 + s.rterm[a][x], s.oblk[a][x]);
            }
        }
        s.row0 += kBK16;
        if (s.w.next(ti)) s.row0 = s.w.b * p.g.G * p.g.P + ti.a1 * kBK16;
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.b = -1; s.m_tile = -2; s.c_lo = 0; s.r_lo = 0; }
    // group thread -> (position row of the 64-row chunk, 64-wide kk block): warp w of the group covers position rows
    // (w & 1)*32 + lane of kk block w >> 1.  The 8 lanes of an STS.128 phase write the same 16-byte chunk index of 8
    // consecutive rows = 8 distinct slots of the 128 B swizzle; their LDS hit consecutive slab words.
    __device__ static int pos_row(int gtid) { return ((gtid >> 5) & 1) * 32 + (gtid & 31); }
    __device__ static void seek_rows(const Params& p, const PairTile& ti, GenState& s, int gtid) {
        const ConvGeom& g = p.g;
        s.pos0 = (ti.a1 + s.w.rel) * kBK16;
        const int pos = s.pos0 + pos_row(gtid);
        const int i = pos / g.d;
        s.j = pos - i * g.d;
        s.off = i * p.sg.pitch + s.j;
    }
    // extra smem: [128] int offsets of this CTA's kk rows (quad mode: first 32 entries, word offsets), then the slab copies
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
        const int step = p.quad ? 4 : 1;
        if (ptid < kBM / step) {
            const int kk = ti.a0 + ptid * step;
            int off = -1;                       // zero row
            if (kk < g.K) {
                const Im2colCursor cur = im2col_cursor(kk, g.k);
                off = ((cur.c - lo.c) * p.sg.rows_max + (cur.v - v_lo)) * p.sg.pitch + cur.u;
                if (p.quad) off >>= 1;
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
        uint32_t* slabw = reinterpret_cast<uint32_t*>(extra + kBM * 4);
        named_bar_sync(1, kGenWarps * 32);
        const int kk_hi = min(ti.a0 + kBM, g.K) - 1;
        const Im2colCursor lo = im2col_cursor(min(ti.a0, g.K - 1), g.k), hi = im2col_cursor(kk_hi, g.k);
        const int nc = hi.c - lo.c + 1;
        const int rows = (nc == 1 ? hi.v - lo.v : g.k - 1) + g.d;      // padded rows [r_lo, r_lo + rows)
        fill_slab16(slabw, p.sg, g, p.y + (long long)s.w.b * g.C * g.n * g.n, lo.c, nc, s.r_lo, rows, ptid, kGenWarps * 32);
        named_bar_sync(1, kGenWarps * 32);
        s.b = s.w.b;
    }
    __device__ static void gen_advance(const Params& p, const PairTile& ti, GenState& s) {
        if (ti.m_tile < 0) return;
        if (s.w.next(ti)) {
            seek_rows(p, ti, s, (threadIdx.x & 127));
            return;
        }
        s.pos0 += kBK16;
        const int d = p.g.d, wrap = p.sg.pitch - d;
        s.j += kBK16; s.off += kBK16;
        while (s.j >= d) { s.j -= d; s.off += wrap; }
    }
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, uint8_t* a_stage, uint8_t* extra, int gtid) {
        const int* tab = reinterpret_cast<const int*>(extra);
        const uint32_t* slabw = reinterpret_cast<const uint32_t*>(extra + kBM * 4);
        const int prow = pos_row(gtid), blk = gtid >> 6;
        uint8_t* dst = a_stage + blk * (kBK16 * 128);
        const bool valid = ti.m_tile >= 0 && s.pos0 + prow < p.g.P;
        if (p.quad) {
            const uint32_t* src = slab16_words(slabw, p.sg, s.off);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int4 oa = *reinterpret_cast<const int4*>(tab + blk * 16 + hf * 8);
                const int4 ob = *reinterpret_cast<const int4*>(tab + blk * 16 + hf * 8 + 4);
                const int o[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
                uint32_t w[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    w[2 * q] = 0u; w[2 * q + 1] = 0u;
                    if (valid && o[q] >= 0) { w[2 * q] = src[o[q]]; w[2 * q + 1] = src[o[q] + 1]; }
                }
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    *reinterpret_cast<uint4*>(dst + sw128_offset(prow, hf * 4 + ch)) = make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
            }
        } else {
            const unsigned short* s0 = reinterpret_cast<const unsigned short*>(slabw) + s.off;
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int o0 = tab[blk * 64 + ch * 8 + 2 * e], o1 = tab[blk * 64 + ch * 8 + 2 * e + 1];
                    const uint32_t lo = (valid && o0 >= 0) ? s0[o0] : 0u;
                    const uint32_t hi = (valid && o1 >= 0) ? s0[o1] : 0u;
                    w[e] = lo | (hi << 16);
                }
                *reinterpret_cast<uint4*>(dst + sw128_offset(prow, ch)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    using EpiState = StagedEpiState;
    static constexpr int kStoreBufs = NB;
    __device__ static void epi_init(const Params&, EpiState& st, uint8_t*, int) { st.blocks = 0; st.sel = 0; }
    __device__ static int store_off(const Params& p) { return p.stage_off; }
    // 32-column pieces of the accumulator starting at column n0 that leave through TMA reduce-adds (uniform over the CTA)
    __device__ static int store_blocks(const Params& p, const PairTile& ti, int n0, bool has_work) {
        if (!p.tma_reduce || !has_work || ti.m_tile < 0) return 0;
        return min(kAccN / 32, (p.g.G * p.g.O - n0 + 31) / 32);
    }
    __device__ static void store_issue(const Params& p, const PairTile& ti, int n0, int blk, uint32_t src) {
        tma_reduce_add_2d(&p.tmD, src, ti.a0, n0 + blk * 32);
    }
    __device__ static void epilogue(const Params& p, const PairTile& ti, EpiState& st, int n0, uint32_t taddr, int row, bool has_work,
                                    uint8_t* extra) {
        if (p.tma_reduce) {
            staged_reduce_epilogue<NB>(taddr, store_blocks(p, ti, n0, has_work), st, extra + p.stage_off, row, __ldg(p.acc_scale));
            return;
        }
        const ConvGeom& g = p.g;
        const int kk = ti.a0 + row;
        const int N = g.G * g.O;
        const bool ok = has_work && ti.m_tile >= 0 && kk < g.K;      // per thread: the warp-collective TMEM loads stay unconditional
        const float acc_scale = __ldg(p.acc_scale);
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
                if (np < N) atomicAdd(p.dbank + (long long)np * g.kpad + kk, __uint_as_float(rr[j]) * acc_scale);
            }
        }
    }
};

using Conv1WgradH = Conv1WgradHT<false>;
using Conv1WgradH128 = Conv1WgradHT<true>;
using Conv1WgradH_1 = Conv1WgradHT<false, 1>;
using Conv1WgradH128_1 = Conv1WgradHT<true, 1>;

}  // namespace tvae
