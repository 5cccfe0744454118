// Generator layer-1 weight gradient on the CTA-pair kernel (tc_gemm2_kernel):
//
//   dW1[j][f] = sum_m dpre[m][j] * cos(phase(m, f))          (models.py:53-58, 112 backward)
//
// accumulator rows = features f (256 per pair, 128 per CTA), columns = ALL hidden units j (two N = 256 accumulators
// fill TMEM), reduction over the B*n^2 pixels in chunks of 64.
//
// Why the pair kernel: the Fourier-feature operand is synthesised by the SM (one MUFU.COS per element, 16 per clock per
// SM).  With one 128 x 256 accumulator per CTA (GenL1Wgrad<256>) a [64 px x 128 f] operand chunk costs 512 clocks of
// MUFU and feeds only 512 clocks of tensor-core work, and every chunk is generated twice (once per 256-column half of
// H = 512): the kernel sat at 22 % tensor-pipe activity, generator-bound.  Here one generated chunk feeds 128 x 512
// accumulator columns (1024 clocks of MMA per SM), cta_group::2 halves the dpre bytes each SM stages, and each feature
// is evaluated once per pixel.
//
// A = feat^T generated MN-major in the 128 B swizzle (two 64-feature blocks [64 px][128 B] per stage),
// B = dpre (fp16 with a power-of-two scale) by 3-D TMA, MN-major, 64 B swizzle; fp32 atomics into dW1.
#pragma once
#include "conv_f16_policies.cuh"
#include "gen_policies.cuh"

namespace tvae {

struct GenL1WgradPairParams {
    CUtensorMap tmQ;          // dpre fp16 [M][H] as {32 j, M rows, H/32 j-blocks}, boxes {32, 64, 4}
    int num_stages, num_tiles, m_pairs, m_tiles, splits, chunks_total, chunks_per_split;
    CoordXform cx;
    const float* wf_scaled; const float* bf;
    int E, H;
    float* dW1;               // [H][E], zero-filled by the caller
    const float* acc_scale;   // device scalar: 1 / (scale dpre was stored with)
    CUtensorMap tmD;          // dW1 as [H][E] fp32: target of the epilogue's TMA reduce-adds (tma_reduce == 1)
    int tma_reduce, stage_off;   // stage_off: byte offset of the two fp32 staging tiles in the extra smem
    const float* zbc;         // FEAT = 1: latent bias (B,E) of the coordinate layer whose activation is the generated operand
};

// FEAT = 0: Fourier features cos(phase).  FEAT = 1 (generators WITHOUT Fourier features, cfg2): the generated operand is the
// first-layer activation itself, a0[m][f] = LeakyReLU(W1[f] . x'_m + b1[f] + zb[image(m)][f]) - two FMAs per element -, so the
// weight gradient of the first hidden layer, dWh[j][f] = sum_m dpre1[m][j] a0[m][f], needs no stored a0 (wf_scaled / bf then
// hold W1 (E,2) / b1 (E), E = H features); see GenL1FwdPairT<., 1>.
template <int FEAT = 0>
struct GenL1WgradPairT : PolicyBase {
    static constexpr const char* kName = "gen_l1_wgrad";
    static constexpr int kProbeSlot = 4;
    using Params = GenL1WgradPairParams;
    static constexpr bool kF16 = true;
    static constexpr bool kAMajorMN = true;
    static constexpr bool kBMajorMN = true;
    struct TmaState { int row0; int jblk[kAcc]; };
    struct GenState {
        int q;                // current chunk
        int b;                // image whose rotation / shift are cached (-1: none)
        float cs, sn, dx0, dx1;
        int pq, pb;           // prefetched coordinate: chunk it belongs to (-1: none), its image
        float2 pv;
    };
    __device__ static void prefetch_descs(const Params& p) { tma_prefetch_desc(&p.tmQ); }
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        load_fourier_table(reinterpret_cast<float4*>(extra), p.wf_scaled, p.bf, p.E, tid, nthreads);
    }
    __device__ static void tile_info(const Params& p, int tile, uint32_t rank, PairTile& ti) {
        const int sp = tile / p.m_pairs;
        const int mp = tile - sp * p.m_pairs;
        ti.n0 = 0;
        ti.n_acc = p.H > kAccN ? 2 : 1;
        const int mt = 2 * mp + static_cast<int>(rank);
        ti.m_tile = mt < p.m_tiles ? mt : -1;
        ti.a0 = mt * kBM;                                  // first feature of this CTA's accumulator rows
        ti.kc_begin = min(sp * p.chunks_per_split, p.chunks_total);
        ti.kc_end = min(ti.kc_begin + p.chunks_per_split, p.chunks_total);
        ti.a1 = ti.a2 = ti.a3 = 0;
    }
    __device__ static void tma_tile_begin(const Params&, const PairTile& ti, uint32_t rank, TmaState& s) {
        s.row0 = ti.kc_begin * kBKh;
#pragma unroll
        for (int a = 0; a < kAcc; ++a) s.jblk[a] = (a * kAccN + static_cast<int>(rank) * 128) >> 5;
    }
    // this CTA's half of each accumulator's B tile: 128 hidden columns x 64 pixel rows (columns >= H: TMA zero fill)
    __device__ static void tma_chunk(const Params& p, const PairTile& ti, TmaState& s, uint32_t sb, uint32_t bar) {
#pragma unroll
        for (int a = 0; a < kAcc; ++a)
            if (a < ti.n_acc) tma_load_3d_pair(sb + a * kBHalfBytes, &p.tmQ, bar, 0, s.row0, s.jblk[a]);
        s.row0 += kBKh;
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) {
        s.q = 0; s.b = -1; s.cs = 1.f; s.sn = 0.f; s.dx0 = 0.f; s.dx1 = 0.f; s.pq = -1; s.pb = 0; s.pv = make_float2(0.f, 0.f);
    }
    __device__ static void gen_tile_begin(const Params&, const PairTile& ti, GenState& s, uint8_t*, int) { s.q = ti.kc_begin; s.pq = -1; }
    __device__ static void gen_prepare(const Params&, const PairTile&, GenState&, uint8_t*, int) {}
    __device__ static void gen_advance(const Params&, const PairTile&, GenState& s) { ++s.q; }
    // one group (128 threads) fills a stage: thread = (pixel row of the chunk, 64-feature block): 64 cosines, 8 swizzled
    // 16-byte stores.  The 8 lanes of a store phase write the same 16-byte column of 8 consecutive rows = 8 distinct
    // slots of the 128 B swizzle.
    // raw coordinate of pixel row m and its image (32-bit division: M < 2^31 is checked by the host)
    __device__ static void load_coord(const Params& p, long long m, float2& v, int& b) {
        if (p.cx.theta == nullptr) {
            v = __ldg(reinterpret_cast<const float2*>(p.cx.x) + m);
            b = FEAT == 1 ? static_cast<int>(static_cast<unsigned>(m) / static_cast<unsigned>(p.cx.N)) : 0;
        } else {
            const unsigned mm = static_cast<unsigned>(m), N = static_cast<unsigned>(p.cx.N);
            b = static_cast<int>(mm / N);
            v = __ldg(reinterpret_cast<const float2*>(p.cx.x) + (mm - static_cast<unsigned>(b) * N));
        }
    }
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, uint8_t* a_stage, uint8_t* extra, int gtid) {
        const float4* tab = reinterpret_cast<const float4*>(extra);
        const int prow = gtid & 63, blk = gtid >> 6;
        const long long m = (long long)s.q * kBKh + prow;
        const bool ok = ti.m_tile >= 0 && m < p.cx.M;
        float x0 = 0.f, x1 = 0.f;
        // The pixel's coordinate was loaded while the group's PREVIOUS chunk was generated (a group owns every other chunk:
        // + 2); the load for the next one is issued now and lands during this chunk's 64 cosines.  Without it every chunk began
        // with an exposed global-load round trip on the warps that pace this kernel.
        float2 v = s.pv;
        int b = s.pb;
        if (s.pq != s.q) { if (ok) load_coord(p, m, v, b); }
        {
            const long long mn = m + 2 * kBKh;
            s.pq = s.q + 2;
            if (ti.m_tile >= 0 && s.q + 2 < ti.kc_end && mn < p.cx.M) load_coord(p, mn, s.pv, s.pb);
            else s.pq = -1;
        }
        if (ok) {
            if (p.cx.theta == nullptr) {
                x0 = v.x; x1 = v.y;
            } else {
                if (b != s.b) {                                   // rotation / shift of the image, cached across chunks
                    const float2 d = __ldg(reinterpret_cast<const float2*>(p.cx.dx) + b);
                    sincosf(__ldg(p.cx.theta + b), &s.sn, &s.cs);
                    s.dx0 = d.x; s.dx1 = d.y; s.b = b;
                }
                const float t0 = v.x - s.dx0, t1 = v.y - s.dx1;
                x0 = t0 * s.cs - t1 * s.sn;                        // x' = (x - dx) [[cos, sin], [-sin, cos]]  (train_mnist.py:234-239)
                x1 = t0 * s.sn + t1 * s.cs;
            }
        }
        const int f0 = ti.a0 + blk * 64;
        uint8_t* dst = a_stage + blk * (kBKh * 128);
        if (ok && f0 + 64 <= p.E) {                                // whole block of live features: no per-feature predicates
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                float e[8];
                if constexpr (FEAT == 1) {
                    const float* zrow = p.zbc + (long long)b * p.E + f0;
                    // same arithmetic as the forward generator (GenL1FwdPairT<., 1>): bias = b1 + zb first, then the two FMAs
                    const float4 z0 = __ldg(reinterpret_cast<const float4*>(zrow + ch * 8)), z1 = __ldg(reinterpret_cast<const float4*>(zrow + ch * 8 + 4));
                    const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 t = tab[f0 + ch * 8 + q];
                        t.z += zz[q];
                        e[q] = lrelu(fourier_phase(t, x0, x1));
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) e[q] = __cosf(fourier_phase(tab[f0 + ch * 8 + q], x0, x1));
                }
                *reinterpret_cast<uint4*>(dst + sw128_offset(prow, ch)) =
                    make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
            }
            return;
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            float e[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = f0 + ch * 8 + q;
                e[q] = 0.f;
                if (ok && f < p.E) {
                    if constexpr (FEAT == 1) {
                        float4 t = tab[f];
                        t.z += __ldg(p.zbc + (long long)b * p.E + f);
                        e[q] = lrelu(fourier_phase(t, x0, x1));
                    } else {
                        e[q] = __cosf(fourier_phase(tab[f], x0, x1));
                    }
                }
            }
            *reinterpret_cast<uint4*>(dst + sw128_offset(prow, ch)) =
                make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
        }
    }
    using EpiState = StagedEpiState;
    static constexpr int kStoreBufs = 2;
    __device__ static void epi_init(const Params&, EpiState& st, uint8_t*, int) { st.blocks = 0; st.sel = 0; }
    __device__ static int store_off(const Params& p) { return p.stage_off; }
    __device__ static int store_blocks(const Params& p, const PairTile& ti, int n0, bool has_work) {
        if (!p.tma_reduce || !has_work || ti.m_tile < 0) return 0;
        return min(kAccN / 32, (p.H - n0 + 31) / 32);
    }
    __device__ static void store_issue(const Params& p, const PairTile& ti, int n0, int blk, uint32_t src) {
        tma_reduce_add_2d(&p.tmD, src, ti.a0, n0 + blk * 32);
    }
    __device__ static void epilogue(const Params& p, const PairTile& ti, EpiState& st, int n0, uint32_t taddr, int row, bool has_work,
                                    uint8_t* extra) {
        if (p.tma_reduce) {
            staged_reduce_epilogue<2>(taddr, store_blocks(p, ti, n0, has_work), st, extra + p.stage_off, row, __ldg(p.acc_scale));
            return;
        }
        const int f = ti.a0 + row;
        const bool ok = has_work && ti.m_tile >= 0 && f < p.E;
        const float acc_scale = __ldg(p.acc_scale);
#pragma unroll 1
        for (int c = 0; c < kAccN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            if (!ok) continue;
            const int j0 = n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j0 + j < p.H) atomicAdd(p.dW1 + (long long)(j0 + j) * p.E + f, __uint_as_float(rr[j]) * acc_scale);
        }
    }
};

using GenL1WgradPair = GenL1WgradPairT<0>;

// ------------------------------------------------------------------------------------------------
// Generator layer 1 forward on the CTA-pair kernel:
//
//   h1[m][j] = LeakyReLU(sum_f cos(phase(m, f)) W1[j][f] + b1[j] + zb[image(m)][j])      (models.py:53-58, 95-117)
//
// accumulator rows = pixels m (256 per pair, 128 per CTA), columns = ALL hidden units j (two N = 256 accumulators), so
// a generated [128 px x 64 f] feature chunk (8192 MUFU.COS = 512 clocks) feeds 1024 clocks of tensor-core work instead
// of 512 (GenL1Fwd<256> evaluates every feature twice, once per 256-column half of H = 512, and is generator-bound).
// A = features generated K-major in the 128 B swizzle, B = W1 fp16 [H][E] by TMA (each CTA stages half of every B tile),
// epilogue = bias + latent bias + LeakyReLU, fp16 tiles through a swizzled staging buffer and TMA stores.
struct GenL1FwdPairParams {
    CUtensorMap tmB;          // W1 fp16 [H][E], boxes {64 f, 128 rows}
    CUtensorMap tmC;          // h1 fp16 [M][H] store view, boxes {64, 128 rows}
    int num_stages, num_tiles, m_tiles, k_chunks;
    int bias_off, tab_off, stage_off;  // byte offsets in the extra smem: [H] b1 floats, [2][H] per-tile additive rows, 3 staging buffers of 16 KB
    CoordXform cx;
    const float* wf_scaled; const float* bf;
    int E, H;
    const float* bias;        // (H)
    const float* zb;          // (B,H) latent_linear(z) or null
    int act;                  // kActTanh or LeakyReLU
    // FEAT = 1: the coordinate layer's activation is the generated operand (wf_scaled / bf = its weight (E,2) / bias (E))
    const float* zbc;         // latent bias (B,E) of the coordinate layer
    unsigned long long* mask_bits;   // out [E/64][M]: bit q of word (kc, m) = (pre-activation of feature 64 kc + q > 0), or null
                                     // FEAT = 0 (LeakyReLU): out [H/64][M], the same one-bit mask of the STORED activation h1, or null
    const float* proj_w; const float* proj_bias; float* proj_out;   // fused output projection proj_out[m][o] = sum_j act[m][j] proj_w[o][j] + proj_bias[o]
    int n_proj, proj_off;     // n_proj <= 4 (0: none); proj_off: byte offset of the [n_proj][H] weight rows in the extra smem
};

// FEAT = 1 - generators WITHOUT Fourier features (cfg2: dSprites): the coordinate layer
//   a0[m][f] = LeakyReLU(W1[f] . x'_m + b1[f] + zb[image(m)][f])          (models.py:95-117, train_dsprites.py:222-239)
// is two FMAs per element, so it is generated straight into the A operand of the FIRST HIDDEN layer's GEMM exactly like the
// Fourier features (the table {w0, w1, b1 + zb} is rebuilt per tile for the at most two images it touches), and the kernel is
//   acts1 = LeakyReLU(a0 Wh^T + bh),   y_hat = acts1 wout^T + bout  (fused projection when this is the last hidden layer)
// - coordinate transform, first layer, hidden layer and output projection in ONE kernel; a0 (0.42 GB at cfg2) is never written
// or read.  The backward pass regenerates a0 for the hidden weight gradient (GenL1WgradPairT<1>) and takes the LeakyReLU mask of
// the input gradient from mask_bits (one bit per element, written here by the generator warps).
// Fused output projection of the staged epilogue (fp32 activations): proj[o] += sum_j v[j] * w[o][col + j].  The first output's
// weights are fetched one column group ahead (NoPost's contract); further outputs (n_proj <= 4) load theirs in place.
struct ProjPost {
    const float* s_proj;      // [n_proj][H] in shared memory, already offset to the accumulator's first column
    int n_proj, H;
    float* proj;              // the row's four accumulators (registers of the caller)
    struct W { float4 a, b; };
    __device__ W load(int col) const {
        W w;
        w.a = *reinterpret_cast<const float4*>(s_proj + col);
        w.b = *reinterpret_cast<const float4*>(s_proj + col + 4);
        return w;
    }
    __device__ static float dot8(const float4& a, const float4& b, const float (&v)[8]) {
        // four independent chains (one warp per scheduler: dependent FMAs cost their latency)
        const float t0 = fmaf(v[1], a.y, v[0] * a.x), t1 = fmaf(v[3], a.w, v[2] * a.z);
        const float t2 = fmaf(v[5], b.y, v[4] * b.x), t3 = fmaf(v[7], b.w, v[6] * b.z);
        return (t0 + t1) + (t2 + t3);
    }
    __device__ void packed(int, int, const uint4&) const {}
    __device__ void apply(const W& w, int col, const float (&v)[8]) const {
        if (n_proj > 0) proj[0] += dot8(w.a, w.b, v);
#pragma unroll
        for (int o = 1; o < 4; ++o)               // static indices: proj[] stays in registers
            if (o < n_proj) proj[o] += dot8(*reinterpret_cast<const float4*>(s_proj + o * H + col), *reinterpret_cast<const float4*>(s_proj + o * H + col + 4), v);
    }
};

template <bool TANH, int FEAT = 0>
struct GenL1FwdPairT : PolicyBase {
    static constexpr const char* kName = "gen_l1_fwd";
    static constexpr int kProbeSlot = 3;
    using Params = GenL1FwdPairParams;
    static constexpr bool kF16 = true;
    struct TmaState { int kc, n_row0; };
    struct GenState { float x0, x1; int kc; int sel; long long m; int b_lo, b_hi; unsigned long long pend_bits; long long pend_idx; };
    using EpiState = StagedEpiState;
    static constexpr int kStoreBufs = 3;      // staging buffers of the h1 stores (tc_gemm2's store issuer)
    // contiguous tile ranges per pair: consecutive tiles lie in the same image (N / 128 tiles per image), so the per-tile
    // tables (generator: b1 + zb of the coordinate layer; epilogue: bias + latent bias) are rebuilt once per image only
    static constexpr bool kContiguousTiles = true;
    __device__ static void prefetch_descs(const Params& p) {
        tma_prefetch_desc(&p.tmB);
        tma_prefetch_desc(&p.tmC);
    }
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        // FEAT = 1: the static part {w0, w1, b1} of the coordinate layer's table sits behind the two per-image tables
        load_fourier_table(reinterpret_cast<float4*>(extra) + (FEAT == 1 ? 2 * p.E : 0), p.wf_scaled, p.bf, p.E, tid, nthreads);
        float* s_bias = reinterpret_cast<float*>(extra + p.bias_off);
        for (int j = tid; j < p.H; j += nthreads) s_bias[j] = __ldg(p.bias + j);
        if (FEAT == 1) {
            float* s_proj = reinterpret_cast<float*>(extra + p.proj_off);
            for (int j = tid; j < p.n_proj * p.H; j += nthreads) s_proj[j] = __ldg(p.proj_w + j);
        }
    }
    __device__ static void epi_init(const Params&, EpiState& st, uint8_t*, int) {
        st.blocks = 0; st.sel = 0; st.b_lo = -1; st.b_hi = -1;
        st.proj[0] = st.proj[1] = st.proj[2] = st.proj[3] = 0.f;
    }
    __device__ static int store_off(const Params& p) { return p.stage_off; }
    __device__ static int store_blocks(const Params& p, const PairTile& ti, int n0, bool has_work) {
        return (has_work && ti.m_tile >= 0) ? min(kAccN / 64, (p.H - n0) / 64) : 0;
    }
    __device__ static void store_issue(const Params& p, const PairTile& ti, int n0, int blk, uint32_t src) {
        tma_store_2d(&p.tmC, src, n0 + blk * 64, ti.a0);
    }
    __device__ static void tile_info(const Params& p, int tile, uint32_t rank, PairTile& ti) {
        ti.n0 = 0;
        ti.n_acc = p.H > kAccN ? 2 : 1;
        const int mt = 2 * tile + static_cast<int>(rank);
        ti.m_tile = mt < p.m_tiles ? mt : -1;
        ti.a0 = mt * kBM;                                  // first pixel row of this CTA's tile
        ti.kc_begin = 0;
        ti.kc_end = p.k_chunks;
        ti.a1 = ti.a2 = ti.a3 = 0;
    }
    __device__ static void tma_tile_begin(const Params&, const PairTile& ti, uint32_t rank, TmaState& s) {
        s.kc = 0;
        s.n_row0 = ti.n0 + static_cast<int>(rank) * 128;
    }
    __device__ static void tma_chunk(const Params& p, const PairTile& ti, TmaState& s, uint32_t sb, uint32_t bar) {
        for (int a = 0; a < ti.n_acc; ++a) tma_load_2d_pair(sb + a * kBHalfBytes, &p.tmB, bar, s.kc * kBKh, s.n_row0 + a * kAccN);
        ++s.kc;
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.x0 = 0.f; s.x1 = 0.f; s.kc = 0; s.sel = 0; s.m = 0; s.b_lo = -1; s.b_hi = -1; s.pend_idx = -1; s.pend_bits = 0; }
    __device__ static void gen_tile_begin(const Params& p, const PairTile& ti, GenState& s, uint8_t* extra, int ptid) {
        s.kc = 0;
        s.sel = 0;
        s.m = (long long)ti.a0 + (ptid & (kBM - 1));
        if (ti.m_tile < 0) { s.x0 = 0.f; s.x1 = 0.f; return; }
        if (FEAT == 1) {
            // feature table of the tile: {w0, w1, b1 + zb[image]} for the (at most two: N >= 128) images of its 128 rows;
            // all 8 generator warps, uniform over the CTA
            float4* tab = reinterpret_cast<float4*>(extra);
            const long long m_last = min((long long)ti.a0 + kBM - 1, p.cx.M - 1);
            const int b_lo = static_cast<int>(static_cast<unsigned>(ti.a0) / static_cast<unsigned>(p.cx.N));
            const int b_hi = static_cast<int>(static_cast<unsigned>(m_last) / static_cast<unsigned>(p.cx.N));
            if (b_lo != s.b_lo || b_hi != s.b_hi) {               // uniform over the CTA
                named_bar_sync(1, kGenWarps * 32);                // the previous tile's table reads are done
                const float4* stat = tab + 2 * p.E;               // {w0, w1, b1, 0} per feature (setup)
                for (int j = ptid; j < 2 * p.E; j += kGenWarps * 32) {
                    const int which = j >= p.E, f = j - which * p.E;
                    float4 t = stat[f];
                    t.z += __ldg(p.zbc + (long long)(which ? b_hi : b_lo) * p.E + f);
                    tab[j] = t;
                }
                named_bar_sync(1, kGenWarps * 32);
                s.b_lo = b_lo; s.b_hi = b_hi;
            }
            s.sel = (b_hi != b_lo && min(s.m, p.cx.M - 1) >= (long long)b_hi * p.cx.N) ? 1 : 0;
        }
        transformed_coord(p.cx, (long long)ti.a0 + (ptid & (kBM - 1)), s.x0, s.x1);
    }
    __device__ static void gen_prepare(const Params&, const PairTile&, GenState&, uint8_t*, int) {}
    __device__ static void gen_advance(const Params&, const PairTile&, GenState& s) { ++s.kc; }
    __device__ static void gen_finish(const Params& p, GenState& s) {
        if (FEAT == 1 && s.pend_idx >= 0) { p.mask_bits[s.pend_idx] = s.pend_bits; s.pend_idx = -1; }
    }
    // one group (128 threads) fills a stage: thread = one pixel row, all 64 features of the chunk (8 swizzled 16-byte stores)
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, uint8_t* a_stage, uint8_t* extra, int gtid) {
        const float4* tab = reinterpret_cast<const float4*>(extra);
        const int row = gtid;
        const bool live = ti.m_tile >= 0;
        const int f0 = s.kc * kBKh;
        if (FEAT == 1) {
            // the coordinate layer's activation (E % 64 == 0: host); sign bits of the pre-activation for the backward mask.
            // The mask word of the PREVIOUS chunk is stored now: the proxy fence that publishes a stage waits for the thread's
            // outstanding global stores too, so a store right before it put a global round trip into every chunk (5.3 k clocks
            // per chunk against 2 k for the cosine features).
            if (s.pend_idx >= 0) { p.mask_bits[s.pend_idx] = s.pend_bits; s.pend_idx = -1; }
            const float4* t2 = tab + s.sel * p.E + f0;
            uint32_t bits[2] = {0u, 0u};
            if (!live) {                                           // padding half of the last pair: zero operand rows
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(a_stage + sw128_offset(row, ch)) = make_uint4(0u, 0u, 0u, 0u);
                return;
            }
            // branch-free: 64 independent {LDS.128, 2 FMA, compare, max} chains (a per-element branch serialised them: 5.3 k
            // clocks per chunk against 2 k for the cosine features)
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                float e[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float pre = fourier_phase(t2[ch * 8 + q], s.x0, s.x1);
                    bits[ch >> 2] |= static_cast<uint32_t>(pre > 0.f) << ((ch & 3) * 8 + q);
                    e[q] = TANH ? tanhf(pre) : lrelu(pre);
                }
                *reinterpret_cast<uint4*>(a_stage + sw128_offset(row, ch)) =
                    make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
            }
            if (p.mask_bits && s.m < p.cx.M) {
                s.pend_bits = bits[0] | (static_cast<unsigned long long>(bits[1]) << 32);
                s.pend_idx = (long long)s.kc * p.cx.M + s.m;
            }
            return;
        }
        if (live && f0 + kBKh <= p.E) {
            // whole chunk of live features (E % 64 == 0 in every reference configuration): no per-feature predicates - they
            // were 4 of the 10 instructions per feature on the warps that pace this kernel
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                float e[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) e[q] = __cosf(fourier_phase(tab[f0 + ch * 8 + q], s.x0, s.x1));
                *reinterpret_cast<uint4*>(a_stage + sw128_offset(row, ch)) =
                    make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
            }
            return;
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            float e[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = f0 + ch * 8 + q;
                e[q] = (live && f < p.E) ? __cosf(fourier_phase(tab[f], s.x0, s.x1)) : 0.f;
            }
            *reinterpret_cast<uint4*>(a_stage + sw128_offset(row, ch)) =
                make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
        }
    }
    // Per tile, while its MMAs run: the additive row of the epilogue, b1 + latent bias of the image, once per image the
    // tile's 128 rows can belong to (N >= 128 pixels per image: at most two).  The epilogue then reads shared memory only -
    // the per-row global loads of the latent bias were the larger half of a latency-bound epilogue.
    __device__ static void epi_tile_begin(const Params& p, const PairTile& ti, EpiState& st, uint8_t* extra, int row) {
        st.sel = 0;
        if (FEAT == 1) st.proj[0] = st.proj[1] = st.proj[2] = st.proj[3] = 0.f;
        if (ti.m_tile < 0) return;                              // uniform over the CTA
        float* tab = reinterpret_cast<float*>(extra + p.tab_off);
        const float* s_bias = reinterpret_cast<const float*>(extra + p.bias_off);
        const long long m_first = ti.a0, m_last = min((long long)ti.a0 + kBM - 1, p.cx.M - 1);
        const int b_lo = static_cast<int>(m_first / p.cx.N), b_hi = static_cast<int>(m_last / p.cx.N);
        const long long m = min((long long)ti.a0 + row, p.cx.M - 1);
        st.sel = (m >= (long long)b_hi * p.cx.N && b_hi != b_lo) ? 1 : 0;
        if (b_lo == st.b_lo && b_hi == st.b_hi) return;         // same images as the previous tile (uniform): table is current
        st.b_lo = b_lo; st.b_hi = b_hi;
        named_bar_sync(2, kEpiWarps * 32);                      // every warp's table reads of the previous tile are done
        for (int j = row; j < 2 * p.H; j += kEpiWarps * 32) {
            const int which = j >= p.H, jj = j - which * p.H;
            const int b = which ? b_hi : b_lo;
            tab[j] = s_bias[jj] + (p.zb ? __ldg(p.zb + (long long)b * p.H + jj) : 0.f);
        }
        named_bar_sync(2, kEpiWarps * 32);
    }
    // 64-column blocks: + (b1 + latent bias) + activation -> fp16 -> staging buffer (ring of three) -> TMA store by the store
    // issuer warp (staged_store_epilogue); rows past M are clipped by the tensor map.
    __device__ static void epilogue(const Params& p, const PairTile& ti, EpiState& st, int n0, uint32_t taddr, int row, bool has_work,
                                    uint8_t* extra) {
        const float* tab = reinterpret_cast<const float*>(extra + p.tab_off) + st.sel * p.H;
        if constexpr (FEAT == 1) {
            ProjPost post{reinterpret_cast<const float*>(extra + p.proj_off) + n0, p.n_proj, p.H, st.proj};
            staged_store_epilogue<TANH, 3, 3>(taddr, store_blocks(p, ti, n0, has_work), st, extra + p.stage_off, row, true,
                                              [&](int blk) { return tab + n0 + blk * 64; }, post);
        } else if constexpr (!TANH) {
            // + the one-bit derivative mask of h1 for the first hidden layer's input gradient (0.1 GB instead of a second
            // 1.7 GB read of h1 at cfg4)
            const long long m = (long long)ti.a0 + row;
            uint32_t lo = 0u, acc = 0u;
            MaskBitsPost post{(p.mask_bits && has_work && ti.m_tile >= 0 && m < p.cx.M) ? p.mask_bits + (long long)(n0 >> 6) * p.cx.M + m : nullptr,
                              p.cx.M, &lo, &acc};
            staged_store_epilogue<TANH, 3, 3>(taddr, store_blocks(p, ti, n0, has_work), st, extra + p.stage_off, row, true,
                                              [&](int blk) { return tab + n0 + blk * 64; }, post);
        } else {
            staged_store_epilogue<TANH, 3, 3>(taddr, store_blocks(p, ti, n0, has_work), st, extra + p.stage_off, row, true,
                                              [&](int blk) { return tab + n0 + blk * 64; });
        }
    }
    // after both accumulators of the tile: the row's projected outputs
    __device__ static void epi_tile_end(const Params& p, const PairTile& ti, EpiState& st, uint8_t*, int row, bool has_work) {
        if (FEAT == 0 || p.n_proj == 0 || !has_work || ti.m_tile < 0) return;
        const long long m = (long long)ti.a0 + row;
        if (m >= p.cx.M) return;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            if (o < p.n_proj) {
                p.proj_out[m * p.n_proj + o] = st.proj[o] + (p.proj_bias ? __ldg(p.proj_bias + o) : 0.f);
            }
        }
    }
};
using GenL1FwdPair = GenL1FwdPairT<false>;

}  // namespace tvae
