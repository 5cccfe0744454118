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
};

struct GenL1WgradPair : PolicyBase {
    static constexpr const char* kName = "gen_l1_wgrad";
    using Params = GenL1WgradPairParams;
    static constexpr bool kF16 = true;
    static constexpr bool kAMajorMN = true;
    static constexpr bool kBMajorMN = true;
    struct TmaState { int row0; int jblk[kAcc]; };
    struct GenState {
        int q;                // current chunk
        int b;                // image whose rotation / shift are cached (-1: none)
        float cs, sn, dx0, dx1;
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
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.q = 0; s.b = -1; s.cs = 1.f; s.sn = 0.f; s.dx0 = 0.f; s.dx1 = 0.f; }
    __device__ static void gen_tile_begin(const Params&, const PairTile& ti, GenState& s, uint8_t*, int) { s.q = ti.kc_begin; }
    __device__ static void gen_prepare(const Params&, const PairTile&, GenState&, uint8_t*, int) {}
    __device__ static void gen_advance(const Params&, const PairTile&, GenState& s) { ++s.q; }
    // one group (128 threads) fills a stage: thread = (pixel row of the chunk, 64-feature block): 64 cosines, 8 swizzled
    // 16-byte stores.  The 8 lanes of a store phase write the same 16-byte column of 8 consecutive rows = 8 distinct
    // slots of the 128 B swizzle.
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, uint8_t* a_stage, uint8_t* extra, int gtid) {
        const float4* tab = reinterpret_cast<const float4*>(extra);
        const int prow = gtid & 63, blk = gtid >> 6;
        const long long m = (long long)s.q * kBKh + prow;
        const bool ok = ti.m_tile >= 0 && m < p.cx.M;
        float x0 = 0.f, x1 = 0.f;
        if (ok) {
            if (p.cx.theta == nullptr) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(p.cx.x) + m);
                x0 = v.x; x1 = v.y;
            } else {
                const int b = static_cast<int>(m / p.cx.N);
                const int px = static_cast<int>(m - (long long)b * p.cx.N);
                if (b != s.b) {                                   // rotation / shift of the image, cached across chunks
                    const float2 d = __ldg(reinterpret_cast<const float2*>(p.cx.dx) + b);
                    sincosf(__ldg(p.cx.theta + b), &s.sn, &s.cs);
                    s.dx0 = d.x; s.dx1 = d.y; s.b = b;
                }
                const float2 v = __ldg(reinterpret_cast<const float2*>(p.cx.x) + px);
                const float t0 = v.x - s.dx0, t1 = v.y - s.dx1;
                x0 = t0 * s.cs - t1 * s.sn;                        // x' = (x - dx) [[cos, sin], [-sin, cos]]  (train_mnist.py:234-239)
                x1 = t0 * s.sn + t1 * s.cs;
            }
        }
        const int f0 = ti.a0 + blk * 64;
        uint8_t* dst = a_stage + blk * (kBKh * 128);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            float e[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = f0 + ch * 8 + q;
                e[q] = (ok && f < p.E) ? __cosf(fourier_phase(tab[f], x0, x1)) : 0.f;
            }
            *reinterpret_cast<uint4*>(dst + sw128_offset(prow, ch)) =
                make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
        }
    }
    __device__ static void epilogue(const Params& p, const PairTile& ti, EpiState&, int n0, uint32_t taddr, int row, bool has_work, uint8_t*) {
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
    int bias_off, stage_off;  // byte offsets in the extra smem: [H] b1 floats, 2 staging buffers of 16 KB
    CoordXform cx;
    const float* wf_scaled; const float* bf;
    int E, H;
    const float* bias;        // (H)
    const float* zb;          // (B,H) latent_linear(z) or null
    int act;                  // kActTanh or LeakyReLU
};

template <bool TANH>
struct GenL1FwdPairT : PolicyBase {
    static constexpr const char* kName = "gen_l1_fwd";
    using Params = GenL1FwdPairParams;
    static constexpr bool kF16 = true;
    struct TmaState { int kc, n_row0; };
    struct GenState { float x0, x1; int kc; };
    struct EpiState { int blocks; };
    __device__ static void prefetch_descs(const Params& p) {
        tma_prefetch_desc(&p.tmB);
        tma_prefetch_desc(&p.tmC);
    }
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        load_fourier_table(reinterpret_cast<float4*>(extra), p.wf_scaled, p.bf, p.E, tid, nthreads);
        float* s_bias = reinterpret_cast<float*>(extra + p.bias_off);
        for (int j = tid; j < p.H; j += nthreads) s_bias[j] = __ldg(p.bias + j);
    }
    __device__ static void epi_init(const Params&, EpiState& st, uint8_t*, int) { st.blocks = 0; }
    __device__ static void epi_finish(const Params&, EpiState&, uint8_t*, int row) {
        if (row == 0) tma_store_wait<0>();
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
    __device__ static void tma_tile_begin(const Params&, const PairTile&, uint32_t rank, TmaState& s) {
        s.kc = 0;
        s.n_row0 = static_cast<int>(rank) * 128;
    }
    __device__ static void tma_chunk(const Params& p, const PairTile& ti, TmaState& s, uint32_t sb, uint32_t bar) {
        for (int a = 0; a < ti.n_acc; ++a) tma_load_2d_pair(sb + a * kBHalfBytes, &p.tmB, bar, s.kc * kBKh, s.n_row0 + a * kAccN);
        ++s.kc;
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.x0 = 0.f; s.x1 = 0.f; s.kc = 0; }
    __device__ static void gen_tile_begin(const Params& p, const PairTile& ti, GenState& s, uint8_t*, int ptid) {
        s.kc = 0;
        if (ti.m_tile < 0) { s.x0 = 0.f; s.x1 = 0.f; return; }
        transformed_coord(p.cx, (long long)ti.a0 + (ptid & (kBM - 1)), s.x0, s.x1);
    }
    __device__ static void gen_prepare(const Params&, const PairTile&, GenState&, uint8_t*, int) {}
    __device__ static void gen_advance(const Params&, const PairTile&, GenState& s) { ++s.kc; }
    // one group (128 threads) fills a stage: thread = one pixel row, all 64 features of the chunk (8 swizzled 16-byte stores)
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, uint8_t* a_stage, uint8_t* extra, int gtid) {
        const float4* tab = reinterpret_cast<const float4*>(extra);
        const int row = gtid;
        const bool live = ti.m_tile >= 0;
        const int f0 = s.kc * kBKh;
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
    // 64-column blocks: bias + latent bias + LeakyReLU -> fp16 -> staging buffer (two, alternating) -> TMA store;
    // rows past M are clipped by the tensor map.  Control flow is uniform over the 128 epilogue threads.
    __device__ static void epilogue(const Params& p, const PairTile& ti, EpiState& st, int n0, uint32_t taddr, int row, bool has_work,
                                    uint8_t* extra) {
        const float* s_bias = reinterpret_cast<const float*>(extra + p.bias_off);
        const bool tile_ok = has_work && ti.m_tile >= 0;
        const long long m = (long long)ti.a0 + row;
        const float* zb = (p.zb && tile_ok && m < p.cx.M) ? p.zb + (m / p.cx.N) * p.H : nullptr;
        uint8_t* stage0 = extra + p.stage_off;
#pragma unroll 1
        for (int blk = 0; blk < kAccN / 64; ++blk) {
            const int j0 = n0 + blk * 64;
            const bool blk_ok = tile_ok && j0 < p.H;                // uniform
            uint32_t rr[2][32];
            tmem_ld_32x32(taddr + blk * 64, rr[0]);
            tmem_ld_32x32(taddr + blk * 64 + 32, rr[1]);
            tmem_ld_wait();
            if (!blk_ok) continue;
            uint8_t* buf = stage0 + (st.blocks & 1) * kStoreBlockBytes;
            if (st.blocks >= 2) {                                   // the store that last used this buffer has read it
                if (row == 0) tma_store_wait_read<1>();
                named_bar_sync(2, kEpiWarps * 32);
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float v[8];
#pragma unroll
                    for (int q = 0; q < 8; q += 4) {
                        const int jj = j0 + hf * 32 + j + q;
                        const float4 bb = *reinterpret_cast<const float4*>(s_bias + jj);
                        const float4 bz = zb ? __ldg(reinterpret_cast<const float4*>(zb + jj)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[q] = __uint_as_float(rr[hf][j + q]) + bb.x + bz.x;
                        v[q + 1] = __uint_as_float(rr[hf][j + q + 1]) + bb.y + bz.y;
                        v[q + 2] = __uint_as_float(rr[hf][j + q + 2]) + bb.z + bz.z;
                        v[q + 3] = __uint_as_float(rr[hf][j + q + 3]) + bb.w + bz.w;
                    }
                    act_vec<TANH>(v);
                    *reinterpret_cast<uint4*>(buf + sw128_offset(row, hf * 4 + (j >> 3))) =
                        make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(2, kEpiWarps * 32);
            if (row == 0) {
                tma_store_2d(&p.tmC, smem_u32(buf), j0, ti.a0);
                tma_store_commit();
            }
            ++st.blocks;
        }
    }
};
using GenL1FwdPair = GenL1FwdPairT<false>;

}  // namespace tvae
