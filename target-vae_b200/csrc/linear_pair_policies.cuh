// Weight gradient of a generator hidden layer on the CTA-pair kernel (tc_gemm2_kernel):
//
//   dW[j][i] += sum_m dpre[m][j] * a_prev[m][i]          (models.py:89 backward;  j, i < H <= 512)
//
// as  C^T[i][j] = sum_m P[m][i] Q[m][j]  with P = a_prev, Q = dpre: accumulator rows = input units i (256 per pair, 128 per
// CTA), columns = ALL output units j (two N = 256 accumulators fill TMEM), reduction over the B*n^2 pixel rows in chunks
// of 64, split across the pairs; fp32 atomics into dW[j][i] - consecutive lanes own consecutive i, so they coalesce.
//
// Why the pair kernel: LinearTN<256> (tc_gemm, 128 x 256 output tiles) re-streams the dpre column slice once per column
// tile and the activation slice once per row tile - 2 x |dpre| + 4 x |a_prev| = 10 GB of L2 -> SM traffic per layer at the
// particle-stack size, which is what bounds it (9.6 TB/s of L2 bandwidth, tensor pipe 36-48 %).  With 256 x 512 pair tiles
// every SM stages its own 128 dpre columns plus HALF of the activation columns: |dpre| + 2 x |a_prev| = 5 GB.
//
// A = P^T MN-major in the 128 B swizzle: two [64 rows][64 columns] blocks per stage, copied from global memory by the
// operand-generator warps (one 128-byte row segment per thread: 8 x LDG.128 + 8 swizzled STS.128 - the layout a TMA box
// {64 columns, 64 rows} / SWIZZLE_128B would produce); B = Q by 3-D TMA, MN-major, 64 B swizzle.
#pragma once
#include "gen_pair_policies.cuh"

namespace tvae {

struct LinearTNPairParams {
    CUtensorMap tmQ;          // Q fp16 [M][Nb] as {32 columns, M rows, Nb/32 blocks}, boxes {32, 64, 4}
    int num_stages, num_tiles, m_pairs, m_tiles, splits, chunks_total, chunks_per_split;
    const __half* P;          // fp16 [M][ldp]
    long long ldp;
    long long M;              // reduction rows
    int Ma, Nb;               // accumulator rows (columns of P) / columns (columns of Q)
    float* C;                 // stored TRANSPOSED: C[nb][ma] at C + nb * ldc + ma; zero-filled by the caller
    long long ldc;
    const float* acc_scale;   // device scalar multiplied into the accumulator (undoes the operands' power-of-two scale), or null
    CUtensorMap tmP;          // a_tma == 1: P as {64 cols, rows, Ma / 64} (make_tmap_3d_mn128_h), boxes {64, 64 rows, 2}: one TMA fills the
    int a_tma;                // CTA's whole A stage (two [64 rows][128 B] blocks, 128 B swizzle) - needs Ma % 64 == 0
};

struct LinearTNPair : PolicyBase {
    static constexpr const char* kName = "linear_tn";
    static constexpr int kProbeSlot = 5;
    using Params = LinearTNPairParams;
    static constexpr bool kF16 = true;
    static constexpr bool kAMajorMN = true;
    static constexpr bool kBMajorMN = true;
    struct TmaState { int row0; int jblk[kAcc]; };
    struct GenState { int q; };
    // bytes the A-operand TMA loads of BOTH CTAs add to a stage's transaction count (tc_gemm2: the leader's expect_tx)
    __device__ static uint32_t a_tx_bytes(const Params& p) { return p.a_tma ? 2u * kAStageBytes : 0u; }
    __device__ static void prefetch_descs(const Params& p) {
        tma_prefetch_desc(&p.tmQ);
        if (p.a_tma) tma_prefetch_desc(&p.tmP);
    }
    __device__ static void tile_info(const Params& p, int tile, uint32_t rank, PairTile& ti) {
        const int sp = tile / p.m_pairs;
        const int mp = tile - sp * p.m_pairs;
        ti.n0 = 0;
        ti.n_acc = p.Nb > kAccN ? 2 : 1;
        const int mt = 2 * mp + static_cast<int>(rank);
        ti.m_tile = mt < p.m_tiles ? mt : -1;
        ti.a0 = mt * kBM;                                  // first accumulator row (column of P) of this CTA
        ti.kc_begin = min(sp * p.chunks_per_split, p.chunks_total);
        ti.kc_end = min(ti.kc_begin + p.chunks_per_split, p.chunks_total);
        ti.a1 = ti.a2 = ti.a3 = 0;
    }
    __device__ static void tma_tile_begin(const Params&, const PairTile& ti, uint32_t rank, TmaState& s) {
        s.row0 = ti.kc_begin * kBKh;
#pragma unroll
        for (int a = 0; a < kAcc; ++a) s.jblk[a] = (a * kAccN + static_cast<int>(rank) * 128) >> 5;
    }
    // this CTA's half of each accumulator's B tile: 128 activation columns x 64 rows (columns >= Nb, rows >= M: TMA zero fill)
    __device__ static void tma_chunk(const Params& p, const PairTile& ti, TmaState& s, uint32_t sb, uint32_t bar) {
#pragma unroll
        for (int a = 0; a < kAcc; ++a)
            if (a < ti.n_acc) tma_load_3d_pair(sb + a * kBHalfBytes, &p.tmQ, bar, 0, s.row0, s.jblk[a]);
        // A operand (this CTA's 128 columns of P, 64 reduction rows) straight into the stage: the generator warps' copy through
        // registers bound this kernel (84 % busy, the MMA issuer waiting for operands half of the time: clock64 probe)
        if (p.a_tma) tma_load_3d_pair(sb - kAStageBytes, &p.tmP, bar, 0, s.row0, ti.a0 >> 6);
        s.row0 += kBKh;
    }
    __device__ static void gen_init(const Params&, GenState& s, uint8_t*, int) { s.q = 0; }
    __device__ static void gen_tile_begin(const Params&, const PairTile& ti, GenState& s, uint8_t*, int) { s.q = ti.kc_begin; }
    __device__ static void gen_prepare(const Params&, const PairTile&, GenState&, uint8_t*, int) {}
    __device__ static void gen_advance(const Params&, const PairTile&, GenState& s) { ++s.q; }
    // one group (128 threads) fills a stage: thread = (reduction row of the chunk, 64-column block): one 128-byte segment
    // of a dpre row.  The 8 lanes of a store phase write the same 16-byte column of 8 consecutive rows = 8 distinct slots
    // of the 128 B swizzle.
    __device__ static void gen_chunk(const Params& p, const PairTile& ti, GenState& s, uint8_t* a_stage, uint8_t*, int gtid) {
        if (p.a_tma) return;                   // the TMA producer loads A; the generator warps only keep the barrier protocol
        const int prow = gtid & 63, blk = gtid >> 6;
        const long long m = (long long)s.q * kBKh + prow;
        const int j0 = ti.a0 + blk * 64;
        const bool row_ok = ti.m_tile >= 0 && m < p.M;
        uint8_t* dst = a_stage + blk * (kBKh * 128);
        const __half* src = p.P + m * p.ldp + j0;
        uint4 v[8];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            v[ch] = make_uint4(0u, 0u, 0u, 0u);
            if (row_ok && j0 + ch * 8 + 8 <= p.Ma) v[ch] = __ldg(reinterpret_cast<const uint4*>(src) + ch);
            else if (row_ok && j0 + ch * 8 < p.Ma) {                       // ragged last 8-column group
                unsigned short h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = (j0 + ch * 8 + e < p.Ma) ? __half_as_ushort(src[ch * 8 + e]) : (unsigned short)0;
                v[ch] = make_uint4(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16), h[4] | (uint32_t(h[5]) << 16), h[6] | (uint32_t(h[7]) << 16));
            }
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(dst + sw128_offset(prow, ch)) = v[ch];
    }
    // fp32 atomics: the epilogue is 3-10 % of this kernel (the generator warps bound it); TMA reduce-adds (measured) made it
    // 4-12 % slower - the reduces share the TMA unit with the operand loads the kernel is starved for
    __device__ static void epilogue(const Params& p, const PairTile& ti, EpiState&, int n0, uint32_t taddr, int row, bool has_work, uint8_t*) {
        const int j = ti.a0 + row;
        const bool ok = has_work && ti.m_tile >= 0 && j < p.Ma;
        const float acc_scale = p.acc_scale ? __ldg(p.acc_scale) : 1.f;
#pragma unroll 1
        for (int c = 0; c < kAccN / 32; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            tmem_ld_wait();
            if (!ok) continue;
            const int i0 = n0 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i0 + i < p.Nb) atomicAdd(p.C + (long long)(i0 + i) * p.ldc + j, __uint_as_float(rr[i]) * acc_scale);
        }
    }
};

}  // namespace tvae
