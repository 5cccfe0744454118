// CTA-pair (cta_group::2) variant of LinearNT for the generator's hidden layers (models.py:84-93 forward and the
// input-gradient GEMM of its backward):
//
//     C[M, N] = epi(A[M, K] * W[N, K]^T)        256 x 256 pair tiles, fp16 operands, fp32 accumulation in TMEM
//
// Why: LinearNT<256> (tc_gemm, one CTA per 128 x 256 tile) stages 48 KB per 64-deep K chunk (16 KB of activations and the
// whole 32 KB weight chunk) for 512 clocks of tensor-core work, and what bounds it is the L2 -> shared-memory stream:
// 768 KB per 128 activation rows of a 512 x 512 layer, 9.8 GB per launch at the particle-stack size, ~9 TB/s with every
// byte of shared memory already in flight (tensor pipe 33-38 %, profiles/r02_linear_layers.md).  Under cta_group::2 the two
// SMs of a TPC share every MMA: each CTA stages its own 128 activation rows and only HALF of the weight chunk (32 KB per
// chunk), so the same shared memory holds a deeper pipeline and the layer moves 1.5 x fewer bytes per FLOP.
// (A weight-resident 4-CTA cluster with multicast activations was built and measured first: 2 x SLOWER - the resident
// 128 KB weight slice left three 16 KB stages, far too shallow for the ~3 us multicast / cross-CTA barrier round trip.)
//
// Unlike the conv pair kernel (tc_gemm2: one 256 x 512 accumulator = all of TMEM, epilogue serialised with the MMAs)
// a tile here is 256 columns wide: TMEM holds TWO accumulator stages and the two epilogue warpgroups take alternate
// tiles, so a tile's epilogue (the longer half of this short-K GEMM) overlaps the next tile's MMAs.  Both column tiles of
// a 256-row block run back to back on the same pair: the second pass re-reads the activations from L2.
//
// Warp roles in BOTH CTAs (384 threads): warp 0 TMA producer (own A rows + own half of B, complete_tx on the LEADER's
// full barrier), warp 1 MMA issuer (leader only), warp 2 TMEM allocator, warps 4-11 two epilogue groups running
// LinearNT<256>::epilogue on this CTA's 128 accumulator rows.
#pragma once
#include "linear_policies.cuh"
#include "tc_gemm2.cuh"

namespace tvae {

constexpr int kNpStageBytes = kAStageBytes + kBHalfBytes;       // 32 KB: [128 rows][64 k] of A + [128 rows][64 k] of W
constexpr int kNpAccStages = 2;
constexpr int kNpEpiGroups = 2;
constexpr int kNpThreads = (kCtrlWarps + kEpiWarps * kNpEpiGroups) * 32;     // 384

struct LinearNTPairParams {
    LinearNTParams nt;        // tmA: boxes {64 k, 128 rows}; tmB: W [N][K] boxes {64 k, 128 rows}; tmC as in LinearNT
    int num_stages;
    int m_pairs;              // ceil(M / 256)
    int items;                // m_pairs * tiles_n work items, item = m_pair * tiles_n + n_tile
    int bar_off, tmem_ptr_off, extra_off;      // shared-memory layout (bytes); stages start at 0
};

struct NpSmem { int bar_off, tmem_ptr_off, extra_off, total; };
__host__ inline NpSmem np_smem_layout(int stages, int extra_bytes) {
    NpSmem L;
    L.bar_off = stages * kNpStageBytes;
    L.tmem_ptr_off = L.bar_off + (2 * kMaxStages + 2 * kNpAccStages) * 8;
    L.extra_off = (L.tmem_ptr_off + 16 + 1023) & ~1023;
    L.total = L.extra_off + extra_bytes;
    return L;
}

template <bool TANH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNpThreads, 1)
linear_nt_pair_kernel(const __grid_constant__ LinearNTPairParams prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    using Epi = LinearNT<256, TANH>;
    const LinearNTParams& p = prm.nt;
    constexpr uint32_t kIdesc = make_idesc_f16(256, 256, false, false, 0, 0);

    const int stages = prm.num_stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + prm.bar_off);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + prm.tmem_ptr_off);
    uint8_t* extra = smem + prm.extra_off;
    const uint32_t full_bar = smem_u32(bars);                                    // [stages]        (used in the leader)
    const uint32_t empty_bar = smem_u32(bars + kMaxStages);                       // [stages]        (each CTA)
    const uint32_t tfull_bar = smem_u32(bars + 2 * kMaxStages);                   // [kNpAccStages]  (each CTA)
    const uint32_t tempty_bar = smem_u32(bars + 2 * kMaxStages + kNpAccStages);   // [kNpAccStages]  (leader)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        if (p.tma_store) tma_prefetch_desc(&p.tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int s = 0; s < kNpAccStages; ++s) {
            mbar_init(tfull_bar + 8 * s, 1);
            mbar_init(tempty_bar + 8 * s, 2 * kEpiWarps);          // one arrive per epilogue warp of the owning group, both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(tmem_ptr_smem), 512);
        tmem_relinquish_pair();
    }
    Epi::setup(p, extra, threadIdx.x, blockDim.x);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // contiguous item ranges per pair: both column tiles of a row block run back to back on the same pair
    const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const int item_begin = static_cast<int>((long long)prm.items * pair / n_pairs);
    const int item_end = static_cast<int>((long long)prm.items * (pair + 1) / n_pairs);
    const int k_chunks = p.k_chunks, tiles_n = p.tiles_n;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full_leader = mapa_rank(full_bar, 0);
            for (int item = item_begin; item < item_end; ++item) {
                const int mp = item / tiles_n, nt = item - mp * tiles_n;
                const int m0 = (2 * mp + static_cast<int>(rank)) * kBM;        // this CTA's activation rows
                const int n0 = nt * 256 + static_cast<int>(rank) * 128;         // this CTA's half of the weight rows
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    if (leader) mbar_arrive_expect_tx(full_bar + 8 * stage, 2u * kNpStageBytes);
                    const uint32_t sa = smem_u32(smem + stage * kNpStageBytes);
                    tma_load_2d_pair(sa, &p.tmA, full_leader + 8 * stage, kc * kBKh, m0);
                    tma_load_2d_pair(sa + kAStageBytes, &p.tmB, full_leader + 8 * stage, kc * kBKh, n0);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader only)
        if (leader && lane == 0) {
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int item = item_begin; item < item_end; ++item) {
                mbar_wait(tempty_bar + 8 * as, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * kNpStageBytes);
                    const uint32_t b_addr = a_addr + kAStageBytes;
#pragma unroll
                    for (int ks = 0; ks < kKSteps; ++ks) {
                        const uint64_t adesc = make_smem_desc(a_addr + ks * 32, 16, 1024, kLayoutSw128);
                        const uint64_t bdesc = make_smem_desc(b_addr + ks * 32, 16, 1024, kLayoutSw128);
                        umma_f16_pair(d_tmem, adesc, bdesc, kIdesc, (kc > 0 || ks > 0) ? 1u : 0u);
                    }
                    umma_commit_pair(empty_bar + 8 * stage, 3);       // frees the slot in both CTAs
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(tfull_bar + 8 * as, 3);              // accumulator complete -> the owning epilogue group of both CTAs
                if (++as == kNpAccStages) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= kFirstEpiWarp) {
        // ------------------------------------------------------------ epilogue (both CTAs; group g takes items begin + g, + 2, ...)
        const int ewarp = (warp - kFirstEpiWarp) & 3;
        const int egrp = (warp - kFirstEpiWarp) >> 2;
        const int row = ewarp * 32 + lane;
        const uint32_t tempty_leader = mapa_rank(tempty_bar, 0);
        typename Epi::EpiState est;
        Epi::epi_init(p, est, extra, egrp * kBM + row);
        int t = egrp;
        for (int item = item_begin + egrp; item < item_end; item += kNpEpiGroups, t += kNpEpiGroups) {
            const int mp = item / tiles_n, nt = item - mp * tiles_n;
            TileInfo ti;
            ti.m0 = (2 * mp + static_cast<int>(rank)) * kBM; ti.n0 = nt * 256; ti.kc_begin = 0; ti.kc_end = k_chunks;
            const int as = t & 1;
            const uint32_t aphase = (t >> 1) & 1;
            mbar_wait(tfull_bar + 8 * as, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + as * 256;
            Epi::epilogue(p, ti, est, taddr, row, extra);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * as);
        }
        Epi::epi_finish(p, est, extra, egrp * kBM + row);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // no CTA exits (or frees TMEM) while its peer can still signal it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace tvae
