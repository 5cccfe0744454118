// CTA-pair (cta_group::2) persistent FP16 GEMM core for the group convolution and its weight gradient.
//
//   D[256 x 512] (fp32, TMEM of both CTAs) += A[256 x K] * B[512 x K]^T        per pair-tile
//
// Two CTAs of one cluster (the two SMs of a TPC) share every tcgen05.mma: the pair computes 256 accumulator
// rows (128 per CTA) and each CTA stages only HALF of the B tile, which halves both the L2->SM operand
// stream and the shared-memory bandwidth the tensor core needs per FLOP.  The accumulator fills TMEM:
// two N = 256 accumulators per CTA, both fed from the same generated A stage, so the synthesised im2col
// operand is produced once per 512 output columns.
//
// Warp roles in BOTH CTAs (512 threads):
//   warp 0      TMA producer: this CTA's half of the B tile, complete_tx on the LEADER's full barrier
//   warp 1      MMA issuer (leader CTA only): tcgen05.mma.cta_group::2.kind::f16, M = 256, N = 256
//   warp 2      TMEM allocator (cta_group::2 alloc / dealloc, same warp in both CTAs)
//   warp 3      store issuer (policies with kStoreBufs > 0): hands the epilogue's staged fp16 blocks to the TMA store unit.
//               Issuing a bulk-tensor store costs the issuing thread 250-800 clocks (clock64 probe, tools/probe_pair.py);
//               on an epilogue warp that sat on the critical path of every 64-column block (16 % of the epilogue, plus a
//               named barrier per block).  The epilogue warps and this warp meet at mbarriers instead: staged[b] (one
//               arrive per epilogue warp: block written) and freed[b] (the store that used buffer b has read it).
//   warps 4-7   epilogue: tcgen05.ld of this CTA's 128 accumulator rows -> policy epilogue
//   warps 8-15  operand generators: synthesise this CTA's 128 A rows straight into swizzled smem (two groups
//               of 4 warps, alternate stages)
// Barriers: full[s] lives in the leader (1 expect_tx arrive + one arrive per generator warp of the owning group in both CTAs);
// empty[s] and tfull live in each CTA and are signalled by a multicast tcgen05.commit; tempty lives in the
// leader and collects one arrive per epilogue warp of both CTAs.
#pragma once
#include "tc_gemm.cuh"

namespace tvae {

constexpr int kPairThreads = 512;
constexpr int kGenWarps = 8;
constexpr int kAcc = 2;                 // accumulators per CTA (N = 256 each)
constexpr int kAccN = 256;
constexpr int kBHalfBytes = 128 * 128;  // this CTA's half of one accumulator's B stage
constexpr int kStage2Bytes = kAStageBytes + kAcc * kBHalfBytes;   // 48 KB
constexpr int kMaxStoreBufs = 4;        // staging buffers of the epilogue's TMA stores (policy: kStoreBufs)
constexpr int kStoreBlockBytes = kBM * 128;     // one staging buffer: 128 rows x 64 halves

// ---------------------------------------------------------------- cluster / pair PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// Default (.release.cta) semantics on purpose: a .release.cluster arrive compiles to MEMBAR.ALL.GPU + ERRBAR per
// call, which throttled the generator warps to ~45 % tensor-pipe activity (profiles/r01_conv_pair_ncu.md).  The
// operand bytes are published to the async proxy by fence.proxy.async before the arrive, and each CTA's tile is
// read by its own SM's tensor core.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D TMA load issued by either CTA of the pair; bytes are credited to the mbarrier at `bar_cluster`
// (a shared::cluster address, normally the leader's full barrier).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once all previously issued MMAs of this thread retire) on the barrier at the same smem offset in
// every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
}

// Development probe (-DTVAE_PROBE, csrc/build.py with TVAE_PROBE=1 -> libtvae_b200_probe.so; never part of the product
// library): clock64() spent by the leader's MMA issuer waiting for a drained accumulator / for operand stages, by epilogue
// thread 0 inside the policy epilogue, by generator thread 0 waiting for free stages / generating, summed over the pairs.
// counters (8-13: segments of staged_store_epilogue, epilogue thread 0): 0 mma total, 1 mma wait tempty, 2 mma wait full, 3 epi wait tfull, 4 epi epilogue, 5 gen wait empty, 6 gen chunk, 7 #pairs
#ifdef TVAE_PROBE
__device__ unsigned long long g_pair_probe[8][16];      // [P::kProbeSlot][counter]
#define TVAE_PROBE_T0() const long long probe_t0 = clock64()
#define TVAE_PROBE_ADD(var) var += clock64() - probe_t0
#else
#define TVAE_PROBE_T0()
#define TVAE_PROBE_ADD(var)
#endif

struct PairTile {
    int n0;              // first accumulator column of the pair-tile
    int n_acc;           // accumulators in use (1 or 2)
    int kc_begin, kc_end;  // compact reduction steps [kc_begin, kc_end); the policy's Tma/Gen states map steps to K chunks
    int m_tile;          // this CTA's own 128-row tile index (policy space), -1 = none (padding half)
    int a0, a1, a2, a3;  // policy scratch
};

struct Smem2Layout {
    uint32_t stage_off, bar_off, tmem_ptr_off, extra_off, total;
};
__host__ __device__ inline Smem2Layout make_smem2_layout(int stages, int extra_bytes) {
    Smem2Layout L;
    L.stage_off = 0;
    L.bar_off = stages * kStage2Bytes;
    L.tmem_ptr_off = L.bar_off + (2 * kMaxStages + 2 + 2 * kMaxStoreBufs) * 8;
    L.extra_off = (L.tmem_ptr_off + 16 + 1023) & ~1023u;   // policies may keep 128 B-swizzled tiles in their extra region
    L.total = L.extra_off + extra_bytes;
    return L;
}

template <class P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
tc_gemm2_kernel(const __grid_constant__ typename P::Params prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // kind::f16: a stage row of 128 B holds 64 reduction elements, one MMA consumes 16.
    constexpr uint32_t kIdesc = make_idesc_f16(256, kAccN, P::kAMajorMN, P::kBMajorMN, P::kAFmt, P::kBFmt);

    const int stages = prm.num_stages;
    const Smem2Layout L = make_smem2_layout(stages, 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L.tmem_ptr_off);
    uint8_t* extra = smem + L.extra_off;

    const uint32_t full_bar = smem_u32(bars);                       // [stages]   (used in the leader)
    const uint32_t empty_bar = smem_u32(bars + kMaxStages);          // [stages]   (each CTA)
    const uint32_t tfull_bar = smem_u32(bars + 2 * kMaxStages);      // each CTA
    const uint32_t tempty_bar = smem_u32(bars + 2 * kMaxStages + 1); // leader
    const uint32_t staged_bar = smem_u32(bars + 2 * kMaxStages + 2); // [kMaxStoreBufs] each CTA: epilogue warps -> store issuer
    const uint32_t freed_bar = staged_bar + 8 * kMaxStoreBufs;       // [kMaxStoreBufs] each CTA: store issuer -> epilogue warps

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) P::prefetch_descs(prm);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar + 8 * s, 1 + kGenWarps);   // expect_tx + 4 generator warps per CTA (one group per stage)
            mbar_init(empty_bar + 8 * s, 1);
        }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 2 * kEpiWarps);
        for (int b = 0; b < kMaxStoreBufs; ++b) {
            mbar_init(staged_bar + 8 * b, kEpiWarps);
            mbar_init(freed_bar + 8 * b, 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(tmem_ptr_smem), 512);
        tmem_relinquish_pair();
    }
    P::setup(prm, extra, threadIdx.x, blockDim.x);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // barrier inits of both CTAs are visible before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int n_pairs = gridDim.x >> 1;
    const int pair = blockIdx.x >> 1;
    // strided schedule: tile costs vary smoothly with the tile index (zero-padding chunks are skipped), so
    // interleaving gives every pair the same mix
    // (policies with kContiguousTiles take contiguous ranges instead: equal-cost tiles whose per-tile tables only change
    // from image to image)
    int tile_begin = pair, tile_end = prm.num_tiles, tile_step = n_pairs;
    if constexpr (P::kContiguousTiles) {
        const int per = (prm.num_tiles + n_pairs - 1) / n_pairs;
        tile_begin = pair * per;
        tile_end = min(prm.num_tiles, tile_begin + per);
        tile_step = 1;
    }

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        // One elected lane: everything it does per chunk is on the critical path of the operand stream, so the
        // policy keeps its coordinates incrementally (TmaState) - no integer division per chunk.
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full_leader = mapa_rank(full_bar, 0);
            typename P::TmaState tst;
            for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
                PairTile ti;
                P::tile_info(prm, tile, rank, ti);
                P::tma_tile_begin(prm, ti, rank, tst);
                for (int q = ti.kc_begin; q < ti.kc_end; ++q) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t fb = full_leader + 8 * stage;
                    if (leader) mbar_arrive_expect_tx(full_bar + 8 * stage, 2u * ti.n_acc * kBHalfBytes + P::a_tx_bytes(prm));
                    const uint32_t sb = smem_u32(smem + L.stage_off + stage * kStage2Bytes + kAStageBytes);
                    P::tma_chunk(prm, ti, tst, sb, fb);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader only)
        if (leader && lane == 0) {
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
#ifdef TVAE_PROBE
            long long pr_tempty = 0, pr_full = 0;
            const long long pr_begin = clock64();
#endif
            for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
                PairTile ti;
                P::tile_info(prm, tile, rank, ti);
                { TVAE_PROBE_T0(); mbar_wait(tempty_bar, tphase ^ 1); TVAE_PROBE_ADD(pr_tempty); }
                tc_fence_after();
                for (int q = ti.kc_begin; q < ti.kc_end; ++q) {
                    { TVAE_PROBE_T0(); mbar_wait(full_bar + 8 * stage, phase); TVAE_PROBE_ADD(pr_full); }
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + L.stage_off + stage * kStage2Bytes);
                    const uint32_t b_addr = a_addr + kAStageBytes;
                    for (int a = 0; a < ti.n_acc; ++a) {
#pragma unroll
                        for (int ks = 0; ks < kKSteps; ++ks) {
                            // MN-major: generated A in 64-element blocks [64 k][128 B] (128 B swizzle), TMA-loaded B in
                            // 32-element blocks [64 k][64 B] (64 B swizzle); 16 k rows per MMA
                            uint64_t adesc, bdesc;
                            const uint32_t bb = b_addr + a * kBHalfBytes;
                            if (P::kAMajorMN) adesc = make_smem_desc(a_addr + ks * 2048, 64 * 128, 1024, kLayoutSw128);
                            else              adesc = make_smem_desc(a_addr + ks * 32, 16, 1024, kLayoutSw128);
                            if (P::kBMajorMN && P::kBSw128MN) bdesc = make_smem_desc(bb + ks * 2048, 64 * 128, 1024, kLayoutSw128);
                            else if (P::kBMajorMN) bdesc = make_smem_desc(bb + ks * 1024, 64 * 64, 512, kLayoutSw64);
                            else              bdesc = make_smem_desc(bb + ks * 32, 16, 1024, kLayoutSw128);
                            umma_f16_pair(tmem_base + a * kAccN, adesc, bdesc, kIdesc, (q > ti.kc_begin || ks > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit_pair(empty_bar + 8 * stage, 3);   // frees the slot in both CTAs
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(tfull_bar, 3);                   // accumulators complete -> both epilogues
                tphase ^= 1;
            }
#ifdef TVAE_PROBE
            atomicAdd(&g_pair_probe[P::kProbeSlot][0], (unsigned long long)(clock64() - pr_begin));
            atomicAdd(&g_pair_probe[P::kProbeSlot][1], (unsigned long long)pr_tempty);
            atomicAdd(&g_pair_probe[P::kProbeSlot][2], (unsigned long long)pr_full);
            atomicAdd(&g_pair_probe[P::kProbeSlot][7], 1ull);
#endif
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ store issuer (both CTAs)
        if constexpr (P::kStoreBufs > 0) {
            if (lane == 0) {
                constexpr int NB = P::kStoreBufs;
                uint32_t cnt = 0;                                  // 64-column blocks stored so far: buffer cnt % NB, use cnt / NB
#ifdef TVAE_PROBE
                long long pr_st[3] = {0, 0, 0};
#endif
                for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
                    PairTile ti;
                    P::tile_info(prm, tile, rank, ti);
                    const bool has_work = ti.kc_end > ti.kc_begin;
                    for (int a = 0; a < ti.n_acc; ++a) {
                        const int n0 = ti.n0 + a * kAccN;
                        const int nblk = P::store_blocks(prm, ti, n0, has_work);
                        for (int blk = 0; blk < nblk; ++blk, ++cnt) {
                            const uint32_t b = cnt % NB;
                            // Free the NEXT block's buffer first: all but the NB - 2 youngest stores have read their
                            // buffers, the oldest of them (NB - 1 back) used the buffer of block cnt + 1.  Doing this
                            // before waiting for block cnt keeps the writers from running in lockstep with this thread.
                            if constexpr (NB >= 2) {
                                { TVAE_PROBE_T0(); tma_store_wait_read<(NB >= 2 ? NB - 2 : 0)>(); TVAE_PROBE_ADD(pr_st[2]); }
                                if (cnt + 1 >= NB) mbar_arrive(freed_bar + 8 * ((cnt + 1) % NB));
                            }
                            { TVAE_PROBE_T0(); mbar_wait(staged_bar + 8 * b, (cnt / NB) & 1); TVAE_PROBE_ADD(pr_st[0]); }
                            { TVAE_PROBE_T0();
                            P::store_issue(prm, ti, n0, blk, smem_u32(extra + P::store_off(prm) + b * kStoreBlockBytes));
                            tma_store_commit();
                            TVAE_PROBE_ADD(pr_st[1]); }
                            if constexpr (NB == 1) {                // a single buffer: the writers wait for this very store
                                tma_store_wait_read<0>();
                                mbar_arrive(freed_bar);
                            }
                        }
                    }
                }
#ifdef TVAE_PROBE
                if (leader) for (int i = 0; i < 3; ++i) atomicAdd(&g_pair_probe[P::kProbeSlot == 1 ? 6 : 7][i], (unsigned long long)pr_st[i]);
                if (leader) atomicAdd(&g_pair_probe[P::kProbeSlot == 1 ? 6 : 7][3], (unsigned long long)cnt);
#endif
                tma_store_wait<0>();                               // staged stores have left shared memory before the CTA exits
            }
        }
    } else if (warp >= kFirstEpiWarp && warp < kFirstProdWarp) {
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        const int ewarp = warp - kFirstEpiWarp;
        const int row = ewarp * 32 + lane;
        uint32_t tphase = 0;
        const uint32_t tempty_leader = mapa_rank(tempty_bar, 0);
        typename P::EpiState est;
        P::epi_init(prm, est, extra, row);
        if constexpr (P::kStoreBufs > 0) { est.staged_bar = staged_bar; est.freed_bar = freed_bar; }
#ifdef TVAE_PROBE
        long long pr_tfull = 0, pr_epi = 0;
#endif
        for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
            PairTile ti;
            P::tile_info(prm, tile, rank, ti);
            P::epi_tile_begin(prm, ti, est, extra, row);     // per-tile tables, built while the tile's MMAs run
            { TVAE_PROBE_T0(); mbar_wait(tfull_bar, tphase); TVAE_PROBE_ADD(pr_tfull); }
            tc_fence_after();
            const bool has_work = ti.kc_end > ti.kc_begin;
            TVAE_PROBE_T0();
            for (int a = 0; a < ti.n_acc; ++a) {
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + a * kAccN;
                P::epilogue(prm, ti, est, ti.n0 + a * kAccN, taddr, row, has_work, extra);
            }
            P::epi_tile_end(prm, ti, est, extra, row, has_work);
            TVAE_PROBE_ADD(pr_epi);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader);
            tphase ^= 1;
        }
        P::epi_finish(prm, est, extra, row);
#ifdef TVAE_PROBE
        if (leader && row == 0) {
            atomicAdd(&g_pair_probe[P::kProbeSlot][3], (unsigned long long)pr_tfull);
            atomicAdd(&g_pair_probe[P::kProbeSlot][4], (unsigned long long)pr_epi);
        }
#endif
    } else if (warp >= kFirstProdWarp) {
        // ------------------------------------------------------------ operand generators (both CTAs)
        // Two groups of 4 warps fill alternate stages: the fixed per-stage latencies (barrier wait, smem round trips,
        // proxy fence, arrive) of one group hide behind the other group's gathers.  Both groups walk every chunk
        // (gen_advance) so that slab refills (gen_prepare, all 8 warps) happen at the same logical point.
        const int ptid = threadIdx.x - kFirstProdWarp * 32;
        const int grp = ptid >> 7, gtid = ptid & 127;
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t full_leader = mapa_rank(full_bar, 0);
        typename P::GenState gst;
        P::gen_init(prm, gst, extra, ptid);
#ifdef TVAE_PROBE
        long long pr_empty = 0, pr_gen = 0, pr_gtb = 0, pr_gprep = 0;
#endif
        for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
            PairTile ti;
            P::tile_info(prm, tile, rank, ti);
            { TVAE_PROBE_T0(); P::gen_tile_begin(prm, ti, gst, extra, ptid); TVAE_PROBE_ADD(pr_gtb); }
            for (int q = ti.kc_begin; q < ti.kc_end; ++q) {
                { TVAE_PROBE_T0(); P::gen_prepare(prm, ti, gst, extra, ptid); TVAE_PROBE_ADD(pr_gprep); }
                if (((q - ti.kc_begin) & 1) == grp) {
                    { TVAE_PROBE_T0(); mbar_wait(empty_bar + 8 * stage, phase ^ 1); TVAE_PROBE_ADD(pr_empty); }
                    TVAE_PROBE_T0();
                    P::gen_chunk(prm, ti, gst, smem + L.stage_off + stage * kStage2Bytes, extra, gtid);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(full_leader + 8 * stage);
                    TVAE_PROBE_ADD(pr_gen);
                }
                P::gen_advance(prm, ti, gst);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
        P::gen_finish(prm, gst);
#ifdef TVAE_PROBE
        if (leader && ptid == 0) {                             // generator thread 0 (group 0: every other chunk)
            atomicAdd(&g_pair_probe[P::kProbeSlot][5], (unsigned long long)pr_empty);
            atomicAdd(&g_pair_probe[P::kProbeSlot][6], (unsigned long long)pr_gen);
            atomicAdd(&g_pair_probe[P::kProbeSlot][14], (unsigned long long)pr_gtb);
            atomicAdd(&g_pair_probe[P::kProbeSlot][15], (unsigned long long)pr_gprep);
        }
#endif
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // no CTA exits (or frees TMEM) while its peer can still signal it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace tvae
