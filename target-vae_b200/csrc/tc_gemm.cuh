// Warp-specialised persistent FP16 GEMM core for sm_100a (tcgen05.mma.kind::f16 + TMEM + TMA).
//
//   D[128 x BN] (fp32, TMEM) += A[128 x K] * B[BN x K]^T        (A, B fp16: the 11-bit significand of TF32 at twice its rate)
//
// One CTA per SM.  Warp roles:
//   warp 0      TMA producer (one lane): bulk-tensor loads into a ring of smem stages
//   warp 1      MMA issuer   (one lane): tcgen05.mma.kind::f16 on swizzled smem descriptors
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld accumulator rows -> registers -> policy epilogue; policies whose epilogue
//               is the bottleneck (streaming GEMMs with a short K loop) run P::kEpiGroups = 2 such warpgroups that
//               take alternate tiles (each tile has its own TMEM accumulator stage)
//   then        optional operand generators (im2col window / Fourier features / ...) that
//               synthesise the A tile straight into swizzled smem instead of loading it
// Three pipelines: smem full/empty (producers <-> MMA), TMEM full/empty (MMA <-> epilogue),
// and a static contiguous tile schedule per CTA.
//
// Operand layouts in smem (fp16 = 2 bytes; a stage covers kBKh = 64 reduction elements, one MMA consumes 16):
//   K-major  operand: [rows][64 k] in the 128 B swizzle, 8-row atoms of 1024 B, SBO = 1024, K step = +32 B
//   MN-major operand: [col-block][64 k rows][32 mn] per 32-wide block of the M/N extent in the 64 B swizzle
//                     (8-row atoms of 512 B): LBO = block stride (4096 B), SBO = 512, K step (16 rows) = +1024 B
// A policy `P` supplies tile geometry, TMA issue, optional generator and the epilogue.
#pragma once
#include "ptx.cuh"

namespace tvae {

constexpr int kBM = 128;        // accumulator rows (TMEM lanes)
constexpr int kBKh = 64;        // fp16 elements per smem stage along the reduction (one 128 B swizzle row)
constexpr int kKSteps = 4;      // tcgen05.mma per stage (16 fp16 reduction elements each)
constexpr int kAStageBytes = kBM * 128;
constexpr int kCtrlWarps = 4;
constexpr int kEpiWarps = 4;
constexpr int kFirstEpiWarp = kCtrlWarps;
constexpr int kFirstProdWarp = kCtrlWarps + kEpiWarps;
constexpr int kMaxStages = 8;

struct TileInfo {
    int m0;        // first accumulator row in the policy's row space
    int n0;        // first accumulator column
    int kc_begin;  // reduction chunks [kc_begin, kc_end) of kBKh elements each
    int kc_end;
    int a0, a1, a2;  // policy scratch (image index, split index, ...)
};

struct SmemLayout {
    uint32_t a_off, b_off, bar_off, tmem_ptr_off, extra_off, total;
};

template <class P>
__host__ __device__ inline SmemLayout make_smem_layout(int stages, int extra_bytes) {
    SmemLayout L;
    L.a_off = 0;
    L.b_off = L.a_off + stages * kAStageBytes;
    // B ring, or (kBResidentChunks > 0) the whole B operand kept resident: its K chunks never change from tile to tile
    L.bar_off = L.b_off + (P::kBResidentChunks > 0 ? P::kBResidentChunks : stages) * (P::kBN * 128);
    L.tmem_ptr_off = L.bar_off + (2 * kMaxStages + 8) * 8;
    L.extra_off = (L.tmem_ptr_off + 16 + 1023) & ~1023u;    // policies may keep 128 B-swizzled tiles in their extra region
    L.total = L.extra_off + extra_bytes;
    return L;
}

template <class P>
__global__ void __launch_bounds__((kCtrlWarps + kEpiWarps * P::kEpiGroups + P::kProdWarps) * 32, 1)
tc_gemm_kernel(const __grid_constant__ typename P::Params prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kBN = P::kBN;
    constexpr int kAccStagesFit = (512 / kBN) > 4 ? 4 : (512 / kBN);
    constexpr int kAccStages = kAccStagesFit < P::kMaxAccStages ? kAccStagesFit : P::kMaxAccStages;   // policies may keep TMEM columns for themselves
    constexpr int kTmemCols = 512;
    constexpr int kBStageBytes = kBN * 128;
    constexpr uint32_t kIdesc = make_idesc_f16(kBM, kBN, P::kAMajorMN, P::kBMajorMN, P::kAFmt, P::kBFmt);
    static_assert(kBN % 16 == 0 && kBN >= 16 && kBN <= 256, "invalid UMMA N");
    constexpr int kEpiGroups = P::kEpiGroups;
    constexpr int kFirstGenWarp = kFirstEpiWarp + kEpiWarps * kEpiGroups;
    static_assert(kEpiGroups == 1 || (kEpiGroups == 2 && kAccStages % 2 == 0), "two epilogue groups need an even number of accumulator stages");

    const int stages = prm.num_stages;
    const SmemLayout L = make_smem_layout<P>(stages, 0);
    uint8_t* smem_a = smem + L.a_off;
    uint8_t* smem_b = smem + L.b_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L.tmem_ptr_off);
    uint8_t* extra = smem + L.extra_off;

    const uint32_t full_bar = smem_u32(bars);                       // [stages]
    const uint32_t empty_bar = smem_u32(bars + kMaxStages);          // [stages]
    const uint32_t tfull_bar = smem_u32(bars + 2 * kMaxStages);      // [kAccStages]
    const uint32_t tempty_bar = smem_u32(bars + 2 * kMaxStages + 4); // [kAccStages]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        P::prefetch_descs(prm);
    }
    if (warp == 1 && lane == 0) {
        const uint32_t full_count = 1 + (P::kAGen ? P::kProdWarps * 32 : 0);
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar + 8 * s, full_count);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(tfull_bar + 8 * s, 1);
            mbar_init(tempty_bar + 8 * s, kEpiWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_ptr_smem), kTmemCols);
        tmem_relinquish();
    }
    P::setup(prm, extra, threadIdx.x, blockDim.x);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const long long nt = prm.num_tiles;
    const int tile_begin = static_cast<int>(nt * blockIdx.x / gridDim.x);
    const int tile_end = static_cast<int>(nt * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                TileInfo ti;
                P::tile_info(prm, tile, ti);
                for (int kc = ti.kc_begin; kc < ti.kc_end; ++kc) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t fb = full_bar + 8 * stage;
                    if constexpr (P::kBResidentChunks > 0) {
                        // resident B: chunk kc is loaded once, with this CTA's first tile; later stages carry A only
                        const bool first = tile == tile_begin;
                        mbar_arrive_expect_tx(fb, kAStageBytes + (first ? kBStageBytes : 0));
                        P::issue_tma_a(prm, ti, kc, smem_u32(smem_a + stage * kAStageBytes), fb);
                        if (first) P::issue_tma_b(prm, ti, kc, smem_u32(smem_b + (kc - ti.kc_begin) * kBStageBytes), fb);
                    } else {
                        mbar_arrive_expect_tx(fb, P::tx_bytes());
                        P::issue_tma(prm, ti, kc, smem_u32(smem_a + stage * kAStageBytes),
                                     smem_u32(smem_b + stage * kBStageBytes), fb);
                    }
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                TileInfo ti;
                P::tile_info(prm, tile, ti);
                mbar_wait(tempty_bar + 8 * as, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * kBN;
                for (int kc = ti.kc_begin; kc < ti.kc_end; ++kc) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * kAStageBytes);
                    const uint32_t b_addr = smem_u32(smem_b + (P::kBResidentChunks > 0 ? kc - ti.kc_begin : stage) * kBStageBytes);
#pragma unroll
                    for (int ks = 0; ks < kKSteps; ++ks) {
                        uint64_t adesc, bdesc;
                        if (P::kAMajorMN) adesc = make_smem_desc(a_addr + ks * 1024, kBKh * 64, 512, kLayoutSw64);
                        else              adesc = make_smem_desc(a_addr + ks * 32, 16, 1024, kLayoutSw128);
                        if (P::kBMajorMN) bdesc = make_smem_desc(b_addr + ks * 1024, kBKh * 64, 512, kLayoutSw64);
                        else              bdesc = make_smem_desc(b_addr + ks * 32, 16, 1024, kLayoutSw128);
                        umma_f16(d_tmem, adesc, bdesc, kIdesc, (kc > ti.kc_begin || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar + 8 * stage);  // frees the smem slot once these MMAs retire
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar + 8 * as);  // accumulator complete -> epilogue
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= kFirstEpiWarp && warp < kFirstGenWarp) {
        // ------------------------------------------------------------ epilogue
        const int ewarp = (warp - kFirstEpiWarp) & 3;  // == warp % 4 -> TMEM lane quarter
        const int egrp = (warp - kFirstEpiWarp) >> 2;  // epilogue group: takes tiles t = egrp, egrp + kEpiGroups, ...
        const int row = ewarp * 32 + lane;
        typename P::EpiState est;
        P::epi_init(prm, est, extra, egrp * kBM + row);     // slot: unique per epilogue thread
        for (int tile = tile_begin + egrp, t = egrp; tile < tile_end; tile += kEpiGroups, t += kEpiGroups) {
            TileInfo ti;
            P::tile_info(prm, tile, ti);
            const int as = t % kAccStages;
            const uint32_t aphase = (t / kAccStages) & 1;
            mbar_wait(tfull_bar + 8 * as, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + as * kBN;
            if constexpr (P::kEpiSelfRelease) {
                // the policy releases the accumulator itself (as soon as it has read it) and may use the TMEM columns
                // past the accumulator stages
                est.tempty = tempty_bar + 8 * as;
                est.tmem_free = tmem_base + kAccStages * kBN;      // lane 0 of the first free column
                P::epilogue(prm, ti, est, taddr, row, extra);
            } else {
                P::epilogue(prm, ti, est, taddr, row, extra);
                tc_fence_before();
                mbar_arrive(tempty_bar + 8 * as);
            }
        }
        P::epi_finish(prm, est, extra, egrp * kBM + row);
    } else if (P::kAGen && warp >= kFirstGenWarp) {
        // ------------------------------------------------------------ operand generators
        const int ptid = threadIdx.x - kFirstGenWarp * 32;
        int stage = 0;
        uint32_t phase = 0;
        typename P::GenState gst;
        P::gen_init(prm, gst, extra, ptid);
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            TileInfo ti;
            P::tile_info(prm, tile, ti);
            P::gen_tile_begin(prm, ti, gst, extra, ptid);
            for (int kc = ti.kc_begin; kc < ti.kc_end; ++kc) {
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                P::gen_chunk(prm, ti, gst, kc, smem_a + stage * kAStageBytes, extra, ptid);
                fence_proxy_async_smem();
                mbar_arrive(full_bar + 8 * stage);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// Defaults a policy can inherit.
struct PolicyBase {
    static constexpr bool kAMajorMN = false;
    static constexpr bool kBMajorMN = false;
    static constexpr bool kAGen = false;
    static constexpr int kProdWarps = 0;
    static constexpr int kEpiGroups = 1;      // tc_gemm: epilogue warpgroups taking alternate tiles
    static constexpr int kMaxAccStages = 4;   // tc_gemm: upper bound on the TMEM accumulator stages
    static constexpr bool kBSw128MN = false;  // tc_gemm2: MN-major B staged in 64-column blocks / 128 B swizzle instead of 32-column / 64 B
    static constexpr int kBResidentChunks = 0;   // tc_gemm: > 0 = B has this many K chunks, identical for every tile, kept in smem
    static constexpr bool kEpiSelfRelease = false;   // tc_gemm: the policy's epilogue arrives on EpiState::tempty itself
    static constexpr uint32_t kAFmt = 0;      // operand formats of kind::f16: 0 = FP16, 1 = BF16 (both operands must agree)
    static constexpr uint32_t kBFmt = 0;
    static constexpr int kProbeSlot = 0;      // tc_gemm2, -DTVAE_PROBE builds: row of the probe counters this policy adds to
    struct EpiState {};
    struct GenState {};
    template <class Prm> __device__ static void setup(const Prm&, uint8_t*, int, int) {}
    template <class Prm, class St> __device__ static void epi_init(const Prm&, St&, uint8_t*, int) {}
    template <class Prm, class St> __device__ static void epi_finish(const Prm&, St&, uint8_t*, int) {}
    template <class Prm, class Ti, class St> __device__ static void epi_tile_begin(const Prm&, const Ti&, St&, uint8_t*, int) {}   // tc_gemm2
    template <class Prm, class Ti, class St> __device__ static void epi_tile_end(const Prm&, const Ti&, St&, uint8_t*, int, bool) {}   // tc_gemm2
    template <class Prm, class St> __device__ static void gen_finish(const Prm&, St&) {}                        // tc_gemm2: after the last tile
    template <class Prm> __device__ static uint32_t a_tx_bytes(const Prm&) { return 0u; }      // tc_gemm2: A operand loaded by TMA
    static constexpr bool kContiguousTiles = false;   // tc_gemm2: a pair takes a contiguous range of tiles instead of every n_pairs-th
    static constexpr int kStoreBufs = 0;      // tc_gemm2: > 0 = the epilogue leaves through that many staging buffers and the store issuer warp
    template <class Prm, class St> __device__ static void gen_init(const Prm&, St&, uint8_t*, int) {}
    template <class Prm, class St> __device__ static void gen_tile_begin(const Prm&, const TileInfo&, St&, uint8_t*, int) {}
    template <class Prm, class St> __device__ static void gen_chunk(const Prm&, const TileInfo&, St&, int, uint8_t*, uint8_t*, int) {}
};

// named barrier among a subset of warps (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Issue helpers shared by the policies -----------------------------------------------------------
// K-major operand tile: `rows` rows x 64 k-elements at (k = kc*64, row = r0): one box.
__device__ __forceinline__ void tma_kmajor_h(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int kc, int r0) {
    tma_load_2d(dst, tm, bar, kc * kBKh, r0);
}
// MN-major operand tile: 64 reduction rows x (32*nblk) features starting at feature f0, rows r0..: one
// {32 feat x 64 rows} box (64 B swizzle) per 32-wide feature block.
__device__ __forceinline__ void tma_mnmajor_h(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int f0, int r0, int nblk) {
    for (int b = 0; b < nblk; ++b) tma_load_2d(dst + b * (kBKh * 64), tm, bar, f0 + 32 * b, r0);
}
}  // namespace tvae
