// Encoder backward, conv2 stage, for O = 128: one pass over dhpre and x1 produces BOTH
//
//   dx1pre[m][o] = (sum_o' dhpre[m][o'] W2[o'][o]) * lrelu'(x1[m][o])      (stored fp16 * s2, operand of the conv1 wgrad)
//   dW2[o'][o]  += sum_m dhpre[m][o'] x1[m][o]                               (conv2.weight gradient)
//   db1[o]      += sum_m dx1pre[m][o]                                        (conv1 bias gradient, column sums)
//
// (autograd of nn.Conv3d(O, O, 1) + LeakyReLU, models.py:347,355-356).  The two-kernel version (LinearTN for dW2, LinearNT
// for dx1) streams dhpre and x1 from HBM twice - 4 x 2 O bytes per row read; here each 128-row tile of dhpre and x1 is
// loaded once by TMA into shared memory and used three ways:
//   * dhpre tile, K-major view   -> A of   D1[128 rows x 128 o]  = dhpre . W2          (W2^T resident in smem)
//   * dhpre tile, MN-major view  -> A of   D2[128 o' x 128 o]   += dhpre^T . x1         (same bytes: a 128 B-swizzled
//     [row][64 halves] tile is both a K-major operand over its columns and an MN-major operand over its rows)
//   * x1 tile, MN-major view     -> B of D2, and the LeakyReLU mask of the epilogue (sign bits read from smem)
// D2 stays in TMEM for the CTA's whole tile range and is flushed once with fp32 atomics.
//
// Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 / 8-11 two epilogue groups on alternate tiles (group g always
// works on tile slot g and accumulator stage g).
#pragma once
#include "linear_policies.cuh"

namespace tvae {

struct EncDx1Dw2Params {
    CUtensorMap tmA;          // dhpre fp16 [R][128], boxes {64, 128 rows}
    CUtensorMap tmX;          // x1 fp16 [R][128], boxes {64, 128 rows}
    CUtensorMap tmW;          // W2^T fp16 [128 o][128 o'], boxes {64, 128 rows}
    CUtensorMap tmC;          // dx1 fp16 [R][128] store view, boxes {64, 128 rows}
    long long R;
    int num_tiles;
    const float* acc_scale;   // device scalar 1 / s1 (dhpre was stored * s1)
    const float* store_scale; // device scalar s2 applied to the fp16 store of dx1
    float* colsum;            // colsum[o * colsum_stride] += sum_m dx1pre[m][o]
    long long colsum_stride;
    float* dw2;               // [128][128] fp32, zero-filled by the caller
};

constexpr int kFuThreads = 384;
constexpr int kFuChunk = kBM * 128;           // [128 rows][64 halves] = 16 KB
constexpr int kFuTile = 2 * kFuChunk;         // 128 x 128 halves = 32 KB
constexpr int kFuStageOff = kFuTile + 4 * kFuTile;            // W | slot0 {A, X} | slot1 {A, X}
constexpr int kFuBarOff = kFuStageOff + 4 * kFuChunk;         // staging: 2 groups x 2 buffers
constexpr int kFuSmemBytes = kFuBarOff + 16 * 8 + 16;

__global__ void __launch_bounds__(kFuThreads, 1) enc_dx1_dw2_kernel(const __grid_constant__ EncDx1Dw2Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem;
    uint8_t* s_slot = smem + kFuTile;
    uint8_t* s_stage = smem + kFuStageOff;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFuBarOff);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);
    const uint32_t full_bar = smem_u32(bars), empty_bar = smem_u32(bars + 2), tfull_bar = smem_u32(bars + 4),
                   tempty_bar = smem_u32(bars + 6), dwfull_bar = smem_u32(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA); tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmW); tma_prefetch_desc(&p.tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1 + kEpiWarps * 32);     // MMA commit + the epilogue group's mask reads
            mbar_init(tfull_bar + 8 * s, 1);
            mbar_init(tempty_bar + 8 * s, kEpiWarps * 32);
        }
        mbar_init(dwfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_ptr_smem), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const long long nt = p.num_tiles;
    const int tile_begin = static_cast<int>(nt * blockIdx.x / gridDim.x);
    const int tile_end = static_cast<int>(nt * (blockIdx.x + 1) / gridDim.x);
    const int n_tiles = tile_end - tile_begin;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                const int m0 = (tile_begin + i) * kBM;
                mbar_wait(empty_bar + 8 * s, ph ^ 1);
                const uint32_t fb = full_bar + 8 * s;
                mbar_arrive_expect_tx(fb, 2 * kFuTile + (i == 0 ? kFuTile : 0));
                if (i == 0) {
                    tma_load_2d(smem_u32(s_w), &p.tmW, fb, 0, 0);
                    tma_load_2d(smem_u32(s_w + kFuChunk), &p.tmW, fb, 64, 0);
                }
                uint8_t* a = s_slot + s * 2 * kFuTile;
                tma_load_2d(smem_u32(a), &p.tmA, fb, 0, m0);
                tma_load_2d(smem_u32(a + kFuChunk), &p.tmA, fb, 64, m0);
                tma_load_2d(smem_u32(a + kFuTile), &p.tmX, fb, 0, m0);
                tma_load_2d(smem_u32(a + kFuTile + kFuChunk), &p.tmX, fb, 64, m0);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t kIdescMain = make_idesc_f16(kBM, 128, false, false, 0, 0);
            constexpr uint32_t kIdescDw = make_idesc_f16(kBM, 128, true, true, 0, 0);
            const uint32_t w_addr = smem_u32(s_w);
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                mbar_wait(tempty_bar + 8 * s, ph ^ 1);
                mbar_wait(full_bar + 8 * s, ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(s_slot + s * 2 * kFuTile), x_addr = a_addr + kFuTile;
                const uint32_t d1 = tmem_base + s * 128;
#pragma unroll
                for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
                    for (int ks = 0; ks < kKSteps; ++ks) {
                        const uint64_t adesc = make_smem_desc(a_addr + kc * kFuChunk + ks * 32, 16, 1024, kLayoutSw128);
                        const uint64_t bdesc = make_smem_desc(w_addr + kc * kFuChunk + ks * 32, 16, 1024, kLayoutSw128);
                        umma_f16(d1, adesc, bdesc, kIdescMain, (kc | ks) ? 1u : 0u);
                    }
                }
                umma_commit(tfull_bar + 8 * s);                       // dx1 accumulator -> epilogue group s
                const uint32_t d2 = tmem_base + 256;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                      // reduction over the tile's 128 rows, 16 per MMA
                    const uint64_t adesc = make_smem_desc(a_addr + ks * 2048, kFuChunk, 1024, kLayoutSw128);
                    const uint64_t bdesc = make_smem_desc(x_addr + ks * 2048, kFuChunk, 1024, kLayoutSw128);
                    umma_f16(d2, adesc, bdesc, kIdescDw, (i > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar + 8 * s);                       // tile slot read by the tensor core
            }
            if (n_tiles > 0) umma_commit(dwfull_bar);
        }
    } else if (warp >= kFirstEpiWarp) {
        // ------------------------------------------------------------ epilogue groups
        const int ewarp = (warp - kFirstEpiWarp) & 3, grp = (warp - kFirstEpiWarp) >> 2;
        const int row = ewarp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(ewarp * 32) << 16;
        const float acc_scale = __ldg(p.acc_scale), store_scale = __ldg(p.store_scale);
        const int bar_id = 2 + grp;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
        int blocks = 0;
        const uint8_t* x_tile = s_slot + grp * 2 * kFuTile + kFuTile;
        for (int i = grp; i < n_tiles; i += 2) {
            const uint32_t ph = (i >> 1) & 1;
            const int m0 = (tile_begin + i) * kBM;
            mbar_wait(tfull_bar + 8 * grp, ph);
            mbar_wait(full_bar + 8 * grp, ph);                        // the TMA-written x1 tile is visible to this thread
            tc_fence_after();
            // LeakyReLU mask of this row: sign bits of x1[m][0..127] from the staged tile
            uint32_t mbits[4] = {0u, 0u, 0u, 0u};
            const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const uint4 q = *reinterpret_cast<const uint4*>(x_tile + (u >> 3) * kFuChunk + sw128_offset(row, u & 7));
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                uint32_t b8 = 0u;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned m = __hgt2_mask(*reinterpret_cast<const __half2*>(&w[e]), zero2);
                    b8 |= ((m & 1u) | ((m >> 15) & 2u)) << (2 * e);
                }
                mbits[u >> 2] |= b8 << (8 * (u & 3));
            }
            mbar_arrive(empty_bar + 8 * grp);                          // this thread is done with the tile slot
            const uint32_t taddr = tmem_base + lane_off + grp * 128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                tmem_ld_wait();
                if (c == 3) {                                          // accumulator stage fully read
                    tc_fence_before();
                    mbar_arrive(tempty_bar + 8 * grp);
                }
                float v[32];
                const uint32_t mb = c == 0 ? mbits[0] : (c == 1 ? mbits[1] : (c == 2 ? mbits[2] : mbits[3]));   // registers, not local memory
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * acc_scale * (((mb >> j) & 1u) ? 1.f : kLreluSlope);
                {
                    const float colsum = warp_colsum32(v, lane);
                    if (c == 0) cs[0] += colsum; else if (c == 1) cs[1] += colsum; else if (c == 2) cs[2] += colsum; else cs[3] += colsum;
                }
                uint8_t* buf = s_stage + (grp * 2 + (blocks & 1)) * kFuChunk;
                if ((c & 1) == 0 && blocks >= 2) {
                    if (row == 0) tma_store_wait_read<1>();
                    named_bar_sync(bar_id, kEpiWarps * 32);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 t;
                    __half2 h;
                    h = __floats2half2_rn(v[j] * store_scale, v[j + 1] * store_scale);     t.x = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[j + 2] * store_scale, v[j + 3] * store_scale); t.y = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[j + 4] * store_scale, v[j + 5] * store_scale); t.z = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[j + 6] * store_scale, v[j + 7] * store_scale); t.w = *reinterpret_cast<uint32_t*>(&h);
                    *reinterpret_cast<uint4*>(buf + sw128_offset(row, (c & 1) * 4 + (j >> 3))) = t;
                }
                if (c & 1) {
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, kEpiWarps * 32);
                    if (row == 0) {
                        tma_store_2d(&p.tmC, smem_u32(buf), (c >> 1) * 64, m0);
                        tma_store_commit();
                    }
                    ++blocks;
                }
            }
        }
        if (row == 0) tma_store_wait<0>();
        // conv1 bias gradient: this warp's column sums (lane l owns column c*32 + l)
        if (p.colsum) {
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(p.colsum + (long long)(c * 32 + lane) * p.colsum_stride, cs[c]);
        }
        // dW2: accumulator rows = o' (this thread's TMEM lane), columns = o; group g flushes columns [64 g, 64 g + 64)
        if (n_tiles > 0) {
            mbar_wait(dwfull_bar, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c = grp * 2; c < grp * 2 + 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + lane_off + 256 + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(p.dw2 + row * 128 + c * 32 + j, __uint_as_float(r[j]) * acc_scale);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// ================================================================================================
// Heads backward for O = 128, NH <= 32 on the tensor core (replaces the warp-MMA thin backward for this shape):
//
//   dhpre[m][o'] = (sum_t d_heads[m][t] Wh[t][o']) * lrelu'(h[m][o'])      (stored fp16 * s1)
//   dWh[t][o']  += sum_m d_heads[m][t] h[m][o']        dbh[t] += sum_m d_heads[m][t]        db2[o'] += sum_m dhpre[m][o']
//
// (autograd of conv_a / conv_r / conv_z + LeakyReLU, models.py:356-358,390-392).  Streaming kernel: one read of h and of
// the planar fp32 head-map gradients, one write of dhpre.  Per 128-row tile
//   * the h tile arrives by TMA (128 B swizzle), the gradient tile S = fp16(d_heads) [128 rows][64 halves, zero beyond NH]
//     is written by four generator warps (thread = row, coalesced planar reads),
//   * D1[128 rows x 128 o'] = S . Wh        (A = S K-major, B = Wh^T tile resident in smem; one or two K = 16 steps)
//   * D2[128 o' x NHP]     += h^T . S       (A = h tile read MN-major, B = S tile read MN-major; resident in TMEM, flushed once)
//   * epilogue: LeakyReLU mask from the sign bits of the staged h tile, column sums (db2), fp16 * s1 through swizzled
//     staging buffers and TMA stores.
// Warps: 0 TMA, 1 MMA, 2 TMEM allocator, 4-7 / 8-11 epilogue groups (alternate tiles), 12-15 generators.
struct EncHeadsBwdParams {
    CUtensorMap tmH;          // h fp16 [R][128], boxes {64, 128 rows}
    CUtensorMap tmC;          // dhpre fp16 [R][128] store view
    long long R;
    int num_tiles, NH, G, P;
    const float* d_heads;     // (B, NH, G, P) fp32
    const float* wh;          // [NH][128] fp32
    const float* store_scale; // device scalar s1
    const float* in_scale;    // device scalar s0 (power of two): S = fp16(d_heads * s0), divided out of dhpre / dWh again
    float* dwh;               // [NH][128], zero-filled by the caller
    float* dbh;               // [NH]
    float* db2;               // [128]
};

constexpr int kHbThreads = 512;
constexpr int kHbSlotBytes = kFuTile + kFuChunk;              // h tile + S tile = 48 KB
constexpr int kHbStageOff = kFuChunk + 2 * kHbSlotBytes;      // Wh tile | slot0 | slot1
constexpr int kHbBarOff = kHbStageOff + 4 * kFuChunk;
constexpr int kHbSmemBytes = kHbBarOff + 16 * 8 + 16;

template <int NHP>
__global__ void __launch_bounds__(kHbThreads, 1) enc_heads_bwd_kernel(const __grid_constant__ EncHeadsBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_wh = smem;
    uint8_t* s_slot = smem + kFuChunk;
    uint8_t* s_stage = smem + kHbStageOff;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHbBarOff);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);
    const uint32_t fullh_bar = smem_u32(bars), fulls_bar = smem_u32(bars + 2), empty_bar = smem_u32(bars + 4),
                   tfull_bar = smem_u32(bars + 6), tempty_bar = smem_u32(bars + 8), d2full_bar = smem_u32(bars + 10);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmH); tma_prefetch_desc(&p.tmC); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(fullh_bar + 8 * s, 1);
            mbar_init(fulls_bar + 8 * s, 128);
            mbar_init(empty_bar + 8 * s, 1 + kEpiWarps * 32);
            mbar_init(tfull_bar + 8 * s, 1);
            mbar_init(tempty_bar + 8 * s, kEpiWarps * 32);
        }
        mbar_init(d2full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_ptr_smem), 512);
        tmem_relinquish();
    }
    // Wh^T as the K-major B operand of D1: row o', element t at 16-byte unit t >> 3 (zero beyond NH)
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
        const int o = i >> 6, t = i & 63;
        const float w = t < p.NH ? __ldg(p.wh + t * 128 + o) : 0.f;
        *reinterpret_cast<__half*>(s_wh + sw128_offset(o, t >> 3) + (t & 7) * 2) = __float2half_rn(w);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const long long nt = p.num_tiles;
    const int tile_begin = static_cast<int>(nt * blockIdx.x / gridDim.x);
    const int tile_end = static_cast<int>(nt * (blockIdx.x + 1) / gridDim.x);
    const int n_tiles = tile_end - tile_begin;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer: h tiles
        if (lane == 0) {
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                const int m0 = (tile_begin + i) * kBM;
                mbar_wait(empty_bar + 8 * s, ph ^ 1);
                const uint32_t fb = fullh_bar + 8 * s;
                mbar_arrive_expect_tx(fb, kFuTile);
                uint8_t* h = s_slot + s * kHbSlotBytes;
                tma_load_2d(smem_u32(h), &p.tmH, fb, 0, m0);
                tma_load_2d(smem_u32(h + kFuChunk), &p.tmH, fb, 64, m0);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t kIdesc1 = make_idesc_f16(kBM, 128, false, false, 0, 0);
            constexpr uint32_t kIdesc2 = make_idesc_f16(kBM, NHP, true, true, 0, 0);
            const uint32_t wh_addr = smem_u32(s_wh);
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                const uint32_t h_addr = smem_u32(s_slot + s * kHbSlotBytes), s_addr = h_addr + kFuTile;
                mbar_wait(tempty_bar + 8 * s, ph ^ 1);
                mbar_wait(fulls_bar + 8 * s, ph);
                tc_fence_after();
                const uint32_t d1 = tmem_base + s * 128;
#pragma unroll
                for (int ks = 0; ks < NHP / 16; ++ks) {
                    const uint64_t adesc = make_smem_desc(s_addr + ks * 32, 16, 1024, kLayoutSw128);
                    const uint64_t bdesc = make_smem_desc(wh_addr + ks * 32, 16, 1024, kLayoutSw128);
                    umma_f16(d1, adesc, bdesc, kIdesc1, ks ? 1u : 0u);
                }
                umma_commit(tfull_bar + 8 * s);
                mbar_wait(fullh_bar + 8 * s, ph);
                tc_fence_after();
                const uint32_t d2 = tmem_base + 256;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                      // reduction over the tile's 128 rows
                    const uint64_t adesc = make_smem_desc(h_addr + ks * 2048, kFuChunk, 1024, kLayoutSw128);
                    const uint64_t bdesc = make_smem_desc(s_addr + ks * 2048, kFuChunk, 1024, kLayoutSw128);
                    umma_f16(d2, adesc, bdesc, kIdesc2, (i > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar + 8 * s);
            }
            if (n_tiles > 0) umma_commit(d2full_bar);
        }
    } else if (warp >= 12) {
        // ------------------------------------------------------------ generators: S = fp16(d_heads) tiles, dbh
        const int row = threadIdx.x - 12 * 32;
        const long long chan = (long long)p.G * p.P;
        const float s0 = __ldg(p.in_scale);
        float dbh[NHP];
#pragma unroll
        for (int t = 0; t < NHP; ++t) dbh[t] = 0.f;
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            const long long m = (long long)(tile_begin + i) * kBM + row;
            float v[NHP];
#pragma unroll
            for (int t = 0; t < NHP; ++t) v[t] = 0.f;
            if (m < p.R) {
                const long long br = m / p.P;
                const int pos = static_cast<int>(m - br * p.P);
                const long long b = br / p.G;
                const int r = static_cast<int>(br - b * p.G);
                const float* src = p.d_heads + b * p.NH * chan + (long long)r * p.P + pos;
#pragma unroll
                for (int t = 0; t < NHP; ++t)
                    if (t < p.NH) v[t] = __ldg(src + t * chan);
            }
#pragma unroll
            for (int t = 0; t < NHP; ++t) dbh[t] += v[t];
            mbar_wait(empty_bar + 8 * s, ph ^ 1);
            uint8_t* st = s_slot + s * kHbSlotBytes + kFuTile;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (u < NHP / 8) {
                    __half2 hh;
                    hh = __floats2half2_rn(v[8 * u] * s0, v[8 * u + 1] * s0);     q.x = *reinterpret_cast<uint32_t*>(&hh);
                    hh = __floats2half2_rn(v[8 * u + 2] * s0, v[8 * u + 3] * s0); q.y = *reinterpret_cast<uint32_t*>(&hh);
                    hh = __floats2half2_rn(v[8 * u + 4] * s0, v[8 * u + 5] * s0); q.z = *reinterpret_cast<uint32_t*>(&hh);
                    hh = __floats2half2_rn(v[8 * u + 6] * s0, v[8 * u + 7] * s0); q.w = *reinterpret_cast<uint32_t*>(&hh);
                }
                *reinterpret_cast<uint4*>(st + sw128_offset(row, u)) = q;
            }
            fence_proxy_async_smem();
            mbar_arrive(fulls_bar + 8 * s);
        }
        if (p.dbh) {
#pragma unroll
            for (int t = 0; t < NHP; ++t) {
                const float sum = warp_sum(dbh[t]);
                if (lane == 0 && t < p.NH) atomicAdd(p.dbh + t, sum);
            }
        }
    } else if (warp >= kFirstEpiWarp) {
        // ------------------------------------------------------------ epilogue groups
        const int ewarp = (warp - kFirstEpiWarp) & 3, grp = (warp - kFirstEpiWarp) >> 2;
        const int row = ewarp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(ewarp * 32) << 16;
        const float store_scale = __ldg(p.store_scale), inv_s0 = 1.f / __ldg(p.in_scale);
        const int bar_id = 2 + grp;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
        int blocks = 0;
        const uint8_t* h_tile = s_slot + grp * kHbSlotBytes;
        for (int i = grp; i < n_tiles; i += 2) {
            const uint32_t ph = (i >> 1) & 1;
            const int m0 = (tile_begin + i) * kBM;
            mbar_wait(tfull_bar + 8 * grp, ph);
            mbar_wait(fullh_bar + 8 * grp, ph);
            tc_fence_after();
            uint32_t mbits[4] = {0u, 0u, 0u, 0u};
            const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const uint4 q = *reinterpret_cast<const uint4*>(h_tile + (u >> 3) * kFuChunk + sw128_offset(row, u & 7));
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                uint32_t b8 = 0u;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned m = __hgt2_mask(*reinterpret_cast<const __half2*>(&w[e]), zero2);
                    b8 |= ((m & 1u) | ((m >> 15) & 2u)) << (2 * e);
                }
                mbits[u >> 2] |= b8 << (8 * (u & 3));
            }
            mbar_arrive(empty_bar + 8 * grp);
            const uint32_t taddr = tmem_base + lane_off + grp * 128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                tmem_ld_wait();
                if (c == 3) {
                    tc_fence_before();
                    mbar_arrive(tempty_bar + 8 * grp);
                }
                float v[32];
                const uint32_t mb = c == 0 ? mbits[0] : (c == 1 ? mbits[1] : (c == 2 ? mbits[2] : mbits[3]));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * (((mb >> j) & 1u) ? inv_s0 : kLreluSlope * inv_s0);
                {
                    const float colsum = warp_colsum32(v, lane);
                    if (c == 0) cs[0] += colsum; else if (c == 1) cs[1] += colsum; else if (c == 2) cs[2] += colsum; else cs[3] += colsum;
                }
                uint8_t* buf = s_stage + (grp * 2 + (blocks & 1)) * kFuChunk;
                if ((c & 1) == 0 && blocks >= 2) {
                    if (row == 0) tma_store_wait_read<1>();
                    named_bar_sync(bar_id, kEpiWarps * 32);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 t;
                    __half2 h;
                    h = __floats2half2_rn(v[j] * store_scale, v[j + 1] * store_scale);     t.x = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[j + 2] * store_scale, v[j + 3] * store_scale); t.y = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[j + 4] * store_scale, v[j + 5] * store_scale); t.z = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[j + 6] * store_scale, v[j + 7] * store_scale); t.w = *reinterpret_cast<uint32_t*>(&h);
                    *reinterpret_cast<uint4*>(buf + sw128_offset(row, (c & 1) * 4 + (j >> 3))) = t;
                }
                if (c & 1) {
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, kEpiWarps * 32);
                    if (row == 0) {
                        tma_store_2d(&p.tmC, smem_u32(buf), (c >> 1) * 64, m0);
                        tma_store_commit();
                    }
                    ++blocks;
                }
            }
        }
        if (row == 0) tma_store_wait<0>();
        if (p.db2) {
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(p.db2 + c * 32 + lane, cs[c]);
        }
        // dWh[t][o'] += D2[o'][t]: accumulator rows = o' (this thread's TMEM lane); group 0 flushes
        if (grp == 0 && n_tiles > 0) {
            mbar_wait(d2full_bar, 0);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < NHP / 16; ++c) {
                uint32_t r[16];
                tmem_ld_32x16(tmem_base + lane_off + 256 + c * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c * 16 + j < p.NH) atomicAdd(p.dwh + (c * 16 + j) * 128 + row, __uint_as_float(r[j]) * inv_s0);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tvae
