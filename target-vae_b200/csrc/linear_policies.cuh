// GEMM policies whose operands are both loaded by TMA (fp16 operands, kind::f16, fp32 accumulation).
//   LinearNT<BN>: C[M,N] = epi(A[M,K] * B[N,K]^T)      (forward linear layers, dgrads with pre-transposed weights)
//   LinearTN<BN>: C[Ma,Nb] += sum_r P[r,Ma] * Q[r,Nb]  (weight gradients; split over r, fp32 atomics)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tc_gemm.cuh"

namespace tvae {

constexpr float kLreluSlope = 0.01f;

// max(x, slope * x) == (x > 0 ? x : slope * x) for every non-NaN x (0 < slope < 1), one instruction shorter
__device__ __forceinline__ float lrelu(float x) { return fmaxf(x, kLreluSlope * x); }
__device__ __forceinline__ float lrelu_grad_from_out(float a) { return a > 0.f ? 1.f : kLreluSlope; }

// Activation selector carried by the shape structs and kernel parameters (--activation leakyrelu | tanh,
// train_mnist.py:423,516-519): kActTanh selects tanh, every other value (zero-initialised parameter blocks) LeakyReLU(0.01).
// Both derivatives are functions of the activation's OUTPUT, which is what the backward kernels have at hand.
constexpr int kActTanh = 2;
__device__ __forceinline__ float act_grad_from_out(float a, int act) { return act == kActTanh ? 1.f - a * a : lrelu_grad_from_out(a); }
// The run-time form is used by the CUDA-core backward kernels only (one FMA + select per element).
// Compile-time variant for the tensor-core epilogues: their LeakyReLU instantiations must stay instruction-identical to
// a build without tanh (a run-time selector, even hoisted, cost gen_l1_fwd +32 % and conv2_heads +11 %).
template <bool TANH, int N>
__device__ __forceinline__ void act_vec(float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = TANH ? tanhf(v[i]) : lrelu(v[i]);
}

struct LinearNTParams {
    CUtensorMap tmA, tmB;
    int num_stages, num_tiles, tiles_n, k_chunks;
    int M, N;
    float* C;                 // [M][ldc] or null
    long long ldc;
    const float* bias;        // [N] or null
    const float* row_bias;    // [M / rows_per_group][ld_rb] or null   (z-conditioned bias)
    int rows_per_group;
    long long ld_rb;
    long long ld_aux;
    int act;                  // 0 = none, else the activation of the instantiation (LinearNT<BN, TANH>)
    int aux_act;              // informational: the host picks LinearNT<BN, true> when act or aux_act is kActTanh
    const float* proj_w;      // [n_proj][N] or null: fused  proj_out[m][o] += sum_n v[m][n] * proj_w[o][n]
    const float* proj_bias;   // [n_proj]
    float* proj_out;          // [M][n_proj], pre-zeroed
    int n_proj;
    void* C16;                // [M][ldc16] fp16 output (value * *store_scale) or null
    long long ldc16;
    const void* aux16;        // [M][ld_aux] fp16 or null: multiply by lrelu'(aux16)
    const unsigned long long* aux_bits;   // [N/64][M] or null: bit q of word (n/64, m) set = derivative 1, clear = LeakyReLU slope
                                          // (the one-bit form of aux16, written by GenL1FwdPairT<., 1>; N % 64 == 0)
    unsigned long long* bits_out;         // [N/64][M] or null: the same one-bit mask of the fp16 values stored here (forward: what the
                                          // NEXT layer's input gradient needs of this layer's activation); needs tma_store
    const float* acc_scale;   // device scalar multiplied into the accumulator first (undoes the operand's scale) or null
    const float* store_scale; // device scalar applied to the fp16 store only (power of two) or null
    float* colsum;            // colsum[n * colsum_stride] += sum_m value[m][n] (bias gradient) or null
    long long colsum_stride;
    CUtensorMap tmC;          // C16 as [M][N] fp16, boxes {64, 128 rows}, for the staged TMA stores (tma_store == 1)
    int tma_store;            // 1: C16 leaves the SM through a swizzled smem staging buffer + TMA store (N % 64 == 0)
    int stage_off;            // byte offset of the staging buffers (two 16 KB buffers per epilogue group) in the extra smem
    int npad;                 // N rounded up to 32: row pitch (floats) of the bias / projection / column-sum rows in the extra smem
    int cs_off;               // float offset of the column-sum rows
};

// In: v[j] of lane l = value (row l, column j) of a 32 x 32 block.  Out (returned): in lane l, the sum over the 32
// rows of column l.  Recursive halving: 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_colsum32(const float (&v)[32], int lane) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const bool up = lane & 16;
        const float keep = up ? v[j + 16] : v[j], send = up ? v[j] : v[j + 16];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {
#pragma unroll
        for (int j = 0; j < w; ++j) {
            const bool up = lane & w;
            const float keep = up ? a[j + w] : a[j], send = up ? a[j] : a[j + w];
            a[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    return a[0];
}

template <int BN, bool TANH = false>
struct LinearNT : PolicyBase {
    static constexpr const char* kName = "linear_nt";
    using Params = LinearNTParams;
    static constexpr int kBN = BN;
    static constexpr bool kF16 = true;         // fp16 A and B (TMA boxes of 64 k-elements)
    static constexpr int kEpiGroups = 2;       // short K loops: the epilogue is the critical path, two warpgroups alternate tiles
    // extra smem, rows of npad floats: bias, n_proj fused-projection weight rows, then (with colsum) one column-sum row
    // per epilogue warp (entry (warp, column) with column % 32 == lane is owned by one thread); sized by the launcher
    // from the actual N so that short rows leave the shared memory to the operand ring
    __host__ static int extra_floats(int N, int n_proj, bool colsum) {
        const int npad = (N + 31) / 32 * 32;
        return (1 + n_proj + (colsum ? kEpiGroups * kEpiWarps : 0)) * npad;
    }
    static constexpr int kStageBytes = kBM * 128;      // [128 rows][64 halves], 128 B swizzle
    struct EpiState { int cs, grp, blocks; };
    __device__ static void setup(const Params& p, uint8_t* extra, int tid, int nthreads) {
        float* s = reinterpret_cast<float*>(extra);
        for (int i = tid; i < p.N; i += nthreads) s[i] = p.bias ? __ldg(p.bias + i) : 0.f;
        for (int i = tid; i < p.n_proj * p.N; i += nthreads) s[(1 + i / p.N) * p.npad + (i % p.N)] = __ldg(p.proj_w + i);
    }
    __device__ static void epi_init(const Params& p, EpiState& st, uint8_t* extra, int slot) {
        st.cs = p.cs_off + (slot >> 5) * p.npad + (slot & 31);
        st.grp = slot >> 7;
        st.blocks = 0;
        if (!p.colsum) return;
        float* cs = reinterpret_cast<float*>(extra) + st.cs;
        for (int c = 0; c * 32 < p.N; ++c) cs[c * 32] = 0.f;
    }
    __device__ static void epi_finish(const Params& p, EpiState& st, uint8_t* extra, int slot) {
        if (p.tma_store && (slot & 127) == 0) tma_store_wait<0>();    // staged stores have left shared memory
        if (!p.colsum) return;
        const int lane = slot & 31;
        const float* cs = reinterpret_cast<const float*>(extra) + st.cs;
        for (int c = 0; c * 32 < p.N; ++c)
            if (c * 32 + lane < p.N) atomicAdd(p.colsum + (long long)(c * 32 + lane) * p.colsum_stride, cs[c * 32]);
    }
    __device__ static void prefetch_descs(const Params& p) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        if (p.tma_store) tma_prefetch_desc(&p.tmC);
    }
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        const int mt = tile / p.tiles_n, nt = tile - mt * p.tiles_n;
        ti.m0 = mt * kBM;
        ti.n0 = nt * BN;
        ti.kc_begin = 0;
        ti.kc_end = p.k_chunks;
    }
    __device__ static constexpr uint32_t tx_bytes() { return kAStageBytes + BN * 128; }
    __device__ static void issue_tma(const Params& p, const TileInfo& ti, int kc, uint32_t sa, uint32_t sb, uint32_t bar) {
        tma_kmajor_h(sa, &p.tmA, bar, kc, ti.m0);
        tma_kmajor_h(sb, &p.tmB, bar, kc, ti.n0);
    }
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState& st, uint32_t taddr, int row, uint8_t* extra) {
        const float* s_bias = reinterpret_cast<const float*>(extra);
        const int m = ti.m0 + row;
        const bool m_ok = m < p.M;
        float proj[4] = {0.f, 0.f, 0.f, 0.f};
        const float acc_scale = p.acc_scale ? __ldg(p.acc_scale) : 1.f;
        const float store_scale = p.store_scale ? __ldg(p.store_scale) : 1.f;
        const float* rb = (p.row_bias && m_ok) ? p.row_bias + (long long)(m / p.rows_per_group) * p.ld_rb : nullptr;
        // LeakyReLU-mask operand (aux16): its global loads are the longest latency of this epilogue.  They run one
        // 32-column block ahead of their use, and the rows the group's NEXT tile will need are pulled into L2 now.
        const __half* ax_row = (p.aux16 && m_ok) ? reinterpret_cast<const __half*>(p.aux16) + (long long)m * p.ld_aux + ti.n0 : nullptr;
        uint4 ax_next[4];
        if (ax_row) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ti.n0 + 8 * j < p.N) ax_next[j] = __ldg(reinterpret_cast<const uint4*>(ax_row) + j);
            const int tile_next = (ti.m0 / kBM) * p.tiles_n + ti.n0 / BN + kEpiGroups;
            const int mt = tile_next / p.tiles_n, nt = tile_next - mt * p.tiles_n;
            const long long mn = (long long)mt * kBM + row;
            if (mn < p.M) {
                const char* pf = reinterpret_cast<const char*>(reinterpret_cast<const __half*>(p.aux16) + mn * p.ld_aux + nt * BN);
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                    if (nt * BN + 64 * j < p.N) prefetch_l2(pf + 128 * j);
            }
        }
        // one-bit LeakyReLU mask: BN / 64 words per row, lanes = consecutive rows (coalesced)
        unsigned long long mbits[BN / 64];
        if (p.aux_bits) {
#pragma unroll
            for (int w = 0; w < BN / 64; ++w)
                mbits[w] = (m_ok && ti.n0 + 64 * w < p.N) ? __ldg(p.aux_bits + (long long)(ti.n0 / 64 + w) * p.M + m) : 0ull;
        }
        uint32_t neg_lo = 0u;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            uint4 ax_cur[4];
            if (ax_row) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ax_cur[j] = ax_next[j];
                if (c + 1 < BN / 32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (ti.n0 + (c + 1) * 32 + 8 * j < p.N) ax_next[j] = __ldg(reinterpret_cast<const uint4*>(ax_row + (c + 1) * 32) + j);
                }
            }
            tmem_ld_wait();
            const int n_base = ti.n0 + c * 32;
            if (p.colsum || p.tma_store) {        // uniform path: every thread takes part in the shuffles / barriers
                if (n_base >= p.N) continue;
            } else if (!m_ok || n_base >= p.N) {
                continue;
            }
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * acc_scale;
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (n_base + j < p.N) {
                        const float4 t = *reinterpret_cast<const float4*>(s_bias + n_base + j);
                        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                    }
                }
            }
            if (rb) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (n_base + j < p.N) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(rb + n_base + j));
                        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                    }
                }
            }
            if (p.act) act_vec<TANH>(v);
            if (ax_row) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (n_base + j < p.N) {
                        const uint4 t = ax_cur[j >> 3];
                        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            // lrelu'(a) from the sign of the packed halves; tanh'(a) = 1 - a^2
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                            v[j + 2 * e] *= TANH ? 1.f - f.x * f.x : lrelu_grad_from_out(f.x);
                            v[j + 2 * e + 1] *= TANH ? 1.f - f.y * f.y : lrelu_grad_from_out(f.y);
                        }
                    }
                }
            }
            if (p.aux_bits) {
                unsigned long long wsel = mbits[0];
#pragma unroll
                for (int w = 1; w < BN / 64; ++w) wsel = (c >> 1) == w ? mbits[w] : wsel;
                const uint32_t bits = static_cast<uint32_t>(wsel >> ((c & 1) * 32));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= ((bits >> j) & 1u) ? 1.f : kLreluSlope;
            }
            if (p.colsum) {
                if (!m_ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                reinterpret_cast<float*>(extra)[st.cs + n_base] += warp_colsum32(v, row & 31);
                if (!m_ok && !p.tma_store) continue;
            }
            if (p.tma_store) {
                // 64-column blocks: the even 32-column group opens a block (the group's staging buffer must have been
                // read by the previous store), the odd one closes it and hands it to the TMA store unit; rows >= M
                // are clipped by the tensor map
                // two staging buffers per group, alternating: a block only waits for the store issued two blocks ago
                uint8_t* buf = extra + p.stage_off + (st.grp * 2 + (st.blocks & 1)) * kStageBytes;
                if ((c & 1) == 0 && st.blocks >= 2) {
                    if (row == 0) tma_store_wait_read<1>();
                    named_bar_sync(2 + st.grp, kEpiWarps * 32);
                }
                uint32_t neg = 0u;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 t;
                    __half2 h;
                    h = __floats2half2_rn(v[j] * store_scale, v[j + 1] * store_scale);     t.x = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[j + 2] * store_scale, v[j + 3] * store_scale); t.y = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[j + 4] * store_scale, v[j + 5] * store_scale); t.z = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[j + 6] * store_scale, v[j + 7] * store_scale); t.w = *reinterpret_cast<uint32_t*>(&h);
                    *reinterpret_cast<uint4*>(buf + sw128_offset(row, (c & 1) * 4 + (j >> 3))) = t;
                    if (p.bits_out) neg |= half8_sign_bits(t) << j;
                }
                if (p.bits_out) {
                    if (c & 1) {
                        if (m_ok) p.bits_out[(long long)((n_base - 32) >> 6) * p.M + m] = ~((static_cast<unsigned long long>(neg) << 32) | neg_lo);
                    } else {
                        neg_lo = neg;
                    }
                }
                if (c & 1) {
                    fence_proxy_async_smem();
                    named_bar_sync(2 + st.grp, kEpiWarps * 32);
                    if (row == 0) {
                        tma_store_2d(&p.tmC, smem_u32(buf), n_base - 32, ti.m0);
                        tma_store_commit();
                    }
                    ++st.blocks;
                }
                if (!m_ok) continue;
            } else if (p.C16) {
                __half* d16 = reinterpret_cast<__half*>(p.C16) + (long long)m * p.ldc16 + n_base;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (n_base + j < p.N) {
                        uint4 t;
                        __half2 h;
                        h = __floats2half2_rn(v[j] * store_scale, v[j + 1] * store_scale);     t.x = *reinterpret_cast<uint32_t*>(&h);
                        h = __floats2half2_rn(v[j + 2] * store_scale, v[j + 3] * store_scale); t.y = *reinterpret_cast<uint32_t*>(&h);
                        h = __floats2half2_rn(v[j + 4] * store_scale, v[j + 5] * store_scale); t.z = *reinterpret_cast<uint32_t*>(&h);
                        h = __floats2half2_rn(v[j + 6] * store_scale, v[j + 7] * store_scale); t.w = *reinterpret_cast<uint32_t*>(&h);
                        *reinterpret_cast<uint4*>(d16 + j) = t;
                    }
                }
            }
            if (p.proj_w) {
                for (int o = 0; o < p.n_proj; ++o) {
                    const float* w = s_bias + (1 + o) * p.npad + n_base;
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (n_base + j < p.N) {
                            const float4 t = *reinterpret_cast<const float4*>(w + j);
                            acc = fmaf(v[j], t.x, acc); acc = fmaf(v[j + 1], t.y, acc);
                            acc = fmaf(v[j + 2], t.z, acc); acc = fmaf(v[j + 3], t.w, acc);
                        }
                    }
                    proj[o] += acc;
                }
            }
            if (p.C) {
                float* dst = p.C + (long long)m * p.ldc + n_base;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (n_base + j < p.N) {
                        *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
            }
        }
        if (p.proj_w && m_ok) {
            for (int o = 0; o < p.n_proj; ++o) {
                float add = proj[o];
                if (ti.n0 == 0 && p.proj_bias) add += __ldg(p.proj_bias + o);
                atomicAdd(p.proj_out + (long long)m * p.n_proj + o, add);
            }
        }
    }
};

struct LinearTNParams {
    CUtensorMap tmP, tmQ;     // P [R][Ma], Q [R][Nb]  (row-major, reduction over R)
    int num_stages, num_tiles, tiles_m, tiles_n, splits, chunks_total, chunks_per_split;
    int Ma, Nb;
    float* C;                 // accumulated with atomics; caller zero-fills
    long long ldc;
    int transpose_out;        // 0: C[ma][nb], 1: C[nb][ma]
    const float* acc_scale;   // device scalar multiplied into the accumulator (undoes the operands' scale) or null
};

template <int BN>
struct LinearTN : PolicyBase {
    static constexpr const char* kName = "linear_tn";
    using Params = LinearTNParams;
    static constexpr int kBN = BN;
    static constexpr bool kF16 = true;         // fp16 P and Q, reduction chunks of 64 rows
    static constexpr bool kAMajorMN = true;
    static constexpr bool kBMajorMN = true;
    __device__ static void prefetch_descs(const Params& p) {
        tma_prefetch_desc(&p.tmP);
        tma_prefetch_desc(&p.tmQ);
    }
    __device__ static void tile_info(const Params& p, int tile, TileInfo& ti) {
        const int per_split = p.tiles_m * p.tiles_n;
        const int sp = tile / per_split;
        const int rem = tile - sp * per_split;
        const int mt = rem / p.tiles_n, nt = rem - mt * p.tiles_n;
        ti.m0 = mt * kBM;
        ti.n0 = nt * BN;
        ti.kc_begin = sp * p.chunks_per_split;
        ti.kc_end = min(ti.kc_begin + p.chunks_per_split, p.chunks_total);
    }
    __device__ static constexpr uint32_t tx_bytes() { return kAStageBytes + BN * 128; }
    __device__ static void issue_tma(const Params& p, const TileInfo& ti, int kc, uint32_t sa, uint32_t sb, uint32_t bar) {
        tma_mnmajor_h(sa, &p.tmP, bar, ti.m0, kc * kBKh, kBM / 32);
        tma_mnmajor_h(sb, &p.tmQ, bar, ti.n0, kc * kBKh, BN / 32);
    }
    __device__ static void epilogue(const Params& p, const TileInfo& ti, EpiState&, uint32_t taddr, int row, uint8_t*) {
        const int ma = ti.m0 + row;
        const bool empty = ti.kc_begin >= ti.kc_end;
        const float acc_scale = p.acc_scale ? __ldg(p.acc_scale) : 1.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
            if (ma >= p.Ma || empty) continue;
            const int nb0 = ti.n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int nb = nb0 + j;
                if (nb < p.Nb) {
                    float* dst = p.transpose_out ? p.C + (long long)nb * p.ldc + ma : p.C + (long long)ma * p.ldc + nb;
                    atomicAdd(dst, __uint_as_float(r[j]) * acc_scale);
                }
            }
        }
    }
};

}  // namespace tvae
