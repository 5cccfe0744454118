// Host-side launch helpers for tc_gemm_kernel policies.
#pragma once
#include "host_utils.cuh"
#include "linear_policies.cuh"
#include "tc_gemm2.cuh"
#include "linear_pair_policies.cuh"

namespace tvae {

constexpr int kMaxSmemBytes = 227 * 1024;

template <class P>
inline int pick_stages(int extra_bytes) {
    const int per_stage = kAStageBytes + (P::kBResidentChunks > 0 ? 0 : P::kBN * 128);   // a resident B is part of the fixed layout
    const SmemLayout L0 = make_smem_layout<P>(0, extra_bytes);
    int s = (kMaxSmemBytes - static_cast<int>(L0.total)) / per_stage;     // 227 KB is the opt-in maximum itself: no further margin
    if (s > kMaxStages) s = kMaxStages;
    return s;
}

template <class P>
inline int launch_gemm(typename P::Params& prm, int extra_bytes, cudaStream_t stream) {
    prm.num_stages = pick_stages<P>(extra_bytes);
    if (prm.num_stages < 2) return fail(-1, "tile does not fit shared memory with >= 2 stages");
    if (prm.num_tiles <= 0) return 0;
    const SmemLayout L = make_smem_layout<P>(prm.num_stages, extra_bytes);
    TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&tc_gemm_kernel<P>), kMaxSmemBytes));
    const int threads = (kCtrlWarps + kEpiWarps * P::kEpiGroups + P::kProdWarps) * 32;
    const int grid = prm.num_tiles < sm_count() ? prm.num_tiles : sm_count();
    ++g_launch_count;
    const int tslot = g_timer.begin(P::kName, stream);
    tc_gemm_kernel<P><<<grid, threads, L.total, stream>>>(prm);
    g_timer.end(tslot, stream);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// CTA-pair kernel: grid = 2 x min(#pairs on the device, tiles), cluster dims (2,1,1) are part of the kernel.
inline int pick_stages2(int extra_bytes) {
    const Smem2Layout L0 = make_smem2_layout(0, extra_bytes);
    int s = (kMaxSmemBytes - static_cast<int>(L0.total) - 1024) / kStage2Bytes;
    if (s > kMaxStages) s = kMaxStages;
    return s;
}
template <class P>
inline int launch_gemm2(typename P::Params& prm, int extra_bytes, cudaStream_t stream, int force_pairs = 0) {
    prm.num_stages = pick_stages2(extra_bytes);
    if (prm.num_stages < 2) return fail(-1, "pair tile does not fit shared memory with >= 2 stages");
    if (prm.num_tiles <= 0) return 0;
    const Smem2Layout L = make_smem2_layout(prm.num_stages, extra_bytes);
    TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&tc_gemm2_kernel<P>), kMaxSmemBytes));
    const int pairs_dev = sm_count() / 2;
    const int pairs = force_pairs > 0 ? force_pairs : (prm.num_tiles < pairs_dev ? prm.num_tiles : pairs_dev);
    ++g_launch_count;
    const int tslot = g_timer.begin(P::kName, stream);
    tc_gemm2_kernel<P><<<2 * pairs, kPairThreads, L.total, stream>>>(prm);
    g_timer.end(tslot, stream);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// C[M,N] = epi(A[M,K] B[N,K]^T); A, B fp16 row-major with leading dims lda, ldb (elements).
struct LinearNTArgs {
    const void* A; long long lda;      // fp16
    const void* B; long long ldb;      // fp16
    int M, N, K;
    float* C = nullptr; long long ldc = 0;
    const float* bias = nullptr;
    const float* row_bias = nullptr; int rows_per_group = 1; long long ld_rb = 0;
    long long ld_aux = 0;
    int act = 0;
    const float* proj_w = nullptr; const float* proj_bias = nullptr; float* proj_out = nullptr; int n_proj = 0;
    void* C16 = nullptr; long long ldc16 = 0;
    const void* aux16 = nullptr; int aux_act = 0;
    const unsigned long long* aux_bits = nullptr;
    unsigned long long* bits_out = nullptr;
    const float* acc_scale = nullptr; const float* store_scale = nullptr;
    float* colsum = nullptr; long long colsum_stride = 1;
};

inline int linear_nt(const LinearNTArgs& a, cudaStream_t stream) {
    TVAE_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "linear_nt: empty problem");
    TVAE_REQUIRE(a.N % 4 == 0 && a.N <= 1024, "linear_nt: N must be a multiple of 4, at most 1024");
    TVAE_REQUIRE(a.n_proj <= 4, "linear_nt: at most 4 fused projection outputs");
    LinearNTParams p{};
    const bool wide = a.N > 128;
    const int BN = wide ? 256 : 128;
    int rc;
    if ((rc = make_tmap_2d_h(&p.tmA, a.A, a.M, a.K, a.lda, kBM))) return rc;
    if ((rc = make_tmap_2d_h(&p.tmB, a.B, a.N, a.K, a.ldb, BN))) return rc;
    p.M = a.M; p.N = a.N;
    p.k_chunks = cdiv(a.K, kBKh);
    p.tiles_n = cdiv(a.N, BN);
    p.num_tiles = cdiv(a.M, kBM) * p.tiles_n;
    p.C = a.C; p.ldc = a.ldc; p.bias = a.bias;
    p.row_bias = a.row_bias; p.rows_per_group = a.rows_per_group; p.ld_rb = a.ld_rb;
    p.ld_aux = a.ld_aux; p.act = a.act; p.aux_act = a.aux_act;
    p.proj_w = a.proj_w; p.proj_bias = a.proj_bias; p.proj_out = a.proj_out; p.n_proj = a.n_proj;
    p.C16 = a.C16; p.ldc16 = a.ldc16; p.colsum = a.colsum; p.colsum_stride = a.colsum_stride;
    p.aux16 = a.aux16; p.aux_bits = a.aux_bits; p.acc_scale = a.acc_scale; p.store_scale = a.store_scale;
    TVAE_REQUIRE(!a.aux_bits || a.N % 64 == 0, "linear_nt: the one-bit mask needs N % 64 == 0");
    // extra smem: bias / projection / column-sum rows sized by the actual N, then (fp16 output with whole 64-column
    // blocks) the staging buffers of the TMA stores: two 16 KB buffers per epilogue group
    p.npad = (a.N + 31) / 32 * 32;
    p.cs_off = (1 + a.n_proj) * p.npad;
    int extra = (LinearNT<128>::extra_floats(a.N, a.n_proj, a.colsum != nullptr) * 4 + 1023) / 1024 * 1024;
    p.tma_store = (a.C16 != nullptr && a.N % 64 == 0) ? 1 : 0;
    TVAE_REQUIRE(!a.bits_out || p.tma_store, "linear_nt: the one-bit mask output needs an fp16 output with N % 64 == 0");
    p.bits_out = a.bits_out;
    p.stage_off = extra;
    if (p.tma_store) {
        if ((rc = make_tmap_2d_h(&p.tmC, a.C16, a.M, a.N, a.ldc16, kBM))) return rc;
        extra += LinearNT<128>::kEpiGroups * 2 * LinearNT<128>::kStageBytes;
    }
    if (a.act == kActTanh || (a.aux16 && a.aux_act == kActTanh))
        return wide ? launch_gemm<LinearNT<256, true>>(p, extra, stream) : launch_gemm<LinearNT<128, true>>(p, extra, stream);
    return wide ? launch_gemm<LinearNT<256>>(p, extra, stream) : launch_gemm<LinearNT<128>>(p, extra, stream);
}

// C[Ma,Nb] (+)= sum_r P[r,Ma] Q[r,Nb], P and Q fp16; caller zero-fills C (accumulated with atomics across row splits).
// CTA-pair variant (linear_pair_policies.cuh) for wide outputs: 256 x 512 pair tiles halve the L2 -> SM operand traffic.
// Computes Ct[nb][ma] (i.e. stores the product transposed); returns 1 when the shape is not covered.
inline bool g_linear_pair_enabled = true;          // test hook
inline int linear_tn_pair(const void* P, long long ldp, const void* Q, long long ldq, int R, int Ma, int Nb, float* Ct, long long ldc,
                          cudaStream_t stream, const float* acc_scale) {
    if (!g_linear_pair_enabled) return 1;
    if (Nb % 32 != 0 || Nb > 2 * kAccN || Nb <= 128 || Ma <= 128 || (ldp & 7) || (reinterpret_cast<uintptr_t>(P) & 15)) return 1;
    LinearTNPairParams p{};
    int rc;
    if ((rc = make_tmap_3d_mn_h(&p.tmQ, Q, R, Nb, ldq, kBKh, 4))) return rc;
    p.P = static_cast<const __half*>(P); p.ldp = ldp; p.M = R; p.Ma = Ma; p.Nb = Nb; p.C = Ct; p.ldc = ldc; p.acc_scale = acc_scale;
    p.m_tiles = cdiv(Ma, kBM);
    p.m_pairs = cdiv(p.m_tiles, 2);
    p.chunks_total = cdiv(R, kBKh);
    const int pairs_dev = sm_count() / 2;
    int splits = pairs_dev / p.m_pairs;
    if (splits > p.chunks_total / 8) splits = p.chunks_total / 8;
    if (splits < 1) splits = 1;
    p.chunks_per_split = cdiv(p.chunks_total, splits);
    p.splits = cdiv(p.chunks_total, p.chunks_per_split);
    p.num_tiles = p.m_pairs * p.splits;
    p.a_tma = (Ma % 64 == 0 && g_dev_knob[3] == 0) ? 1 : 0;
    if (p.a_tma && (rc = make_tmap_3d_mn128_h(&p.tmP, P, R, Ma, ldp, kBKh, 2))) return rc;
    return launch_gemm2<LinearTNPair>(p, 0, stream);
}

inline int linear_tn(const void* P, long long ldp, const void* Q, long long ldq, int R, int Ma, int Nb,
                     float* C, long long ldc, int transpose_out, cudaStream_t stream, const float* acc_scale = nullptr) {
    TVAE_REQUIRE(R > 0 && Ma > 0 && Nb > 0, "linear_tn: empty problem");
    {
        // the pair kernel stores its accumulator transposed (coalesced atomics): C[ma][nb] is the transposed store of the
        // product with the operands swapped
        const int rc = transpose_out ? linear_tn_pair(P, ldp, Q, ldq, R, Ma, Nb, C, ldc, stream, acc_scale)
                                     : linear_tn_pair(Q, ldq, P, ldp, R, Nb, Ma, C, ldc, stream, acc_scale);
        if (rc <= 0) return rc;
    }
    LinearTNParams p{};
    const bool wide = Nb > 128;
    const int BN = wide ? 256 : 128;
    int rc;
    if ((rc = make_tmap_2d_mn_h(&p.tmP, P, R, Ma, ldp, kBKh))) return rc;
    if ((rc = make_tmap_2d_mn_h(&p.tmQ, Q, R, Nb, ldq, kBKh))) return rc;
    p.Ma = Ma; p.Nb = Nb;
    p.acc_scale = acc_scale;
    p.tiles_m = cdiv(Ma, kBM);
    p.tiles_n = cdiv(Nb, BN);
    p.chunks_total = cdiv(R, kBKh);
    const int out_tiles = p.tiles_m * p.tiles_n;
    int splits = cdiv(2 * sm_count(), out_tiles);        // ~2 waves of work items
    const int min_chunks = 16;                            // keep >= 512 reduction rows per split
    if (splits > cdiv(p.chunks_total, min_chunks)) splits = cdiv(p.chunks_total, min_chunks);
    if (splits < 1) splits = 1;
    p.chunks_per_split = cdiv(p.chunks_total, splits);
    p.splits = cdiv(p.chunks_total, p.chunks_per_split);
    p.num_tiles = out_tiles * p.splits;
    p.C = C; p.ldc = ldc; p.transpose_out = transpose_out;
    return wide ? launch_gemm<LinearTN<256>>(p, 0, stream) : launch_gemm<LinearTN<128>>(p, 0, stream);
}

}  // namespace tvae
