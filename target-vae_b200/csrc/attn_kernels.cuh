// a-3 tail + a-4: attention inference over the head maps (models.py:383-387, train_mnist.py:187-282), HBM-bound.
//
//   heads (B, NH, G, P) planar, channel order [attn(+p_r), theta_mu(+offset), theta_logstd, z_mu[0..z), z_logstd[0..z)],
//   L = G*P cells per image.
//
// Forward: a thread-block CLUSTER per image (1, 2, 4 or 8 CTAs, chosen so that B images fill the 148 SMs several times
// over - one CTA per image left a third of the SMs idle at B = 100 and gave each SM one latency-bound stream).  Every CTA
// owns a contiguous slice of the L cells and reads it with 16-byte loads:
//   sweep 1  online max / sum-exp of the two softmaxes (logits, logits + Gumbel)        -> per-CTA partials in shared memory
//            cluster barrier; every CTA combines the partials of all ranks through distributed shared memory
//   sweep 2  expectations under the Gumbel-softmax sample and the KL sums under q        -> per-CTA partials, cluster barrier,
//            rank 0 combines them in rank order (bit-deterministic) and writes the per-image outputs; a last cluster
//            barrier keeps every CTA's shared memory alive until rank 0 has read it
// The slice's logits / noise are read twice; the second read hits L1/L2 (a slice is a few tens of KB).
// Backward: elementwise over (image, cell) with 16-byte loads and stores.
// Per-cell exponentials / logarithms use the MUFU approximations (__expf / __logf, ~2 ulp): with 3 + 2z of each per cell the
// precise library forms made the kernels instruction-bound; the sums they feed are compared with fp64 at 1e-4.
#pragma once
#include <cooperative_groups.h>

#include "simt_kernels.cuh"

namespace tvae {

namespace cg = cooperative_groups;

constexpr int kAttnThreads = 256;     // 3 CTAs per SM: the sweeps are latency-bound streams, occupancy is what hides them

// online softmax statistics: (m, s) with s = sum exp(x - m)
__device__ __forceinline__ void online_add(float& m, float& s, float x) {
    if (x > m) { s = s * __expf(m - x) + 1.f; m = x; }
    else s += __expf(x - m);
}
__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
    if (m2 == -CUDART_INF_F) return;
    if (m == -CUDART_INF_F) { m = m2; s = s2; return; }
    if (m2 > m) { s = s * __expf(m - m2) + s2; m = m2; }
    else s += s2 * __expf(m2 - m);
}

template <int VEC> struct AttnVec;
template <> struct AttnVec<4> {
    __device__ static void load(const float* p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ static void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct AttnVec<1> {
    __device__ static void load(const float* p, float (&v)[1]) { v[0] = *p; }
    __device__ static void store(float* p, const float (&v)[1]) { *p = v[0]; }
};

// grid = B * CL CTAs, cluster (CL,1,1); image b = blockIdx.x / CL
template <int Z, int VEC>
__global__ void __launch_bounds__(kAttnThreads, 3) attn_fwd_kernel(AttnParams p) {
    constexpr int NA = 6 + 4 * Z;
    __shared__ float scratch[32 * NA];
    __shared__ float s_stat[4];                     // this CTA's softmax statistics, read by its cluster peers after sweep 1
    __shared__ float s_part[NA];                    // this CTA's partial sums, read by rank 0 after sweep 2
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = static_cast<int>(cluster.num_blocks());
    const int rank = static_cast<int>(cluster.block_rank());
    const int b = blockIdx.x / CL;
    const int P = p.d * p.d, L = p.G * P, NH = 3 + 2 * Z;
    const float* hb = p.heads + (long long)b * NH * L;
    const float* gb = p.gumbel + (long long)b * L;
    // this CTA's slice [l0, l1): whole vectors
    const int nvec = L / VEC;
    const int v0 = static_cast<int>((long long)nvec * rank / CL), v1 = static_cast<int>((long long)nvec * (rank + 1) / CL);
    const int l0 = v0 * VEC, l1 = v1 * VEC;

    // ---- sweep 1: online (max, sum-exp) of logits and logits + Gumbel
    float mq = -CUDART_INF_F, sq = 0.f, ma = -CUDART_INF_F, sa = 0.f;
    for (int l = l0 + threadIdx.x * VEC; l < l1; l += 2 * kAttnThreads * VEC) {
        // two vectors per thread and round: twice the bytes in flight per memory-latency round
        const int l2 = l + kAttnThreads * VEC;
        const bool two = l2 < l1;
        float a[VEC], g[VEC], a2[VEC], g2[VEC];
        AttnVec<VEC>::load(hb + l, a);
        AttnVec<VEC>::load(gb + l, g);
        if (two) {
            AttnVec<VEC>::load(hb + l2, a2);
            AttnVec<VEC>::load(gb + l2, g2);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            online_add(mq, sq, a[e]);
            online_add(ma, sa, a[e] + g[e]);
        }
        if (two) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                online_add(mq, sq, a2[e]);
                online_add(ma, sa, a2[e] + g2[e]);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        online_merge(mq, sq, __shfl_xor_sync(0xffffffffu, mq, o), __shfl_xor_sync(0xffffffffu, sq, o));
        online_merge(ma, sa, __shfl_xor_sync(0xffffffffu, ma, o), __shfl_xor_sync(0xffffffffu, sa, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kAttnThreads / 32;
    if (lane == 0) { scratch[warp * 4 + 0] = mq; scratch[warp * 4 + 1] = sq; scratch[warp * 4 + 2] = ma; scratch[warp * 4 + 3] = sa; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m1 = -CUDART_INF_F, s1 = 0.f, m2 = -CUDART_INF_F, s2 = 0.f;
        for (int w = 0; w < nw; ++w) {
            online_merge(m1, s1, scratch[w * 4 + 0], scratch[w * 4 + 1]);
            online_merge(m2, s2, scratch[w * 4 + 2], scratch[w * 4 + 3]);
        }
        s_stat[0] = m1; s_stat[1] = s1; s_stat[2] = m2; s_stat[3] = s2;
    }
    cluster.sync();
    {
        float m1 = -CUDART_INF_F, s1 = 0.f, m2 = -CUDART_INF_F, s2 = 0.f;
        for (int r = 0; r < CL; ++r) {                        // rank order: every CTA of the cluster gets the same bits
            const float* rp = cluster.map_shared_rank(s_stat, r);
            online_merge(m1, s1, rp[0], rp[1]);
            online_merge(m2, s2, rp[2], rp[3]);
        }
        mq = m1; sq = s1; ma = m2; sa = s2;
    }
    const float lse_q = mq + logf(sq), lse_a = ma + logf(sa);

    // ---- sweep 2: expectations under a (Gumbel-softmax sample) and KL sums under pi = exp(q)
    float acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.f;
    const float inv2s2 = 1.f / (2.f * p.theta_prior_std * p.theta_prior_std);
    const float log_sp = logf(p.theta_prior_std);
    for (int l = l0 + threadIdx.x * VEC; l < l1; l += kAttnThreads * VEC) {
        float lg[VEC], g[VEC], tm[VEC], tl[VEC], lp[VEC];
        AttnVec<VEC>::load(hb + l, lg);
        AttnVec<VEC>::load(gb + l, g);
        AttnVec<VEC>::load(hb + L + l, tm);
        AttnVec<VEC>::load(hb + 2 * (long long)L + l, tl);
        AttnVec<VEC>::load(p.log_prior + l, lp);
        float pi[VEC], a[VEC], f[VEC];
        bool dead[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int r = (l + e) / P, t = (l + e) - r * P;
            const float q = lg[e] - lse_q;
            pi[e] = __expf(q);
            a[e] = __expf(lg[e] + g[e] - lse_a);
            float gx, gy;
            grid_xy(t, p.d, p.s, gx, gy);
            acc[0] += a[e] * gx;
            acc[1] += a[e] * gy;
            dead[e] = (pi[e] == 0.f);  // train_mnist.py:246-254 guards
            const float th_std = __expf(tl[e]) + kEpsStd;
            acc[2] += a[e] * tm[e];
            acc[3] += a[e] * th_std;
            f[e] = q - lp[e];
            if (!dead[e]) {
                const float dm = tm[e] - p.offsets[r];
                f[e] += log_sp - __logf(th_std) + (th_std * th_std + dm * dm) * inv2s2 - 0.5f;
            }
        }
#pragma unroll
        for (int k = 0; k < Z; ++k) {
            float zm[VEC], zl[VEC];
            AttnVec<VEC>::load(hb + (long long)(3 + k) * L + l, zm);
            AttnVec<VEC>::load(hb + (long long)(3 + Z + k) * L + l, zl);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float zs = __expf(zl[e]) + kEpsStd;
                acc[6 + k] += a[e] * zm[e];
                acc[6 + Z + k] += a[e] * zs;
                if (!dead[e]) f[e] += -__logf(zs) + 0.5f * (zs * zs + zm[e] * zm[e]) - 0.5f;
            }
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[4] += pi[e] * f[e];
    }
    block_reduce<NA, false>(acc, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) s_part[i] = acc[i];
    }
    cluster.sync();
    if (rank == 0 && threadIdx.x == 0) {
        float tot[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) tot[i] = 0.f;
        for (int r = 0; r < CL; ++r) {
            const float* rp = cluster.map_shared_rank(s_part, r);
#pragma unroll
            for (int i = 0; i < NA; ++i) tot[i] += rp[i];
        }
        p.stats[b * 4 + 0] = mq;
        p.stats[b * 4 + 1] = lse_q;
        p.stats[b * 4 + 2] = ma;
        p.stats[b * 4 + 3] = lse_a;
        p.dx[b * 2 + 0] = tot[0];
        p.dx[b * 2 + 1] = tot[1];
        p.theta_b[b] = tot[3] * p.r_theta[b] + tot[2];
        p.kl[b] = tot[4];
        for (int k = 0; k < Z; ++k) p.zb[b * Z + k] = tot[6 + Z + k] * p.r_z[b * Z + k] + tot[6 + k];
    }
    cluster.sync();                                           // peers keep their shared memory alive until rank 0 has read it
}

// elementwise over (image b = blockIdx.y, cells): one read of the maps, one write of their gradients
template <int Z, int VEC>
__global__ void __launch_bounds__(256, 3) attn_bwd_kernel(AttnBwdParams p) {
    const int P = p.d * p.d, L = p.G * P, NH = 3 + 2 * Z;
    const int b = blockIdx.y;
    const float* hb = p.heads + (long long)b * NH * L;
    const float* gb = p.gumbel + (long long)b * L;
    float* db = p.d_heads + (long long)b * NH * L;
    const float lse_q = p.stats[b * 4 + 1], lse_a = p.stats[b * 4 + 3];
    const float K = p.kl[b];
    const float g_kl = __ldg(p.g_kl);
    const float g_th = p.g_theta[b], r_th = p.r_theta[b];
    const float gdx = p.g_dx[b * 2], gdy = p.g_dx[b * 2 + 1];
    float gz[Z], rz[Z];
    float s_ac = g_th * p.theta_b[b] + gdx * p.dx[b * 2] + gdy * p.dx[b * 2 + 1];
#pragma unroll
    for (int k = 0; k < Z; ++k) {
        gz[k] = p.g_zb[b * Z + k];
        rz[k] = p.r_z[b * Z + k];
        s_ac += gz[k] * p.zb[b * Z + k];
    }
    const float inv_s2 = 1.f / (p.theta_prior_std * p.theta_prior_std);
    const float log_sp = logf(p.theta_prior_std);
    for (int l = (blockIdx.x * blockDim.x + threadIdx.x) * VEC; l + VEC <= L; l += gridDim.x * blockDim.x * VEC) {
        float lg[VEC], gm[VEC], tm[VEC], tl[VEC], lp[VEC];
        AttnVec<VEC>::load(hb + l, lg);
        AttnVec<VEC>::load(gb + l, gm);
        AttnVec<VEC>::load(hb + L + l, tm);
        AttnVec<VEC>::load(hb + 2 * (long long)L + l, tl);
        AttnVec<VEC>::load(p.log_prior + l, lp);
        float q[VEC], pi[VEC], a[VEC], c[VEC], f[VEC], o_tm[VEC], o_tl[VEC];
        bool dead[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int r = (l + e) / P, t = (l + e) - r * P;
            q[e] = lg[e] - lse_q;
            pi[e] = __expf(q[e]);
            a[e] = __expf(lg[e] + gm[e] - lse_a);
            dead[e] = (pi[e] == 0.f);
            float gx, gy;
            grid_xy(t, p.d, p.s, gx, gy);
            const float e_th = __expf(tl[e]), th_std = e_th + kEpsStd;
            c[e] = g_th * (th_std * r_th + tm[e]) + gdx * gx + gdy * gy;
            f[e] = q[e] - lp[e];
            o_tm[e] = g_th * a[e];
            o_tl[e] = g_th * r_th * a[e] * e_th;
            if (!dead[e]) {
                const float dm = tm[e] - p.offsets[r];
                f[e] += log_sp - __logf(th_std) + 0.5f * (th_std * th_std + dm * dm) * inv_s2 - 0.5f;
                o_tm[e] += g_kl * pi[e] * dm * inv_s2;
                o_tl[e] += g_kl * pi[e] * (-1.f / th_std + th_std * inv_s2) * e_th;
            }
        }
#pragma unroll
        for (int k = 0; k < Z; ++k) {
            float zm[VEC], zl[VEC], o_zm[VEC], o_zl[VEC];
            AttnVec<VEC>::load(hb + (long long)(3 + k) * L + l, zm);
            AttnVec<VEC>::load(hb + (long long)(3 + Z + k) * L + l, zl);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float e_z = __expf(zl[e]), zs = e_z + kEpsStd;
                c[e] += gz[k] * (zs * rz[k] + zm[e]);
                o_zm[e] = gz[k] * a[e];
                o_zl[e] = gz[k] * rz[k] * a[e] * e_z;
                if (!dead[e]) {
                    f[e] += -__logf(zs) + 0.5f * (zs * zs + zm[e] * zm[e]) - 0.5f;
                    o_zm[e] += g_kl * pi[e] * zm[e];
                    o_zl[e] += g_kl * pi[e] * (-1.f / zs + zs) * e_z;
                }
            }
            AttnVec<VEC>::store(db + (long long)(3 + k) * L + l, o_zm);
            AttnVec<VEC>::store(db + (long long)(3 + Z + k) * L + l, o_zl);
        }
        AttnVec<VEC>::store(db + L + l, o_tm);
        AttnVec<VEC>::store(db + 2 * (long long)L + l, o_tl);
        float o_a[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) o_a[e] = a[e] * (c[e] - s_ac) + g_kl * pi[e] * (f[e] - K);
        AttnVec<VEC>::store(db + l, o_a);
    }
}

}  // namespace tvae
