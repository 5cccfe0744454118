// CUDA-core (HBM-bound) kernels of the hot path: rotated filter bank fwd/bwd, attention inference
// (softmax / Gumbel-softmax / expectations / KL) fwd/bwd, likelihoods, thin-layer backward, small helpers.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "ptx.cuh"
#include "linear_policies.cuh"   // act_apply / act_grad_from_out

namespace tvae {

constexpr int kMaxG = 16;
constexpr int kMaxHeads = 3 + 2 * 16;  // attn, theta_mu, theta_logstd, 2*z  (z <= 16)
constexpr float kEpsStd = 1e-6f;       // train_mnist.py:197
constexpr float kSlope = 0.01f;

struct RotTable {
    float cs[kMaxG];
    float sn[kMaxG];
};

// ------------------------------------------------------------------------------------------
// a-1  rotated filter bank (models.py:174-197), closed form of affine_grid + grid_sample
//      (bilinear, zeros padding, align_corners=False).
//   weight (O,C,k,k)  ->  bank [G*O][kpad16] fp16, row n' = r*O + o, column kk = (c*k + v)*k + u;
//   columns kk >= C*k*k are zero.  (fp16 keeps the 11-bit significand of a TF32 operand.)
// ------------------------------------------------------------------------------------------
struct BilinearTap {
    int x0, y0;
    float wx0, wx1, wy0, wy1;
};
__device__ __forceinline__ BilinearTap rot_tap(int u, int v, int k, float cs, float sn) {
    const float c0 = 0.5f * (k - 1);
    const float cx = u - c0, cy = v - c0;
    const float ix = cs * cx + sn * cy + c0;
    const float iy = -sn * cx + cs * cy + c0;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    BilinearTap t;
    t.x0 = static_cast<int>(fx0);
    t.y0 = static_cast<int>(fy0);
    t.wx1 = ix - fx0; t.wx0 = 1.f - t.wx1;
    t.wy1 = iy - fy0; t.wy0 = 1.f - t.wy1;
    return t;
}

__global__ void filter_bank_fwd_kernel(const float* __restrict__ w, __half* __restrict__ bank, int O, int C, int k, int G,
                                       int kpad, RotTable rot) {
    const int K = C * k * k;
    const long long total = (long long)G * O * kpad;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int kk = static_cast<int>(idx % kpad);
        const int np = static_cast<int>(idx / kpad);
        float val = 0.f;
        if (kk < K) {
            const int r = np / O, o = np - r * O;
            const int u = kk % k, v = (kk / k) % k, c = kk / (k * k);
            const BilinearTap t = rot_tap(u, v, k, rot.cs[r], rot.sn[r]);
            const float* wp = w + ((long long)o * C + c) * k * k;
            const bool x0ok = t.x0 >= 0 && t.x0 < k, x1ok = t.x0 + 1 >= 0 && t.x0 + 1 < k;
            const bool y0ok = t.y0 >= 0 && t.y0 < k, y1ok = t.y0 + 1 >= 0 && t.y0 + 1 < k;
            if (y0ok && x0ok) val += wp[t.y0 * k + t.x0] * (t.wy0 * t.wx0);
            if (y0ok && x1ok) val += wp[t.y0 * k + t.x0 + 1] * (t.wy0 * t.wx1);
            if (y1ok && x0ok) val += wp[(t.y0 + 1) * k + t.x0] * (t.wy1 * t.wx0);
            if (y1ok && x1ok) val += wp[(t.y0 + 1) * k + t.x0 + 1] * (t.wy1 * t.wx1);
        }
        bank[idx] = __float2half_rn(val);
    }
}

// adjoint: dweight (O,C,k,k) += bilinear-scatter of dbank [G*O][ld] (first C*k*k columns).
__global__ void filter_bank_bwd_kernel(const float* __restrict__ dbank, long long ld, float* __restrict__ dw, int O, int C,
                                       int k, int G, RotTable rot) {
    const int K = C * k * k;
    const long long total = (long long)G * O * K;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int kk = static_cast<int>(idx % K);
        const int np = static_cast<int>(idx / K);
        const int r = np / O, o = np - r * O;
        const int u = kk % k, v = (kk / k) % k, c = kk / (k * k);
        const float g = dbank[(long long)np * ld + kk];
        const BilinearTap t = rot_tap(u, v, k, rot.cs[r], rot.sn[r]);
        float* wp = dw + ((long long)o * C + c) * k * k;
        const bool x0ok = t.x0 >= 0 && t.x0 < k, x1ok = t.x0 + 1 >= 0 && t.x0 + 1 < k;
        const bool y0ok = t.y0 >= 0 && t.y0 < k, y1ok = t.y0 + 1 >= 0 && t.y0 + 1 < k;
        if (y0ok && x0ok) atomicAdd(wp + t.y0 * k + t.x0, g * (t.wy0 * t.wx0));
        if (y0ok && x1ok) atomicAdd(wp + t.y0 * k + t.x0 + 1, g * (t.wy0 * t.wx1));
        if (y1ok && x0ok) atomicAdd(wp + (t.y0 + 1) * k + t.x0, g * (t.wy1 * t.wx0));
        if (y1ok && x1ok) atomicAdd(wp + (t.y0 + 1) * k + t.x0 + 1, g * (t.wy1 * t.wx1));
    }
}

// Plane variants of the two kernels above (k * k floats fit shared memory).
// Forward, grid (O * C planes, G rotations): the k x k plane is read once into shared memory and sampled from there (the
// grid-stride form paid 64-bit index arithmetic and four scattered global loads per output).
// dynamic smem: k * k floats.
__global__ void __launch_bounds__(256) filter_bank_fwd_plane_kernel(const float* __restrict__ w, __half* __restrict__ bank, int O, int C,
                                                                    int k, int kpad, RotTable rot) {
    extern __shared__ float s_plane[];
    const int oc = blockIdx.x, o = oc / C, c = oc - o * C, kk2 = k * k, r = blockIdx.y;
    const float* wp = w + (long long)oc * kk2;
    for (int i = threadIdx.x; i < kk2; i += blockDim.x) s_plane[i] = __ldg(wp + i);
    __syncthreads();
    __half* out = bank + (long long)(r * O + o) * kpad + c * kk2;
    const float cs = rot.cs[r], sn = rot.sn[r];
    for (int kk = threadIdx.x; kk < kk2; kk += blockDim.x) {
        const int v = kk / k, u = kk - v * k;
        const BilinearTap t = rot_tap(u, v, k, cs, sn);
        const bool x0ok = t.x0 >= 0 && t.x0 < k, x1ok = t.x0 + 1 >= 0 && t.x0 + 1 < k;
        const bool y0ok = t.y0 >= 0 && t.y0 < k, y1ok = t.y0 + 1 >= 0 && t.y0 + 1 < k;
        float val = 0.f;
        if (y0ok && x0ok) val += s_plane[t.y0 * k + t.x0] * (t.wy0 * t.wx0);
        if (y0ok && x1ok) val += s_plane[t.y0 * k + t.x0 + 1] * (t.wy0 * t.wx1);
        if (y1ok && x0ok) val += s_plane[(t.y0 + 1) * k + t.x0] * (t.wy1 * t.wx0);
        if (y1ok && x1ok) val += s_plane[(t.y0 + 1) * k + t.x0 + 1] * (t.wy1 * t.wx1);
        out[kk] = __float2half_rn(val);
    }
    if (c == C - 1)                                        // the row's padding columns [C k^2, kpad)
        for (int kk = C * kk2 + threadIdx.x; kk < kpad; kk += blockDim.x) bank[(long long)(r * O + o) * kpad + kk] = __float2half_rn(0.f);
}
// Backward as a GATHER, thread = one weight pixel (x, y) of plane (o, c): for every rotation the output taps (u, v) whose
// bilinear footprint covers (x, y) lie within sqrt(2) of the inversely rotated pixel - at most 4 x 4 candidates, each
// contributing dbank * max(0, 1 - |ix - x|) * max(0, 1 - |iy - y|), which is exactly the forward's corner weight.  No atomics
// (the scatter form issued 16.8 M of them at cfg2: 73 us), no zero-fill of dweight, bit-deterministic.
// grid (O * C, ceil(k * k / 256)).
__global__ void __launch_bounds__(256) filter_bank_bwd_gather_kernel(const float* __restrict__ dbank, long long ld, float* __restrict__ dw,
                                                                     int O, int C, int k, int G, RotTable rot) {
    const int oc = blockIdx.x, o = oc / C, c = oc - o * C, kk2 = k * k;
    const int px = blockIdx.y * blockDim.x + threadIdx.x;
    if (px >= kk2) return;
    const int y = px / k, x = px - y * k;
    const float c0 = 0.5f * (k - 1);
    const float fx = x - c0, fy = y - c0;
    float acc = 0.f;
    for (int r = 0; r < G; ++r) {
        const float cs = rot.cs[r], sn = rot.sn[r];
        const float* gp = dbank + (long long)(r * O + o) * ld + c * kk2;
        // inverse of (ix, iy) = (cs cx + sn cy, -sn cx + cs cy) + c0
        const float cu = cs * fx - sn * fy + c0, cv = sn * fx + cs * fy + c0;
        const int u_lo = max(0, static_cast<int>(ceilf(cu - 1.5f))), u_hi = min(k - 1, static_cast<int>(floorf(cu + 1.5f)));
        const int v_lo = max(0, static_cast<int>(ceilf(cv - 1.5f))), v_hi = min(k - 1, static_cast<int>(floorf(cv + 1.5f)));
        for (int v = v_lo; v <= v_hi; ++v) {
            const float cy = v - c0;
            for (int u = u_lo; u <= u_hi; ++u) {
                const float cx = u - c0;
                const float ix = cs * cx + sn * cy + c0;            // the forward's expressions (rot_tap), term by term
                const float iy = -sn * cx + cs * cy + c0;
                const float wx = 1.f - fabsf(ix - x), wy = 1.f - fabsf(iy - y);
                if (wx > 0.f && wy > 0.f) acc = fmaf(__ldg(gp + v * k + u), wx * wy, acc);
            }
        }
    }
    dw[(long long)oc * kk2 + px] = acc;
}

// ------------------------------------------------------------------------------------------
// block-wide reductions (blockDim.x multiple of 32, <= 1024)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// reduces NV values per thread; result valid in all threads. scratch: NV * 32 floats.
template <int NV, bool IS_MAX>
__device__ __forceinline__ void block_reduce(float (&v)[NV], float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = IS_MAX ? warp_max(v[i]) : warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[i * 32 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float x = lane < nw ? scratch[i * 32 + lane] : (IS_MAX ? -CUDART_INF_F : 0.f);
        v[i] = IS_MAX ? warp_max(x) : warp_sum(x);
    }
}

// ------------------------------------------------------------------------------------------
// a-3 tail + a-4: attention inference over the head maps of one image.
//   heads (B, NH, G, P) planar, channel order [attn(+p_r), theta_mu(+offset), theta_logstd,
//   z_mu[0..z), z_logstd[0..z)], P = H'*W'.  One CTA per image, two passes (second hits L2).
//   Outputs per image (fp32): stats[b] = {max_q, lse_q, max_a, lse_a}, zb (z), theta_b, dx (2), kl_b.
// ------------------------------------------------------------------------------------------
struct AttnParams {
    const float* heads;     // (B, NH, G, P)
    const float* gumbel;    // (B, G*P)
    const float* r_z;       // (B, z)
    const float* r_theta;   // (B)
    const float* log_prior; // (G*P) : log_softmax over (r,t) of p_t + p_r  (train_mnist.py:258-262)
    float* stats;           // (B, 4)
    float* zb;              // (B, z)
    float* theta_b;         // (B)
    float* dx;              // (B, 2)
    float* kl;              // (B)
    int B, G, d, z;         // d = H' = W'
    float s;                // pixel spacing x[1,0]-x[0,0]
    float theta_prior_std;  // pi / G (train_mnist.py:269-272)
    float offsets[kMaxG];
};

__device__ __forceinline__ void grid_xy(int t, int d, float s, float& gx, float& gy) {
    // train_mnist.py:209-217: cell (i,j) -> ((j - d/2) s, (d-1-i - d/2) s)
    const int i = t / d, j = t - i * d;
    gx = (j - d / 2) * s;
    gy = (d - 1 - i - d / 2) * s;
}

struct AttnBwdParams {
    const float* heads; const float* gumbel; const float* r_z; const float* r_theta; const float* log_prior;
    const float* stats; const float* zb; const float* theta_b; const float* dx; const float* kl;
    const float* g_zb;     // (B, z)   dLoss/dz_b
    const float* g_theta;  // (B)
    const float* g_dx;     // (B, 2)
    const float* g_kl;     // device scalar dLoss/dkl_b (same for every image)
    float* d_heads;        // (B, NH, G, P)
    int B, G, d, z;
    float s, theta_prior_std;
    float offsets[kMaxG];
};

// log_prior[l] = log_softmax_{(r,t)}( sum_xy N(grid; 0, 0.1).log_prob + p_r[r] )   (one small CTA)
__global__ void log_prior_kernel(float* __restrict__ out, int G, int d, float s, RotTable p_r_in_cs) {
    __shared__ float scratch[32];
    const int P = d * d, L = G * P;
    const float sig = 0.1f;
    const float c0 = -logf(sig) - 0.5f * logf(2.f * CUDART_PI_F);
    float mx[1] = {-CUDART_INF_F};
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const int r = l / P, t = l - r * P;
        float gx, gy;
        grid_xy(t, d, s, gx, gy);
        const float v = (-(gx * gx) / (2.f * sig * sig) + c0) + (-(gy * gy) / (2.f * sig * sig) + c0) + p_r_in_cs.cs[r];
        out[l] = v;
        mx[0] = fmaxf(mx[0], v);
    }
    block_reduce<1, true>(mx, scratch);
    float se[1] = {0.f};
    for (int l = threadIdx.x; l < L; l += blockDim.x) se[0] += expf(out[l] - mx[0]);
    block_reduce<1, false>(se, scratch);
    const float lse = mx[0] + logf(se[0]);
    for (int l = threadIdx.x; l < L; l += blockDim.x) out[l] -= lse;
}

// ------------------------------------------------------------------------------------------
// module-interface tail (models.py:382-388): q_t_r = log_softmax(attn), a_sampled = softmax(attn + g)
// from the stats of attn_fwd-style reductions.  One CTA per image.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) softmax_pair_kernel(const float* __restrict__ heads, const float* __restrict__ gumbel,
                                                            float* __restrict__ q_out, float* __restrict__ a_out, int NH, int L) {
    __shared__ float scratch[64];
    const int b = blockIdx.x;
    const float* hb = heads + (long long)b * NH * L;
    const float* gb = gumbel + (long long)b * L;
    float mx[2] = {-CUDART_INF_F, -CUDART_INF_F};
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        mx[0] = fmaxf(mx[0], hb[l]);
        mx[1] = fmaxf(mx[1], hb[l] + gb[l]);
    }
    block_reduce<2, true>(mx, scratch);
    float se[2] = {0.f, 0.f};
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        se[0] += __expf(hb[l] - mx[0]);
        se[1] += __expf(hb[l] + gb[l] - mx[1]);
    }
    block_reduce<2, false>(se, scratch);
    const float lse_q = mx[0] + logf(se[0]), lse_a = mx[1] + logf(se[1]);
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        q_out[(long long)b * L + l] = hb[l] - lse_q;
        a_out[(long long)b * L + l] = expf(hb[l] + gb[l] - lse_a);
    }
}

// ------------------------------------------------------------------------------------------
// a-10 get_latent (clustering_mnist.py:122-161): argmax over (r,t) of attn, z / theta at the argmax,
// softmax-expected translation.  One CTA per image.
// ------------------------------------------------------------------------------------------
template <int Z>
__global__ void __launch_bounds__(1024) get_latent_kernel(const float* __restrict__ heads, int G, int d, float s,
                                                          float* __restrict__ z_content, float* __restrict__ theta_mu,
                                                          float* __restrict__ dx, int* __restrict__ argmax_out) {
    __shared__ float scratch[64];
    __shared__ int sidx[32];
    const int b = blockIdx.x;
    const int P = d * d, L = G * P, NH = 3 + 2 * Z;
    const float* hb = heads + (long long)b * NH * L;
    float best = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const float v = hb[l];
        if (v > best) { best = v; bi = l; }  // ascending l per thread: first maximum wins
    }
    // first-index argmax across the block (torch.max returns the first maximal index)
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { scratch[warp] = best; sidx[warp] = bi; }
    __syncthreads();
    best = lane < nw ? scratch[lane] : -CUDART_INF_F;
    bi = lane < nw ? sidx[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __syncthreads();
    float se[1] = {0.f};
    for (int l = threadIdx.x; l < L; l += blockDim.x) se[0] += __expf(hb[l] - best);
    block_reduce<1, false>(se, scratch);
    float acc[2] = {0.f, 0.f};
    const float inv = 1.f / se[0];
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const int t = l % P;
        float gx, gy;
        grid_xy(t, d, s, gx, gy);
        const float pi = __expf(hb[l] - best) * inv;
        acc[0] += pi * gx;
        acc[1] += pi * gy;
    }
    block_reduce<2, false>(acc, scratch);
    if (threadIdx.x == 0) {
        argmax_out[b] = bi;
        dx[b * 2] = acc[0];
        dx[b * 2 + 1] = acc[1];
        theta_mu[b] = hb[L + bi];
        for (int k = 0; k < Z; ++k) {
            z_content[b * 2 * Z + k] = hb[(3 + k) * L + bi];
            z_content[b * 2 * Z + Z + k] = expf(hb[(3 + Z + k) * L + bi]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// a-7 Bernoulli likelihood (train_mnist.py:288-291) fwd + dLoss/dy_hat in one pass.
//   ll[b] = -sum_e (softplus(yh) - y*yh);  d_yhat = g * (y - sigmoid(yh)), g = dLoss/d(ll_b) (device scalar)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bernoulli_kernel(const float* __restrict__ yhat, const float* __restrict__ y,
                                                        float* __restrict__ ll, float* __restrict__ d_yhat, int E, const float* __restrict__ g_ptr) {
    __shared__ float scratch[32];
    const int b = blockIdx.y;
    const float g = d_yhat ? __ldg(g_ptr) : 0.f;
    const float* yh = yhat + (long long)b * E;
    const float* yy = y + (long long)b * E;
    float acc[1] = {0.f};
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const float x = yh[e], t = yy[e];
        // BCE-with-logits: max(x,0) - x*t + log1p(exp(-|x|))
        acc[0] += fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        if (d_yhat) d_yhat[(long long)b * E + e] = g * (t - 1.f / (1.f + expf(-x)));
    }
    block_reduce<1, false>(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(ll + b, -acc[0]);
}

// ------------------------------------------------------------------------------------------
// a-8 CTF point-spread correlation (train_particles.py:298-302): per-image (n-1)x(n-1) kernel,
// zero padding (n-1)/2, cross-correlation.  out[b,i,j] = sum_{v,u} in[b, i+v-pad, j+u-pad] * ctf[b,v,u].
// TRANSPOSED = adjoint (gradient w.r.t. in).  One CTA per (image, 16x16 output tile); the
// needed input window and the filter stream through shared memory in row slabs.
// ------------------------------------------------------------------------------------------
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256) ctf_apply_kernel(const float* __restrict__ in, const float* __restrict__ ctf,
                                                        float* __restrict__ out, int n, int mf, int m) {
    // mf: filter size as stored, pad = mf / 2 (train_particles.py:301 `padding=ctf.size(2)//2`); m <= mf: the centred
    // m x m part of it that can meet the image (taps further than n - 1 from the centre only ever multiply padding)
    extern __shared__ float sm[];
    const int pad = m / 2, off = (mf - m) / 2;
    const int b = blockIdx.z;
    const int ti = blockIdx.y * 16, tj = blockIdx.x * 16;
    const int ly = threadIdx.x / 16, lx = threadIdx.x % 16;
    const int W = 16 + m - 1;              // window width/height
    float* win = sm;                       // W*W input window (zero padded)
    float* flt = sm + W * W;               // m*m filter
    const float* inb = in + (long long)b * n * n;
    const float* fb = ctf + (long long)b * mf * mf;
    // forward:  out[i,j] = sum in[i+v-pad, j+u-pad] f[v,u]       window origin (ti-pad, tj-pad)
    // adjoint:  out[i,j] = sum in[i-v+pad, j-u+pad] f[v,u]       window origin (ti+pad-(m-1), tj+pad-(m-1)), f flipped
    const int oy = TRANSPOSED ? ti + pad - (m - 1) : ti - pad;
    const int ox = TRANSPOSED ? tj + pad - (m - 1) : tj - pad;
    for (int idx = threadIdx.x; idx < W * W; idx += blockDim.x) {
        const int yy = oy + idx / W, xx = ox + idx % W;
        win[idx] = (yy >= 0 && yy < n && xx >= 0 && xx < n) ? inb[yy * n + xx] : 0.f;
    }
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
        const int v = idx / m, u = idx % m;
        flt[idx] = TRANSPOSED ? fb[(off + m - 1 - v) * mf + (off + m - 1 - u)] : fb[(off + v) * mf + off + u];
    }
    __syncthreads();
    float acc = 0.f;
    for (int v = 0; v < m; ++v) {
        const float* wr = win + (ly + v) * W + lx;
        const float* fr = flt + v * m;
#pragma unroll 4
        for (int u = 0; u < m; ++u) acc = fmaf(wr[u], fr[u], acc);
    }
    const int i = ti + ly, j = tj + lx;
    if (i < n && j < n) out[(long long)b * n * n + i * n + j] = acc;
}

// Gaussian likelihood with optional mask (train_particles.py:326-338, no learned variance):
//   ll[b] = -0.5 sum_px mask*(mu - y)^2 ; d_mu = g * mask * (y - mu), g = dLoss/d(ll_b) (device scalar)
// mask: pixels within `radius` of dx/s on the integer grid x in [-n/2, n/2), y in (-n/2, n/2]  (:309-324)
__global__ void __launch_bounds__(256) gaussian_kernel(const float* __restrict__ mu, const float* __restrict__ y,
                                                       const float* __restrict__ dx, float s, int n, int radius,
                                                       float* __restrict__ ll, float* __restrict__ d_mu, const float* __restrict__ g_ptr) {
    __shared__ float scratch[32];
    const int b = blockIdx.y, E = n * n;
    const float g = d_mu ? __ldg(g_ptr) : 0.f;
    float cx = 0.f, cy = 0.f;
    if (radius > 0) { cx = dx[b * 2] / s; cy = dx[b * 2 + 1] / s; }
    float acc[1] = {0.f};
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        bool keep = true;
        if (radius > 0) {
            const int i = e / n, j = e - i * n;
            // np.arange(-n//2, n//2) / np.arange(n//2, -n//2, -1) with python floor division
            const int half_lo = -((n + 1) / 2);      // -n//2 for positive n  (floor)
            const float gx = static_cast<float>(half_lo + j);
            const float gy = static_cast<float>(n / 2 - i);
            const double ddx = (double)cx - gx, ddy = (double)cy - gy;
            keep = sqrt(ddx * ddx + ddy * ddy) < (double)radius;
        }
        const float diff = keep ? mu[(long long)b * E + e] - y[(long long)b * E + e] : 0.f;
        acc[0] += diff * diff;
        if (d_mu) d_mu[(long long)b * E + e] = -g * diff;
    }
    block_reduce<1, false>(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(ll + b, -0.5f * acc[0]);
}

// Gaussian likelihood with a learned per-pixel variance (--fit-noise, train_particles.py:289-296, 333-334; no CTF and no
// mask: with either, the reference's own tensor shapes no longer line up for B > 1).  The generator emits (B, N, 2); the
// reference flattens that to (B, 2N) and takes the FIRST N values as the mean and the LAST N as the log-variance, so
//   mu_e = f[b][e], lv_e = f[b][N + e]   (not the per-pixel pairs) - reproduced as is:
//   ll[b] = -0.5 sum_e ((mu_e - y_e)^2 exp(-lv_e) + lv_e)
//   d f[b][e] = -g (mu_e - y_e) exp(-lv_e),  d f[b][N + e] = 0.5 g ((mu_e - y_e)^2 exp(-lv_e) - 1),  g = dLoss/d(ll_b)
// One CTA per image: deterministic ll[b].
__global__ void __launch_bounds__(256) gaussian_fit_noise_kernel(const float* __restrict__ f, const float* __restrict__ y, int N,
                                                                 float* __restrict__ ll, float* __restrict__ d_f,
                                                                 const float* __restrict__ g_ptr) {
    __shared__ float scratch[32];
    const int b = blockIdx.x;
    const float g = d_f ? __ldg(g_ptr) : 0.f;
    const float* fb = f + (long long)b * 2 * N;
    const float* yb = y + (long long)b * N;
    float acc[1] = {0.f};
    for (int e = threadIdx.x; e < N; e += blockDim.x) {
        const float lv = fb[N + e];
        const float diff = fb[e] - yb[e];
        const float w = diff * expf(-lv);       // (mu - y) / var
        acc[0] += fmaf(diff, w, lv);
        if (d_f) {
            d_f[(long long)b * 2 * N + e] = -g * w;
            d_f[(long long)b * 2 * N + N + e] = 0.5f * g * (diff * w - 1.f);
        }
    }
    block_reduce<1, false>(acc, scratch);
    if (threadIdx.x == 0) ll[b] = -0.5f * acc[0];
}

// ------------------------------------------------------------------------------------------
// Thin-layer backward.  A "thin" layer maps a wide activation a[m][W] to T outputs t[m][j] =
// sum_c a[m][c] Wt[j][c] + bt[j]  (attention/theta/z heads: T = 3+2z; generator output layer: T = n_out).
// Given dt it produces, in one streaming pass over a (HBM-bound: one read of a, one write of dpre):
//   dpre[m][c] = (sum_j dt[m][j] Wt[j][c]) * lrelu'(a[m][c])      (gradient w.r.t. the pre-activation of a)
//   dWt[j][c] += sum_m dt[m][j] a[m][c],  dbt[j] += sum_m dt[m][j],  dcol[c] += sum_m dpre[m][c]
// Thread = VEC adjacent columns x one row slot; a CTA streams a contiguous chunk of rows in blocks of
// kThinRB rows whose dt values are staged in shared memory.  dt is addressed as
//   dt[b * dt_outer + j * dt_chan + r * P + pos],  m = (b*G + r)*P + pos   (planar head maps), or row-major dt[m][T].
// Per-CTA partial sums are combined in shared memory, then one global atomicAdd per output element.
// ------------------------------------------------------------------------------------------
struct ThinBwdParams {
    const void* a;       // [M][W] post-activation (fp32, or fp16 when H16)
    const float* dt;
    const float* Wt;     // [T][W]
    void* dpre;          // [M][W] (fp32, or fp16 = value * *store_scale when H16)
    const float* store_scale;
    float* dWt;          // [T][W]
    float* dbt;          // [T]
    float* dcol;         // [W]
    long long M;
    int W, T, P;
    long long dt_outer, dt_chan;
    int rows_per_cta;
    int act;             // activation that produced `a`: kActTanh or LeakyReLU (linear_policies.cuh)
    const float* dt_scale;   // thin_bwd_mma only: device scalar (power of two) applied to dt before its fp16 rounding and divided
                             // out of dpre / dWt again (dt spans more than fp16's exponent range), or null
};
constexpr int kThinRB = 64;

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };
template <int VEC>
__device__ __forceinline__ void vec_load(const float* p, float (&v)[VEC]) {
    using V = typename VecT<VEC>::type;
    const V t = *reinterpret_cast<const V*>(p);
    if constexpr (VEC == 4) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (VEC == 2) { v[0] = t.x; v[1] = t.y; }
    else v[0] = t;
}
template <int VEC>
__device__ __forceinline__ void vec_store(float* p, const float (&v)[VEC]) {
    using V = typename VecT<VEC>::type;
    V t;
    if constexpr (VEC == 4) { t.x = v[0]; t.y = v[1]; t.z = v[2]; t.w = v[3]; }
    else if constexpr (VEC == 2) { t.x = v[0]; t.y = v[1]; }
    else t = v[0];
    *reinterpret_cast<V*>(p) = t;
}

// 16-bit variants: VEC fp16 values <-> floats
template <int VEC>
__device__ __forceinline__ void vec_load_h(const __half* p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else if constexpr (VEC == 2) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(p));
        v[0] = a.x; v[1] = a.y;
    } else {
        v[0] = __half2float(*p);
    }
}
template <int VEC>
__device__ __forceinline__ void vec_store_h(__half* p, const float (&v)[VEC], float scale) {
    if constexpr (VEC == 4) {
        const __half2 a = __floats2half2_rn(v[0] * scale, v[1] * scale), b = __floats2half2_rn(v[2] * scale, v[3] * scale);
        uint2 t;
        t.x = *reinterpret_cast<const uint32_t*>(&a); t.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v[0] * scale, v[1] * scale);
    } else {
        *p = __float2half_rn(v[0] * scale);
    }
}

// dynamic smem: [kThinRB * T] staged dt, then [(T + 1) * W + T] CTA partial sums (dWt, dcol, dbt)
template <int TMAX, int VEC, bool PLANAR, bool H16>
__global__ void __launch_bounds__(256) thin_bwd_kernel(ThinBwdParams p, int G) {
    extern __shared__ float s_thin[];
    float* s_dt = s_thin;
    float* s_red = s_thin + kThinRB * p.T;
    const int cgs = p.W / VEC;                 // column groups
    const int rpp = blockDim.x / cgs;          // rows per pass
    const int cg = threadIdx.x % cgs, rs = threadIdx.x / cgs;
    const int c0 = cg * VEC;
    const int n_red = (p.T + 1) * p.W + p.T;
    for (int i = threadIdx.x; i < n_red; i += blockDim.x) s_red[i] = 0.f;
    const long long m_begin = (long long)blockIdx.x * p.rows_per_cta;
    const long long m_end = min(m_begin + p.rows_per_cta, p.M);
    float wt[TMAX][VEC], dw[TMAX][VEC], dcol[VEC];
#pragma unroll
    for (int j = 0; j < TMAX; ++j) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            wt[j][v] = j < p.T ? p.Wt[(long long)j * p.W + c0 + v] : 0.f;
            dw[j][v] = 0.f;
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) dcol[v] = 0.f;
    float dbt = 0.f;   // threads with tid < T accumulate column tid of the staged dt block
    const float store_scale = (H16 && p.store_scale) ? __ldg(p.store_scale) : 1.f;
    for (long long m0 = m_begin; m0 < m_end; m0 += kThinRB) {
        const int rows = static_cast<int>(min((long long)kThinRB, m_end - m0));
        __syncthreads();
        for (int idx = threadIdx.x; idx < rows * p.T; idx += blockDim.x) {
            long long addr;
            int j, rr;
            if (PLANAR) {
                j = idx / rows; rr = idx - j * rows;   // consecutive threads -> consecutive rows (coalesced planar read)
                const long long m = m0 + rr;
                const long long br = m / p.P;
                const int pos = static_cast<int>(m - br * p.P);
                const long long b = br / G;
                const int r = static_cast<int>(br - b * G);
                addr = b * p.dt_outer + (long long)j * p.dt_chan + (long long)r * p.P + pos;
            } else {
                rr = idx / p.T; j = idx - rr * p.T;
                addr = (m0 + rr) * p.T + j;
            }
            s_dt[rr * p.T + j] = p.dt[addr];
        }
        __syncthreads();
        if (rs < rpp) {
            for (int rr0 = rs; rr0 < rows; rr0 += 4 * rpp) {
                float av[4][VEC];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = rr0 + q * rpp;
                    if (rr < rows) {
                        if constexpr (H16) vec_load_h<VEC>(static_cast<const __half*>(p.a) + (m0 + rr) * p.W + c0, av[q]);
                        else vec_load<VEC>(static_cast<const float*>(p.a) + (m0 + rr) * p.W + c0, av[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = rr0 + q * rpp;
                    if (rr < rows) {
                        float g[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) g[v] = 0.f;
#pragma unroll
                        for (int j = 0; j < TMAX; ++j) {
                            if (j < p.T) {
                                const float d = s_dt[rr * p.T + j];
#pragma unroll
                                for (int v = 0; v < VEC; ++v) {
                                    g[v] = fmaf(d, wt[j][v], g[v]);
                                    dw[j][v] = fmaf(d, av[q][v], dw[j][v]);
                                }
                            }
                        }
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            g[v] *= act_grad_from_out(av[q][v], p.act);
                            dcol[v] += g[v];
                        }
                        if constexpr (H16) vec_store_h<VEC>(static_cast<__half*>(p.dpre) + (m0 + rr) * p.W + c0, g, store_scale);
                        else vec_store<VEC>(static_cast<float*>(p.dpre) + (m0 + rr) * p.W + c0, g);
                    }
                }
            }
        }
        if (p.dbt && threadIdx.x < p.T) {
            for (int rr = 0; rr < rows; ++rr) dbt += s_dt[rr * p.T + threadIdx.x];
        }
    }
    if (rs < rpp) {
#pragma unroll
        for (int j = 0; j < TMAX; ++j) {
            if (j < p.T) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) atomicAdd(s_red + j * p.W + c0 + v, dw[j][v]);
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) atomicAdd(s_red + p.T * p.W + c0 + v, dcol[v]);
    }
    if (p.dbt && threadIdx.x < p.T) s_red[(p.T + 1) * p.W + threadIdx.x] = dbt;
    __syncthreads();
    for (int i = threadIdx.x; i < p.T * p.W; i += blockDim.x) atomicAdd(p.dWt + i, s_red[i]);
    if (p.dcol)
        for (int i = threadIdx.x; i < p.W; i += blockDim.x) atomicAdd(p.dcol + i, s_red[p.T * p.W + i]);
    if (p.dbt && threadIdx.x < p.T) atomicAdd(p.dbt + threadIdx.x, s_red[(p.T + 1) * p.W + threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// Warp-MMA variant of thin_bwd_kernel for fp16 activations, T <= 16 thin outputs and W % 128 == 0.
// The CUDA-core kernel above spends 16 FMA + ~6 other instructions per element and is latency / issue bound
// (ncu: 46 % issue slots, 18 % DRAM); here both thin contractions are warp-level mma.sync.m16n8k16 (the operands are
// far too thin for tcgen05 tiles: K = T <= 16 for dpre, M = T <= 16 for dWt), which leaves a streaming kernel.
//   grid = (row chunks, W / 128); CTA = 8 warps; per 64-row block warp w owns rows (w & 3) * 16 .. +16 and the 64
//   columns (w >> 2) * 64 .. of the CTA's 128-column slice.
//   dpre  = S[16 rows x 16 t] . Wt[16 t x 8 c]     A fragment from the staged dt block, B fragments pre-packed in smem
//   dWt  += S^T[16 t x 16 rows] . a[16 rows x 8 c]  B fragments by ldmatrix.trans from the staged activation tile
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
constexpr int kThinPitch = 136;    // halves per staged activation row: 272 B keeps ldmatrix rows and fragment reads conflict-free

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
    const int sz = valid ? 16 : 0;     // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc, bool valid) {
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Three-buffer ring: the dt block and the activation tile of row blocks i+1 and i+2 are in flight (cp.async) while block i
// is computed (two CTAs per SM x two 18 KB blocks in flight cover the HBM latency-bandwidth product); dpre is written back into the activation tile in place (same thread, same element) and leaves the CTA as
// coalesced 16-byte stores.
// dynamic smem: [(T + 1) * 128 + T] partial sums | [16][32] uint2 Wt fragments | kThinBufs x { [64 * T] dt | a tile [64][136] halves }
constexpr int kThinBufs = 3;
// TT = 1: T <= 16.  TT = 2: 16 < T <= 24 (z = 8: 19 heads) - a second k-step for dpre and a second m-tile for dWt of
// which only rows 16..23 are kept (the accumulator rows 24..31 multiply zero rows of S^T).
template <bool PLANAR, int TT = 1>
__global__ void __launch_bounds__(256) thin_bwd_mma_kernel(ThinBwdParams p, int G) {
    extern __shared__ __align__(16) float s_thin[];
    const int T = p.T;
    float* s_red = s_thin;                                           // [(T + 1) * 128 + T]
    const int n_red = (T + 1) * 128 + T;
    uint2* s_wb = reinterpret_cast<uint2*>(s_red + ((n_red + 3) & ~3));   // [TT][16][32]
    const int dt_floats = (kThinRB * T + 3) & ~3;
    const int buf_bytes = dt_floats * 4 + kThinRB * kThinPitch * 2;
    uint8_t* bufs = reinterpret_cast<uint8_t*>(s_wb + TT * 16 * 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp & 3, cgp = warp >> 2;
    const int slice0 = blockIdx.y * 128;                             // first column of this CTA
    const int lr = lane >> 2, lc = (lane & 3) * 2;                   // fragment row / column-pair of this lane
    const float dt_scale = p.dt_scale ? __ldg(p.dt_scale) : 1.f, dt_inv = 1.f / dt_scale;
    const float store_scale = (p.store_scale ? __ldg(p.store_scale) : 1.f);
    const __half* a_g = static_cast<const __half*>(p.a);
    __half* dpre_g = static_cast<__half*>(p.dpre);
    for (int i = tid; i < n_red; i += blockDim.x) s_red[i] = 0.f;
    // B fragments of Wt (k = t, n = c): b0 = {Wt[lc][c], Wt[lc+1][c]}, b1 = {Wt[lc+8][c], Wt[lc+9][c]}, c = slice0 + nt*8 + lr
    for (int i = tid; i < TT * 16 * 32; i += blockDim.x) {
        const int ks = i >> 9, nt = (i >> 5) & 15, l = i & 31;
        const int c = slice0 + nt * 8 + (l >> 2), t0 = ks * 16 + (l & 3) * 2;
        auto w = [&](int t) { return t < T ? p.Wt[(long long)t * p.W + c] : 0.f; };
        s_wb[i] = make_uint2(pack_h2(w(t0), w(t0 + 1)), pack_h2(w(t0 + 8), w(t0 + 9)));
    }
    const long long m_begin = (long long)blockIdx.x * p.rows_per_cta;
    const long long m_end = min(m_begin + p.rows_per_cta, p.M);
    // asynchronous fetch of one 64-row block into buffer `bi` (rows past the end are zero-filled)
    auto prefetch = [&](long long m0, int bi) {
        float* d_dt = reinterpret_cast<float*>(bufs + bi * buf_bytes);
        __half* d_a = reinterpret_cast<__half*>(bufs + bi * buf_bytes + dt_floats * 4);
        const int rows = static_cast<int>(min((long long)kThinRB, m_end - m0));
        for (int idx = tid; idx < kThinRB * T; idx += blockDim.x) {
            int j, rr;
            long long addr;
            if (PLANAR) {
                j = idx / kThinRB; rr = idx - j * kThinRB;
                const long long m = min(m0 + rr, p.M - 1);
                const long long br = m / p.P;
                const int pos = static_cast<int>(m - br * p.P);
                const long long b = br / G;
                const int r = static_cast<int>(br - b * G);
                addr = b * p.dt_outer + (long long)j * p.dt_chan + (long long)r * p.P + pos;
            } else {
                rr = idx / T; j = idx - rr * T;
                addr = min(m0 + rr, p.M - 1) * T + j;
            }
            cp_async_4(d_dt + rr * T + j, p.dt + addr, rr < rows);
        }
        for (int idx = tid; idx < kThinRB * 16; idx += blockDim.x) {     // 16 x 16 bytes per 128-column row
            const int rr = idx >> 4, q = idx & 15;
            cp_async_16(d_a + rr * kThinPitch + q * 8, a_g + min(m0 + rr, p.M - 1) * p.W + slice0 + q * 8, rr < rows);
        }
        cp_async_commit();
    };
    float dw[8][4], dw2[TT == 2 ? 8 : 1][2], dcol[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        dw[nt][0] = dw[nt][1] = dw[nt][2] = dw[nt][3] = 0.f;
        dcol[nt][0] = dcol[nt][1] = 0.f;
        if (TT == 2) dw2[nt][0] = dw2[nt][1] = 0.f;
    }
    float dbt = 0.f;
    if (m_begin < m_end) prefetch(m_begin, 0);
    if (m_begin + kThinRB < m_end) prefetch(m_begin + kThinRB, 1);
    int bi = 0;
    for (long long m0 = m_begin; m0 < m_end; m0 += kThinRB, bi = (bi + 1 == kThinBufs ? 0 : bi + 1)) {
        const int rows = static_cast<int>(min((long long)kThinRB, m_end - m0));
        const bool more1 = m0 + kThinRB < m_end, more2 = m0 + 2 * kThinRB < m_end;
        if (more2) prefetch(m0 + 2 * kThinRB, bi + 2 >= kThinBufs ? bi + 2 - kThinBufs : bi + 2);
        if (more2) cp_async_wait<2>(); else if (more1) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        const float* s_dt = reinterpret_cast<const float*>(bufs + bi * buf_bytes);
        __half* s_a = reinterpret_cast<__half*>(bufs + bi * buf_bytes + dt_floats * 4);
        // ---- fragments of the staged dt block for this warp's 16 rows
        const int r0 = rg * 16 + lr;
        auto dval = [&](int rr, int t) { return t < T ? s_dt[rr * T + t] * dt_scale : 0.f; };
        uint32_t fa[TT][4], ft[TT][4];
        const int rk = rg * 16 + lc;
#pragma unroll
        for (int ks = 0; ks < TT; ++ks) {
            const int tc = ks * 16 + lc, tr = ks * 16 + lr;
            fa[ks][0] = pack_h2(dval(r0, tc), dval(r0, tc + 1));              // S[row][t]        (dpre:  A = S)
            fa[ks][1] = pack_h2(dval(r0 + 8, tc), dval(r0 + 8, tc + 1));
            fa[ks][2] = pack_h2(dval(r0, tc + 8), dval(r0, tc + 9));
            fa[ks][3] = pack_h2(dval(r0 + 8, tc + 8), dval(r0 + 8, tc + 9));
            ft[ks][0] = pack_h2(dval(rk, tr), dval(rk + 1, tr));              // S^T[t][row]      (dWt:   A = S^T)
            ft[ks][1] = pack_h2(dval(rk, tr + 8), dval(rk + 1, tr + 8));
            ft[ks][2] = pack_h2(dval(rk + 8, tr), dval(rk + 9, tr));
            ft[ks][3] = pack_h2(dval(rk + 8, tr + 8), dval(rk + 9, tr + 8));
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {                                   // pairs of 8-column n-tiles
            // B fragments of the activation tile for two n-tiles: ldmatrix.x4.trans, matrix q = (rows (q&1)*8.., cols (q>>1)*8..)
            uint32_t bm[4];
            {
                const int q = lane >> 3, rrow = rg * 16 + (q & 1) * 8 + (lane & 7), ccol = cgp * 64 + np * 16 + (q >> 1) * 8;
                const uint32_t addr = smem_u32(s_a + rrow * kThinPitch + ccol);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bm[0]), "=r"(bm[1]), "=r"(bm[2]), "=r"(bm[3]) : "r"(addr) : "memory");
            }
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int nt = np * 2 + h2;
                mma_m16n8k16(dw[nt], ft[0], bm[2 * h2], bm[2 * h2 + 1]);
                if (TT == 2) {
                    float d2[4] = {dw2[nt][0], dw2[nt][1], 0.f, 0.f};        // rows 24..31 of S^T are zero
                    mma_m16n8k16(d2, ft[TT - 1], bm[2 * h2], bm[2 * h2 + 1]);
                    dw2[nt][0] = d2[0]; dw2[nt][1] = d2[1];
                }
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                const uint2 wb = s_wb[(cgp * 8 + nt) * 32 + lane];
                mma_m16n8k16(c, fa[0], wb.x, wb.y);
                if (TT == 2) {
                    const uint2 wb1 = s_wb[16 * 32 + (cgp * 8 + nt) * 32 + lane];
                    mma_m16n8k16(c, fa[TT - 1], wb1.x, wb1.y);
                }
                const int col = cgp * 64 + nt * 8 + lc;
                __half2* pa0 = reinterpret_cast<__half2*>(s_a + r0 * kThinPitch + col);
                __half2* pa1 = reinterpret_cast<__half2*>(s_a + (r0 + 8) * kThinPitch + col);
                const float2 a0 = __half22float2(*pa0), a1 = __half22float2(*pa1);
                c[0] *= dt_inv * act_grad_from_out(a0.x, p.act); c[1] *= dt_inv * act_grad_from_out(a0.y, p.act);
                c[2] *= dt_inv * act_grad_from_out(a1.x, p.act); c[3] *= dt_inv * act_grad_from_out(a1.y, p.act);
                dcol[nt][0] += c[0] + c[2];
                dcol[nt][1] += c[1] + c[3];
                // in place: this element of the tile is read by no later ldmatrix of the warp (they move on to other columns)
                *pa0 = __floats2half2_rn(c[0] * store_scale, c[1] * store_scale);
                *pa1 = __floats2half2_rn(c[2] * store_scale, c[3] * store_scale);
            }
        }
        if (p.dbt && blockIdx.y == 0 && tid < T) {
            for (int rr = 0; rr < rows; ++rr) dbt += s_dt[rr * T + tid];
        }
        __syncthreads();
        for (int idx = tid; idx < rows * 16; idx += blockDim.x) {          // coalesced 16-byte stores of the dpre tile
            const int rr = idx >> 4, q = idx & 15;
            *reinterpret_cast<uint4*>(dpre_g + (m0 + rr) * p.W + slice0 + q * 8) = *reinterpret_cast<const uint4*>(s_a + rr * kThinPitch + q * 8);
        }
        __syncthreads();                                                   // the tile is free for the prefetch of block i + 2
    }
    // ---- CTA reduction, then one global atomic per output element
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int col = cgp * 64 + nt * 8 + lc;
        if (lr < T) { atomicAdd(s_red + lr * 128 + col, dw[nt][0] * dt_inv); atomicAdd(s_red + lr * 128 + col + 1, dw[nt][1] * dt_inv); }
        if (lr + 8 < T) { atomicAdd(s_red + (lr + 8) * 128 + col, dw[nt][2] * dt_inv); atomicAdd(s_red + (lr + 8) * 128 + col + 1, dw[nt][3] * dt_inv); }
        if (TT == 2 && lr + 16 < T) { atomicAdd(s_red + (lr + 16) * 128 + col, dw2[nt][0] * dt_inv); atomicAdd(s_red + (lr + 16) * 128 + col + 1, dw2[nt][1] * dt_inv); }
        atomicAdd(s_red + T * 128 + col, dcol[nt][0]);
        atomicAdd(s_red + T * 128 + col + 1, dcol[nt][1]);
    }
    if (p.dbt && blockIdx.y == 0 && tid < T) s_red[(T + 1) * 128 + tid] = dbt;
    __syncthreads();
    for (int i = tid; i < T * 128; i += blockDim.x) atomicAdd(p.dWt + (long long)(i >> 7) * p.W + slice0 + (i & 127), s_red[i]);
    if (p.dcol)
        for (int i = tid; i < 128; i += blockDim.x) atomicAdd(p.dcol + slice0 + i, s_red[T * 128 + i]);
    if (p.dbt && blockIdx.y == 0 && tid < T) atomicAdd(p.dbt + tid, s_red[(T + 1) * 128 + tid]);
}

// ------------------------------------------------------------------------------------------
// Streaming variant for the generator's OUTPUT layer (T = n_out <= 4 thin outputs, row-major dt[m][T], fp16 activation):
// with one to four outputs the two thin contractions are 2 T FMAs per element, far below what the load / store stream costs,
// so there is nothing to stage: thread = (8 adjacent columns, row slot), one 16-byte load and one 16-byte store per row, UNR
// rows in flight per thread, Wt / dWt / column sums in registers.  (thin_bwd_mma pads T to 16 for mma.sync and passes every
// tile through shared memory behind two CTA barriers per 64 rows: 0.66 of the HBM floor on this shape.)
// blockDim.x = (W / 8) * slots; dynamic smem: [(T + 1) * W + T] CTA partial sums.
// ------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(256) thin_bwd_stream_kernel(ThinBwdParams p) {
    extern __shared__ float s_red[];
    constexpr int UNR = 4;
    const int cgs = p.W / 8, slots = blockDim.x / cgs;
    const int cg = threadIdx.x % cgs, slot = threadIdx.x / cgs, c0 = cg * 8;
    const int n_red = (T + 1) * p.W + T;
    for (int i = threadIdx.x; i < n_red; i += blockDim.x) s_red[i] = 0.f;
    const float store_scale = p.store_scale ? __ldg(p.store_scale) : 1.f;
    const __half* a_g = static_cast<const __half*>(p.a);
    __half* dpre_g = static_cast<__half*>(p.dpre);
    float w[T][8], dw[T][8], dcol[8], dbt[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        dbt[t] = 0.f;
#pragma unroll
        for (int v = 0; v < 8; ++v) { w[t][v] = __ldg(p.Wt + (long long)t * p.W + c0 + v); dw[t][v] = 0.f; }
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) dcol[v] = 0.f;
    const long long m_begin = (long long)blockIdx.x * p.rows_per_cta;
    const long long m_end = min(m_begin + p.rows_per_cta, p.M);
    const bool tanh_act = p.act == kActTanh;
    for (long long m0 = m_begin + slot; m0 < m_end; m0 += (long long)slots * UNR) {
        uint4 av[UNR];
        float dv[UNR][T];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long long m = m0 + (long long)u * slots;
            if (m < m_end) {
                av[u] = __ldg(reinterpret_cast<const uint4*>(a_g + m * p.W + c0));
#pragma unroll
                for (int t = 0; t < T; ++t) dv[u][t] = __ldg(p.dt + m * T + t);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long long m = m0 + (long long)u * slots;
            if (m >= m_end) break;
            const uint32_t aw[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
            float a8[8], g[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&aw[e]));
                a8[2 * e] = f.x; a8[2 * e + 1] = f.y;
            }
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    acc = fmaf(dv[u][t], w[t][v], acc);
                    dw[t][v] = fmaf(dv[u][t], a8[v], dw[t][v]);
                }
                acc *= tanh_act ? 1.f - a8[v] * a8[v] : lrelu_grad_from_out(a8[v]);
                dcol[v] += acc;
                g[v] = acc * store_scale;
            }
            if (cg == 0) {
#pragma unroll
                for (int t = 0; t < T; ++t) dbt[t] += dv[u][t];
            }
            uint4 o;
            o.x = pack_h2(g[0], g[1]); o.y = pack_h2(g[2], g[3]); o.z = pack_h2(g[4], g[5]); o.w = pack_h2(g[6], g[7]);
            *reinterpret_cast<uint4*>(dpre_g + m * p.W + c0) = o;
        }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int v = 0; v < 8; ++v) atomicAdd(s_red + t * p.W + c0 + v, dw[t][v]);
        if (cg == 0) atomicAdd(s_red + (T + 1) * p.W + t, dbt[t]);
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) atomicAdd(s_red + T * p.W + c0 + v, dcol[v]);
    __syncthreads();
    for (int i = threadIdx.x; i < T * p.W; i += blockDim.x) atomicAdd(p.dWt + i, s_red[i]);
    if (p.dcol)
        for (int i = threadIdx.x; i < p.W; i += blockDim.x) atomicAdd(p.dcol + i, s_red[T * p.W + i]);
    if (p.dbt && threadIdx.x < T) atomicAdd(p.dbt + threadIdx.x, s_red[(T + 1) * p.W + threadIdx.x]);
}

// column sums per group of rows: out[g][c] = sum_{m in group g} x[m][c]   (z-conditioned bias gradient),
// and total[c] += sum over all rows.  grid = (chunks, groups), blockDim.x = W.
// x is fp16 holding value * (1 / *inv_scale).
__global__ void __launch_bounds__(1024) group_colsum_kernel(const __half* __restrict__ x, const float* __restrict__ inv_scale,
                                                           float* __restrict__ out, float* __restrict__ total, int rows_per_group,
                                                           int W, int rows_per_cta) {
    const int c = threadIdx.x, g = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, rows_per_group);
    const __half* xg = x + ((long long)g * rows_per_group) * W;
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) acc += __half2float(xg[(long long)r * W + c]);
    acc *= __ldg(inv_scale);
    if (out) atomicAdd(out + (long long)g * W + c, acc);
    if (total) atomicAdd(total + c, acc);
}

// the same for W % 8 == 0, W <= 2048: thread = (8 adjacent columns, row slot), 16-byte loads, blockDim.x = (W / 8) * slots
// with slots = 256 / (W / 8); the row slots are combined through shared memory before the atomics.
__global__ void __launch_bounds__(256) group_colsum8_kernel(const __half* __restrict__ x, const float* __restrict__ inv_scale,
                                                            float* __restrict__ out, float* __restrict__ total, int rows_per_group,
                                                            int W, int rows_per_cta) {
    extern __shared__ float s_part[];                    // [slots][W]
    const int cgs = W / 8, slots = blockDim.x / cgs;
    const int cg = threadIdx.x % cgs, slot = threadIdx.x / cgs, g = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, rows_per_group);
    const __half* xg = x + ((long long)g * rows_per_group) * W + cg * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int r = r0 + slot; r < r1; r += slots) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(xg + (long long)r * W));
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
            acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s_part[slot * W + cg * 8 + e] = acc[e];
    __syncthreads();
    const float inv = __ldg(inv_scale);
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        float v = 0.f;
        for (int q = 0; q < slots; ++q) v += s_part[q * W + c];
        v *= inv;
        if (out) atomicAdd(out + (long long)g * W + c, v);
        if (total) atomicAdd(total + c, v);
    }
}

// ------------------------------------------------------------------------------------------
// Power-of-two scales that keep the encoder's backward intermediates inside fp16's range.
//   amax[0] = max |x| over a tensor (float bits compared as ints: all values are >= 0)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, float* __restrict__ amax) {
    float m = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));
}

// largest power of two s with s * bound <= 2^14 (fp16 overflows at 65504; the headroom covers fp32 summation slack)
__device__ __forceinline__ float pow2_scale_for(float bound) {
    if (!(bound > 0.f) || !isfinite(bound)) return 1.f;
    int e;
    frexpf(bound, &e);                   // bound = f * 2^e, f in [0.5, 1)
    int k = 14 - e;
    k = k > 100 ? 100 : (k < -100 ? -100 : k);
    return ldexpf(1.f, k);
}

// scales[0] = s1, [1] = 1/s1 : dhpre = (d_heads . Wh) * lrelu'   is stored as fp16(dhpre * s1),  |dhpre| <= amax * max_c sum_t |Wh[t][c]|
// scales[2] = s2, [3] = 1/s2 : dx1   = (dhpre . W2) * lrelu'     is stored as fp16(dx1 * s2),    |dx1| <= bound1 * max_o sum_j |W2[j][o]|
// scales[4] = s2/s1 is not needed: the dx1 GEMM rescales with acc_scale = 1/s1 then store_scale = s2.
// one CTA of 256 threads; O <= 256.
__global__ void __launch_bounds__(256) enc_bwd_scales_kernel(const float* __restrict__ amax, const float* __restrict__ wh, int NH,
                                                             const float* __restrict__ w2, int O, float* __restrict__ scales) {
    __shared__ float red[2][8];
    const int c = threadIdx.x;
    float a = 0.f, b = 0.f;
    if (c < O) {
        for (int t = 0; t < NH; ++t) a += fabsf(wh[t * O + c]);
        for (int j = 0; j < O; ++j) b += fabsf(w2[j * O + c]);    // column c of W2 (O,O): dx1[., c] = sum_j dhpre[., j] W2[j][c]
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((c & 31) == 0) { red[0][c >> 5] = a; red[1][c >> 5] = b; }
    __syncthreads();
    if (c == 0) {
        float ma = 0.f, mb = 0.f;
        for (int w = 0; w < 8; ++w) { ma = fmaxf(ma, red[0][w]); mb = fmaxf(mb, red[1][w]); }
        const float bound1 = amax[0] * ma, bound2 = bound1 * mb;
        const float s1 = pow2_scale_for(bound1), s2 = pow2_scale_for(bound2);
        scales[0] = s1; scales[1] = 1.f / s1; scales[2] = s2; scales[3] = 1.f / s2;
        // scales[6] = s0: the head-map gradients themselves are an fp16 MMA operand (S = fp16(d_heads * s0)).  Unscaled, the
        // z / theta entries - q(t, r) / B times an O(1) factor, 2.6e-8 at cfg5 (150 k cells, B = 256) - fall below fp16's
        // smallest subnormal and the conv_z weight gradient loses 20 % (measured against 8 shards of 32 images).
        scales[6] = pow2_scale_for(amax[0]);
    }
}

// ------------------------------------------------------------------------------------------
// Rotation pooling of InferenceNetwork_AttentionTranslation_UnimodalRotation with groupconv = G > 0 (models.py:301-304):
//   xp[(b,pos)][o] = sum_r fc_w[r] * x1[(b*G + r)*P + pos][o] + fc_b          (nn.Linear(G, 1) over the rotation axis;
// x1 = act(conv1) is the group-conv activation, xp feeds conv2 un-activated).  HBM-bound: one read of x1, one write of xp.
// Thread = (row of the pass, group of 8 channels); a CTA walks `rows_per_cta` consecutive (b,pos) rows.
// ------------------------------------------------------------------------------------------
struct RotPoolParams {
    const __half* x1;        // [B*G*P][O]
    const float* fc_w;       // [G]
    const float* fc_b;       // [1]
    __half* xp;              // fwd out [B*P][O]
    // backward
    const __half* dxp;       // [B*P][O] = d(xp) * s2
    __half* dx1;             // out [B*G*P][O] = d(conv1 pre-activation) * s2 * c
    const float* scales;     // [2] = s2, [3] = 1/s2, [4] = c (power of two, set by rot_pool_scales_kernel), [5] = 1/(s2 c)
    float* dfc_w;            // out [G]  (pre-zeroed)
    float* dfc_b;            // out [1]  (pre-zeroed)
    float* db1;              // conv1 bias gradient slot, stride db1_stride, += sum over rows of d(conv1 pre-activation)
    long long db1_stride;
    int B, G, P, O, act;
    int rows_per_cta;
};

__global__ void rot_pool_scales_kernel(const float* __restrict__ fc_w, int G, float* __restrict__ scales) {
    float mw = 0.f;
    for (int r = 0; r < G; ++r) mw = fmaxf(mw, fabsf(fc_w[r]));
    float c = 1.f;
    while (c * mw > 1.f && c > 1e-30f) c *= 0.5f;        // |w_r| c <= 1: the scaled fp16 gradient cannot grow past d(xp)'s bound
    scales[4] = c;
    scales[5] = scales[3] / c;
}

__device__ __forceinline__ void load8h(const __half* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
        v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
}
__device__ __forceinline__ void store8h(__half* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        w[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256) rot_pool_fwd_kernel(RotPoolParams p) {
    const int cgs = p.O / 8, rpp = 256 / cgs;
    const int cg = threadIdx.x % cgs, rs = threadIdx.x / cgs;
    if (rs >= rpp) return;
    const long long rows = (long long)p.B * p.P;
    const long long row0 = (long long)blockIdx.x * p.rows_per_cta;
    const long long row1 = min(rows, row0 + p.rows_per_cta);
    const float fb = __ldg(p.fc_b);
    for (long long m = row0 + rs; m < row1; m += rpp) {
        const long long b = m / p.P;
        const int pos = static_cast<int>(m - b * p.P);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fb;
        for (int r = 0; r < p.G; ++r) {
            float v[8];
            load8h(p.x1 + (((b * p.G + r) * p.P + pos) * p.O + cg * 8), v);
            const float w = __ldg(p.fc_w + r);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, v[e], acc[e]);
        }
        store8h(p.xp + m * p.O + cg * 8, acc);
    }
}

// Adjoint: d(conv1 pre)[(b,r,pos)][o] = fc_w[r] * d(xp)[(b,pos)][o] * act'(x1[(b,r,pos)][o]),
// dfc_w[r] = sum d(xp) * x1_r, dfc_b = sum d(xp), conv1 bias gradient = column sums of d(conv1 pre).
__global__ void __launch_bounds__(256) rot_pool_bwd_kernel(RotPoolParams p) {
    extern __shared__ float s_red[];                      // [O] bias-gradient partial sums, then [G + 1] fc gradients
    float* s_db = s_red;
    float* s_fc = s_red + p.O;
    for (int i = threadIdx.x; i < p.O + p.G + 1; i += blockDim.x) s_red[i] = 0.f;
    __syncthreads();
    const int cgs = p.O / 8, rpp = 256 / cgs;
    const int cg = threadIdx.x % cgs, rs = threadIdx.x / cgs;
    const long long rows = (long long)p.B * p.P;
    const long long row0 = (long long)blockIdx.x * p.rows_per_cta;
    const long long row1 = min(rows, row0 + p.rows_per_cta);
    const float inv_s2 = p.scales[3], c = p.scales[4], inv_s3 = p.scales[5];
    float db[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) db[e] = 0.f;
    float dfb = 0.f;
    // uniform trip counts for the whole CTA (the warp shuffles below need every lane): lanes without a row contribute zeros
    for (long long mb = row0; mb < row1; mb += rpp) {
        const long long m = mb + rs;
        const bool valid = rs < rpp && m < row1;
        const long long b = valid ? m / p.P : 0;
        const int pos = valid ? static_cast<int>(m - b * p.P) : 0;
        float d[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] = 0.f;
        if (valid) load8h(p.dxp + m * p.O + cg * 8, d);   // d(xp) * s2
#pragma unroll
        for (int e = 0; e < 8; ++e) dfb += d[e];
        for (int r = 0; r < p.G; ++r) {
            float dw = 0.f;
            if (valid) {
                const long long off = ((b * p.G + r) * p.P + pos) * p.O + cg * 8;
                float v[8], g[8];
                load8h(p.x1 + off, v);
                const float w = __ldg(p.fc_w + r) * c;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    dw = fmaf(d[e], v[e], dw);
                    g[e] = d[e] * w * act_grad_from_out(v[e], p.act);      // * s2 * c
                    db[e] += g[e];
                }
                store8h(p.dx1 + off, g);
            }
            // warp-level sum of this rotation's fc_w gradient, then one shared atomic per warp
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) dw += __shfl_xor_sync(0xffffffffu, dw, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(s_fc + r, dw);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(s_db + cg * 8 + e, db[e]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) dfb += __shfl_xor_sync(0xffffffffu, dfb, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(s_fc + p.G, dfb);
    __syncthreads();
    for (int i = threadIdx.x; i < p.O; i += blockDim.x) atomicAdd(p.db1 + (long long)i * p.db1_stride, s_db[i] * inv_s3);
    for (int i = threadIdx.x; i < p.G; i += blockDim.x) atomicAdd(p.dfc_w + i, s_fc[i] * inv_s2);
    if (threadIdx.x == 0) atomicAdd(p.dfc_b, s_fc[p.G] * inv_s2);
}

// Generator backward: dpre_L = (d_yhat . Wout) lrelu', dpre_{i-1} = (dpre_i . W_i) lrelu'; dpre_i is stored as
// fp16(dpre_i * s_i) with s_i = scales[2i], 1/s_i = scales[2i + 1], i = 0..L.  Bounds: |dpre_L| <= amax max_c sum_o |Wout[o][c]|,
// |dpre_{i-1}| <= bound_i max_c sum_j |W_i[j][c]|.
// Step 1: colmax[layer] = max_c sum_j |W[j][c]|.  grid = (column slices of 32, L + 1 layers), block = 32 columns x 32 row
// parts (one CTA per layer walked the whole 1 MB matrix alone: 23 us); colmax zero-filled by the caller.
__global__ void __launch_bounds__(1024) gen_colmax_kernel(const float* __restrict__ wout, int n_out, const float* __restrict__ wh,
                                                          int L, int H, float* __restrict__ colmax) {
    __shared__ float s_part[32][33];
    const int layer = blockIdx.y;                        // layer == L: Wout (n_out x H); else W_{layer+1} = wh[layer] (H x H)
    const float* w = layer == L ? wout : wh + (long long)layer * H * H;
    const int rows = layer == L ? n_out : H;
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float a = 0.f;
    if (c < H)
        for (int j = part; j < rows; j += 32) a += fabsf(__ldg(w + (long long)j * H + c));
    s_part[part][lane] = a;
    __syncthreads();
    if (part == 0) {
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) m += s_part[q][lane];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(colmax + layer), __float_as_int(m));
    }
}
// Step 2 (one thread): chain the bounds.  colmax must be zero-filled before step 1.
__global__ void gen_bwd_scales_kernel(const float* __restrict__ amax, const float* __restrict__ colmax, int L, float* __restrict__ scales) {
    float bound = amax[0];
    for (int layer = L; layer >= 0; --layer) {
        bound *= colmax[layer];
        const float sc = pow2_scale_for(bound);
        scales[2 * layer] = sc;
        scales[2 * layer + 1] = 1.f / sc;
    }
}

// scales[2] = s, [3] = 1/s from amax alone (standalone GroupConv backward)
__global__ void single_scale_kernel(const float* __restrict__ amax, float* __restrict__ scales) {
    const float s = pow2_scale_for(amax[0]);
    scales[2] = s; scales[3] = 1.f / s;
}

__global__ void to_half_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = __float2half_rn(in[i]);
}

// out[c][r] = fp16(in[r][c]) : small weight transposes for the fp16 dgrad GEMMs
__global__ void transpose_half_kernel(const float* __restrict__ in, __half* __restrict__ out, int rows, int cols) {
    const long long total = (long long)rows * cols;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = static_cast<int>(idx % cols);
        const int r = static_cast<int>(idx / cols);
        out[(long long)c * rows + r] = __float2half_rn(in[idx]);
    }
}

// ------------------------------------------------------------------------------------------
// Multi-tensor Adam (torch/optim/adam.py::_single_tensor_adam arithmetic), one launch for every parameter tensor.
// Each CTA owns one kAdamChunk-element slice of one tensor; the slice table travels in the kernel parameters.
// HBM-bound: 16 B read + 12 B written per parameter.
constexpr int kAdamMaxTensors = 48;
constexpr int kAdamChunk = 4096;
struct AdamTable {
    float* param[kAdamMaxTensors];
    const float* grad[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    long long numel[kAdamMaxTensors];
    int block_start[kAdamMaxTensors + 1];     // first CTA of each tensor
    int n;
};
struct AdamHyper {
    float beta1, beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay, step_size, bc2_sqrt;   // 1 - beta formed in double on the host, like torch's Python floats
    int zero_grad;
};
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamHyper& h) {
    if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, p, g);
    m = m + (g - m) * h.one_minus_beta1;                             // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(h.one_minus_beta2 * g, g, v * h.beta2);                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;              // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p - h.step_size * (m / denom);                               // param.addcdiv_(exp_avg, denom, value = -step_size)
}
__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamTable tab, const AdamHyper h) {
    int t = 0;
    while (t + 1 < tab.n && static_cast<int>(blockIdx.x) >= tab.block_start[t + 1]) ++t;
    const long long off = (long long)(blockIdx.x - tab.block_start[t]) * kAdamChunk;
    const long long cnt = min((long long)kAdamChunk, tab.numel[t] - off);
    float* p = tab.param[t] + off;
    const float* g = tab.grad[t] + off;
    float* m = tab.m[t] + off;
    float* v = tab.v[t] + off;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const long long n4 = vec ? cnt / 4 : 0;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        adam_update(pp.x, gg.x, mm.x, vv.x, h); adam_update(pp.y, gg.y, mm.y, vv.y, h);
        adam_update(pp.z, gg.z, mm.z, vv.z, h); adam_update(pp.w, gg.w, mm.w, vv.w, h);
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
        if (h.zero_grad) reinterpret_cast<float4*>(const_cast<float*>(g))[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long i = n4 * 4 + threadIdx.x; i < cnt; i += blockDim.x) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_update(pp, g[i], mm, vv, h);
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (h.zero_grad) const_cast<float*>(g)[i] = 0.f;
    }
}
// running means of train_mnist.py:326-338 on the device: state = {c, elbo, gen_loss, kl}
__global__ void running_means_kernel(const float* __restrict__ elbo, const float* __restrict__ log_p, const float* __restrict__ kl, float b,
                                     float* __restrict__ state) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float c = state[0] + b;
    state[0] = c;
    const float x[3] = {*elbo, -*log_p, *kl};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float delta = b * (x[i] - state[1 + i]);
        state[1 + i] += delta / c;
    }
}

// ------------------------------------------------------------------------------------------
// Real-space CTF filters (src/ctf.py:6-55): one CTA per micrograph, fp64.
//   c[k1][k2] = CTF(fftfreq(n)[k1] / apix, fftfreq(m)[k2] / apix)          kept in shared memory (n*m doubles)
//   out[a][b] = -Re ifft2(c)[(a - n/2) mod n][(b - m/2) mod m]             exact separable inverse DFT, row by row
// dynamic smem: c [n*m] | row transform T [m] complex | twiddles w_n [n], w_m [m] complex
struct CtfFilterParams { const double* params; float* out; int B, n, m; double scale; };
__global__ void __launch_bounds__(256) ctf_filter_kernel(CtfFilterParams p) {
    extern __shared__ __align__(16) double s_ctf[];
    const int n = p.n, m = p.m, b = blockIdx.x;
    double* c = s_ctf;
    double2* T = reinterpret_cast<double2*>(c + ((n * m + 1) & ~1));      // 16-byte aligned
    double2* wn = T + m;
    double2* wm = wn + n;
    const double* q = p.params + 8 * b;
    const double defocus = q[0], cs = q[1] * 1e7, volt = q[2] * 1000.0, apix = q[3] * p.scale, bfac = q[4], w = q[5] / 100.0;
    const double ang0 = 2.0 * 3.14159265358979323846 * q[7] / 360.0;
    const double dfu = defocus * 10000.0, dfv = defocus * 10000.0;
    const double lam = 12.2639 / sqrt(volt + 0.97845e-6 * volt * volt);
    const double a1 = sqrt(1.0 - w * w);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { double sn, cs_; sincospi(2.0 * i / n, &sn, &cs_); wn[i] = make_double2(cs_, sn); }
    for (int i = threadIdx.x; i < m; i += blockDim.x) { double sn, cs_; sincospi(2.0 * i / m, &sn, &cs_); wm[i] = make_double2(cs_, sn); }
    for (int i = threadIdx.x; i < n * m; i += blockDim.x) {
        const int k1 = i / m, k2 = i - k1 * m;
        const double x = static_cast<double>(k1 <= (n - 1) / 2 ? k1 : k1 - n) / n / apix;     // np.fft.fftfreq
        const double y = static_cast<double>(k2 <= (m - 1) / 2 ? k2 : k2 - m) / m / apix;
        const double ang = atan2(y, x), s2 = x * x + y * y;
        const double df = 0.5 * (dfu + dfv + (dfu - dfv) * cos(2.0 * (ang - ang0)));
        const double gamma = 2.0 * 3.14159265358979323846 * (-0.5 * df * lam * s2 + 0.25 * cs * lam * lam * lam * s2 * s2);
        c[i] = (a1 * sin(gamma) - w * cos(gamma)) * exp(-bfac / 4.0 * s2);
    }
    __syncthreads();
    const double inv = 1.0 / (static_cast<double>(n) * m);
    for (int a = 0; a < n; ++a) {
        const int xo = ((a - n / 2) % n + n) % n;                    // fftshift: shifted[a] = orig[(a - n/2) mod n]
        for (int k2 = threadIdx.x; k2 < m; k2 += blockDim.x) {       // T[k2] = sum_k1 c[k1][k2] w_n^(k1 xo)
            double re = 0.0, im = 0.0;
            int j = 0;
            for (int k1 = 0; k1 < n; ++k1) {
                const double v = c[k1 * m + k2];
                re = fma(v, wn[j].x, re);
                im = fma(v, wn[j].y, im);
                j += xo; if (j >= n) j -= n;
            }
            T[k2] = make_double2(re, im);
        }
        __syncthreads();
        for (int bb = threadIdx.x; bb < m; bb += blockDim.x) {       // out = Re sum_k2 T[k2] w_m^(k2 yo)
            const int yo = ((bb - m / 2) % m + m) % m;
            double re = 0.0;
            int j = 0;
            for (int k2 = 0; k2 < m; ++k2) {
                re = fma(T[k2].x, wm[j].x, re);
                re = fma(-T[k2].y, wm[j].y, re);
                j += yo; if (j >= m) j -= m;
            }
            p.out[((long long)b * n + a) * m + bb] = static_cast<float>(-re * inv);
        }
        __syncthreads();
    }
}

// centre crop + per-image standardisation (src/image.py:30-42, train_particles.py:592-600); one CTA per image, fp64 sums
// T = element type of the stack as stored: float (arrays / MRC mode 2) or the integer types of MRC modes 0, 1, 6
// (src/mrc.py:118-131); every value is widened to double first, like numpy's mean / std over an integer array.
template <typename T>
__global__ void __launch_bounds__(256) crop_normalize_kernel(const T* __restrict__ in, float* __restrict__ out, int n, int m, int c0,
                                                             int c1, int si, int sj, int normalize) {
    __shared__ double s_red[2][8];
    __shared__ double s_stat[2];
    const int b = blockIdx.x, total = c0 * c1;
    const T* src = in + (long long)b * n * m;
    double sum = 0.0, sq = 0.0;
    if (normalize) {
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int r = i / c1, cc = i - r * c1;
            const double v = static_cast<double>(src[(si + r) * m + sj + cc]);
            sum += v; sq += v * v;
        }
        for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
        if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = sum; s_red[1][threadIdx.x >> 5] = sq; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, q = 0.0;
            for (int w = 0; w < 8; ++w) { a += s_red[0][w]; q += s_red[1][w]; }
            const double mu = a / total;
            double var = q / total - mu * mu;
            // second pass for the variance would be exact; the two-moment form in fp64 is accurate to ~1e-13 relative here
            if (var < 0.0) var = 0.0;
            s_stat[0] = mu; s_stat[1] = sqrt(var);
        }
        __syncthreads();
    }
    const double mu = normalize ? s_stat[0] : 0.0, sd = normalize ? s_stat[1] : 1.0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i / c1, cc = i - r * c1;
        const double v = static_cast<double>(src[(si + r) * m + sj + cc]);
        out[(long long)b * total + i] = static_cast<float>((v - mu) / sd);
    }
}

}  // namespace tvae
