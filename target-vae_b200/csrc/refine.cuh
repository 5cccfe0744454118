// Exact (fp32-operand) re-evaluation of the attention logits at candidate cells: the argmax (rotation, translation)
// assignment of clustering_*.get_latent (clustering_mnist.py:127, `attn.view(B,-1).max(1)`).
//
// The tensor-core encoder computes its contractions with FP16 operands; its logits move by ~1e-4 of the logit range
// against an fp32 evaluation, so on an image whose two best cells are closer than that the argmax can flip.  Instead of
// running the whole conv1 -> conv2 -> heads chain error-compensated (3 MMAs per product, 3x the encoder), the fast maps
// are used as a FILTER: every cell whose fast logit is within `rel_tol` x (max - min) of the fast maximum - a band many
// times wider than the fast path's error - is a candidate, and the logit chain is re-evaluated there from the fp32
// image, the fp32 bilinear-rotated filter (models.py:174-197) and the fp32 weights with fp32 products summed in DOUBLE,
// rounded to fp32 where the reference's fp32 layers round (after each convolution, bias add and activation).  The
// refined argmax and the z / theta values at it are then those of an fp32 evaluation.  Cost: one 2 * O * C * k^2 flop
// dot-product block per candidate (typically 1-3 per image) against L2-resident filters.
//
//   refine_select_kernel   one CTA per image: candidates (<= kRefineMaxCand) inside the band; the band is halved until they fit
//   refine_eval_kernel     one CTA per (candidate, image): conv1 at that cell for all O channels -> act -> conv2 -> act -> heads
//   refine_pick_kernel     one warp per image: first-index argmax of the refined logits, outputs of get_latent
#pragma once
#include "simt_kernels.cuh"

namespace tvae {

constexpr int kRefineMaxCand = 32;

// fp32 rotated bank [G*O][K] (the value filter_bank_fwd_kernel rounds to fp16), row n' = r*O + o
__global__ void filter_bank_f32_kernel(const float* __restrict__ w, float* __restrict__ bank, int O, int C, int k, int G, RotTable rot) {
    const int K = C * k * k;
    const long long total = (long long)G * O * K;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int kk = static_cast<int>(idx % K);
        const int np = static_cast<int>(idx / K);
        const int r = np / O, o = np - r * O;
        const int u = kk % k, v = (kk / k) % k, c = kk / (k * k);
        const BilinearTap t = rot_tap(u, v, k, rot.cs[r], rot.sn[r]);
        const float* wp = w + ((long long)o * C + c) * k * k;
        const bool x0ok = t.x0 >= 0 && t.x0 < k, x1ok = t.x0 + 1 >= 0 && t.x0 + 1 < k;
        const bool y0ok = t.y0 >= 0 && t.y0 < k, y1ok = t.y0 + 1 >= 0 && t.y0 + 1 < k;
        float val = 0.f;
        if (y0ok && x0ok) val += wp[t.y0 * k + t.x0] * (t.wy0 * t.wx0);
        if (y0ok && x1ok) val += wp[t.y0 * k + t.x0 + 1] * (t.wy0 * t.wx1);
        if (y1ok && x0ok) val += wp[(t.y0 + 1) * k + t.x0] * (t.wy1 * t.wx0);
        if (y1ok && x1ok) val += wp[(t.y0 + 1) * k + t.x0 + 1] * (t.wy1 * t.wx1);
        bank[idx] = val;
    }
}

// heads (B, NH, G2, P): channel 0 is the attention map (already + p_r).  cand (B, kRefineMaxCand), n_cand (B).
__global__ void __launch_bounds__(1024) refine_select_kernel(const float* __restrict__ heads, int NH, int L, float rel_tol,
                                                             int* __restrict__ cand, int* __restrict__ n_cand) {
    __shared__ float scratch[64];
    __shared__ int s_n;
    __shared__ int s_list[kRefineMaxCand];
    const int b = blockIdx.x;
    const float* hb = heads + (long long)b * NH * L;
    float mx[1] = {-CUDART_INF_F}, mn[1] = {-CUDART_INF_F};
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const float v = hb[l];
        mx[0] = fmaxf(mx[0], v);
        mn[0] = fmaxf(mn[0], -v);
    }
    block_reduce<1, true>(mx, scratch);
    block_reduce<1, true>(mn, scratch);
    float band = rel_tol * (mx[0] + mn[0]);          // rel_tol x (max - min)
    for (int it = 0; it < 40; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const float thr = mx[0] - band;
        for (int l = threadIdx.x; l < L; l += blockDim.x) {
            if (hb[l] >= thr) {
                const int slot = atomicAdd(&s_n, 1);
                if (slot < kRefineMaxCand) s_list[slot] = l;
            }
        }
        __syncthreads();
        if (s_n <= kRefineMaxCand) break;
        band *= 0.5f;                                  // more near-ties than slots: narrow the band (uniform decision)
    }
    // (a band of 0 keeps the cells equal to the maximum; more than kRefineMaxCand exact ties keeps an arbitrary subset plus
    // - below - the fast argmax itself, which is then what the refined argmax can fall back to)
    const int n = s_n < kRefineMaxCand ? s_n : kRefineMaxCand;
    if (threadIdx.x < n) cand[b * kRefineMaxCand + threadIdx.x] = s_list[threadIdx.x];
    if (threadIdx.x == 0) n_cand[b] = n;
}

struct RefineEvalParams {
    const float* y;          // (B,C,n,n)
    const float* bank32;     // [G*O][K] fp32 rotated filters
    const float* conv1_bias; // (O)
    const float* w2;         // (O,O)
    const float* b2;         // (O)
    const float* wh;         // [NH][O]
    const float* bh;         // [NH]
    const float* head_add;   // [NH][G2]
    const float* fc_w;       // (G) or null: rotation pooling between conv1 and conv2 (attention space has one slot)
    const float* fc_b;       // (1)
    const int* cand;         // (B, kRefineMaxCand) cell index l = r*P + pos in the attention space
    const int* n_cand;       // (B)
    float* cand_heads;       // (B, kRefineMaxCand, NH)
    int C, n, k, p, G, O, d, NH, act;
};

__device__ __forceinline__ float refine_act(float x, int act) { return act == kActTanh ? tanhf(x) : lrelu(x); }

// dynamic smem: patch [K] floats, x1 [O], xacc [O], h [O]
__global__ void __launch_bounds__(256) refine_eval_kernel(RefineEvalParams p) {
    extern __shared__ float s_ref[];
    const int b = blockIdx.y, ci = blockIdx.x;
    if (ci >= p.n_cand[b]) return;
    const int K = p.C * p.k * p.k, P = p.d * p.d;
    float* patch = s_ref;
    float* x1 = patch + ((K + 3) & ~3);
    float* xacc = x1 + p.O;
    float* h = xacc + p.O;
    const int l = p.cand[b * kRefineMaxCand + ci];
    const int slot = l / P, pos = l - slot * P;          // slot = rotation (or 0 with pooling)
    const int i = pos / p.d, j = pos - i * p.d;
    // im2col row of this cell: patch[(c*k + v)*k + u] = y[b, c, i+v-p, j+u-p] (zero outside)   (models.py:215)
    const float* yb = p.y + (long long)b * p.C * p.n * p.n;
    for (int kk = threadIdx.x; kk < K; kk += blockDim.x) {
        const Im2colCursor cur = im2col_cursor(kk, p.k);
        const int iy = i + cur.v - p.p, ix = j + cur.u - p.p;
        patch[kk] = (iy >= 0 && iy < p.n && ix >= 0 && ix < p.n) ? yb[((long long)cur.c * p.n + iy) * p.n + ix] : 0.f;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const bool pooled = p.fc_w != nullptr;
    const int r_begin = pooled ? 0 : slot, r_end = pooled ? p.G : slot + 1;
    for (int o = threadIdx.x; o < p.O; o += blockDim.x) xacc[o] = 0.f;
    __syncthreads();
    for (int r = r_begin; r < r_end; ++r) {
        // conv1 at this cell, rotation r, all O channels: fp32 products, double accumulation, 4 channels per warp pass
        for (int o0 = warp * 4; o0 < p.O; o0 += nw * 4) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const float* w0 = p.bank32 + ((long long)r * p.O + o0) * K;
            for (int kk = lane; kk < K; kk += 32) {
                const float a = patch[kk];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (o0 + q < p.O) acc[q] += (double)a * (double)__ldg(w0 + (long long)q * K + kk);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], s);
            }
            if (lane < 4 && o0 + lane < p.O) {
                const int o = o0 + lane;
                const double a = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
                // the reference's layers are fp32: the conv output, the bias add and the activation each round to fp32
                const float v = refine_act(static_cast<float>(a) + p.conv1_bias[o], p.act);
                if (pooled) xacc[o] += p.fc_w[r] * v;       // fc_r over the rotation axis (models.py:301-304), same thread per o
                else x1[o] = v;
            }
        }
        __syncthreads();
    }
    if (pooled) {
        for (int o = threadIdx.x; o < p.O; o += blockDim.x) x1[o] = xacc[o] + p.fc_b[0];
        __syncthreads();
    }
    // conv2 (1x1x1): h = act(W2 x1 + b2)
    for (int o2 = warp; o2 < p.O; o2 += nw) {
        double acc = 0.0;
        for (int o = lane; o < p.O; o += 32) acc += (double)x1[o] * (double)__ldg(p.w2 + (long long)o2 * p.O + o);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) h[o2] = refine_act(static_cast<float>(acc) + p.b2[o2], p.act);
    }
    __syncthreads();
    // heads (+ p_r / rotation offset table)
    const int G2 = pooled ? 1 : p.G;
    for (int t = warp; t < p.NH; t += nw) {
        double acc = 0.0;
        for (int o = lane; o < p.O; o += 32) acc += (double)h[o] * (double)__ldg(p.wh + (long long)t * p.O + o);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0)
            p.cand_heads[((long long)b * kRefineMaxCand + ci) * p.NH + t] = (static_cast<float>(acc) + p.bh[t]) + p.head_add[t * G2 + slot];
    }
}

// one warp per image: first-index argmax of the refined logits (torch.max returns the first maximal index), then
// z_content = [z_mu, exp(z_logstd)], theta_mu at it (clustering_mnist.py:140-161)
__global__ void refine_pick_kernel(const int* __restrict__ cand, const int* __restrict__ n_cand, const float* __restrict__ cand_heads,
                                   int NH, int Z, float* __restrict__ z_content, float* __restrict__ theta_mu, int* __restrict__ argmax_out,
                                   float* __restrict__ refined_logit) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int n = n_cand[b];
    float best = -CUDART_INF_F;
    int bi = 0x7fffffff, bslot = -1;
    if (lane < n) {
        best = cand_heads[((long long)b * kRefineMaxCand + lane) * NH];
        bi = cand[b * kRefineMaxCand + lane];
        bslot = lane;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int os = __shfl_xor_sync(0xffffffffu, bslot, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; bslot = os; }
    }
    if (lane == 0 && bslot >= 0) {
        const float* hv = cand_heads + ((long long)b * kRefineMaxCand + bslot) * NH;
        argmax_out[b] = bi;
        if (refined_logit) refined_logit[b] = best;
        theta_mu[b] = hv[1];
        for (int k = 0; k < Z; ++k) {
            z_content[b * 2 * Z + k] = hv[3 + k];
            z_content[b * 2 * Z + Z + k] = expf(hv[3 + Z + k]);
        }
    }
}

}  // namespace tvae
