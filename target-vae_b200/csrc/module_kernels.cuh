// Kernels behind the STANDALONE module interfaces of src/models.py - the calls a user of the reference can make outside the
// fused training step: RandomFourierEmbedding2d.forward (models.py:53-58), ResidLinear.forward (models.py:29-30) and the
// module-level tail of the encoder, q_t_r = log_softmax(attn), a_sampled = gumbel_softmax(attn) (models.py:383-388), with
// their backward passes.  None of them is on the timed path (the fused step generates the Fourier features inside the layer-1
// GEMM and never materialises q_t_r / a_sampled); they exist so that every forward of the drop-in classes runs on the device
// library instead of raising.
#pragma once
#include "simt_kernels.cuh"

namespace tvae {

// out[m][f] = cos(x[m][0] * w[f][0] + x[m][1] * w[f][1] + b[f])          w = embed_latent.weight / sigma
__global__ void __launch_bounds__(256) fourier_embed_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                                float* __restrict__ out, long long M, int E) {
    const long long total = M * E;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / E;
        const int f = static_cast<int>(i - m * E);
        const float2 xv = __ldg(reinterpret_cast<const float2*>(x) + m);
        // F.linear(x, w, b): fp32 dot product, then the bias
        out[i] = cosf(fmaf(xv.y, __ldg(w + 2 * f + 1), xv.x * __ldg(w + 2 * f)) + __ldg(b + f));
    }
}
// dx[m] = sum_f g[m][f] * (-sin(phase)) * (w[f][0], w[f][1]); one warp per row
__global__ void __launch_bounds__(256) fourier_embed_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                                const float* __restrict__ g, float* __restrict__ dx, long long M, int E) {
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    const float2 xv = __ldg(reinterpret_cast<const float2*>(x) + m);
    float a0 = 0.f, a1 = 0.f;
    for (int f = lane; f < E; f += 32) {
        const float w0 = __ldg(w + 2 * f), w1 = __ldg(w + 2 * f + 1);
        const float t = -sinf(fmaf(xv.y, w1, xv.x * w0) + __ldg(b + f)) * g[m * E + f];
        a0 = fmaf(t, w0, a0);
        a1 = fmaf(t, w1, a1);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) { dx[2 * m] = a0; dx[2 * m + 1] = a1; }
}

// fp16 operand copies of a Linear weight [N][K]: w16 = fp16(W (+ I)), wt16 = its transpose [K][N] (either may be null)
__global__ void weight_to_half_kernel(const float* __restrict__ w, __half* __restrict__ w16, __half* __restrict__ wt16, int N, int K, int resid) {
    const long long total = (long long)N * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = static_cast<int>(i / K), k = static_cast<int>(i - (long long)n * K);
        const float v = w[i] + ((resid && n == k) ? 1.f : 0.f);
        if (w16) w16[i] = __float2half_rn(v);
        if (wt16) wt16[(long long)k * N + n] = __float2half_rn(v);
    }
}
// y += x (the residual add of ResidLinear happens before the activation, so it is folded into the weight; this kernel is
// only the elementwise part of the backward): dpre = g * act'(y); out16 = fp16(dpre * scales[2]); colsum[c] += sum_r dpre[r][c]
__global__ void __launch_bounds__(256) actgrad_to_half_colsum_kernel(const float* __restrict__ g, const float* __restrict__ y, __half* __restrict__ out16,
                                                                     const float* __restrict__ scales, float* __restrict__ colsum, long long R, int W,
                                                                     int rows_per_cta, int act) {
    extern __shared__ float sm_cs2[];
    for (int c = threadIdx.x; c < W; c += blockDim.x) sm_cs2[c] = 0.f;
    __syncthreads();
    const float sc = __ldg(scales + 2);
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        float s = 0.f;
        for (long long r = r0; r < r0 + rows_per_cta && r < R; ++r) {
            const float d = g[r * W + c] * act_grad_from_out(y[r * W + c], act);
            out16[r * W + c] = __float2half_rn(d * sc);
            s += d;
        }
        sm_cs2[c] = s;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) atomicAdd(colsum + c, sm_cs2[c]);
}
// amax of |g * act'(y)| (bound for the fp16 scale)
__global__ void __launch_bounds__(256) actgrad_absmax_kernel(const float* __restrict__ g, const float* __restrict__ y, long long n, int act, float* __restrict__ amax) {
    float m = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(g[i] * act_grad_from_out(y[i], act)));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));
}

// backward of softmax_pair_kernel: d_attn = dq - exp(q) * sum(dq) + a * (da - sum(a * da)); one CTA per image
__global__ void __launch_bounds__(1024) softmax_pair_bwd_kernel(const float* __restrict__ q, const float* __restrict__ a, const float* __restrict__ dq,
                                                                const float* __restrict__ da, float* __restrict__ d_attn, int L) {
    __shared__ float scratch[64];
    const long long o = (long long)blockIdx.x * L;
    float s[2] = {0.f, 0.f};
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        if (dq) s[0] += dq[o + l];
        if (da) s[1] += a[o + l] * da[o + l];
    }
    block_reduce<2, false>(s, scratch);
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        float d = 0.f;
        if (dq) d += dq[o + l] - expf(q[o + l]) * s[0];
        if (da) d += a[o + l] * (da[o + l] - s[1]);
        d_attn[o + l] = d;
    }
}

// Gradient of GroupConv.forward (models.py:202-225) w.r.t. its INPUT image - only needed when GroupConv is used on its own
// with an input that requires a gradient (the training step never does: the image is data).  CUDA-core kernel, fp32:
//   dy[b][c][iy][ix] = sum_{r,o} sum_{u,v} g[(b,r,u*d+v)][o] * bank[(r,o)][c][iy-u+p][ix-v+p]
// one CTA per (image row iy, b*C + c); thread = (ix, half of the o range), four o per step (16-byte loads of g).
__global__ void __launch_bounds__(256) groupconv_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ bank, float* __restrict__ dy,
                                                              int C, int n, int k, int p, int G, int O, int d) {
    __shared__ float part[256];
    const int iy = blockIdx.x, bc = blockIdx.y, b = bc / C, c = bc - b * C;
    const int K = C * k * k, P = d * d;
    const int xl = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int o_begin = half * (O / 2), o_end = o_begin + O / 2;           // O % 8 == 0 (host)
    const int u_lo = max(0, iy + p - k + 1), u_hi = min(d - 1, iy + p);
    for (int ix0 = 0; ix0 < n; ix0 += 128) {
        const int ix = ix0 + xl;
        float acc = 0.f;
        if (ix < n) {
            const int v_lo = max(0, ix + p - k + 1), v_hi = min(d - 1, ix + p);
            for (int r = 0; r < G; ++r) {
                const float* gr = g + ((long long)(b * G + r) * P) * O;
                const float* br = bank + (long long)r * O * K + (long long)c * k * k;
                for (int u = u_lo; u <= u_hi; ++u) {
                    const int ky = iy - u + p;
                    for (int o = o_begin; o < o_end; o += 4) {
                        const float* b0 = br + (long long)o * K + ky * k + ix + p;       // + (-v) below
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                        for (int v = v_lo; v <= v_hi; ++v) {
                            const float4 gq = __ldg(reinterpret_cast<const float4*>(gr + (long long)(u * d + v) * O + o));
                            const float* bb = b0 - v;
                            a0 = fmaf(gq.x, __ldg(bb), a0);
                            a1 = fmaf(gq.y, __ldg(bb + K), a1);
                            a2 = fmaf(gq.z, __ldg(bb + 2 * K), a2);
                            a3 = fmaf(gq.w, __ldg(bb + 3 * K), a3);
                        }
                        acc += (a0 + a1) + (a2 + a3);
                    }
                }
            }
        }
        part[threadIdx.x] = acc;
        __syncthreads();
        if (half == 0 && ix < n) dy[((long long)bc * n + iy) * n + ix] = part[xl] + part[xl + 128];
        __syncthreads();
    }
}

}  // namespace tvae
