// CUDA-core helpers around the generator GEMMs: latent bias, no-Fourier first layer, coordinate-gradient
// reductions, thin output layer.
#pragma once
#include "gen_policies.cuh"
#include "simt_kernels.cuh"

namespace tvae {

// zb[b][j] = sum_d z[b][d] Wz[j][d]   (latent_linear, models.py:114)
__global__ void latent_bias_kernel(const float* __restrict__ z, const float* __restrict__ wz, float* __restrict__ zb, int B, int H, int zdim) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    float acc = 0.f;
    for (int d = 0; d < zdim; ++d) acc = fmaf(z[b * zdim + d], wz[j * zdim + d], acc);
    zb[idx] = acc;
}
// dWz[j][d] = sum_b dzb[b][j] z[b][d];  dz[b][d] = sum_j dzb[b][j] Wz[j][d]
__global__ void latent_bias_bwd_kernel(const float* __restrict__ dzb, const float* __restrict__ z, const float* __restrict__ wz,
                                       float* __restrict__ dwz, float* __restrict__ dz, int B, int H, int zdim) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < H * zdim) {
        const int j = idx / zdim, d = idx - j * zdim;
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc = fmaf(dzb[b * H + j], z[b * zdim + d], acc);
        dwz[idx] = acc;
    }
    if (idx < B * zdim) {
        const int b = idx / zdim, d = idx - b * zdim;
        float acc = 0.f;
        for (int j = 0; j < H; ++j) acc = fmaf(dzb[b * H + j], wz[j * zdim + d], acc);
        dz[idx] = acc;
    }
}

// No Fourier expansion (cfg2): a0[m][j] = LeakyReLU(x'0 W1[j][0] + x'1 W1[j][1] + b1[j] + zb[b][j]).  blockDim.x = H.
__global__ void __launch_bounds__(1024) coord_layer_fwd_kernel(CoordXform cx, const float* __restrict__ w1, const float* __restrict__ b1,
                                                               const float* __restrict__ zb, float* __restrict__ a0, int H, int rows_per_cta) {
    const int j = threadIdx.x;
    const float wx = w1[2 * j], wy = w1[2 * j + 1], bb = b1[j];
    const long long m0 = (long long)blockIdx.x * rows_per_cta;
    for (long long m = m0; m < min(m0 + rows_per_cta, cx.M); ++m) {
        float x0, x1;
        transformed_coord(cx, m, x0, x1);
        const float zz = zb ? zb[(m / cx.N) * H + j] : 0.f;
        a0[m * H + j] = to_tf32(lrelu(fmaf(x1, wy, fmaf(x0, wx, bb)) + zz));
    }
}
// backward of the above, weight part: dW1[j][0..1] += sum_m dpre[m][j] x'[m]   (thread = column j)
__global__ void __launch_bounds__(1024) coord_layer_bwd_w_kernel(CoordXform cx, const float* __restrict__ dpre, float* __restrict__ dw1,
                                                                 int H, int rows_per_cta) {
    const int j = threadIdx.x;
    float dwx = 0.f, dwy = 0.f;
    const long long m0 = (long long)blockIdx.x * rows_per_cta;
    for (long long m = m0; m < min(m0 + rows_per_cta, cx.M); ++m) {
        float x0, x1;
        transformed_coord(cx, m, x0, x1);
        const float g = dpre[m * H + j];
        dwx = fmaf(g, x0, dwx);
        dwy = fmaf(g, x1, dwy);
    }
    atomicAdd(dw1 + 2 * j, dwx);
    atomicAdd(dw1 + 2 * j + 1, dwy);
}
// coordinate part: dxp[m] = sum_j dpre[m][j] W1[j][:]   (warp per row)
__global__ void __launch_bounds__(256) coord_layer_bwd_x_kernel(const float* __restrict__ w1, const float* __restrict__ dpre,
                                                                float* __restrict__ dxp, long long M, int H) {
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    float g0 = 0.f, g1 = 0.f;
    for (int j = lane; j < H; j += 32) {
        const float g = dpre[m * H + j];
        g0 = fmaf(g, __ldg(w1 + 2 * j), g0);
        g1 = fmaf(g, __ldg(w1 + 2 * j + 1), g1);
    }
    g0 = warp_sum(g0);
    g1 = warp_sum(g1);
    if (lane == 0) { dxp[2 * m] = g0; dxp[2 * m + 1] = g1; }
}

// dxp (B*N,2) -> d_theta (B), d_dx (B,2) through x' = (x - dx) R(theta)   (train_mnist.py:222,234-239).
// One CTA per image.
__global__ void __launch_bounds__(256) coord_xform_bwd_kernel(CoordXform cx, const float* __restrict__ dxp, float* __restrict__ d_theta,
                                                              float* __restrict__ d_dx) {
    __shared__ float scratch[96];
    const int b = blockIdx.x;
    float sn, cs;
    sincosf(cx.theta[b], &sn, &cs);
    float acc[3] = {0.f, 0.f, 0.f};
    for (int px = threadIdx.x; px < cx.N; px += blockDim.x) {
        const long long m = (long long)b * cx.N + px;
        float x0, x1;
        transformed_coord(cx, m, x0, x1);
        const float g0 = dxp[2 * m], g1 = dxp[2 * m + 1];
        acc[0] += -g0 * x1 + g1 * x0;              // d/dtheta
        acc[1] += -(g0 * cs + g1 * sn);            // d/d dx_0
        acc[2] += -(-g0 * sn + g1 * cs);           // d/d dx_1
    }
    block_reduce<3, false>(acc, scratch);
    if (threadIdx.x == 0) {
        d_theta[b] = acc[0];
        d_dx[2 * b] = acc[1];
        d_dx[2 * b + 1] = acc[2];
    }
}

// thin forward (output layer straight after layer 1 when num_layers == 1): y[m][o] = sum_c a[m][c] W[o][c] + b[o]; warp per row.
__global__ void __launch_bounds__(256) thin_fwd_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ y, long long M, int W, int T) {
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    for (int o = 0; o < T; ++o) {
        float acc = 0.f;
        for (int c = lane; c < W; c += 32) acc = fmaf(a[m * W + c], w[(long long)o * W + c], acc);
        acc = warp_sum(acc);
        if (lane == 0) y[m * T + o] = acc + bias[o];
    }
}

__global__ void round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = to_tf32(in[i]);
}

// dbias[o] = sum_r dbank[(r*O + o)][K]
__global__ void bank_bias_grad_kernel(const float* __restrict__ dbank, float* __restrict__ dbias, int O, int G, int kpad, int K) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    float acc = 0.f;
    for (int r = 0; r < G; ++r) acc += dbank[((long long)r * O + o) * kpad + K];
    dbias[o] = acc;
}

}  // namespace tvae
