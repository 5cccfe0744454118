// CUDA-core helpers around the generator GEMMs: latent bias, no-Fourier first layer, coordinate-gradient
// reductions, thin output layer.
#pragma once
#include <cuda_bf16.h>
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
// dWz[j][d] = sum_b dzb[b][j] z[b][d];  dz[b][d] = sum_j dzb[b][j] Wz[j][d].  One warp per output element (H zdim + B zdim
// warps, lanes stride over the reduction index, shuffle reduction): the one-thread-per-output form ran a dependent chain of H
// strided loads per thread - 36 us for 1 224 outputs.
__global__ void __launch_bounds__(128) latent_bias_bwd_kernel(const float* __restrict__ dzb, const float* __restrict__ z, const float* __restrict__ wz,
                                                              float* __restrict__ dwz, float* __restrict__ dz, int B, int H, int zdim) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float acc = 0.f;
    float* dst = nullptr;
    if (w < H * zdim) {
        const int j = w / zdim, d = w - j * zdim;
        for (int b = lane; b < B; b += 32) acc = fmaf(dzb[b * H + j], z[b * zdim + d], acc);
        dst = dwz + w;
    } else if (w < (H + B) * zdim) {
        const int i = w - H * zdim;
        const int b = i / zdim, d = i - b * zdim;
        for (int j = lane; j < H; j += 32) acc = fmaf(dzb[b * H + j], wz[j * zdim + d], acc);
        dst = dz + i;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (dst && lane == 0) *dst = acc;
}

// No Fourier expansion (cfg2): a0[m][j] = LeakyReLU(x'0 W1[j][0] + x'1 W1[j][1] + b1[j] + zb[b][j]), stored fp16.
// CTA = kCoordRB rows; the transformed coordinates of the block are computed once into shared memory, then
// thread = VEC adjacent columns x one row slot streams 16-byte (VEC = 8) or 8-byte (VEC = 4) stores (HBM-bound: one
// write of a0).
constexpr int kCoordRB = 64;
template <int VEC, bool TANH>
__global__ void __launch_bounds__(256) coord_layer_fwd_kernel(CoordXform cx, const float* __restrict__ w1, const float* __restrict__ b1,
                                                              const float* __restrict__ zb, __half* __restrict__ a0, int H) {
    __shared__ float2 s_x[kCoordRB];
    const int cgs = H / VEC, rpp = blockDim.x / cgs;
    const int cg = threadIdx.x % cgs, rs = threadIdx.x / cgs;
    const long long m0 = (long long)blockIdx.x * kCoordRB;
    if (threadIdx.x < kCoordRB) {
        float x0, x1;
        transformed_coord(cx, m0 + threadIdx.x, x0, x1);
        s_x[threadIdx.x] = make_float2(x0, x1);
    }
    const int c0 = cg * VEC;
    float wx[VEC], wy[VEC], bb[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { wx[v] = w1[2 * (c0 + v)]; wy[v] = w1[2 * (c0 + v) + 1]; bb[v] = b1[c0 + v]; }
    __syncthreads();
    const int rows = static_cast<int>(min((long long)kCoordRB, cx.M - m0));
    // image of a row without a 64-bit division per row: one division per thread, then a running remainder.  A block that
    // lies inside one image (always, when the pixels per image are a multiple of kCoordRB) folds the latent bias of that
    // image into the bias registers and its row loop issues no loads at all.
    const long long b_first = m0 / cx.N;
    const long long rem0 = m0 - b_first * cx.N;
    const bool one_image = rem0 + rows <= cx.N;
    if (zb && one_image) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) bb[v] += __ldg(zb + b_first * H + c0 + v);
    }
    for (int rr = rs; rr < rows; rr += rpp) {
        const long long m = m0 + rr;
        const float2 x = s_x[rr];
        long long b = b_first;
        if (zb && !one_image) {
            long long rem = rem0 + rr;
            while (rem >= cx.N) { rem -= cx.N; ++b; }
        }
        float o[VEC];
#pragma unroll
        for (int q = 0; q < VEC; q += 4) {
            float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (zb && !one_image) z4 = __ldg(reinterpret_cast<const float4*>(zb + b * H + c0 + q));
            o[q] = fmaf(x.y, wy[q], fmaf(x.x, wx[q], bb[q])) + z4.x;
            o[q + 1] = fmaf(x.y, wy[q + 1], fmaf(x.x, wx[q + 1], bb[q + 1])) + z4.y;
            o[q + 2] = fmaf(x.y, wy[q + 2], fmaf(x.x, wx[q + 2], bb[q + 2])) + z4.z;
            o[q + 3] = fmaf(x.y, wy[q + 3], fmaf(x.x, wx[q + 3], bb[q + 3])) + z4.w;
        }
        act_vec<TANH>(o);
        if constexpr (VEC == 8)
            *reinterpret_cast<uint4*>(a0 + m * H + c0) = make_uint4(pack_half2(o[0], o[1]), pack_half2(o[2], o[3]), pack_half2(o[4], o[5]), pack_half2(o[6], o[7]));
        else
            *reinterpret_cast<uint2*>(a0 + m * H + c0) = make_uint2(pack_half2(o[0], o[1]), pack_half2(o[2], o[3]));
    }
}
// backward of the above in one pass over dpre:  dW1[j][0..1] += sum_m dpre[m][j] x'[m]  and
// dxp[m] = sum_j dpre[m][j] W1[j][:]   (CTA = rows_per_cta rows in blocks of kCoordRB).  When dzb != null every CTA lies
// inside one image (rows_per_cta divides the pixels per image) and the same pass also yields the bias gradients
// db1[j] += sum_m dpre[m][j] and dzb[b][j] += sum_{m in image b} dpre[m][j].
__global__ void __launch_bounds__(256) coord_layer_bwd_kernel(CoordXform cx, const float* __restrict__ w1, const __half* __restrict__ dpre,
                                                              const float* __restrict__ inv_scale, float* __restrict__ dw1,
                                                              float* __restrict__ dxp, float* __restrict__ dzb, float* __restrict__ db1,
                                                              int H, int rows_per_cta) {
    extern __shared__ float s_cl[];
    float2* s_x = reinterpret_cast<float2*>(s_cl);                 // [kCoordRB]
    float* s_dx = s_cl + 2 * kCoordRB;                              // [kCoordRB][2]
    float* s_dw = s_dx + 2 * kCoordRB;                              // [H][2]
    const int cgs = H / 4, rpp = blockDim.x / cgs;
    const int cg = threadIdx.x % cgs, rs = threadIdx.x / cgs;
    const int c0 = cg * 4, lane = threadIdx.x & 31;
    const bool warp_rows = (cgs % 32) == 0;                         // every warp lies inside one row
    const float inv = __ldg(inv_scale);
    float* s_sum = s_dw + 2 * H;                                    // [H] column sums (only with dzb)
    float wx[4], wy[4], dwx[4], dwy[4], dsum[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) { wx[v] = w1[2 * (c0 + v)]; wy[v] = w1[2 * (c0 + v) + 1]; dwx[v] = 0.f; dwy[v] = 0.f; dsum[v] = 0.f; }
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) s_dw[i] = 0.f;
    if (dzb) for (int i = threadIdx.x; i < H; i += blockDim.x) s_sum[i] = 0.f;
    const long long m_begin = (long long)blockIdx.x * rows_per_cta;
    const long long m_end = min(m_begin + rows_per_cta, cx.M);
    for (long long m0 = m_begin; m0 < m_end; m0 += kCoordRB) {
        const int rows = static_cast<int>(min((long long)kCoordRB, m_end - m0));
        __syncthreads();
        if (threadIdx.x < kCoordRB) {
            float x0, x1;
            transformed_coord(cx, m0 + threadIdx.x, x0, x1);
            s_x[threadIdx.x] = make_float2(x0, x1);
            s_dx[2 * threadIdx.x] = 0.f;
            s_dx[2 * threadIdx.x + 1] = 0.f;
        }
        __syncthreads();
        for (int rr0 = rs; rr0 < rows; rr0 += 4 * rpp) {
            float4 g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int rr = rr0 + q * rpp;
                if (rr < rows) {
                    const uint2 t = *reinterpret_cast<const uint2*>(dpre + (m0 + rr) * H + c0);
                    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
                    g[q] = make_float4(a.x * inv, a.y * inv, b.x * inv, b.y * inv);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int rr = rr0 + q * rpp;
                if (rr < rows) {
                    const float2 x = s_x[rr];
                    const float gv[4] = {g[q].x, g[q].y, g[q].z, g[q].w};
                    float p0 = 0.f, p1 = 0.f;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        dwx[v] = fmaf(gv[v], x.x, dwx[v]);
                        dwy[v] = fmaf(gv[v], x.y, dwy[v]);
                        dsum[v] += gv[v];
                        p0 = fmaf(gv[v], wx[v], p0);
                        p1 = fmaf(gv[v], wy[v], p1);
                    }
                    if (warp_rows) {
                        p0 = warp_sum(p0);
                        p1 = warp_sum(p1);
                        if (lane == 0) { atomicAdd(s_dx + 2 * rr, p0); atomicAdd(s_dx + 2 * rr + 1, p1); }
                    } else {
                        atomicAdd(s_dx + 2 * rr, p0);
                        atomicAdd(s_dx + 2 * rr + 1, p1);
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < rows) {
            dxp[2 * (m0 + threadIdx.x)] = s_dx[2 * threadIdx.x];
            dxp[2 * (m0 + threadIdx.x) + 1] = s_dx[2 * threadIdx.x + 1];
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        atomicAdd(s_dw + 2 * (c0 + v), dwx[v]);
        atomicAdd(s_dw + 2 * (c0 + v) + 1, dwy[v]);
        if (dzb) atomicAdd(s_sum + c0 + v, dsum[v]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) atomicAdd(dw1 + i, s_dw[i]);
    if (dzb) {
        const long long b = m_begin / cx.N;
        for (int i = threadIdx.x; i < H; i += blockDim.x) {
            atomicAdd(dzb + b * H + i, s_sum[i]);
            atomicAdd(db1 + i, s_sum[i]);
        }
    }
}

// Row-streaming variant of the above for H % 8 == 0, H <= 512 (cfg2: H = 512): a warp owns a 256-column slice of whole
// rows (lane l holds columns 256 g + 8 l + [0,8), g = warp % ceil(H / 256)), kCoordU rows = kCoordU 16-byte loads in flight
// per lane at 3 CTAs per SM, the transformed coordinates of the CTA's rows computed once into shared memory, one warp
// reduction per row slice for dxp and per-lane register accumulators for dW1 / db1 / dzb.
// HBM-bound: one read of dpre (2 H bytes per row).
constexpr int kCoordMaxRows = 2048;
constexpr int kCoordU = 4;            // rows in flight per lane of coord_layer_bwd_rows_kernel
__global__ void __launch_bounds__(256) coord_layer_bwd_rows_kernel(CoordXform cx, const float* __restrict__ w1, const __half* __restrict__ dpre,
                                                                   const float* __restrict__ inv_scale, float* __restrict__ dw1,
                                                                   float* __restrict__ dxp, float* __restrict__ dzb, float* __restrict__ db1,
                                                                   int H, int rows_per_cta) {
    extern __shared__ float s_cl[];
    float2* s_x = reinterpret_cast<float2*>(s_cl);                 // [rows_per_cta] transformed coordinates
    float* s_dx = s_cl + 2 * rows_per_cta;                          // [rows_per_cta][2] dxp of the CTA's rows
    float* s_dw = s_dx + 2 * rows_per_cta;                          // [H][2]
    float* s_sum = s_dw + 2 * H;                                    // [H]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncg = (H + 255) / 256;                                // column groups (1 or 2)
    const int cg = warp % ncg, rw = warp / ncg, nrw = 8 / ncg;      // this warp: column group, row slot, row slots per pass
    const int c0 = cg * 256 + lane * 8;
    const bool col_ok = c0 < H;
    const float inv = __ldg(inv_scale);
    const long long m_begin = (long long)blockIdx.x * rows_per_cta;
    const int rows = static_cast<int>(min((long long)rows_per_cta, cx.M - m_begin));
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        float x0, x1;
        transformed_coord(cx, m_begin + i, x0, x1);
        s_x[i] = make_float2(x0, x1);
        s_dx[2 * i] = 0.f;
        s_dx[2 * i + 1] = 0.f;
    }
    for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) s_dw[i] = 0.f;
    float wx[8], wy[8], dwx[8], dwy[8], dsum[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        wx[v] = col_ok ? __ldg(w1 + 2 * (c0 + v)) : 0.f;
        wy[v] = col_ok ? __ldg(w1 + 2 * (c0 + v) + 1) : 0.f;
        dwx[v] = 0.f; dwy[v] = 0.f; dsum[v] = 0.f;
    }
    __syncthreads();
    const __half* base = dpre + m_begin * H + c0;
    for (int r0 = rw; r0 < rows; r0 += kCoordU * nrw) {                  // rows r0 + nrw * {0 .. kCoordU-1} of this warp: kCoordU 16-byte loads in flight per lane
        uint4 t[kCoordU];
#pragma unroll
        for (int q = 0; q < kCoordU; ++q) {
            const int r = r0 + nrw * q;
            t[q] = (r < rows && col_ok) ? __ldg(reinterpret_cast<const uint4*>(base + (long long)r * H)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < kCoordU; ++q) {
            const int r = r0 + nrw * q;
            if (r >= rows) break;                                    // uniform per warp
            const float2 x = s_x[r];
            float p0 = 0.f, p1 = 0.f;
            const uint32_t w[4] = {t[q].x, t[q].y, t[q].z, t[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                dwx[2 * e] = fmaf(f.x, x.x, dwx[2 * e]);         dwx[2 * e + 1] = fmaf(f.y, x.x, dwx[2 * e + 1]);
                dwy[2 * e] = fmaf(f.x, x.y, dwy[2 * e]);         dwy[2 * e + 1] = fmaf(f.y, x.y, dwy[2 * e + 1]);
                dsum[2 * e] += f.x;                              dsum[2 * e + 1] += f.y;
                p0 = fmaf(f.x, wx[2 * e], p0);                   p0 = fmaf(f.y, wx[2 * e + 1], p0);
                p1 = fmaf(f.x, wy[2 * e], p1);                   p1 = fmaf(f.y, wy[2 * e + 1], p1);
            }
            p0 = warp_sum(p0);
            p1 = warp_sum(p1);
            if (lane == 0) {
                if (ncg == 1) { s_dx[2 * r] = p0; s_dx[2 * r + 1] = p1; }
                else { atomicAdd(s_dx + 2 * r, p0); atomicAdd(s_dx + 2 * r + 1, p1); }
            }
        }
    }
    if (col_ok) {
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            atomicAdd(s_dw + 2 * (c0 + v), dwx[v] * inv);
            atomicAdd(s_dw + 2 * (c0 + v) + 1, dwy[v] * inv);
            atomicAdd(s_sum + c0 + v, dsum[v] * inv);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * rows; i += blockDim.x) dxp[2 * m_begin + i] = s_dx[i] * inv;
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) atomicAdd(dw1 + i, s_dw[i]);
    if (dzb) {
        const long long b = m_begin / cx.N;
        for (int i = threadIdx.x; i < H; i += blockDim.x) {
            atomicAdd(dzb + b * H + i, s_sum[i]);
            atomicAdd(db1 + i, s_sum[i]);
        }
    }
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
__global__ void __launch_bounds__(256) thin_fwd_kernel(const __half* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ y, long long M, int W, int T) {
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    for (int o = 0; o < T; ++o) {
        float acc = 0.f;
        for (int c = lane; c < W; c += 32) acc = fmaf(__half2float(a[m * W + c]), w[(long long)o * W + c], acc);
        acc = warp_sum(acc);
        if (lane == 0) y[m * T + o] = acc + bias[o];
    }
}

// dbias[o] = sum_r dbank[(r*O + o)][K]
__global__ void bank_bias_grad_kernel(const float* __restrict__ dbank, float* __restrict__ dbias, int O, int G, int kpad, int K) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    float acc = 0.f;
    for (int r = 0; r < G; ++r) acc += dbank[((long long)r * O + o) * kpad + K];
    dbias[o] = acc;
}

// out16[r][c] = fp16(in[r][c] * scales[2]) and colsum[c * cs_stride] += sum_r in[r][c]   (standalone GroupConv
// backward: the gradient arrives in fp32; the wgrad GEMM consumes scaled fp16, the conv1 bias gradient is the column sum)
__global__ void __launch_bounds__(256) rows_to_half_colsum_kernel(const float* __restrict__ in, __half* __restrict__ out16,
                                                                  const float* __restrict__ scales, float* __restrict__ colsum,
                                                                  long long cs_stride, long long R, int W, int rows_per_cta) {
    extern __shared__ float sm_cs[];
    for (int c = threadIdx.x; c < W; c += blockDim.x) sm_cs[c] = 0.f;
    __syncthreads();
    const float sc = __ldg(scales + 2);
    const int cpr = W / 2;                          // column pairs per row
    const int rpp = blockDim.x / cpr > 0 ? blockDim.x / cpr : 1;
    const int cp = threadIdx.x % cpr, rr = threadIdx.x / cpr;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    float s0 = 0.f, s1 = 0.f;
    if (rr < rpp) {
        for (long long r = r0 + rr; r < r0 + rows_per_cta && r < R; r += rpp) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(in + r * W) + cp);
            reinterpret_cast<__half2*>(out16 + r * W)[cp] = __floats2half2_rn(v.x * sc, v.y * sc);
            s0 += v.x; s1 += v.y;
        }
        atomicAdd(&sm_cs[2 * cp], s0);
        atomicAdd(&sm_cs[2 * cp + 1], s1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) atomicAdd(colsum + c * cs_stride, sm_cs[c]);
}

}  // namespace tvae
