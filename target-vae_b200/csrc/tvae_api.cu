// C-ABI entry points of libtvae_b200.so (declared in include/tvae_b200.h).
#include "../../include/tvae_b200.h"

#include "conv_policies.cuh"
#include "conv2_policies.cuh"
#include "conv_f16_policies.cuh"
#include "ctf_policies.cuh"
#include "gen_policies.cuh"
#include "gen_pair_policies.cuh"
#include "launch.cuh"
#include "simt_gen.cuh"
#include "simt_kernels.cuh"
#include "attn_kernels.cuh"
#include "enc_bwd_fused.cuh"
#include "refine.cuh"
#include "module_kernels.cuh"

using namespace tvae;

namespace {

inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
inline int blocks_for(long long n, int threads, int cap = 148 * 8) {
    long long b = (n + threads - 1) / threads;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

// event-times the launches issued while it is alive (tvae_profile_enable; no-op otherwise)
struct Timed {
    int slot;
    cudaStream_t st;
    Timed(const char* name, cudaStream_t s) : slot(g_timer.begin(name, s)), st(s) {}
    ~Timed() { g_timer.end(slot, st); }
};

RotTable make_rot_table(int G) {
    // theta accumulated in double exactly like models.py:181-195, cos/sin rounded to fp32 (models.py:186-190)
    RotTable t{};
    const double d_theta = 2.0 * 3.14159265358979323846 / G;
    double theta = 0.0;
    for (int i = 0; i < G && i < kMaxG; ++i) {
        t.cs[i] = static_cast<float>(cos(theta));
        t.sn[i] = static_cast<float>(sin(theta));
        theta += d_theta;
    }
    return t;
}

ConvGeom make_geom(const tvae_enc_shape* s) {
    ConvGeom g{};
    g.B = s->B; g.C = s->C; g.n = s->n; g.k = s->k; g.p = s->p; g.G = s->G; g.O = s->O;
    g.d = s->n + 2 * s->p - s->k + 1;
    g.P = g.d * g.d;
    g.K = s->C * s->k * s->k;
    g.kpad = s->kpad;
    return g;
}

int check_enc_shape(const tvae_enc_shape* s) {
    TVAE_REQUIRE(s->B > 0 && s->C > 0 && s->n > 0 && s->k > 0 && s->p >= 0, "encoder: non-positive dimension");
    TVAE_REQUIRE(s->G >= 1 && s->G <= kMaxG, "encoder: groupconv must be in 1..16");
    TVAE_REQUIRE(s->act == TVAE_ACT_LEAKYRELU || s->act == TVAE_ACT_TANH, "encoder: act must be TVAE_ACT_LEAKYRELU or TVAE_ACT_TANH");
    TVAE_REQUIRE(s->O % 32 == 0 && s->O >= 32 && s->O <= 256, "encoder: kernel count must be a multiple of 32 in [32,256]");
    TVAE_REQUIRE(s->z >= 1 && 3 + 2 * s->z <= kMaxNH, "encoder: latent dim out of range");
    TVAE_REQUIRE(s->n + 2 * s->p - s->k + 1 >= 1, "encoder: kernel larger than padded image");
    TVAE_REQUIRE(s->kpad == tvae_bank_pitch(s->C, s->k), "encoder: kpad must come from tvae_bank_pitch");
    return 0;
}

// warp-MMA thin backward (fp16 activations, T <= 16, W % 128 == 0)
template <bool PLANAR, int TT = 1>
int launch_thin_bwd_mma(ThinBwdParams& p, int G, cudaStream_t st) {
    const int slices = p.W / 128;
    long long rows = (p.M * slices + 148LL * 8 - 1) / (148LL * 8);
    rows = (rows + kThinRB - 1) / kThinRB * kThinRB;
    p.rows_per_cta = static_cast<int>(rows);
    const int n_red = (p.T + 1) * 128 + p.T;
    const size_t sm = sizeof(float) * ((n_red + 3) & ~3) + TT * 16 * 32 * 8 + kThinBufs * (((kThinRB * p.T + 3) & ~3) * 4 + kThinRB * kThinPitch * 2);
    if (sm > 48 * 1024) {
        TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&thin_bwd_mma_kernel<PLANAR, TT>), 100 * 1024));
    }
    TVAE_REQUIRE(sm <= 100 * 1024, "thin backward (mma): shared memory");
    ++g_launch_count;
    thin_bwd_mma_kernel<PLANAR, TT><<<dim3(cdiv(p.M, rows), slices), 256, sm, st>>>(p, G);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// generator output layer: n_out <= 4 thin outputs, row-major dt, fp16 activation (thin_bwd_stream_kernel)
inline int launch_thin_bwd_stream(ThinBwdParams& p, cudaStream_t st) {
    const int cgs = p.W / 8, slots = 256 / cgs;
    const int per_pass = slots * 4;
    long long rows = (p.M + 148LL * 8 - 1) / (148LL * 8);
    rows = (rows + per_pass - 1) / per_pass * per_pass;
    p.rows_per_cta = static_cast<int>(rows);
    const size_t sm = sizeof(float) * ((p.T + 1) * p.W + p.T);
    TVAE_REQUIRE(sm <= 48 * 1024, "thin backward (stream): shared memory");
    const dim3 grid(static_cast<unsigned>(cdiv(p.M, rows)));
    ++g_launch_count;
    switch (p.T) {
        case 1: thin_bwd_stream_kernel<1><<<grid, cgs * slots, sm, st>>>(p); break;
        case 2: thin_bwd_stream_kernel<2><<<grid, cgs * slots, sm, st>>>(p); break;
        case 3: thin_bwd_stream_kernel<3><<<grid, cgs * slots, sm, st>>>(p); break;
        default: thin_bwd_stream_kernel<4><<<grid, cgs * slots, sm, st>>>(p); break;
    }
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int TMAX, int VEC, bool PLANAR, bool H16 = false>
int launch_thin_bwd(ThinBwdParams& p, int G, cudaStream_t st) {
    if (H16 && !PLANAR && p.T >= 1 && p.T <= 4 && p.W % 8 == 0 && p.W / 8 <= 256 && g_dev_knob[4] == 0) return launch_thin_bwd_stream(p, st);
    if (H16 && p.T <= 16 && p.W % 128 == 0) return launch_thin_bwd_mma<PLANAR, 1>(p, G, st);
    if (H16 && p.T <= 24 && p.W % 128 == 0) return launch_thin_bwd_mma<PLANAR, 2>(p, G, st);
    TVAE_REQUIRE(p.W % VEC == 0 && p.W / VEC <= 256 && p.T <= TMAX, "thin backward: unsupported width");
    const int cgs = p.W / VEC;
    const int rpp = 256 / cgs > 0 ? 256 / cgs : 1;
    long long rows = (p.M + 148LL * 8 - 1) / (148LL * 8);
    rows = (rows + kThinRB - 1) / kThinRB * kThinRB;
    p.rows_per_cta = static_cast<int>(rows);
    const size_t sm = sizeof(float) * (kThinRB * p.T + (p.T + 1) * p.W + p.T);
    TVAE_REQUIRE(sm <= 48 * 1024, "thin backward: shared memory");
    ++g_launch_count;
    thin_bwd_kernel<TMAX, VEC, PLANAR, H16><<<cdiv(p.M, rows), cgs * rpp, sm, st>>>(p, G);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <class P>
int launch_split_tn(typename P::Params& p, int out_tiles, int chunks_total, int align, int extra, cudaStream_t st) {
    int splits = cdiv(2 * sm_count(), out_tiles);
    const int min_chunks = 16;
    if (splits > cdiv(chunks_total, min_chunks)) splits = cdiv(chunks_total, min_chunks);
    if (splits < 1) splits = 1;
    int cps = cdiv(chunks_total, splits);
    if (align > 1) cps = cdiv(cps, align) * align;   // whole images per split where possible
    p.chunks_total = chunks_total;
    p.chunks_per_split = cps;
    p.splits = cdiv(chunks_total, cps);
    p.num_tiles = out_tiles * p.splits;
    return launch_gemm<P>(p, extra, st);
}

}  // namespace

extern "C" {

const char* tvae_last_error(void) { return g_last_error.c_str(); }
int tvae_version(void) { return 102; }   // 102: tvae_gen_fwd_args.mask_bits; 101: act in the shape structs, rotation-pooling fields in the encoder arguments
long long tvae_launch_count(void) { return g_launch_count.load(); }

// per-kernel timing of the tensor-core GEMM launches (CUDA events on the launching stream)
void tvae_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_timer.mu);
    g_timer.enabled.store(on != 0);
    if (on) g_timer.n_events = 0;
}
// Synchronises, then fills up to `cap` entries: names[i] (static strings), total_ms[i], launches[i]. Returns count.
int tvae_profile_collect(const char** names, float* total_ms, int* launches, int cap) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_timer.mu);
    int n = g_timer.n_names < cap ? g_timer.n_names : cap;
    for (int i = 0; i < n; ++i) { names[i] = g_timer.names[i]; total_ms[i] = 0.f; launches[i] = 0; }
    for (int e = 0; e < g_timer.n_events; ++e) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_timer.start[e], g_timer.stop[e]) == cudaSuccess && g_timer.name_of[e] < n) {
            total_ms[g_timer.name_of[e]] += ms;
            launches[g_timer.name_of[e]] += 1;
        }
    }
    g_timer.n_events = 0;
    return n;
}

int tvae_bank_pitch(int C, int k) {
    const int K = C * k * k;
    return (K / 32 + 1) * 32;   // always leaves >= 1 spare column for the bias-gradient ones column
}

int tvae_bank16_pitch(int C, int k) {
    const int K = C * k * k;
    return (K + 63) / 64 * 64;   // whole 64-tap stage rows of the fp16 bank
}

// ================================================================================ filter bank
int tvae_filter_bank_fwd(const tvae_enc_shape* s, const float* weight, void* bank, void* stream) {
    int rc = check_enc_shape(s);
    if (rc) return rc;
    const int kpad16 = tvae_bank16_pitch(s->C, s->k);
    const long long total = (long long)s->G * s->O * kpad16;
    Timed tm("filter_bank_fwd", S(stream));
    ++g_launch_count;
    const size_t plane = sizeof(float) * s->k * s->k;
    if (plane <= 48 * 1024)
        filter_bank_fwd_plane_kernel<<<dim3(s->O * s->C, s->G), 256, plane, S(stream)>>>(weight, static_cast<__half*>(bank), s->O, s->C, s->k, kpad16, make_rot_table(s->G));
    else
        filter_bank_fwd_kernel<<<blocks_for(total, 256), 256, 0, S(stream)>>>(weight, static_cast<__half*>(bank), s->O, s->C, s->k, s->G, kpad16, make_rot_table(s->G));
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tvae_filter_bank_bwd(const tvae_enc_shape* s, const float* dbank, float* dweight, float* dbias, void* stream) {
    int rc = check_enc_shape(s);
    if (rc) return rc;
    const int K = s->C * s->k * s->k;
    Timed tm("filter_bank_bwd", S(stream));
    ++g_launch_count;
    if (g_dev_knob[5] == 0) {
        filter_bank_bwd_gather_kernel<<<dim3(s->O * s->C, cdiv(s->k * s->k, 256)), 256, 0, S(stream)>>>(dbank, s->kpad, dweight, s->O, s->C, s->k, s->G, make_rot_table(s->G));
    } else {
        TVAE_CHECK_CUDA(cudaMemsetAsync(dweight, 0, sizeof(float) * s->O * K, S(stream)));
        const long long total = (long long)s->G * s->O * K;
        filter_bank_bwd_kernel<<<blocks_for(total, 256), 256, 0, S(stream)>>>(dbank, s->kpad, dweight, s->O, s->C, s->k, s->G, make_rot_table(s->G));
    }
    if (dbias) { ++g_launch_count; bank_bias_grad_kernel<<<cdiv(s->O, 128), 128, 0, S(stream)>>>(dbank, dbias, s->O, s->G, s->kpad, K); }
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ================================================================================ encoder
}  // extern "C"

namespace {
// conv1: implicit GEMM with the im2col operand generated on chip (CTA-pair kernel, tc_gemm2.cuh), fp16 operands
Slab16Geom make_slab16(const ConvGeom& g, int rows_max, int channels) {
    Slab16Geom sg{};
    sg.Wp = g.n + 2 * g.p;
    // even pitch (a quad of taps keeps its 4-byte alignment class on every slab row) with pitch - d = 0 or +-1 (mod 64):
    // the 32 consecutive output cells of a warp keep (almost) consecutive words across image-row wraps
    int pitch = sg.Wp + (sg.Wp & 1);
    for (;; pitch += 2) {
        const int m = (((pitch - g.d) % 64) + 64) % 64;
        if (m == 0 || m == 1 || m == 63) break;
    }
    sg.pitch = pitch;
    sg.rows_max = rows_max < sg.Wp ? rows_max : sg.Wp;
    int words = (channels * sg.rows_max * sg.pitch + 1) / 2 + 4;
    words += ((16 - words % 32) + 32) % 32;          // copy 1 starts 16 banks after copy 0
    sg.copy_words = words;
    return sg;
}
int conv1_forward(const ConvGeom& g, const float* y, const void* bank16, const float* bias, float* x1, __half* x1h, int act, cudaStream_t st) {
    Conv1FwdHParams p{};
    const int N = g.G * g.O;
    const int kpad16 = tvae_bank16_pitch(g.C, g.k);
    int rc;
    if ((rc = make_tmap_2d_h(&p.tmB, bank16, N, kpad16, kpad16, 128))) return rc;
    p.g = g;
    p.y = y; p.bias = bias; p.x1 = x1; p.x1h = x1h; p.act = act;
    p.n_passes = cdiv(N, kAcc * kAccN);
    p.tiles_per_image = cdiv(g.P, kBM);
    p.m_tiles = g.B * p.tiles_per_image;
    p.k_chunks = cdiv(g.K, kBK16);
    p.m_pairs = cdiv(p.m_tiles, 2);
    const int pairs_dev = sm_count() / 2;
    p.pairs = p.m_pairs < pairs_dev ? p.m_pairs : pairs_dev;
    p.num_tiles = p.pairs * cdiv(p.m_pairs, p.pairs) * p.n_passes;   // (m-pair, pass) grid padded to whole rounds of the pairs
    p.quad = (g.k % 4 == 0) ? 1 : 0;
    p.skip = ((g.k * g.k) % kBK16 == 0) ? 1 : 0;
    p.chunks_per_channel = p.skip ? (g.k * g.k) / kBK16 : p.k_chunks;
    // multi-channel images whose channels are chunk-aligned keep one channel in the slab at a time (cfg3: three
    // 64x64-tap channels would need 101 KB of slab and leave room for a single pipeline stage)
    p.per_channel = (g.C > 1 && p.skip) ? 1 : 0;
    p.sg = make_slab16(g, (kBM - 1) / g.d + 2 + g.k - 1, p.per_channel ? 1 : g.C);
    p.tab_entries = (p.per_channel ? p.chunks_per_channel : p.k_chunks) * (p.quad ? kBK16 / 4 : kBK16);
    int extra = p.tab_entries * 4 + 2 * p.sg.copy_words * 4;
    p.bias_off = extra;
    extra += ((g.O * 4 + 1023) / 1024) * 1024;
    extra = (extra + 1023) / 1024 * 1024;            // staging buffers need the 128 B swizzle's 1024-byte alignment
    p.stage_off = extra;
    p.tma_store = (x1h != nullptr && g.O % 64 == 0) ? 1 : 0;
    if (p.tma_store) {
        if ((rc = make_tmap_3d_store_h(&p.tmX, x1h, (uint64_t)g.B * g.G, g.P, g.O, kBM))) return rc;
        extra += 2 * kStoreBlockBytes;
    }
    if (act == kActTanh) return launch_gemm2<Conv1FwdHT<true>>(p, extra, st, p.pairs);
    return launch_gemm2<Conv1FwdH>(p, extra, st, p.pairs);
}
// conv1 weight gradient w.r.t. the rotated bank (dbank pre-zeroed); dx1 is fp16 [(b,r,pos)][O] times 1 / *acc_scale
int conv1_wgrad(const ConvGeom& g, const float* y, const void* dx1_16, const float* acc_scale, float* dbank, cudaStream_t st) {
    Conv1WgradHParams p{};
    const int N = g.G * g.O;
    const long long R = (long long)g.B * g.G * g.P;
    int rc;
    const bool b128 = g.O % 64 == 0;     // dX1 staged in 64-column blocks / 128 B swizzle (full 128-byte operand rows)
    if (b128) {
        p.nb = (g.O % 128 == 0) ? 2 : 1;                              // 128 accumulator columns = 2 / nb boxes of nb 64-column blocks
        if ((rc = make_tmap_3d_mn128_h(&p.tmQ, dx1_16, R, g.O, g.O, kBK16, p.nb))) return rc;
    } else {
        const int oblocks = g.O / 32;
        p.nb = (oblocks % 4 == 0) ? 4 : (oblocks % 2 == 0 ? 2 : 1);   // 128 accumulator columns = 4 / nb boxes of nb o-blocks
        if ((rc = make_tmap_3d_mn_h(&p.tmQ, dx1_16, R, g.O, g.O, kBK16, p.nb))) return rc;
    }
    p.g = g; p.y = y; p.dbank = dbank; p.acc_scale = acc_scale;
    p.m_tiles = cdiv(g.K, kBM);
    p.m_pairs = cdiv(p.m_tiles, 2);
    p.n_passes = cdiv(N, kAcc * kAccN);
    p.chunks_per_image = cdiv(g.P, kBK16);
    p.chunks_total = g.B * p.chunks_per_image;
    p.quad = (g.k % 4 == 0) ? 1 : 0;
    // slab: one channel and the few padded rows a 128-wide kk tile touches, or whole padded channels when a tile can
    // straddle channels
    const bool straddles = g.C > 1 && (g.k * g.k) % kBM != 0;
    const int nc_max = straddles ? (g.C < kBM / (g.k * g.k) + 2 ? g.C : kBM / (g.k * g.k) + 2) : 1;
    p.sg = make_slab16(g, straddles ? g.n + 2 * g.p : g.d + (kBM - 1) / g.k + 1, nc_max);
    int extra = kBM * 4 + 2 * p.sg.copy_words * 4;
    // epilogue: TMA reduce-adds of fp32 staging tiles into dbank [N][K] (pitch kpad); per-element atomics if dbank is not
    // 16-byte aligned
    rc = make_tmap_2d_f32_reduce(&p.tmD, dbank, N, g.K, g.kpad);
    if (rc < 0) return rc;
    p.tma_reduce = rc == 0;
    int nbuf = 0;
    if (p.tma_reduce) {
        extra = (extra + 127) / 128 * 128;
        p.stage_off = extra;
        // two staging tiles unless the second costs a pipeline stage (then one; if even one does: atomics)
        const int s0 = pick_stages2(extra);
        nbuf = pick_stages2(extra + 2 * kStoreBlockBytes) == s0 ? 2 : (pick_stages2(extra + kStoreBlockBytes) == s0 ? 1 : 0);
        extra += nbuf * kStoreBlockBytes;
        p.tma_reduce = nbuf > 0;
    }
    // Reduction splits: minimise waves x (chunks per split + epilogue).  The epilogue adds a 256 x 512 fp32 tile per pair into
    // dbank; its cost grows with the number of pairs that add into the SAME output tile at the same time (L2 serialises
    // reduces on one line): measured 16 k clocks per tile with 2.3 concurrent pairs per output tile (cfg2), 49 k with 9.3
    // (cfg1).  Few, long splits win (clock64 probe, profiles/r02_pair_kernel_probe.md: 9 instead of 46 splits = -8 % at
    // cfg2, -10 % at cfg4 / cfg5, -37 % at cfg1): the out-tiles of one split walk the same dX1 rows in near lockstep and
    // share them through L2 even when a split's slice is far larger than L2, so round 1's L2-sized lower bound on the
    // split count only bought more epilogues.
    const int out_tiles = p.m_pairs * p.n_passes;
    const int pairs_dev = sm_count() / 2;
    const double inflight = out_tiles < pairs_dev ? static_cast<double>(pairs_dev) / out_tiles : 1.0;
    const double epi_clocks = 6000.0 + 5000.0 * inflight;
    const int s_cap = p.chunks_total / 32 > 1 ? p.chunks_total / 32 : 1;      // keep >= 32 chunks (2048 positions) per split
    // With the zero-padding skip a tile's chunk count depends on its kk rows (36 .. 65 live output rows of 65 at cfg2), and a
    // pair's strided tile list only averages that out over enough tiles: 4 tiles per pair measured 6 % slower than 10
    // although the mean per-pair time was lower.  Without the skip (cfg1) all tiles are equal and one wave is best.
    const bool uneven = (g.k * g.k) % kBK16 == 0;
    const int min_waves = uneven ? 8 : 1;
    double best = 1e30;
    int best_cps = p.chunks_total;
    for (int s = 1; s <= s_cap && s <= 4 * pairs_dev; ++s) {
        const int cps = cdiv(p.chunks_total, s);
        const int s_eff = cdiv(p.chunks_total, cps);
        const int waves = cdiv((long long)out_tiles * s_eff, pairs_dev);
        if (waves < min_waves && s < s_cap) continue;
        const double cost = waves * (cps * 1024.0 + epi_clocks);
        if (cost < best) { best = cost; best_cps = cps; }
    }
    if (g_dev_knob[0] > 0) best_cps = cdiv(p.chunks_total, g_dev_knob[0]);
    p.chunks_per_split = best_cps;
    p.splits = cdiv(p.chunks_total, best_cps);
    p.num_tiles = out_tiles * p.splits;
    p.skip = 1;
    if (nbuf == 1) return b128 ? launch_gemm2<Conv1WgradH128_1>(p, extra, st) : launch_gemm2<Conv1WgradH_1>(p, extra, st);
    return b128 ? launch_gemm2<Conv1WgradH128>(p, extra, st) : launch_gemm2<Conv1WgradH>(p, extra, st);
}
// rotation pooling between conv1 and conv2 (simt_kernels.cuh: rot_pool_fwd_kernel / rot_pool_bwd_kernel)
inline int rot_pool_rows_per_cta(long long rows, int O) {
    const int rpp = 256 / (O / 8);
    long long per = (rows + 8LL * sm_count() - 1) / (8LL * sm_count());   // 8 resident CTAs of 256 threads per SM = full occupancy
    per = (per + rpp - 1) / rpp * rpp;
    return static_cast<int>(per < rpp ? rpp : per);
}
inline int rot_pool_forward(const ConvGeom& g, const __half* x1, const float* fc_w, const float* fc_b, __half* xp, cudaStream_t st) {
    TVAE_REQUIRE(g.O % 8 == 0 && g.O / 8 <= 256, "rotation pooling: kernel count must be a multiple of 8");
    RotPoolParams p{};
    p.x1 = x1; p.fc_w = fc_w; p.fc_b = fc_b; p.xp = xp;
    p.B = g.B; p.G = g.G; p.P = g.P; p.O = g.O;
    const long long rows = (long long)g.B * g.P;
    p.rows_per_cta = rot_pool_rows_per_cta(rows, g.O);
    ++g_launch_count; rot_pool_fwd_kernel<<<cdiv(rows, p.rows_per_cta), 256, 0, st>>>(p);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}
// conv1 bias gradient slot: column K of dbank row (r = 0, o)  (summed over r by tvae_filter_bank_bwd)
inline float* bias_grad_slot(const ConvGeom& g, float* dbank) { return dbank + g.K; }
}  // namespace

extern "C" {

// Fraction of the dense K-chunk count the conv1 kernels actually execute (same arithmetic as Conv1FwdH::tile_info /
// Conv1WgradH::tile_info: chunks that only meet zero padding are skipped).
double tvae_conv1_executed_fraction(const tvae_enc_shape* s, int wgrad) {
    if (check_enc_shape(s)) return -1.0;
    const ConvGeom g = make_geom(s);
    auto imin = [](int a, int b) { return a < b ? a : b; };
    auto imax = [](int a, int b) { return a > b ? a : b; };
    double live = 0.0, dense = 0.0;
    if (!wgrad) {
        const int tiles_per_image = cdiv(g.P, kBM), m_tiles = g.B * tiles_per_image, m_pairs = cdiv(m_tiles, 2);
        const int k_chunks = cdiv(g.K, kBK16);
        const bool skip = (g.k * g.k) % kBK16 == 0;
        for (int mp = 0; mp < m_pairs; ++mp) {
            dense += k_chunks;
            if (!skip) { live += k_chunks; continue; }
            int v_lo = g.k, v_hi = 0;
            for (int h = 0; h < 2; ++h) {
                const int t = 2 * mp + h;
                if (t >= m_tiles) continue;
                const int pos0 = (t % tiles_per_image) * kBM;
                const int i_first = pos0 / g.d, i_last = imin(pos0 + kBM - 1, g.P - 1) / g.d;
                v_lo = imin(v_lo, imax(0, g.p - i_last));
                v_hi = imax(v_hi, imin(g.k, g.p - i_first + g.n));
            }
            const int lo = (v_lo * g.k) / kBK16, hi = (v_hi * g.k + kBK16 - 1) / kBK16;
            live += hi > lo ? g.C * (hi - lo) : 0;
        }
    } else {
        const int m_tiles = cdiv(g.K, kBM), m_pairs = cdiv(m_tiles, 2), cpi = cdiv(g.P, kBK16);
        for (int mp = 0; mp < m_pairs; ++mp) {
            dense += cpi;
            const int kk0 = 2 * mp * kBM, kk1 = imin(kk0 + 2 * kBM, g.K) - 1;
            int cnt = cpi;
            {
                const int c0 = kk0 / (g.k * g.k), c1 = kk1 / (g.k * g.k);
                const int v0 = (kk0 - c0 * g.k * g.k) / g.k, v1 = (kk1 - c1 * g.k * g.k) / g.k;
                if (c0 == c1) {
                    const int i_lo = imax(0, g.p - v1), i_hi = imin(g.d - 1, g.p - v0 + g.n - 1);
                    cnt = i_hi >= i_lo ? ((i_hi + 1) * g.d + kBK16 - 1) / kBK16 - (i_lo * g.d) / kBK16 : 0;
                }
            }
            live += cnt;
        }
        // dense count of the contract has K rows, the kernel's tile grid has 2 * m_pairs * 128
        dense *= static_cast<double>(g.K) / (2.0 * m_pairs * kBM);
    }
    return dense > 0 ? live / dense : 1.0;
}

int tvae_groupconv_fwd(const tvae_enc_shape* s, const float* y, const void* bank, const float* bias, float* out, void* stream) {
    int rc = check_enc_shape(s);
    if (rc) return rc;
    return conv1_forward(make_geom(s), y, bank, bias, out, nullptr, 0, S(stream));
}

int tvae_groupconv_wgrad(const tvae_enc_shape* s, const float* y, const float* dout, void* dout16, float* scales, float* dbank, void* stream) {
    int rc = check_enc_shape(s);
    if (rc) return rc;
    const ConvGeom g = make_geom(s);
    cudaStream_t st = S(stream);
    TVAE_CHECK_CUDA(cudaMemsetAsync(dbank, 0, sizeof(float) * g.G * g.O * g.kpad, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(scales, 0, sizeof(float) * 8, st));
    const long long R = (long long)g.B * g.G * g.P;
    ++g_launch_count; absmax_kernel<<<blocks_for(R * g.O, 256), 256, 0, st>>>(dout, R * g.O, scales + 7);
    ++g_launch_count; single_scale_kernel<<<1, 1, 0, st>>>(scales + 7, scales);
    const int rows_per_cta = static_cast<int>((R + 148LL * 8 - 1) / (148LL * 8));
    ++g_launch_count;
    rows_to_half_colsum_kernel<<<cdiv(R, rows_per_cta), 256, g.O * sizeof(float), st>>>(
        dout, static_cast<__half*>(dout16), scales, bias_grad_slot(g, dbank), g.kpad, R, g.O, rows_per_cta);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return conv1_wgrad(g, y, dout16, scales + 3, dbank, st);
}

int tvae_encoder_fwd(const tvae_enc_shape* s, const tvae_enc_fwd_args* a, void* stream) {
    int rc = check_enc_shape(s);
    if (rc) return rc;
    ConvGeom g = make_geom(s);
    const int NH = 3 + 2 * s->z;
    long long R = (long long)g.B * g.G * g.P;
    cudaStream_t st = S(stream);
    const int act = s->act == TVAE_ACT_TANH ? kActTanh : 1;
    if ((rc = conv1_forward(g, a->y, a->bank, a->conv1_bias, nullptr, static_cast<__half*>(a->x1), act, st))) return rc;
    const void* x_in = a->x1;
    if (a->fc_w) {
        // ---- rotation pooling (attention/unimodal encoder, groupconv > 0): conv2 and the heads see one rotation slot
        TVAE_REQUIRE(a->fc_b && a->xp, "encoder: rotation pooling needs fc_b and xp");
        if ((rc = rot_pool_forward(g, static_cast<const __half*>(a->x1), a->fc_w, a->fc_b, static_cast<__half*>(a->xp), st))) return rc;
        x_in = a->xp;
        R = (long long)g.B * g.P;
        g.G = 1;
    }
    // ---- conv2 (1x1x1) + heads
    {
        ++g_launch_count; to_half_kernel<<<blocks_for((long long)g.O * g.O, 256), 256, 0, st>>>(a->w2, static_cast<__half*>(a->w2_h), (long long)g.O * g.O);
        if (g.O == 128 && NH <= 32) {
            // heads on the tensor core, h through staged TMA stores (Conv2HeadsTC)
            Conv2HeadsTCParams q{};
            if ((rc = make_tmap_2d_h(&q.tmA, x_in, R, g.O, g.O, kBM))) return rc;
            if ((rc = make_tmap_2d_h(&q.tmB, a->w2_h, g.O, g.O, g.O, 128))) return rc;
            q.store_h = a->h != nullptr;
            if (q.store_h && (rc = make_tmap_2d_h(&q.tmH, a->h, R, g.O, g.O, kBM))) return rc;
            q.R = R; q.O = g.O; q.NH = NH; q.NHpad = NH <= 16 ? 16 : 32; q.G = g.G; q.P = g.P;
            q.k_chunks = 2;
            q.num_tiles = cdiv(R, kBM);
            q.b2 = a->b2; q.wh = a->wh; q.bh = a->bh; q.head_add = a->head_add; q.heads = a->heads;
            q.act = act;
            if (act == kActTanh) return launch_gemm<Conv2HeadsTCT<true>>(q, Conv2HeadsTC::kExtraBytes, st);
            return launch_gemm<Conv2HeadsTC>(q, Conv2HeadsTC::kExtraBytes, st);
        }
        Conv2HeadsParams p{};
        const bool wide = g.O > 128;
        const int BN = wide ? 256 : 128;
        if ((rc = make_tmap_2d_h(&p.tmA, x_in, R, g.O, g.O, kBM))) return rc;
        if ((rc = make_tmap_2d_h(&p.tmB, a->w2_h, g.O, g.O, g.O, BN))) return rc;
        p.R = R; p.O = g.O; p.NH = NH; p.G = g.G; p.P = g.P;
        p.k_chunks = cdiv(g.O, kBKh);
        p.num_tiles = cdiv(R, kBM);
        p.b2 = a->b2; p.wh = a->wh; p.bh = a->bh; p.head_add = a->head_add; p.h = static_cast<__half*>(a->h); p.heads = a->heads;
        p.act = act;
        const int extra = (NH * g.O + g.O) * static_cast<int>(sizeof(float));
        if (act == kActTanh) {
            if (wide) rc = launch_gemm<Conv2Heads<256, kMaxNH, true>>(p, extra, st);
            else if (NH <= 8) rc = launch_gemm<Conv2Heads<128, 8, true>>(p, extra, st);
            else if (NH <= 20) rc = launch_gemm<Conv2Heads<128, 20, true>>(p, extra, st);
            else rc = launch_gemm<Conv2Heads<128, kMaxNH, true>>(p, extra, st);
        }
        else if (wide) rc = launch_gemm<Conv2Heads<256, kMaxNH>>(p, extra, st);
        else if (NH <= 8) rc = launch_gemm<Conv2Heads<128, 8>>(p, extra, st);
        else if (NH <= 20) rc = launch_gemm<Conv2Heads<128, 20>>(p, extra, st);
        else rc = launch_gemm<Conv2Heads<128, kMaxNH>>(p, extra, st);
        if (rc) return rc;
    }
    return 0;
}

int tvae_encoder_bwd(const tvae_enc_shape* s, const tvae_enc_bwd_args* a, void* stream) {
    int rc = check_enc_shape(s);
    if (rc) return rc;
    const ConvGeom g = make_geom(s);
    const int NH = 3 + 2 * s->z;
    // rotation pooling (attention/unimodal encoder, groupconv > 0): everything after conv1 sees ONE rotation slot
    const bool pooled = a->fc_w != nullptr;
    TVAE_REQUIRE(!pooled || (a->xp && a->dxp16 && a->dfc_w && a->dfc_b), "encoder backward: rotation pooling needs xp, dxp16, dfc_w, dfc_b");
    const int G2 = pooled ? 1 : g.G;
    const long long R = (long long)g.B * G2 * g.P;          // rows of h / dhpre / the conv2 input
    cudaStream_t st = S(stream);
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->dwh, 0, sizeof(float) * NH * g.O, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->dbh, 0, sizeof(float) * NH, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->db2, 0, sizeof(float) * g.O, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->dw2, 0, sizeof(float) * g.O * g.O, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->dbank, 0, sizeof(float) * g.G * g.O * g.kpad, st));
    // ---- power-of-two scales that keep the fp16 gradient operands in range (computed on the device, no sync)
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->scales, 0, sizeof(float) * 8, st));
    {
        const long long n = (long long)g.B * NH * G2 * g.P;
        ++g_launch_count; absmax_kernel<<<blocks_for(n, 256), 256, 0, st>>>(a->d_heads, n, a->scales + 7);
        ++g_launch_count; enc_bwd_scales_kernel<<<1, 256, 0, st>>>(a->scales + 7, a->wh, NH, a->w2, g.O, a->scales);
    }
    // ---- heads backward: dhpre = (d_heads . Wh) * act'(h), stored fp16 * s1; dWh, dbh, db2
    // (the two fused kernels take the LeakyReLU derivative from sign bits of the staged tiles; tanh needs the values and
    // runs on the general kernels)
    const bool tanh_act = s->act == TVAE_ACT_TANH;
    if (g.O == 128 && NH <= 32 && !tanh_act) {
        // tensor-core streaming kernel (enc_bwd_fused.cuh): one read of h and d_heads, one write of dhpre
        EncHeadsBwdParams q{};
        if ((rc = make_tmap_2d_h(&q.tmH, a->h, R, 128, 128, kBM))) return rc;
        if ((rc = make_tmap_2d_h(&q.tmC, a->dhpre, R, 128, 128, kBM))) return rc;
        q.R = R; q.num_tiles = static_cast<int>(cdiv(R, kBM)); q.NH = NH; q.G = G2; q.P = g.P;
        q.d_heads = a->d_heads; q.wh = a->wh; q.store_scale = a->scales + 0; q.in_scale = a->scales + 6;
        q.dwh = a->dwh; q.dbh = a->dbh; q.db2 = a->db2;
        const int grid = q.num_tiles < sm_count() ? q.num_tiles : sm_count();
        TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&enc_heads_bwd_kernel<16>), kHbSmemBytes));
        TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&enc_heads_bwd_kernel<32>), kHbSmemBytes));
        ++g_launch_count;
        const int tslot = g_timer.begin("enc_heads_bwd", st);
        if (NH <= 16) enc_heads_bwd_kernel<16><<<grid, kHbThreads, kHbSmemBytes, st>>>(q);
        else enc_heads_bwd_kernel<32><<<grid, kHbThreads, kHbSmemBytes, st>>>(q);
        g_timer.end(tslot, st);
        TVAE_CHECK_CUDA(cudaGetLastError());
    } else {
        ThinBwdParams p{};
        p.a = a->h; p.dt = a->d_heads; p.Wt = a->wh; p.dpre = a->dhpre; p.dWt = a->dwh; p.dbt = a->dbh; p.dcol = a->db2;
        p.store_scale = a->scales + 0;
        p.dt_scale = a->scales + 6;       // used by the mma.sync kernel (fp16 operand); the CUDA-core kernel reads d_heads in fp32
        p.act = tanh_act ? kActTanh : 0;
        p.M = R; p.W = g.O; p.T = NH; p.P = g.P;
        // row m = (b*G + r)*P + pos  ->  d_heads[((b*NH + j)*G + r)*P + pos]
        p.dt_outer = (long long)NH * G2 * g.P;    // stride of b
        p.dt_chan = (long long)G2 * g.P;          // stride of channel j
        if (NH <= 8) rc = launch_thin_bwd<8, 4, true, true>(p, G2, st);
        else if (NH <= 20) rc = launch_thin_bwd<20, 2, true, true>(p, G2, st);
        else rc = launch_thin_bwd<kMaxHeads + 1, 1, true, true>(p, G2, st);
        if (rc) return rc;
    }
    ++g_launch_count; transpose_half_kernel<<<blocks_for((long long)g.O * g.O, 256), 256, 0, st>>>(a->w2, static_cast<__half*>(a->w2t_h), g.O, g.O);
    if (pooled) {
        // ---- conv2 input is the pooled map xp (no activation in between): dW2 = dhpre^T xp, d(xp) = dhpre W2 (fp16 * s2),
        // then the adjoint of the pooling spreads d(xp) over the rotations, applies act'(x1) and yields dfc_r
        if ((rc = linear_tn(a->dhpre, g.O, a->xp, g.O, static_cast<int>(R), g.O, g.O, a->dw2, g.O, 0, st, a->scales + 1))) return rc;
        LinearNTArgs l{};
        l.A = a->dhpre; l.lda = g.O; l.B = a->w2t_h; l.ldb = g.O;
        l.M = static_cast<int>(R); l.N = g.O; l.K = g.O;
        l.C = nullptr; l.C16 = a->dxp16; l.ldc16 = g.O;
        l.acc_scale = a->scales + 1; l.store_scale = a->scales + 2;
        if ((rc = linear_nt(l, st))) return rc;
        TVAE_CHECK_CUDA(cudaMemsetAsync(a->dfc_w, 0, sizeof(float) * g.G, st));
        TVAE_CHECK_CUDA(cudaMemsetAsync(a->dfc_b, 0, sizeof(float), st));
        ++g_launch_count; rot_pool_scales_kernel<<<1, 1, 0, st>>>(a->fc_w, g.G, a->scales);
        RotPoolParams p{};
        p.x1 = static_cast<const __half*>(a->x1); p.fc_w = a->fc_w; p.dxp = static_cast<const __half*>(a->dxp16);
        p.dx1 = static_cast<__half*>(a->dx1_16); p.scales = a->scales; p.dfc_w = a->dfc_w; p.dfc_b = a->dfc_b;
        p.db1 = bias_grad_slot(g, a->dbank); p.db1_stride = g.kpad;
        p.B = g.B; p.G = g.G; p.P = g.P; p.O = g.O; p.act = tanh_act ? kActTanh : 0;
        TVAE_REQUIRE(g.O % 8 == 0, "rotation pooling: kernel count must be a multiple of 8");
        p.rows_per_cta = rot_pool_rows_per_cta(R, g.O);
        ++g_launch_count;
        rot_pool_bwd_kernel<<<cdiv(R, p.rows_per_cta), 256, sizeof(float) * (g.O + g.G + 1), st>>>(p);
        TVAE_CHECK_CUDA(cudaGetLastError());
        return conv1_wgrad(g, a->y, a->dx1_16, a->scales + 5, a->dbank, st);
    }
    if (g.O == 128 && !tanh_act) {
        // ---- one pass over dhpre and x1: dW2 = dhpre^T x1, dx1pre = (dhpre W2) * lrelu'(x1) (fp16 * s2) and its column
        // sums (the conv1 bias gradient)   (enc_bwd_fused.cuh)
        EncDx1Dw2Params q{};
        if ((rc = make_tmap_2d_h(&q.tmA, a->dhpre, R, 128, 128, kBM))) return rc;
        if ((rc = make_tmap_2d_h(&q.tmX, a->x1, R, 128, 128, kBM))) return rc;
        if ((rc = make_tmap_2d_h(&q.tmW, a->w2t_h, 128, 128, 128, 128))) return rc;
        if ((rc = make_tmap_2d_h(&q.tmC, a->dx1_16, R, 128, 128, kBM))) return rc;
        q.R = R; q.num_tiles = static_cast<int>(cdiv(R, kBM));
        q.acc_scale = a->scales + 1; q.store_scale = a->scales + 2;
        q.colsum = bias_grad_slot(g, a->dbank); q.colsum_stride = g.kpad;
        q.dw2 = a->dw2;
        TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&enc_dx1_dw2_kernel), kFuSmemBytes));
        const int grid = q.num_tiles < sm_count() ? q.num_tiles : sm_count();
        ++g_launch_count;
        const int tslot = g_timer.begin("enc_dx1_dw2", st);
        enc_dx1_dw2_kernel<<<grid, kFuThreads, kFuSmemBytes, st>>>(q);
        g_timer.end(tslot, st);
        TVAE_CHECK_CUDA(cudaGetLastError());
    } else {
    // ---- dW2 = dhpre^T x1   (fp16 x fp16, accumulators carry s1)
    if ((rc = linear_tn(a->dhpre, g.O, a->x1, g.O, static_cast<int>(R), g.O, g.O, a->dw2, g.O, 0, st, a->scales + 1))) return rc;
    // ---- dx1pre = (dhpre W2) * lrelu'(x1), stored fp16 * s2; its column sums are the conv1 bias gradient
    {
        LinearNTArgs l{};
        l.A = a->dhpre; l.lda = g.O; l.B = a->w2t_h; l.ldb = g.O;
        l.M = static_cast<int>(R); l.N = g.O; l.K = g.O;
        l.C = nullptr; l.C16 = a->dx1_16; l.ldc16 = g.O; l.aux16 = a->x1; l.ld_aux = g.O;
        l.aux_act = tanh_act ? kActTanh : 0;
        l.acc_scale = a->scales + 1; l.store_scale = a->scales + 2;
        l.colsum = bias_grad_slot(g, a->dbank); l.colsum_stride = g.kpad;
        if ((rc = linear_nt(l, st))) return rc;
    }
    }
    // ---- conv1 weight gradient (w.r.t. the rotated bank)
    return conv1_wgrad(g, a->y, a->dx1_16, a->scales + 3, a->dbank, st);
}

// ================================================================================ attention
static AttnParams to_attn_params(const tvae_attn_shape* s, const tvae_attn_fwd_args* a) {
    AttnParams p{};
    p.heads = a->heads; p.gumbel = a->gumbel; p.r_z = a->r_z; p.r_theta = a->r_theta; p.log_prior = a->log_prior;
    p.stats = a->stats; p.zb = a->zb; p.theta_b = a->theta_b; p.dx = a->dx; p.kl = a->kl;
    p.B = s->B; p.G = s->G; p.d = s->d; p.z = s->z; p.s = s->s; p.theta_prior_std = s->theta_prior_std;
    for (int i = 0; i < kMaxG; ++i) p.offsets[i] = s->offsets[i];
    return p;
}

int tvae_attn_log_prior(const tvae_attn_shape* s, const float* p_r_host16, float* log_prior, void* stream) {
    RotTable t{};
    for (int i = 0; i < s->G && i < kMaxG; ++i) t.cs[i] = p_r_host16[i];
    ++g_launch_count; log_prior_kernel<<<1, 1024, 0, S(stream)>>>(log_prior, s->G, s->d, s->s, t);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

#define TVAE_DISPATCH_Z(zval, CALL)                                                        \
    switch (zval) {                                                                        \
        case 1: { constexpr int ZZ = 1; CALL; } break;                                     \
        case 2: { constexpr int ZZ = 2; CALL; } break;                                     \
        case 3: { constexpr int ZZ = 3; CALL; } break;                                     \
        case 4: { constexpr int ZZ = 4; CALL; } break;                                     \
        case 5: { constexpr int ZZ = 5; CALL; } break;                                     \
        case 6: { constexpr int ZZ = 6; CALL; } break;                                     \
        case 8: { constexpr int ZZ = 8; CALL; } break;                                     \
        case 10: { constexpr int ZZ = 10; CALL; } break;                                   \
        case 16: { constexpr int ZZ = 16; CALL; } break;                                   \
        default: return fail(-1, "latent dim not instantiated (supported: 1-6, 8, 10, 16)"); \
    }

}  // extern "C"

namespace {
// cluster size of the attention forward: the smallest power of two (<= 8, the portable maximum) that gives every SM
// about six CTAs of 256 threads (two resident rounds), never so large that a CTA's slice falls below two sweeps of its threads
int attn_cluster_size(int B, int L) {
    int cl = 1;
    while (cl < 8 && B * cl < 6 * sm_count() && L / (2 * cl) >= kAttnThreads * 4) cl *= 2;
    return cl;
}
template <int Z, int VEC>
int launch_attn_fwd(const AttnParams& p, int B, int L, cudaStream_t st) {
    const int cl = attn_cluster_size(B, L);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * cl, 1, 1);
    cfg.blockDim = dim3(kAttnThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TVAE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_fwd_kernel<Z, VEC>, p));
    return 0;
}
}  // namespace

extern "C" {

int tvae_attn_fwd(const tvae_attn_shape* s, const tvae_attn_fwd_args* a, void* stream) {
    TVAE_REQUIRE(s->G >= 1 && s->G <= kMaxG && s->d >= 1 && s->B >= 1, "attention: bad shape");
    const AttnParams p = to_attn_params(s, a);
    const int L = s->G * s->d * s->d;
    // 16-byte accesses need whole float4 planes (every plane starts at a multiple of L floats)
    const bool vec = L % 4 == 0 && (reinterpret_cast<uintptr_t>(a->heads) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->gumbel) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(a->log_prior) & 15) == 0;
    Timed tm("attn_fwd", S(stream));
    ++g_launch_count;
    int rc = 0;
    if (vec) { TVAE_DISPATCH_Z(s->z, (rc = launch_attn_fwd<ZZ, 4>(p, s->B, L, S(stream)))); }
    else { TVAE_DISPATCH_Z(s->z, (rc = launch_attn_fwd<ZZ, 1>(p, s->B, L, S(stream)))); }
    if (rc) return rc;
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tvae_attn_bwd(const tvae_attn_shape* s, const tvae_attn_bwd_args* a, void* stream) {
    TVAE_REQUIRE(s->G >= 1 && s->G <= kMaxG && s->d >= 1 && s->B >= 1, "attention: bad shape");
    AttnBwdParams p{};
    p.heads = a->f.heads; p.gumbel = a->f.gumbel; p.r_z = a->f.r_z; p.r_theta = a->f.r_theta; p.log_prior = a->f.log_prior;
    p.stats = a->f.stats; p.zb = a->f.zb; p.theta_b = a->f.theta_b; p.dx = a->f.dx; p.kl = a->f.kl;
    p.g_zb = a->g_zb; p.g_theta = a->g_theta; p.g_dx = a->g_dx; p.g_kl = a->g_kl; p.d_heads = a->d_heads;
    p.B = s->B; p.G = s->G; p.d = s->d; p.z = s->z; p.s = s->s; p.theta_prior_std = s->theta_prior_std;
    for (int i = 0; i < kMaxG; ++i) p.offsets[i] = s->offsets[i];
    const int L = s->G * s->d * s->d;
    const bool vec = L % 4 == 0 && (reinterpret_cast<uintptr_t>(a->f.heads) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->f.gumbel) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(a->f.log_prior) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->d_heads) & 15) == 0;
    // enough CTAs of 256 threads to give every SM ~8 of them, one vector per thread and sweep
    const int per_image = cdiv(L, 256 * (vec ? 4 : 1));
    int gx = cdiv(8LL * sm_count(), s->B);
    if (gx > per_image) gx = per_image;
    if (gx < 1) gx = 1;
    dim3 grid(gx, s->B);
    Timed tm("attn_bwd", S(stream));
    ++g_launch_count;
    if (vec) { TVAE_DISPATCH_Z(s->z, (attn_bwd_kernel<ZZ, 4><<<grid, 256, 0, S(stream)>>>(p))); }
    else { TVAE_DISPATCH_Z(s->z, (attn_bwd_kernel<ZZ, 1><<<grid, 256, 0, S(stream)>>>(p))); }
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tvae_attn_softmax_pair(const float* heads, const float* gumbel, float* q_t_r, float* a_sampled, int B, int NH, int L, void* stream) {
    ++g_launch_count; softmax_pair_kernel<<<B, 1024, 0, S(stream)>>>(heads, gumbel, q_t_r, a_sampled, NH, L);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tvae_get_latent(const tvae_attn_shape* s, const float* heads, float* z_content, float* theta_mu, float* dx, int* argmax, void* stream) {
    ++g_launch_count;
    TVAE_DISPATCH_Z(s->z, (get_latent_kernel<ZZ><<<s->B, 1024, 0, S(stream)>>>(heads, s->G, s->d, s->s, z_content, theta_mu, dx, argmax)));
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// exact re-evaluation of the attention logits at the near-maximal cells of the fast maps (refine.cuh)
int tvae_refine_argmax(const tvae_enc_shape* s, const tvae_refine_args* a, void* stream) {
    int rc = check_enc_shape(s);
    if (rc) return rc;
    TVAE_REQUIRE(a->y && a->weight && a->conv1_bias && a->w2 && a->b2 && a->wh && a->bh && a->head_add && a->heads,
                 "refine_argmax: null input");
    TVAE_REQUIRE(a->bank32 && a->cand && a->n_cand && a->cand_heads && a->z_content && a->theta_mu && a->argmax,
                 "refine_argmax: null scratch / output");
    TVAE_REQUIRE(a->rel_tol >= 0.f && a->rel_tol < 1.f, "refine_argmax: rel_tol must be in [0, 1)");
    const ConvGeom g = make_geom(s);
    const int NH = 3 + 2 * s->z;
    const bool pooled = a->fc_w != nullptr;
    TVAE_REQUIRE(!pooled || a->fc_b, "refine_argmax: rotation pooling needs fc_b");
    const int G2 = pooled ? 1 : g.G;
    const int L = G2 * g.P;
    cudaStream_t st = S(stream);
    Timed tm("refine_argmax", st);
    ++g_launch_count;
    filter_bank_f32_kernel<<<blocks_for((long long)g.G * g.O * g.K, 256), 256, 0, st>>>(a->weight, a->bank32, g.O, g.C, g.k, g.G, make_rot_table(g.G));
    ++g_launch_count;
    refine_select_kernel<<<g.B, 1024, 0, st>>>(a->heads, NH, L, a->rel_tol, a->cand, a->n_cand);
    RefineEvalParams p{};
    p.y = a->y; p.bank32 = a->bank32; p.conv1_bias = a->conv1_bias; p.w2 = a->w2; p.b2 = a->b2; p.wh = a->wh; p.bh = a->bh;
    p.head_add = a->head_add; p.fc_w = a->fc_w; p.fc_b = a->fc_b; p.cand = a->cand; p.n_cand = a->n_cand; p.cand_heads = a->cand_heads;
    p.C = g.C; p.n = g.n; p.k = g.k; p.p = g.p; p.G = g.G; p.O = g.O; p.d = g.d; p.NH = NH;
    p.act = s->act == TVAE_ACT_TANH ? kActTanh : 0;
    const size_t sm = sizeof(float) * (((g.K + 3) & ~3) + 3 * g.O);
    TVAE_REQUIRE(sm <= 227 * 1024, "refine_argmax: filter window does not fit shared memory");
    TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&refine_eval_kernel), 227 * 1024));
    ++g_launch_count;
    refine_eval_kernel<<<dim3(kRefineMaxCand, g.B), 256, sm, st>>>(p);
    ++g_launch_count;
    refine_pick_kernel<<<g.B, 32, 0, st>>>(a->cand, a->n_cand, a->cand_heads, NH, s->z, a->z_content, a->theta_mu, a->argmax, a->refined_logit);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ================================================================================ generator
static int check_gen_shape(const tvae_gen_shape* s) {
    TVAE_REQUIRE(s->B > 0 && s->N > 0, "generator: empty batch");
    TVAE_REQUIRE(s->H % 32 == 0 && s->H >= 32 && s->H <= 1024, "generator: hidden dim must be a multiple of 32 in [32,1024]");
    TVAE_REQUIRE(s->E == 0 || (s->E % 32 == 0 && s->E <= 4096), "generator: Fourier dim must be a multiple of 32");
    TVAE_REQUIRE(s->L >= 0 && s->L <= 8, "generator: 0..8 hidden layers");
    TVAE_REQUIRE(s->n_out >= 1 && s->n_out <= 4 && s->zdim >= 1, "generator: n_out in 1..4");
    TVAE_REQUIRE(s->act == TVAE_ACT_LEAKYRELU || s->act == TVAE_ACT_TANH, "generator: act must be TVAE_ACT_LEAKYRELU or TVAE_ACT_TANH");
    return 0;
}

// Generators without Fourier features whose coordinate layer is fused into the first hidden layer's GEMM (forward) and
// regenerated in its weight gradient (backward): LeakyReLU, 128 | H <= 512, whole 128-row tiles touch at most two images.
static bool gen_coord_fused(const tvae_gen_shape* s) {
    return g_dev_knob[1] == 0 && s->E == 0 && s->L >= 1 && s->act == TVAE_ACT_LEAKYRELU && s->H % 128 == 0 && s->H <= 2 * kAccN &&
           s->N >= kBM && (long long)s->B * s->N < (1LL << 31);
}

// Fourier generators on the pair kernel (GenL1FwdPairT<., 0>)
static bool gen_l1_pair(const tvae_gen_shape* s) {
    return s->E > 0 && s->H % 64 == 0 && s->H <= 2 * kAccN && s->H > 128 && s->N >= kBM;
}
// One-bit LeakyReLU masks (tvae_gen_fwd_args.mask_bits): is the mask of acts[layer] written by the forward pass?  Layer 0 by the
// pair kernel of the Fourier layer, layers 1 .. L-1 by the hidden layers' LinearNT epilogues (acts[L] feeds the output layer's
// backward kernel, which reads the activation itself).  The coordinate-fused path keeps its own mask in the acts[0] buffer.
static unsigned long long* gen_mask_words(const tvae_gen_shape* s, const tvae_gen_fwd_args* a, int layer) {
    if (!a->mask_bits || s->act != TVAE_ACT_LEAKYRELU || s->H % 64 != 0 || layer >= s->L) return nullptr;
    if (layer == 0 && !gen_l1_pair(s)) return nullptr;
    if (layer == 1 && gen_coord_fused(s)) return nullptr;            // acts[1] is written by the fused coordinate-layer kernel
    return static_cast<unsigned long long*>(a->mask_bits) + (long long)layer * (s->H / 64) * s->B * s->N;
}

static CoordXform make_xform(const tvae_gen_shape* s, const tvae_gen_fwd_args* a) {
    CoordXform c{};
    c.x = a->x; c.theta = a->theta; c.dx = a->dx; c.N = s->N; c.M = (long long)s->B * s->N;
    return c;
}

int tvae_generator_fwd(const tvae_gen_shape* s, const tvae_gen_fwd_args* a, void* stream) {
    int rc = check_gen_shape(s);
    if (rc) return rc;
    cudaStream_t st = S(stream);
    const long long M = (long long)s->B * s->N;
    const int H = s->H, E = s->E;
    const CoordXform cx = make_xform(s, a);
    const int gact = s->act == TVAE_ACT_TANH ? kActTanh : 1;
    ++g_launch_count; latent_bias_kernel<<<cdiv(s->B * H, 256), 256, 0, st>>>(a->z, a->wz, a->zb, s->B, H, s->zdim);
    __half* w1h = static_cast<__half*>(a->w_h);                                    // [H][E]
    __half* whh = static_cast<__half*>(a->w_h) + (long long)H * (E > 0 ? E : 2);   // [L][H][H]
    __half* acts = static_cast<__half*>(a->acts);
    if (s->L > 0) { ++g_launch_count; to_half_kernel<<<blocks_for((long long)s->L * H * H, 256), 256, 0, st>>>(a->wh, whh, (long long)s->L * H * H); }
    __half* a0 = acts;
    int first_hidden = 1;
    if (E > 0) {
        ++g_launch_count; to_half_kernel<<<blocks_for((long long)H * E, 256), 256, 0, st>>>(a->w1, w1h, (long long)H * E);
        if (gen_l1_pair(s)) {
            // CTA-pair kernel: every generated feature chunk feeds all H hidden columns (gen_pair_policies.cuh)
            GenL1FwdPairParams q{};
            q.mask_bits = gen_mask_words(s, a, 0);
            if ((rc = make_tmap_2d_h(&q.tmB, w1h, H, E, E, 128))) return rc;
            if ((rc = make_tmap_2d_h(&q.tmC, a0, M, H, H, kBM))) return rc;
            q.cx = cx; q.wf_scaled = a->wf_scaled; q.bf = a->bf; q.E = E; q.H = H;
            q.bias = a->b1; q.zb = a->zb; q.act = gact;
            q.m_tiles = static_cast<int>(cdiv(M, kBM));
            q.k_chunks = cdiv(E, kBKh);
            q.num_tiles = cdiv(q.m_tiles, 2);
            int extra = E * 16;
            q.bias_off = extra;
            extra += H * 4;
            q.tab_off = extra;
            extra += 2 * H * 4;
            extra = (extra + 1023) / 1024 * 1024;
            q.stage_off = extra;
            extra += 3 * kStoreBlockBytes;
            rc = gact == kActTanh ? launch_gemm2<GenL1FwdPairT<true>>(q, extra, st) : launch_gemm2<GenL1FwdPair>(q, extra, st);
            if (rc) return rc;
        } else {
        GenL1FwdParams p{};
        const bool wide = H > 128;
        const int BN = wide ? 256 : 128;
        if ((rc = make_tmap_2d_h(&p.tmB, w1h, H, E, E, BN))) return rc;
        p.cx = cx; p.wf_scaled = a->wf_scaled; p.bf = a->bf; p.E = E; p.H = H;
        p.bias = a->b1; p.zb = a->zb; p.h1 = a0; p.act = gact;
        p.tiles_n = cdiv(H, BN);
        p.k_chunks = cdiv(E, kBKh);
        p.num_tiles = cdiv(M, kBM) * p.tiles_n;
        const int extra = E * 16;
        if (gact == kActTanh) rc = wide ? launch_gemm<GenL1Fwd<256, true>>(p, extra, st) : launch_gemm<GenL1Fwd<128, true>>(p, extra, st);
        else rc = wide ? launch_gemm<GenL1Fwd<256>>(p, extra, st) : launch_gemm<GenL1Fwd<128>>(p, extra, st);
        if (rc) return rc;
        }
    } else if (gen_coord_fused(s)) {
        // no Fourier features: coordinate layer generated as the A operand of the first hidden layer's GEMM, output
        // projection fused when that is the last hidden layer (gen_pair_policies.cuh, GenL1FwdPairT<., 1>); acts[0] is never
        // written - its buffer holds the one-bit LeakyReLU mask of the coordinate layer for the backward pass
        GenL1FwdPairParams q{};
        __half* a1 = acts + M * H;
        if ((rc = make_tmap_2d_h(&q.tmB, whh, H, H, H, 128))) return rc;
        if ((rc = make_tmap_2d_h(&q.tmC, a1, M, H, H, kBM))) return rc;
        q.cx = cx; q.wf_scaled = a->w1; q.bf = a->b1; q.zbc = a->zb; q.E = H; q.H = H;
        q.bias = a->bh; q.zb = nullptr; q.act = gact;
        q.mask_bits = reinterpret_cast<unsigned long long*>(acts);
        if (s->L == 1) { q.proj_w = a->wout; q.proj_bias = a->bout; q.proj_out = a->y_hat; q.n_proj = s->n_out; }
        q.m_tiles = static_cast<int>(cdiv(M, kBM));
        q.k_chunks = cdiv(H, kBKh);
        q.num_tiles = cdiv(q.m_tiles, 2);
        int extra = 3 * H * 16;                    // feature tables {w0, w1, b1 + zb} of two images + the static {w0, w1, b1}
        q.bias_off = extra;
        extra += H * 4;
        q.tab_off = extra;
        extra += 2 * H * 4;
        q.proj_off = extra;
        extra += q.n_proj * H * 4;
        extra = (extra + 1023) / 1024 * 1024;
        q.stage_off = extra;
        extra += 3 * kStoreBlockBytes;
        if ((rc = launch_gemm2<GenL1FwdPairT<false, 1>>(q, extra, st))) return rc;
        first_hidden = 2;
    } else {
        ++g_launch_count;
        if (H % 8 == 0 && H / 8 <= 256) {          // 16-byte stores
            const int cgs = H / 8, rpp = 256 / cgs > 0 ? 256 / cgs : 1;
            if (gact == kActTanh) coord_layer_fwd_kernel<8, true><<<cdiv(M, kCoordRB), cgs * rpp, 0, st>>>(cx, a->w1, a->b1, a->zb, a0, H);
            else coord_layer_fwd_kernel<8, false><<<cdiv(M, kCoordRB), cgs * rpp, 0, st>>>(cx, a->w1, a->b1, a->zb, a0, H);
        } else {
            const int cgs = H / 4, rpp = 256 / cgs > 0 ? 256 / cgs : 1;
            if (gact == kActTanh) coord_layer_fwd_kernel<4, true><<<cdiv(M, kCoordRB), cgs * rpp, 0, st>>>(cx, a->w1, a->b1, a->zb, a0, H);
            else coord_layer_fwd_kernel<4, false><<<cdiv(M, kCoordRB), cgs * rpp, 0, st>>>(cx, a->w1, a->b1, a->zb, a0, H);
        }
        TVAE_CHECK_CUDA(cudaGetLastError());
    }
    if (s->L == 0) {
        ++g_launch_count; thin_fwd_kernel<<<cdiv(M, 8), 256, 0, st>>>(a0, a->wout, a->bout, a->y_hat, M, H, s->n_out);
        TVAE_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (first_hidden > s->L) return 0;             // the fused kernel was the last hidden layer and wrote y_hat itself
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->y_hat, 0, sizeof(float) * M * s->n_out, st));
    for (int i = first_hidden; i <= s->L; ++i) {
        LinearNTArgs l{};
        l.A = acts + (long long)(i - 1) * M * H; l.lda = H;
        l.B = whh + (long long)(i - 1) * H * H; l.ldb = H;
        l.M = static_cast<int>(M); l.N = H; l.K = H;
        l.C16 = acts + (long long)i * M * H; l.ldc16 = H;
        l.bias = a->bh + (long long)(i - 1) * H;
        l.act = gact;
        l.bits_out = gen_mask_words(s, a, i);
        if (i == s->L) { l.proj_w = a->wout; l.proj_bias = a->bout; l.proj_out = a->y_hat; l.n_proj = s->n_out; }
        if ((rc = linear_nt(l, st))) return rc;
    }
    return 0;
}

int tvae_generator_bwd(const tvae_gen_shape* s, const tvae_gen_bwd_args* a, void* stream) {
    int rc = check_gen_shape(s);
    if (rc) return rc;
    cudaStream_t st = S(stream);
    const long long M = (long long)s->B * s->N;
    const int H = s->H, E = s->E, L = s->L;
    const CoordXform cx = make_xform(s, &a->f);
    const int gact = s->act == TVAE_ACT_TANH ? kActTanh : 0;
    const __half* acts = static_cast<const __half*>(a->f.acts);
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->dwout, 0, sizeof(float) * s->n_out * H, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->dbout, 0, sizeof(float) * s->n_out, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->db1, 0, sizeof(float) * H, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->dzb, 0, sizeof(float) * s->B * H, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->dw1, 0, sizeof(float) * H * (E > 0 ? E : 2), st));
    if (L > 0) {
        TVAE_CHECK_CUDA(cudaMemsetAsync(a->dwh, 0, sizeof(float) * L * H * H, st));
        TVAE_CHECK_CUDA(cudaMemsetAsync(a->dbh, 0, sizeof(float) * L * H, st));
    }
    // ---- power-of-two scales s_i of the fp16 gradient tensors dpre_i (i = L .. 0): scales[2i] = s_i, scales[2i+1] = 1/s_i
    TVAE_CHECK_CUDA(cudaMemsetAsync(a->scales, 0, sizeof(float) * 32, st));
    ++g_launch_count; absmax_kernel<<<blocks_for(M * s->n_out, 256), 256, 0, st>>>(a->d_yhat, M * s->n_out, a->scales + 31);
    {   // scales[20 .. 20 + L]: per-layer max column abs-sum of the weights (L <= 8)
        ++g_launch_count; gen_colmax_kernel<<<dim3(cdiv(H, 32), L + 1), 1024, 0, st>>>(a->f.wout, s->n_out, a->f.wh, L, H, a->scales + 20);
        ++g_launch_count; gen_bwd_scales_kernel<<<1, 1, 0, st>>>(a->scales + 31, a->scales + 20, L, a->scales);
    }
    __half* dcur = static_cast<__half*>(a->dpre0);
    __half* dnext = static_cast<__half*>(a->dpre1);
    __half* wt_h = static_cast<__half*>(a->wt_h);
    // ---- output layer backward -> dpre of the last hidden activation
    {
        ThinBwdParams p{};
        p.a = acts + (long long)L * M * H; p.dt = a->d_yhat; p.Wt = a->f.wout; p.dpre = dcur;
        p.store_scale = a->scales + 2 * L;
        p.dWt = a->dwout; p.dbt = a->dbout; p.dcol = (L > 0) ? a->dbh + (long long)(L - 1) * H : nullptr;
        p.M = M; p.W = H; p.T = s->n_out; p.P = 1; p.dt_outer = s->n_out; p.dt_chan = 1;
        p.act = gact;
        if ((rc = launch_thin_bwd<4, 4, false, true>(p, 1, st))) return rc;
    }
    // ---- hidden layers, last to first
    const bool coord_fused = gen_coord_fused(s);
    for (int i = L; i >= 1; --i) {
        const __half* a_prev = acts + (long long)(i - 1) * M * H;
        const float* w = a->f.wh + (long long)(i - 1) * H * H;
        if (i == 1 && coord_fused) {
            // dWh[0][j][f] = sum_m dpre1[m][j] a0[m][f] with a0 (the coordinate layer's activation) regenerated on chip
            GenL1WgradPairParams p{};
            if ((rc = make_tmap_3d_mn_h(&p.tmQ, dcur, M, H, H, kBKh, 4))) return rc;
            p.cx = cx; p.wf_scaled = a->f.w1; p.bf = a->f.b1; p.zbc = a->f.zb; p.E = H; p.H = H; p.dW1 = a->dwh;
            p.acc_scale = a->scales + 2 * i + 1;
            p.m_tiles = cdiv(H, kBM);
            p.m_pairs = cdiv(p.m_tiles, 2);
            p.chunks_total = static_cast<int>(cdiv(M, kBKh));
            const int pairs_dev = sm_count() / 2;
            int splits = pairs_dev / p.m_pairs;
            if (splits > p.chunks_total / 8) splits = p.chunks_total / 8;
            if (splits < 1) splits = 1;
            p.chunks_per_split = cdiv(p.chunks_total, splits);
            p.splits = cdiv(p.chunks_total, p.chunks_per_split);
            p.num_tiles = p.m_pairs * p.splits;
            int extra = H * 16;
            rc = make_tmap_2d_f32_reduce(&p.tmD, a->dwh, H, H, H);
            if (rc < 0) return rc;
            p.tma_reduce = rc == 0;
            if (p.tma_reduce) { p.stage_off = extra; extra += 2 * kStoreBlockBytes; }
            if ((rc = launch_gemm2<GenL1WgradPairT<1>>(p, extra, st))) return rc;
        } else
        if ((rc = linear_tn(dcur, H, a_prev, H, static_cast<int>(M), H, H, a->dwh + (long long)(i - 1) * H * H, H, 0, st,
                            a->scales + 2 * i + 1))) return rc;
        ++g_launch_count; transpose_half_kernel<<<blocks_for((long long)H * H, 256), 256, 0, st>>>(w, wt_h, H, H);
        LinearNTArgs l{};
        l.A = dcur; l.lda = H; l.B = wt_h; l.ldb = H;
        l.M = static_cast<int>(M); l.N = H; l.K = H;
        l.C16 = dnext; l.ldc16 = H; l.aux16 = a_prev; l.ld_aux = H; l.aux_act = gact;
        if (i == 1 && coord_fused) { l.aux16 = nullptr; l.aux_bits = reinterpret_cast<const unsigned long long*>(acts); }   // one-bit mask written by the forward kernel
        else if (const unsigned long long* mb = gen_mask_words(s, &a->f, i - 1)) { l.aux16 = nullptr; l.aux_bits = mb; }
        l.acc_scale = a->scales + 2 * i + 1; l.store_scale = a->scales + 2 * (i - 1);
        if (i - 1 >= 1) { l.colsum = a->dbh + (long long)(i - 2) * H; l.colsum_stride = 1; }   // bias gradient of hidden layer i-1
        if ((rc = linear_nt(l, st))) return rc;
        __half* t = dcur; dcur = dnext; dnext = t;
    }
    // ---- dcur == dpre of layer 1 (scaled by s_0): bias / latent-bias gradients.  Without Fourier features and with
    // whole 64-row blocks per image they come out of the coordinate-layer backward pass below instead.
    int coord_rows = 0;
    if (E == 0 && s->N % kCoordRB == 0) {
        const int cap = g_dev_knob[6] > 0 ? g_dev_knob[6] : 1024;   // measured at cfg2: 1024 rows per CTA 166 us, 512 176 us, 256 194 us, 2048 278 us (8 rows in flight per lane: 210 us - the kernel is issue-bound, not latency-bound)
        for (int r = kCoordRB; r <= s->N && r <= cap && r <= kCoordMaxRows; r += kCoordRB)
            if (s->N % r == 0) coord_rows = r;
    }
    if (!coord_rows) {
        ++g_launch_count;
        if (H % 8 == 0 && H / 8 <= 256) {
            const int cgs = H / 8, slots = 256 / cgs;
            group_colsum8_kernel<<<dim3(cdiv(s->N, 512), s->B), cgs * slots, sizeof(float) * slots * H, st>>>(dcur, a->scales + 1, a->dzb, a->db1, s->N, H, 512);
        } else {
            group_colsum_kernel<<<dim3(cdiv(s->N, 512), s->B), H, 0, st>>>(dcur, a->scales + 1, a->dzb, a->db1, s->N, H, 512);
        }
    }
    // ---- layer 1 weight and coordinate gradients
    if (E > 0) {
        if (H % 128 == 0 && H <= 2 * kAccN) {
            // CTA-pair kernel: all H hidden columns per generated feature chunk (gen_pair_policies.cuh)
            GenL1WgradPairParams p{};
            if ((rc = make_tmap_3d_mn_h(&p.tmQ, dcur, M, H, H, kBKh, 4))) return rc;
            p.cx = cx; p.wf_scaled = a->f.wf_scaled; p.bf = a->f.bf; p.E = E; p.H = H; p.dW1 = a->dw1;
            p.acc_scale = a->scales + 1;
            p.m_tiles = cdiv(E, kBM);
            p.m_pairs = cdiv(p.m_tiles, 2);
            p.chunks_total = static_cast<int>(cdiv(M, kBKh));
            const int pairs_dev = sm_count() / 2;
            int splits = pairs_dev / p.m_pairs;
            if (splits > p.chunks_total / 8) splits = p.chunks_total / 8;
            if (splits < 1) splits = 1;
            p.chunks_per_split = cdiv(p.chunks_total, splits);
            p.splits = cdiv(p.chunks_total, p.chunks_per_split);
            p.num_tiles = p.m_pairs * p.splits;
            int extra = E * 16;
            rc = make_tmap_2d_f32_reduce(&p.tmD, a->dw1, H, E, E);
            if (rc < 0) return rc;
            p.tma_reduce = rc == 0;
            if (p.tma_reduce) { p.stage_off = extra; extra += 2 * kStoreBlockBytes; }
            if ((rc = launch_gemm2<GenL1WgradPair>(p, extra, st))) return rc;
        } else {
            GenL1WgradParams p{};
            const bool wide = H > 128;
            const int BN = wide ? 256 : 128;
            if ((rc = make_tmap_2d_mn_h(&p.tmQ, dcur, M, H, H, kBKh))) return rc;
            p.cx = cx; p.wf_scaled = a->f.wf_scaled; p.bf = a->f.bf; p.E = E; p.H = H; p.dW1 = a->dw1;
            p.acc_scale = a->scales + 1;
            p.tiles_m = cdiv(E, kBM);
            p.tiles_n = cdiv(H, BN);
            const int extra = E * 16;
            const int out_tiles = p.tiles_m * p.tiles_n;
            rc = wide ? launch_split_tn<GenL1Wgrad<256>>(p, out_tiles, cdiv(M, kBKh), 1, extra, st)
                      : launch_split_tn<GenL1Wgrad<128>>(p, out_tiles, cdiv(M, kBKh), 1, extra, st);
            if (rc) return rc;
        }
        {
            TVAE_CHECK_CUDA(cudaMemsetAsync(a->dxp, 0, sizeof(float) * M * 2, st));
            ++g_launch_count; transpose_half_kernel<<<blocks_for((long long)H * E, 256), 256, 0, st>>>(a->f.w1, wt_h, H, E);   // -> [E][H]
            GenL1DgradParams p{};
            const bool wide = E > 128;
            const int BN = wide ? 256 : 128;
            if ((rc = make_tmap_2d_h(&p.tmA, dcur, M, H, H, kBM))) return rc;
            if ((rc = make_tmap_2d_h(&p.tmB, wt_h, E, H, H, BN))) return rc;
            p.cx = cx; p.wf_scaled = a->f.wf_scaled; p.bf = a->f.bf; p.E = E; p.H = H; p.dxp = a->dxp;
            p.acc_scale = a->scales + 1;
            p.tiles_n = cdiv(E, BN);
            p.k_chunks = cdiv(H, kBKh);
            p.num_tiles = cdiv(M, kBM) * p.tiles_n;
            const int extra = E * 16;
            rc = wide ? launch_gemm<GenL1Dgrad<256>>(p, extra, st) : launch_gemm<GenL1Dgrad<128>>(p, extra, st);
            if (rc) return rc;
        }
    } else {
        const int cgs = H / 4, rpp = 256 / cgs > 0 ? 256 / cgs : 1;
        long long rows = (M + 148LL * 8 - 1) / (148LL * 8);
        rows = (rows + kCoordRB - 1) / kCoordRB * kCoordRB;
        if (coord_rows) rows = coord_rows;
        const size_t sm = sizeof(float) * (4 * kCoordRB + 3 * H);
        ++g_launch_count;
        if (H % 8 == 0 && H <= 512 && rows <= kCoordMaxRows) {
            // row-streaming kernel: warps own whole rows, 16-byte loads, register accumulators
            const size_t smr = sizeof(float) * (4 * rows + 3 * H);
            coord_layer_bwd_rows_kernel<<<cdiv(M, rows), 256, smr, st>>>(cx, a->f.w1, dcur, a->scales + 1, a->dw1, a->dxp,
                                                                          coord_rows ? a->dzb : nullptr, a->db1, H, static_cast<int>(rows));
        } else
        coord_layer_bwd_kernel<<<cdiv(M, rows), cgs * rpp, sm, st>>>(cx, a->f.w1, dcur, a->scales + 1, a->dw1, a->dxp,
                                                                      coord_rows ? a->dzb : nullptr, a->db1, H, static_cast<int>(rows));
        TVAE_CHECK_CUDA(cudaGetLastError());
    }
    ++g_launch_count; latent_bias_bwd_kernel<<<cdiv((s->B + H) * s->zdim * 32, 128), 128, 0, st>>>(a->dzb, a->f.z, a->f.wz, a->dwz, a->d_z, s->B, H, s->zdim);
    TVAE_CHECK_CUDA(cudaGetLastError());
    if (a->f.theta && a->d_theta) {
        ++g_launch_count; coord_xform_bwd_kernel<<<s->B, 256, 0, st>>>(cx, a->dxp, a->d_theta, a->d_dx);
        TVAE_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

// ================================================================================ likelihoods
int tvae_bernoulli(const float* y_hat, const float* y, float* ll, float* d_yhat, const float* g, int B, int E, void* stream) {
    cudaStream_t st = S(stream);
    TVAE_CHECK_CUDA(cudaMemsetAsync(ll, 0, sizeof(float) * B, st));
    dim3 grid(1, B);   // one CTA per image: ll[b] is a fixed-order sum (bit-deterministic forward)
    ++g_launch_count; bernoulli_kernel<<<grid, 256, 0, st>>>(y_hat, y, ll, d_yhat, E, g);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

namespace {
// tensor-core CTF path: even n <= 128 (one 128 x 128 accumulator per image, the padded filter fits shared memory)
bool ctf_gemm_ok(int n) { return n >= 4 && n <= 128 && n % 2 == 0; }
CtfGeom make_ctf_geom(int B, int n) {
    CtfGeom g{};
    g.B = B; g.n = n; g.m = n - 1;
    g.Hp = n + g.m - 1;
    g.Wp = (g.Hp + 63) / 64 * 64;
    g.CP = (g.m + 1 + 7) / 8 * 8;
    const int fp = g.Wp > g.Hp ? g.Wp : g.Hp;
    g.FP = (fp + 7) / 8 * 8;
    g.WQ = g.Wp / 64;
    return g;
}
struct CtfWorkspace { __half* ypad; __half* ctf16; __half* flip16; float* scales; long long bytes; };
CtfWorkspace ctf_workspace(const CtfGeom& g, void* base) {
    auto up = [](long long x) { return (x + 255) / 256 * 256; };
    CtfWorkspace w{};
    char* p = static_cast<char*>(base);
    long long off = 0;
    w.ypad = reinterpret_cast<__half*>(p + off);   off += up(2LL * g.B * g.Hp * g.Wp);
    w.ctf16 = reinterpret_cast<__half*>(p + off);  off += up(2LL * g.B * g.m * g.CP);
    w.flip16 = reinterpret_cast<__half*>(p + off); off += up(2LL * g.B * g.m * g.CP);
    w.scales = reinterpret_cast<float*>(p + off);  off += 256;
    w.bytes = off;
    return w;
}
// out (B,n,n) = correlation of in (B,n,n) with the fp16 filters `flt` (already flipped for the adjoint)
int ctf_apply_gemm(const CtfGeom& g, const CtfWorkspace& w, const float* in, const __half* flt, const float* in_scale,
                   const float* out_scale, float* out, cudaStream_t st) {
    ++g_launch_count;
    ctf_pad_input_kernel<<<blocks_for((long long)g.B * g.Hp * g.Wp, 256), 256, 0, st>>>(in, w.ypad, g, in_scale);
    TVAE_CHECK_CUDA(cudaGetLastError());
    CtfApplyParams p{};
    int rc;
    if ((rc = make_tmap_2d_h(&p.tmB, w.ypad, (uint64_t)g.B * g.Hp, g.Wp, g.Wp, 128))) return rc;
    p.g = g; p.ctf16 = flt; p.out = out; p.acc_scale = out_scale;
    p.num_tiles = g.B;
    const int extra = (256 + g.m * g.FP + 256) * 2;
    return launch_gemm<CtfApply>(p, extra, st);
}
}  // namespace

extern "C" {

long long tvae_gaussian_workspace_bytes(int B, int n) {
    if (!ctf_gemm_ok(n)) return 0;
    return ctf_workspace(make_ctf_geom(B, n), nullptr).bytes;
}

int tvae_gaussian(const float* y_hat, const float* y, const float* ctf, int ctf_size, const float* dx, float s, int radius,
                  float* mu, float* dmu, float* ll, float* d_yhat, const float* g, int B, int n, void* ws, void* stream) {
    cudaStream_t st = S(stream);
    TVAE_REQUIRE(B >= 1 && n >= 1, "gaussian: empty batch");
    TVAE_CHECK_CUDA(cudaMemsetAsync(ll, 0, sizeof(float) * B, st));
    // filter size: the reference applies any (B,1,m,m) filter with padding m // 2 (train_particles.py:298-302); the
    // output keeps the image size only for odd m (and for m = n - 1, the trainer's default, with even n)
    const int mf = (ctf && ctf_size > 0) ? ctf_size : n - 1;
    TVAE_REQUIRE(!ctf || mf == n - 1 || (mf % 2 == 1), "gaussian: CTF filter size must be odd (or n - 1)");
    const int m = (mf != n - 1 && mf > 2 * n - 1) ? 2 * n - 1 : mf;     // taps beyond n - 1 from the centre never meet the image
    const int W = 16 + m - 1;
    const size_t sm = sizeof(float) * (W * W + m * m);
    dim3 cgrid(cdiv(n, 16), cdiv(n, 16), B);
    const bool gemm = ctf && ws && mf == n - 1 && ctf_gemm_ok(n);
    const CtfGeom cg = make_ctf_geom(B, n);
    const CtfWorkspace cw = ctf_workspace(cg, ws);
    int rc;
    const float* mu_in = y_hat;
    if (gemm) {
        ++g_launch_count;
        ctf_to_half_kernel<<<blocks_for((long long)B * cg.m * cg.CP, 256), 256, 0, st>>>(ctf, cw.ctf16, cw.flip16, cg);
        TVAE_CHECK_CUDA(cudaGetLastError());
        // y_hat == NULL: mu already holds ctf (*) y_hat from the forward call (the backward pass does not recompute it)
        if (y_hat && (rc = ctf_apply_gemm(cg, cw, y_hat, cw.ctf16, nullptr, nullptr, mu, st))) return rc;
        mu_in = mu;
    } else if (ctf) {
        TVAE_REQUIRE(sm <= 227 * 1024, "gaussian: CTF window does not fit shared memory (filter too large for this image size)");
        TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&ctf_apply_kernel<false>), 227 * 1024));
        TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&ctf_apply_kernel<true>), 227 * 1024));
        if (y_hat) { ++g_launch_count; ctf_apply_kernel<false><<<cgrid, 256, sm, st>>>(y_hat, ctf, mu, n, mf, m); }
        mu_in = mu;
    }
    dim3 grid(1, B);   // one CTA per image: deterministic ll[b]
    float* dmu_out = d_yhat ? (ctf ? dmu : d_yhat) : nullptr;
    ++g_launch_count; gaussian_kernel<<<grid, 256, 0, st>>>(mu_in, y, dx, s, n, radius, ll, dmu_out, g);
    TVAE_CHECK_CUDA(cudaGetLastError());
    if (ctf && d_yhat) {
        if (gemm) {
            // adjoint = correlation with the flipped filters; dmu is scaled into fp16's range by a power of two
            TVAE_CHECK_CUDA(cudaMemsetAsync(cw.scales, 0, sizeof(float) * 8, st));
            ++g_launch_count; absmax_kernel<<<blocks_for((long long)B * n * n, 256), 256, 0, st>>>(dmu, (long long)B * n * n, cw.scales + 7);
            ++g_launch_count; single_scale_kernel<<<1, 1, 0, st>>>(cw.scales + 7, cw.scales);
            if ((rc = ctf_apply_gemm(cg, cw, dmu, cw.flip16, cw.scales + 2, cw.scales + 3, d_yhat, st))) return rc;
        } else {
            ++g_launch_count; ctf_apply_kernel<true><<<cgrid, 256, sm, st>>>(dmu, ctf, d_yhat, n, mf, m);
        }
    }
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tvae_gaussian_fit_noise(const float* y_hat2, const float* y, float* ll, float* d_yhat2, const float* g, int B, int N,
                            void* stream) {
    TVAE_REQUIRE(B >= 1 && N >= 1, "gaussian (fit-noise): empty batch");
    TVAE_REQUIRE(y_hat2 && y && ll, "gaussian (fit-noise): null pointer");
    TVAE_REQUIRE(!d_yhat2 || g, "gaussian (fit-noise): gradient requested without its scale");
    ++g_launch_count; gaussian_fit_noise_kernel<<<B, 256, 0, S(stream)>>>(y_hat2, y, N, ll, d_yhat2, g);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ================================================================================ optimiser step
int tvae_adam_step(const tvae_adam_tensor* tensors, int n, double lr, double beta1, double beta2, double eps, double weight_decay,
                   int step, int zero_grad, void* stream) {
    TVAE_REQUIRE(tensors != nullptr && n >= 0, "adam: null tensor table");
    TVAE_REQUIRE(step >= 1, "adam: step is 1-based");
    TVAE_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && lr >= 0.0 && eps >= 0.0, "adam: invalid hyper-parameter");
    // bias corrections in double like the Python-float arithmetic of torch/optim/adam.py
    const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
    AdamHyper h{};
    h.beta1 = static_cast<float>(beta1); h.beta2 = static_cast<float>(beta2); h.eps = static_cast<float>(eps);
    h.weight_decay = static_cast<float>(weight_decay);
    h.step_size = static_cast<float>(lr / bc1);
    h.bc2_sqrt = static_cast<float>(sqrt(bc2));
    h.one_minus_beta1 = static_cast<float>(1.0 - beta1);
    h.one_minus_beta2 = static_cast<float>(1.0 - beta2);
    h.zero_grad = zero_grad;
    for (int i = 0; i < n;) {          // `i` is the consumed index: empty tensors are skipped, never re-visited
        AdamTable tab{};
        int blocks = 0;
        tab.n = 0;
        for (; i < n && tab.n < kAdamMaxTensors; ++i) {
            const tvae_adam_tensor& t = tensors[i];
            if (t.numel <= 0) continue;
            TVAE_REQUIRE(t.param && t.grad && t.exp_avg && t.exp_avg_sq, "adam: null tensor pointer");
            const int k = tab.n++;
            tab.param[k] = t.param; tab.grad[k] = t.grad; tab.m[k] = t.exp_avg; tab.v[k] = t.exp_avg_sq; tab.numel[k] = t.numel;
            tab.block_start[k] = blocks;
            blocks += static_cast<int>((t.numel + kAdamChunk - 1) / kAdamChunk);
        }
        tab.block_start[tab.n] = blocks;
        if (blocks == 0) continue;
        ++g_launch_count; adam_kernel<<<blocks, 256, 0, S(stream)>>>(tab, h);
        TVAE_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

int tvae_running_means(const float* elbo, const float* log_p, const float* kl, float b, float* state4, void* stream) {
    TVAE_REQUIRE(elbo && log_p && kl && state4, "running means: null pointer");
    ++g_launch_count; running_means_kernel<<<1, 32, 0, S(stream)>>>(elbo, log_p, kl, b, state4);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ================================================================================ particle-stack input pipeline
int tvae_ctf_filter(const double* params, int B, int n, int m, double scale, float* out, void* stream) {
    TVAE_REQUIRE(B >= 0 && n >= 1 && m >= 1 && n <= 255 && m <= 255, "ctf_filter: filter size must be in 1..255");
    TVAE_REQUIRE(scale > 0.0, "ctf_filter: scale must be positive");
    if (B == 0) return 0;                       // empty parameter table: nothing to do (pointers may be null)
    TVAE_REQUIRE(params && out, "ctf_filter: null pointer");
    const size_t sm = sizeof(double) * ((((size_t)n * m + 1) & ~(size_t)1) + 2 * (m + n + m));
    TVAE_REQUIRE(sm <= 227 * 1024, "ctf_filter: filter does not fit shared memory");
    TVAE_CHECK_CUDA(smem_optin(reinterpret_cast<const void*>(&ctf_filter_kernel), 227 * 1024));
    CtfFilterParams p{params, out, B, n, m, scale};
    ++g_launch_count; ctf_filter_kernel<<<B, 256, sm, S(stream)>>>(p);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tvae_mrc_crop_normalize(const void* in, int mode, int B, int n, int m, int crop, int normalize, float* out, void* stream) {
    TVAE_REQUIRE(B >= 0 && n >= 1 && m >= 1 && crop >= 0 && crop <= n && crop <= m, "crop_normalize: crop larger than the image");
    TVAE_REQUIRE(mode == 0 || mode == 1 || mode == 2 || mode == 6, "crop_normalize: MRC mode must be 0 (int8), 1 (int16), 2 (float32) or 6 (uint16)");
    if (B == 0) return 0;
    TVAE_REQUIRE(in && out, "crop_normalize: null pointer");
    const int c0 = crop > 0 ? crop : n, c1 = crop > 0 ? crop : m;
    const int si = crop > 0 ? (n - crop) / 2 : 0, sj = crop > 0 ? (m - crop) / 2 : 0;
    cudaStream_t st = S(stream);
    ++g_launch_count;
    switch (mode) {
        case 0: crop_normalize_kernel<signed char><<<B, 256, 0, st>>>(static_cast<const signed char*>(in), out, n, m, c0, c1, si, sj, normalize); break;
        case 1: crop_normalize_kernel<short><<<B, 256, 0, st>>>(static_cast<const short*>(in), out, n, m, c0, c1, si, sj, normalize); break;
        case 6: crop_normalize_kernel<unsigned short><<<B, 256, 0, st>>>(static_cast<const unsigned short*>(in), out, n, m, c0, c1, si, sj, normalize); break;
        default: crop_normalize_kernel<float><<<B, 256, 0, st>>>(static_cast<const float*>(in), out, n, m, c0, c1, si, sj, normalize); break;
    }
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tvae_crop_normalize(const float* in, int B, int n, int m, int crop, int normalize, float* out, void* stream) {
    return tvae_mrc_crop_normalize(in, 2, B, n, m, crop, normalize, out, stream);
}

// ================================================================================ standalone module interfaces
int tvae_fourier_embed_fwd(const float* x, const float* w_scaled, const float* b, float* out, long long M, int E, void* stream) {
    TVAE_REQUIRE(M >= 0 && E >= 1, "fourier_embed: bad shape");
    if (M == 0) return 0;
    TVAE_REQUIRE(x && w_scaled && b && out, "fourier_embed: null pointer");
    ++g_launch_count; fourier_embed_fwd_kernel<<<blocks_for(M * E, 256), 256, 0, S(stream)>>>(x, w_scaled, b, out, M, E);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int tvae_fourier_embed_bwd(const float* x, const float* w_scaled, const float* b, const float* g, float* dx, long long M, int E, void* stream) {
    TVAE_REQUIRE(M >= 0 && E >= 1, "fourier_embed: bad shape");
    if (M == 0) return 0;
    TVAE_REQUIRE(x && w_scaled && b && g && dx, "fourier_embed: null pointer");
    ++g_launch_count; fourier_embed_bwd_kernel<<<cdiv(M, 8), 256, 0, S(stream)>>>(x, w_scaled, b, g, dx, M, E);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tvae_linear_act_fwd(const float* x, const float* w, const float* bias, int M, int N, int K, int resid, int act, float* y,
                        void* x16, void* w16, void* stream) {
    TVAE_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 4 == 0 && N <= 1024, "linear_act: K must be a multiple of 8, N a multiple of 4 (<= 1024)");
    TVAE_REQUIRE(!resid || N == K, "linear_act: a residual layer is square");
    TVAE_REQUIRE(act == TVAE_ACT_LEAKYRELU || act == TVAE_ACT_TANH, "linear_act: act must be TVAE_ACT_LEAKYRELU or TVAE_ACT_TANH");
    TVAE_REQUIRE(x && w && y && x16 && w16, "linear_act: null pointer");
    cudaStream_t st = S(stream);
    ++g_launch_count; to_half_kernel<<<blocks_for((long long)M * K, 256), 256, 0, st>>>(x, static_cast<__half*>(x16), (long long)M * K);
    ++g_launch_count; weight_to_half_kernel<<<blocks_for((long long)N * K, 256), 256, 0, st>>>(w, static_cast<__half*>(w16), nullptr, N, K, resid);
    TVAE_CHECK_CUDA(cudaGetLastError());
    LinearNTArgs a{};
    a.A = x16; a.lda = K; a.B = w16; a.ldb = K; a.M = M; a.N = N; a.K = K; a.C = y; a.ldc = N; a.bias = bias;
    a.act = act == TVAE_ACT_TANH ? kActTanh : 1;
    return linear_nt(a, st);
}

int tvae_linear_act_bwd(const void* x16, const float* w, const float* y, const float* g, int M, int N, int K, int resid, int act,
                        void* dpre16, void* wt16, float* scales8, float* dx, float* dw, float* db, void* stream) {
    TVAE_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 8 == 0 && N <= 1024 && K <= 1024, "linear_act: bad shape");
    TVAE_REQUIRE(x16 && w && y && g && dpre16 && wt16 && scales8 && dw && db, "linear_act: null pointer");
    cudaStream_t st = S(stream);
    const int kact = act == TVAE_ACT_TANH ? kActTanh : 0;
    TVAE_CHECK_CUDA(cudaMemsetAsync(scales8, 0, sizeof(float) * 8, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, st));
    TVAE_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * N * K, st));
    ++g_launch_count; actgrad_absmax_kernel<<<blocks_for((long long)M * N, 256), 256, 0, st>>>(g, y, (long long)M * N, kact, scales8 + 7);
    ++g_launch_count; single_scale_kernel<<<1, 1, 0, st>>>(scales8 + 7, scales8);
    const int rows_per_cta = static_cast<int>(((long long)M + 8LL * sm_count() - 1) / (8LL * sm_count()));
    ++g_launch_count;
    actgrad_to_half_colsum_kernel<<<cdiv(M, rows_per_cta), 256, N * sizeof(float), st>>>(g, y, static_cast<__half*>(dpre16), scales8, db, M, N, rows_per_cta, kact);
    TVAE_CHECK_CUDA(cudaGetLastError());
    int rc;
    // dW[n][k] = sum_m dpre[m][n] x[m][k]
    if ((rc = linear_tn(dpre16, N, x16, K, M, N, K, dw, K, 0, st, scales8 + 3))) return rc;
    if (dx) {
        // dx = dpre (W + I): B operand = (W + I)^T as [K][N]
        ++g_launch_count; weight_to_half_kernel<<<blocks_for((long long)N * K, 256), 256, 0, st>>>(w, nullptr, static_cast<__half*>(wt16), N, K, resid);
        TVAE_CHECK_CUDA(cudaGetLastError());
        LinearNTArgs a{};
        a.A = dpre16; a.lda = N; a.B = wt16; a.ldb = N; a.M = M; a.N = K; a.K = N; a.C = dx; a.ldc = K; a.acc_scale = scales8 + 3;
        if ((rc = linear_nt(a, st))) return rc;
    }
    return 0;
}

int tvae_groupconv_dgrad(const tvae_enc_shape* s, const float* weight, const float* dout, float* bank32, float* dy, void* stream) {
    int rc = check_enc_shape(s);
    if (rc) return rc;
    TVAE_REQUIRE(weight && dout && bank32 && dy, "groupconv_dgrad: null pointer");
    TVAE_REQUIRE(s->O % 8 == 0, "groupconv_dgrad: kernel count must be a multiple of 8");
    const ConvGeom g = make_geom(s);
    cudaStream_t st = S(stream);
    ++g_launch_count;
    filter_bank_f32_kernel<<<blocks_for((long long)g.G * g.O * g.K, 256), 256, 0, st>>>(weight, bank32, g.O, g.C, g.k, g.G, make_rot_table(g.G));
    ++g_launch_count;
    groupconv_dgrad_kernel<<<dim3(g.n, g.B * g.C), 256, 0, st>>>(dout, bank32, dy, g.C, g.n, g.k, g.p, g.G, g.O, g.d);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tvae_attn_softmax_pair_bwd(const float* q_t_r, const float* a_sampled, const float* d_q, const float* d_a, float* d_attn, int B, int L,
                               void* stream) {
    TVAE_REQUIRE(B >= 1 && L >= 1 && q_t_r && a_sampled && d_attn && (d_q || d_a), "softmax_pair_bwd: bad arguments");
    ++g_launch_count; softmax_pair_bwd_kernel<<<B, 1024, 0, S(stream)>>>(q_t_r, a_sampled, d_q, d_a, d_attn, L);
    TVAE_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ================================================================================ test hooks
#ifdef TVAE_PROBE
// development probe of the CTA-pair kernel (tc_gemm2.cuh): read and clear the 8 x 16 counters (synchronises the device)
extern "C" int tvae_probe_read(unsigned long long* out128) {
    TVAE_CHECK_CUDA(cudaDeviceSynchronize());
    TVAE_CHECK_CUDA(cudaMemcpyFromSymbol(out128, tvae::g_pair_probe, sizeof(unsigned long long) * 128));
    unsigned long long zero[128] = {};
    TVAE_CHECK_CUDA(cudaMemcpyToSymbol(tvae::g_pair_probe, zero, sizeof(zero)));
    return 0;
}
#endif
int tvae_test_linear_nt(const void* A, const void* B, float* C, int M, int N, int K, const float* bias, int act,
                        void* stream) {
    LinearNTArgs a{};
    a.A = A; a.lda = K; a.B = B; a.ldb = K; a.M = M; a.N = N; a.K = K; a.C = C; a.ldc = N; a.bias = bias; a.act = act;
    return linear_nt(a, S(stream));
}

// full epilogue of the hidden-layer GEMM: fp16 output (x *store_scale), bias, activation, derivative mask from aux16,
// accumulator scale, column sums, fused projection
int tvae_test_linear_nt_full(const void* A, const void* B, int M, int N, int K, const float* bias, int act, void* C16,
                             const void* aux16, int aux_act, const float* acc_scale, const float* store_scale, float* colsum,
                             const float* proj_w, const float* proj_bias, float* proj_out, int n_proj, void* stream) {
    LinearNTArgs a{};
    a.A = A; a.lda = K; a.B = B; a.ldb = K; a.M = M; a.N = N; a.K = K; a.bias = bias; a.act = act;
    a.C16 = C16; a.ldc16 = N; a.aux16 = aux16; a.ld_aux = N; a.aux_act = aux_act;
    a.acc_scale = acc_scale; a.store_scale = store_scale; a.colsum = colsum; a.colsum_stride = 1;
    a.proj_w = proj_w; a.proj_bias = proj_bias; a.proj_out = proj_out; a.n_proj = n_proj;
    return linear_nt(a, S(stream));
}

// A/B switch for the wide weight-gradient kernel (default on): 0 sends it back to the tc_gemm LinearTN policy
void tvae_test_set_fast_paths(int pair_tn) { g_linear_pair_enabled = pair_tn != 0; }
// development knobs for A/B measurements (0 = default heuristics): 0 = conv1 weight-gradient reduction splits
void tvae_test_set_knob(int id, int value) { if (id >= 0 && id < 8) g_dev_knob[id] = value; }

int tvae_test_linear_tn(const void* P, const void* Q, float* C, int R, int Ma, int Nb, int transpose_out, void* stream) {
    return linear_tn(P, Ma, Q, Nb, R, Ma, Nb, C, transpose_out ? Ma : Nb, transpose_out, S(stream));
}

}  // extern "C"
