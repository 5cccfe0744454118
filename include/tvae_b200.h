/* tvae_b200.h - C ABI of libtvae_b200.so: the B200 (sm_100a) kernels behind the TARGET-VAE training hot path.
 *
 * The reference (SMLC-NYSBC/TARGET-VAE) has no FFI / plugin layer: its only stable boundary is the Python
 * surface of src/models.py and the trainers' eval_minibatch.  These entry points are what the
 * drop-in Python classes in target-vae_b200/src/models.py bind through ctypes; each comment names the reference
 * code the call replaces (file:line relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 data unless noted; `stream` is a cudaStream_t
 *   - calls never allocate, never synchronise, and are re-entrant; all scratch is passed in by the caller
 *   - return 0 on success, < 0 on error; tvae_last_error() returns the thread-local message
 *   - fp16 tensors (MMA operands, stored activations, scaled gradients) are passed as void*
 *
 * Precision: every dense contraction runs on the tensor cores with FP16 operands and FP32 accumulation
 * (tcgen05.mma.kind::f16).  An fp16 operand has the 11-bit significand of a TF32 one (the precision the reference's
 * cuDNN path uses); gradients are kept inside fp16's exponent range by power-of-two scales chosen on the device
 * from max|.| bounds and divided out (exactly) in the consuming kernel's epilogue.
 *
 * Internal activation layout (rows are (b, r, pos) with pos = i*W' + j, P = H'*W'):
 *   x1, h  : fp16 [(b*G + r)*P + pos][O]
 *   heads  : (B, NH, G, P) planar, NH = 3 + 2*z, channels = [attn, theta_mu, theta_logstd, z_mu.., z_logstd..]
 *            == attn (B,G,H',W'), theta (B,2,G,H',W'), z (B,2z,G,H',W') of models.py:403 stacked on dim 1
 */
#ifndef TVAE_B200_H
#define TVAE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* tvae_last_error(void);
int tvae_version(void);
long long tvae_launch_count(void);   /* kernels launched by this library so far (process-wide) */
/* Optional timing of every tensor-core GEMM launch with CUDA events on its own stream (used by bench.py for the
 * roofline of the dominant kernel).  collect() synchronises the device and returns the number of kernel names. */
void tvae_profile_enable(int on);
int tvae_profile_collect(const char** names, float* total_ms, int* launches, int cap);

/* ------------------------------------------------------------------ encoder (models.py:132-225, 326-403) */
/* --activation (train_mnist.py:423,516-519): the module-level activation of the encoder / generator */
#define TVAE_ACT_LEAKYRELU 0     /* nn.LeakyReLU(), negative slope 0.01 */
#define TVAE_ACT_TANH 1          /* nn.Tanh() */

typedef struct {
    int B, C, n, k, p, G, O, z;
    int kpad;            /* row pitch of the filter bank: multiple of 32 and > C*k*k (tvae_bank_pitch) */
    int act;             /* TVAE_ACT_* */
} tvae_enc_shape;

int tvae_bank_pitch(int C, int k);     /* row pitch (floats) of the fp32 bank GRADIENT dbank */
int tvae_bank16_pitch(int C, int k);   /* row pitch (halves) of the fp16 bank: C*k*k rounded up to 64 */

/* GroupConv.trans_filter (models.py:174-197): weight (O,C,1,k,k) -> bank fp16 [G*O][kpad16] (row r*O + o).
 * The conv GEMMs run kind::f16: an fp16 operand carries the same 11-bit significand as a TF32 one. */
int tvae_filter_bank_fwd(const tvae_enc_shape* s, const float* weight, void* bank, void* stream);
/* adjoint of the above; the conv1 bias gradient is read from column C*k*k of dbank (rows r*O + o, summed over r),
 * where the backward kernels deposit it: dbank [G*O][kpad] -> dweight (O,C,1,k,k), dbias (O).  Both overwritten. */
int tvae_filter_bank_bwd(const tvae_enc_shape* s, const float* dbank, float* dweight, float* dbias, void* stream);

/* GroupConv.forward alone (models.py:202-225): out [(b*G + r)*P + pos][O] = conv + bias, no activation.
 * bias may be NULL.  tvae_groupconv_wgrad: dout in the same layout -> dbank [G*O][kpad] (overwritten);
 * dout16 is scratch for the scaled fp16 copy of dout the GEMM consumes (B*G*P*O halves), scales8 scratch of 8 floats. */
int tvae_groupconv_fwd(const tvae_enc_shape* s, const float* y, const void* bank, const float* bias, float* out, void* stream);
int tvae_groupconv_wgrad(const tvae_enc_shape* s, const float* y, const float* dout, void* dout16, float* scales8, float* dbank, void* stream);

/* Executed / dense-count ratio of the K chunks of tvae_groupconv_fwd (wgrad = 0) or tvae_groupconv_wgrad (wgrad = 1):
 * chunks that only meet zero padding are skipped.  > 1 is possible for the wgrad (tile-grid padding of the kk rows).
 * Host-only arithmetic, used by bench.py to report the executed-MAC roofline next to the dense-count one. */
double tvae_conv1_executed_fraction(const tvae_enc_shape* s, int wgrad);

typedef struct {
    const float* y;          /* (B,C,n,n) */
    const void* bank;        /* fp16 [G*O][kpad16] from tvae_filter_bank_fwd */
    const float* conv1_bias; /* (O) */
    const float* w2;         /* conv2.weight (O,O) */
    const float* b2;         /* (O) */
    const float* wh;         /* [NH][O]: conv_a, conv_r, conv_z weights stacked */
    const float* bh;         /* [NH] */
    const float* head_add;   /* [NH][G]: p_r on channel 0, rotation offsets on channel 1, else 0 */
    void* x1;                /* out fp16 [B*G*P][O]  LeakyReLU(conv1) */
    void* h;                 /* out fp16 [B*G*P][O]  LeakyReLU(conv2); NULL = inference only (clustering_*.get_latent,
                              * clustering_mnist.py:81-161): the hidden map feeds the heads on chip and is not written */
    float* heads;            /* out (B,NH,G,P); (B,NH,1,P) with rotation pooling */
    void* w2_h;              /* scratch fp16 (O,O) */
    /* Rotation pooling of InferenceNetwork_AttentionTranslation_UnimodalRotation with groupconv > 0 (models.py:301-304):
     * fc_w != NULL inserts xp = fc_r(x1 over the rotation axis) between conv1 and conv2; conv2, the heads and h then have
     * B*P rows (one rotation slot) and head_add is [NH][1]. */
    const float* fc_w;       /* fc_r.weight (G) or NULL */
    const float* fc_b;       /* fc_r.bias (1) */
    void* xp;                /* out fp16 [B*P][O] */
} tvae_enc_fwd_args;
/* GroupConv.forward + InferenceNetwork_AttentionTranslation_AttentionRotation.forward up to the head maps
 * (models.py:202-225, 355-358, 382, 390-399). */
int tvae_encoder_fwd(const tvae_enc_shape* s, const tvae_enc_fwd_args* a, void* stream);

typedef struct {
    const float* y;
    const float* w2;
    const float* wh;
    const void* x1;          /* in: saved fp16 activation LeakyReLU(conv1) */
    const void* h;           /* in: saved fp16 activation LeakyReLU(conv2) */
    const float* d_heads;    /* (B,NH,G,P) */
    void* dhpre;             /* scratch fp16 [B*G*P][O]: d(conv2 pre-activation) * s1 */
    void* dx1_16;            /* scratch fp16 [B*G*P][O]: d(conv1 pre-activation) * s2, the wgrad GEMM's operand */
    void* w2t_h;             /* scratch fp16 (O,O) */
    float* scales;           /* scratch, 8 floats: the power-of-two scales s1, 1/s1, s2, 1/s2 (computed on the device) */
    float* dbank;            /* out [G*O][kpad] (feed to tvae_filter_bank_bwd) */
    float* dw2;              /* out (O,O) */
    float* db2;              /* out (O) */
    float* dwh;              /* out [NH][O] */
    float* dbh;              /* out [NH] */
    /* rotation pooling (fc_w != NULL, see tvae_enc_fwd_args): h, d_heads, dhpre have B*P rows */
    const float* fc_w;       /* fc_r.weight (G) or NULL */
    const void* xp;          /* in: saved fp16 [B*P][O] */
    void* dxp16;             /* scratch fp16 [B*P][O]: d(xp) * s2 */
    float* dfc_w;            /* out (G) */
    float* dfc_b;            /* out (1) */
} tvae_enc_bwd_args;
/* autograd of the above (convolution_backward x5, train_mnist.py:321). */
int tvae_encoder_bwd(const tvae_enc_shape* s, const tvae_enc_bwd_args* a, void* stream);

/* ------------------------------------------------------------------ attention inference + KL
 * models.py:383-387 (log_softmax, gumbel_softmax) and train_mnist.py:187-282. */
typedef struct {
    int B, G, d, z;          /* d = H' = W' */
    float s;                 /* pixel spacing x[1,0] - x[0,0] */
    float theta_prior_std;   /* pi / G */
    float offsets[16];       /* rotation offsets (0 when no refinement) */
} tvae_attn_shape;

/* log_prior[G*d*d] = log_softmax over (r,t) of [sum_xy N(0,0.1).log_prob(grid_t) + p_r[r]] (train_mnist.py:258-262) */
int tvae_attn_log_prior(const tvae_attn_shape* s, const float* p_r_host16, float* log_prior, void* stream);

typedef struct {
    const float* heads; const float* gumbel; const float* r_z; const float* r_theta; const float* log_prior;
    float* stats;            /* (B,4) max/lse of the two softmaxes */
    float* zb;               /* (B,z) sampled z */
    float* theta_b;          /* (B) sampled rotation */
    float* dx;               /* (B,2) expected translation */
    float* kl;               /* (B) per-image KL (val1 + val2) */
} tvae_attn_fwd_args;
int tvae_attn_fwd(const tvae_attn_shape* s, const tvae_attn_fwd_args* a, void* stream);

typedef struct {
    tvae_attn_fwd_args f;    /* the forward's inputs and outputs */
    const float* g_zb; const float* g_theta; const float* g_dx;   /* dLoss/d(zb, theta_b, dx) */
    const float* g_kl;       /* device scalar: dLoss/d(kl_b) (already divided by B) */
    float* d_heads;          /* out (B,NH,G,P) */
} tvae_attn_bwd_args;
int tvae_attn_bwd(const tvae_attn_shape* s, const tvae_attn_bwd_args* a, void* stream);

/* module-interface tail: q_t_r = log_softmax(attn), a_sampled = softmax(attn + gumbel) (models.py:383-388) */
int tvae_attn_softmax_pair(const float* heads, const float* gumbel, float* q_t_r, float* a_sampled, int B, int NH, int L, void* stream);
/* clustering_mnist.py:122-161: argmax (r,t), z/theta at the argmax, softmax-expected translation */
int tvae_get_latent(const tvae_attn_shape* s, const float* heads, float* z_content, float* theta_mu, float* dx, int* argmax, void* stream);

/* Argmax (rotation, translation) assignment at fp32 accuracy (clustering_mnist.py:127 `attn.view(B,-1).max(1)`, z / theta
 * gathered there, :140-161).  The tensor-core encoder's FP16-operand logits are used as a FILTER: every cell of the fast
 * attention map within rel_tol x (max - min) of its maximum is a candidate (at most TVAE_REFINE_MAX_CAND per image; the
 * band is halved until they fit), and conv1 -> act -> [fc_r] -> conv2 -> act -> heads is re-evaluated at those cells from
 * the fp32 image, the fp32 bilinear-rotated filters (models.py:174-197) and the fp32 weights, fp32 products summed in
 * double and rounded to fp32 where the reference's fp32 layers round.  Outputs overwrite those of tvae_get_latent:
 * argmax (B) cell index r*P + pos, z_content (B,2z) = [z_mu, exp(z_logstd)], theta_mu (B); the softmax-expected
 * translation stays the one of tvae_get_latent.  `weight` is conv1.weight (O,C,1,k,k) itself, not the fp16 bank. */
#define TVAE_REFINE_MAX_CAND 32
typedef struct {
    const float* y;          /* (B,C,n,n) */
    const float* weight;     /* conv1.weight (O,C,1,k,k) */
    const float* conv1_bias; /* (O) */
    const float* w2;         /* conv2.weight (O,O) */
    const float* b2;         /* (O) */
    const float* wh;         /* [NH][O] stacked head weights */
    const float* bh;         /* [NH] */
    const float* head_add;   /* [NH][G] ([NH][1] with rotation pooling) */
    const float* fc_w;       /* fc_r.weight (G) or NULL */
    const float* fc_b;       /* fc_r.bias (1) */
    const float* heads;      /* fast head maps (B,NH,G,P) from tvae_encoder_fwd */
    float rel_tol;           /* candidate band as a fraction of the map's range; 2e-3 is 4x the largest flip observed */
    float* bank32;           /* scratch fp32 [G*O][C*k*k] */
    int* cand;               /* scratch/out (B, TVAE_REFINE_MAX_CAND) candidate cells */
    int* n_cand;             /* scratch/out (B) */
    float* cand_heads;       /* scratch/out (B, TVAE_REFINE_MAX_CAND, NH) refined head values at the candidates */
    float* z_content;        /* out (B,2z) */
    float* theta_mu;         /* out (B) */
    int* argmax;             /* out (B) */
    float* refined_logit;    /* out (B) or NULL: the refined maximal logit */
} tvae_refine_args;
int tvae_refine_argmax(const tvae_enc_shape* s, const tvae_refine_args* a, void* stream);

/* ------------------------------------------------------------------ generator (models.py:53-58, 95-123) */
typedef struct {
    int B, N;                /* images, pixels per image (M = B*N rows) */
    int E;                   /* Fourier features (1024) or 0 when no expansion (coord_linear in_dim = 2) */
    int H;                   /* hidden width */
    int L;                   /* number of hidden Linear(H,H) layers = num_layers - 1 */
    int n_out, zdim;
    int act;                 /* TVAE_ACT_* */
} tvae_gen_shape;

typedef struct {
    const float* x;          /* (N,2) base coords, or (B*N,2) explicit coords when theta == NULL */
    const float* theta;      /* (B) or NULL */
    const float* dx;         /* (B,2) or NULL */
    const float* z;          /* (B,zdim) */
    const float* wf_scaled;  /* (E,2) embed_latent.weight / sigma */
    const float* bf;         /* (E) */
    const float* w1;         /* coord_linear.weight (H, E or 2) */
    const float* b1;         /* (H) */
    const float* wz;         /* latent_linear.weight (H,zdim) */
    const float* wh;         /* [L][H][H] hidden weights; ResidLinear layers (--generator-resid-layers, models.py:22-30:
                              * act(W x + b + x)) are passed as W + I - dwh is then the gradient w.r.t. W itself */
    const float* bh;         /* [L][H] */
    const float* wout;       /* (n_out,H) */
    const float* bout;       /* (n_out) */
    float* zb;               /* out (B,H) latent_linear(z) */
    void* acts;              /* out fp16 [L+1][B*N][H] post-activation of every hidden layer */
    float* y_hat;            /* out (B*N, n_out) */
    void* w_h;               /* scratch fp16: H*max(E,2) + L*H*H halves (fp16 copies of the weights) */
    void* mask_bits;         /* optional out, max(L,1) * (H/64) * B*N 64-bit words (LeakyReLU, H % 64 == 0): word (layer l, column
                              * block c, row m) holds the derivative mask of acts[l][m][64c .. 64c+63] (bit set = 1, clear = the
                              * LeakyReLU slope), written by the forward kernels; the backward pass then reads one bit per element
                              * instead of the fp16 activation in the hidden layers' input gradients.  NULL: masks from acts. */
} tvae_gen_fwd_args;
int tvae_generator_fwd(const tvae_gen_shape* s, const tvae_gen_fwd_args* a, void* stream);

typedef struct {
    tvae_gen_fwd_args f;
    const float* d_yhat;     /* (B*N, n_out) */
    void* dpre0; void* dpre1;     /* scratch fp16 [B*N][H] each: layer gradients times a power-of-two scale */
    void* wt_h;              /* scratch fp16: max(E*H, H*H) halves (transposed weights) */
    float* scales;           /* scratch, 32 floats: the per-layer scales (computed on the device) */
    float* dxp;              /* scratch/out (B*N,2): gradient w.r.t. the transformed coordinates */
    float* dzb;              /* scratch (B,H) */
    float* dw1; float* db1; float* dwz; float* dwh; float* dbh; float* dwout; float* dbout;  /* out, same shapes as weights */
    float* d_theta;          /* out (B)   (NULL when explicit coords: dxp is the coordinate gradient) */
    float* d_dx;             /* out (B,2) */
    float* d_z;              /* out (B,zdim) */
} tvae_gen_bwd_args;
int tvae_generator_bwd(const tvae_gen_shape* s, const tvae_gen_bwd_args* a, void* stream);

/* ------------------------------------------------------------------ likelihoods
 * Bernoulli (train_mnist.py:288-291, train_galaxy.py:288-292): ll[b] = -sum_e BCEWithLogits; d_yhat = g*(sigmoid - y).
 * g is a device scalar = dLoss/d(ll_b) * (-1) (i.e. 1/B for loss = -elbo). d_yhat may be NULL. */
int tvae_bernoulli(const float* y_hat, const float* y, float* ll, float* d_yhat, const float* g, int B, int E, void* stream);
/* Gaussian with optional CTF and mask (train_particles.py:298-338): mu = ctf (*) y_hat (scratch (B,n,n)),
 * ll[b] = -0.5 sum mask (mu - y)^2, d_yhat = adjoint-ctf(g * mask * (mu - y)).  ctf may be NULL, radius 0 = no mask.
 * ctf is (B, ctf_size, ctf_size), applied with zero padding ctf_size / 2 exactly like
 * F.conv2d(..., padding=ctf.size(2)//2, groups=B) (train_particles.py:301): ctf_size = n - 1 (the trainer's default;
 * 0 means n - 1) or any odd size (--crop keeps the filters of the uncropped micrograph size, train_particles.py:543-547).
 * With a ctf, y_hat may be NULL when `mu` still holds ctf (*) y_hat from the forward call (backward pass). */
/* ws: scratch of tvae_gaussian_workspace_bytes(B, n) bytes for the tensor-core CTF path (banded-Toeplitz GEMM with
 * fp16 operands; ctf_size = n - 1, even n <= 128).  ws == NULL, a size query of 0 or another filter size selects the
 * CUDA-core correlation kernel (the centred part of the filter that can meet the image must fit shared memory). */
long long tvae_gaussian_workspace_bytes(int B, int n);
int tvae_gaussian(const float* y_hat, const float* y, const float* ctf, int ctf_size, const float* dx, float s, int radius,
                  float* mu, float* dmu, float* ll, float* d_yhat, const float* g, int B, int n, void* ws, void* stream);

/* Gaussian with a learned per-pixel variance (--fit-noise: train_particles.py:289-296, 333-334, 663-666; generator
 * n_out = 2, no CTF, no mask - with either the reference itself fails for B > 1).  y_hat2 is the generator output
 * (B, N, 2) read the way the reference reads it: flattened to (B, 2N), first N values = mean, last N = log-variance.
 * ll[b] = -0.5 sum_e ((mu_e - y_e)^2 / exp(lv_e) + lv_e); d_yhat2 (same layout, may be NULL) = g * d ll_b / d y_hat2. */
int tvae_gaussian_fit_noise(const float* y_hat2, const float* y, float* ll, float* d_yhat2, const float* g, int B, int N,
                            void* stream);

/* ------------------------------------------------------------------ optimiser step and running statistics
 * torch.optim.Adam(params, lr).step() [+ zero_grad()] of train_mnist.py:323-324,579 for ALL generator and encoder
 * parameters in ONE launch (the reference's optimiser launches several kernels per parameter tensor).  Arithmetic of
 * torch/optim/adam.py::_single_tensor_adam (torch 2.11, amsgrad = False, maximize = False):
 *     g  = grad (+ weight_decay * p);  m += (g - m)(1 - beta1);  v = beta2 v + (1 - beta2) g g
 *     p -= (lr / (1 - beta1^step)) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 * `tensors` is a HOST array of n descriptors holding DEVICE pointers (fp32); `step` is the 1-based step count.
 * zero_grad != 0 also clears the gradients (optim.zero_grad(set_to_none=False)). */
typedef struct {
    float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
    long long numel;
} tvae_adam_tensor;
int tvae_adam_step(const tvae_adam_tensor* tensors, int n, double lr, double beta1, double beta2, double eps, double weight_decay,
                   int step, int zero_grad, void* stream);   /* doubles: 1 - beta, beta^step are formed like torch's Python floats */
/* Running means of train_mnist.py:326-338 kept on the device (the reference syncs three .item() per step):
 * state = {c, elbo_accum, gen_loss_accum, kl_loss_accum};  c += b;  x_accum += b (x - x_accum) / c  with
 * x = elbo, -log_p_x_g_z, kl_div (device scalars).  Zero the state to start an epoch. */
int tvae_running_means(const float* elbo, const float* log_p, const float* kl, float b, float* state4, void* stream);

/* ------------------------------------------------------------------ particle-stack input pipeline
 * src/ctf.py:32-55 `ctf_filter` (with `compute_2d_ctf`, src/ctf.py:6-23), called at train_particles.py:543-547:
 * params = DEVICE array (B, 8) of doubles in parse_ctf's column order {defocus [um], cs, voltage, apix, bfactor, ampcont,
 * dfdiff, dfang [deg]} -> out (B, n, m) fp32 = -fftshift(ifft2(CTF on the fftfreq grid / (apix * scale))).real.
 * fp64 arithmetic on the device (the reference computes in numpy float64), exact separable inverse DFT (n, m <= 255). */
int tvae_ctf_filter(const double* params, int B, int n, int m, double scale, float* out, void* stream);
/* src/image.py:30-42 `crop` (centre crop to `crop` x `crop`, 0 = none) followed by train_particles.py:592-600 --normalize
 * (per-image (x - mean) / std with the population std; normalize = 0 copies the crop): in (B, n, m) -> out (B, c, c). */
int tvae_crop_normalize(const float* in, int B, int n, int m, int crop, int normalize, float* out, void* stream);
/* The same on the image payload of an MRC / MRCS container as stored (src/mrc.py:108-140 `parse`: the array is
 * np.frombuffer(content[1024 + next:], dtype by header.mode) reshaped (nz, ny, nx)): `in` = DEVICE copy of the payload
 * bytes of B consecutive images - this rank's shard of the stack -, mode = 0 (int8), 1 (int16), 2 (float32), 6 (uint16);
 * decoded, cropped and standardised in one pass (values widened to double like numpy does for an integer stack).
 * The 1024-byte header is parsed on the host (tvae_b200/mrc.py). */
int tvae_mrc_crop_normalize(const void* in, int mode, int B, int n, int m, int crop, int normalize, float* out, void* stream);

/* ------------------------------------------------------------------ standalone module interfaces of src/models.py
 * (calls a user can make outside the fused step; not on the timed path)
 * RandomFourierEmbedding2d.forward (models.py:53-58): out (M,E) = cos(x (M,2) . w_scaled^T + b), w_scaled = weight / sigma;
 * backward: dx (M,2) from g (M,E). */
int tvae_fourier_embed_fwd(const float* x, const float* w_scaled, const float* b, float* out, long long M, int E, void* stream);
int tvae_fourier_embed_bwd(const float* x, const float* w_scaled, const float* b, const float* g, float* dx, long long M, int E, void* stream);
/* ResidLinear.forward (models.py:29-30): y (M,N) = act(x (M,K) . W^T + bias [+ x]) on the LinearNT tensor-core kernel with fp16
 * operands (resid = 1: W + I, N == K; resid = 0: a plain Linear + activation).  x16 (M*K halves) and w16 (N*K halves) are
 * scratch; x16 is the saved operand of the backward call.  Backward: g (M,N) = dLoss/dy -> dx (M,K) (may be NULL), dw (N,K),
 * db (N); dpre16 (M*N halves), wt16 (N*K halves) and scales8 (8 floats) are scratch. */
int tvae_linear_act_fwd(const float* x, const float* w, const float* bias, int M, int N, int K, int resid, int act, float* y,
                        void* x16, void* w16, void* stream);
int tvae_linear_act_bwd(const void* x16, const float* w, const float* y, const float* g, int M, int N, int K, int resid, int act,
                        void* dpre16, void* wt16, float* scales8, float* dx, float* dw, float* db, void* stream);
/* GroupConv.forward's gradient w.r.t. its input image (models.py:202-225 under autograd; the training step never needs it -
 * the image is data): dout fp32 [(b*G + r)*P + pos][O] (the layout tvae_groupconv_fwd writes) -> dy (B,C,n,n).  bank32 is
 * scratch, G*O*C*k*k floats (the fp32 rotated bank).  CUDA-core kernel. */
int tvae_groupconv_dgrad(const tvae_enc_shape* s, const float* weight, const float* dout, float* bank32, float* dy, void* stream);
/* backward of tvae_attn_softmax_pair (models.py:383-388 under autograd): d_attn (B,L) from d_q and / or d_a (either may be NULL) */
int tvae_attn_softmax_pair_bwd(const float* q_t_r, const float* a_sampled, const float* d_q, const float* d_a, float* d_attn, int B, int L,
                               void* stream);

/* ------------------------------------------------------------------ test hooks for the GEMM core (fp16 operands, fp32 out)
 * nt: C[M,N] = act(A[M,K] B[N,K]^T + bias);  tn: C[Ma,Nb] (or its transpose) += sum_r P[r,Ma] Q[r,Nb], C zero-filled */
int tvae_test_linear_nt(const void* A, const void* B, float* C, int M, int N, int K, const float* bias, int act, void* stream);
int tvae_test_linear_tn(const void* P, const void* Q, float* C, int R, int Ma, int Nb, int transpose_out, void* stream);
/* A/B switch (process-wide, for measurements and tests): 0 sends wide weight gradients back to the tc_gemm LinearTN policy
 * instead of the CTA-pair kernel (linear_pair_policies.cuh).  Default: 1. */
void tvae_test_set_fast_paths(int pair_tn);
void tvae_test_set_knob(int id, int value);   /* development knobs for A/B measurements; 0 = default heuristics */
/* the hidden-layer GEMM with its whole epilogue: C16[M,N] fp16 = store_scale * mask(aux16) * act(acc_scale * A B^T + bias);
 * colsum[N] += column sums, proj_out[M,n_proj] (pre-zeroed) += value . proj_w^T (+ proj_bias).  Any pointer may be NULL. */
int tvae_test_linear_nt_full(const void* A, const void* B, int M, int N, int K, const float* bias, int act, void* C16,
                             const void* aux16, int aux_act, const float* acc_scale, const float* store_scale, float* colsum,
                             const float* proj_w, const float* proj_bias, float* proj_out, int n_proj, void* stream);

#ifdef __cplusplus
}
#endif
#endif
