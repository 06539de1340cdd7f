/*
 * include/sdfg.h -- C-ABI of libsdfg.so: the sm_100a kernels of the SDF-generator hot path
 * (per-ray-sample neural-field evaluation + volume rendering of SDFace-GAN's im2scene SDF generator).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it is marked "host";
 *   - the caller allocates everything; the library never allocates device memory and never synchronises;
 *   - every entry point takes the CUDA stream to launch on (`stream`, a cudaStream_t passed as void*) -- the
 *     reference extensions launch on the legacy default stream (gridencoder.cu:377-380, shencoder.cu:389), which is
 *     what this replaces;
 *   - return value: SDFG_OK (0) or a negative SDFG_ERR_*; sdfg_last_error() returns a thread-local message.  The
 *     Python shim turns non-zero into RuntimeError, as TORCH_CHECK does in the reference (gridencoder.cu:15-18);
 *   - re-entrant per device/stream (no global mutable state except the immutable SH coefficient table).
 *
 * "ref:" comments cite the reference interface each entry point replaces, relative to
 * /root/reference/im2scene/sdf/models/.
 */
#ifndef SDFG_H_
#define SDFG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFG_OK 0
#define SDFG_ERR_INVALID (-1)     /* bad argument (null pointer, zero size where not allowed, ...) */
#define SDFG_ERR_UNSUPPORTED (-2) /* shape / option outside what the kernels are built for */
#define SDFG_ERR_CUDA (-3)        /* a CUDA runtime call or kernel launch failed */

/* layouts of the hash-grid feature tensor */
#define SDFG_LAYOUT_NLC 0 /* [N, L*C]  sample-major (what GridEncoder.forward returns, grid.py:57) */
#define SDFG_LAYOUT_LNC 1 /* [L, N, C] level-major  (what grid_encode_forward writes, grid.py:47)  */

const char* sdfg_last_error(void);
int sdfg_version(void);
/* number of kernels launched by this library (all threads of the process) since the last reset (bench.py's gpu_launches) */
int64_t sdfg_launch_count(void);
void sdfg_launch_count_reset(void);
/* per-kernel device timing for the roofline line of bench.py: while enabled, every launch whose tag contains
 * `tag_substring` is bracketed by CUDA events on its own stream; collect() synchronises on them, returns the summed
 * milliseconds and the number of launches, and clears the list. */
void sdfg_prof_enable(int on, const char* tag_substring);
int sdfg_prof_collect(double* total_ms, int64_t* launches);

/* ---------------------------------------------------------------------------------------------------------------
 * Ray generation + depth sampling + point construction.
 * ref: VolumeFeatureRenderer.get_rays sdf_model.py:207-222, render :363-378, render_rays :310-351 (sampling part),
 *      mlp_init_pass :380-398 (stratified mode).
 *   c2w [B,3,4], focal/near/far [B]            camera (generate_camera_params, sdf_utils.py:97-159)
 *   t_vals [S]                                  the renderer's `t_vals` buffer (sdf_model.py:174-179): linspace(0,1-1/S,S)
 *                                               for offset sampling, linspace(0,1,S) for stratified
 *   t_rand  NULL (jitter_mode 0) | [B,R,R] (jitter_mode 1: one offset per ray, upper end = far, :326-331)
 *                                | [B,R,R,S] (jitter_mode 2: stratified between mid-points, :332-338 and :389-396)
 *   outputs (each may be NULL): z_vals [B,R,R,S]; pts, npts [B,R,R,S,3] (world / normalised by 2/(far-near));
 *                               viewdirs [B,R,R,3] (unit); rays_d [B,R,R,3]
 */
int sdfg_sample_rays(const float* c2w, const float* focal, const float* near, const float* far, const float* t_vals,
                     const float* t_rand, int jitter_mode, int static_viewdirs, int z_normalize,
                     uint32_t B, uint32_t R, uint32_t S,
                     float* z_vals, float* pts, float* npts, float* viewdirs, float* rays_d, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Multi-resolution hash grid.
 * ref: grid_encode_forward gridencoder/src/gridencoder.h:12, gridencoder.cu:448-469 (kernel_grid :87-245) and the
 *      affine map of GridEncoder.forward gridencoder/grid.py:149.
 *   inputs [N,D] f32; if bound > 0 the kernel first maps x -> (x + bound) / (2*bound) (grid.py:149), else x is in [0,1]
 *   embeddings [offsets[L], C] f32; offsets [L+1] i32 (device)
 *   outputs f32, layout SDFG_LAYOUT_*;  dy_dx NULL or [L, D, C, N] (component-major, so that a warp of samples writes and reads
 *   coalesced lines; the reference keeps [N, L*D*C], grid.py:50, and pays 32 partial sectors per store)
 *   S = log2(per_level_scale) as float, H = base resolution, gridtype 0 hash / 1 tiled, interp 0 linear / 1 smoothstep
 *   D in {2,3}; C in {1,2,4,8}; L <= 32.
 */
int sdfg_grid_encode_forward(const float* inputs, const float* embeddings, const int* offsets, float* outputs,
                             uint32_t N, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, float bound,
                             float* dy_dx, uint32_t gridtype, int align_corners, uint32_t interp, int out_layout,
                             void* stream);

/* ref: grid_encode_backward gridencoder.h:13, gridencoder.cu:472-503 (kernel_grid_backward :248-340,
 *      kernel_input_backward :343-369).
 *   grad (layout grad_layout) is scattered INTO grad_embeddings (accumulated: the caller pre-zeroes it, or passes a
 *   live .grad buffer to skip the reference's zeros_like + add, grid.py:77).
 *   grad_inputs NULL or [N,D] (overwritten; needs dy_dx).  If bound > 0, grad_inputs is already divided by 2*bound.
 */
int sdfg_grid_encode_backward(const float* grad, const float* inputs, const float* embeddings, const int* offsets,
                              float* grad_embeddings, uint32_t N, uint32_t D, uint32_t C, uint32_t L, float S,
                              uint32_t H, float bound, const float* dy_dx, float* grad_inputs, uint32_t gridtype,
                              int align_corners, uint32_t interp, int grad_layout, void* stream);

/* ref: grad_total_variation gridencoder.h:15, gridencoder.cu:612-645 (kernel_grad_tv :506-610). inputs in [0,1]. */
int sdfg_grad_total_variation(const float* inputs, const float* embeddings, float* grad, const int* offsets,
                              float weight, uint32_t N, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                              uint32_t gridtype, int align_corners, void* stream);

/* the per-level `scale = exp2f(level*S)*H - 1` table exactly as the device computes it (gridencoder.cu:138); the CPU
 * oracle takes it as an input because libm's exp2f may differ from CUDA's in the last bit.  out [L] f32 (device). */
int sdfg_grid_level_scales(float* out, uint32_t L, float S, uint32_t H, void* stream);

/* per-level integer cell coordinates and corner rows: the bit-exactness probe used by the parity tests.
 *   corner_idx [N, L, 2^D] u32 (table row within the level, 0xFFFFFFFF if the sample is out of bounds) */
int sdfg_grid_corner_indices(const float* inputs, const int* offsets, uint32_t* corner_idx, float* corner_w,
                             uint32_t N, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, float bound,
                             uint32_t gridtype, int align_corners, void* stream);

/* roofline probe (measurement aid, not part of the reference interface): `threads` threads each issue 8*rounds random 8-byte
 * gathers from buf[n_rows][2]; time it with events to get the device's random-gather rate over a table-sized buffer. */
int sdfg_l2_gather_probe(const float* buf, uint32_t n_rows, uint32_t threads, uint32_t rounds, float* sink, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Spherical harmonics of the view direction.
 * ref: sh_encode_forward shencoder/src/shencoder.h:9, shencoder.cu:400-416 (kernel_sh :27-355);
 *      sh_encode_backward shencoder.h:10, shencoder.cu:419-438 (kernel_sh_backward :358-382).
 *   inputs [N,3]; outputs [N, degree^2]; dy_dx NULL or [N, 3, degree^2]; degree in 1..8
 *   backward: grad_inputs [N,3] += sum_ch grad[n,ch] * dy_dx[n,d,ch]   (accumulates; caller pre-zeroes)
 */
int sdfg_sh_encode_forward(const float* inputs, float* outputs, uint32_t N, uint32_t degree, float* dy_dx, void* stream);
int sdfg_sh_encode_backward(const float* grad, const float* dy_dx, float* grad_inputs, uint32_t N, uint32_t degree,
                            void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Style-modulated SIREN field (NGPSIRENGenerator sdf_model.py:1534-1592 / SirenGenerator :101-139).
 *
 * Network description (host struct, passed by pointer; all weight pointers are device pointers, row-major [out,in]):
 *   x_in [N, in_dim]  -> (optional) input_linear [W, in_dim] -> n_film FiLM-SIREN layers (first has K = in_dim if there is
 *   no input_linear, else W) -> sdf head [1,W]; cat(h, view feature [rays, view_dim] broadcast over the S samples of a
 *   ray) -> views FiLM-SIREN [W, W+view_dim] -> rgb head [3,W].
 *   FiLM: h' = sin(gamma * (W h + b) + beta), gamma/beta [B, W] per image (FiLMSiren.forward :61-69); the image of
 *   sample n is n / samples_per_image.
 */
#define SDFG_MAX_FILM 9
typedef struct {
    uint32_t width;              /* W (256) */
    uint32_t in_dim;             /* 32 (hash features) or 3 (raw points) */
    uint32_t view_dim;           /* 16 (SH degree 4) or 3 (raw dirs) */
    uint32_t n_film;             /* trunk FiLM layers: 3 (ngp) or 8 (siren) */
    uint32_t has_input_linear;   /* 1 for ngp */
    uint32_t samples_per_image;  /* R*R*S */
    uint32_t samples_per_ray;    /* S */
    uint32_t reserved;
    const float* input_w;        /* [W, in_dim] or NULL */
    const float* input_b;        /* [W] */
    const float* film_w[SDFG_MAX_FILM];   /* trunk layers 0..n_film-1, then the views layer at index n_film */
    const float* film_b[SDFG_MAX_FILM];
    const float* gamma;          /* [B, n_film+1, W]  (15*Lin(w)+30, already evaluated) */
    const float* beta;           /* [B, n_film+1, W]  (0.25*Lin(w)) */
    const float* sigma_w;        /* [1, W] */
    const float* sigma_b;        /* [1] */
    const float* rgb_w;          /* [3, W] */
    const float* rgb_b;          /* [3] */
} sdfg_field_params;

/* gradient sinks, same shapes as the parameters; all accumulated INTO (caller pre-zeroes). NULL = not wanted. */
typedef struct {
    float* input_w;
    float* input_b;
    float* film_w[SDFG_MAX_FILM];
    float* film_b[SDFG_MAX_FILM];
    float* gamma;                /* [B, n_film+1, W] */
    float* beta;
    float* sigma_w;
    float* sigma_b;
    float* rgb_w;
    float* rgb_b;
} sdfg_field_grads;

/* bytes of activation workspace sdfg_field_forward needs; `save_for_backward` keeps every layer's pre-activation. */
uint64_t sdfg_field_workspace_bytes(const sdfg_field_params* p, uint64_t N, int save_for_backward, int precision);

#define SDFG_PRECISION_FP32 0   /* SIMT fp32 FMA path: parity <= 1e-3 max-abs with the reference fp32 path */
#define SDFG_PRECISION_TC16 1   /* tcgen05 path: fp16 operands (activations, weights, loss-scaled gradients), fp32 accumulate in
                                   TMEM (sm_100a); needs width == 256 and samples_per_image % 128 == 0 */

/* forward.  x_in [N,in_dim]; view_feat [N/S, view_dim]; out_sdf [N]; out_rgb [N,3] (NULL ok); out_feat [N,W] (NULL ok);
 * workspace as sized above (kept by the caller until backward). */
int sdfg_field_forward(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N,
                       float* out_sdf, float* out_rgb, float* out_feat, void* workspace, int save_for_backward,
                       int precision, void* stream);

/* inference variant of the tensor-core path (SDFG_PRECISION_TC16 constraints apply): the [N,W] feature output leaves the chip
 * as fp16 (written by TMA straight from the last layer's operand tile), half the HBM bytes of the fp32 copy; consumed by
 * sdfg_composite_forward_h.  No save_for_backward.  Replaces the same torch op chain as sdfg_field_forward
 * (sdf_model.py:1566-1592). */
int sdfg_field_forward_h(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N,
                         float* out_sdf, float* out_rgb, uint16_t* out_feat16, void* workspace, void* stream);

/* backward.  d_sdf [N], d_rgb [N,3], d_feat [N,W] (each NULL = zero; at least one given) -> parameter grads (g NULL =
 * none wanted, e.g. for the eikonal pass) + d_x_in [N,in_dim] (NULL ok).  Reads the workspace written by the matching
 * forward (save_for_backward = 1), `out_feat` = the pointer that forward was given (NULL if none) and uses `scratch`
 * (sdfg_field_backward_scratch_bytes) for the two [N,W] gradient ping-pong buffers. */
uint64_t sdfg_field_backward_scratch_bytes(const sdfg_field_params* p, uint64_t N, int precision);
int sdfg_field_backward(const sdfg_field_params* p, const sdfg_field_grads* g, const float* x_in, const float* view_feat,
                        uint64_t N, const float* d_sdf, const float* d_rgb, const float* d_feat, const float* out_feat,
                        const void* workspace, void* scratch, float* d_x_in, int precision, void* stream);

/* two-call variant: `phases` selects SDFG_BWD_CHAIN (loss scale, gradient chain
 * through the layers -> d_x_in, and the gradient tiles the second phase reads from `scratch`), SDFG_BWD_WGRAD (parameter
 * gradients into g) or both.  A data-parallel caller runs CHAIN, enqueues the hash-table scatter of d_x_in, starts the
 * all-reduce of the table gradient and then calls WGRAD with the same arguments: the exchange overlaps the weight-gradient
 * kernels, and scatter and contractions never contend for L2.  `g` must be the same in both calls (NULL = no parameter
 * gradients at all).  Paths whose two halves interleave (fp32 path, per-layer tensor-core kernels) do all their work in the call
 * that carries SDFG_BWD_CHAIN and nothing in a WGRAD-only call. */
#define SDFG_BWD_CHAIN 1
#define SDFG_BWD_WGRAD 2
#define SDFG_BWD_BOTH 3
int sdfg_field_backward_phase(const sdfg_field_params* p, const sdfg_field_grads* g, const float* x_in, const float* view_feat,
                              uint64_t N, const float* d_sdf, const float* d_rgb, const float* d_feat, const float* out_feat,
                              const void* workspace, void* scratch, float* d_x_in, int precision, int phases, void* stream);

/* eikonal pass (get_eikonal_term sdf_model.py:224-229 through _grid_encode.backward's grad_inputs, grid.py:77-93 / kernel_input_backward
 * gridencoder.cu:343-369): the trunk-only backward with upstream d_sdf [N] and no parameter gradients, contracted ON CHIP with the hash
 * encoder's dy_dx (sdfg_grid_encode_forward, [L * D * C, N] component-major, D = 3, C = 2, in_dim = L * C):
 *     d_pts[n, d] += scale * sum_{l, c} d_x_in[n, l * C + c] * dy_dx[(l * D + d) * C + c, n]
 * d_pts [N, 3] must be zeroed by the caller; d_x_in [N, in_dim] is never written.  Tensor-core path on shapes the two-tile chain takes
 * (SDFG_ERR_UNSUPPORTED otherwise: run sdfg_field_backward with d_x_in and sdfg_grid_encode_backward with grad_inputs instead). */
int sdfg_field_eikonal(const sdfg_field_params* p, const float* x_in, const float* view_feat, uint64_t N, const float* d_sdf,
                       const void* workspace, void* scratch, const float* dy_dx, uint32_t D, uint32_t C, float scale, float* d_pts,
                       int precision, void* stream);

/* probe of the tcgen05 pipeline for the parity tests: out[M,N] (fp32) = fp16(x)[M,K] * fp16(w)[N,K]^T with fp32 accumulation.
 * N multiple of 32 in 32..256, K <= 320. */
uint64_t sdfg_tc_linear_probe_workspace_bytes(uint32_t M, uint32_t K, uint32_t N);
int sdfg_tc_linear_probe(const float* x, const float* w, float* out, uint32_t M, uint32_t K, uint32_t N, void* workspace,
                         void* stream);

/* probe of the sample-axis (MN-major) weight-gradient contraction: G [B, 256, *ldg] (pre-zeroed, B = N / rows_per_image)
 * G[b,j,k<Kx] = sum_{n in image b} x16(dz)[n,j] * x16(x)[n,k];  G[b,j,*ones_col] = sum_n x16(dz)[n,j].  x_fmt (both operands): 0 fp16
 * (what the backward uses), 1 bf16.
 * workspace >= 2*N*(256 + Kx + 8) + 512 bytes. */
int sdfg_tc_wgrad_probe(const float* dz, const float* x, float* G, uint32_t N, uint32_t Kx, uint32_t rows_per_image, uint32_t x_fmt,
                        uint32_t* ldg, uint32_t* ones_col, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * SDF -> density -> alpha -> front-to-back compositing.
 * ref: VolumeFeatureRenderer.sdf_activation sdf_model.py:231-234 + volume_integration :236-301.
 *   sdf [NR,S]  (or raw density when with_sdf = 0); rgb [NR,S,3]; feat [NR,S,F] or NULL; z_vals [NR,S];
 *   rays_d [NR,3]; pts [NR,S,3] or NULL (needed for xyz); noise [NR,S] or NULL (non-sdf branch raw_noise)
 *   sigmoid_beta: device pointer to the learnable scalar (sdf_model.py:163-164)
 *   outputs: rgb_map [NR,3]; feat_map [NR,F] or NULL; xyz_map [NR,3] or NULL; mask [NR] or NULL;
 *            weights [NR,S] or NULL (kept for backward)
 *   rgb == rgb_map == NULL: sdf-only query (sdf_mesh.py's surface pass needs xyz / mask / sdf only).
 */
int sdfg_composite_forward(const float* sdf, const float* rgb, const float* feat, const float* z_vals,
                           const float* rays_d, const float* pts, const float* noise, const float* sigmoid_beta,
                           uint64_t NR, uint32_t S, uint32_t F, int with_sdf, int force_background,
                           float* rgb_map, float* feat_map, float* xyz_map, float* mask, float* weights, void* stream);

/* same, with fp16 features [NR,S,F] as written by sdfg_field_forward_h (feat16 and feat_map required). */
int sdfg_composite_forward_h(const float* sdf, const float* rgb, const uint16_t* feat16, const float* z_vals,
                             const float* rays_d, const float* pts, const float* noise, const float* sigmoid_beta,
                             uint64_t NR, uint32_t S, uint32_t F, int with_sdf, int force_background,
                             float* rgb_map, float* feat_map, float* xyz_map, float* mask, float* weights, void* stream);

/* backward of the above wrt sdf, rgb, feat, sigmoid_beta (accumulated into d_sigmoid_beta[0]) and pts (NULL ok).
 * d_* map gradients may be NULL (= zero). */
int sdfg_composite_backward(const float* sdf, const float* rgb, const float* feat, const float* z_vals,
                            const float* rays_d, const float* pts, const float* noise, const float* sigmoid_beta,
                            uint64_t NR, uint32_t S, uint32_t F, int with_sdf, int force_background,
                            const float* d_rgb_map, const float* d_feat_map, const float* d_xyz_map, const float* d_mask,
                            float* d_sdf, float* d_rgb, float* d_feat, float* d_pts, float* d_sigmoid_beta, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Frustum -> box resampling of an SDF volume for marching cubes.
 * ref: align_volume sdf_utils.py:164-184 (torch.meshgrid + F.grid_sample(padding_mode="border", align_corners=True) + the
 *      out-of-frustum fill with 1).  volume, out: [B, H, W, D, C] fp32 (the renderer's `sdf` output has C = 1).
 *      out[b,y,x,z,:] = trilinear sample of volume at (x', y', z) with x' = lin(x) * k(z), y' = lin(y) * k(z),
 *      k(z) = linspace(far/near, 1, D)[z], lin = linspace(-1, 1, .); cells with |x'| > 1 or |y'| > 1 are set to 1.
 */
int sdfg_align_volume(const float* volume, float* out, uint32_t B, uint32_t H, uint32_t W, uint32_t D, uint32_t C, float near_, float far_,
                      void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * StyleGAN2 decoder, forward (ref Decoder sdf_model.py:883-1056, ModulatedConv2d :614-704, StyledConv :793-818,
 * ToRGB :821-843, Upsample / Blur :480-538, fused_leaky_relu + upfirdn2d sdf_op.py).  Activations are channels-last fp16
 * [B, H, W, C]; the caller allocates everything.
 */
/* fp32 -> fp16, same element order (the renderer's feature map [B, H*W, C] is channels-last already) */
int sdfg_nhwc16(const float* in, uint16_t* out, uint64_t n_elems, void* stream);

/* per-sample weights of a ModulatedConv2d (:655-669): out[b][tap][o][i] = fp16(scale * weight[o][i][tap] * style[b][i] * demod[b][o]),
 * demod[b][o] = rsqrt(sum_{i,tap} (scale * weight * style)^2 + 1e-8) when `demodulate`, else 1.  weight [Cout, Cin, taps] (the
 * reference parameter [1, Cout, Cin, k, k]), style [B, Cin] (output of the layer's `modulation` EqualLinear), demod_scratch [B, Cout]. */
int sdfg_modconv_fold(const float* weight, const float* style, float scale, uint32_t B, uint32_t Cin, uint32_t Cout, uint32_t taps,
                      int demodulate, float* demod_scratch, uint16_t* out, void* stream);

/* 3 x 3 (taps = 9) / 1 x 1 (taps = 1) convolution, stride 1, zero padding, on per-sample weights wf [B, taps, Cout, Cin] (sdfg_modconv_fold),
 * fused with out = leaky_relu(conv + noise_w[0] * noise[b,y,x] + bias[o], 0.2) * sqrt(2) (NoiseInjection + FusedLeakyReLU);
 * out [B, H, W, Cout].  noise [B, H, W] / noise_w (device scalar) / bias [Cout] may be NULL.  Cin % 64 == 0, Cout % 64 == 0 (N tile 256 / 128 / 64), W a power of two >= 8.
 * ref ModulatedConv2d.forward :685-704 + StyledConv.forward :812-816. */
int sdfg_conv_forward(const uint16_t* x, const uint16_t* wf, uint32_t B, uint32_t H, uint32_t W, uint32_t Cin, uint32_t Cout, uint32_t taps,
                      const float* bias, const float* noise, const float* noise_w, uint16_t* out, void* stream);

/* up-sampling StyledConv: conv_transpose2d(x, wf, stride 2) -> Blur(outer([1,3,3,1]) / 16 * 4, pad (1,1)) -> + noise_w[0] * noise[b, Y, X] + bias[o]
 * -> leaky_relu(0.2) * sqrt(2).  x [B, H, W, Cin], wf [B, 9, Cout, Cin] (sdfg_modconv_fold), out [B, 2H, 2W, Cout],
 * t_scratch [B, 2H + 1, 2W + 1, Cout] fp16 (the transposed convolution before the blur).  noise [B, 2H, 2W].  Same shape limits as
 * sdfg_conv_forward.  ref ModulatedConv2d.forward :671-684 + StyledConv.forward :812-816. */
int sdfg_upconv_forward(const uint16_t* x, const uint16_t* wf, uint32_t B, uint32_t H, uint32_t W, uint32_t Cin, uint32_t Cout, uint16_t* t_scratch,
                        const float* bias, const float* noise, const float* noise_w, uint16_t* out, void* stream);

/* ToRGB (:821-843): 1 x 1 modulated convolution without demodulation to 3 channels + bias [3] + (skip != NULL) the previous level's
 * rgb [B, H/2, W/2, 3] up-sampled by Upsample (upfirdn2d up = 2, outer([1,3,3,1]) / 16, pad (2, 1)).  weight [3, C], style [B, C],
 * wrgb_scratch [B, C, 4] floats (16-byte aligned).  Outputs (either may be NULL): out_nhwc [B, H, W, 3] (skip of the next level), out_nchw [B, 3, H, W]. */
int sdfg_to_rgb(const uint16_t* x, const float* weight, const float* style, float scale, const float* bias, const float* skip, uint32_t B,
                uint32_t H, uint32_t W, uint32_t C, float* wrgb_scratch, float* out_nhwc, float* out_nchw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDFG_H_ */
