/*
 * gs_b200.h — C ABI of the B200-native 3D Gaussian-splatting rasterizer.
 *
 * Drop-in boundary for the path ComfyUI-3D-Pack reaches through the third-party
 * Python package `diff_gaussian_rasterization` (ashawkey fork), whose two
 * torch ops `rasterize_gaussians` / `rasterize_gaussians_backward` are what
 *   MVs_Algorithms/GaussianSplatting/main_3DGS_renderer.py:840-936   (render)
 *   MVs_Algorithms/GaussianSplatting/main_3DGS.py:205               (loss.backward)
 *   Gen_3D_Modules/LGM/core/gs.py:57-84, TriplaneGaussian/models/renderer.py:209-304,
 *   Gen_3D_Modules/TRELLIS/trellis/renderers/gaussian_render.py:62-137
 * end up calling.  The reference binds that package by Python import name;
 * this header is the C-level surface a binding (ctypes / pybind / cgo) links.
 * Plain pointers and sizes only — no torch types.  All pointers are DEVICE
 * pointers on the current CUDA device unless a function name ends in `_host`.
 * Every function returns 0 on success, non-zero on error (message via
 * gs_b200_last_error()).  Work is enqueued on `stream` (a cudaStream_t passed
 * as void*; NULL = legacy default stream).
 */
#ifndef GS_B200_H
#define GS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_B200_ABI_VERSION 1
#define GS_B200_TILE 16

/* ---- view / settings ------------------------------------------------------
 * Mirrors GaussianRasterizationSettings, the 12-field NamedTuple built at
 * main_3DGS_renderer.py:849-862.  bg/viewmatrix/projmatrix/campos stay device
 * pointers exactly as the reference passes CUDA tensors (no host sync).
 *   viewmatrix = w2c^T   row-major 4x4   (camera_utils.py:205)
 *   projmatrix = w2c^T @ P^T row-major    (camera_utils.py:213)
 */
typedef struct gs_b200_view {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    const float* bg;          /* [3]  */
    float scale_modifier;
    const float* viewmatrix;  /* [16] */
    const float* projmatrix;  /* [16] */
    int32_t sh_degree;
    const float* campos;      /* [3]  */
    int32_t prefiltered;      /* accepted, ignored (reference passes False) */
    int32_t debug;            /* accepted; non-zero adds a stream sync + error check per stage */
} gs_b200_view;

/* ---- allocator callback ----------------------------------------------------
 * Replaces the std::function<char*(size_t)> resize callbacks of the package's
 * RasterizeGaussiansCUDA: the caller owns every buffer (SURVEY §8b ownership).
 * tag says what the buffer is for; the callback returns >=256-byte aligned
 * device memory valid on `stream`.  If `alloc` is NULL the library uses
 * cudaMallocAsync on the stream and the caller releases saved state with
 * gs_b200_state_free().
 */
enum { GS_B200_BUF_GEOM = 0, GS_B200_BUF_BINNING = 1, GS_B200_BUF_IMAGE = 2, GS_B200_BUF_SCRATCH = 3 };
typedef void* (*gs_b200_alloc_fn)(void* user, int32_t tag, size_t bytes);

/* ---- state kept between forward and backward -------------------------------
 * (geomBuffer / binningBuffer / imgBuffer of the package.)  Filled by
 * gs_b200_rasterize_forward; pointers live in buffers obtained through the
 * callback, so freeing those buffers frees the state.
 */
typedef struct gs_b200_state {
    void* geom;              /* N x 48 B splat records {px,py,depth,radius | conic a,b,c,opacity | r,g,b,tiles} */
    uint32_t* point_list;    /* [P] Gaussian index per (tile,splat) pair, sorted by (tile, depth, index) */
    uint32_t* tile_keys;     /* [P] tile id per sorted pair */
    uint32_t* ranges;        /* [tiles][2] start,end into point_list */
    uint32_t* n_contrib;     /* [H*W] 1-based position of the last blended splat */
    float* final_T;          /* [H*W] transmittance after the last blended splat */
    int64_t num_rendered;    /* P */
    int32_t num_gaussians;   /* N */
    int32_t tiles_x, tiles_y;
    void* owned[4];          /* non-NULL only when the internal allocator was used */
} gs_b200_state;

int32_t gs_b200_abi_version(void);
const char* gs_b200_last_error(void);

/* Tile culling (no reference counterpart; an optimisation the caller can switch off).  The package's binning lists
 * every tile of each splat's 3-sigma bounding square (duplicateWithKeys); with culling on, tiles in which the
 * splat cannot reach alpha >= 1/255 at any pixel are left out of the list.  Rendered images are bit-identical and
 * gradients equal up to summation order; only point_list / ranges / n_contrib (list positions) differ from the
 * package's.  mode 0: off everywhere; 1 (default): gs_b200_step_* entries only; 2: also rasterize_forward.
 * Initial value can be set with the environment variable GS_B200_TILE_CULLING. */
int32_t gs_b200_set_tile_culling(int32_t mode);
int32_t gs_b200_get_tile_culling(void);

/* Measurement switch (no reference counterpart): which tile-composite kernels run.  0 (default): the round-2
 * kernels (gs_composite.cu: two pixels per lane in packed f32x2, cp.async-fed mbarrier ring, queued gradient phase);
 * 1: the round-1 kernels (gs_render.cu), kept for A/B timing.  Results agree (images bit for bit).
 * Initial value from the environment variable GS_B200_COMPOSITE ("r1" selects 1). */
int32_t gs_b200_debug_set_composite(int32_t mode);

/* Forward — replaces `_C.rasterize_gaussians` as called from
 * GaussianRasterizer.forward (main_3DGS_renderer.py:927-936).
 *   M        number of SH coefficients per channel in `shs` ([N,M,3]); ignored when colors_precomp != NULL
 *   exactly one of shs / colors_precomp, exactly one of (scales,rotations) / cov3D_precomp
 * Outputs (caller-allocated): out_color[3,H,W], out_depth[1,H,W], out_alpha[1,H,W], radii[N] int32.
 */
int32_t gs_b200_rasterize_forward(
    const gs_b200_view* view, int32_t N, int32_t M,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
    const float* scales, const float* rotations, const float* cov3D_precomp,
    float* out_color, float* out_depth, float* out_alpha, int32_t* radii,
    gs_b200_alloc_fn alloc, void* alloc_user, gs_b200_state* state, void* stream);

/* Backward — replaces `_C.rasterize_gaussians_backward` (triggered by
 * loss.backward(), main_3DGS.py:205).  Gradient outputs are caller-allocated
 * and OVERWRITTEN (accumulate==0) or ADDED TO (accumulate!=0; used by the
 * multi-view optimisation step, which sums views in place):
 *   dL_dmeans3D[N,3] dL_dmeans2D[N,3] (= NDC-space screen gradient, z=0; consumed at main_3DGS_renderer.py:767-769)
 *   dL_dshs[N,M,3] | dL_dcolors[N,3]   dL_dopacities[N,1]
 *   dL_dscales[N,3], dL_drotations[N,4] | dL_dcov3D[N,6]
 * NULL gradient pointers for the unused alternative of each exclusive pair.
 */
int32_t gs_b200_rasterize_backward(
    const gs_b200_view* view, int32_t N, int32_t M,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
    const float* scales, const float* rotations, const float* cov3D_precomp,
    const int32_t* radii, const gs_b200_state* state,
    const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
    float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors, float* dL_dopacities,
    float* dL_dscales, float* dL_drotations, float* dL_dcov3D, int32_t accumulate,
    gs_b200_alloc_fn alloc, void* alloc_user, void* stream);

/* Release state buffers obtained with the internal allocator (alloc == NULL). */
int32_t gs_b200_state_free(gs_b200_state* state, void* stream);

/* Parity/debug: reconstruct the 64-bit sorted keys (tile<<32 | float_bits(depth))
 * the package's binning stage would hold, for bit-exact comparison. keys_out[P]. */
int32_t gs_b200_debug_sorted_keys(const gs_b200_state* state, uint64_t* keys_out, void* stream);

/* CUB-free onesweep LSD radix sort of (u32 key, u32 value) pairs on bits
 * [begin_bit,end_bit) — the replacement for cub::DeviceRadixSort::SortPairs in
 * the package's binning stage.  keys/vals are ping-pong pairs of n elements;
 * *result_in_alt receives 1 when the sorted data ended in keys_alt/vals_alt.
 * scratch must hold gs_b200_sort_scratch_bytes(n) bytes. vals may be NULL (keys only). */
size_t gs_b200_sort_scratch_bytes(int64_t n);
int32_t gs_b200_sort_pairs_u32(uint32_t* keys, uint32_t* keys_alt, uint32_t* vals, uint32_t* vals_alt,
                               int64_t n, int32_t begin_bit, int32_t end_bit, void* scratch,
                               int32_t* result_in_alt, void* stream);

/* distCUDA2 replacement (simple_knn._C.distCUDA2, main_3DGS_renderer.py:408,419):
 * mean squared distance to the 3 nearest neighbours. points[N,3] -> out[N]. */
int32_t gs_b200_knn_mean_dist2(const float* points, int32_t N, float* out, void* stream);

/* Multi-view optimisation step with DEVICE-resident buffers: V x (forward + backward), gradients
 * summed over views into `grads` (zeroed first), packed as
 *   means3D[N,3] | shs[N,M,3] | opacities[N] | scales[N,3] | rotations[N,4] | means2D[N,3]
 * (the buffer a data-parallel caller all-reduces).  Views are software-pipelined over two internal
 * streams with persistent workspaces; all work is ordered after `stream` and `stream` continues after it.
 *   views_host / views_dev: the same V x 40 floats (layout below) in host and device memory
 *   dL_dout: device V x [5,H,W]; images: optional device V x [5,H,W] output (NULL = not kept) */
int32_t gs_b200_step_device(
    int32_t V, int32_t H, int32_t W, int32_t sh_degree, float scale_modifier, const float* views_host,
    const float* views_dev, int32_t N, int32_t M, const float* means3D, const float* shs, const float* opacities,
    const float* scales, const float* rotations, const float* dL_dout, float* grads, float* images,
    int64_t* num_rendered_out, void* stream);

/* Multi-view optimisation step with HOST buffers (the e2e entry: H2D of the
 * Gaussians, V x (forward + backward) with gradients summed over views on the
 * device, D2H of the summed gradients).  Host pointers should be pinned.
 *   views_host: V consecutive records of 40 floats:
 *      [0..15] viewmatrix, [16..31] projmatrix, [32..34] campos, [35..37] bg, [38] tanfovx, [39] tanfovy
 *   dL_dout_host: V x [5,H,W] (3 colour planes, depth, alpha) upstream gradients
 *   grads_host:   [N,(3+3M+1+3+4)] packed per parameter group in the order
 *                 means3D | shs | opacities | scales | rotations  (each contiguous), then means2D [N,3]
 *   images_host:  optional V x [5,H,W] rendered (colour, depth, alpha); may be NULL
 * Upload order: views, geometry parameters, SH block, upstream gradients view by view.  For V <= 8 the step starts
 * as soon as the geometry is resident (projection, sorting and binning of every view need nothing else) and fills the
 * colours into the records when the SH block has landed; images are bit-identical to gs_b200_step_device.
 * GS_B200_HOST_LATE_SH=0 in the environment makes it wait for the whole upload first.
 */
int32_t gs_b200_step_host(
    int32_t V, int32_t H, int32_t W, int32_t sh_degree, float scale_modifier, const float* views_host,
    int32_t N, int32_t M, const float* means3D_host, const float* shs_host, const float* opacities_host,
    const float* scales_host, const float* rotations_host, const float* dL_dout_host,
    float* grads_host, float* images_host, int64_t* num_rendered_out, void* stream);
/* Same step, but the summed gradients are left on the device: copied (D2D) into
 * grads_dev[N*(11+3M)+3N] instead of going to the host, so a multi-GPU caller can
 * all-reduce them before its own D2H.  grads_host may then be NULL. */
int32_t gs_b200_step_host_dev_grads(
    int32_t V, int32_t H, int32_t W, int32_t sh_degree, float scale_modifier, const float* views_host,
    int32_t N, int32_t M, const float* means3D_host, const float* shs_host, const float* opacities_host,
    const float* scales_host, const float* rotations_host, const float* dL_dout_host,
    float* grads_dev, float* images_host, int64_t* num_rendered_out, void* stream);

/* Same step with the LOSS inside the pipeline (the reference computes its loss between render and backward,
 * main_3DGS.py:176-205): after view v's forward has been enqueued on an internal stream, `hook(user, v, stream)` is
 * called on the host; it must enqueue ON THAT STREAM the work that reads images[v] ([5,H,W]: rgb, depth, alpha)
 * and writes dL_dout[v]; view v's backward is ordered after it.  Non-zero return aborts the step. */
typedef int32_t (*gs_b200_view_hook)(void* user, int32_t view_index, void* stream);
int32_t gs_b200_step_device_hook(
    int32_t V, int32_t H, int32_t W, int32_t sh_degree, float scale_modifier, const float* views_host,
    const float* views_dev, int32_t N, int32_t M, const float* means3D, const float* shs, const float* opacities,
    const float* scales, const float* rotations, float* dL_dout, float* grads, float* images,
    int32_t* radii /* optional V x [N] */, gs_b200_view_hook hook, void* hook_user, int64_t* num_rendered_out,
    void* stream);

/* The reference's training loss and its gradient for ONE view, on the device (GaussianSplatting3D.training,
 * main_3DGS.py:184-192; batch of one):
 *   loss = scale * [ (1-lambda_ssim) * L1(img*m, ref*m) + lambda_alpha * MSE(alpha, m)
 *                    + lambda_ssim * (1 - MS_SSIM(ref*m, img*m)) ],   img = clamp(rgb, 0, 1)
 * MS_SSIM as pytorch_msssim.MS_SSIM(data_range=1, size_average=True, channel=3) (11-tap Gaussian, 5 scales), skipped when
 * lambda_ssim == 0 as the reference does.  image: [5,H,W] (rgb | depth | alpha, what the rasterizer entries emit),
 * ref_image [3,H,W], ref_mask [1,H,W]; dL_dimage [5,H,W] (depth plane = 0); loss_out: 1 device float.
 * H and W must exceed 160 when lambda_ssim > 0 (the package's own assertion). */
int32_t gs_b200_image_loss(
    int32_t H, int32_t W, const float* image, const float* ref_image, const float* ref_mask, float lambda_ssim,
    float lambda_alpha, float scale, float* dL_dimage, float* loss_out, void* stream);

/* gs_b200_step_device with that loss evaluated inside the pipeline (per view: forward -> loss + gradient ->
 * backward), no host callback: ref_images V x [3,H,W], ref_masks V x [1,H,W]; loss_scale = 1 / (views in the
 * global batch); losses: V device floats (their sum over all ranks' views is the batch loss); dL_dout V x [5,H,W]
 * receives the upstream gradients that were used; images V x [5,H,W]; radii optional V x [N]. */
int32_t gs_b200_step_device_train(
    int32_t V, int32_t H, int32_t W, int32_t sh_degree, float scale_modifier, const float* views_host,
    const float* views_dev, int32_t N, int32_t M, const float* means3D, const float* shs, const float* opacities,
    const float* scales, const float* rotations, const float* ref_images, const float* ref_masks, float lambda_ssim,
    float lambda_alpha, float loss_scale, float* dL_dout, float* grads, float* images, int32_t* radii, float* losses,
    int64_t* num_rendered_out, void* stream);

/* Data-parallel hook (no reference counterpart): overlap the gradient all-reduce with the tail of the step.
 * The last pass of every gs_b200_step_device* call (chain rule from per-view splat gradients to the packed parameter
 * gradients) is cut into `nchunks` Gaussian ranges; after the kernel of range [first, first+count) has been enqueued
 * on `stream`, `sink(user, first, count, stream)` is called on the host.  Rows first..first+count of every group of
 * `grads` are final once the work enqueued on `stream` so far has run, so the callee can start the all-reduce of those
 * six slices (ordered after an event on `stream`) while the next range is still being computed.
 * Per calling thread; sink == NULL removes it.  Applies to gs_b200_step_device, _hook and _train. */
typedef void (*gs_b200_grad_sink)(void* user, int32_t first, int32_t count, void* stream);
int32_t gs_b200_set_grad_sink(gs_b200_grad_sink sink, void* user, int32_t nchunks);

/* Forward only over V views that share the Gaussians (render nodes: orbit previews, LGM / TRELLIS style multi-view
 * renders — nodes.py:1130-1163, Gen_3D_Modules/LGM/core/gs.py:41-92): one pass over the parameters for all
 * views, then the per-view pipeline.  Exactly one of shs [N,M,3] / colors_precomp [N,3] (LGM passes colours,
 * core/gs.py:75-84).  images: V x [5,H,W]; radii: optional V x [N] int32 (NULL = not kept). */
int32_t gs_b200_render_views(
    int32_t V, int32_t H, int32_t W, int32_t sh_degree, float scale_modifier, const float* views_host,
    const float* views_dev, int32_t N, int32_t M, const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, const float* rotations, float* images, int32_t* radii,
    int64_t* num_rendered_out, void* stream);

/* ---- optimisation step around the rasterizer (GaussianModel, main_3DGS_renderer.py) -------------
 * Raw (pre-activation) parameters live in ONE packed buffer laid out like the gradient buffer:
 *   xyz[N,3] | shs[N,M,3] (dc = coefficient 0) | opacity[N] | scaling[N,3] | rotation[N,4]
 * gs_b200_activate: exp / sigmoid / normalize of :293-321 (means and SHs are used as they are).
 * gs_b200_adam_step: chain rule through those activations + torch.optim.Adam (betas, eps=1e-15 as :446) with the
 *   six per-group learning rates {xyz, f_dc, f_rest, opacity, scaling, rotation} (host array), `step` = 1-based
 *   Adam step, grad_scale multiplies the incoming gradient (1/world for averaged data-parallel gradients).
 *   grads_packed holds gradients wrt the ACTIVATED values (what gs_b200_step_device produces).
 * gs_b200_densify_stats: add_densification_stats (:767-769) + max_radii2D update (main_3DGS.py:212) for radii > 0. */
int32_t gs_b200_activate(int32_t N, const float* raw_opacities, const float* raw_scales, const float* raw_rotations,
                         float* opacities, float* scales, float* rotations, void* stream);
int32_t gs_b200_adam_step(int32_t N, int32_t M, const float* lrs6_host, float beta1, float beta2, float eps, int32_t step,
                          float grad_scale, const float* grads_packed, float* params_packed, float* exp_avg,
                          float* exp_avg_sq, void* stream);
int32_t gs_b200_densify_stats(int32_t N, const float* dL_dmeans2D, const int32_t* radii, float* xyz_gradient_accum,
                              float* denom, float* max_radii2D, void* stream);

/* ---- densification as stream compaction in the packed layout -----------------------------------------
 * Replaces GaussianModel.densify_and_prune and the optimizer surgery behind it (main_3DGS_renderer.py:543-688,752-781).
 * gs_b200_densify_plan: classifies all N Gaussians (clone / split / prune rules of :641-668,752-781 on the raw
 *   parameters and the statistics of gs_b200_densify_stats) and scans the flags into destination rows.
 *   work: [8][N] u32 (flags, then offsets); counts_dev: 5 x u64 = survivors, kept clones, split parents, split parents
 *   whose children are kept, clones selected before the prune; scratch: gs_b200_densify_scratch_bytes(N).
 *   The caller reads counts_dev (the one host sync), sizes the destination for
 *   N' = survivors + kept clones + 2 x kept split parents and draws z ~ N(0,1) [2 x split parents, 3] (the draw
 *   torch.normal(mean=0, std=stds) makes in densify_and_split, :653, before scaling by std).
 * gs_b200_densify_apply: scatters parameters and both Adam moments from buffers packed for N rows into buffers packed
 *   for N' rows: survivors in order | kept clones | kept split children (first copies, then second copies); new rows
 *   get zero moments (cat_tensors_to_optimizer, :606-625); children = R(q/|q|) (exp(scaling) * z) + xyz with scaling
 *   log(exp(scaling) / 1.6).  Source and destination must not overlap. */
size_t gs_b200_densify_scratch_bytes(int32_t N);
int32_t gs_b200_densify_plan(int32_t N, const float* raw_opacity, const float* raw_scaling, const float* grad_accum,
                             const float* denom, float max_grad, float min_opacity, float extent, float percent_dense,
                             uint32_t* work, uint64_t* counts_dev, void* scratch, void* stream);
int32_t gs_b200_densify_apply(int32_t N, int32_t M, const float* src_raw, const float* src_m1, const float* src_m2,
                              const uint32_t* work, int32_t n_keep, int32_t n_clone, int32_t n_split_parents,
                              int32_t n_split_keep, const float* z, float* dst_raw, float* dst_m1, float* dst_m2, void* stream);

/* ---- instrumentation (bench.py): kernel-launch counter and per-stage CUDA-event timing ------
 * Stages: 0 preprocess, 1 depth sort, 2 scan, 3 emit, 4 tile sort, 5 ranges, 6 composite fwd,
 *         7 composite bwd, 8 preprocess bwd.  Events are recorded on the stream each stage is
 * launched on; gs_b200_profile_read synchronises those events and returns accumulated
 * milliseconds and call counts since the last gs_b200_profile_enable(1). */
#define GS_B200_NSTAGES 9
int64_t gs_b200_launch_count(void);
void gs_b200_profile_enable(int32_t on);
int32_t gs_b200_profile_read(float* ms_out, int32_t* calls_out);

#ifdef __cplusplus
}
#endif
#endif /* GS_B200_H */
