/* gsr.h -- C ABI of the B200-native differentiable 3DGS tile rasterizer (libgsr_b200.so).
 *
 * This is the drop-in seam for the torch-free core of the reference rasterizer,
 *   CudaRasterizer::Rasterizer::{forward, backward, markVisible}
 *   ($RAST/cuda_rasterizer/rasterizer.h:24-91, $RAST = submodules/gaustudio-diff-gaussian-rasterization),
 * which the reference's pybind layer ($RAST/rasterize_points.cu:35-231, $RAST/ext.cpp:15-19) binds.
 * Plain pointers and sizes only: no torch / glm / std:: types.  All float* are DEVICE pointers to
 * contiguous float32 (the reference's convention, rasterize_points.cu:97-117); optional inputs are NULL
 * when absent (the reference passes the data pointer of an empty tensor, i.e. nullptr).
 * Every entry point enqueues its work on `stream` (a cudaStream_t passed as void*); the reference uses
 * the legacy default stream everywhere (forward.cu:416,459; rasterizer_impl.cu:148,280,292,306,317).
 *
 * Return convention: >= 0 success, < 0 error (text via gsr_last_error()).  With `debug` != 0 each stage is
 * followed by a stream synchronize + error check, mirroring CHECK_CUDA (auxiliary.h:166-173).
 */
#ifndef GSR_H_INCLUDED
#define GSR_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_ABI_VERSION 1

/* Scratch allocator callback.  Replaces the three `std::function<char*(size_t)>` resize closures the
 * reference takes (rasterizer.h:35-37, built by resizeFunctional at rasterize_points.cu:27-33).
 * Must return a device pointer to at least `bytes` bytes (any alignment >= 16), valid until the matching
 * backward call has completed.  The buffers are opaque to the caller, exactly like the reference's
 * geomBuffer / binningBuffer / imgBuffer (gaustudio_diff_gaussian_rasterization/__init__.py:97,106). */
typedef char* (*gsr_alloc_fn)(void* user, size_t bytes);

int gsr_abi_version(void);
const char* gsr_last_error(void);

/* Sizes of the three opaque buffers (what the allocator callbacks will be asked for). */
size_t gsr_geometry_bytes(int P);
size_t gsr_image_bytes(int width, int height);
size_t gsr_binning_bytes(int64_t num_rendered);

/* Replaces Rasterizer::forward (rasterizer.h:34-60, impl rasterizer_impl.cu:198-343).
 * Same arguments in the same order, plus:
 *   r_capacity : 0  -> exact mode: one blocking 8-byte device->host read of num_rendered sizes the binning
 *                      buffer (the reference does the same at rasterizer_impl.cu:284);
 *                >0 -> pipelined mode: no host sync; the binning buffer is sized for r_capacity tile
 *                      instances; if the view needs more, nothing is rendered for the overflowing tiles and
 *                      *r_host (if given) receives the true count so the caller can detect it
 *                      (num_rendered > r_capacity) and re-run.
 *   r_host     : optional PINNED host int64 that asynchronously receives num_rendered (either mode).
 * Returns num_rendered (exact mode), r_capacity (pipelined mode), or < 0 on error. */
int64_t gsr_forward(gsr_alloc_fn geometry_alloc, void* geometry_user, gsr_alloc_fn binning_alloc,
                    void* binning_user, gsr_alloc_fn image_alloc, void* image_user, int P, int D, int M,
                    const float* background, int width, int height, const float* means3D, const float* shs,
                    const float* colors_precomp, const float* opacities, const float* scales,
                    float scale_modifier, const float* rotations, const float* cov3D_precomp,
                    const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                    float tan_fovy, int prefiltered, float* out_color, float* out_depth,
                    float* out_median_depth, float* out_opacity, int* radii, int debug, int64_t r_capacity,
                    int64_t* r_host, void* stream);

/* Replaces Rasterizer::backward (rasterizer.h:62-91, impl rasterizer_impl.cu:347-452).  Same arguments in
 * the same order (+ stream).  All ten outputs are fully written (no pre-zeroing needed, unlike the
 * reference which accumulates into torch::zeros tensors, rasterize_points.cu:160-169); dL_dconic and
 * dL_ddepth are internal in the reference's Python API and may be NULL here.
 * Shapes: dL_dmean2D[P,3] dL_dconic[P,2,2] dL_dopacity[P] dL_dcolor[P,3] dL_ddepth[P] dL_dmean3D[P,3]
 *         dL_dcov3D[P,6] dL_dsh[P,M,3] dL_dscale[P,3] dL_drot[P,4]. */
int gsr_backward(int P, int D, int M, int64_t R, const float* background, int width, int height,
                 const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
                 float scale_modifier, const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                 float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                 const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dpix_median_depth,
                 const float* dL_dpix_final_opacity, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                 float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                 float* dL_dscale, float* dL_drot, int debug, void* stream);

/* Fused-activation variants -- the step immediately before the path in every caller:
 * gaustudio's VanillaRenderer.get_gaussians_properties (gaustudio/renderers/vanilla_renderer.py:28-52) runs
 * exp(_scale), sigmoid(_opacity), normalize(_rot) and cat(_f_dc, _f_rest) as separate elementwise kernels on
 * every view (gaustudio/models/vanilla_sg.py:58-63,102-106; models/utils.py:6-32).  These entry points take the
 * model's RAW attributes instead -- log_scales[P,3], raw_rotations[P,4], opacity_logits[P], f_dc[P,1,3],
 * f_rest[P,M-1,3] -- and apply the activations inside the projection kernel; the backward returns gradients
 * w.r.t. the raw attributes (dL_dlog_scale, dL_draw_rot, dL_dopacity_logit, dL_df_dc, dL_df_rest).  Everything
 * else is identical to gsr_forward / gsr_backward; the opaque buffers are interchangeable. */
int64_t gsr_forward_fused(gsr_alloc_fn geometry_alloc, void* geometry_user, gsr_alloc_fn binning_alloc,
                          void* binning_user, gsr_alloc_fn image_alloc, void* image_user, int P, int D, int M,
                          const float* background, int width, int height, const float* means3D, const float* f_dc,
                          const float* f_rest, const float* opacity_logits, const float* log_scales,
                          float scale_modifier, const float* raw_rotations, const float* viewmatrix,
                          const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                          int prefiltered, float* out_color, float* out_depth, float* out_median_depth,
                          float* out_opacity, int* radii, int debug, int64_t r_capacity, int64_t* r_host,
                          void* stream);
int gsr_backward_fused(int P, int D, int M, int64_t R, const float* background, int width, int height,
                       const float* means3D, const float* f_dc, const float* f_rest, const float* opacity_logits,
                       const float* log_scales, float scale_modifier, const float* raw_rotations,
                       const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                       float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                       const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dpix_median_depth,
                       const float* dL_dpix_final_opacity, float* dL_dmean2D, float* dL_dopacity_logit,
                       float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_df_dc, float* dL_df_rest,
                       float* dL_dlog_scale, float* dL_draw_rot, int debug, void* stream);

/* Replaces Rasterizer::markVisible (rasterizer.h:27-32, impl rasterizer_impl.cu:54-66,141-153).
 * present: device bool[P] (1 byte each). */
int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present, void* stream);

/* Depth -> normal map (the "rendered normal" of gaustudio's vanilla path:
 * gaustudio/datasets/__init__.py:106-112,307-380, Camera.depth2point + Camera.depth2normal with k=3).
 * depth: device float[H*W]; out: device float[H*W*3]; rot: optional device float[9] (row-major 3x3, the
 * `inverse(extrinsics[:3,:3]).t()` of coordinate='world') or NULL for camera coordinates. */
int gsr_depth2normal(const float* depth, int width, int height, float fx, float fy, float cx, float cy,
                     float d_min, float d_max, const float* rot, float* out, void* stream);

/* Depth -> 3-D points (Camera.depth2point, gaustudio/datasets/__init__.py:307-339; the step after the path in
 * gaustudio/scripts/extract_mesh.py:95-115).  out: device float[H*W*3]; cam_to_world: optional device float[16]
 * (row-major inverse(extrinsics)) for coordinate='world', NULL for camera coordinates. */
int gsr_depth2point(const float* depth, int width, int height, float fx, float fy, float cx, float cy,
                    const float* cam_to_world, float* out, void* stream);

/* Introspection for parity tests: copies internal state of the last forward out of the opaque buffers into
 * caller-provided DEVICE arrays (any may be NULL):
 *   point_list  uint32[R]   Gaussian index per sorted tile instance (== BinningState::point_list)
 *   ranges      uint32[T*2] per-tile [start,end) (== ImageState::ranges; empty tiles are (0,0))
 *   n_contrib   uint32[H*W] final_T float[H*W]     (== ImageState::n_contrib / accum_alpha)
 *   means2D float[P*2] conic_opacity float[P*4] depths float[P] rgb float[P*3] cov3D float[P*6]
 *   tiles_touched uint32[P] clamped uint8[P*3]     (== GeometryState members, rasterizer_impl.h:33-47) */
int gsr_debug_export(int P, int width, int height, int64_t R, const char* geom_buffer,
                     const char* binning_buffer, const char* image_buffer, uint32_t* point_list,
                     uint32_t* ranges, uint32_t* n_contrib, float* final_T, float* means2D, float* conic_opacity,
                     float* depths, float* rgb, float* cov3D, uint32_t* tiles_touched, unsigned char* clamped,
                     void* stream);

/* Optional per-stage device timing (CUDA events recorded on the caller's stream around each kernel).
 * Stages: 0 preprocess_fwd, 1 tile_scan, 2 scatter, 3 tile_sort, 4 render_fwd, 5 render_bwd,
 *         6 preprocess_bwd, 7 depth2normal.  gsr_profile_read synchronises the recorded events, adds their
 * elapsed times (ms) into ms[GSR_NUM_STAGES] / launch counts into counts[GSR_NUM_STAGES] and clears them. */
#define GSR_NUM_STAGES 8
int gsr_profile_enable(int on);
int gsr_profile_read(float* ms, int* counts);

#ifdef __cplusplus
}
#endif
#endif /* GSR_H_INCLUDED */
