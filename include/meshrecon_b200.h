/*
 * meshrecon_b200.h -- C ABI of the B200-native dense-correspondence hot path of
 * addam/mesh-reconstruction (libmeshrecon_b200.so).
 *
 * This is the drop-in boundary: every entry point replaces one C++ interface of the
 * reference (cited as file:line into the upstream tree).  Signatures are OpenCV-free:
 * plain pointers and sizes.  The cv::Mat shim a maintainer adds on the reference side
 * is shown in INTEGRATION.md; the in-repo C++ mirror of recon.hpp lives in
 * mesh_reconstruction_b200/csrc/recon_b200.hpp.
 *
 * Conventions (same as the reference, recon.hpp / SURVEY.md 8b):
 *   - images are row-major, top-down, densely packed (stride == width * channels);
 *   - camera matrices are 4x4 float32 row-major, acting on column vectors, rows giving
 *     clip x, y, z, w (io_export_tracks.py:22-28,59-66);
 *   - depth maps hold NDC z, background == 1.0f exactly (recon.hpp:30);
 *   - a flow record is 4 floats (u, v, variance, 0) per pixel (flow.cpp:37-40);
 *   - a point row is 7 floats (x, y, z, w, nx, ny, nz) (recon.cpp:152-155).
 *
 * Every buffer argument may be a HOST pointer (pageable or pinned) or a DEVICE pointer
 * of the context's GPU; the library detects which (cudaPointerGetAttributes) and stages
 * host buffers through the context's stream.  Calls on one context are serialised by the
 * caller (the reference is single-threaded); different contexts are independent.
 *
 * Error convention: every function returns MR_OK (0) or a negative MR_E* code and never
 * calls exit(); mr_last_error() returns a human readable message.  There is NO CPU
 * fallback: without a usable CUDA device mr_create() fails with MR_ENODEVICE.
 */
#ifndef MESHRECON_B200_H
#define MESHRECON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MR_OK 0
#define MR_EINVAL (-1)    /* bad argument (null pointer, size mismatch, S out of range) */
#define MR_ENODEVICE (-2) /* no CUDA device / device index out of range */
#define MR_ECUDA (-3)     /* a CUDA runtime call or kernel failed; see mr_last_error */
#define MR_ENOMESH (-4)   /* a render call before mr_load_mesh */
#define MR_ENOMEM (-5)    /* device allocation failed */

#define MR_MAX_SIDE 16              /* max side cameras per main camera */
#define MR_BACKGROUND_DEPTH 1.0f    /* recon.hpp:30 */

typedef struct mr_context mr_context;

/* ---- lifetime ------------------------------------------------------------------
 * Replaces spawnRender(hint) / RenderGLX::RenderGLX (render_glx.cpp:57-62,152-208):
 * the render size is fixed per context (heuristic.cpp:548-551). */
int mr_create(mr_context **ctx, int device, int width, int height);
void mr_destroy(mr_context *ctx);                    /* RenderGLX::~RenderGLX render_glx.cpp:210-227 */
const char *mr_last_error(const mr_context *ctx);    /* ctx may be NULL: last creation error */
int mr_version(void);
/* The CUDA stream (cudaStream_t) all work of this context is enqueued on. */
void *mr_stream(mr_context *ctx);
/* Block until all enqueued work of the context finished. */
int mr_synchronize(mr_context *ctx);

/* ---- Render (recon.hpp:93-99) --------------------------------------------------- */
/* Render::loadMesh  render_glx.cpp:230-258.  vertices: V x 4 homogeneous float32,
 * faces: F x 3 int32 vertex indices. */
int mr_load_mesh(mr_context *ctx, const float *vertices_xyzw, int n_vertices, const int32_t *faces, int n_faces);
/* Render::depth  render_glx.cpp:369-397.  out: H*W float32 NDC z. */
int mr_depth(mr_context *ctx, const float camera[16], float *out_depth);
/* Batched single-pixel depth queries for Heuristic::chooseCameras / filterCameras (heuristic.cpp:285-341,
 * 456): for each of n_cameras "viewer" matrices (n_cameras*16 floats) the scene is rasterised on the device
 * and only out[i*n + j] = depth.at<float>(rows[i*n + j], cols[i*n + j]) is returned (rows/cols: n_cameras *
 * n_per_camera int32).  Same values as mr_depth() followed by host indexing; rows outside [0,H) or cols
 * outside [0,W] give MR_BACKGROUND_DEPTH, col == W reads the next row's first pixel like the reference's
 * continuous cv::Mat does. */
int mr_depth_samples(mr_context *ctx, const float *cameras, int n_cameras, const int32_t *rows, const int32_t *cols,
                     int n_per_camera, float *out);
/* Render::projected  render_glx.cpp:261-367 + shader.vert:9-13 + shader.frag:11-25.
 * frame: H*W uint8 (side camera's gray frame); out_rgb: H*W*3 uint8 (R = predicted gray,
 * G = B = 255 where visible and in-frame, else 0,0,0). */
int mr_projected(mr_context *ctx, const float camera[16], const uint8_t *frame, const float projector[16],
                 uint8_t *out_rgb);

/* ---- util.cpp / flow.cpp free functions (recon.hpp:40-55) ----------------------- */
/* mixBackground  util.cpp:366-387.  depth is IN/OUT (masked pixels are set to 1.0f). */
int mr_mix_background(mr_context *ctx, const uint8_t *image_rgb, const uint8_t *background, float *depth_inout,
                      uint8_t *out_mixed);
/* calculateFlow  flow.cpp:19-42.  out: H*W*4 float32 (u, v, variance, 0).
 * use_farneback == 0: cv::optflow VariationalRefinement (flow.cpp:29, the reference's default);
 * use_farneback != 0: cv::FarnebackOpticalFlow with the reference's parameters (flow.cpp:22-26). */
int mr_calculate_flow(mr_context *ctx, const uint8_t *prev, const uint8_t *next, int use_farneback, float *out_flow4);
/* flowRemap  util.cpp:390-403.  flow: H*W*stride_floats float32 with (u, v) first
 * (stride_floats = 2 for CV_32FC2 or 4 for the flow record); out: H*W uint8. */
int mr_flow_remap(mr_context *ctx, const float *flow, int stride_floats, const uint8_t *image, uint8_t *out);
/* compare  util.cpp:332-361.  out: H*W float32. */
int mr_compare(mr_context *ctx, const uint8_t *prev, const uint8_t *next, float *out);
/* imageGradient  util.cpp:465-479 (single-channel float input). out: H*W*2 float32. */
int mr_image_gradient(mr_context *ctx, const float *image, float *out_grad2);
/* triangulatePixels  util.cpp:167-329.  flows: n_side pointers to H*W*4 float32;
 * cameras: n_side*16 float32; out_points: capacity H*W*7 float32; *out_count = M rows,
 * in row-major pixel order. */
int mr_triangulate_pixels(mr_context *ctx, const float *const *flows, int n_side, const float main_camera[16],
                          const float *cameras, const float *depth, float *out_points, int *out_count);
/* extractCameraCenter  util.cpp:33-41 (host-side helper, dehomogenised). */
int mr_extract_camera_center(const float camera[16], float out_center3[3]);

/* ---- fused, device-resident main-frame step (recon.cpp:65-119) ------------------
 * One iteration of the reference's outer loop for main camera `main_camera`:
 *   depth = render->depth(main)                                   recon.cpp:70
 *   for each side i: projected -> mixBackground -> calculateFlow   recon.cpp:85-89
 *   triangulatePixels(flows, main, sides, depth)                   recon.cpp:114
 * without materialising intermediates on the host.  Results are identical to calling
 * the individual entry points in that order.  out_points (capacity H*W*7 floats, host or
 * device) may be NULL to keep the rows in the context (mr_points_device). */
int mr_process_main_frame(mr_context *ctx, const uint8_t *main_frame, const float main_camera[16], int n_side,
                          const uint8_t *const *side_frames, const float *side_cameras, float *out_points,
                          int *out_count);
/* The reference's `-f` switch (Configuration::useFarneback, configuration.cpp:26,92) for
 * mr_process_main_frame[_async]; off by default like the reference. */
int mr_set_use_farneback(mr_context *ctx, int on);
/* Pipelined variant for HOST (ideally pinned) out_points: returns as soon as the rows are complete on
 * the device and *out_count is known; their device->host copy into out_points runs on a second
 * stream and overlaps the NEXT call's compute (two internal row buffers are ping-ponged).  The caller
 * alternates between (at least) two host buffers; a buffer's contents are valid after mr_wait_copies()
 * or mr_synchronize().  With a device out_points it behaves exactly like mr_process_main_frame. */
int mr_process_main_frame_async(mr_context *ctx, const uint8_t *main_frame, const float main_camera[16], int n_side,
                                const uint8_t *const *side_frames, const float *side_cameras, float *out_points,
                                int *out_count);
/* Fully asynchronous variant: enqueues the whole main-frame step and returns WITHOUT waiting for the GPU, so the
 * host can queue many main frames back to back.  out_points and out_count (may be NULL) must be device memory
 * or PINNED host memory (cudaHostAlloc / cudaHostRegister; torch pin_memory), both in the same memory space.
 * Pinned buffers are filled by a copy-engine DMA of the full H*W*7-float capacity (the row count is not known on
 * the host without a synchronisation; only the first *out_count rows are meaningful) that overlaps the next
 * frames' compute.  Input frames in pinned host memory are uploaded asynchronously and, like the outputs, must
 * stay alive and untouched until mr_synchronize() (or mr_wait_copies() after the stream has drained). */
int mr_submit_main_frame(mr_context *ctx, const uint8_t *main_frame, const float main_camera[16], int n_side,
                         const uint8_t *const *side_frames, const float *side_cameras, float *out_points,
                         int *out_count);
/* mr_submit_main_frame can replay its launch sequence as one CUDA graph from the second submission of a given shape
 * (same n_side / flow method / host-vs-device inputs) on: one front-end launch per main frame instead of ~35, which
 * keeps the small kernels at full speed while the PCIe link is busy with the previous frame's rows.
 * mode 0 = never, 1 = when out_points is host memory or the Farneback branch is selected (default), 2 = always.  Results are identical in every mode.
 * mr_graph_launch_count: main frames replayed as a graph so far. */
int mr_set_use_graphs(mr_context *ctx, int mode);
uint64_t mr_graph_launch_count(const mr_context *ctx);
/* Block until every outstanding row copy of mr_process_main_frame_async / mr_submit_main_frame has landed. */
int mr_wait_copies(mr_context *ctx);
/* Block until at most max_in_flight of the row copies queued so far are still outstanding (they complete in
 * submission order): with a ring of R pinned output buffers per context, calling this with R - 1 before re-using a
 * buffer -- or with the number of frames queued since the results one wants to read -- lets the host consume one
 * main frame's rows while the following ones are still being computed and copied.  0 == mr_wait_copies. */
int mr_wait_copies_until(mr_context *ctx, int max_in_flight);
/* The path's one exchange step when main frames are sharded over GPUs, one rank per GPU (the reference is
 * single-process and appends every main frame's rows, recon.cpp:115-116): variable-length all-gather of point rows
 * over an NCCL communicator created by the host (nccl_comm is its ncclComm_t).  rows: this rank's `count` rows
 * (count x 7 floats, device memory).  out_rows (device, capacity out_capacity_rows rows -- at least the gathered
 * total on every rank) receives the rows of all ranks concatenated in RANK ORDER, which with contiguous blocks of
 * main frames per rank is the reference's append order; out_counts[world] / out_total (host, may be NULL) the
 * per-rank counts and their sum (64-bit: the cloud of a long clip exceeds 2^31 rows).  One small all-gather of every
 * rank's (count, capacity) -- so that a bad argument or a too small out_rows on ANY rank makes EVERY rank return
 * MR_EINVAL before the row exchange instead of leaving the others waiting in it (one stream synchronisation) -- then one
 * grouped NCCL operation of per-rank broadcasts with the exact counts (no padding), enqueued on mr_stream(ctx):
 * out_rows is complete after mr_synchronize().  NCCL is resolved at run time from the host process (or
 * libnccl.so.2); MR_ENODEVICE if it cannot be found. */
int mr_allgather_points(mr_context *ctx, void *nccl_comm, const float *rows, int count, float *out_rows,
                        size_t out_capacity_rows, int *out_counts, long long *out_total);
/* The same exchange WITHOUT kernels, for one process per GPU on an NVLink / NVSwitch node: every rank allocates a
 * receive buffer with one slot per rank (mr_xchg_alloc returns the device pointer and its 64-byte CUDA IPC handle),
 * the host passes the handles around (MPI, torch.distributed, a pipe ...), every rank maps its peers' buffers
 * (mr_xchg_open) and, per batch of main frames, DMAs its rows into its slot of every peer's buffer with mr_xchg_push:
 * copy-engine traffic over NVLink, ordered after the work queued so far on mr_stream(ctx), running on
 * mr_xchg_stream(ctx).  No SMs are used, so the exchange cannot slow the path's kernels down (a kernel-based
 * collective takes SMs from the 1-CTA-per-SM variational-refinement kernel).  Completion is the host's barrier:
 * enter it stream-ordered after mr_xchg_stream (or after cudaStreamSynchronize of it); when every rank has passed
 * it, every slot has landed.  The normals kernel can write a rank's own rows straight into its own slot (pass
 * the slot as out_points). */
int mr_xchg_alloc(mr_context *ctx, size_t bytes, void **dev_ptr, unsigned char ipc_handle[64]);
int mr_xchg_free(mr_context *ctx, void *dev_ptr);
int mr_xchg_open(mr_context *ctx, const unsigned char ipc_handle[64], void **peer_ptr);
int mr_xchg_close(mr_context *ctx, void *peer_ptr);
int mr_xchg_push(mr_context *ctx, void *peer_dst, const void *src, size_t bytes);
void *mr_xchg_stream(mr_context *ctx);
/* The same push through an NVSwitch MULTICAST mapping of the ranks' receive buffers (cuMulticastCreate / cuMulticastBindMem
 * by the host, e.g. torch.distributed's symmetric memory): one 128-bit multimem.st per 16 bytes, replicated by the switch
 * into every rank's buffer, so the rows leave the GPU once instead of world - 1 times.  A few CTAs on the high-priority
 * push stream (MR_MCAST_CTAS, default 32), ordered after mr_stream(ctx); completion as for mr_xchg_push.  bytes and
 * both pointers must be multiples of 16. */
int mr_xchg_push_mcast(mr_context *ctx, void *mcast_dst, const void *src, size_t bytes);
/* ---- frame ingest  configuration.cpp:226-245 -------------------------------------------------------
 * What Configuration does to every decoded frame before the path sees it: cv::resize(frame, Size(width, height),
 * CV_INTER_AREA) when the clip is larger than the render size (the -s scaling factor), then cv::cvtColor(CV_BGR2GRAY).
 * bgr: src_height x src_width x 3 uint8 (host -- ideally a pinned ring buffer the decoder writes into -- or device), at
 * least the context's size in both directions: integer factors take OpenCV's fast area path (integer box sums), anything
 * else (`-s 1.5`: the reference only warns when the size is not divisible, configuration.cpp:149-151) its general one
 * (float cell weights);
 * out_gray: H x W uint8, host or device (pass the device frame buffer mr_process_main_frame will read).  Arithmetic is
 * OpenCV's (integer box sums + its rounding, fixed-point gray), bit-identical to the cv2 binary.  Video decoding and
 * estimateExposure (off by default, configuration.cpp:25) stay on the host.  Asynchronous for device outputs
 * (enqueued on mr_stream); a host input must stay untouched until the stream has drained. */
int mr_ingest_frame(mr_context *ctx, const uint8_t *bgr, int src_width, int src_height, uint8_t *out_gray);
/* The same when Configuration::estimateExposure is on (configuration.cpp:270-426, off by default): its last step
 * (configuration.cpp:417-425) replaces the gray conversion by  frame = sum_c channel[c] * exposure[c]  in 8-bit cv::Mat
 * arithmetic (each product rounded and saturated to uchar, the additions saturating, channel order B, G, R).  The
 * exposure estimate itself -- a 100-iteration alternating least-squares over the ~20 bundle points of every frame
 * (3-column pseudo-inverses) -- is KB-sized host work and stays with the caller; exposure_bgr are its three weights for
 * this frame. */
int mr_ingest_frame_exposure(mr_context *ctx, const uint8_t *bgr, int src_width, int src_height, const float exposure_bgr[3],
                             uint8_t *out_gray);
/* BGR2GRAY coefficients: 15 (default) = the 15-bit ones of OpenCV >= 3.4.6 / 4.x (3735, 19235, 9798), 14 = the 14-bit
 * ones of OpenCV 3.0 - 3.4.5 (1868, 9617, 4899) -- the reference does not pin its OpenCV version (Makefile:10). */
int mr_set_gray_shift(mr_context *ctx, int shift);

/* ---- Heuristic::filterPoints  heuristic.cpp:55-176 ------------------------------------------------
 * The step that consumes the gathered cloud after every pass over the main frames (recon.cpp:125): outlier / redundancy
 * filter by local density -- radius-neighbour table restricted to j < i, clamped power iteration (up to 200 sweeps),
 * greedy thinning along descending density, compaction in ascending index order.  Running it on the device means that
 * only the SURVIVING rows have to cross PCIe for the CPU meshing step.
 *   points_xyzw: n x 4 homogeneous (dehomogenised like util.cpp:16-29), normals_xyz: n x 3 or NULL, host or device.
 *   radius: alphaVals.back() / 4 (heuristic.cpp:63).  NOTE the reference compares it with FLANN L2_Simple distances,
 *   which are SQUARED Euclidean distances; so does this function.
 *   out_points_xyzw / out_normals_xyz (capacity n rows) / out_keep (capacity n, the surviving indices, ascending): host
 *   or device, each may be NULL; *out_count = number of survivors.
 * Bit-identical to the CPU restatement (oracle/filter_oracle.cpp) including the double accumulations over all pairs.
 * Two things the reference leaves to its libraries are defined: the neighbour set is the
 * exact radius set (the reference's FLANN kd-tree search is randomised and approximate), and points of EQUAL density
 * are visited in descending index order (cv::sortIdx leaves their order to std::sort).
 * mr_filter_rows is the same for n x 7 point rows (x, y, z, w, nx, ny, nz) as produced by the path. */
int mr_filter_points(mr_context *ctx, const float *points_xyzw, const float *normals_xyz, size_t n, float radius,
                     float *out_points_xyzw, float *out_normals_xyz, int32_t *out_keep, size_t *out_count);
int mr_filter_rows(mr_context *ctx, const float *rows7, size_t n, float radius, float *out_rows7, int32_t *out_keep,
                   size_t *out_count);
/* Facts about the last mr_filter_points / mr_filter_rows of the context: info3 = neighbour pairs (j < i), power
 * iterations run, thinning rounds; density / score (n floats each, host or device, may be NULL) = the converged density
 * and the raw score of the last iteration (heuristic.cpp:104-138), for tests. */
int mr_filter_info(mr_context *ctx, long long info3[3], float *density, float *score);
/* Test hook: left-to-right double accumulation of n non-negative float terms, evaluated in parallel but bit-identical to
 * the sequential loop (the primitive behind the reference's `sum +=` / `change +=`, heuristic.cpp:117,133). */
int mr_debug_seqsum(mr_context *ctx, const float *terms, size_t n, double *out);

/* Device pointer to the point rows produced by the last mr_process_main_frame /
 * mr_triangulate_pixels (valid until the next call), and their count. */
const float *mr_points_device(mr_context *ctx, int *out_count);
/* Intermediates of the last mr_process_main_frame (device pointers, valid until the next
 * call): the final (mixed) depth, and the flow record of side camera i. For tests. */
const float *mr_last_depth_device(mr_context *ctx);
const float *mr_last_flow_device(mr_context *ctx, int side);
const uint8_t *mr_last_mixed_device(mr_context *ctx, int side);

/* Debug / benchmarking knob (process-wide): 0 = plane-per-stage variational-refinement
 * kernels, 1 = fused shared-memory tile kernel with TMA-staged frames (default; falls back to
 * plain loads when width % 16 != 0), 2 = fused kernel with plain loads.  Identical bits. */
int mr_set_vr_impl(int impl);

/* Per-stage device timing for bench.py (CUDA events on the context stream).  Stages:
 * 0 raster, 1 shade_mix, 2 variational_refinement, 3 cubic_remap, 4 pyramid_compare,
 * 5 triangulate, 6 normals.  mr_profile_read synchronises, returns the milliseconds and kernel
 * launches accumulated per stage since the last read, and resets them; returns the stage count. */
int mr_profile_enable(mr_context *ctx, int on);
int mr_profile_read(mr_context *ctx, double *ms_by_stage, uint64_t *launches_by_stage, int n);
const char *mr_stage_name(int stage);

/* Test hook: runs the host build of the kernels' 3x3 symmetric eigen-solver (cv::eigen restatement,
 * csrc/jacobi3.cuh).  cov6 = c00,c01,c02,c11,c12,c22; eigenvalues descending, eigenvectors in rows. */
int mr_debug_jacobi3(const float cov6[6], float evals3[3], float evecs9[9]);

/* Diagnostics of the window-PCA normals (util.cpp:250-326) since mr_create.  The covariance of a pixel's window is
 * evaluated from exact integer box sums for every coordinate whose values keep one sign and stay within a factor of two
 * over the pixel's 32x24 tile (+ halo), and with the reference's own sample-by-sample double accumulation for the other
 * coordinates; the rows are bit-identical to the reference's evaluation order either way (csrc/normals.cu).
 * out5[0] = tiles holding valid pixels; out5[1], out5[2], out5[3] = those with one, two, three coordinates on the
 * sample-by-sample route; out5[4] = pixels of the remaining entries that had to take that route as well. */
int mr_normals_stats(mr_context *ctx, uint64_t out5[5]);

/* Number of kernels this library launched on the context since creation (bench.py's
 * gpu_launches claim). */
uint64_t mr_launch_count(const mr_context *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MESHRECON_B200_H */
