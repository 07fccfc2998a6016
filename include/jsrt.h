/* libjsrt — C ABI of the B200-native render path for alitteneker/jsraytracer.
 *
 * The reference has no FFI or plugin interface: its render
 * path is `renderer.render(img, timelimit, callback, x_offset, x_delt)`
 * (src/renderers.js:10,70) called from a web worker (src/worker.js:30-32) that
 * then posts `img.imgdata` back (src/worker.js:31,37).  A `CUDARenderer` class
 * next to those renderers binds the entry points below (N-API addon under Node,
 * ctypes in this repository's Python host mirror); INTEGRATION.md shows both.
 *
 * Conventions: plain pointers and sizes only; the caller owns every host
 * buffer; the library owns device memory behind the opaque handles.  Functions
 * returning int return 0 on success and non-zero on failure, with the message
 * available from jsrt_last_error() on the calling thread (the reference's
 * convention is `throw "<string>"`, e.g. src/aggregates.js:39; the binding turns
 * the message into a thrown Error).  A scene handle must be used from one
 * thread at a time.  There is no CPU fallback: every entry point that renders
 * fails if no CUDA device is present.
 */
#ifndef JSRT_H
#define JSRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JSRT_FORMAT_JSON 0      /* JSON.stringify(new Serializer(test).plain())   tests/test_to_json.js:36 */
#define JSRT_FORMAT_MSGPACK 1   /* msgpack.encode(plain)                          tests/test_to_json.js:38 */

/* jsrt_render flags */
#define JSRT_FLAG_NO_JITTER 1   /* SimpleRenderer sampling: pixel corner, no jitter (src/renderers.js:21-25) */
#define JSRT_FLAG_COUNT_WORK 2  /* instrumented kernels: count BVH nodes visited / primitives tested / SDF evaluations
                                   per ray class (slower; defines the roofline's algorithmic bytes, never timed) */

#define JSRT_FLAG_AOV 4         /* also accumulate the GL path's auxiliary buffers for these passes: first-hit normal and
                                   distance sums and the running per-pixel variance (gl/src/WebGLRendererAdapter.js:352-356,
                                   376-379); read them with jsrt_read_aov.  Radiance then goes through a per-sample buffer
                                   (slower); the image is the same up to FP32 summation order */

typedef struct jsrt_scene jsrt_scene;

/* Number of CUDA devices visible to the process (0 if none / no driver). */
int jsrt_device_count(void);

/* Replaces: worker.js:23-26 (each worker rebuilding the scene from test.mjs) and the
 * worker pool of src/raytrace_launcher.js:65-101.
 * Parses the serializer blob `{renderer:{world,camera,maxRecursionDepth
 * [,samplesPerPixel]},width,height}` (src/serializer.js:4-63), flattens it
 * (BVH nodes in reference visit order, triangle constants, material / light
 * tables, SDF bytecode) and uploads it to every device in `devices[0..ndev)`
 * (NULL / 0 = device 0).  With ndev > 1 the scene is replicated, every
 * jsrt_render call deals its passes to the devices in contiguous blocks, and
 * jsrt_resolve_rgba8 / jsrt_read_accum sum the per-device accumulation buffers
 * on devices[0] by reading the others over NVLink peer access inside the
 * resolve kernel (the devices must be peers).  Returns NULL on failure. */
jsrt_scene* jsrt_scene_create(const uint8_t* blob, size_t len, int format, const int* devices, int ndev);

/* Parse + flatten only (no CUDA calls): lets hosts without a GPU validate a
 * blob.  Rendering calls on such a handle fail. */
jsrt_scene* jsrt_scene_create_host(const uint8_t* blob, size_t len, int format);

void jsrt_scene_destroy(jsrt_scene*);

/* Re-uploads the flattened scene arrays host->device (what a second scene of
 * the same size would cost); used by the end-to-end benchmark leg. */
int jsrt_scene_upload(jsrt_scene*);

/* Use an existing CUDA stream (cudaStream_t passed as void*) for all work of
 * this scene; NULL restores the scene's own stream. */
int jsrt_scene_set_stream(jsrt_scene*, void* cuda_stream);

/* Replaces: IncrementalMultisamplingRenderer.render's pass / column / row loops
 * (src/renderers.js:87-98) and SimpleRenderer.render (:21-26).
 * Renders passes [first_pass, first_pass + n_passes) of every pixel whose
 * column px satisfies px >= x_offset && (px - x_offset) % x_delt == 0 (the
 * web-worker striping, src/renderers.js:21,88) and adds the samples into the
 * HBM-resident accumulation buffer.  `seed` keys the counter-based RNG that
 * stands in for Math.random().  Asynchronous on the scene's stream. */
int jsrt_render(jsrt_scene*, int first_pass, int n_passes, uint64_t seed, int x_offset, int x_delt, int flags);

/* Zeroes the accumulation buffer and the pass counter (`buffer` in src/renderers.js:81-86). */
int jsrt_reset_accum(jsrt_scene*);

/* Blocks until all queued work of this scene has finished. */
int jsrt_synchronize(jsrt_scene*);

/* Replaces: PixelBuffer.setColor on `buffer[px][py].times(1/(iter+1))`
 * (src/renderers.js:98, src/pixelbuffer.js:39-49): out[(y*W+x)*4+c] =
 * round(255*clamp(sum/passes, 0, 1)), alpha 255; pixels never rendered
 * (column striping) are written as 0,0,0,0 like a fresh ImageData.
 * `out` holds W*H*4 bytes.  Synchronous (device->host copy). */
int jsrt_resolve_rgba8(jsrt_scene*, uint8_t* out);

/* Reads the raw accumulation buffer: out = W*H*4 floats (sum r, g, b, sample
 * count per pixel); *passes = passes accumulated since the last reset. */
int jsrt_read_accum(jsrt_scene*, float* out, int* passes);

/* Replaces: the GL renderer's auxiliary render targets (gl/src/WebGLRendererAdapter.js:352-356 `outNormalSum`,
 * `outSampleVariance`; :376-379 first-hit distance / normal), filled by passes rendered with JSRT_FLAG_AOV.
 * normal_depth = W*H*4 floats: sum over samples of the first hit's world normal (xyz) and of its distance from the
 * ray origin (w).  variance = W*H*4 floats: xyz = the GL shader's running variance sums (divide by the sample count
 * like its display pass does), w = number of samples whose camera ray hit something.  Zero if no AOV pass was rendered. */
int jsrt_read_aov(jsrt_scene*, float* normal_depth, float* variance);

/* Replaces: the GL path's display pass with its variance-guided denoiser (gl/src/WebGLRendererAdapter.js:183-246: `smartDeNoise`
 * and `main` of the passthrough shader; defaults there: sigma 1, kSigma 2, threshold 5, colorLogScale 0, :15-22).  Filters
 * the running mean with a circular Gaussian window of radius round(kSigma * sigma) whose taps are weighted down by their
 * own per-pixel standard deviation (the variance buffer of JSRT_FLAG_AOV passes: render with that flag first).  Writes the
 * filtered mean as W*H*4 floats (rgb, mean sample weight) to out_rgba and / or as 8-bit RGBA (clamped, rounded, alpha 255) to
 * out_rgba8; either may be NULL.  Scenes created on several devices: AOV passes run on devices[0] only, and so does this. */
int jsrt_denoise(jsrt_scene*, float sigma, float k_sigma, float threshold, float color_log_scale, float* out_rgba, uint8_t* out_rgba8);

/* One process per GPU: every process exports its accumulation buffer as a 64-byte CUDA
 * IPC handle (jsrt_accum_export), the handles travel over the host's own channel, and the process that owns the image
 * maps the others' buffers with jsrt_accum_attach(handles = n x 64 bytes); from then on its jsrt_resolve_rgba8 /
 * jsrt_read_accum sum them like the helper devices of an ndev > 1 scene.  The caller orders the processes (a barrier
 * after every process's jsrt_synchronize, before the owner resolves).  n = 0 detaches. */
int jsrt_accum_export(jsrt_scene*, uint8_t* handle /* 64 bytes */);
int jsrt_accum_attach(jsrt_scene*, const uint8_t* handles, int n);

/* Device pointer of the W*H float4 accumulation buffer (for NCCL reductions
 * across GPUs issued by the host process). */
void* jsrt_accum_device_ptr(jsrt_scene*);
/* Adds `passes` to the pass counter after the host reduced other GPUs' buffers into this one. */
int jsrt_add_passes(jsrt_scene*, int passes);

/* Parity probe: un-jittered pinhole/DOF-less primary rays through every pixel
 * (SimpleRenderer sampling); prim_id[y*W+x] = index of the hit Primitive in a
 * depth-first walk of world.objects descending into Aggregate.objects (-1 =
 * miss), t = hit distance in ray-parameter units. */
int jsrt_primary_hits(jsrt_scene*, int32_t* prim_id, float* t);

typedef struct jsrt_info {
    int width, height, samples_per_pixel, max_depth;
    int jitter;                 /* renderer class samples with jitter */
    int n_top, n_prims, n_ext_prims, n_nodes, n_tris, n_materials, n_lights, n_sdfs, n_sdf_instrs;
    int light_samples;          /* shadow rays per shaded hit */
    int fanout;                 /* 1 or 2 children per shaded hit */
    int max_bvh_depth;
    int batch_samples;          /* camera samples per wavefront batch */
    uint64_t scene_bytes;       /* flattened scene bytes resident on the device */
    uint64_t queue_bytes;       /* wavefront queue bytes */
} jsrt_info;
int jsrt_scene_info(jsrt_scene*, jsrt_info*);

/* Diagnostic (works on host-only handles): the padded world-space box the device tests in front of each
 * BVHAggregate's ray transform (Aggregate.intersect, src/aggregates.js:14-18,43-46), in world.objects order.
 * out = 8 floats per aggregate (centre xyz, 0, half-size xyz, 0), at most `cap` aggregates are written;
 * returns the number of BVHAggregates (-1 on error). */
int jsrt_bvh_world_boxes(jsrt_scene*, float* out, int cap);

typedef struct jsrt_stats {
    uint64_t rays_primary, rays_secondary, rays_shadow;   /* World.cast calls, src/world.js:28-30 */
    uint64_t shaded_hits;
    uint64_t launches;                                    /* kernel launches issued */
    uint64_t camera_samples;
    double ms_generate, ms_extend, ms_shade, ms_shadow;   /* device time per kernel class; filled only when profiling is on */
    uint64_t n_generate, n_extend, n_shade, n_shadow;     /* launches per kernel class */
    /* JSRT_FLAG_COUNT_WORK renders only; index = ray class (0 primary, 1 secondary, 2 shadow) */
    uint64_t bvh_nodes[3];      /* BVHAggregateNode.intersect calls = AABB slab tests, src/aggregates.js:207-209 */
    uint64_t bvh_prims[3];      /* leaf object intersect calls, src/aggregates.js:211-212 */
    uint64_t top_prims[3];      /* top-level / list primitive intersect calls, src/world.js:9-10 */
    uint64_t sdf_evals[3];      /* root_sdf.distance() calls while marching, src/sdf.js:24 */
    /* profiling only: ms_extend / ms_shadow split by kernel (prims_kernel, bvh_kernel, sdf_kernel) */
    double ms_extend_prims, ms_extend_bvh, ms_extend_sdf, ms_shadow_prims, ms_shadow_bvh, ms_shadow_sdf;
} jsrt_stats;
/* Counters accumulated since the last jsrt_stats_reset; synchronises the scene's stream. */
int jsrt_stats_get(jsrt_scene*, jsrt_stats*);
int jsrt_stats_reset(jsrt_scene*);
/* Per-kernel CUDA-event timing: records an event pair around every launch on the
 * launching stream (no host sync until jsrt_stats_get). */
int jsrt_set_profiling(jsrt_scene*, int on);

const char* jsrt_last_error(void);

/* Measurement helper for the traversal roofline (the config scenes are L2-resident, so the
 * memory-side ceiling that binds the trace kernels is L2 bandwidth, "which the builder must measure on the box"):
 * reads a `bytes`-sized device buffer `iters` times with 128-bit loads from every SM (after one untimed pass)
 * and stores the achieved read bandwidth in GB/s.  bytes <= ~64 MB stays L2-resident; a multi-GB buffer
 * measures HBM instead. */
int jsrt_measure_read_bandwidth(int device, size_t bytes, int iters, double* gb_per_s);

/* ---- scene-build helper (host only) --------------------
 * BVHAggregateNode.build / split_objects (src/aggregates.js:65-185) over n
 * object boxes given as the reference stores them (centre, half_size, min, max;
 * 3 floats each).  Same topology as the reference's JS build. */
typedef struct jsrt_bvh jsrt_bvh;
typedef struct jsrt_bvh_node {
    int32_t depth, is_leaf, obj_first, obj_count, lesser, greater;
    float center[4], half_size[4], min[4], max[4];
} jsrt_bvh_node;
jsrt_bvh* jsrt_bvh_build(int n, const float* center, const float* half_size, const float* bmin, const float* bmax,
                         double max_depth, int min_node_size);
int jsrt_bvh_node_count(const jsrt_bvh*);
int jsrt_bvh_leaf_object_count(const jsrt_bvh*);
int jsrt_bvh_copy(const jsrt_bvh*, jsrt_bvh_node* nodes, int32_t* leaf_objects);
void jsrt_bvh_free(jsrt_bvh*);

/* The geometry half of parseObjFile (src/objloader.js:149-238) on OBJ text: positions (x y z w), texcoords (u v w),
 * normals (x y z 0) as f32, fan-triangulated faces as 9 ints per triangle ((v, vt, vn) x 3, 0-based, -1 = absent),
 * the `usemtl` index of every triangle (-1 = none, names in order of first appearance) and the `mtllib` arguments.
 * Same tokenisation, Number.parseFloat and index-regex semantics as the reference (csrc/obj_parse.cpp).
 * jsrt_obj_parse always returns a handle; jsrt_obj_error is NULL on success, else the reference's parse-error text.
 * counts = {positions, texcoords, normals, triangles, material names, mtllibs}. */
typedef struct jsrt_obj jsrt_obj;
jsrt_obj* jsrt_obj_parse(const char* text, size_t len);
const char* jsrt_obj_error(const jsrt_obj*);
void jsrt_obj_counts(const jsrt_obj*, int32_t counts[6]);
void jsrt_obj_copy(const jsrt_obj*, float* positions, float* texcoords, float* normals, int32_t* faces, int32_t* face_material);
const char* jsrt_obj_material_name(const jsrt_obj*, int i);
const char* jsrt_obj_mtllib(const jsrt_obj*, int i);
void jsrt_obj_free(jsrt_obj*);

#ifdef __cplusplus
}
#endif
#endif /* JSRT_H */
