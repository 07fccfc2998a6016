// N-API shim between Node and libjsrt (include/jsrt.h): the binding a maintainer of the
// reference would add so that js/cuda_renderer.js can run under Node next to the web-worker
// renderer.  Pure C N-API (node_api.h is a stable C ABI); one function per C entry point,
// errors become thrown JS Errors carrying jsrt_last_error() (the reference throws strings,
// e.g. src/aggregates.js:39).
//
// Not compiled in this repository's CI: neither node nor node_api.h exist in the build image
// (SURVEY.md Appendix A).  Build under Node with:  npx node-gyp configure build   (binding.gyp).
#include <node_api.h>

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/jsrt.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, NULL, "N-API call failed: " #call); return NULL; } } while (0)

static napi_value throw_last(napi_env env) { napi_throw_error(env, "JSRT", jsrt_last_error()); return NULL; }

// The external wraps a small box, not the scene itself: destroyScene() empties the box, so a stale handle throws instead of
// double-freeing, and the finalizer frees whatever is still alive when the JS object is collected.
typedef struct { jsrt_scene* scene; } scene_box;
static void scene_finalize(napi_env env, void* data, void* hint) {
    (void)env; (void)hint;
    scene_box* b = (scene_box*)data;
    if (b) { if (b->scene) jsrt_scene_destroy(b->scene); free(b); }
}
static scene_box* box_arg(napi_env env, napi_value v) {
    void* p = NULL;
    if (napi_get_value_external(env, v, &p) != napi_ok || !p) { napi_throw_type_error(env, NULL, "scene handle expected"); return NULL; }
    return (scene_box*)p;
}
static jsrt_scene* scene_arg(napi_env env, napi_value v) {
    scene_box* b = box_arg(env, v);
    if (!b) return NULL;
    if (!b->scene) { napi_throw_error(env, "JSRT", "scene handle used after destroyScene()"); return NULL; }
    return b->scene;
}
static int32_t int_arg(napi_env env, napi_value v) { int32_t x = 0; napi_get_value_int32(env, v, &x); return x; }

// createScene(blob: Buffer, format: 0|1, device: number) -> external
static napi_value CreateScene(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value a[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    void* data = NULL; size_t len = 0;
    NAPI_OK(napi_get_buffer_info(env, a[0], &data, &len));
    // device: a number, or an Int32Array of device indices (the scene is replicated and every render call's passes are
    // dealt to them: jsrt_scene_create with ndev > 1)
    int devs[16]; int ndev = 1; devs[0] = 0;
    if (argc > 2) {
        napi_typedarray_type ty; size_t n = 0; void* d = NULL; napi_value ab; size_t off = 0;
        if (napi_get_typedarray_info(env, a[2], &ty, &n, &d, &ab, &off) == napi_ok) {
            if (ty != napi_int32_array || n < 1 || n > 16) { napi_throw_range_error(env, NULL, "devices must be an Int32Array of 1..16 indices"); return NULL; }
            ndev = (int)n; memcpy(devs, d, n * sizeof(int));
        } else devs[0] = int_arg(env, a[2]);
    }
    jsrt_scene* s = jsrt_scene_create((const uint8_t*)data, len, int_arg(env, a[1]), devs, ndev);
    if (!s) return throw_last(env);
    scene_box* b = (scene_box*)malloc(sizeof *b);
    if (!b) { jsrt_scene_destroy(s); napi_throw_error(env, NULL, "out of memory"); return NULL; }
    b->scene = s;
    napi_value ext;
    if (napi_create_external(env, b, scene_finalize, NULL, &ext) != napi_ok) { scene_finalize(env, b, NULL); napi_throw_error(env, NULL, "napi_create_external failed"); return NULL; }
    return ext;
}
static napi_value DestroyScene(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    scene_box* b = box_arg(env, a[0]);
    if (b && b->scene) { jsrt_scene_destroy(b->scene); b->scene = NULL; }       // idempotent
    return NULL;
}
// render(scene, firstPass, nPasses, seed, xOffset, xDelt, flags)
static napi_value Render(napi_env env, napi_callback_info info) {
    size_t argc = 7; napi_value a[7]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    jsrt_scene* s = scene_arg(env, a[0]); if (!s) return NULL;
    int64_t seed = 1; napi_get_value_int64(env, a[3], &seed);
    if (jsrt_render(s, int_arg(env, a[1]), int_arg(env, a[2]), (uint64_t)seed, int_arg(env, a[4]), int_arg(env, a[5]), argc > 6 ? int_arg(env, a[6]) : 0))
        return throw_last(env);
    return NULL;
}
static napi_value ResetAccum(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    jsrt_scene* s = scene_arg(env, a[0]); if (!s) return NULL;
    if (jsrt_reset_accum(s)) return throw_last(env);
    return NULL;
}
static napi_value Synchronize(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value a[1]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    jsrt_scene* s = scene_arg(env, a[0]); if (!s) return NULL;
    if (jsrt_synchronize(s)) return throw_last(env);
    return NULL;
}
// resolveRGBA8(scene, out: Uint8ClampedArray | Buffer of W*H*4)
static napi_value ResolveRGBA8(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    jsrt_scene* s = scene_arg(env, a[0]); if (!s) return NULL;
    napi_typedarray_type ty; size_t n = 0; void* data = NULL; napi_value ab; size_t off = 0;
    NAPI_OK(napi_get_typedarray_info(env, a[1], &ty, &n, &data, &ab, &off));
    jsrt_info inf; if (jsrt_scene_info(s, &inf)) return throw_last(env);
    if ((ty != napi_uint8_clamped_array && ty != napi_uint8_array) || n < (size_t)inf.width * inf.height * 4) {
        napi_throw_range_error(env, NULL, "output must be a Uint8ClampedArray of width*height*4"); return NULL;
    }
    if (jsrt_resolve_rgba8(s, (uint8_t*)data)) return throw_last(env);
    return NULL;
}
// primaryHits(scene, primId: Int32Array, t: Float32Array)  — parity probe
static napi_value PrimaryHits(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value a[3]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    jsrt_scene* s = scene_arg(env, a[0]); if (!s) return NULL;
    napi_typedarray_type ty1, ty2; size_t n1 = 0, n2 = 0; void *d1 = NULL, *d2 = NULL; napi_value ab; size_t off;
    NAPI_OK(napi_get_typedarray_info(env, a[1], &ty1, &n1, &d1, &ab, &off));
    NAPI_OK(napi_get_typedarray_info(env, a[2], &ty2, &n2, &d2, &ab, &off));
    jsrt_info inf; if (jsrt_scene_info(s, &inf)) return throw_last(env);
    const size_t npix = (size_t)inf.width * inf.height;
    if (ty1 != napi_int32_array || ty2 != napi_float32_array || n1 < npix || n2 < npix) {
        napi_throw_range_error(env, NULL, "primaryHits needs an Int32Array and a Float32Array of width*height"); return NULL;
    }
    if (jsrt_primary_hits(s, (int32_t*)d1, (float*)d2)) return throw_last(env);
    return NULL;
}
// sceneHeader(blob: Buffer, format: 0|1, out: Int32Array(5)) — width, height, samplesPerPixel, maxRecursionDepth, jitter of a
// serialised scene, read by the library's own parser without touching a GPU (jsrt_scene_create_host).  What
// CUDARenderer.fromWire needs for a scene that only exists in wire form (tests/dragon_json, tests/toledo_json).
static napi_value SceneHeader(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value a[3]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    void* data = NULL; size_t len = 0;
    NAPI_OK(napi_get_buffer_info(env, a[0], &data, &len));
    napi_typedarray_type ty; size_t n = 0; void* out = NULL; napi_value ab; size_t off = 0;
    NAPI_OK(napi_get_typedarray_info(env, a[2], &ty, &n, &out, &ab, &off));
    if (ty != napi_int32_array || n < 5) { napi_throw_range_error(env, NULL, "output must be an Int32Array of 5"); return NULL; }
    jsrt_scene* s = jsrt_scene_create_host((const uint8_t*)data, len, int_arg(env, a[1]));
    if (!s) return throw_last(env);
    jsrt_info inf; const int rc = jsrt_scene_info(s, &inf);
    jsrt_scene_destroy(s);
    if (rc) return throw_last(env);
    int32_t* o = (int32_t*)out;
    o[0] = inf.width; o[1] = inf.height; o[2] = inf.samples_per_pixel; o[3] = inf.max_depth; o[4] = inf.jitter;
    return NULL;
}
// readAccum(scene, out: Float32Array of W*H*4) -> passes accumulated so far (the f32 sums behind the 8-bit image)
static napi_value ReadAccum(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value a[2]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    jsrt_scene* s = scene_arg(env, a[0]); if (!s) return NULL;
    napi_typedarray_type ty; size_t n = 0; void* data = NULL; napi_value ab; size_t off = 0;
    NAPI_OK(napi_get_typedarray_info(env, a[1], &ty, &n, &data, &ab, &off));
    jsrt_info inf; if (jsrt_scene_info(s, &inf)) return throw_last(env);
    if (ty != napi_float32_array || n < (size_t)inf.width * inf.height * 4) {
        napi_throw_range_error(env, NULL, "output must be a Float32Array of width*height*4"); return NULL;
    }
    int passes = 0;
    if (jsrt_read_accum(s, (float*)data, &passes)) return throw_last(env);
    napi_value v; NAPI_OK(napi_create_int32(env, passes, &v)); return v;
}
// readAov(scene, normalDepth: Float32Array(W*H*4), variance: Float32Array(W*H*4)): the GL path's auxiliary buffers
// (gl/src/WebGLRendererAdapter.js:352-356,376-379) of the passes rendered with flag 4 (JSRT_FLAG_AOV)
static int float_out(napi_env env, napi_value v, size_t need, float** out) {
    napi_typedarray_type ty; size_t n = 0; void* data = NULL; napi_value ab; size_t off = 0;
    if (napi_get_typedarray_info(env, v, &ty, &n, &data, &ab, &off) != napi_ok || ty != napi_float32_array || n < need) {
        napi_throw_range_error(env, NULL, "output must be a Float32Array of width*height*4"); return 1;
    }
    *out = (float*)data; return 0;
}
static napi_value ReadAov(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value a[3]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    jsrt_scene* s = scene_arg(env, a[0]); if (!s) return NULL;
    jsrt_info inf; if (jsrt_scene_info(s, &inf)) return throw_last(env);
    const size_t need = (size_t)inf.width * inf.height * 4;
    float *nd = NULL, *var = NULL;
    if (float_out(env, a[1], need, &nd) || float_out(env, a[2], need, &var)) return NULL;
    if (jsrt_read_aov(s, nd, var)) return throw_last(env);
    return NULL;
}
// denoise(scene, sigma, kSigma, threshold, colorLogScale, out: Float32Array(W*H*4) | Uint8ClampedArray(W*H*4)): the GL path's
// display pass with its variance-guided filter (gl/src/WebGLRendererAdapter.js:183-246; its defaults: 1, 2, 5, 0)
static napi_value Denoise(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value a[6]; NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
    jsrt_scene* s = scene_arg(env, a[0]); if (!s) return NULL;
    double p[4] = {1, 2, 5, 0};
    for (int i = 0; i < 4; ++i)
        if (napi_get_value_double(env, a[1 + i], &p[i]) != napi_ok) { napi_throw_type_error(env, NULL, "sigma, kSigma, threshold, colorLogScale must be numbers"); return NULL; }
    jsrt_info inf; if (jsrt_scene_info(s, &inf)) return throw_last(env);
    napi_typedarray_type ty; size_t n = 0; void* data = NULL; napi_value ab; size_t off = 0;
    if (napi_get_typedarray_info(env, a[5], &ty, &n, &data, &ab, &off) != napi_ok || n < (size_t)inf.width * inf.height * 4 ||
        (ty != napi_float32_array && ty != napi_uint8_clamped_array && ty != napi_uint8_array)) {
        napi_throw_range_error(env, NULL, "output must be a Float32Array or Uint8ClampedArray of width*height*4"); return NULL;
    }
    const int f32 = ty == napi_float32_array;
    if (jsrt_denoise(s, (float)p[0], (float)p[1], (float)p[2], (float)p[3], f32 ? (float*)data : NULL, f32 ? NULL : (uint8_t*)data)) return throw_last(env);
    return NULL;
}
static napi_value DeviceCount(napi_env env, napi_callback_info info) {
    (void)info; napi_value v; NAPI_OK(napi_create_int32(env, jsrt_device_count(), &v)); return v;
}

static napi_value Init(napi_env env, napi_value exports) {
    const struct { const char* name; napi_callback fn; } fns[] = {
        {"createScene", CreateScene}, {"destroyScene", DestroyScene}, {"render", Render}, {"resetAccum", ResetAccum},
        {"synchronize", Synchronize}, {"resolveRGBA8", ResolveRGBA8}, {"primaryHits", PrimaryHits}, {"deviceCount", DeviceCount},
        {"sceneHeader", SceneHeader}, {"readAccum", ReadAccum}, {"readAov", ReadAov}, {"denoise", Denoise}};
    for (size_t i = 0; i < sizeof fns / sizeof fns[0]; ++i) {
        napi_value f; NAPI_OK(napi_create_function(env, fns[i].name, NAPI_AUTO_LENGTH, fns[i].fn, NULL, &f));
        NAPI_OK(napi_set_named_property(env, exports, fns[i].name, f));
    }
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
