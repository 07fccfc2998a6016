// C ABI of libjsrt (include/jsrt.h).  Thin: argument checks, exception -> error
// code + thread-local message, handle ownership.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <chrono>
#include <cstdio>
#include "../../include/jsrt.h"
#include "host_scene.h"
#include "render.h"

using namespace jsrt;

// One scene handle = the flattened scene + one Renderer per CUDA device it was created for.  `renderer` (device
// devices[0]) owns the image: the other devices render their share of every call's passes into their own accumulation
// buffers, which the first device's resolve / read-back kernels sum over NVLink peer access (render.cu: PeerAccum) —
// the in-process replacement of the reference's worker pool + compositing (src/raytrace_launcher.js:65-101).
struct jsrt_scene {
    HostScene host;
    std::unique_ptr<Renderer> renderer;                 // null for host-only handles
    std::vector<std::unique_ptr<Renderer>> helpers;     // devices[1..]
    std::vector<Renderer*> all() { std::vector<Renderer*> v; if (renderer) v.push_back(renderer.get()); for (auto& h : helpers) v.push_back(h.get()); return v; }
    // everything queued on the helpers happens before what the owner queues next (resolve, read-back)
    void joinHelpers() { for (auto& h : helpers) orderAfter(*renderer, *h); }
};

static thread_local std::string g_error;

static int failWith(const std::exception& e) { g_error = e.what(); return 1; }
static int needDevice(jsrt_scene* s) {
    if (!s) { g_error = "jsrt: null scene handle"; return 1; }
    if (!s->renderer) { g_error = "jsrt: scene was created without a CUDA device (jsrt_scene_create_host); there is no CPU fallback"; return 1; }
    return 0;
}
#define JSRT_TRY(body) try { body; return 0; } catch (const std::exception& e) { return failWith(e); }

static size_t queueBudget() {
    if (const char* e = getenv("JSRT_QUEUE_BYTES")) { const double v = atof(e); if (v >= 1e6) return (size_t)v; }
    return (size_t)64 << 30;      // of the 180 GB of HBM3e; see render.cu for why big waves pay
}

extern "C" {

const char* jsrt_last_error(void) { return g_error.c_str(); }

int jsrt_device_count(void) { return deviceCount(); }

jsrt_scene* jsrt_scene_create_host(const uint8_t* blob, size_t len, int format) {
    try {
        std::unique_ptr<jsrt_scene> s(new jsrt_scene);
        const bool timing = getenv("JSRT_HOST_TIMING") != nullptr;     // diagnostics: where scene creation spends its time
        const auto t0 = std::chrono::steady_clock::now();
        WireDoc doc(blob, len, format);
        const auto t1 = std::chrono::steady_clock::now();
        flattenScene(doc, s->host);
        if (timing) {
            const auto t2 = std::chrono::steady_clock::now();
            fprintf(stderr, "jsrt: scene blob %.1f MB: parse %.3f s, flatten %.3f s\n", len / 1e6,
                    std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
        }
        return s.release();
    } catch (const std::exception& e) { failWith(e); return nullptr; }
}

jsrt_scene* jsrt_scene_create(const uint8_t* blob, size_t len, int format, const int* devices, int ndev) {
    try {
        if (ndev < 0 || ndev > 16) fail("jsrt: ndev must be in 0..16");
        const int have = deviceCount();
        if (have <= 0) fail("jsrt: no CUDA device available; this library has no CPU fallback");
        std::vector<int> devs;
        for (int i = 0; devices && i < ndev; ++i) devs.push_back(devices[i]);
        if (devs.empty()) devs.push_back(0);
        for (size_t i = 0; i < devs.size(); ++i) {
            if (devs[i] < 0 || devs[i] >= have) fail("jsrt: CUDA device index out of range");
            for (size_t k = 0; k < i; ++k) if (devs[k] == devs[i]) fail("jsrt: the same CUDA device listed twice");
        }
        std::unique_ptr<jsrt_scene> s(jsrt_scene_create_host(blob, len, format));
        if (!s) return nullptr;
        s->renderer.reset(new Renderer(s->host, devs[0], queueBudget()));
        std::vector<const void*> peer_ptrs;
        for (size_t i = 1; i < devs.size(); ++i) {
            enablePeerAccess(devs[0], devs[i]);
            s->helpers.emplace_back(new Renderer(s->host, devs[i], queueBudget()));
            peer_ptrs.push_back(s->helpers.back()->accumPtr());
        }
        if (!peer_ptrs.empty()) s->renderer->setPeers(peer_ptrs.data(), (int)peer_ptrs.size());
        return s.release();
    } catch (const std::exception& e) { failWith(e); return nullptr; }
}

void jsrt_scene_destroy(jsrt_scene* s) { delete s; }

int jsrt_scene_upload(jsrt_scene* s) { if (needDevice(s)) return 1; JSRT_TRY(for (Renderer* r : s->all()) r->upload()) }
int jsrt_scene_set_stream(jsrt_scene* s, void* st) { if (needDevice(s)) return 1; JSRT_TRY(s->renderer->setStream(st)) }

int jsrt_render(jsrt_scene* s, int first_pass, int n_passes, uint64_t seed, int x_offset, int x_delt, int flags) {
    if (needDevice(s)) return 1;
    if (n_passes < 0 || first_pass < 0) { g_error = "jsrt: negative pass range"; return 1; }
    try {
        // The passes of one call are dealt to the scene's devices in contiguous blocks (big waves per GPU); the RNG is keyed
        // by the absolute pass index, so the image does not depend on the split (up to FP32 summation order).  The AOV
        // buffers live on the first device only, so AOV passes are not split.
        std::vector<Renderer*> devs = s->all();
        const int nd = (flags & JSRT_FLAG_AOV) ? 1 : (int)devs.size();
        const int chunk = (n_passes + nd - 1) / nd;
        int own = 0;
        for (int g = 0; g < nd; ++g) {
            const int first = first_pass + g * chunk, n = std::min(chunk, first_pass + n_passes - first);
            if (n <= 0) break;
            devs[g]->render(first, n, seed, x_offset, x_delt, flags);
            if (g == 0) own = n;
        }
        if (nd > 1 && n_passes > own) s->renderer->addPasses((x_offset == 0 && x_delt <= 1) ? n_passes - own : 0);
        return 0;
    } catch (const std::exception& e) { return failWith(e); }
}
int jsrt_reset_accum(jsrt_scene* s) { if (needDevice(s)) return 1; JSRT_TRY(for (Renderer* r : s->all()) r->resetAccum()) }
int jsrt_synchronize(jsrt_scene* s) { if (needDevice(s)) return 1; JSRT_TRY(for (Renderer* r : s->all()) r->synchronize()) }
int jsrt_resolve_rgba8(jsrt_scene* s, uint8_t* out) {
    if (needDevice(s)) return 1;
    if (!out) { g_error = "jsrt: null output buffer"; return 1; }
    JSRT_TRY(s->joinHelpers(); s->renderer->resolve(out))
}
int jsrt_read_accum(jsrt_scene* s, float* out, int* passes) {
    if (needDevice(s)) return 1;
    if (!out) { g_error = "jsrt: null output buffer"; return 1; }
    JSRT_TRY(s->joinHelpers(); s->renderer->readAccum(out, passes))
}
int jsrt_accum_export(jsrt_scene* s, uint8_t* handle) {
    if (needDevice(s)) return 1;
    if (!handle) { g_error = "jsrt: null output buffer"; return 1; }
    JSRT_TRY(s->renderer->exportAccum(handle))
}
int jsrt_accum_attach(jsrt_scene* s, const uint8_t* handles, int n) {
    if (needDevice(s)) return 1;
    if (n < 0 || (n > 0 && !handles)) { g_error = "jsrt: bad peer handle list"; return 1; }
    if (!s->helpers.empty()) { g_error = "jsrt: a scene created on several devices cannot also attach remote buffers"; return 1; }
    JSRT_TRY(s->renderer->attachAccum(handles, n))
}
int jsrt_read_aov(jsrt_scene* s, float* normal_depth, float* variance) {
    if (needDevice(s)) return 1;
    if (!normal_depth || !variance) { g_error = "jsrt: null output buffer"; return 1; }
    JSRT_TRY(s->renderer->readAov(normal_depth, variance))
}
int jsrt_denoise(jsrt_scene* s, float sigma, float k_sigma, float threshold, float color_log_scale, float* out_rgba, uint8_t* out_rgba8) {
    if (needDevice(s)) return 1;
    if (!out_rgba && !out_rgba8) { g_error = "jsrt: null output buffer"; return 1; }
    JSRT_TRY(s->renderer->denoise(sigma, k_sigma, threshold, color_log_scale, out_rgba, out_rgba8))
}
void* jsrt_accum_device_ptr(jsrt_scene* s) { return (s && s->renderer) ? s->renderer->accumPtr() : nullptr; }
int jsrt_add_passes(jsrt_scene* s, int n) { if (needDevice(s)) return 1; JSRT_TRY(s->renderer->addPasses(n)) }
int jsrt_primary_hits(jsrt_scene* s, int32_t* prim_id, float* t) { if (needDevice(s)) return 1; if (!prim_id || !t) { g_error = "jsrt: null output buffer"; return 1; } JSRT_TRY(s->renderer->primaryHits(prim_id, t)) }

int jsrt_scene_info(jsrt_scene* s, jsrt_info* o) {
    if (!s || !o) { g_error = "jsrt: null argument"; return 1; }
    const HostScene& h = s->host;
    memset(o, 0, sizeof *o);
    o->width = h.width; o->height = h.height; o->samples_per_pixel = h.samples_per_pixel; o->max_depth = h.max_depth; o->jitter = h.jitter;
    o->n_top = h.world_object_count; o->n_prims = (int)h.prims.size(); o->n_ext_prims = h.ext_prim_count; o->n_nodes = h.tree_node_count;
    o->n_tris = (int)h.tris.size(); o->n_materials = (int)h.materials.size(); o->n_lights = (int)h.lights.size();
    o->n_sdfs = (int)h.sdfs.size(); o->n_sdf_instrs = (int)h.sdf_code.size(); o->light_samples = h.light_samples; o->fanout = h.fanout;
    o->max_bvh_depth = h.max_bvh_depth;
    if (s->renderer) { o->batch_samples = s->renderer->batchSamples(); o->scene_bytes = s->renderer->sceneBytes(); o->queue_bytes = s->renderer->queueBytes(); }
    return 0;
}

int jsrt_bvh_world_boxes(jsrt_scene* s, float* out, int cap) {
    if (!s) { g_error = "jsrt: null scene handle"; return -1; }
    try {
        std::vector<float> wb;
        computeWorldBoxes(s->host, wb);
        const int n = (int)(wb.size() / 8);
        if (out) memcpy(out, wb.data(), sizeof(float) * 8 * (size_t)(n < cap ? n : (cap < 0 ? 0 : cap)));
        return n;
    } catch (const std::exception& e) { failWith(e); return -1; }
}

int jsrt_stats_get(jsrt_scene* s, jsrt_stats* o) {
    if (needDevice(s)) return 1;
    try {
        RenderStats r; s->renderer->getStats(r);
        for (auto& h : s->helpers) {          // rays and work of the helper devices are added; times stay those of the first device
            RenderStats q; h->getStats(q);
            r.rays_primary += q.rays_primary; r.rays_secondary += q.rays_secondary; r.rays_shadow += q.rays_shadow; r.shaded_hits += q.shaded_hits;
            r.camera_samples += q.camera_samples; r.launches += q.launches;
            for (int k = 0; k < 3; ++k) { r.nodes[k] += q.nodes[k]; r.leaf_prims[k] += q.leaf_prims[k]; r.top_prims[k] += q.top_prims[k]; r.sdf_evals[k] += q.sdf_evals[k]; }
        }
        memset(o, 0, sizeof *o);
        o->rays_primary = r.rays_primary; o->rays_secondary = r.rays_secondary; o->rays_shadow = r.rays_shadow; o->shaded_hits = r.shaded_hits;
        o->launches = r.launches; o->camera_samples = r.camera_samples;
        o->ms_generate = r.ms[0]; o->ms_extend = r.ms[1]; o->ms_shade = r.ms[2]; o->ms_shadow = r.ms[3];
        for (int k = 0; k < 3; ++k) { o->bvh_nodes[k] = r.nodes[k]; o->bvh_prims[k] = r.leaf_prims[k]; o->top_prims[k] = r.top_prims[k]; o->sdf_evals[k] = r.sdf_evals[k]; }
        o->ms_extend_prims = r.ms_part[0]; o->ms_extend_bvh = r.ms_part[1]; o->ms_extend_sdf = r.ms_part[2];
        o->ms_shadow_prims = r.ms_part[3]; o->ms_shadow_bvh = r.ms_part[4]; o->ms_shadow_sdf = r.ms_part[5];
        o->n_generate = r.kernel_launches[0]; o->n_extend = r.kernel_launches[1]; o->n_shade = r.kernel_launches[2]; o->n_shadow = r.kernel_launches[3];
        return 0;
    } catch (const std::exception& e) { return failWith(e); }
}
int jsrt_stats_reset(jsrt_scene* s) { if (needDevice(s)) return 1; JSRT_TRY(for (Renderer* r : s->all()) r->resetStats()) }
int jsrt_set_profiling(jsrt_scene* s, int on) { if (needDevice(s)) return 1; JSRT_TRY(s->renderer->setProfiling(on != 0)) }

int jsrt_measure_read_bandwidth(int device, size_t bytes, int iters, double* gb_per_s) {
    try {
        if (jsrt::deviceCount() < 1) throw std::runtime_error("jsrt: no CUDA device");
        if (!gb_per_s || bytes < 4096 || iters < 1) throw std::runtime_error("jsrt: bad arguments");
        *gb_per_s = jsrt::measureReadBandwidth(device, bytes, iters);
        return 0;
    } catch (const std::exception& e) { return failWith(e); }
}

}  // extern "C"
