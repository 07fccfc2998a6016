// Host interface of the wavefront renderer (render.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>

#include "host_scene.h"

namespace jsrt {

struct RenderStats {
    uint64_t rays_primary = 0, rays_secondary = 0, rays_shadow = 0, shaded_hits = 0, camera_samples = 0, launches = 0;
    double ms[4] = {0, 0, 0, 0};   // generate, extend, shade, shadow (profiling mode only)
    double ms_part[6] = {0, 0, 0, 0, 0, 0};   // extend: prims, bvh, sdf kernels | shadow: prims, bvh, sdf kernels
    uint64_t kernel_launches[4] = {0, 0, 0, 0};
    // per ray class (primary, secondary, shadow); only counted by JSRT_FLAG_COUNT_WORK renders
    uint64_t nodes[3] = {0, 0, 0}, leaf_prims[3] = {0, 0, 0}, top_prims[3] = {0, 0, 0}, sdf_evals[3] = {0, 0, 0};
};

class Renderer {
public:
    Renderer(const HostScene& hs, int device, size_t queue_budget_bytes);
    ~Renderer();
    void render(int first_pass, int n_passes, uint64_t seed, int x_offset, int x_delt, int flags);
    void flushPending();          // launches what jsrt_render calls have been holding back (submission coalescing, render.cu)
    void upload();
    void setStream(void* cuda_stream);
    void synchronize();
    void resetAccum();
    void resolve(uint8_t* out_rgba);
    void readAccum(float* out, int* passes);
    void readAov(float* normal_depth, float* variance);
    void denoise(float sigma, float k_sigma, float threshold, float color_log_scale, float* out_rgba, uint8_t* out_rgba8);
    void* accumPtr();
    void addPasses(int n);
    // other GPUs' accumulation buffers, summed into this one's by resolve() / readAccum() (read over NVLink by the kernels)
    void setPeers(const void* const* device_ptrs, int n);
    void exportAccum(void* ipc_handle_64_bytes);                  // cudaIpcGetMemHandle of the accumulation buffer
    void attachAccum(const void* ipc_handles, int n);             // maps n exported buffers of other processes as peers
    void* stream() const;
    int device() const;
    void primaryHits(int32_t* prim_id, float* t);
    void getStats(RenderStats& s);
    void resetStats();
    void setProfiling(bool on);
    int passes() const;
    int batchSamples() const;
    size_t sceneBytes() const;
    size_t queueBytes() const;

private:
    struct Impl;
    Impl* impl_;
};

int deviceCount();
// device `a` may read device `b`'s memory (cudaDeviceEnablePeerAccess); throws if the GPUs are not peers
void enablePeerAccess(int a, int b);
// work queued so far on `producer`'s stream happens before anything queued later on `consumer`'s stream (cross-device event)
void orderAfter(Renderer& consumer, Renderer& producer);
double measureReadBandwidth(int device, size_t bytes, int iters);

}  // namespace jsrt
