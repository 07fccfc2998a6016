"""ctypes binding of the C-ABI library (include/jsrt.h).  This is the Python
stand-in for the N-API addon a Node host would load: same
entry points, plain pointers and sizes.

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a) as
`jsraytracer_b200/libjsrt.so`.  There is no fallback: a missing library or a
missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JSRT_LIB") or os.path.join(_HERE, "libjsrt.so")   # JSRT_LIB: A/B-test another build of the same ABI

FORMAT_JSON = 0
FORMAT_MSGPACK = 1
FLAG_NO_JITTER = 1
FLAG_COUNT_WORK = 2
FLAG_AOV = 4

_LIB = None


class JsrtError(RuntimeError):
    pass


class Info(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "width", "height", "samples_per_pixel", "max_depth", "jitter", "n_top", "n_prims", "n_ext_prims", "n_nodes",
        "n_tris", "n_materials", "n_lights", "n_sdfs", "n_sdf_instrs", "light_samples", "fanout", "max_bvh_depth",
        "batch_samples")] + [("scene_bytes", C.c_uint64), ("queue_bytes", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays_primary", "rays_secondary", "rays_shadow", "shaded_hits", "launches",
                                          "camera_samples")] + \
               [(n, C.c_double) for n in ("ms_generate", "ms_extend", "ms_shade", "ms_shadow")] + \
               [(n, C.c_uint64) for n in ("n_generate", "n_extend", "n_shade", "n_shadow")] + \
               [(n, C.c_uint64 * 3) for n in ("bvh_nodes", "bvh_prims", "top_prims", "sdf_evals")] + \
               [(n, C.c_double) for n in ("ms_extend_prims", "ms_extend_bvh", "ms_extend_sdf",
                                          "ms_shadow_prims", "ms_shadow_bvh", "ms_shadow_sdf")]

    def as_dict(self):
        d = {n: (list(getattr(self, n)) if hasattr(getattr(self, n), "__len__") else getattr(self, n))
             for n, _ in self._fields_}
        d["rays"] = d["rays_primary"] + d["rays_secondary"] + d["rays_shadow"]
        return d


class BvhNode(C.Structure):
    _fields_ = [("depth", C.c_int32), ("is_leaf", C.c_int32), ("obj_first", C.c_int32), ("obj_count", C.c_int32),
                ("lesser", C.c_int32), ("greater", C.c_int32),
                ("center", C.c_float * 4), ("half_size", C.c_float * 4), ("min", C.c_float * 4), ("max", C.c_float * 4)]


EXPORTS = [
    "jsrt_device_count", "jsrt_scene_create", "jsrt_scene_create_host", "jsrt_scene_destroy", "jsrt_scene_upload",
    "jsrt_scene_set_stream", "jsrt_render", "jsrt_reset_accum", "jsrt_synchronize", "jsrt_resolve_rgba8",
    "jsrt_read_accum", "jsrt_read_aov", "jsrt_accum_device_ptr", "jsrt_add_passes", "jsrt_primary_hits", "jsrt_scene_info",
    "jsrt_stats_get", "jsrt_stats_reset", "jsrt_set_profiling", "jsrt_last_error", "jsrt_bvh_build",
    "jsrt_bvh_node_count", "jsrt_bvh_leaf_object_count", "jsrt_bvh_copy", "jsrt_bvh_free", "jsrt_measure_read_bandwidth",
    "jsrt_bvh_world_boxes", "jsrt_obj_parse", "jsrt_obj_error", "jsrt_obj_counts", "jsrt_obj_copy", "jsrt_obj_material_name",
    "jsrt_obj_mtllib", "jsrt_obj_free", "jsrt_accum_export", "jsrt_accum_attach", "jsrt_denoise",
]


def load():
    """Loads libjsrt.so; raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise JsrtError("libjsrt.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_size_t
    L.jsrt_device_count.restype = i32
    L.jsrt_scene_create.restype = vp
    L.jsrt_scene_create.argtypes = [vp, sz, i32, vp, i32]
    L.jsrt_scene_create_host.restype = vp
    L.jsrt_scene_create_host.argtypes = [vp, sz, i32]
    L.jsrt_scene_destroy.argtypes = [vp]
    L.jsrt_scene_destroy.restype = None
    L.jsrt_scene_upload.argtypes = [vp]
    L.jsrt_scene_set_stream.argtypes = [vp, vp]
    L.jsrt_render.argtypes = [vp, i32, i32, u64, i32, i32, i32]
    L.jsrt_reset_accum.argtypes = [vp]
    L.jsrt_synchronize.argtypes = [vp]
    L.jsrt_resolve_rgba8.argtypes = [vp, vp]
    L.jsrt_read_accum.argtypes = [vp, vp, vp]
    L.jsrt_read_aov.argtypes = [vp, vp, vp]
    L.jsrt_accum_device_ptr.restype = vp
    L.jsrt_accum_device_ptr.argtypes = [vp]
    L.jsrt_add_passes.argtypes = [vp, i32]
    L.jsrt_denoise.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp]
    L.jsrt_accum_export.argtypes = [vp, vp]
    L.jsrt_accum_attach.argtypes = [vp, vp, i32]
    L.jsrt_primary_hits.argtypes = [vp, vp, vp]
    L.jsrt_scene_info.argtypes = [vp, vp]
    L.jsrt_stats_get.argtypes = [vp, vp]
    L.jsrt_stats_reset.argtypes = [vp]
    L.jsrt_set_profiling.argtypes = [vp, i32]
    L.jsrt_last_error.restype = C.c_char_p
    L.jsrt_bvh_build.restype = vp
    L.jsrt_bvh_build.argtypes = [i32, vp, vp, vp, vp, C.c_double, i32]
    L.jsrt_bvh_node_count.argtypes = [vp]
    L.jsrt_bvh_leaf_object_count.argtypes = [vp]
    L.jsrt_bvh_copy.argtypes = [vp, vp, vp]
    L.jsrt_bvh_free.argtypes = [vp]
    L.jsrt_bvh_free.restype = None
    L.jsrt_measure_read_bandwidth.argtypes = [i32, sz, i32, vp]
    L.jsrt_obj_parse.restype = vp
    L.jsrt_obj_parse.argtypes = [C.c_char_p, sz]
    L.jsrt_obj_error.restype = C.c_char_p
    L.jsrt_obj_error.argtypes = [vp]
    L.jsrt_obj_counts.restype = None
    L.jsrt_obj_counts.argtypes = [vp, vp]
    L.jsrt_obj_copy.restype = None
    L.jsrt_obj_copy.argtypes = [vp, vp, vp, vp, vp, vp]
    L.jsrt_obj_material_name.restype = C.c_char_p
    L.jsrt_obj_material_name.argtypes = [vp, i32]
    L.jsrt_obj_mtllib.restype = C.c_char_p
    L.jsrt_obj_mtllib.argtypes = [vp, i32]
    L.jsrt_obj_free.restype = None
    L.jsrt_obj_free.argtypes = [vp]
    L.jsrt_bvh_world_boxes.restype = i32
    L.jsrt_bvh_world_boxes.argtypes = [vp, vp, i32]
    _LIB = L
    return L


def last_error():
    return load().jsrt_last_error().decode("utf8", "replace")


def measure_read_bandwidth(device=0, nbytes=32 << 20, iters=200):
    """GB/s of 128-bit reads over an `nbytes` device buffer (L2-resident when small): the measured L2 ceiling."""
    out = C.c_double()
    if load().jsrt_measure_read_bandwidth(device, nbytes, iters, C.byref(out)):
        raise JsrtError(last_error())
    return out.value


def device_count():
    return load().jsrt_device_count()


class Scene:
    """Owns one `jsrt_scene*`."""

    def __init__(self, blob: bytes, fmt: int = FORMAT_MSGPACK, device=0):
        """device=None: parse + flatten only (no CUDA); rendering then fails.  A list / tuple of device indices
        replicates the scene on all of them (jsrt_scene_create with ndev > 1)."""
        self._L = load()
        if isinstance(blob, str):
            blob = blob.encode("utf8")
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        if device is None:
            self._h = self._L.jsrt_scene_create_host(buf, len(blob), fmt)
        else:
            devs = list(device) if isinstance(device, (list, tuple)) else [device]
            dev = (C.c_int * len(devs))(*devs)
            self._h = self._L.jsrt_scene_create(buf, len(blob), fmt, dev, len(devs))
        if not self._h:
            raise JsrtError(last_error())
        self.info = self.get_info()
        self.size = (self.info["width"], self.info["height"])

    def close(self):
        if getattr(self, "_h", None):
            self._L.jsrt_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise JsrtError(last_error())

    def get_info(self):
        i = Info()
        self._ck(self._L.jsrt_scene_info(self._h, C.byref(i)))
        return i.as_dict()

    def bvh_world_boxes(self):
        """(n, 2, 3) array: padded world-space (centre, half size) of every BVHAggregate (works on host-only handles)."""
        import numpy as np
        n = self._L.jsrt_bvh_world_boxes(self._h, None, 0)
        if n < 0:
            raise JsrtError(last_error())
        out = np.zeros((max(n, 1), 8), dtype=np.float32)
        self._L.jsrt_bvh_world_boxes(self._h, out.ctypes.data, n)
        return out[:n].reshape(n, 2, 4)[:, :, :3].copy()

    def upload(self):
        self._ck(self._L.jsrt_scene_upload(self._h))

    def set_stream(self, cuda_stream: int | None):
        self._ck(self._L.jsrt_scene_set_stream(self._h, cuda_stream))

    def render(self, first_pass, n_passes, seed=1, x_offset=0, x_delt=1, flags=0):
        self._ck(self._L.jsrt_render(self._h, first_pass, n_passes, seed, x_offset, x_delt, flags))

    def reset_accum(self):
        self._ck(self._L.jsrt_reset_accum(self._h))

    def synchronize(self):
        self._ck(self._L.jsrt_synchronize(self._h))

    def resolve_rgba8(self, out=None):
        W, H = self.size
        if out is None:
            out = np.empty(W * H * 4, dtype=np.uint8)
        assert out.dtype == np.uint8 and out.size == W * H * 4 and out.flags["C_CONTIGUOUS"]
        self._ck(self._L.jsrt_resolve_rgba8(self._h, out.ctypes.data))
        return out.reshape(H, W, 4)

    def read_accum(self):
        W, H = self.size
        out = np.empty((H, W, 4), dtype=np.float32)
        passes = C.c_int()
        self._ck(self._L.jsrt_read_accum(self._h, out.ctypes.data, C.byref(passes)))
        return out, passes.value

    def read_aov(self):
        """(normal_depth, variance): the GL path's auxiliary buffers of the passes rendered with FLAG_AOV (include/jsrt.h)."""
        W, H = self.size
        nd = np.empty((H, W, 4), dtype=np.float32)
        var = np.empty((H, W, 4), dtype=np.float32)
        self._ck(self._L.jsrt_read_aov(self._h, nd.ctypes.data, var.ctypes.data))
        return nd, var

    def denoise(self, sigma=1.0, k_sigma=2.0, threshold=5.0, color_log_scale=0.0, rgba8=False):
        """The GL path's variance-guided denoiser over the passes rendered with FLAG_AOV (include/jsrt.h: jsrt_denoise);
        returns the filtered mean (H, W, 4) f32, and the 8-bit image too if `rgba8`."""
        W, H = self.size
        out = np.empty((H, W, 4), dtype=np.float32)
        img = np.empty((H, W, 4), dtype=np.uint8) if rgba8 else None
        self._ck(self._L.jsrt_denoise(self._h, sigma, k_sigma, threshold, color_log_scale, out.ctypes.data, img.ctypes.data if rgba8 else None))
        return (out, img) if rgba8 else out

    def accum_device_ptr(self):
        return self._L.jsrt_accum_device_ptr(self._h)

    def accum_export(self) -> bytes:
        """64-byte CUDA IPC handle of the accumulation buffer (one process per GPU: jsrt_accum_export)."""
        h = (C.c_ubyte * 64)()
        self._ck(self._L.jsrt_accum_export(self._h, h))
        return bytes(h)

    def accum_attach(self, handles):
        """Map other processes' accumulation buffers (list of 64-byte handles) as peers of this scene."""
        blob = b"".join(handles)
        assert len(blob) == 64 * len(handles)
        buf = (C.c_ubyte * max(1, len(blob))).from_buffer_copy(blob or b"\0")
        self._ck(self._L.jsrt_accum_attach(self._h, buf, len(handles)))

    def add_passes(self, n):
        self._ck(self._L.jsrt_add_passes(self._h, n))

    def primary_hits(self):
        W, H = self.size
        ids = np.empty((H, W), dtype=np.int32)
        t = np.empty((H, W), dtype=np.float32)
        self._ck(self._L.jsrt_primary_hits(self._h, ids.ctypes.data, t.ctypes.data))
        return ids, t

    def stats(self):
        s = Stats()
        self._ck(self._L.jsrt_stats_get(self._h, C.byref(s)))
        return s.as_dict()

    def stats_reset(self):
        self._ck(self._L.jsrt_stats_reset(self._h))

    def set_profiling(self, on):
        self._ck(self._L.jsrt_set_profiling(self._h, int(bool(on))))
