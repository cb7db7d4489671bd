"""refrakt_b200 — ctypes harness over the C ABI (include/refrakt_b200.h).

The product is the C++/CUDA library librefrakt_b200.so; this module only mirrors the
reference's host interface (flame_compiler, flame: load_flame / warmup / draw_to_bins,
src/flame.hpp, src/variation_table.hpp) for the tests and the bench. There is no CPU
path: GPU entry points raise RefraktError when the CUDA device or the library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librefrakt_b200.so")
DATA_DIR = os.path.join(_HERE, "data")
OVERLAY_YAML = os.path.join(DATA_DIR, "variations_b200.yaml")


class RefraktError(RuntimeError):
    pass


class FlameInfo(C.Structure):
    _fields_ = [("size", C.c_uint32 * 2), ("center", C.c_float * 2), ("scale", C.c_float), ("rotate", C.c_float),
                ("estimator_min", C.c_int32), ("estimator_radius", C.c_int32), ("estimator_curve", C.c_float),
                ("gamma", C.c_float), ("vibrancy", C.c_float), ("brightness", C.c_float),
                ("num_xforms", C.c_int32), ("has_final_xform", C.c_int32), ("param_count", C.c_int32)]


class XformInfo(C.Structure):
    _fields_ = [("affine", C.c_float * 6), ("has_post", C.c_int32), ("post", C.c_float * 6),
                ("weight", C.c_float), ("color", C.c_float), ("color_speed", C.c_float),
                ("rotation_frequency", C.c_float), ("opacity", C.c_float),
                ("num_variations", C.c_int32), ("num_params", C.c_int32)]


class KernelOptions(C.Structure):
    _fields_ = [("math_mode", C.c_int32), ("fmad", C.c_int32), ("per_lane_xform", C.c_int32), ("warp_aggregate", C.c_int32),
                ("deterministic", C.c_int32), ("count_xforms", C.c_int32), ("min_blocks", C.c_int32), ("block_width", C.c_int32), ("deal_period", C.c_int32), ("l2_hints", C.c_int32), ("staged_bins", C.c_int32), ("specialize", C.c_int32), ("pair_particles", C.c_int32)]


class HotMapInfo(C.Structure):
    _fields_ = [("tiles_x", C.c_int32), ("tiles_y", C.c_int32), ("hot_tiles", C.c_uint32), ("threshold_bucket", C.c_uint32), ("budget_bytes", C.c_uint64)]


class PostParams(C.Structure):
    _fields_ = [("estimator_radius", C.c_int32), ("estimator_min", C.c_int32), ("estimator_curve", C.c_float),
                ("gamma", C.c_float), ("brightness", C.c_float), ("vibrancy", C.c_float), ("scale_constant", C.c_float)]


class FrameRequest(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("warmup_passes", C.c_uint32), ("drawing_passes", C.c_uint32),
                ("tss_width", C.c_float), ("target_binned", C.c_uint64), ("max_draw_calls", C.c_uint32), ("scale_constant_exp", C.c_float),
                ("supersample", C.c_uint32), ("filter_radius", C.c_float)]


class FrameStats(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("binned", C.c_uint64), ("draw_calls", C.c_uint32),
                ("ms_warmup", C.c_float), ("ms_draw", C.c_float), ("ms_post", C.c_float), ("ms_readback", C.c_float)]


class MotionInfo(C.Structure):
    _fields_ = [("freq", C.c_float), ("amplitude", C.c_float), ("function", C.c_char * 16), ("target", C.c_char * 64)]


class RowSlab(C.Structure):
    _fields_ = [("y0", C.c_uint32), ("y1", C.c_uint32), ("src_y0", C.c_uint32), ("src_y1", C.c_uint32)]


class ShardedRequest(C.Structure):
    _fields_ = [("frame", FrameRequest), ("want_rgba8", C.c_uint32), ("want_image", C.c_uint32)]


class ShardedStats(C.Structure):
    _fields_ = [("iterations_global", C.c_uint64), ("binned_global", C.c_uint64), ("passes", C.c_uint64), ("draw_calls", C.c_uint32), ("p2p", C.c_uint32),
                ("y0", C.c_uint32), ("y1", C.c_uint32),
                ("ms_warmup", C.c_float), ("ms_draw", C.c_float), ("ms_reduce", C.c_float), ("ms_post", C.c_float), ("ms_readback", C.c_float)]


_vp, _cp, _sz, _i, _f = C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_float
_fpp, _ipp, _upp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint32)

# name -> (restype, argtypes); every symbol include/refrakt_b200.h declares
SIGNATURES = {
    "rfk_abi_version": (_i, []),
    "rfk_last_error": (_cp, []),
    "rfk_set_device": (_i, [_i]),
    "rfk_set_stream": (_i, [_vp]),
    "rfk_synchronize": (_i, []),
    "rfk_device_alloc": (_vp, [_sz]),
    "rfk_device_free": (_i, [_vp]),
    "rfk_device_zero": (_i, [_vp, _sz]),
    "rfk_memcpy_to_device": (_i, [_vp, _vp, _sz]),
    "rfk_memcpy_to_host": (_i, [_vp, _vp, _sz]),
    "rfk_release_buffers": (C.c_int, []),
    "rfk_kernel_launch_count": (C.c_uint64, []),
    "rfk_set_sim_parameters": (_i, [_sz, _sz, _sz, C.c_uint64]),
    "rfk_compiler_create": (_vp, [_cp]),
    "rfk_compiler_create_from_text": (_vp, [_cp]),
    "rfk_compiler_load_overlay": (_i, [_vp, _cp]),
    "rfk_compiler_destroy": (None, [_vp]),
    "rfk_compiler_is_param": (_i, [_vp, _cp]),
    "rfk_compiler_is_variation": (_i, [_vp, _cp]),
    "rfk_compiler_is_common": (_i, [_vp, _cp]),
    "rfk_compiler_variation_count": (_i, [_vp]),
    "rfk_compiler_variation_name": (_cp, [_vp, _i]),
    "rfk_compiler_get_parameters_for_variation": (_i, [_vp, _cp, _cp, _sz]),
    "rfk_compiler_param_owner": (_cp, [_vp, _cp]),
    "rfk_flame_load": (_vp, [_cp, _vp]),
    "rfk_flame_load_string": (_vp, [_cp, _vp]),
    "rfk_flame_destroy": (None, [_vp]),
    "rfk_flame_get_info": (_i, [_vp, C.POINTER(FlameInfo)]),
    "rfk_flame_set_info": (_i, [_vp, C.POINTER(FlameInfo)]),
    "rfk_flame_get_xform": (_i, [_vp, _i, C.POINTER(XformInfo)]),
    "rfk_flame_set_xform": (_i, [_vp, _i, C.POINTER(XformInfo)]),
    "rfk_flame_variation_name": (_cp, [_vp, _i, _i]),
    "rfk_flame_param_name": (_cp, [_vp, _i, _i]),
    "rfk_flame_get_variation": (_i, [_vp, _i, _cp, _fpp]),
    "rfk_flame_set_variation": (_i, [_vp, _i, _cp, _f]),
    "rfk_flame_get_param": (_i, [_vp, _i, _cp, _fpp]),
    "rfk_flame_set_param": (_i, [_vp, _i, _cp, _f]),
    "rfk_flame_get_palette": (_i, [_vp, _fpp]),
    "rfk_flame_set_palette": (_i, [_vp, _fpp]),
    "rfk_flame_buffer_map_json": (_cp, [_vp]),
    "rfk_flame_copy_params": (_i, [_vp, _fpp]),
    "rfk_flame_glsl_source": (_cp, [_vp]),
    "rfk_flame_cuda_source": (_cp, [_vp]),
    "rfk_flame_get_cubin": (_i, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "rfk_flame_variant_source": (_cp, [_vp, _i, _i]),
    "rfk_flame_get_variant_cubin": (_i, [_vp, _i, _i, _vp, _sz, C.POINTER(_sz)]),
    "rfk_flame_uses_specialised": (_i, [_vp]),
    "rfk_flame_pair_particles_state": (_i, [_vp, _fpp]),
    "rfk_flame_get_options": (_i, [_vp, C.POINTER(KernelOptions)]),
    "rfk_flame_set_options": (_i, [_vp, C.POINTER(KernelOptions)]),
    "rfk_flame_kernel_info": (_i, [_vp, _cp, _ipp, _ipp, _ipp]),
    "rfk_flame_needs_warmup": (_i, [_vp]),
    "rfk_flame_warmup": (_i, [_vp, _sz, _f]),
    "rfk_flame_draw_to_bins": (C.c_int64, [_vp, _vp, _sz, _sz, _i]),
    "rfk_flame_draw_to_bins_async": (_i, [_vp, _vp, _sz, _sz, _i]),
    "rfk_flame_binned_total": (C.c_int64, [_vp]),
    "rfk_flame_build_hot_map": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_uint64, C.POINTER(HotMapInfo)]),
    "rfk_flame_clear_hot_map": (C.c_int, [_vp]),
    "rfk_flame_copy_hot_map": (C.c_int64, [_vp, C.POINTER(C.c_uint32), C.c_size_t]),
    "rfk_flame_reset_animation": (_i, [_vp]),
    "rfk_flame_xform_counts": (_i, [_vp, C.POINTER(C.c_uint64), _i]),
    "rfk_flame_screen_affine": (_i, [_vp, _sz, _sz, _fpp]),
    "rfk_flame_rotate_xforms": (_i, [_vp, _f]),
    "rfk_flame_motion_count": (_i, [_vp, _i]),
    "rfk_flame_get_motion": (_i, [_vp, _i, _i, C.POINTER(MotionInfo)]),
    "rfk_flame_apply_motion": (_i, [_vp, _f]),
    "rfk_motion_function": (_f, [_cp, _f]),
    "rfk_rotate_affine": (None, [_fpp, _f, _fpp]),
    "rfk_scale_affine": (None, [_fpp, _f, _fpp]),
    "rfk_translate_affine": (None, [_fpp, _fpp, _fpp]),
    "rfk_flame_post_params": (_i, [_vp, C.POINTER(PostParams)]),
    "rfk_density_estimate": (_i, [_vp, _vp, _sz, _sz, C.POINTER(PostParams)]),
    "rfk_tonemap": (_i, [_vp, _vp, _vp, _sz, _sz, C.POINTER(PostParams)]),
    "rfk_density_tonemap": (_i, [_vp, _vp, _vp, _sz, _sz, C.POINTER(PostParams)]),
    "rfk_density_tonemap_rows": (_i, [_vp, _vp, _vp, _sz, _sz, C.POINTER(PostParams), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "rfk_downsample2x": (_i, [_vp, _vp, _sz, _sz]),
    "rfk_spatial_downsample": (_i, [_vp, _vp, _sz, _sz, _i, _f]),
    "rfk_spatial_filter_taps": (_i, [_i, _f, _fpp]),
    "rfk_seed_rng_states": (_i, [_vp, _sz, C.c_uint32]),
    "rfk_make_sample_points": (_i, [_vp, C.c_uint32]),
    "rfk_make_shuffle_buffers": (_i, [_vp, C.c_uint32, C.c_uint32, C.c_uint64]),
    "rfk_copy_rng_states": (_i, [_upp, _sz, _sz]),
    "rfk_render_frame": (_i, [_vp, C.POINTER(FrameRequest), _vp, _vp, C.POINTER(FrameStats)]),
    "rfk_comm_unique_id": (_i, [_vp]),
    "rfk_comm_init": (_i, [_vp, _i, _i]),
    "rfk_comm_destroy": (_i, []),
    "rfk_comm_rank": (_i, []),
    "rfk_comm_world": (_i, []),
    "rfk_comm_p2p": (_i, []),
    "rfk_comm_barrier": (_i, []),
    "rfk_comm_reduce_histogram": (_i, [_vp, _sz, _i]),
    "rfk_comm_row_slab": (_i, [C.c_uint32, C.c_uint32, _i, _i, C.POINTER(RowSlab)]),
    "rfk_render_frame_sharded": (_i, [_vp, C.POINTER(ShardedRequest), _vp, _vp, C.POINTER(ShardedStats)]),
    "rfk_cache_write_buffer": (_i, [_cp, _cp, _cp, _vp, _sz, _cp, _cp, _sz]),
    "rfk_cache_read_buffer": (C.c_int64, [_cp, _cp, _cp, _cp, _vp, _sz]),
    "rfk_cache_list": (_i, [_cp, _cp, _cp, _cp, _sz]),
    "rfk_export_sim_cache": (_i, [_cp, C.c_uint64]),
    "rfk_import_rng_states": (_i, [_cp, _cp]),
    "rfk_set_rng_states": (_i, [_upp, _sz, _sz]),
    "rfk_write_png": (_i, [_cp, _vp, _sz, _sz]),
    "rfk_write_exr": (_i, [_cp, _vp, _sz, _sz]),
    "rfk_set_shuffle_buffers": (_i, [_upp, _sz, C.c_uint64]),
    "rfk_flame_reference_warmup": (_i, [_vp, _sz, _f, _upp]),
    "rfk_flame_reference_draw_to_bins": (C.c_int64, [_vp, _vp, _sz, _sz, _i, _upp]),
    "rfk_flame_copy_particles": (_i, [_vp, _fpp]),
    "rfk_flame_single_step": (_i, [_vp, _i, _fpp, _ipp, _upp, _fpp, _i, _fpp]),
    "rfk_flame_select_xform": (_i, [_vp, _i, _fpp, _fpp, _ipp]),
    "rfk_flame_bucket_index": (_i, [_vp, _i, _fpp, _fpp, _i, _i, _ipp, _ipp]),
    "rfk_flame_animate": (_i, [_vp, _f, _i, _fpp]),
    "rfk_text_replace_macro": (_i, [_cp, _cp, _cp, _cp, _sz]),
    "rfk_text_find_macros": (_i, [_cp, _cp, _sz]),
}

_lib = None


def lib():
    """The C-ABI library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RefraktError("librefrakt_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def _check(rc, what=""):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise RefraktError("%s failed (%s): %s" % (what, rc, lib().rfk_last_error().decode(errors="replace")))
    return rc


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def set_sim_parameters(total_particles: int, temporal_samples: int, shuffle_count: int = 1024, seed: int = 0):
    """flame::set_sim_parameters (src/flame.hpp:77)"""
    _check(lib().rfk_set_sim_parameters(total_particles, temporal_samples, shuffle_count, seed), "set_sim_parameters")


def set_shuffle_buffers(tables: Optional[np.ndarray] = None, count: int = 64, seed: int = 0):
    """shuffle buffers of the reference pass mode (binding 4): `tables` is count x (P/TS) uint32, or None for device-generated"""
    if tables is not None:
        tables = np.ascontiguousarray(tables, dtype=np.uint32)
        count = tables.shape[0]
    _check(lib().rfk_set_shuffle_buffers(None if tables is None else _ptr(tables, C.c_uint32), count, seed), "set_shuffle_buffers")


def release_buffers():
    """frees the library-owned device memory (simulation state, cached frame buffers); set_sim_parameters + warmup again before drawing"""
    _check(lib().rfk_release_buffers(), "release_buffers")


def kernel_launch_count() -> int:
    return int(lib().rfk_kernel_launch_count())


class DeviceBuffer:
    """storage_buffer<T> (src/buffer_objects.hpp:9-60) on cudaMalloc."""

    def __init__(self, nbytes: int):
        self.nbytes = nbytes
        self.ptr = lib().rfk_device_alloc(nbytes)
        if not self.ptr:
            raise RefraktError("device_alloc: " + lib().rfk_last_error().decode())

    def zero_out(self):
        _check(lib().rfk_device_zero(self.ptr, self.nbytes), "device_zero")

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        _check(lib().rfk_memcpy_to_device(self.ptr, a.ctypes.data, a.nbytes), "memcpy_to_device")

    def download(self, dtype, shape) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        _check(lib().rfk_memcpy_to_host(out.ctypes.data, self.ptr, out.nbytes), "memcpy_to_host")
        return out

    def free(self):
        if self.ptr:
            lib().rfk_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FlameCompiler:
    """flame_compiler (src/variation_table.hpp:20-60)"""

    def __init__(self, variations_yaml: str, overlay: Optional[str] = None):
        self.handle = lib().rfk_compiler_create(variations_yaml.encode())
        if not self.handle:
            raise RefraktError("flame_compiler: " + lib().rfk_last_error().decode())
        if overlay:
            _check(lib().rfk_compiler_load_overlay(self.handle, overlay.encode()), "load_overlay")

    def is_param(self, n): return bool(lib().rfk_compiler_is_param(self.handle, n.encode()))
    def is_variation(self, n): return bool(lib().rfk_compiler_is_variation(self.handle, n.encode()))
    def is_common(self, n): return bool(lib().rfk_compiler_is_common(self.handle, n.encode()))

    def variations(self):
        n = lib().rfk_compiler_variation_count(self.handle)
        return [lib().rfk_compiler_variation_name(self.handle, i).decode() for i in range(n)]

    def get_parameters_for_variation(self, name):
        buf = C.create_string_buffer(4096)
        n = _check(lib().rfk_compiler_get_parameters_for_variation(self.handle, name.encode(), buf, 4096), "get_parameters_for_variation")
        return buf.value.decode().split("\n") if n else []

    def param_owner(self, param):
        r = lib().rfk_compiler_param_owner(self.handle, param.encode())
        return r.decode() if r else None

    def __del__(self):
        try:
            if self.handle:
                lib().rfk_compiler_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Flame:
    """flame (src/flame.hpp:41-174)"""

    def __init__(self, handle, compiler):
        self.handle = handle
        self._compiler = compiler  # keep alive

    @staticmethod
    def load_flame(path: str, compiler: FlameCompiler) -> Optional["Flame"]:
        """flame::load_flame: None on failure, message in last_error()."""
        h = lib().rfk_flame_load(path.encode(), compiler.handle)
        return Flame(h, compiler) if h else None

    @staticmethod
    def load_flame_string(xml_text: str, compiler: FlameCompiler) -> Optional["Flame"]:
        h = lib().rfk_flame_load_string(xml_text.encode(), compiler.handle)
        return Flame(h, compiler) if h else None

    @staticmethod
    def last_error() -> str:
        return lib().rfk_last_error().decode(errors="replace")

    # --- fields
    def info(self) -> FlameInfo:
        i = FlameInfo()
        _check(lib().rfk_flame_get_info(self.handle, C.byref(i)), "get_info")
        return i

    def set_info(self, i: FlameInfo):
        _check(lib().rfk_flame_set_info(self.handle, C.byref(i)), "set_info")

    def xform(self, index: int) -> XformInfo:
        x = XformInfo()
        _check(lib().rfk_flame_get_xform(self.handle, index, C.byref(x)), "get_xform")
        return x

    def set_xform(self, index: int, x: XformInfo):
        _check(lib().rfk_flame_set_xform(self.handle, index, C.byref(x)), "set_xform")

    def variations(self, index: int) -> dict:
        x = self.xform(index)
        out = {}
        for k in range(x.num_variations):
            name = lib().rfk_flame_variation_name(self.handle, index, k)
            v = C.c_float()
            _check(lib().rfk_flame_get_variation(self.handle, index, name, C.byref(v)), "get_variation")
            out[name.decode()] = v.value
        return out

    def params_of(self, index: int) -> dict:
        x = self.xform(index)
        out = {}
        for k in range(x.num_params):
            name = lib().rfk_flame_param_name(self.handle, index, k)
            v = C.c_float()
            _check(lib().rfk_flame_get_param(self.handle, index, name, C.byref(v)), "get_param")
            out[name.decode()] = v.value
        return out

    def set_variation(self, index, name, value): _check(lib().rfk_flame_set_variation(self.handle, index, name.encode(), value), "set_variation")
    def set_param(self, index, name, value): _check(lib().rfk_flame_set_param(self.handle, index, name.encode(), value), "set_param")

    def palette(self) -> np.ndarray:
        out = np.zeros((256, 4), dtype=np.float32)
        _check(lib().rfk_flame_get_palette(self.handle, _ptr(out)), "get_palette")
        return out

    def set_palette(self, pal):
        pal = _f32(pal).reshape(256, 4)
        _check(lib().rfk_flame_set_palette(self.handle, _ptr(pal)), "set_palette")

    # --- generated code
    def buffer_map_json(self) -> str: return lib().rfk_flame_buffer_map_json(self.handle).decode()
    def glsl_source(self) -> str: return lib().rfk_flame_glsl_source(self.handle).decode()
    def cuda_source(self) -> str: return lib().rfk_flame_cuda_source(self.handle).decode()

    def copy_flame_data_to_buffer(self) -> np.ndarray:
        out = np.zeros(1024, dtype=np.float32)
        _check(lib().rfk_flame_copy_params(self.handle, _ptr(out)), "copy_params")
        return out

    def cubin(self) -> bytes:
        size = C.c_size_t()
        _check(lib().rfk_flame_get_cubin(self.handle, None, 0, C.byref(size)), "get_cubin")
        buf = C.create_string_buffer(size.value)
        _check(lib().rfk_flame_get_cubin(self.handle, buf, size.value, C.byref(size)), "get_cubin")
        return buf.raw

    def variant_source(self, staged=False, specialised=False) -> str:
        r = lib().rfk_flame_variant_source(self.handle, int(staged), int(specialised))
        if r is None:
            raise RefraktError("variant_source: " + Flame.last_error())
        return r.decode()

    def variant_cubin(self, staged=False, specialised=False) -> bytes:
        size = C.c_size_t()
        _check(lib().rfk_flame_get_variant_cubin(self.handle, int(staged), int(specialised), None, 0, C.byref(size)), "get_variant_cubin")
        buf = C.create_string_buffer(size.value)
        _check(lib().rfk_flame_get_variant_cubin(self.handle, int(staged), int(specialised), buf, size.value, C.byref(size)), "get_variant_cubin")
        return buf.raw

    def pair_particles_state(self):
        """(state, probe ms): 0 generic kernels, 1 two particles per thread, 2 one; the two probe times that decided it"""
        ms = np.zeros(2, dtype=np.float32)
        st = _check(lib().rfk_flame_pair_particles_state(self.handle, _ptr(ms)), "pair_particles_state")
        return int(st), ms.tolist()

    def uses_specialised(self) -> bool: return bool(_check(lib().rfk_flame_uses_specialised(self.handle), "uses_specialised"))

    def options(self) -> KernelOptions:
        o = KernelOptions()
        _check(lib().rfk_flame_get_options(self.handle, C.byref(o)), "get_options")
        return o

    def set_options(self, **kw):
        o = self.options()
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, int(v))
        _check(lib().rfk_flame_set_options(self.handle, C.byref(o)), "set_options")

    def kernel_info(self, kernel="rfk_draw"):
        r, s, b = C.c_int(), C.c_int(), C.c_int()
        _check(lib().rfk_flame_kernel_info(self.handle, kernel.encode(), C.byref(r), C.byref(s), C.byref(b)), "kernel_info")
        return {"regs": r.value, "smem": s.value, "blocks_per_sm": b.value}

    # --- run
    def needs_warmup(self) -> bool: return bool(_check(lib().rfk_flame_needs_warmup(self.handle), "needs_warmup"))
    def warmup(self, num_passes: int, tss_width: float): _check(lib().rfk_flame_warmup(self.handle, num_passes, tss_width), "warmup")

    def draw_to_bins(self, bins_ptr: int, bins_len: int, bins_width: int, num_iter: int) -> int:
        """bins_ptr: device pointer to bins_len float4. Returns the samples binned by this call."""
        return int(_check(lib().rfk_flame_draw_to_bins(self.handle, bins_ptr, bins_len, bins_width, num_iter), "draw_to_bins"))

    def draw_to_bins_async(self, bins_ptr, bins_len, bins_width, num_iter):
        _check(lib().rfk_flame_draw_to_bins_async(self.handle, bins_ptr, bins_len, bins_width, num_iter), "draw_to_bins_async")

    # --- reference pass mode (same-hardware baseline, pass-level parity hook)
    def reference_warmup(self, num_passes: int, tss_width: float, shuffle_ids=None):
        ids = None if shuffle_ids is None else np.ascontiguousarray(shuffle_ids, dtype=np.uint32).reshape(-1)
        assert ids is None or ids.size == 2 * (1 + num_passes)
        _check(lib().rfk_flame_reference_warmup(self.handle, num_passes, tss_width, None if ids is None else _ptr(ids, C.c_uint32)), "reference_warmup")

    def reference_draw_to_bins(self, bins_ptr, bins_len, bins_width, num_iter, shuffle_ids=None) -> int:
        ids = None if shuffle_ids is None else np.ascontiguousarray(shuffle_ids, dtype=np.uint32).reshape(-1)
        assert ids is None or ids.size == 2 * num_iter
        return int(_check(lib().rfk_flame_reference_draw_to_bins(self.handle, bins_ptr, bins_len, bins_width, num_iter, None if ids is None else _ptr(ids, C.c_uint32)), "reference_draw_to_bins"))

    def copy_particles(self, total_particles: int) -> np.ndarray:
        out = np.zeros((total_particles, 4), dtype=np.float32)
        _check(lib().rfk_flame_copy_particles(self.handle, _ptr(out)), "copy_particles")
        return out

    def build_hot_map(self, bins_ptr, bins_len, bins_width, budget_bytes=0) -> HotMapInfo:
        """kernel option l2_hints: mark the densest 16 x 16-bin tiles of the histogram as worth keeping in L2"""
        info = HotMapInfo()
        _check(lib().rfk_flame_build_hot_map(self.handle, bins_ptr, bins_len, bins_width, budget_bytes, C.byref(info)), "build_hot_map")
        return info

    def clear_hot_map(self): _check(lib().rfk_flame_clear_hot_map(self.handle), "clear_hot_map")

    def hot_map(self) -> np.ndarray:
        """the map as a (tiles_y, tiles_x) bool array; empty when none is active"""
        n = int(_check(lib().rfk_flame_copy_hot_map(self.handle, None, 0), "copy_hot_map"))
        words = np.zeros(n, dtype=np.uint32)
        if n:
            _check(lib().rfk_flame_copy_hot_map(self.handle, _ptr(words, C.c_uint32), n), "copy_hot_map")
        return np.unpackbits(words.view(np.uint8), bitorder="little").astype(bool)

    def binned_total(self) -> int: return int(_check(lib().rfk_flame_binned_total(self.handle), "binned_total"))
    def reset_animation(self): _check(lib().rfk_flame_reset_animation(self.handle), "reset_animation")
    def motion(self, index: int) -> dict:
        """the <motion> entries of an xform: target -> (frequency, function, amplitude)"""
        out = {}
        for k in range(_check(lib().rfk_flame_motion_count(self.handle, index), "motion_count")):
            m = MotionInfo()
            _check(lib().rfk_flame_get_motion(self.handle, index, k, C.byref(m)), "get_motion")
            out[m.target.decode()] = (m.freq, m.function.decode(), m.amplitude)
        return out

    def apply_motion(self, time_seconds: float) -> int:
        return int(_check(lib().rfk_flame_apply_motion(self.handle, time_seconds), "apply_motion"))

    def rotate_xforms(self, degrees: float): _check(lib().rfk_flame_rotate_xforms(self.handle, degrees), "rotate_xforms")

    def xform_counts(self, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint64)
        _check(lib().rfk_flame_xform_counts(self.handle, _ptr(out, C.c_uint64), n), "xform_counts")
        return out

    def screen_space_affine(self, W, H) -> np.ndarray:
        out = np.zeros(6, dtype=np.float32)
        _check(lib().rfk_flame_screen_affine(self.handle, W, H, _ptr(out)), "screen_affine")
        return out

    def post_params(self) -> PostParams:
        p = PostParams()
        _check(lib().rfk_flame_post_params(self.handle, C.byref(p)), "post_params")
        return p

    def render_frame(self, width, height, target_binned=0, max_draw_calls=0, warmup_passes=16, drawing_passes=128, tss_width=1.2 / 60.0,
                     scale_constant_exp=4.0, rgba8_out: Optional[np.ndarray] = None, image_out: Optional[np.ndarray] = None,
                     supersample=1, filter_radius=1.0):
        req = FrameRequest(width, height, warmup_passes, drawing_passes, tss_width, target_binned, max_draw_calls, scale_constant_exp, supersample, filter_radius)
        if rgba8_out is None and image_out is None:
            rgba8_out = np.empty((height, width, 4), dtype=np.uint8)
        stats = FrameStats()
        _check(lib().rfk_render_frame(self.handle, C.byref(req), rgba8_out.ctypes.data if rgba8_out is not None else None,
                                      image_out.ctypes.data if image_out is not None else None, C.byref(stats)), "render_frame")
        return (rgba8_out if rgba8_out is not None else image_out), stats

    def render_frame_sharded(self, width, height, target_binned=0, max_draw_calls=0, warmup_passes=16, drawing_passes=128, tss_width=1.2 / 60.0,
                             scale_constant_exp=4.0, want_rgba8=True, want_image=False, supersample=1, filter_radius=1.0,
                             rgba8_out: Optional[np.ndarray] = None, image_out: Optional[np.ndarray] = None):
        """rfk_render_frame_sharded: collective over the ranks of comm_init; rank 0 gets the image(s), the others None"""
        req = ShardedRequest(FrameRequest(width, height, warmup_passes, drawing_passes, tss_width, target_binned, max_draw_calls, scale_constant_exp, supersample, filter_radius),
                             int(want_rgba8), int(want_image))
        root = comm_rank() == 0
        if root and want_rgba8 and rgba8_out is None:
            rgba8_out = np.empty((height, width, 4), dtype=np.uint8)
        if root and want_image and image_out is None:
            image_out = np.empty((height, width, 4), dtype=np.float32)
        stats = ShardedStats()
        _check(lib().rfk_render_frame_sharded(self.handle, C.byref(req), rgba8_out.ctypes.data if (root and want_rgba8) else None,
                                              image_out.ctypes.data if (root and want_image) else None, C.byref(stats)), "render_frame_sharded")
        return (rgba8_out if root else None), (image_out if root else None), stats

    # --- test hooks
    def single_step(self, xyz, xid, rng, fp=None, first_run=False):
        xyz = _f32(xyz)
        xid = np.ascontiguousarray(xid, dtype=np.int32)
        rng = np.ascontiguousarray(rng, dtype=np.uint32).copy()
        n = xid.shape[0]
        out = np.zeros((n, 4), dtype=np.float32)
        fpp = None
        if fp is not None:
            fp = _f32(fp)
            assert fp.size == 1024
            fpp = _ptr(fp)
        _check(lib().rfk_flame_single_step(self.handle, n, _ptr(xyz), _ptr(xid, C.c_int), _ptr(rng, C.c_uint32), fpp, int(first_run), _ptr(out)), "single_step")
        return out, rng

    def select_xform(self, ratio, fp=None):
        ratio = _f32(ratio)
        out = np.zeros(ratio.shape[0], dtype=np.int32)
        fpp = None
        if fp is not None:
            fp = _f32(fp)
            fpp = _ptr(fp)
        _check(lib().rfk_flame_select_xform(self.handle, ratio.shape[0], _ptr(ratio), fpp, _ptr(out, C.c_int)), "select_xform")
        return out

    def bucket_index(self, xyzw, ss_affine, W, H):
        xyzw = _f32(xyzw)
        ss = _f32(ss_affine)
        n = xyzw.shape[0]
        idx, pal = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        _check(lib().rfk_flame_bucket_index(self.handle, n, _ptr(xyzw), _ptr(ss), W, H, _ptr(idx, C.c_int), _ptr(pal, C.c_int)), "bucket_index")
        return idx, pal

    def animate(self, temporal_samples: int, tss_width: float) -> np.ndarray:
        n = self.info().param_count
        out = np.zeros((temporal_samples, n), dtype=np.float32)
        _check(lib().rfk_flame_animate(self.handle, tss_width, temporal_samples, _ptr(out)), "animate")
        return out

    def __del__(self):
        try:
            if self.handle:
                lib().rfk_flame_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# --- several GPUs, one process per GPU (include/refrakt_b200.h "rfk_comm_*")
def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().rfk_comm_unique_id(buf), "comm_unique_id")
    return buf.raw


def comm_init(unique_id: bytes, rank: int, world: int):
    assert len(unique_id) == 128
    _check(lib().rfk_comm_init(C.create_string_buffer(unique_id, 128), rank, world), "comm_init")


def comm_destroy(): _check(lib().rfk_comm_destroy(), "comm_destroy")
def comm_rank() -> int: return int(lib().rfk_comm_rank())
def comm_world() -> int: return int(lib().rfk_comm_world())
def comm_p2p() -> bool: return bool(lib().rfk_comm_p2p())
def comm_barrier(): _check(lib().rfk_comm_barrier(), "comm_barrier")


def comm_reduce_histogram(bins_ptr: int, bins_len: int, root: int = 0):
    """in-place sum of the per-rank float4 histograms onto `root` (every rank when root < 0); stream-ordered"""
    _check(lib().rfk_comm_reduce_histogram(bins_ptr, bins_len, root), "comm_reduce_histogram")


def comm_row_slab(height: int, halo: int, rank: int, world: int) -> RowSlab:
    s = RowSlab()
    _check(lib().rfk_comm_row_slab(height, halo, rank, world, C.byref(s)), "comm_row_slab")
    return s


def rotate_affine(a, deg):
    a, out = _f32(a), np.zeros(6, dtype=np.float32)
    lib().rfk_rotate_affine(_ptr(a), deg, _ptr(out))
    return out


def scale_affine(a, s):
    a, out = _f32(a), np.zeros(6, dtype=np.float32)
    lib().rfk_scale_affine(_ptr(a), s, _ptr(out))
    return out


def translate_affine(a, t):
    a, t, out = _f32(a), _f32(t), np.zeros(6, dtype=np.float32)
    lib().rfk_translate_affine(_ptr(a), _ptr(t), _ptr(out))
    return out


def density_estimate(bins_ptr, image_ptr, W, H, p: PostParams):
    _check(lib().rfk_density_estimate(bins_ptr, image_ptr, W, H, C.byref(p)), "density_estimate")


def tonemap(in_ptr, out_ptr, rgba8_ptr, W, H, p: PostParams):
    _check(lib().rfk_tonemap(in_ptr, out_ptr, rgba8_ptr, W, H, C.byref(p)), "tonemap")


def density_tonemap(bins_ptr, out_ptr, rgba8_ptr, W, H, p: PostParams):
    _check(lib().rfk_density_tonemap(bins_ptr, out_ptr, rgba8_ptr, W, H, C.byref(p)), "density_tonemap")


def density_tonemap_rows(bins_rows_ptr, out_ptr, rgba8_ptr, W, H, p: PostParams, y0, y1, src_y0, src_y1, out_y0):
    _check(lib().rfk_density_tonemap_rows(bins_rows_ptr, out_ptr, rgba8_ptr, W, H, C.byref(p), y0, y1, src_y0, src_y1, out_y0), "density_tonemap_rows")


def downsample2x(in_ptr, out_ptr, W, H):
    _check(lib().rfk_downsample2x(in_ptr, out_ptr, W, H), "downsample2x")


def spatial_downsample(in_ptr, out_ptr, W, H, supersample, filter_radius):
    _check(lib().rfk_spatial_downsample(in_ptr, out_ptr, W, H, supersample, filter_radius), "spatial_downsample")


def spatial_filter_taps(supersample: int, filter_radius: float) -> np.ndarray:
    taps = np.zeros(64, dtype=np.float32)
    n = _check(lib().rfk_spatial_filter_taps(supersample, filter_radius, _ptr(taps)), "spatial_filter_taps")
    return taps[:n].copy()


def seed_rng_states(count: int, seed_base: int = 0) -> np.ndarray:
    buf = DeviceBuffer(count * 16)
    _check(lib().rfk_seed_rng_states(buf.ptr, count, seed_base), "seed_rng_states")
    out = buf.download(np.uint32, (count, 4))
    buf.free()
    return out


def make_sample_points(count: int) -> np.ndarray:
    buf = DeviceBuffer(count * 16)
    _check(lib().rfk_make_sample_points(buf.ptr, count), "make_sample_points")
    out = buf.download(np.float32, (count, 4))
    buf.free()
    return out


def make_shuffle_buffers(size: int, count: int, seed: int = 0) -> np.ndarray:
    buf = DeviceBuffer(size * count * 4)
    _check(lib().rfk_make_shuffle_buffers(buf.ptr, size, count, seed), "make_shuffle_buffers")
    out = buf.download(np.uint32, (count, size))
    buf.free()
    return out


def replace_macro(text: str, name: str, value: str) -> str:
    """util.cpp:6-9"""
    buf = C.create_string_buffer(len(text) + (len(value) + 2) * (text.count("$") + 1) + 16)
    _check(lib().rfk_text_replace_macro(text.encode(), name.encode(), value.encode(), buf, len(buf)), "replace_macro")
    return buf.value.decode()


def find_macros(text: str) -> set:
    """util.cpp:11-23"""
    buf = C.create_string_buffer(len(text) + 16)
    _check(lib().rfk_text_find_macros(text.encode(), buf, len(buf)), "find_macros")
    return set(buf.value.decode().split("\n")) - {""}


class BufferGroup:
    """buffer_cache::buffer_group (src/buffer_cache.hpp:12-60): <root>/cache/<type>/<group>/<name>.bin"""

    def __init__(self, root: str, type_: str, group: str):
        self.root, self.type, self.group = root.encode(), type_.encode(), str(group).encode()

    def write_buffer(self, data: np.ndarray, name: Optional[str] = None) -> str:
        data = np.ascontiguousarray(data)
        out = C.create_string_buffer(128)
        _check(lib().rfk_cache_write_buffer(self.root, self.type, self.group, data.ctypes.data, data.nbytes, name.encode() if name else None, out, 128), "cache_write_buffer")
        return out.value.decode()

    def read_buffer(self, name: str, dtype=np.uint8) -> np.ndarray:
        n = _check(lib().rfk_cache_read_buffer(self.root, self.type, self.group, name.encode(), None, 0), "cache_read_buffer")
        raw = np.empty(n, dtype=np.uint8)
        _check(lib().rfk_cache_read_buffer(self.root, self.type, self.group, name.encode(), raw.ctypes.data, n), "cache_read_buffer")
        return raw.view(dtype)

    def cached_buffers(self):
        buf = C.create_string_buffer(1 << 20)
        n = _check(lib().rfk_cache_list(self.root, self.type, self.group, buf, len(buf)), "cache_list")
        return buf.value.decode().split("\n") if n else []


def export_sim_cache(root: str, shuffle_seed: int = 0):
    _check(lib().rfk_export_sim_cache(root.encode(), shuffle_seed), "export_sim_cache")


def import_rng_states(root: str, name: Optional[str] = None):
    _check(lib().rfk_import_rng_states(root.encode(), name.encode() if name else None), "import_rng_states")


def set_rng_states(states: np.ndarray, first: int = 0):
    states = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, 4)
    _check(lib().rfk_set_rng_states(_ptr(states, C.c_uint32), first, states.shape[0]), "set_rng_states")


def write_png(path: str, rgba8: np.ndarray):
    """screenshot (src/main.cpp:590-593): rgba8 is H x W x 4 uint8, rows top to bottom"""
    rgba8 = np.ascontiguousarray(rgba8, dtype=np.uint8)
    h, w = rgba8.shape[:2]
    _check(lib().rfk_write_png(path.encode(), rgba8.ctypes.data, w, h), "write_png")


def write_exr(path: str, rgba32f: np.ndarray):
    """the float frame as an uncompressed scanline OpenEXR file; rgba32f is H x W x 4 float32, rows top to bottom"""
    rgba32f = np.ascontiguousarray(rgba32f, dtype=np.float32)
    h, w = rgba32f.shape[:2]
    _check(lib().rfk_write_exr(path.encode(), rgba32f.ctypes.data, w, h), "write_exr")


def copy_rng_states(first: int, count: int) -> np.ndarray:
    out = np.zeros((count, 4), dtype=np.uint32)
    _check(lib().rfk_copy_rng_states(_ptr(out, C.c_uint32), first, count), "copy_rng_states")
    return out
