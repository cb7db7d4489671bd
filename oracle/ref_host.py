"""TEST INFRASTRUCTURE — Python driver of oracle/_ref/libref_host.so: the reference's own host code (flame.cpp,
variation_table.cpp, ...) compiled from /root/reference and run over the software GL of oracle/softgl/, which executes the
reference's GLSL on the CPU (see oracle/ref_host.cpp). Only available where /root/reference exists (the build container);
tests/golden/make_reference_golden.py uses it to write the committed fixtures, tests use it directly when present.
Never imported by the product."""
import ctypes
import json
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RFK_REFERENCE", "/root/reference")
LIB = os.path.join(HERE, "_ref", "libref_host.so")


def libm_probe() -> str:
    """Digest of this machine's libm on fixed inputs. The device fixtures were produced by g++ + glibc; a bit-for-bit
    comparison against them is only meaningful where libm rounds the same way (glibc picks FMA / non-FMA variants per CPU)."""
    import hashlib
    libm = ctypes.CDLL("libm.so.6")
    x = np.random.default_rng(1234).normal(0, 3, 4096).astype(np.float32)
    h = hashlib.sha256()
    for name in ("sinf", "cosf", "tanf", "expf", "logf", "sinhf", "coshf", "atanf", "sqrtf"):
        f = getattr(libm, name)
        f.restype, f.argtypes = ctypes.c_float, [ctypes.c_float]
        h.update(np.array([f(float(abs(v)) if name in ("logf", "sqrtf") else float(v)) for v in x], dtype=np.float32).tobytes())
    for name in ("powf", "atan2f"):
        f = getattr(libm, name)
        f.restype, f.argtypes = ctypes.c_float, [ctypes.c_float, ctypes.c_float]
        h.update(np.array([f(float(abs(a)), float(b)) for a, b in zip(x, x[::-1])], dtype=np.float32).tobytes())
    return h.hexdigest()


def available() -> bool:
    return os.path.exists(LIB) and os.path.isdir(os.path.join(REF, "shaders"))


class ReferenceHost:
    """One process-wide instance: the reference keeps its simulation buffers in class statics."""

    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
            cls._instance._setup()
        return cls._instance

    def _setup(self):
        # the reference reads shaders/ and variations.yaml and writes cache/ relative to the working directory
        self.sandbox = tempfile.mkdtemp(prefix="rfk_refhost_")
        os.symlink(os.path.join(REF, "shaders"), os.path.join(self.sandbox, "shaders"))
        os.symlink(os.path.join(REF, "variations.yaml"), os.path.join(self.sandbox, "variations.yaml"))
        shader_dir = os.path.join(HERE, "_ref", "softgl")
        os.makedirs(shader_dir, exist_ok=True)
        os.environ["RFK_SOFTGL_DIR"] = shader_dir
        os.environ["RFK_SOFTGL_TRANSLATOR"] = os.path.join(HERE, "softgl", "glsl_to_cpp.py")
        self.lib = ctypes.CDLL(LIB)
        for f in ("ref_host_describe", "ref_host_run", "ref_host_draw", "ref_host_buffer", "ref_host_set_bins", "ref_host_post"):
            getattr(self.lib, f).restype = ctypes.c_long

    def _call(self, fn, *args):
        buf = ctypes.create_string_buffer(1 << 24)
        cwd = os.getcwd()
        os.chdir(self.sandbox)
        try:
            n = fn(*args, buf, len(buf))
        finally:
            os.chdir(cwd)
        assert n > 0, "libref_host call failed"
        return json.loads(buf.value.decode())

    def describe(self, genome_path, W, H):
        return self._call(self.lib.ref_host_describe, os.path.abspath(genome_path).encode(), ctypes.c_ulong(W), ctypes.c_ulong(H))

    def run(self, genome_path, P, TS, n_shuffle, warmup_passes, tss_width, W, H, draw_passes):
        """flame::set_sim_parameters(P, TS, n_shuffle); load_flame; warmup(warmup_passes, tss_width); draw_to_bins(W x H, draw_passes)"""
        self.dims = (H, W)
        return self._call(self.lib.ref_host_run, os.path.abspath(genome_path).encode(), ctypes.c_ulong(P), ctypes.c_ulong(TS), ctypes.c_ulong(n_shuffle),
                          ctypes.c_ulong(warmup_passes), ctypes.c_float(tss_width), ctypes.c_ulong(W), ctypes.c_ulong(H), ctypes.c_int(draw_passes))

    def draw(self, W, H, draw_passes):
        """a further flame::draw_to_bins(bins(W x H, zeroed), W, draw_passes) on the flame of the last run()"""
        self.dims = (H, W)
        return self._call(self.lib.ref_host_draw, ctypes.c_ulong(W), ctypes.c_ulong(H), ctypes.c_int(draw_passes))

    def buffer(self, which, dtype, shape=None):
        n = self.lib.ref_host_buffer(which.encode(), None, 0)
        assert n >= 0, which
        out = np.empty(n // np.dtype(dtype).itemsize, dtype=dtype)
        if n:
            self.lib.ref_host_buffer(which.encode(), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(n))
        return out if shape is None else out.reshape(shape)

    def set_bins(self, bins):
        bins = np.ascontiguousarray(bins, dtype=np.float32)
        H, W = bins.shape[:2]
        self.lib.ref_host_set_bins(bins.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_ulong(W), ctypes.c_ulong(H))
        self.dims = (H, W)

    def post(self, estimator_radius, estimator_min, estimator_curve, gamma, brightness, vibrancy, scale_constant=4.0):
        """the density-estimation draw + tonemap dispatch of main.cpp:490-535 on the current bins; returns (density, tonemapped)"""
        cwd = os.getcwd()
        os.chdir(self.sandbox)
        try:
            rc = self.lib.ref_host_post(int(estimator_radius), int(estimator_min), ctypes.c_float(estimator_curve), ctypes.c_float(gamma),
                                        ctypes.c_float(brightness), ctypes.c_float(vibrancy), ctypes.c_double(scale_constant))
        finally:
            os.chdir(cwd)
        assert rc == 0
        H, W = self.dims
        return self.buffer("density", np.float32, (H, W, 4)), self.buffer("tonemapped", np.float32, (H, W, 4))
