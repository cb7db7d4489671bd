"""The reference-side binding of INTEGRATION.md, compiled and run: tests/integration/flame_b200.cpp implements the member
functions of refrakt's own `struct flame` (its header included from /root/reference, not copied) over the C ABI, and
binding_demo.cpp drives them the way src/main.cpp does. The binary is built where the reference is present
(tests/integration/Makefile -> oracle/_ref/ref_binding_demo) and travels to the GPU box."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import FIXTURES, GENOME, ROOT

DEMO = os.path.join(ROOT, "oracle", "_ref", "ref_binding_demo")


def _demo(*args):
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "integration")], check=True)
    if not os.path.exists(DEMO):
        pytest.skip("no prebuilt oracle/_ref/ref_binding_demo and no reference to build it from")
    r = subprocess.run([DEMO, os.path.basename(GENOME)] + [str(a) for a in args], cwd=FIXTURES, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]), r.stdout


def test_binding_compiles_against_the_reference_headers_and_loads_a_genome(rfk):
    """no GPU needed: refrakt's flame::load_flame, bound to the library, returns refrakt's own flame object with the parsed
    fields; the reference's flame_compiler (its variation_table.cpp, compiled from source) lives beside it"""
    d, out = _demo(64, 36, 256 * 4, 4, 2)
    assert d["loaded"] and d["xforms"] == 10 and d["has_final"] == 1 and d["variation_julian_known"] == 1
    assert abs(d["scale"] - 215.921005) < 1e-4 and d["gamma"] == 4
    if rfk.lib().rfk_set_device(0) != 0:  # no device here: the binding must report it, not fall back
        assert d["warmed"] == 0 and d["binned"] == 0 and "CUDA" in d["error"] and "set_sim_parameters:" in out


@pytest.mark.gpu
def test_binding_renders_through_refrakts_own_flame_interface(gpu_ready, rfk, compiler, tmp_path):
    """flame::set_sim_parameters / load_flame / warmup / draw_to_bins called on refrakt's class, executed by the B200 library:
    same binned count as the same calls made directly on the C ABI (same seed, same parameters), and a PNG comes out"""
    W, H, P, TS, passes = 320, 180, 256 * 64, 16, 32
    png = tmp_path / "binding.png"
    d, out = _demo(W, H, P, TS, passes, png)
    assert d["loaded"] and d["warmed"] == 1 and d["image"] == 1 and d["binned"] > 0, out
    assert png.exists() and png.read_bytes()[:8] == b"\x89PNG\r\n\x1a\n"
    f = rfk.Flame.load_flame(GENOME, compiler)
    rfk.set_sim_parameters(P, TS, 1024, seed=0)
    f.warmup(16, np.float32(1.2) / np.float32(60.0))
    buf = rfk.DeviceBuffer(W * H * 16)
    buf.zero_out()
    assert f.draw_to_bins(buf.ptr, W * H, W, passes) == d["binned"]
    buf.free()
