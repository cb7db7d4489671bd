import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

FIXTURES = os.path.join(ROOT, "tests", "fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden")
GENOME = os.path.join(FIXTURES, "electricsheep.247.11256.flam3")
VARIATIONS = os.path.join(FIXTURES, "variations.yaml")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    import refrakt_b200
    if not os.path.exists(refrakt_b200.LIB_PATH):
        from refrakt_b200 import build
        build.build()


@pytest.fixture(scope="session")
def rfk():
    _ensure_built()
    import refrakt_b200
    return refrakt_b200


@pytest.fixture(scope="session")
def compiler(rfk):
    return rfk.FlameCompiler(VARIATIONS)


@pytest.fixture(scope="session")
def flame(rfk, compiler):
    f = rfk.Flame.load_flame(GENOME, compiler)
    assert f is not None, rfk.Flame.last_error()
    return f


@pytest.fixture(scope="session")
def oracle_mod():
    import refrakt_oracle
    return refrakt_oracle


@pytest.fixture(scope="session")
def vt(oracle_mod):
    return oracle_mod.VariationTable(VARIATIONS)


@pytest.fixture(scope="session")
def oracle(oracle_mod, vt):
    return oracle_mod.Oracle(oracle_mod.load_flame(GENOME, vt), vt)


@pytest.fixture(scope="session")
def gpu_ready(rfk):
    """Fails (not skips) when the CUDA path cannot run: GPU tests must never pass on a fallback."""
    import ctypes
    rc = rfk.lib().rfk_set_device(0)
    assert rc == 0, "no CUDA device: " + rfk.lib().rfk_last_error().decode()
    return True


COMPILE_CLEAN = """linear sinusoidal spherical swirl horseshoe polar handkerchief heart disc spiral hyperbolic diamond ex julia bent
waves fisheye popcorn exponential power cosine rings fan blob pdj fan2 rings2 eyefish bubble perspective noise julian juliascope blur
gaussian_blur radial_blur pie ngon curl rectangles arch tangent square rays blade secant2 cross disc2 super_shape flower conic parabola
boarders butterfly curve foci loonie exp log sin cos sinh pre_blur waves2 cylinder auger flux mobius""".split()
BROKEN = "twintrian bent2 bipolar cell cpow edisc oscope coth".split()

GENOME_TEMPLATE = """<flame name="t" size="640 480" center="0 0" scale="120" rotate="0" brightness="4" gamma="4" vibrancy="1"
 estimator_radius="9" estimator_curve="0.4">
%s
 <color index="0" rgb="255 0 0"/><color index="255" rgb="0 0 255"/>
</flame>"""


def xform_xml(names, vt, rng, tag="xform"):
    attrs = []
    for n in names:
        attrs.append('%s="%.4f"' % (n, rng.uniform(0.1, 0.9)))
        for p in vt.vars[n].param:
            attrs.append('%s="%.4f"' % (p, rng.uniform(0.5, 3.0)))
    coefs = " ".join("%.4f" % v for v in rng.normal(0, 0.6, 6))
    weight = 'weight="%.3f" ' % rng.uniform(0.2, 1) if tag == "xform" else ""
    return '<%s %scolor="%.3f" color_speed="0.5" animate="1" %s coefs="%s" opacity="1"/>' % (tag, weight, rng.random(), " ".join(attrs), coefs)


def chunk_genome(chunk, vt, names=None):
    """synthetic genome using every third compile-clean variation of `chunk` (3 variations per xform)"""
    import numpy as np
    rng = np.random.default_rng(chunk)
    names = COMPILE_CLEAN[chunk::6] if names is None else names
    return GENOME_TEMPLATE % "\n".join(xform_xml(names[i:i + 3], vt, rng) for i in range(0, len(names), 3))


def stress_genome(vt):
    """BASELINE configs[4]: 12 xforms + final xform mixing divergent variations (julian, juliascope, trig, bipolar ...),
    weights/affines from numpy default_rng(247), palette of the shipped genome"""
    import numpy as np
    rng = np.random.default_rng(247)
    pool = ["julian", "juliascope", "sin", "cos", "sinh", "exp", "log", "swirl", "sinusoidal", "polar", "disc", "heart", "bipolar"]
    xforms = []
    for i in range(12):
        k = 2 + int(rng.integers(0, 2))
        names = sorted(set(["linear"] + [pool[int(j)] for j in rng.choice(len(pool), k, replace=False)]))
        xforms.append(xform_xml(names, vt, rng))
    xforms.append(xform_xml(["linear", "julian"], vt, rng, tag="finalxform"))
    palette = [l for l in open(GENOME).read().split("\n") if "<color " in l]
    body = "\n".join(xforms + palette)
    return ('<flame name="stress247" size="800 592" center="0 0" scale="140" rotate="0" brightness="20" gamma="4" vibrancy="1" '
            'estimator_radius="11" estimator_curve="0.6">\n%s\n</flame>' % body)


@pytest.fixture(scope="session")
def overlay_compiler(rfk):
    return rfk.FlameCompiler(VARIATIONS, overlay=rfk.OVERLAY_YAML)


@pytest.fixture(scope="session")
def overlay_vt(rfk, oracle_mod):
    return oracle_mod.VariationTable(VARIATIONS, overlay=rfk.OVERLAY_YAML)
