"""TEST INFRASTRUCTURE — parity oracle for refrakt_b200. Not part of the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. The product (refrakt_b200/) never does and has no CPU path.

A CPU restatement of the reference's render path (untrioctium/refrakt), following the
reference source function by function:

  flame_compiler ctor        /root/reference/src/variation_table.cpp:174-215  -> VariationTable
  replace_macro/find_macros  /root/reference/src/util.cpp:6-23                -> replace_macro, find_macros
  flame::load_flame          /root/reference/src/flame.cpp:160-226            -> load_flame
  make_shader_buffer_map     /root/reference/src/flame.cpp:33-71              -> make_buffer_map
  copy_flame_data_to_buffer  /root/reference/src/flame.cpp:73-103             -> copy_flame_data_to_buffer
  compile_flame_xforms       /root/reference/src/variation_table.cpp:78-169, :217-265 -> compile_flame_xforms
  rotate/scale/translate     /root/reference/src/flame.hpp:97-128             -> rotate_affine, ...
  draw_to_bins ss_affine     /root/reference/src/flame.cpp:289-296            -> screen_space_affine
  the GLSL itself            oracle_core.hpp (see its header for the shader-by-shader map)

The generated GLSL is compiled for the CPU against glsl_shim.hpp with g++ (-ffp-contract=off)
into oracle/_build/ and driven through ctypes.

PARITY PINNED against the reference's own code run here. The reference ships no tests, golden vectors or fixtures
 and cannot be built as a whole (GL + 12 fetched dependencies). What compiles from the reference sources
where they lie (oracle/Makefile -> oracle/_ref/):
  libref_pins*.so   jsf32, hammersley, replace_macro/find_macros, the buffer_cache reader, the
                    affine helpers of flame.hpp  -> tests/test_oracle_golden.py
  libref_host.so    the reference's own flame.cpp + variation_table.cpp + ... against stand-in third-party headers
                    (oracle/stubs/) and a software GL that runs its GLSL on the CPU (oracle/softgl/)
                    -> tests/golden/reference_host_*.json.gz, reference_device_*.npz
                    -> tests/test_reference_golden.py, tests/test_reference_device_golden.py compare this restatement
                       AND the product, bit for bit where the arithmetic is the same.
"""
from __future__ import annotations

import ctypes
import hashlib
import math
import os
import re
import subprocess
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from fractions import Fraction
from typing import Dict, List, Optional

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
f32 = np.float32


# ----------------------------------------------------------------------------------------
# util.cpp:6-23
# ----------------------------------------------------------------------------------------
def replace_macro(s: str, name: str, value: str) -> str:
    return re.sub(r"\$" + re.escape(name) + r"([^a-zA-Z0-9_])", lambda m: value + m.group(1), s)


def find_macros(s: str) -> set:
    return set(re.findall(r"\$([a-z0-9_]+)", s))


# ----------------------------------------------------------------------------------------
# number parsing with the C library's semantics
# ----------------------------------------------------------------------------------------
_NUM = re.compile(r"^[ \t\n\r\f\v]*([+-]?(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?))")


def strtod_prefix(s: str) -> float:
    """strtod on the longest numeric prefix; 0.0 when there is none (pugixml as_float, as_int)."""
    m = _NUM.match(s)
    return float(m.group(1)) if m else 0.0


def strtof(s: str) -> np.float32:
    """Correctly rounded decimal -> binary32 (std::stof), avoiding double rounding."""
    m = _NUM.match(s)
    if not m:
        raise ValueError("stof: no conversion")
    exact = Fraction(m.group(1))
    guess = f32(float(exact))
    best = guess
    for cand in (np.nextafter(guess, f32(-np.inf)), np.nextafter(guess, f32(np.inf))):
        if not np.isfinite(cand):
            continue
        dc, db = abs(Fraction(float(cand)) - exact), abs(Fraction(float(best)) - exact)
        if dc < db or (dc == db and (int(f32(cand).view(np.uint32)) & 1) == 0):
            best = f32(cand)
    return best


def stod_to_float(s: str) -> np.float32:
    """(float) std::stod(v)  (flame.cpp:194-195)"""
    m = _NUM.match(s)
    if not m:
        raise ValueError("stod: no conversion")
    return f32(float(m.group(1)))


def stoi(s: str) -> int:
    m = re.match(r"^\s*([+-]?\d+)", s)
    if not m:
        raise ValueError("stoi: no conversion")
    return int(m.group(1))


# ----------------------------------------------------------------------------------------
# variation table (variation_table.cpp:174-215)
# ----------------------------------------------------------------------------------------
def default_replace_macros(s: str) -> str:
    s = replace_macro(s, "x", "v.x")
    s = replace_macro(s, "y", "v.y")
    s = replace_macro(s, "v", "v.xy")
    s = replace_macro(s, "result", "result")
    return s


@dataclass
class VariationDefinition:
    source: str = ""
    result: str = ""
    param: List[str] = field(default_factory=list)
    flags: set = field(default_factory=set)


class VariationTable:
    def __init__(self, path: str, overlay: Optional[str] = None):
        self.common: Dict[str, str] = {}
        self.vars: Dict[str, VariationDefinition] = {}
        self.param_owners: Dict[str, str] = {}
        self._load(path)
        if overlay:
            self._load(overlay)

    def _load(self, path: str):
        with open(path) as fh:
            defs = yaml.safe_load(fh)
        for name, body in (defs.get("variations") or {}).items():
            body = body or {}
            vd = VariationDefinition()
            vd.source = default_replace_macros(str(body.get("src", "") or ""))
            vd.result = default_replace_macros(str(body.get("result", "") or ""))
            for pname in (body.get("param") or {}):
                vd.param.append(pname)
                self.param_owners[pname] = name
            for flag in (body.get("flags") or []):
                vd.flags.add(str(flag))
            self.vars[name] = vd
        for name, src in (defs.get("common") or {}).items():
            self.common[name] = default_replace_macros(str(src))

    def is_param(self, n): return n in self.param_owners
    def is_variation(self, n): return n in self.vars
    def is_common(self, n): return n in self.common


# ----------------------------------------------------------------------------------------
# flame model (flame.hpp:21-60) and parser (flame.cpp:160-226)
# ----------------------------------------------------------------------------------------
@dataclass
class Xform:
    affine: List[np.float32] = field(default_factory=lambda: [f32(0)] * 6)
    post: Optional[List[np.float32]] = None
    variations: Dict[str, np.float32] = field(default_factory=dict)
    var_param: Dict[str, np.float32] = field(default_factory=dict)
    weight: np.float32 = f32(0)
    color: np.float32 = f32(0)
    color_speed: np.float32 = f32(0)
    rotation_frequency: np.float32 = f32(0)
    opacity: np.float32 = f32(0)
    motion: Dict[str, tuple] = field(default_factory=dict)  # flame.hpp:36: target -> (freq, function, amplitude); flame.cpp:199-210 as intended


@dataclass
class Flame:
    xforms: List[Xform] = field(default_factory=list)
    final_xform: Optional[Xform] = None
    palette: np.ndarray = field(default_factory=lambda: np.zeros((256, 4), dtype=np.float32))
    size: List[int] = field(default_factory=lambda: [0, 0])
    center: List[np.float32] = field(default_factory=lambda: [f32(0), f32(0)])
    scale: np.float32 = f32(0)
    rotate: np.float32 = f32(0)
    estimator_min: int = 0
    estimator_radius: int = 0
    estimator_curve: np.float32 = f32(0)
    gamma: np.float32 = f32(0)
    vibrancy: np.float32 = f32(0)
    brightness: np.float32 = f32(0)
    buffer_map: dict = field(default_factory=dict)


def _parse_strings(text: Optional[str], n: int, conv, zero):
    out = [zero] * n
    if text is None:
        return out
    for i, tok in enumerate(text.split()[:n]):
        out[i] = conv(tok)
    return out


def load_flame(path: str, vt: VariationTable) -> Optional[Flame]:
    with open(path) as fh:
        return load_flame_string(fh.read(), vt)


def load_flame_string(text: str, vt: VariationTable) -> Optional[Flame]:
    root = ET.fromstring(text)
    node = root if root.tag == "flame" else root.find("flame")
    if node is None:
        return None
    a = node.attrib
    as_float = lambda k: f32(strtod_prefix(a.get(k, "")))
    as_int = lambda k: int(strtod_prefix(a.get(k, "")))
    f = Flame()
    f.center = _parse_strings(a.get("center"), 2, strtof, f32(0))
    f.scale, f.rotate = as_float("scale"), as_float("rotate")
    f.estimator_curve = as_float("estimator_curve")
    f.estimator_min = as_int("estimator_min")  # sic (flame.cpp:171)
    f.estimator_radius = as_int("estimator_radius")
    f.brightness, f.gamma, f.vibrancy = as_float("brightness"), as_float("gamma"), as_float("vibrancy")
    f.size = _parse_strings(a.get("size"), 2, stoi, 0)
    bad = False
    for child in node:
        if child.tag in ("xform", "finalxform"):
            x = Xform()
            for name, val in child.attrib.items():
                fv = f32(strtod_prefix(val))
                if name == "weight": x.weight = fv
                elif name == "color": x.color = fv
                elif name == "color_speed": x.color_speed = fv
                elif name == "animate": x.rotation_frequency = f32(1.0) if (fv > 0 and child.tag != "final_xform") else f32(0.0)
                elif name == "opacity": x.opacity = fv
                elif vt.is_param(name): x.var_param[name] = fv
                elif vt.is_variation(name): x.variations[name] = fv
                elif name == "coefs": x.affine = _parse_strings(val, 6, stod_to_float, f32(0))
                elif name == "post": x.post = _parse_strings(val, 6, stod_to_float, f32(0))
                else: bad = True
            for m in child:  # the commented-out block of flame.cpp:199-210, with the attribute value as the amplitude
                if m.tag != "motion":
                    continue
                freq = f32(strtod_prefix(m.attrib.get("motion_frequency", "")))
                fn = m.attrib.get("motion_function", "")
                for name, val in m.attrib.items():
                    if name not in ("motion_frequency", "motion_function"):
                        x.motion[name] = (freq, fn, f32(strtod_prefix(val)))
            x.motion = dict(sorted(x.motion.items()))
            # std::map iterates alphabetically
            x.variations = dict(sorted(x.variations.items()))
            x.var_param = dict(sorted(x.var_param.items()))
            if child.tag == "finalxform": f.final_xform = x
            else: f.xforms.append(x)
        elif child.tag == "color":
            idx = int(strtod_prefix(child.attrib.get("index", "0")))
            rgb = _parse_strings(child.attrib.get("rgb"), 4, lambda v: f32(stoi(v)) / f32(256.0), f32(0))
            f.palette[idx] = rgb
            f.palette[idx][3] = 1.0
    if bad:
        return None
    f.buffer_map = make_buffer_map(f)
    return f


def motion_function(name: str, x) -> np.float32:
    """flam3's motion functions of the phase x = frequency * time (the reference declares motion_info and never evaluates it)"""
    import math
    x = float(f32(x))
    if name == "sin":
        return f32(math.sin(2.0 * math.pi * x))
    if name == "hill":
        return f32((1.0 - math.cos(2.0 * math.pi * x)) * 0.5)
    if name == "triangle":
        fr = math.fmod(x, 1.0)
        if fr < 0.0:
            fr += 1.0
        if fr <= 0.25:
            return f32(4.0 * fr)
        if fr <= 0.75:
            return f32(-4.0 * fr + 2.0)
        return f32(4.0 * fr - 4.0)
    return f32(0.0)


def apply_motion(f: Flame, base: Flame, time) -> int:
    """f's animated fields = base's loaded values + amplitude * function(freq * time); returns the fields written"""
    written = 0
    pairs = list(zip(f.xforms, base.xforms)) + ([(f.final_xform, base.final_xform)] if f.final_xform is not None else [])
    for x, b in pairs:
        for name, (freq, fn, amp) in b.motion.items():
            delta = f32(amp * motion_function(fn, f32(freq * f32(time))))
            if name in ("weight", "color", "color_speed", "opacity"):
                setattr(x, name, f32(getattr(b, name) + delta))
            elif name in x.variations:
                x.variations[name] = f32(b.variations[name] + delta)
            elif name in x.var_param:
                x.var_param[name] = f32(b.var_param[name] + delta)
            else:
                continue
            written += 1
    return written


# ----------------------------------------------------------------------------------------
# flame.cpp:33-103
# ----------------------------------------------------------------------------------------
def make_buffer_map(f: Flame) -> dict:
    counter = 0

    def make_xform_map(x: Xform):
        nonlocal counter
        m = {"meta": {"start": counter}}
        m["weight"] = counter; counter += 1
        m["affine"] = list(range(counter, counter + 6)); counter += 6
        if x.post is not None:
            m["post"] = list(range(counter, counter + 6)); counter += 6
        m["variations"] = {}
        for k in x.variations:
            m["variations"][k] = counter; counter += 1
        m["param"] = {}
        for k in x.var_param:
            m["param"][k] = counter; counter += 1
        for key in ("color", "color_speed", "opacity", "rotation_frequency"):
            m[key] = counter; counter += 1
        m["meta"]["end"] = counter - 1
        m["meta"]["size"] = counter - m["meta"]["start"]
        return m

    bm = {"xforms": [make_xform_map(x) for x in f.xforms]}
    if f.final_xform is not None:
        bm["final_xform"] = make_xform_map(f.final_xform)
    bm["size"] = counter
    return bm


def copy_flame_data_to_buffer(f: Flame) -> np.ndarray:
    buf = np.zeros(1024, dtype=np.float32)
    normal_weight = f32(0.0)
    for x in f.xforms:
        normal_weight = f32(normal_weight + x.weight)

    def push(x: Xform, m: dict):
        with np.errstate(divide="ignore", invalid="ignore"):
            buf[m["weight"]] = f32(x.weight) / normal_weight
        for i in range(6): buf[m["affine"][i]] = x.affine[i]
        for n, w in x.variations.items(): buf[m["variations"][n]] = w
        for n, v in x.var_param.items(): buf[m["param"][n]] = v
        if x.post is not None:
            for i in range(6): buf[m["post"][i]] = x.post[i]
        buf[m["color"]] = x.color
        buf[m["opacity"]] = x.opacity
        buf[m["color_speed"]] = x.color_speed
        buf[m["rotation_frequency"]] = x.rotation_frequency

    for x, m in zip(f.xforms, f.buffer_map["xforms"]):
        push(x, m)
    if f.final_xform is not None:
        push(f.final_xform, f.buffer_map["final_xform"])
    return buf


# ----------------------------------------------------------------------------------------
# flame.hpp:97-128, flame.cpp:289-296 — binary32 arithmetic with the C library's sinf/cosf
# ----------------------------------------------------------------------------------------
_libm = ctypes.CDLL("libm.so.6")
_libm.sinf.restype = ctypes.c_float; _libm.sinf.argtypes = [ctypes.c_float]
_libm.cosf.restype = ctypes.c_float; _libm.cosf.argtypes = [ctypes.c_float]


def rotate_affine(a, deg):
    a = [f32(v) for v in a]
    rad = f32(0.01745329251) * f32(deg)
    sino, coso = f32(_libm.sinf(float(rad))), f32(_libm.cosf(float(rad)))
    r = list(a)
    r[0] = f32(f32(a[0] * coso) + f32(a[2] * sino))
    r[1] = f32(f32(a[1] * coso) + f32(a[3] * sino))
    r[2] = f32(f32(a[2] * coso) - f32(a[0] * sino))
    r[3] = f32(f32(a[3] * coso) - f32(a[1] * sino))
    return r


def scale_affine(a, s):
    a = [f32(v) for v in a]
    s = f32(s)
    return [f32(a[0] * s), f32(a[1] * s), f32(a[2] * s), f32(a[3] * s), a[4], a[5]]


def translate_affine(a, t):
    a = [f32(v) for v in a]
    t = [f32(t[0]), f32(t[1])]
    r = list(a)
    r[4] = f32(f32(f32(r[0] * t[0]) + f32(r[2] * t[1])) + a[4])
    r[5] = f32(f32(f32(r[1] * t[0]) + f32(r[3] * t[1])) + a[5])
    return r


def screen_space_affine(f: Flame, W: int, H: int):
    base = [f32(1), f32(0), f32(0), f32(1), f32(0), f32(0)]
    base = translate_affine(base, [f32(W) / f32(2.0), f32(H) / f32(2.0)])
    base = scale_affine(base, f32(f32(f.scale * f32(H)) / f32(f.size[1])))
    base = rotate_affine(base, f.rotate)
    base = translate_affine(base, [-f.center[0], -f.center[1]])
    return np.array(base, dtype=np.float32)


# ----------------------------------------------------------------------------------------
# compile_flame_xforms (variation_table.cpp:78-169, :217-265)
# ----------------------------------------------------------------------------------------
def _fp(slot: int) -> str:
    return "fp[%d]" % slot


def _get_ordering(adj: Dict[str, set]) -> List[str]:
    """variation_table.cpp:25-47; returns the stack bottom-to-top."""
    visited = {k: False for k in adj}
    stack: List[str] = []

    def recurse(v):
        visited[v] = True
        for con in sorted(adj[v]):
            if not visited.get(con, False):
                recurse(con)
        stack.append(v)

    for v in sorted(adj):
        if not visited[v]:
            recurse(v)
    return stack


def _resolve_coefs(src: str, slots: List[int], prefix: str) -> str:
    for c in range(3):
        for r in range(2):
            src = replace_macro(src, "%s%d%d" % (prefix, c, r), _fp(slots[c * 2 + r]))
    return src


def _xform_text(x: Xform, m: dict, vt: VariationTable):
    if x.post is None and list(x.variations) == ["linear"]:
        src = "$weight * vec2(fma($c00, v.x, fma($c10, v.y, $c20)), fma($c01, v.x, fma($c11, v.y, $c21)))"
        src = replace_macro(src, "weight", _fp(m["variations"]["linear"]))
        return True, _resolve_coefs(src, m["affine"], "c")
    result = ""
    first_var = True
    affine = "\tv.xy = vec2(fma($c00, v.x, fma($c10, v.y, $c20)), fma($c01, v.x, fma($c11, v.y, $c21)));\n"
    for name in x.variations:
        vd = vt.vars[name]
        w = _fp(m["variations"][name])
        var_src = "// variation: " + name + "\n"
        if vd.source:
            var_src += replace_macro(vd.source, "weight", w) + "\n"
        weight_str = "" if "no_weight_mul" in vd.flags else "$weight *"
        if "pre_xform" in vd.flags:
            affine += replace_macro("v.xy += " + weight_str + vd.result + ";", "weight", w) + "\n"
        else:
            var_src += replace_macro(("vec2 result = " if first_var else "result += ") + weight_str + vd.result + ";", "weight", w) + "\n"
            result += var_src
            first_var = False
    for pname in x.var_param:
        result = replace_macro(result, pname, _fp(m["param"][pname]))
    macros = {n for n in find_macros(result) if vt.is_common(n)}
    adj: Dict[str, set] = {}
    for mac in sorted(macros):
        deps = find_macros(vt.common[mac])
        for d in sorted(deps):
            if vt.is_common(d) and d not in adj:
                adj[d] = find_macros(vt.common[mac])
        adj[mac] = find_macros(vt.common[mac])
    order = _get_ordering(adj)
    while order:
        top = order.pop()
        result = "float " + top + " = " + vt.common[top] + ";\n" + result
    result = affine + result
    result = _resolve_coefs(result, m["affine"], "c")
    if x.post is not None:
        result += "\tresult = vec2(fma($p00, result.x, fma($p10, result.y, $p20)), fma($p01, result.x, fma($p11, result.y, $p21)));\n"
        result = _resolve_coefs(result, m["post"], "p")
    result = result.replace("\n", "\n\t").replace("$", "")
    return False, result


def _xform_select(bm: dict) -> str:
    """shaders/templates/xform_select.tpl.glsl (inja: loop.index is 0-based)."""
    lines = ["int get_xform_id(float ratio) {", ""]
    n = len(bm["xforms"])
    for i, xm in enumerate(bm["xforms"]):
        if i == 0:
            lines += ["\tfloat sum = fp[%d];" % xm["weight"], "\tif(sum >= ratio) return 0;"]
        elif i == n - 1:
            lines += ["\treturn %d;" % i]
        else:
            lines += ["\tsum += fp[%d];" % xm["weight"], "\tif(sum >= ratio) return %d;" % i]
    lines += ["}"]
    return "\n".join(lines)


def _cases(f: Flame, vt: VariationTable):
    """[(label, text)] with label in {'case -1', 'case 0', ..., 'default'}"""
    out = []
    for i in range(-1, len(f.xforms)):
        if i == -1 and f.final_xform is None:
            continue
        x = f.final_xform if i == -1 else f.xforms[i]
        m = f.buffer_map["final_xform"] if i == -1 else f.buffer_map["xforms"][i]
        inlined, src = _xform_text(x, m, vt)
        invoke = "return vec4(%s, mix(((first_run)? randf(): v.z), %s, %s), %s);\n" % (
            src if inlined else "result", _fp(m["color"]), _fp(m["color_speed"]), _fp(m["opacity"]))
        if not inlined:
            invoke = src + invoke
        out.append(("default" if i + 1 == len(f.xforms) else "case %d" % i, invoke))
    return out


def compile_flame_xforms(f: Flame, vt: VariationTable) -> str:
    disp = "vec4 dispatch(vec3 v, int xform){\nswitch(xform){\n"
    for label, invoke in _cases(f, vt):
        if label == "default":
            disp += "default: {\n" + invoke + "\n}}\n"
        else:
            disp += label + ": {\n" + invoke + "\n}\n"
    disp = disp.replace("\n", "\n\t") + "\n}"
    return _xform_select(f.buffer_map) + disp


# ----------------------------------------------------------------------------------------
# GLSL -> C++ for the CPU build (the shim provides swizzles, so only two rewrites are needed)
# ----------------------------------------------------------------------------------------
_FLOAT_LIT = re.compile(r"(?<![A-Za-z0-9_.])((?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?)(lf|LF|f|F)?(?![A-Za-z0-9_])")


def _suffix_literals(src: str) -> str:
    def one_line(line: str) -> str:
        code, sep, comment = line.partition("//")

        def repl(m):
            tok = m.group(1)
            if "." in tok or "e" in tok or "E" in tok:
                return tok + "f"
            return m.group(0)
        return _FLOAT_LIT.sub(repl, code) + sep + comment
    return "\n".join(one_line(l) for l in src.split("\n"))


def _sequence_randf(body: str, counter: List[int]) -> str:
    """Statements with >= 2 randf() calls (and no `?`) get their draws hoisted, left to right."""
    out, chunk = [], []
    i = 0
    pieces = re.split(r"([;{}])", body)
    # re-assemble into delimiter-terminated chunks
    chunks = []
    for k in range(0, len(pieces) - 1, 2):
        chunks.append(pieces[k] + pieces[k + 1])
    if len(pieces) % 2:
        chunks.append(pieces[-1])
    for ch in chunks:
        code_lines = [l for l in ch.split("\n") if not l.strip().startswith("//")]
        code = "\n".join(code_lines)
        n = code.count("randf()")
        if n >= 2 and "?" not in code:
            names = []
            def repl(_m):
                names.append("_rf%d" % counter[0]); counter[0] += 1
                return names[-1]
            lines = ch.split("\n")
            new_lines, started = [], False
            for l in lines:
                if l.strip().startswith("//") or (not started and not l.strip()):
                    new_lines.append(l); continue
                if not started:
                    started = True
                    new_lines.append("@@DECL@@")
                new_lines.append(re.sub(r"randf\(\)", repl, l))
            decl = "float " + ", ".join("%s = randf()" % nme for nme in names) + ";"
            ch = "\n".join(new_lines).replace("@@DECL@@", decl)
        out.append(ch)
    return "".join(out)


def generate_cpp(f: Flame, vt: VariationTable) -> str:
    counter = [0]
    disp = "vec4 dispatch(vec3 v, int xform){\nswitch(xform){\n"
    for label, invoke in _cases(f, vt):
        invoke = _sequence_randf(_suffix_literals(invoke), counter)
        if label == "default":
            disp += "default: {\n" + invoke + "\n}}\n"
        else:
            disp += label + ": {\n" + invoke + "\n}\n"
    if not f.xforms:
        disp += "default: { return vec4(v.xy, v.z, 0.0f); }}\n"
    disp += "return vec4(0.0f, 0.0f, 0.0f, 0.0f);\n}"
    sel = _xform_select(f.buffer_map)
    if len(f.buffer_map["xforms"]) <= 1:
        sel = sel[:-1] + "\treturn 0;\n}"
    return "\n".join([
        "// generated by oracle/refrakt_oracle.py — TEST INFRASTRUCTURE",
        '#include "glsl_shim.hpp"',
        "#define ORC_TOTAL_PARAMS %d" % f.buffer_map["size"],
        "#define ORC_HAS_FINAL %d" % (1 if f.final_xform is not None else 0),
        '#include "oracle_core_pre.hpp"',
        "namespace glsl {",
        sel, disp,
        "}  // namespace glsl",
        '#include "oracle_core.hpp"',
        '#include "oracle_api.inc"', ""])


# ----------------------------------------------------------------------------------------
# compiled instance
# ----------------------------------------------------------------------------------------
def _compile(source: str, native: bool) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    deps = "".join(open(os.path.join(HERE, n)).read() for n in ("glsl_shim.hpp", "oracle_core_pre.hpp", "oracle_core.hpp", "oracle_api.inc"))
    tag = hashlib.sha256((source + deps + str(native)).encode()).hexdigest()[:16]
    so = os.path.join(BUILD_DIR, "oracle_%s.so" % tag)
    if not os.path.exists(so):
        cpp = os.path.join(BUILD_DIR, "oracle_%s.cpp" % tag)
        with open(cpp, "w") as fh:
            fh.write(source)
        cmd = ["g++", "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I" + HERE, cpp, "-o", so + ".tmp"]
        if native:
            cmd.insert(3, "-march=native")
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout[-6000:])
        os.replace(so + ".tmp", so)
    return so


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


class Oracle:
    """One compiled CPU instance of the reference path for one genome."""

    def __init__(self, flame: Flame, vt: VariationTable, native: bool = False):
        self.flame = flame
        self.vt = vt
        self.glsl = compile_flame_xforms(flame, vt)
        self.cpp = generate_cpp(flame, vt)
        self.lib = ctypes.CDLL(_compile(self.cpp, native))
        self.lib.orc_draw_to_bins.restype = ctypes.c_ulonglong
        self.total_params = flame.buffer_map["size"]
        maps = list(flame.buffer_map["xforms"]) + ([flame.buffer_map["final_xform"]] if flame.final_xform is not None else [])
        self.slots7 = np.array([m["affine"] + [m["rotation_frequency"]] for m in maps], dtype=np.int32).reshape(-1, 7)

    # --- small functions
    def jsf32_warmup(self, seed: int) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint32)
        self.lib.orc_jsf32_warmup(ctypes.c_uint32(seed), _p(out, ctypes.c_uint32))
        return out

    def device_randf(self, state: np.ndarray, n: int):
        st = np.ascontiguousarray(state, dtype=np.uint32).copy()
        out = np.zeros(n, dtype=np.float32)
        self.lib.orc_device_randf(_p(st, ctypes.c_uint32), n, _p(out, ctypes.c_float))
        return out, st

    def make_sample_points(self, count: int) -> np.ndarray:
        out = np.zeros((count, 4), dtype=np.float32)
        self.lib.orc_make_sample_points(ctypes.c_uint32(count), _p(out, ctypes.c_float))
        return out

    def params(self) -> np.ndarray:
        return copy_flame_data_to_buffer(self.flame)

    def single_step(self, xyz, xid, rng, fp=None, first_run=False):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        xid = np.ascontiguousarray(xid, dtype=np.int32)
        rng = np.ascontiguousarray(rng, dtype=np.uint32).copy()
        fp = np.ascontiguousarray(self.params() if fp is None else fp, dtype=np.float32)
        n = xid.shape[0]
        out = np.zeros((n, 4), dtype=np.float32)
        self.lib.orc_single_step(n, _p(xyz, ctypes.c_float), _p(xid, ctypes.c_int), _p(rng, ctypes.c_uint32), _p(fp, ctypes.c_float), int(first_run), _p(out, ctypes.c_float))
        return out, rng

    def select_xform(self, ratio, fp=None):
        ratio = np.ascontiguousarray(ratio, dtype=np.float32)
        fp = np.ascontiguousarray(self.params() if fp is None else fp, dtype=np.float32)
        out = np.zeros(ratio.shape[0], dtype=np.int32)
        self.lib.orc_select_xform(ratio.shape[0], _p(ratio, ctypes.c_float), _p(fp, ctypes.c_float), _p(out, ctypes.c_int))
        return out

    def bucket_index(self, xyzw, ss_affine, W, H):
        xyzw = np.ascontiguousarray(xyzw, dtype=np.float32)
        ss = np.ascontiguousarray(ss_affine, dtype=np.float32)
        n = xyzw.shape[0]
        idx, pal = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        self.lib.orc_bucket_index(n, _p(xyzw, ctypes.c_float), _p(ss, ctypes.c_float), W, H, _p(idx, ctypes.c_int), _p(pal, ctypes.c_int))
        return idx, pal

    def animate(self, temporal_samples: int, tss_width: float, fp=None) -> np.ndarray:
        fp = np.ascontiguousarray(self.params() if fp is None else fp, dtype=np.float32)
        out = np.zeros((temporal_samples, self.total_params), dtype=np.float32)
        self.lib.orc_animate(_p(fp, ctypes.c_float), _p(out, ctypes.c_float), temporal_samples, ctypes.c_float(tss_width), _p(self.slots7, ctypes.c_int), self.slots7.shape[0])
        return out

    # --- the render path
    def set_sim_parameters(self, total_particles, temporal_samples, shuffle_count, shuffle_seed=0x5EED0000, rng_seed=0, pass_seed=0x5EED0001):
        self.P, self.TS = total_particles, temporal_samples
        self.lib.orc_set_sim_parameters(ctypes.c_size_t(total_particles), ctypes.c_size_t(temporal_samples), ctypes.c_size_t(shuffle_count),
                                        ctypes.c_uint64(shuffle_seed), ctypes.c_uint32(rng_seed), ctypes.c_uint32(pass_seed))

    def rng_states(self, first, count):
        out = np.zeros((count, 4), dtype=np.uint32)
        self.lib.orc_get_rng_states(_p(out, ctypes.c_uint32), ctypes.c_size_t(first), ctypes.c_size_t(count))
        return out

    def shuffle_table(self, shuffle_count):
        """the nshuf x PPT permutation table (binding 4)"""
        out = np.zeros((shuffle_count, self.P // self.TS), dtype=np.uint32)
        self.lib.orc_get_shuffle_table(_p(out, ctypes.c_uint32))
        return out

    def set_shuffle_table(self, tables):
        tables = np.ascontiguousarray(tables, dtype=np.uint32)
        assert tables.size == self.shuffle_table(tables.shape[0]).size
        self.lib.orc_set_shuffle_table(_p(tables, ctypes.c_uint32))

    def force_pass_ids(self, ids):
        """the next host loops take their (in, out) shuffle ids from `ids` (flattened pairs) instead of the seeded generator"""
        ids = np.ascontiguousarray(ids, dtype=np.uint32).reshape(-1)
        self.lib.orc_force_pass_ids(_p(ids, ctypes.c_uint32), ctypes.c_size_t(ids.size))

    def fp_inflated(self):
        out = np.zeros((self.TS, self.total_params), dtype=np.float32)
        self.lib.orc_get_fp_inflated(_p(out, ctypes.c_float))
        return out

    def id_log(self, clear=False):
        """every (shuf_buf_idx_in, shuf_buf_idx_out) pair the host loops drew so far (flame.cpp:261-262, :274-275, :320-321)"""
        self.lib.orc_id_log_size.restype = ctypes.c_size_t
        n = self.lib.orc_id_log_size()
        out = np.zeros(n, dtype=np.uint32)
        if n:
            self.lib.orc_get_id_log(_p(out, ctypes.c_uint32))
        if clear:
            self.lib.orc_clear_id_log()
        return out.reshape(-1, 2)

    def particles(self):
        out = np.zeros((self.P, 4), dtype=np.float32)
        self.lib.orc_get_particles(_p(out, ctypes.c_float))
        return out

    def warmup(self, num_passes: int, tss_width: float):
        fp = self.params()
        pal = np.ascontiguousarray(self.flame.palette, dtype=np.float32)
        self.lib.orc_warmup(_p(fp, ctypes.c_float), _p(pal, ctypes.c_float), _p(self.slots7, ctypes.c_int), self.slots7.shape[0],
                            ctypes.c_size_t(num_passes), ctypes.c_float(tss_width))

    def draw_to_bins(self, bins: np.ndarray, W: int, num_iter: int, count_xforms=False) -> int:
        assert bins.dtype == np.float32 and bins.flags.c_contiguous
        H = bins.size // 4 // W
        ss = screen_space_affine(self.flame, W, H)
        return int(self.lib.orc_draw_to_bins(_p(bins, ctypes.c_float), ctypes.c_size_t(W), ctypes.c_size_t(H), _p(ss, ctypes.c_float), num_iter, int(count_xforms)))

    def draw_accumulate(self, W: int, H: int, num_iter: int) -> int:
        """the passes of draw_to_bins into per-thread histograms that persist until merge_private (bench.py's CPU baseline)"""
        ss = screen_space_affine(self.flame, W, H)
        self.lib.orc_draw_accumulate.restype = ctypes.c_ulonglong
        return int(self.lib.orc_draw_accumulate(ctypes.c_size_t(W), ctypes.c_size_t(H), _p(ss, ctypes.c_float), num_iter))

    def merge_private(self, bins: np.ndarray):
        assert bins.dtype == np.float32 and bins.flags.c_contiguous
        self.lib.orc_merge_private(_p(bins, ctypes.c_float))

    def release_private(self):
        self.lib.orc_release_private()

    def xform_picks(self, n):
        out = np.zeros(n, dtype=np.uint64)
        self.lib.orc_get_xform_picks(_p(out, ctypes.c_ulonglong), n)
        return out

    def density_estimate(self, bins, W, H, radius=None, min_=None, curve=None):
        bins = np.ascontiguousarray(bins, dtype=np.float32)
        out = np.zeros((H, W, 4), dtype=np.float32)
        fl = self.flame
        self.lib.orc_density_estimate(_p(bins, ctypes.c_float), _p(out, ctypes.c_float), W, H,
                                      int(fl.estimator_radius if radius is None else radius), int(fl.estimator_min if min_ is None else min_),
                                      ctypes.c_float(fl.estimator_curve if curve is None else curve))
        return out

    def tonemap(self, image, gamma=None, scale_constant=1e-4, brightness=None, vibrancy=None):
        image = np.ascontiguousarray(image, dtype=np.float32)
        out = np.zeros_like(image)
        fl = self.flame
        self.lib.orc_tonemap(_p(image, ctypes.c_float), _p(out, ctypes.c_float), ctypes.c_size_t(image.size // 4),
                             ctypes.c_float(fl.gamma if gamma is None else gamma), ctypes.c_float(scale_constant),
                             ctypes.c_float(fl.brightness if brightness is None else brightness), ctypes.c_float(fl.vibrancy if vibrancy is None else vibrancy))
        return out

    def max_threads(self):
        return int(self.lib.orc_max_threads())

    def set_threads(self, n):
        self.lib.orc_set_threads(int(n))


def to_rgba8(image: np.ndarray) -> np.ndarray:
    """texture::get_pixels (buffer_objects.hpp:113-119): float -> UNORM8, round to nearest."""
    return np.rint(np.clip(image, 0.0, 1.0) * 255.0).astype(np.uint8)
