#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — turns one GLSL 4.60 shader of the reference into a shared library for the soft GL.

    glsl_to_cpp.py <shader.glsl> <stage: compute|vertex|fragment> <out.so>

The input is the exact string the reference hands to glShaderSource (oracle/softgl/softgl.cpp writes it to disk), after
the reference's own `$include_` / `$varsource` / `$block_width` / `$final_xform_call` substitution. Two things happen:

1. inja templates. The stand-in for inja (oracle/stubs/inja/inja.hpp) does not render; it returns the template and the
   JSON data the reference passed, between sentinels. `render_inja` evaluates the subset of inja 3.1 the two templates
   (xform_select.tpl.glsl, animate.tpl.glsl) use: {{ dotted.path }}, {% for x in list %}, {% for k, v in object %}
   (objects iterate in key order, as nlohmann::json stores them), loop.index (0-based) / loop.is_first / loop.is_last,
   {% if %} / {% else if %} / {% else %}, exists("name"), existsIn(obj, "key").

2. transliteration. GLSL is close enough to C++ that the shader BODY compiles unchanged against oracle/glsl_shim.hpp
   (vector types with swizzles, built-ins). Only declarations are rewritten, mechanically:
     layout(local_size_*) in;              -> RFK_LOCAL_X/Y/Z
     layout(std430, binding=N) buffer ..   -> a pointer the soft GL sets from glBindBufferBase
     uniform T name [= default];           -> a static the soft GL sets from glUniform* (bool is stored as int)
     shared T name;                        -> a static (work groups run one at a time)
     T name;  /  [flat] in|out T name;     -> a member of the per-invocation struct rfk_private
     void main()                           -> void rfk_shader_main()
     discard                               -> return rfk_discard()
     1.0 / .5 / 1e10                       -> float literals get their `f` (GLSL literals are float, C++ ones double)
     several randf() in one statement      -> hoisted left to right (GLSL leaves the order open; same rule as the oracle)
   randd() (random.glsl, unused, double arithmetic) is dropped.

No reference text is stored in the repository: inputs are read where they lie, outputs go to oracle/_ref/.
"""
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

INJA_OPEN, INJA_MID, INJA_CLOSE = "\x01INJA\x02", "\x02DATA\x02", "\x03"


# ------------------------------------------------------------------------------------------------ inja subset
def _lookup(path, scopes):
    parts = path.strip().split(".")
    for scope in reversed(scopes):
        if parts[0] in scope:
            cur = scope[parts[0]]
            break
    else:
        raise KeyError(path)
    for p in parts[1:]:
        if isinstance(cur, list):
            cur = cur[int(p)]
        else:
            cur = cur[p]
    return cur


def _exists(path, scopes):
    try:
        _lookup(path, scopes)
        return True
    except (KeyError, IndexError, ValueError):
        return False


def _eval(expr, scopes):
    expr = expr.strip()
    m = re.fullmatch(r'existsIn\(\s*([\w.]+)\s*,\s*"([^"]*)"\s*\)', expr)
    if m:
        obj = _lookup(m.group(1), scopes)
        return isinstance(obj, dict) and m.group(2) in obj
    m = re.fullmatch(r'exists\(\s*"([^"]*)"\s*\)', expr)
    if m:
        return _exists(m.group(1), scopes)
    return _lookup(expr, scopes)


def _fmt(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float) and v == int(v):
        return str(int(v))
    return str(v)


def render_inja(template, data):
    tokens = re.split(r"(\{\{.*?\}\}|\{%.*?%\})", template, flags=re.S)
    pos = 0

    def parse(stop):
        nonlocal pos
        nodes = []
        while pos < len(tokens):
            t = tokens[pos]
            if t.startswith("{%"):
                stmt = t[2:-2].strip()
                if any(stmt == s or stmt.startswith(s + " ") for s in stop):
                    return nodes, stmt
                pos += 1
                if stmt.startswith("for "):
                    m = re.fullmatch(r"for\s+(\w+)\s*(?:,\s*(\w+))?\s+in\s+([\w.]+)", stmt)
                    body, _ = parse(("endfor",))
                    pos += 1
                    nodes.append(("for", m.group(1), m.group(2), m.group(3), body))
                elif stmt.startswith("if "):
                    branches, cond = [], stmt[3:]
                    while True:
                        body, end = parse(("else if", "else", "endif"))
                        pos += 1
                        branches.append((cond, body))
                        if end == "endif":
                            break
                        cond = end[len("else if"):].strip() if end.startswith("else if") else None
                    nodes.append(("if", branches))
                else:
                    raise ValueError("inja subset: unsupported statement " + stmt)
            elif t.startswith("{{"):
                nodes.append(("expr", t[2:-2]))
                pos += 1
            else:
                nodes.append(("text", t))
                pos += 1
        return nodes, None

    tree, _ = parse(())

    def run(nodes, scopes, out):
        for n in nodes:
            if n[0] == "text":
                out.append(n[1])
            elif n[0] == "expr":
                out.append(_fmt(_eval(n[1], scopes)))
            elif n[0] == "if":
                for cond, body in n[1]:
                    if cond is None or _eval(cond, scopes):
                        run(body, scopes, out)
                        break
            else:
                _, a, b, src, body = n
                seq = _lookup(src, scopes)
                items = [(k, seq[k]) for k in sorted(seq)] if isinstance(seq, dict) else list(enumerate(seq))
                for i, (k, v) in enumerate(items):
                    scope = {"loop": {"index": i, "index1": i + 1, "is_first": i == 0, "is_last": i == len(items) - 1}}
                    if b is None:
                        scope[a] = v
                    else:
                        scope[a], scope[b] = k, v
                    run(body, scopes + [scope], out)

    out = []
    run(tree, [data], out)
    return "".join(out)


def expand_inja(text):
    def repl(m):
        return render_inja(m.group(2), json.loads(m.group(1)))
    return re.sub(re.escape(INJA_OPEN) + "(.*?)" + re.escape(INJA_MID) + "(.*?)" + re.escape(INJA_CLOSE), repl, text, flags=re.S)


# ------------------------------------------------------------------------------------------------ transliteration
_CTYPE = {"bool": "int"}
_SIZES = {"bool": 4, "int": 4, "uint": 4, "float": 4, "uvec2": 8, "ivec2": 8, "mat4": 64, "image2D": 24}
_VALUE_TYPES = r"(?:float|int|uint|bool|vec2|vec3|vec4|uvec2|uvec3|uvec4|ivec2|mat4)"


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def transliterate(src, stage):
    from refrakt_oracle import _suffix_literals, _sequence_randf  # the oracle's two documented text rules

    src = expand_inja(src).replace("\r", "")
    src = _strip_comments(src)
    src = re.sub(r"^\s*#\s*(version|line)[^\n]*$", "", src, flags=re.M)
    src = re.sub(r"double\s+randd\s*\(\s*\)\s*\{[^}]*\}", "", src)

    local = [1, 1, 1]
    m = re.search(r"layout\s*\(\s*local_size_x\s*=\s*(\d+)\s*,\s*local_size_y\s*=\s*(\d+)\s*,\s*local_size_z\s*=\s*(\d+)\s*\)\s*in\s*;", src)
    if m:
        local = [int(m.group(k)) for k in (1, 2, 3)]
        src = src[:m.start()] + src[m.end():]

    bindings, uniforms, privates, varyings = [], [], [], []

    def buffer_block(m):
        binding, ctype, name = int(m.group(1)), m.group(2), m.group(3)
        bindings.append((binding, name))
        return "static %s* %s;" % (ctype, name)
    src = re.sub(r"layout\s*\(\s*std430\s*,\s*binding\s*=\s*(\d+)\s*\)\s*buffer\s+\w+\s*\{\s*(\w+)\s+(\w+)\s*\[\s*\]\s*;\s*\}\s*;", buffer_block, src)

    def image_uniform(m):
        uniforms.append((m.group(1), "image2D", 1))
        return "static image2D %s;" % m.group(1)
    src = re.sub(r"layout\s*\(\s*rgba32f\s*\)\s*uniform\s+image2D\s+(\w+)\s*;", image_uniform, src)

    def uniform(m):
        gtype, name, count, default = m.group(1), m.group(2), m.group(3), m.group(4)
        uniforms.append((name, gtype, int(count) if count else 1))
        decl = "static %s %s%s" % (_CTYPE.get(gtype, gtype), name, "[%s]" % count if count else "")
        return decl + (" = %s;" % default.strip() if default else (" = {};" if count or gtype in ("uvec2", "ivec2", "mat4") else " = 0;"))
    src = re.sub(r"\buniform\s+(\w+)\s+(\w+)\s*(?:\[\s*(\d+)\s*\])?\s*(?:=\s*([^;]+))?;", uniform, src)

    src = re.sub(r"\bshared\s+(\w+\s+\w+\s*(?:\[\s*\d+\s*\])?\s*;)", r"static \1", src)

    def varying(m):
        direction, gtype, name = m.group(1), m.group(2), m.group(3)
        privates.append((gtype, name))
        varyings.append((name, gtype, 1 if direction == "out" else 0))
        return ""
    src = re.sub(r"^\s*(?:flat\s+)?(in|out)\s+(" + _VALUE_TYPES + r")\s+(\w+)\s*;", varying, src, flags=re.M)

    # unqualified file-scope variables are per-invocation in GLSL
    depth, out_lines = 0, []
    for line in src.split("\n"):
        m = re.match(r"^\s*(" + _VALUE_TYPES + r")\s+(\w+)\s*;\s*$", line) if depth == 0 else None
        if m:
            privates.append((m.group(1), m.group(2)))
            line = ""
        depth += line.count("{") - line.count("}")
        out_lines.append(line)
    src = "\n".join(out_lines)

    src = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void rfk_shader_main()", src)
    src = re.sub(r"\bdiscard\b", "return rfk_discard()", src)
    src = re.sub(r"(\d)lf\b", r"\1", src)
    src = _sequence_randf(_suffix_literals(src), [0])

    head = ["// generated by oracle/softgl/glsl_to_cpp.py from a shader of the reference — TEST INFRASTRUCTURE, never committed",
            "#define RFK_LOCAL_X %d" % local[0], "#define RFK_LOCAL_Y %d" % local[1], "#define RFK_LOCAL_Z %d" % local[2],
            "#define RFK_STAGE %d" % {"compute": 0, "vertex": 1, "fragment": 2}[stage],
            "#define RFK_USES_BARRIER %d" % (1 if re.search(r"\bbarrier\s*\(", src) else 0),
            '#include "shader_rt_pre.hpp"', "#include <new>", "namespace glsl {",
            "struct rfk_private { " + " ".join("%s %s;" % (_CTYPE.get(t, t), n) for t, n in privates) + " };"]
    head += ["#define %s (static_cast<rfk_private*>(rfk_cur->priv)->%s)" % (n, n) for _, n in privates]
    tail = ["static rfk_binding rfk_bindings[] = {" + "".join("{%d, (void**)&%s}, " % b for b in bindings) + "{-1, nullptr}};",
            "static rfk_uniform rfk_uniforms[] = {" + "".join('{"%s", (void*)&%s, sizeof(%s)}, ' % (n, n, n) for n, _, _ in uniforms) + "{nullptr, nullptr, 0}};"]
    for _, n in privates:
        tail.append("#undef " + n)
    tail += ["struct rfk_varying { const char* name; size_t offset, bytes; int is_out; };",
             "static rfk_varying rfk_varyings[] = {" + "".join('{"%s", offsetof(rfk_private, %s), sizeof(rfk_private::%s), %d}, ' % (n, n, n, o) for n, _, o in varyings) + "{nullptr, 0, 0, 0}};",
             "}  // namespace glsl", '#include "shader_rt_post.hpp"', ""]
    return "\n".join(head) + "\n" + src + "\n" + "\n".join(tail)


def build(glsl_path, stage, so_path):
    cpp = transliterate(open(glsl_path, newline="").read(), stage)
    cpp_path = os.path.splitext(so_path)[0] + ".cpp"
    with open(cpp_path, "w") as fh:
        fh.write(cpp)
    cmd = ["g++", "-std=c++17", "-O2", "-w", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I" + HERE, cpp_path, "-o", so_path + ".tmp"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-8000:])
        return 1
    os.replace(so_path + ".tmp", so_path)
    return 0


if __name__ == "__main__":
    sys.exit(build(sys.argv[1], sys.argv[2], sys.argv[3]))
