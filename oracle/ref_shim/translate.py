#!/usr/bin/env python
"""GLSL ES 3.00 -> C++ for the reference's simulation shaders.  TEST INFRASTRUCTURE (oracle/_ref build).

The shader text is read from the reference checkout where it lies (never copied into this repo) and rewritten
mechanically — no shader logic is restated here:

  * `#include "common.glsl"` is expanded the way the reference's loader does (app.js:6690-6692);
  * `#version`, `precision` lines dropped; comments stripped;
  * every floating literal gets an `f` suffix (GLSL literals are fp32; C++ would promote to double);
  * global `uniform` declarations become static members of the generated struct (they persist between
    invocations), every other global (`in`, `out`, plain) a member that is initialised anew for every
    invocation — zero unless the shader gives an initialiser ("out varyings start at zero");
  * scalar locals declared without an initialiser start at zero (lightingShader.frag:90 IR_up: undefined in
    GLSL where no case assigns it; canonical 0);
  * the bound of the per-row uniform tables (`vec4 ...[126]`, 504 rows) is raised to WSB_REF_PROFILE_VEC4S (grids taller
    than the reference's own 503-row cap: every BASELINE grid); indexing is untouched;
  * every `case` of a `switch` gets its own block (C++ forbids jumping over initialised declarations);
  * the right operand of integer `%` is wrapped in glsl_nz() (`% 0` is undefined in GLSL; canonical "non-zero");
  * the whole shader becomes `struct Shader` in namespace glsl::ref_<name>, with `main()` as a member, plus
    generated glue: set_varyings(vs) copies every `in` from the vertex shader's `out` of the same name,
    bind_uniforms(bag) copies every non-sampler uniform from a bag by NAME (as gl.getUniformLocation does).

usage: translate.py <reference shaders dir> <out dir>
"""
import hashlib
import os
import re
import sys

SHADERS = [  # (file under shaders/, struct namespace)
    ("vertex/simShader.vert", "simVert"),
    ("fragment/velocityShader.frag", "velocity"),
    ("fragment/curlShader.frag", "curl"),
    ("fragment/vorticityShader.frag", "vorticity"),
    ("fragment/boundaryShader.frag", "boundary"),
    ("fragment/advectionShader.frag", "advection"),
    ("fragment/pressureShader.frag", "pressure"),
    ("fragment/lightingShader.frag", "lighting"),
    ("vertex/precipitationShader.vert", "precipVert"),
    ("fragment/precipitationShader.frag", "precipFrag"),
    ("fragment/lightningLocationShader.frag", "lightningLocation"),
    ("fragment/setupShader.frag", "setup"),
]

SAMPLERS = ("sampler2D", "isampler2D")
FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def strip_comments(src: str) -> str:
    out, i, n = [], 0, len(src)
    while i < n:
        if src.startswith("//", i):
            while i < n and src[i] != "\n":
                i += 1
        elif src.startswith("/*", i):
            j = src.find("*/", i + 2)
            j = n if j < 0 else j + 2
            out.append("\n" * src.count("\n", i, j))
            i = j
        else:
            out.append(src[i])
            i += 1
    return "".join(out)


def wrap_mod_operands(line: str) -> str:
    """a % b  ->  a % glsl_nz(b) ; b is a number, an identifier or a parenthesised expression."""
    out, i = [], 0
    while True:
        j = line.find("%", i)
        if j < 0:
            out.append(line[i:])
            return "".join(out)
        out.append(line[i:j + 1])
        k = j + 1
        while k < len(line) and line[k] == " ":
            k += 1
        if k < len(line) and line[k] == "(":
            depth, e = 0, k
            while e < len(line):
                depth += line[e] == "("
                depth -= line[e] == ")"
                e += 1
                if depth == 0:
                    break
        else:
            m = re.match(r"\w+", line[k:])
            if not m:
                raise SystemExit("translate.py: cannot parse the right operand of %% in: %s" % line)
            e = k + m.end()
        out.append(" glsl_nz(" + line[k:e] + ")")
        i = e


def zero_init_scalars(line: str) -> str:
    m = re.match(r"^(\s*)(const\s+)?(float|int|bool|uint)\s+([^;()]*);\s*$", line)
    if not m or m.group(2):
        return line
    decls = [d.strip() for d in m.group(4).split(",")]
    if not all(re.match(r"^\w+(\s*=.*)?$", d) for d in decls):
        return line
    zero = {"float": "0.0f", "int": "0", "uint": "0u", "bool": "false"}[m.group(3)]
    decls = [d if "=" in d else d + " = " + zero for d in decls]
    return "%s%s %s;" % (m.group(1), m.group(3), ", ".join(decls))


def translate(shader_dir: str, rel: str, ns: str) -> str:
    path = os.path.join(shader_dir, rel)
    raw = open(path).read()
    common = open(os.path.join(shader_dir, "common.glsl")).read()
    sha = hashlib.sha1((raw + common).encode()).hexdigest()
    src = raw.replace('#include "common.glsl"', common)  # app.js:6690-6692
    src = strip_comments(src)
    src = FLOAT_LIT.sub(lambda m: m.group(1) + "f", src)
    is_vertex = rel.endswith(".vert")

    body, ins, outs, uniforms = [], [], [], []
    depth = 0
    switches = []  # [body depth, a case block is open]
    for line in src.split("\n"):
        s = line.strip()
        # GLSL lets a `case` label jump over a declaration with an initialiser (boundaryShader.frag:437-476,
        # lightingShader.frag:108-113); C++ does not: give every case its own block.  Fall-through is unaffected.
        if switches and depth == switches[-1][0]:
            if re.match(r"^(case\b[^:]*|default\s*):$", s):
                line = ("} " if switches[-1][1] else "") + s + " {"
                switches[-1][1] = True
            elif s.startswith("}") and switches[-1][1]:
                line = "} " + line
                switches[-1][1] = False
        if depth == 0:
            if s.startswith("#version") or s.startswith("precision"):
                continue
            m = re.match(r"^(?:layout\s*\([^)]*\)\s*)?(uniform|in|out)\s+(\w+)\s+(\w+)\s*(\[\s*\d+\s*\])?\s*;$", s)
            if m:
                q, ty, name, arr = m.group(1), m.group(2), m.group(3), m.group(4) or ""
                if q == "uniform":
                    if ty == "vec4" and arr.replace(" ", "") == "[126]":
                        # the per-row profile tables (initial_Tv, realWorldSounding_*v): 504 rows in the reference, which
                        # caps its grids at 503 rows; the BASELINE grids are taller, so ONLY the array bound is raised
                        arr = "[WSB_REF_PROFILE_VEC4S]"
                    uniforms.append((ty, name, arr))
                    body.append("  static inline %s %s%s;" % (ty, name, arr))
                else:
                    (ins if q == "in" else outs).append((ty, name))
                    init = " = 0.0f" if ty == "float" else ""
                    body.append("  %s %s%s;" % (ty, name, init))
                continue
        if "%" in s and not s.startswith("#"):
            line = wrap_mod_operands(line)
        line = zero_init_scalars(line)
        body.append(line)
        depth += s.count("{") - s.count("}")
        if re.search(r"\bswitch\s*\(.*\)\s*\{$", s):
            switches.append([depth, False])
        elif switches and depth < switches[-1][0]:
            switches.pop()
    if depth != 0:
        raise SystemExit("translate.py: unbalanced braces in " + rel)

    glue = ["  // ---- generated glue ----"]
    glue.append("  template <class VS> void set_varyings(const VS& vs) {" + " ".join("%s = vs.%s;" % (n, n) for _, n in ins) + " (void)vs; }")
    binds = []
    for ty, name, arr in uniforms:
        if ty in SAMPLERS:
            continue
        if arr:
            cnt = arr.strip("[] ")
            binds.append("for (int i = 0; i < %s; i++) %s[i] = u.%s[i];" % (cnt, name, name))
        else:
            binds.append("%s = u.%s;" % (name, name))
    glue.append("  template <class Bag> static void bind_uniforms(const Bag& u) { " + " ".join(binds) + " (void)u; }")
    head = [
        "// GENERATED by oracle/ref_shim/translate.py from the reference checkout — build artefact, never committed.",
        "// source: shaders/%s (+ common.glsl), sha1 %s" % (rel, sha),
        "#pragma once",
        '#include "glsl_shim.h"',
        "namespace glsl { namespace ref_%s {" % ns,
        "struct Shader {",
        "  bool glsl_discarded = false;",
    ]
    if is_vertex:
        head += ["  vec4 gl_Position;", "  float gl_PointSize = 0.0f;"]
    return "\n".join(head + body + glue + ["};", "} }", ""])


def main():
    shader_dir, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    for rel, ns in SHADERS:
        with open(os.path.join(out_dir, "gen_%s.h" % ns), "w") as f:
            f.write(translate(shader_dir, rel, ns))
    print("translated %d shaders into %s" % (len(SHADERS), out_dir))


if __name__ == "__main__":
    main()
