#!/usr/bin/env python
"""Builds oracle/_ref/libref_shaders.so: the reference's own GLSL simulation shaders, translated mechanically
(translate.py) from the files where they lie in the reference checkout and compiled for the host against
glsl_shim.h + ref_driver.cpp.  TEST INFRASTRUCTURE.  Outputs only under oracle/_ref/ (git-ignored).

    python oracle/ref_shim/build_ref.py [<reference checkout>]      (default /root/reference)

Returns 0 and does nothing when the checkout is absent (the GPU box): tests that need the library skip there
and use the golden vectors generated from it (tests/golden/make_ref_shader_golden.py) instead."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
LIB = os.path.join(OUT, "libref_shaders.so")
CXX = "/usr/bin/g++"  # the image's $CXX wrapper has no libgomp.spec (see oracle/Makefile)
# -ffp-contract=off: one fp32 rounding per written operation.  -fpermissive: GLSL lets a `case` label jump over a
# declaration with an initialiser (boundaryShader.frag:437, lightingShader.frag:103), C++ calls that ill-formed.
FLAGS = ["-O3", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fpermissive", "-w"]


def build(ref_root: str = "/root/reference", force: bool = False, extra_flags=(), out: str | None = None) -> str | None:
    """Build (when stale) and return the library path; None when there is neither a checkout nor a prebuilt library.
    extra_flags / out: experiment builds beside the canonical library (e.g. -DWSB_REF_LIBM, -ffp-contract=fast;
    profiles/tools/freeze_sensitivity.py)."""
    lib = out or LIB
    shaders = os.path.join(ref_root, "shaders")
    if not os.path.isdir(shaders):
        return lib if os.path.exists(lib) else None
    srcs = [os.path.join(HERE, f) for f in ("translate.py", "glsl_shim.h", "ref_driver.cpp")]
    srcs += [os.path.join(dp, f) for dp, _, fs in os.walk(shaders) for f in fs]
    if not (force or out) and os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(p) for p in srcs):
        return lib
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    gen = tempfile.mkdtemp(prefix="wsb_ref_gen_")  # the translated shader text is an intermediate: only the .so is kept
    try:
        subprocess.check_call([sys.executable, os.path.join(HERE, "translate.py"), shaders, gen], stdout=subprocess.DEVNULL)
        subprocess.check_call([CXX] + FLAGS + list(extra_flags) + ["-I", HERE, "-I", gen, "-o", lib, os.path.join(HERE, "ref_driver.cpp")])
    finally:
        if os.environ.get("WSB_KEEP_REF_GEN"):
            print("generated headers kept in", gen)
        else:
            shutil.rmtree(gen, ignore_errors=True)
    return lib


if __name__ == "__main__":
    lib = build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", force=True)
    print(lib or "reference checkout absent: nothing built")
