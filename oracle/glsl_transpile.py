#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY. Turns the reference's compute shaders into C++ translation units WITHOUT copying them
into the repository: the shader text is read where it lies ($REF/shaders/*.glsl), rewritten by the textual rules
below and written to oracle/_ref/ (git-ignored), where oracle/Makefile compiles it against oracle/glsl_shim.inc and
oracle/glsl_driver.inc into oracle/_ref/libadypt_glsl.so. The result executes the REFERENCE'S OWN traversal, primary-
ray and path-tracing code on the CPU; tests/test_glsl_reference.py holds the oracle (and through the golden fixtures
the CUDA kernels) to it bit for bit.

Rewrites (every rule asserts how often it fired, so a different shader version fails loudly):
  syntax   S1 buffer / image / sampler / uniform-block declarations -> plain globals set by the driver
           S2 `in const T` / `in T` -> by value, `out T` / `inout T` -> T&
           S3 swizzles: `.xyz` -> `.xyz()`, `v.zw = e;` -> `v.set_zw(e);`
           S4 unsuffixed floating literals get an `f` (GLSL has no double promotion)
           S5 declarations at case level of a switch are hoisted in front of it (C++ forbids jumping over them)
           S6 `main` -> `shader_main`; kPixel and the traversal stack become thread_local
  FP policy (the choices GLSL leaves to the driver, made the same way as in the oracle and in CUDA)
           P1 slab tests  `float(q) * adjusted_idir_c + adjusted_origin.c` -> fma(float(q), adjusted_idir_c, adjusted_origin.c)
           P2 Woop test   `tox + tt*tdx` -> fma(tt, tdx, tox) (same for y);  dot() inside traversal.glsl -> dot_fma()
           P3 min / max of the slab tests -> slab_min / slab_max (IEEE minNum / maxNum)
"""
import os
import re
import sys


def sub(pattern, repl, text, expect=None, flags=0, what=""):
    out, n = re.subn(pattern, repl, text, flags=flags)
    if expect is not None and n != expect:
        raise SystemExit(f"glsl_transpile: rule {what or pattern!r} fired {n} times, expected {expect}")
    return out


def common(text, counts):
    c = counts
    # S1
    text = sub(r"layout\(std430, binding = \d+\) (?:readonly )?buffer \w+ \{ (\w+) (\w+)\[\]; \};", r"static const \1 *\2;", text, c["ssbo"], what="S1 ssbo")
    text = sub(r"layout\((rgba32f|rg8), binding = \d+\) uniform image2D (\w+);", r"static image2D \2;", text, c["image"], what="S1 image")
    text = sub(r"layout\(binding = 3\) uniform uuTextures \{ sampler2D uTextures\[TEXTURE_COUNT\]; \};", "static const sampler2D *uTextures;", text, c["tex"], what="S1 textures")
    text = sub(r"layout\(std140, binding = \d+\) uniform \w+\s*\{([^}]*)\};", lambda m: re.sub(r"(?m)^(\s*)(vec4|mat4|int|float) ", r"\1static \2 ", m.group(1)), text, c["ubo"], what="S1 ubo")
    # S2
    text = sub(r"\bin const ", "const ", text, what="S2 in const")
    text = sub(r"\binout (\w+) ", r"\1 &", text, c["inout"], what="S2 inout")
    text = sub(r"\bout (\w+) ", r"\1 &", text, c["out"], what="S2 out")
    # S6
    text = sub(r"const ivec2 kPixel = ivec2\(gl_GlobalInvocationID\.xy\);", "static thread_local ivec2 kPixel;", text, c["kpixel"], what="S6 kPixel")
    text = sub(r"\bvoid main\(\)", "void shader_main()", text, c["main"], what="S6 main")
    text = sub(r"(?m)^uvec2 stack\[TRAVERSAL_STACK_SIZE\];", "static thread_local uvec2 stack[TRAVERSAL_STACK_SIZE];", text, c["stack"], what="S6 stack")
    # S3
    text = sub(r"(\w+)\.(xyz|zw)\s*=(?!=)\s*([^;]+);", r"\1.set_\2(\3);", text, c["swz_set"], what="S3 swizzle store")
    text = sub(r"\.(xyz|xy|zw|rgb)\b(?!\()", r".\1()", text, what="S3 swizzle load")
    # S4
    text = sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?)(?![\w.])", r"\1f", text, what="S4 literals")
    return text


def traversal(text):
    text = common(text, dict(ssbo=3, image=0, tex=0, ubo=0, inout=2, out=0, kpixel=0, main=0, stack=1, swz_set=0))
    # P1: 48 slab evaluations in each of the two BVHIntersection overloads
    text = sub(r"= float\((.*?& 0xffu)\) \* (adjusted_idir_[xyz]) \+ (adjusted_origin\.[xyz]);", r"= fma(float(\1), \2, \3);", text, 96, what="P1 slab fma")
    # P2
    text = sub(r"\btu = tox \+ tt\*tdx;", "tu = fma(tt, tdx, tox);", text, 2, what="P2 tu")
    text = sub(r"\btv = toy \+ tt\*tdy;", "tv = fma(tt, tdy, toy);", text, 2, what="P2 tv")
    text = sub(r"\bdot\(", "dot_fma(", text, 12, what="P2 dot")
    # P3
    text = sub(r"\bmax\(", "slab_max(", text, 48, what="P3 max")
    text = sub(r"\bmin\(", "slab_min(", text, 48, what="P3 min")
    return text


def hoist_switch_decls(text, expect):
    """S5: `float x = e;` / `float x;` at case-body level (4 tabs) inside `switch(...) {` (2 tabs) -> hoisted."""
    lines = text.split("\n")
    out, names, sw_at, depth_in = [], [], None, False
    for ln in lines:
        if re.match(r"^\t\tswitch\(", ln):
            sw_at = len(out)
            depth_in = True
        elif depth_in and re.match(r"^\t\t\}", ln):
            depth_in = False
        elif depth_in:
            m = re.match(r"^(\t\t\t\t)float (\w+)(\s*=\s*.*)?;\s*$", ln)
            if m:
                names.append(m.group(2))
                ln = f"{m.group(1)}{m.group(2)}{m.group(3)};" if m.group(3) else ""
        out.append(ln)
    if len(names) != expect or sw_at is None:
        raise SystemExit(f"glsl_transpile: S5 hoisted {names}, expected {expect} declarations")
    out.insert(sw_at, "\t\tfloat " + ", ".join(names) + ";")
    return "\n".join(out)


def pathtracer(text):
    text = common(text, dict(ssbo=3, image=3, tex=1, ubo=2, inout=1, out=5, kpixel=1, main=1, stack=0, swz_set=2))
    return hoist_switch_decls(text, 5)


def primaryray(text):
    return common(text, dict(ssbo=2, image=1, tex=1, ubo=2, inout=0, out=0, kpixel=1, main=1, stack=0, swz_set=0))


HEAD = """// GENERATED by oracle/glsl_transpile.py from {ref}/shaders -- do not commit (oracle/_ref/ is git-ignored).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include "../detmath.h"
#define TRAVERSAL_STACK_SIZE 64
#define TEXTURE_COUNT 1 /* the texture branch is compiled in; it is taken only where m_dtex != -1 */
#define IMG_SIZE g_img_size
namespace {ns} {{
#include "../glsl_shim.inc"
static ivec2 g_img_size;
// ---------------------------------------------------------------- {ref}/shaders/traversal.glsl
{traversal}
// ---------------------------------------------------------------- {ref}/shaders/{name}.glsl
{body}
#define GLSL_DRIVER_{NAME}
#include "../glsl_driver.inc"
}} // namespace {ns}
"""


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    outdir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    os.makedirs(outdir, exist_ok=True)
    rd = lambda n: open(os.path.join(ref, "shaders", n + ".glsl")).read()
    trav = traversal(rd("traversal"))
    for name, fn in (("pathtracer", pathtracer), ("primaryray", primaryray)):
        src = HEAD.format(ref=ref, ns="glsl_" + name, traversal=trav, name=name, body=fn(rd(name)), NAME=name.upper())
        with open(os.path.join(outdir, f"glsl_{name}.gen.cpp"), "w") as f:
            f.write(src)
    print(f"[glsl_transpile] wrote {outdir}/glsl_pathtracer.gen.cpp and glsl_primaryray.gen.cpp")


if __name__ == "__main__":
    main()
