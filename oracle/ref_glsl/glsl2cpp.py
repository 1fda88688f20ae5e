#!/usr/bin/env python
"""glsl2cpp.py -- TEST INFRASTRUCTURE.  Reads one of the reference's shader files WHERE IT LIES (with its #includes, resolved against the
shader directory like the reference's own loader does) and writes the same text to stdout with the handful of token-level changes that make
it C++ for oracle/ref_glsl/glsl_shim.hpp -- nothing is reordered, no expression is touched:
  * `#version`, `#extension` and `layout (local_size...) in;` lines are dropped;
  * `layout(...)`, `uniform`, `restrict`, `readonly`, `volatile`, `coherent` qualifiers are dropped; the braces of uniform / buffer BLOCKS are
    dropped so that their members become globals (an unsized array `T name[];` becomes `T* name;`);
  * parameter qualifiers: `out T x` / `inout T x` -> `T& x`, `in T x` -> `T x`;
  * a mutable variable declared at file scope (private to an invocation in GLSL) becomes thread_local;
  * floating-point literals get the `f` suffix (a GLSL literal is a float, a C++ literal a double);
  * two RNG calls inside one vec2 constructor are put in a braced list (GLSL orders argument evaluation left to right, C++ does not);
  * `void main()` -> `void shader_main()`; `gl_GlobalInvocationID.xy` / `.x` -> the shim's per-invocation variables;
  * `ivec2(gl_GlobalInvocationID.xy)` style conversions of unsigned vectors go through to_ivec2().
The output is piped straight into g++ by oracle/Makefile; no copy of the reference's source is written anywhere.
usage: glsl2cpp.py <shader file> [-I<dir> ...] [-DNAME[=VALUE] ...]"""
import os
import re
import sys

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])(?!f)")
QUALIFIERS = re.compile(r"\b(?:uniform|restrict|readonly|writeonly|volatile|coherent)\b\s*")
LAYOUT = re.compile(r"\blayout\s*\([^)]*\)\s*")
OUT_PARAM = re.compile(r"\b(?:inout|out)\s+(\w+)\s+(\w+)")
IN_PARAM = re.compile(r"\bin\s+(?=(?:const\s+)?\w+\s+\w+\s*(?:[,)]|$))")
BLOCK_HEAD = re.compile(r"^\s*(?:uniform|buffer)\s+\w+\s*$")
PRIVATE_GLOBAL = re.compile(r"^\s*(?:float|int|uint|bool|vec[234]|ivec[234]|uvec[234])\s+\w+\s*(?:\[[^\]]*\])?\s*(?:=[^;]*)?;")
STATE = {"depth": 0}
UNSIZED = re.compile(r"^(\s*)(\w+)\s+(\w+)\s*\[\s*\]\s*;")


INCLUDE_DIRS = []


def resolve(base, name):
    for d in [base] + INCLUDE_DIRS:
        if os.path.exists(os.path.join(d, name)):
            return os.path.join(d, name)
    raise SystemExit("glsl2cpp: cannot find %s" % name)


def convert(path, out, seen_depth=0):
    base = os.path.dirname(path)
    in_block = 0      # 1: block header seen, waiting for '{'; 2: inside the block
    in_comment = False
    out.append('#line 1 "%s"' % path)
    for no, raw in enumerate(open(path, encoding="latin-1").read().split("\n"), 1):
        line = raw.rstrip("\r")
        s = line.strip()
        m = re.match(r'#\s*include\s+"([^"]+)"', s)
        if m:
            convert(resolve(base, m.group(1)), out, seen_depth + 1)
            out.append('#line %d "%s"' % (no + 1, path))
            continue
        # (block comments never contain code we need: nothing inside them is transformed)
        if in_comment:
            if "*/" in line:
                in_comment = False
            out.append(line)
            continue
        if s.startswith("#version") or s.startswith("#extension"):
            out.append("")
            continue
        if s.startswith("#"):
            out.append(FLOAT_LIT.sub(r"\1f", line) if s.startswith("#define") else line)
            continue
        # a block comment that opens on this line (before any // ) and does not close on it
        i_block, i_line = line.find("/*"), line.find("//")
        if i_block >= 0 and (i_line < 0 or i_block < i_line) and "*/" not in line[i_block + 2:]:
            in_comment = True
            out.append(line)
            continue
        code, sep, comment = line.partition("//")
        if re.search(r"\blayout\s*\(\s*local_size", code):
            out.append("")
            continue
        code = LAYOUT.sub("", code)
        had_block_kw = re.search(r"\b(uniform|buffer)\s+\w+\s*$", code) is not None and "(" not in code and ";" not in code
        if had_block_kw:
            in_block = 1
            out.append("// " + line)
            continue
        code = QUALIFIERS.sub("", code)
        code = re.sub(r"\bbuffer\s+(?=\w+\s*$)", "", code)
        if in_block == 1 and code.strip().startswith("{"):
            in_block = 2
            code = code.replace("{", "", 1)
        elif in_block == 2 and code.strip().startswith("}"):
            in_block = 0
            code = re.sub(r"\}\s*\w*\s*;", "", code, count=1)
        if in_block == 2:
            code = UNSIZED.sub(r"\1\2* \3;", code)
        # a mutable variable at file scope is private to the invocation in GLSL: one per thread here
        if STATE["depth"] == 0 and in_block == 0 and PRIVATE_GLOBAL.match(code):
            code = "thread_local " + code.lstrip()
        STATE["depth"] += code.count("{") - code.count("}")
        code = OUT_PARAM.sub(r"\1& \2", code)
        code = IN_PARAM.sub("", code)
        code = FLOAT_LIT.sub(r"\1f", code)
        # GLSL evaluates function arguments left to right; C++ leaves the order open (g++: right to left).  The only arguments with side effects
        # in the reference's shaders are calls that advance the RNG: two of them inside one constructor go through a braced list (ordered in C++)
        code = re.sub(r"\bvec2\(\s*(Random\w*\([^()]*\))\s*,\s*(Random\w*\([^()]*\))\s*\)", r"vec2(glsl_ordered2{\1, \2})", code)
        code = re.sub(r"\bvoid\s+main\s*\(", "void shader_main(", code)
        code = code.replace("ivec2(gl_GlobalInvocationID.xy)", "to_ivec2(gl_GlobalInvocationID_xy)")
        code = code.replace("gl_GlobalInvocationID.xy", "gl_GlobalInvocationID_xy").replace("gl_GlobalInvocationID.x", "gl_GlobalInvocationID_x")
        out.append(code + sep + comment)


def main():
    args = sys.argv[1:]
    path = args[0]
    out = ['#include "glsl_shim.hpp"']
    for a in args[1:]:
        if a.startswith("-I"):
            INCLUDE_DIRS.append(a[2:])
        elif a.startswith("-D"):
            name, _, val = a[2:].partition("=")
            out.append("#define %s %s" % (name, val))
    convert(path, out)
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
