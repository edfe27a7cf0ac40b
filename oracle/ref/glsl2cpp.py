#!/usr/bin/env python3
"""GLSL -> C++ pre-pass for oracle/_ref (TEST INFRASTRUCTURE, not product code).

Reads the reference's shader sources where they lie (default /root/reference/src/gpu), applies a handful of
purely lexical rewrites and writes the result under oracle/_ref/gen/ (git-ignored: reference sources are never
committed).  The output is compiled as C++20 against oracle/ref/glsl_shim.hpp, so the arithmetic that runs is the
reference's own text.  Nothing here touches an expression's operators, operands or order; the rewrites are:

  1. `#version`, `#pragma`, `#extension` lines dropped.
  2. floating literals get an `f` suffix (GLSL literals are 32-bit floats; C++ would make them double).
  3. `out T x` / `inout T x` parameters become `T& x`.
  4. `layout(...)` resource declarations become plain C++ globals the binder fills in:
       layout(local_size_x = ..) in;                          -> GLSL_LOCAL_SIZE(x, y, z)
       layout(binding = N) [qualifiers] uniform TYPE NAME;    -> TYPE NAME;
       layout(..) uniform BLOCK { members } INST;             -> struct BLOCK_block { members } INST;
       layout(..) uniform BLOCK { members };                  -> members
       layout(constant_id = N) const T NAME = d;              -> const T NAME = GLSL_SPEC_CONSTANT_N;
  5. `shared` -> `static` (one work group runs at a time), `void main()` -> `void shader_main()`.
  6. vector/matrix constructor calls `vec3(a, b)` become braced `vec3{a, b}`: C++ leaves the evaluation order of
     function arguments unspecified (g++ goes right to left), braces pin GLSL's left-to-right order -- it matters
     in secondaryRays.comp:72, where both arguments advance the RNG.
"""
import argparse
import os
import re
import sys

CTOR_TYPES = ("vec2", "vec3", "vec4", "uvec2", "uvec3", "uvec4", "ivec2", "ivec3", "ivec4", "mat3", "mat4")

FLOAT_LIT = re.compile(
    r"(?<![A-Za-z0-9_.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![A-Za-z0-9_.])")


def split_code_comments(src):
    """Yield (is_code, text) pieces so rewrites never touch comments or string literals."""
    i, n, start = 0, len(src), 0
    while i < n:
        two = src[i:i + 2]
        if two == "//":
            if start < i:
                yield True, src[start:i]
            j = src.find("\n", i)
            j = n if j < 0 else j
            yield False, src[i:j]
            i = start = j
        elif two == "/*":
            if start < i:
                yield True, src[start:i]
            j = src.find("*/", i + 2)
            j = n if j < 0 else j + 2
            yield False, src[i:j]
            i = start = j
        elif src[i] == '"':
            if start < i:
                yield True, src[start:i]
            j = src.find('"', i + 1)
            j = n if j < 0 else j + 1
            yield False, src[i:j]
            i = start = j
        else:
            i += 1
    if start < n:
        yield True, src[start:]


def brace_ctors(code):
    """vecN( ... ) -> vecN{ ... } with parenthesis matching (code has no comments/strings here)."""
    pat = re.compile(r"\b(" + "|".join(CTOR_TYPES) + r")\s*\(")
    out, pos = [], 0
    while True:
        m = pat.search(code, pos)
        if not m:
            out.append(code[pos:])
            break
        # a declaration `vec3 name(...)`/function definition never has '(' right after the type name
        depth, j = 1, m.end()
        while j < len(code) and depth:
            depth += code[j] == "("
            depth -= code[j] == ")"
            j += 1
        if depth:  # unbalanced inside this piece (split by a comment): leave it alone
            out.append(code[pos:m.end()])
            pos = m.end()
            continue
        out.append(code[pos:m.start()] + m.group(1) + "{")
        out.append(brace_ctors(code[m.end():j - 1]) + "}")
        pos = j
    return "".join(out)


def rewrite_layouts(code):
    code = re.sub(r"layout\s*\(\s*local_size_x\s*=\s*(\d+)\s*(?:,\s*local_size_y\s*=\s*(\d+)\s*)?"
                  r"(?:,\s*local_size_z\s*=\s*(\d+)\s*)?\)\s*in\s*;",
                  lambda m: "GLSL_LOCAL_SIZE(%s, %s, %s)" % (m.group(1), m.group(2) or "1", m.group(3) or "1"), code)
    code = re.sub(r"layout\s*\(\s*constant_id\s*=\s*(\d+)\s*\)\s*const\s+(\w+)\s+(\w+)\s*=\s*[^;]+;",
                  r"const \2 \3 = GLSL_SPEC_CONSTANT_\1;", code)

    def block(m):
        name, body, inst = m.group(1), m.group(2), m.group(3)
        if inst:
            return "struct %s_block {%s} %s;" % (name, body, inst)
        return body.strip("\n")
    code = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+(\w+)\s*\{(.*?)\}\s*(\w*)\s*;", block, code, flags=re.S)
    code = re.sub(r"layout\s*\([^)]*\)\s*(?:(?:restrict|writeonly|readonly|coherent)\s+)*uniform\s+(\w+)\s+(\w+)\s*;",
                  r"\1 \2;", code)
    return code


def translate(src):
    src = re.sub(r"^[ \t]*#[ \t]*(version|pragma|extension)\b[^\n]*$", "", src, flags=re.M)
    pieces = list(split_code_comments(src))
    # layout blocks and constructor calls can span comments: do those on a comment-stripped view, keeping the
    # comments only where they sit between whole statements is not worth it -- drop comments from the output.
    code = "".join(t if is_code else (" " if t.startswith("/*") or t.startswith("//") else t)
                   for is_code, t in pieces)
    # `#include "x"` lines carry string literals: protect them
    lines = []
    for line in code.split("\n"):
        if re.match(r"\s*#\s*include\b", line):
            lines.append(line)
            continue
        lines.append(FLOAT_LIT.sub(lambda m: m.group(1) + "f", line))
    code = "\n".join(lines)
    code = rewrite_layouts(code)
    code = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", code)
    code = re.sub(r"\bshared\s+", "static ", code)
    code = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", code)
    code = brace_ctors(code)
    return code


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference/src/gpu")
    ap.add_argument("--dst", required=True)
    a = ap.parse_args()
    if not os.path.isdir(a.src):
        sys.exit("reference shaders not found at %s" % a.src)
    n = 0
    for root, _, files in os.walk(a.src):
        for f in files:
            if not f.endswith((".comp", ".glsl")):
                continue
            rel = os.path.relpath(os.path.join(root, f), a.src)
            out = os.path.join(a.dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            with open(os.path.join(root, f), encoding="utf-8") as fh:
                text = fh.read()
            with open(out, "w", encoding="utf-8") as fh:
                fh.write("// GENERATED from the reference's %s by oracle/ref/glsl2cpp.py -- do not commit\n" % rel)
                fh.write(translate(text))
            n += 1
    print("glsl2cpp: %d shader files -> %s" % (n, a.dst))


if __name__ == "__main__":
    main()
