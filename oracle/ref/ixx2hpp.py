#!/usr/bin/env python3
"""C++23 named-module interface (.ixx, MSVC) -> plain header, for oracle/_ref (TEST INFRASTRUCTURE).

The reference's host math (src/stx/math.ixx: vec/mat/look/perspective/inverse) and camera (src/gfx/camera.ixx) decide
the ray-generation matrices bit for bit.  g++ 13 cannot consume MSVC header units, so the module
syntax -- and only that -- is rewritten; every expression stays the reference's own text:
    export module X;        -> (dropped)
    import <header>;        -> #include <header>
    import minote.name;     -> #include "minote.name.hpp"     (generated next to this one)
    export <declaration>    -> <declaration>
Additionally the nested `struct Params { ... };` of src/gfx/modules/sky.ixx (the std140 mirror of the atmosphere block with
its `earth()` factory, :28-84) is cut out by brace matching into sky_params.hpp -- the rest of that file is Vulkan glue.
Output goes under oracle/_ref/gen/host/ (git-ignored).
"""
import argparse
import os
import re
import sys

MODULES = {  # module name -> path under src/
    "minote.types": "stx/types.ixx",
    "minote.concepts": "stx/concepts.ixx",
    "minote.ranges": "stx/ranges.ixx",
    "minote.math": "stx/math.ixx",
    "minote.camera": "gfx/camera.ixx",
}


def translate(text):
    text = re.sub(r"^\s*export\s+module\s+[\w.]+\s*;\s*$", "", text, flags=re.M)
    text = re.sub(r"^\s*(?:export\s+)?import\s+<([^>]+)>\s*;", r"#include <\1>", text, flags=re.M)
    text = re.sub(r"^\s*(?:export\s+)?import\s+([\w.]+)\s*;", r'#include "\1.hpp"', text, flags=re.M)
    text = re.sub(r"\bexport\s+", "", text)
    return "#pragma once\n" + text


def extract_block(text, head):
    """The text of `head { ... };` (balanced braces)."""
    i = text.index(head)
    j = text.index("{", i)
    depth, k = 1, j + 1
    while depth:
        depth += text[k] == "{"
        depth -= text[k] == "}"
        k += 1
    return text[i:k] + ";"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference/src")
    ap.add_argument("--dst", required=True)
    a = ap.parse_args()
    os.makedirs(a.dst, exist_ok=True)
    for mod, rel in MODULES.items():
        path = os.path.join(a.src, rel)
        if not os.path.exists(path):
            sys.exit("missing " + path)
        with open(path, encoding="utf-8-sig") as fh:
            text = fh.read()
        with open(os.path.join(a.dst, mod + ".hpp"), "w", encoding="utf-8") as fh:
            fh.write("// GENERATED from the reference's src/%s by oracle/ref/ixx2hpp.py -- do not commit\n" % rel)
            fh.write(translate(text))
    with open(os.path.join(a.src, "gfx/modules/sky.ixx"), encoding="utf-8-sig") as fh:
        params = extract_block(fh.read(), "struct Params")
    with open(os.path.join(a.dst, "sky_params.hpp"), "w", encoding="utf-8") as fh:
        fh.write("// GENERATED from the reference's src/gfx/modules/sky.ixx (struct Atmosphere::Params) -- do not commit\n")
        fh.write('#pragma once\n#include "minote.math.hpp"\n' + params + "\n")
    print("ixx2hpp: %d modules -> %s" % (len(MODULES), a.dst))


if __name__ == "__main__":
    main()
