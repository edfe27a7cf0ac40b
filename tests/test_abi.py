"""The C-ABI library loads on a CPU-only box and exports every symbol include/minotert.h declares.
No compute calls here (no GPU); the only call is mrt_create, which must fail loudly without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "minotert.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from minotert_b200 import capi
    L = capi.load()
    names = declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(L, n), f"libminotert.so does not export {n}"
    assert set(capi.EXPORTS) <= set(names)
    assert L.mrt_abi_version() == 1


def test_pod_layouts_match_reference_ubos():
    from minotert_b200 import capi
    assert C.sizeof(capi.PrimaryConstants) == 5 * 64 + 4   # primaryRay.comp:14-21
    assert C.sizeof(capi.SecondaryConstants) == 4 * 64 + 16  # secondaryRays.comp:25-32
    assert C.sizeof(capi.Sphere) == 28                       # intersect.glsl:9-13
    assert C.sizeof(capi.AtmosphereParams) == 144            # sky.ixx:28-56 (std140)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from minotert_b200 import capi
    with pytest.raises(capi.MinoteError, match="no CPU fallback"):
        capi.Context(0)


def test_partition_rows_host_logic():
    from minotert_b200 import capi
    L = capi.load()
    import numpy as np
    for H in (1, 7, 64, 101, 2160):
        for nranks in (1, 2, 3, 4, 8):
            for slab in (1, 8, 32):
                seen = np.zeros(H, int)
                for r in range(nranks):
                    n = C.c_uint32()
                    assert L.mrt_partition_rows_for(r, nranks, slab, H, None, C.byref(n)) == 0
                    rows = np.zeros(n.value, np.uint32)
                    L.mrt_partition_rows_for(r, nranks, slab, H, rows.ctypes.data_as(C.c_void_p), C.byref(n))
                    assert np.all(np.diff(rows.astype(int)) > 0) if n.value > 1 else True
                    assert np.all((rows // slab) % nranks == r)
                    seen[rows] += 1
                assert np.all(seen == 1)
    n = C.c_uint32()
    assert L.mrt_partition_rows_for(2, 2, 8, 64, None, C.byref(n)) != 0
