"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports
every symbol include/b200vor.h declares, and fails LOUDLY (no CPU fallback) when no
CUDA device is present."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200vor.h")


@pytest.fixture(scope="module")
def lib():
    from geogram_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return capi.load_library()


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200vor_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from geogram_b200 import capi
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(capi.EXPORTS) == names


def test_version(lib):
    assert lib.b200vor_version() == 1


def test_no_silent_cpu_fallback(lib):
    """Without a CUDA device the product path must refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from geogram_b200 import Context, B200VorError
    with pytest.raises(B200VorError) as e:
        Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_link_oracle():
    """The shipped library must not depend on anything under oracle/."""
    from geogram_b200 import capi
    out = subprocess.check_output(["ldd", capi.LIB_PATH]).decode()
    assert "oracle" not in out and "georef" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "geogram_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                src = open(os.path.join(root, f)).read()
                assert "voro_oracle" not in src and "georef" not in src, f


def test_sass_is_sm100a():
    from geogram_b200 import capi
    out = subprocess.check_output(["cuobjdump", "-lelf", capi.LIB_PATH]).decode()
    assert "sm_100a" in out
    assert "sm_90" not in out and "sm_80" not in out


def test_ctypes_structs_match_the_header(tmp_path):
    """The ctypes mirrors in geogram_b200/capi.py have the size and the field offsets the C compiler gives
    the structs of include/b200vor.h (an ABI drift would corrupt results silently)."""
    from geogram_b200 import capi
    pairs = [("b200vor_cell_params", capi.CellParams), ("b200vor_stats", capi.Stats),
             ("b200vor_otm_params", capi.OtmParams), ("b200vor_otm_result", capi.OtmResult),
             ("b200vor_peaks", capi.Peaks), ("b200vor_slab_stats", capi.SlabStats)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200vor.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append('printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (cname, fname))
        lines.append('printf("\\n");')
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    for (cname, cls), line in zip(pairs, out):
        f = line.split()
        assert f[0] == cname
        assert int(f[1]) == ctypes.sizeof(cls), cname
        for (fname, _), off in zip(cls._fields_, f[2:]):
            assert int(off) == getattr(cls, fname).offset, (cname, fname)
