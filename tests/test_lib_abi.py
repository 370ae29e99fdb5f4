"""CPU: libssdhead.so loads without a GPU and exports every symbol include/ssdhead.h declares; the ctypes
table mirrors the header; no compute entry point is called."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ssdhead.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssdhead_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from objectdetection_ssd_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ssdhead.h but not exported by libssdhead.so"


def test_ctypes_table_covers_the_header(lib):
    from objectdetection_ssd_b200 import _lib
    assert set(header_symbols()) == set(_lib.SIGNATURES), set(header_symbols()) ^ set(_lib.SIGNATURES)


def test_metadata_entry_points(lib):
    from objectdetection_ssd_b200 import _lib
    assert lib.ssdhead_abi_version() == 4
    assert lib.ssdhead_error_string(0) == b"ok"
    assert b"workspace" in lib.ssdhead_error_string(-3)
    assert lib.ssdhead_workspace_bytes(_lib.WS_MATCH, 32, 8732, 21, 200) > 0
    assert lib.ssdhead_workspace_bytes(_lib.WS_LOSS, 32, 8732, 21, 0) >= 32 * 8732 * 4
    assert lib.ssdhead_workspace_bytes(_lib.WS_LOSS, 1, 100000, 21, 0) == 0          # beyond the shared-memory bound
    assert lib.ssdhead_workspace_bytes(_lib.WS_DETECT, 64, 8732, 21, 0) > 64 * 8732 * 16
    assert lib.ssdhead_workspace_bytes(_lib.WS_ROWS, 256, 8732, 21, 0) >= 256 * (8732 * 2 + 8)
    assert lib.ssdhead_workspace_bytes(_lib.WS_ROWS, 1, 70000, 21, 0) == 0           # row indices are 16 bits
    assert lib.ssdhead_workspace_bytes(_lib.WS_DETECT, 64, 8732, 21, 1024) < lib.ssdhead_workspace_bytes(_lib.WS_DETECT, 64, 8732, 21, 0)
    # three key buffers of 20 keys per prior (every row of the short-list route owns its 20 slots: nothing can overflow)
    assert lib.ssdhead_workspace_bytes(_lib.WS_DETECT, 256, 8732, 21, 0) >= 3 * 256 * 20 * 8732 * 8
    # CE + best-gt map + the 65 536 seeds of the fused match's cull
    assert lib.ssdhead_workspace_bytes(_lib.WS_LOSS, 128, 24564, 21, 0) >= 128 * 24564 * 6 + 65536 * 4


def test_bad_arguments_are_rejected_without_touching_the_gpu(lib):
    assert lib.ssdhead_match(None, None, None, None, 4, 8732, 21, 10, 0.5, None, None, None, None, None, None, 0, None) == -1
    assert lib.ssdhead_ce_stream(None, 4, 8732, 21, None, None, None, None, 0, None) == -1
    assert lib.ssdhead_iou_matrix(None, -1, None, 3, None, None) == -1
    assert lib.ssdhead_ctx_multibox_loss_dev(None, None, None, None, None, None, 1, 1, 3, 0.5, None, None, None, None, None) == -1


def test_levels_struct_layout_matches_the_header(tmp_path):
    """ctypes `Levels` must have the layout a C compiler gives `ssdhead_levels` (offsets and size), else the per-level
    entry points would read garbage."""
    import ctypes as C
    import shutil
    import subprocess
    from objectdetection_ssd_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ssdhead.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %d\\n", sizeof(ssdhead_levels), offsetof(ssdhead_levels, num_levels),\n'
        '  offsetof(ssdhead_levels, count), offsetof(ssdhead_levels, conf), offsetof(ssdhead_levels, loc),\n'
        '  offsetof(ssdhead_levels, grad_conf), offsetof(ssdhead_levels, grad_loc), SSDHEAD_MAX_LEVELS); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    L = _lib.Levels
    want = [C.sizeof(L), L.num_levels.offset, L.count.offset, L.conf.offset, L.loc.offset, L.grad_conf.offset,
            L.grad_loc.offset, _lib.MAX_LEVELS]
    assert got == want, (got, want)
