"""The C-ABI shared library loads on a CPU box and exports every symbol include/pose2sim_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from pose2sim_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pose2sim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p2s_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("p2s_create", "p2s_destroy", "p2s_stage_observations_device", "p2s_triangulate_device",
                 "p2s_triangulate_host", "p2s_associate_device", "p2s_associate_host", "p2s_measure_fp64_peak"):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_status_strings_and_pure_helpers():
    lib = _lib.load()
    assert lib.p2s_status_string(0) == b"ok"
    assert b"no CPU fallback" in lib.p2s_status_string(2)
    assert lib.p2s_obs_bytes(1000, 8) == 1000 * 8 * 16


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device creating a handle must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pose2sim_b200 import ops
    with pytest.raises(_lib.P2SError) as ei:
        ops.Engine(0)
    assert ei.value.status == 2


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and a C program links against the library and calls the
    entry points that need no GPU."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    src = tmp_path / "use.c"
    src.write_text("""
#include <stdio.h>
#include "pose2sim_b200.h"
int main(void) {
    p2s_handle *h = 0;
    int rc = p2s_create(-1, &h);                 /* invalid device: must fail with a status, not crash */
    printf("%d %s %zu\\n", rc, p2s_status_string(rc), p2s_obs_bytes(10, 8));
    return rc == P2S_OK;
}
""")
    exe = tmp_path / "use"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-lp2s_b200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code, rest = r.stdout.split(" ", 1)
    assert int(code) == 2 and "no CPU fallback" in rest and rest.strip().endswith("1280")
