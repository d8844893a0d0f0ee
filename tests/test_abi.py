"""CPU: the C-ABI shared library loads and exports every symbol include/gs_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from gs_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gs_b200.h")).read()
    return re.findall(r"GSB_API\s+[\w\s\*]+?\b(gsb_\w+)\s*\(", text)


def test_every_declared_symbol_is_exported():
    L = lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert getattr(L, s) is not None, s
    assert set(lib.EXPORTED_SYMBOLS) == set(syms)


def test_host_only_entry_points():
    L = lib.lib()
    assert b"sm_100a" in L.gsb_version()
    g1, g2 = L.gsb_geom_bytes(1000), L.gsb_geom_bytes(2000)
    assert 0 < g1 < g2 and g2 - g1 >= 1000 * 104          # rec 48 + rect 8 + clamped 1 + accumulator 48, modulo alignment
    assert L.gsb_image_bytes(1920, 1080) >= 1920 * 1080 * 8
    assert L.gsb_binning_bytes(10 ** 6) >= 10 ** 6 * 20
    assert L.gsb_launch_count() >= 0


def test_struct_layouts_match_header():
    # field order / sizes of the ctypes mirrors against the C declarations (x86-64 SysV)
    assert ctypes.sizeof(lib.GsbQuant) == 6 * 8
    assert ctypes.sizeof(lib.GsbCamera) == 4 * 4 + 4 * 8 + 8
    assert ctypes.sizeof(lib.GsbGrads) == 9 * 8 + 8
    assert ctypes.sizeof(lib.GsbDebug) == 7 * 8
    assert lib.GsbScene.means3D.offset == 8 and lib.GsbScene.scale_modifier.offset == 8 + 8 * 8
    assert lib.GsbScene.band_count.offset == lib.GsbScene.scale_modifier.offset + 8
    assert ctypes.sizeof(lib.GsbScene) == lib.GsbScene.quant.offset + 8


def test_sass_is_blackwell_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out
