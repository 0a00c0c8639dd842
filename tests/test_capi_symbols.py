"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/aarmvs.h declares
(no compute calls here)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "aarmvs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#ifdef AARMVS_DEBUG.*?#endif", "", text, flags=re.S)      # debug hooks: -DAARMVS_DEBUG builds only
    return sorted(set(re.findall(r"\b(aarmvs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from aarmvsnet_b200 import build, capi
    path = build.build_library()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(capi.PROTOTYPES) == names        # the ctypes table and the header agree
    assert lib.aarmvs_version() == capi.ABI_VERSION
    for n in capi.DEBUG_PROTOTYPES:                 # the product library carries no process-global debug switches
        assert not hasattr(lib, n), n


def test_library_is_sm100a_and_has_no_torch_dependency():
    from aarmvsnet_b200 import build
    path = build.build_library()
    out = subprocess.run(["cuobjdump", "--list-elf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    dyn = subprocess.run(["readelf", "-d", path], capture_output=True, text=True).stdout
    needed = [ln for ln in dyn.splitlines() if "NEEDED" in ln]
    assert needed and not any("torch" in ln or "c10" in ln or "cudart" in ln for ln in needed)     # static cudart, no torch


def test_bad_arguments_are_rejected_without_a_gpu():
    from aarmvsnet_b200 import capi
    lib = capi.load()
    assert lib.aarmvs_warp(None, None, 4, 4, None, 1.0, 0, None) == capi.ERR_ARG
    assert lib.aarmvs_next_plane(None, None, 3, None) == capi.ERR_ARG
    assert lib.aarmvs_omega_stats_bytes(6) == 6 * 6 * 8
