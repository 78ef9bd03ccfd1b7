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
    # the C compiler's own layout of the header's structs (a plain C translation unit: the header is C)
    import subprocess, tempfile
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "esa_pose_b200.h"\n'
           'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(epb_voting_params), sizeof(epb_voting_io),'
           ' offsetof(epb_voting_params, sb), offsetof(epb_voting_params, philox_seed),'
           ' offsetof(epb_voting_params, stage)); return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o",
                        os.path.join(d, "t")], check=True)
        sz_p, sz_io, off_sb, off_seed, off_stage = map(int, subprocess.run([os.path.join(d, "t")], capture_output=True,
                                                                           text=True, check=True).stdout.split())
    assert ctypes.sizeof(_lib.VotingParams) == sz_p
    assert ctypes.sizeof(_lib.VotingIO) == sz_io == 15 * ctypes.sizeof(ctypes.c_void_p)
    assert _lib.VotingParams.sb.offset == off_sb and _lib.VotingParams.philox_seed.offset == off_seed
    assert _lib.VotingParams.stage.offset == off_stage
    lib = _lib.load()
    p = _lib.VotingParams()
    assert lib.epb_voting_workspace_bytes(p) == 0                      # invalid (all-zero) parameters
    p.mode, p.B, p.H, p.W, p.vn, p.hn, p.rounds = 0, 2, 64, 64, 11, 128, 1
    n = lib.epb_voting_workspace_bytes(p)
    assert n >= 2 * 64 * 64 * 4 + 2 * 11 * 128 * 12
    # config[1] with torch-compatible Philox draws: the workspace is bounded by max_num, not by H*W
    # (373 MB when sized by the field): voting pixels per image <= 30000 + 8 sigma + 1024
    p.B, p.H, p.W, p.hn, p.max_num, p.rng_mode = 64, 256, 256, 512, 30000, _lib.RNG_PHILOX
    cap = (30000 + int(8 * 30000 ** 0.5) + 1024 + 127) // 128 * 128
    n1 = lib.epb_voting_workspace_bytes(p)
    assert 64 * cap * (4 + 11 * 8) <= n1 <= 64 * cap * (4 + 11 * 8) + 64 * 11 * 512 * 12 + (1 << 20)
    p.H = p.W = 128                      # an image smaller than max_num: bounded by the image
    assert lib.epb_voting_workspace_bytes(p) <= 64 * 128 * 128 * (4 + 11 * 8) + 64 * 11 * 512 * 12 + (1 << 20)


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "esa_pose_estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src, f
                # /root/reference may be CITED (file:line in comments and docstrings) but never opened or imported
                for ln in src.splitlines():
                    if "/root/reference" in ln:
                        assert not re.search(r"\b(open|import|sys\.path|CDLL|dlopen|fopen|include)\b", ln), (f, ln)


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
