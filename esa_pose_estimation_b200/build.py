"""Builds libesa_pose_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  `python -m esa_pose_estimation_b200.build` or `build()` from __graft_entry__.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libesa_pose_b200.so")
SOURCES = ["api.cu", "decode.cu", "voting.cu", "pose.cu", "metrics.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "esa_pose_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]
# --use_fast_math is never enabled: the voting kernels rely on IEEE sqrt/div.
NVCC_FLAGS = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
# tuning experiments only (e.g. EPB_EXTRA_NVCC_FLAGS="-DEPB_VOTE_MINB=8")
NVCC_FLAGS += os.environ.get("EPB_EXTRA_NVCC_FLAGS", "").split()


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, lib=LIB, extra_flags=(), objdir="build"):
    if not force and lib == LIB and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, objdir), exist_ok=True)
    for s in SOURCES:
        obj = os.path.join(HERE, objdir, s.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, out))
        if verbose and out.strip():
            print(out)
    cmd = [_nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("link failed:\n" + out.stdout + out.stderr)
    return lib


TUNING_LIB = os.path.join(HERE, "libesa_pose_b200_tuning.so")


def build_tuning(force=False, verbose=False):
    """The same library with -DEPB_TUNING: the only build in which the EPB_* environment knobs of
    csrc/common.cuh exist.  Used by the measurement scripts under tools/, never by the package."""
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    if not force and os.path.exists(TUNING_LIB) and all(
            os.path.getmtime(d) <= os.path.getmtime(TUNING_LIB) for d in deps if os.path.exists(d)):
        return TUNING_LIB
    return build(force=True, verbose=verbose, lib=TUNING_LIB, extra_flags=["-DEPB_TUNING"], objdir="build/tuning")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    if "--tuning" in sys.argv:
        print(build_tuning(force="--force" in sys.argv, verbose=True))
