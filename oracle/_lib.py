"""TEST INFRASTRUCTURE ONLY: ctypes loaders for oracle/liboracle.so and oracle/_ref/*.so."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _load(path, required=True):
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        if path.endswith("liboracle.so"):
            build()
        elif required:
            raise FileNotFoundError(path + " (built by `make -C oracle` where /root/reference exists)")
        else:
            return None
    lib = ctypes.CDLL(path)
    _cache[path] = lib
    return lib


def oracle_lib():
    return _load(os.path.join(_HERE, "liboracle.so"))


def ref_pnp_lib(required=True):
    return _load(os.path.join(_HERE, "_ref", "libuncertainty_pnp_ref.so"), required)


def ref_voting_lib(required=True):
    return _load(os.path.join(_HERE, "_ref", "libref_voting.so"), required)


def ref_nearest_lib(required=True):
    return _load(os.path.join(_HERE, "_ref", "libref_nearest.so"), required)


def fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def u8ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
