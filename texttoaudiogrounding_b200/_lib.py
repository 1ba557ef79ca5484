"""Build and load libtag_b200.so (the C-ABI CUDA library) and bind it with ctypes.

The ctypes signatures are derived from ``include/tag_b200.h`` so that the header is the
single source of truth for the ABI.  There is no CPU fallback: if the library is missing or a
call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess
import sys
from typing import Dict, List, Tuple

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
HEADER = os.path.join(ROOT, "include", "tag_b200.h")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtag_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]

_CTYPES = {
    "int": ctypes.c_int, "long": ctypes.c_long, "long long": ctypes.c_longlong,
    "float": ctypes.c_float, "double": ctypes.c_double, "uint64_t": ctypes.c_uint64,
    "cudaStream_t": ctypes.c_void_p,
}


def parse_header(path: str = HEADER) -> Dict[str, List[Tuple[str, str]]]:
    """name -> [(ctype string, arg name)] for every ``int tag_*(...)`` prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(tag_\w+)\s*\(([^)]*)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    parsed.append(("ptr", a.split("*")[-1].strip()))
                else:
                    toks = a.replace("const ", "").split()
                    parsed.append((" ".join(toks[:-1]), toks[-1]))
        protos[name] = parsed
    return protos


def sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link lib/libtag_b200.so (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    objs = []
    flags = list(NVCC_FLAGS)
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + flags + ["-I", CSRC, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out.decode()}")
        if verbose and out:
            print(out.decode(), file=sys.stderr)
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout.decode()}")
    return LIB_PATH


_LIB = None


def lib() -> ctypes.CDLL:
    """The loaded library with argtypes set.  Raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the CUDA kernels)")
    cdll = ctypes.CDLL(LIB_PATH)
    for name, args in parse_header().items():
        fn = getattr(cdll, name)      # AttributeError here == header/library mismatch
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p if t == "ptr" else _CTYPES[t] for t, _ in args]
    _LIB = cdll
    return cdll


class TagError(RuntimeError):
    pass


def check(code: int, name: str) -> None:
    if code == 0:
        return
    if code >= 10001:
        what = {10001: "bad argument", 10002: "unsupported configuration"}.get(code, "error")
        raise TagError(f"{name}: {what} (code {code})")
    raise TagError(f"{name}: CUDA error {code}")
