"""Loader for the in-tree sm_100a extension (``feddrift_b200/_C/fdb200_C*.so``).

The extension is built by ``__graft_entry__.build()`` (nvcc cross-compiles
without a GPU).  Policy: on a CUDA machine a missing extension is a hard error
(no silent eager fallback); on CPU-only machines every op falls back to its
plain-PyTorch fp32 reference, which is also the numerics oracle of the tests.
"""
from __future__ import annotations

import glob
import importlib.util
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(os.path.dirname(_HERE), "_C")
CSRC_DIR = os.path.join(os.path.dirname(_HERE), "csrc")
EXT_NAME = "fdb200_C"
_lock = threading.Lock()
_mod = None
_tried = False

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--use_fast_math=false",
    "-std=c++17", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC_DIR, "*.cu")) + glob.glob(os.path.join(CSRC_DIR, "*.cpp")))


def build(verbose: bool = False):
    """Compile the extension in-tree (works on a CPU-only box; nvcc cross-compiles)."""
    from torch.utils.cpp_extension import load
    os.makedirs(BUILD_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    mod = load(name=EXT_NAME, sources=sources(), extra_cuda_cflags=flags, extra_cflags=["-O3", "-std=c++17"],
               build_directory=BUILD_DIR, verbose=verbose, with_cuda=True)
    global _mod, _tried
    _mod, _tried = mod, True
    return mod


def _so_path():
    hits = sorted(glob.glob(os.path.join(BUILD_DIR, EXT_NAME + "*.so")))
    return hits[-1] if hits else None


def load(required: bool = False):
    """Return the extension module or None (CPU-only, not built)."""
    global _mod, _tried
    if _mod is not None:
        return _mod
    with _lock:
        if _mod is None and not _tried:
            _tried = True
            path = _so_path()
            if path is not None:
                spec = importlib.util.spec_from_file_location(EXT_NAME, path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _mod = mod
    if _mod is None and (required or torch.cuda.is_available()):
        raise RuntimeError(
            f"feddrift_b200: native extension {EXT_NAME} not found in {BUILD_DIR}; run "
            "`python -c 'import __graft_entry__ as g; g.build()'` — refusing to fall back to eager on a GPU box")
    return _mod


def available() -> bool:
    try:
        return load() is not None
    except RuntimeError:
        return False


def use_native(*tensors) -> bool:
    """True when the inputs live on CUDA (then the extension MUST be present)."""
    if any(t is not None and t.is_cuda for t in tensors):
        load(required=True)
        return True
    return False
