"""CPU-side checks of the drop-in boundary: libkvae.so loads, exports every symbol include/kvae.h
declares, and the Python binding table covers exactly those symbols.  No compute calls."""
import ctypes
import os
import re

import pytest

import helpers  # noqa: F401  (sys.path)
from kalle_audio_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "kvae.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kvae_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/kvae.h but not exported by libkvae.so"


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES.keys()) == _declared()


def test_version_and_no_gpu_behaviour():
    L = _lib.lib()
    assert L.kvae_version() == 100
    import torch
    if not torch.cuda.is_available():
        assert L.kvae_device_count() == 0
        arch = _lib.KvaeArch()
        arch.io_channels, arch.channels, arch.latent_dim, arch.n_stages = 2, 8, 4, 1
        arch.c_mults[0], arch.strides[0] = 2, 2
        h = ctypes.c_void_p()
        rc = L.kvae_plan_create(ctypes.byref(arch), _lib.KVAE_DECODER, _lib.KVAE_PREC_F32, 0, ctypes.byref(h))
        assert rc != 0 and b"no CUDA device" in L.kvae_last_error()


def test_bad_arguments_rejected_without_gpu():
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.kvae_plan_create(None, 0, 0, 0, ctypes.byref(h)) != 0
    arch = _lib.KvaeArch()
    arch.io_channels, arch.channels, arch.latent_dim, arch.n_stages = 2, 8, 4, 9
    assert L.kvae_plan_create(ctypes.byref(arch), 1, 0, 0, ctypes.byref(h)) != 0
    assert b"n_stages" in L.kvae_last_error()
    arch.n_stages = 1
    arch.c_mults[0], arch.strides[0] = 2, 16
    assert L.kvae_plan_create(ctypes.byref(arch), 1, 0, 0, ctypes.byref(h)) != 0
    assert b"stride" in L.kvae_last_error()
