"""Test helper: the two builds of the SAME kernel sources the parity tests run against.

* ``gpu``  sygnals_b200/libsygb200.so (nvcc, sm_100a) -- the product; tests using it are marked ``gpu``.
* ``emu``  tests/emu/libsygb200_emu.so (g++ -DSYG_EMU, one fiber per CUDA thread) -- TEST INFRASTRUCTURE ONLY, built on
           demand; lets the kernels' index arithmetic / barriers / host logic be checked in the GPU-less container.
           The sygnals_b200 package never loads it.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

from sygnals_b200 import _ffi  # noqa: E402

_cache = {}


def get_engine(kind: str) -> "_ffi.Engine":
    if kind in _cache:
        return _cache[kind]
    if kind == "gpu":
        eng = _ffi.engine(0)
    elif kind == "emu":
        import build_emu
        lib = _ffi.Library(build_emu.build_emu())
        eng = _ffi.Engine(0, lib)
    else:
        raise ValueError(kind)
    eng.test_backend = kind                 # tests ask the engine which build it is (never load the emulator to compare)
    _cache[kind] = eng
    return eng


BACKENDS = ["emu", pytest.param("gpu", marks=pytest.mark.gpu)]
