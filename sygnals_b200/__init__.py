"""sygnals_b200 -- B200-native (sm_100a) engine for the Sygnals segment->features hot path.

Layout (only what the path needs):
  csrc/                     hand-written CUDA kernels + the C ABI (include/sygb200.h) -> libsygb200.so
  _ffi.py                   ctypes binding of the C ABI
  core/                     host-side mirror of the reference interface for this path (same names / signatures /
                            errors as sygnals.core.{dsp,segmentation,features.*,audio.features})
  batch.py                  additive batched entry points (clips / segment tables, torch tensors in HBM)
  dist.py                   unit sharding over ranks + the final NCCL gather of the feature matrix
  plugin.py, plugin.toml    SygnalsPluginBase plugin that rebinds the reference's functions to this engine

There is no CPU fallback: every compute call goes through libsygb200.so and raises if it (or a CUDA device) is missing.
"""
__version__ = "0.1.0"

from . import _ffi  # noqa: F401  (does not load the shared library until first use)


def library_path() -> str:
    return _ffi.default_library_path()
