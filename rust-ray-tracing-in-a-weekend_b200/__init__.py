"""B200-native path-tracing backend for the rust-ray-tracing-in-a-weekend scene API.

Only what the render hot path needs lives here:
  csrc/      hand-written CUDA (sm_100a) + the C ABI of include/rtw.h  -> librtw.so
  host/      C++ mirror of the reference's constructor surface (flatten -> C ABI)
  api.py     ctypes binding of the C ABI
  scenes.py  the reference's scene functions restated as seeded input generators
"""
from . import api, scenes  # noqa: F401
from .api import Camera, RenderParams, RtwError, Scene, load_rtw, make_params  # noqa: F401

__all__ = ["api", "scenes", "Camera", "RenderParams", "RtwError", "Scene", "load_rtw", "make_params"]
