"""Loader for the in-tree sm_100a extension (``lstm_tensorspark_b200/_C*.so``).

There is no eager fallback on the GPU: if a CUDA tensor reaches an op and the extension cannot be imported,
``ext()`` raises with the build command.  ``LSTM_TS_BUILD=1`` builds on demand (used by tests / CI)."""
from __future__ import annotations

import importlib
import os

_EXT = None
_ERR = None
LAUNCHES = {"n": 0}          # number of OUR kernels launched through the extension (bench.py "gpu_launches")
_NO_KERNEL = {"ar_max_blocks", "ar_flag_words", "ar_slots"}


class _Counting:
    """Thin proxy over the pybind module that counts kernel launches."""

    def __init__(self, mod):
        object.__setattr__(self, "_m", mod)

    def __getattr__(self, name):
        fn = getattr(self._m, name)
        if name in _NO_KERNEL or not callable(fn):
            return fn

        def call(*a, **k):
            LAUNCHES["n"] += 1
            return fn(*a, **k)
        object.__setattr__(self, name, call)
        return call



def ext():
    global _EXT, _ERR
    if _EXT is not None:
        return _EXT
    try:
        _EXT = _Counting(importlib.import_module("lstm_tensorspark_b200._C"))
        return _EXT
    except Exception as e:                       # noqa: BLE001
        _ERR = e
    if os.environ.get("LSTM_TS_BUILD", "0") == "1":
        from .. import build as _b
        _b.build()
        importlib.invalidate_caches()
        _EXT = _Counting(importlib.import_module("lstm_tensorspark_b200._C"))
        return _EXT
    raise RuntimeError(
        "lstm_tensorspark_b200._C (the hand-written sm_100a kernels) is not built/importable: "
        f"{_ERR!r}.  Run `python -m lstm_tensorspark_b200.build` (or __graft_entry__.build()).  "
        "There is deliberately no PyTorch fallback on CUDA tensors.")


def available() -> bool:
    try:
        ext()
        return True
    except Exception:                            # noqa: BLE001
        return False
