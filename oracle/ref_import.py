"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference module ``/root/reference/utils_supersdr.py`` with its GUI /
audio / network dependencies (pygame, sounddevice, tkinter, xmltodict, requests) replaced by
``MagicMock`` modules, so the reference's own numeric functions can be run headless:

* ``filtering``                          utils_supersdr.py:333-348
* ``kiwi_waterfall.spectrum_db2col``     utils_supersdr.py:787-813
* ``kiwi_waterfall.run`` averaging       utils_supersdr.py:879-888
* ``kiwi_sound.play_buffer``             utils_supersdr.py:1106-1148
* ``kiwi_sound.process_audio_stream``    utils_supersdr.py:1044-1076

``/root/reference`` exists only in the build container, not on the GPU box; this module is used
by ``oracle/make_golden.py`` (to generate the committed fixtures under ``tests/golden/``) and by
the CPU tests that are skipped when the reference tree is absent.
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_DIR = os.environ.get("SSDR_REFERENCE_DIR", "/root/reference")

_STUBBED = ["pygame", "pygame.locals", "pygame.font", "pygame.event", "pygame.draw",
            "pygame.freetype", "sounddevice", "tkinter", "tkinter.ttk", "tkinter.font",
            "tkinter.messagebox", "xmltodict", "requests"]

_cached = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "utils_supersdr.py"))


def _pygame_locals():
    """utils_supersdr.py:69-71 builds key tables from pygame.locals at import time."""
    mod = types.ModuleType("pygame.locals")
    code = 1000
    for i in range(10):
        setattr(mod, "K_%d" % i, code); code += 1
        setattr(mod, "K_KP%d" % i, code); code += 1
    for name in ("K_BACKSPACE", "K_RETURN", "K_ESCAPE", "K_KP_ENTER", "K_MINUS", "K_PERIOD",
                 "K_KP_PERIOD", "K_KP_MINUS"):
        setattr(mod, name, code); code += 1
    mod.__all__ = [k for k in vars(mod) if k.startswith("K_")]
    return mod


def load():
    """Return the reference ``utils_supersdr`` module (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    saved = {k: sys.modules.get(k) for k in _STUBBED}
    for name in _STUBBED:
        sys.modules[name] = MagicMock(name=name)
    sys.modules["pygame.locals"] = _pygame_locals()
    tk = types.ModuleType("tkinter")          # `from tkinter import *` needs a real module
    tk.__all__ = []
    tk.Tk = MagicMock(); tk.Toplevel = MagicMock()
    sys.modules["tkinter"] = tk
    cwd = os.getcwd()
    sys.path.insert(0, REFERENCE_DIR)
    try:
        os.chdir(REFERENCE_DIR)               # font files are opened relative to cwd at import
        mod = importlib.import_module("utils_supersdr")
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE_DIR)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached = mod
    return mod
