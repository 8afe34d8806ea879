"""CPU: the C-ABI shared library loads and exports every symbol include/ssdr_b200.h declares; the
product path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

from conftest import HAS_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "ssdr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ssdr_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import supersdr_b200 as S
    lib = ctypes.CDLL(S.LIB_PATH)
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(S.EXPORTS) == names, "ctypes prototypes and header disagree"
    assert lib.ssdr_abi_version() == 1


def _py_code_tokens(path):
    """Source of a .py file with comments and string literals (docstrings) removed."""
    import io
    import tokenize
    out = []
    with open(path, "rb") as fh:
        for tok in tokenize.tokenize(fh.readline):
            if tok.type not in (tokenize.COMMENT, tokenize.STRING, tokenize.ENCODING):
                out.append(tok.string)
    return " ".join(out)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the product package imports, includes, loads or executes it.
    Comments and docstrings may cite oracle files; code may not name them."""
    import subprocess
    pkg = os.path.join(ROOT, "supersdr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dp, f)
            if f.endswith(".py"):
                src = open(path).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "oracle" not in _py_code_tokens(path), f          # no identifier / attribute named oracle
                for lit in re.findall(r"(?:CDLL|open|Popen|check_call|run)\(([^)]*)\)", src):
                    assert "oracle" not in lit, (f, lit)
            elif f.endswith((".cu", ".cuh", ".h", "Makefile")):
                src = open(path).read()
                code = re.sub(r"//[^\n]*|/\*.*?\*/", "", src, flags=re.S)      # strip C/C++ comments
                code = re.sub(r"(?m)^#(?!include).*$", "", code) if f != "Makefile" else re.sub(r"(?m)#.*$", "", src)
                assert "oracle" not in code, f
    import supersdr_b200 as S
    needed = subprocess.run(["readelf", "-d", S.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed                                        # the shared library does not link the checker


@pytest.mark.skipif(HAS_GPU, reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    import numpy as np
    import supersdr_b200 as S
    with pytest.raises(S.SsdrError):
        S.WaterfallBank(1024, 1, 1)
    with pytest.raises(S.SsdrError):
        S.DemodBank(1, 512)
    with pytest.raises(S.SsdrError):
        S.filtering(6000, 48000).lowpass(np.zeros(100))


def test_host_side_design_and_params():
    import numpy as np
    import supersdr_b200 as S
    from oracle import tier_p, tier_u
    f = S.filtering(6000, 48000)                       # utils_supersdr.py:333-344 known answers
    assert f.n_tap == 33 and np.array_equal(f.h, tier_p.fir_design(6000, 48000))
    for n in (256, 1024, 2048, 16384):
        lo, g = S.percentile_index(n)
        lo2, g2 = tier_p.percentile_virtual_index(n, 40.0)
        assert lo == lo2 and np.float32(g) == g2
    assert S.percentile_index(1024) == (409, float(np.float32(409.2) - np.float32(409)))
    p = S.demod_params("cw")
    assert (p.low_cut_hz, p.high_cut_hz) == (400.0, 800.0) and p.agc_thresh_dbm == -80 and p.agc_decay_ms == 4000
    taps = np.array(p.taps[:], np.float32)
    assert np.array_equal(taps, tier_u.demod_taps(400, 800).astype(np.float32))
    assert S.default_passband("lsb") == (-3000, -30) and S.default_passband("am") == (-6000, 6000)
