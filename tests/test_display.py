"""Display epilogues: palette against the reference's create_cm (CPU), the scrolling image /
RGB / spectrum trace on the GPU against a numpy restatement of kiwi_waterfall.run's bookkeeping (utils_supersdr.py:
893-897) and display_stuff.plot_spectrum's arithmetic (:1678-1679)."""
import os
from collections import deque

import numpy as np
import pytest

from oracle import ref_import

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_palette_matches_reference_golden():
    import supersdr_b200 as S
    g = np.load(os.path.join(GOLD, "palette_cutesdr.npz"))["colormap"]
    cm = np.asarray(S.create_cm("cutesdr"), dtype=np.float64)
    assert cm.shape == (255, 3) and np.array_equal(cm, g)
    pal = S.palette_u8(cm)
    assert pal.shape == (256, 3) and pal.dtype == np.uint8
    assert tuple(pal[0]) == (0, 0, 0) and tuple(pal[86]) == (0, 255, 255) and tuple(pal[255]) == (255, 255, 255)
    assert np.array_equal(pal[:255], np.trunc(g).astype(np.uint8))


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_palette_matches_reference_live():
    import supersdr_b200 as S
    m = ref_import.load()
    disp = m.display_stuff.__new__(m.display_stuff)
    assert np.array_equal(np.asarray(disp.create_cm("cutesdr"), np.float64), np.asarray(S.create_cm(), np.float64))


@pytest.mark.gpu
def test_image_ring_rgb_and_trace(ssdr):
    B, H, W, SH = 3, 12, 1024, 200
    rng = np.random.default_rng(4)
    img = ssdr.WaterfallImage(B, H, W)
    # the reference's bookkeeping, per channel (utils_supersdr.py:692-693,893-897)
    wf_data = np.zeros((B, H, W))
    tmp = [deque([], 3) for _ in range(B)]
    run_index = 0
    for it in range(20):
        rows = (rng.uniform(0, 254, (B, W))).astype(np.float32)
        if it == 7:
            rows[1, 5:9] = np.nan                      # plot_spectrum uses nanmean
        img.push(rows)
        run_index += 1
        for b in range(B):
            tmp[b].appendleft(rows[b])
            if len(tmp[b]) > 0 and run_index > 3:
                wf_data[b, 1:, :] = wf_data[b, 0:-1, :]
                wf_data[b, 0, :] = tmp[b].pop()
        if it in (0, 2, 3, 4, 11, 19):
            rgb, data = img.image(want_rgb=True, want_data=True)
            assert np.array_equal(data, wf_data, equal_nan=True)
            idx = np.clip(np.rint(np.nan_to_num(wf_data, nan=0.0)), 0, 255).astype(np.uint8)
            assert np.array_equal(rgb, img.palette[idx])
            v, y = img.trace(15, SH)
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ref_v = np.stack([np.nanmean(wf_data[b].T[:, :15], axis=1) for b in range(B)])
            assert np.array_equal(v, ref_v, equal_nan=True)
            ref_y = np.array([[SH - 1 - int(x / 255 * SH) if x == x else -1 for x in ref_v[b]] for b in range(B)])
            assert np.array_equal(y, ref_y)
    img.set_white_flag()                                # utils_supersdr.py:875-877
    wf_data[:, 0, :] = 255
    rgb, data = img.image(want_rgb=True, want_data=True)
    assert np.array_equal(data, wf_data, equal_nan=True) and np.all(rgb[:, 0] == 255)
    img.close()
