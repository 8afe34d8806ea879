"""CPU: the host-side work plan of the tcgen05 demodulator engine (ssdr_demod_plan, no device needed): channels grouped
by bitwise-equal taps into quads (one tensor-core tile each), quads of one filter into rounds, dearest detectors first,
the last partial wave spread over narrower rounds."""
import numpy as np
import pytest

import supersdr_b200 as S
from supersdr_b200.sound import MODE_IDS, demod_params, demod_plan


def _check_plan(params, n_sm):
    qc, qf, tiles, fill = demod_plan(params, n_sm)
    B = len(params)
    assert qc.shape[0] % tiles == 0
    used = qc[qc >= 0]
    assert sorted(used.tolist()) == list(range(B))                      # every channel exactly once
    taps = [tuple(p.taps) for p in params]
    for q, f in zip(qc, qf):
        chans = q[q >= 0]
        assert len(set(taps[c] for c in chans)) <= 1                     # one filter per quad (the tile's B operand)
        k = int((q >= 0).sum())
        assert np.all(q[:k] >= 0) and np.all(q[k:] < 0)                 # slots filled in order
    rounds = qc.reshape(-1, tiles, 4)
    rf = qf.reshape(-1, tiles)
    for r, f in zip(rounds, rf):
        assert r[0, 0] >= 0                                             # a round starts with a real quad
        assert len(set(f.tolist())) == 1                                # one filter id per round
        real = [taps[q[0]] for q in r if q[0] >= 0]
        assert len(set(real)) == 1
    nonempty = int((qc[:, 0] >= 0).sum())
    assert fill == pytest.approx(B / (4.0 * nonempty))
    return qc, qf, tiles, fill


def test_plan_uniform_batch_splits_the_tail_wave():
    B, n_sm = 4096, 148
    qc, qf, tiles, fill = _check_plan([demod_params("usb", 300, 2700)] * B, n_sm)
    assert fill == 1.0 and len(set(qf.tolist())) == 1
    rounds = qc.reshape(-1, tiles, 4)
    per_round = (rounds[:, :, 0] >= 0).sum(1)
    n_quads = B // 4
    full = (n_quads // tiles) // n_sm * n_sm                              # whole waves of full rounds ...
    assert np.all(per_round[:full] == tiles)
    tail_quads = n_quads - full * tiles                                  # ... then the rest spread over (nearly) all SMs
    per = -(-tail_quads // n_sm)
    assert per < tiles and np.all(per_round[full:] <= per) and per_round[full:].sum() == tail_quads
    assert len(per_round) - full == -(-tail_quads // per)


def test_plan_mixed_modes_cost_order_and_padding():
    modes = ["am", "lsb", "usb", "cw", "nbfm"]
    B = 203
    params = [demod_params(modes[c % 5]) for c in range(B)]
    qc, qf, tiles, fill = _check_plan(params, 8)
    assert len(set(qf.tolist())) == 3                                    # AM / NBFM share +-6 kHz, USB / LSB share, CW
    first_mode = [params[r[0, 0]].mode for r in qc.reshape(-1, tiles, 4)]
    cost = [2 if m == MODE_IDS["am"] else 0 if m == MODE_IDS["nbfm"] else 1 for m in first_mode]      # measured per-mode cost, round 2
    full = [c for c, r in zip(cost, qc.reshape(-1, tiles, 4)) if (r[:, 0] >= 0).all()]
    assert full == sorted(full, reverse=True)                            # dearest detectors first
    assert 0.5 < fill <= 1.0


def test_plan_distinct_filters_degenerate_to_single_channel_quads():
    params = [demod_params("usb", 300, 2700 + 10 * c) for c in range(9)]
    qc, qf, tiles, fill = _check_plan(params, 148)
    assert fill == pytest.approx(0.25)                                   # AUTO falls back to the FFMA engine below 0.5
    assert int((qc >= 0).sum()) == 9 and len(set(qf[qc[:, 0] >= 0].tolist())) == 9


def test_plan_many_filters_never_adds_a_wave():
    """A round cannot mix filters: with 64 filters x 64 channels the narrow tail rounds (3 quads each) would outnumber the
    SMs (162 on 148) and add a mostly idle wave; the planner then keeps the full rounds (scripts/demod_hetero.py)."""
    B, n_sm = 4096, 148
    uniq = [demod_params("usb", 300, 2700 + g) for g in range(64)]
    qc, qf, tiles, fill = _check_plan([uniq[c % 64] for c in range(B)], n_sm)
    per_round = (qc.reshape(-1, tiles, 4)[:, :, 0] >= 0).sum(1)
    full = int((per_round == tiles).sum())
    assert len(per_round) - (full // n_sm) * n_sm <= n_sm                # whatever follows the whole waves fits one wave
    assert len(per_round) == 256 and full == 256
