"""TEST INFRASTRUCTURE ONLY -- generates the committed fixtures under tests/golden/.

Tier-P fixtures (``tier_p_*.npz``) are OUTPUTS OF THE UNMODIFIED REFERENCE: its functions are imported
headless from /root/reference (oracle/ref_import.py) and run on seeded inputs; this only works in
the build container, which is why the vectors are committed.  Tier-U fixtures (``tier_u_*.npz``)
come from the builder-defined oracles (oracle/c/ssdr_oracle.c, oracle/tier_u.py) -- parity unpinned.

    python -m oracle.make_golden
"""
import os
import queue
import sys
from collections import deque

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_import, tier_u, c_oracle   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_waterfall_line(m, lines_u8, zoom, auto, dlow, dhigh, low_clip=None, dyn=None):
    """Run the reference's own averaging (utils:881-886) + spectrum_db2col (utils:787-813)."""
    wf = m.kiwi_waterfall.__new__(m.kiwi_waterfall)
    wf.zoom, wf.wf_auto_scaling = zoom, auto
    wf.delta_low_db, wf.delta_high_db = dlow, dhigh
    wf.dynamic_range = wf.MIN_DYN_RANGE if dyn is None else dyn
    if low_clip is not None:
        wf.low_clip_db = low_clip
    n = lines_u8.shape[0]
    if n > 1:
        dq = deque([], n)
        for k in range(n):
            # receive_spectrum, utils:783-784
            dq.append(np.ndarray(len(lines_u8[k].tobytes()), dtype='B', buffer=lines_u8[k].tobytes()).astype(np.float32))
        wf.spectrum = np.mean(dq, axis=0)
    else:
        wf.spectrum = lines_u8[0].astype(np.float32)
    spectrum = wf.spectrum.copy()
    wf.spectrum_db2col()
    return spectrum, wf.wf_color.copy(), np.array([wf.low_clip_db, wf.high_clip_db, wf.dynamic_range,
                                                   wf.wf_min_db, wf.wf_max_db], np.float32)


def gen_tier_p_waterfall(m):
    rng = np.random.default_rng(20261017)
    cases = []
    specs = [(1024, 1, 0), (1024, 10, 3), (1024, 100, 14), (256, 7, 1), (2048, 3, 5), (16384, 10, 2), (1024, 2, 0)]
    for W, n, zoom in specs:
        for kind in range(3):
            if kind == 0:
                lines = np.clip(rng.normal(110, 7, (n, W)), 0, 255)
                lines[:, rng.integers(0, W, 5)] += 90
            elif kind == 1:
                lines = rng.integers(0, 256, (n, W))
            else:
                lines = np.full((n, W), int(rng.integers(0, 256)))      # flat line: dyn range = 40 floor
            lines = np.clip(lines, 0, 255).astype(np.uint8)
            auto = kind != 1 or n == 1
            dlow, dhigh = (int(rng.integers(-15, 15)), int(rng.integers(-15, 15))) if kind == 0 else (0, 0)
            spec, col, sc = ref_waterfall_line(m, lines, zoom, auto, dlow, dhigh, low_clip=-120 if not auto else None)
            cases.append(dict(lines=lines, zoom=zoom, auto=auto, dlow=dlow, dhigh=dhigh, spectrum=spec, colour=col, scalars=sc))
    flat = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat["c%02d_%s" % (i, k)] = np.asarray(v)
    flat["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "tier_p_waterfall.npz"), **flat)
    return len(cases)


def gen_tier_p_audio(m):
    rng = np.random.default_rng(77)
    f = m.filtering(6000, 48000)
    snd = m.kiwi_sound.__new__(m.kiwi_sound)
    snd.kiwi_filter = f
    snd.n_tap = f.n_tap
    snd.old_buffer = np.zeros(f.n_tap - 1)
    snd.late_flag = False
    snd.audio_buffer = queue.Queue()
    snd.rssi, snd.mute_counter, snd.max_rssi_before_mute, snd.muting_delay = -90, 0, -20, 15
    snd.audio_rec = type("AR", (), {"recording_flag": False})()
    xs, vols, bals, outs = [], [], [], []
    for k in range(12):
        x = rng.integers(-32768, 32768, 512).astype(np.int16)
        if k == 3:
            x[:] = 32767                      # full-scale DC: exercises the int16 wrap at volume 150
        snd.volume = int(rng.integers(0, 16)) * 10 if k != 3 else 150
        snd.audio_balance = float(np.round(rng.uniform(-1, 1), 2))
        snd.audio_buffer.put(x)
        out = np.zeros((2048, 2), np.int16)
        snd.play_buffer(out, 2048, None, None)
        xs.append(x); vols.append(snd.volume); bals.append(snd.audio_balance); outs.append(out.copy())
    # SND frame parse, utils:1065-1072
    body = rng.integers(-32768, 32768, 512).astype(">i2").tobytes()
    msg = b"SND" + bytes([2]) + (1234).to_bytes(4, "little") + (870).to_bytes(2, "big") + body
    np.savez_compressed(os.path.join(OUT, "tier_p_audio.npz"), h=f.h, x=np.stack(xs), volume=np.array(vols),
                        balance=np.array(bals), out=np.stack(outs), snd_msg=np.frombuffer(msg, np.uint8),
                        snd_samples=np.frombuffer(body, ">i2").astype(np.int16))
    return len(xs)


def gen_tier_p_audio_resample(m):
    """kiwi_sound.play_buffer on its non-integer-ratio path (utils_supersdr.py:1125-1126): a 20.25 kHz Kiwi into a
    48 kHz sound card, resample_poly(x, 64, 27, padtype="line")[:-1]."""
    rng = np.random.default_rng(78)
    snd = m.kiwi_sound.__new__(m.kiwi_sound)
    snd.SAMPLE_RATIO = 48000 / 20250
    snd.n_low, snd.n_high = 27, 64
    snd.late_flag = False
    snd.audio_buffer = queue.Queue()
    snd.rssi, snd.mute_counter, snd.max_rssi_before_mute, snd.muting_delay = -90, 0, -20, 15
    snd.audio_rec = type("AR", (), {"recording_flag": False})()
    xs, vols, bals, outs = [], [], [], []
    t = np.arange(512)
    for k in range(8):
        x = rng.integers(-20000, 20000, 512).astype(np.int16)
        if k == 2:
            x = np.rint(15000 * np.sin(2 * np.pi * 0.01 * t) + 20 * t).astype(np.int16)      # a ramp: the "line" extension matters
        snd.volume = int(rng.integers(1, 16)) * 10
        snd.audio_balance = float(np.round(rng.uniform(-1, 1), 2))
        snd.audio_buffer.put(x)
        out = np.zeros((1213, 2), np.int16)
        snd.play_buffer(out, 1213, None, None)
        xs.append(x); vols.append(snd.volume); bals.append(snd.audio_balance); outs.append(out.copy())
    np.savez_compressed(os.path.join(OUT, "tier_p_audio_resample.npz"), x=np.stack(xs), volume=np.array(vols),
                        balance=np.array(bals), out=np.stack(outs), up=64, down=27)
    return len(xs)


def gen_palette(m):
    """display_stuff.create_cm('cutesdr') (utils_supersdr.py:1391-1412) -> tests/golden/palette_cutesdr.npz."""
    disp = m.display_stuff.__new__(m.display_stuff)
    cm = np.asarray(disp.create_cm("cutesdr"), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "palette_cutesdr.npz"), colormap=cm)
    return cm.shape


def gen_adpcm():
    """kiwi/client.py ImaAdpcmDecoder (:58-87) on random codes, streamed in three calls -> tests/golden/adpcm.npz."""
    sys.path.insert(0, ref_import.REFERENCE_DIR)
    from kiwi import client
    rng = np.random.default_rng(21)
    streams = []
    for s_ in range(4):
        data = rng.integers(0, 256, 3 * 333).astype(np.uint8)
        if s_ == 1:
            data[:] = 0x77                       # step index pinned at its upper clamp, samples saturate
        if s_ == 2:
            data[:] = 0x00                       # step index pinned at 0
        dec = client.ImaAdpcmDecoder()
        out, states = [], []
        for c in range(3):
            out.append(np.array(dec.decode(bytes(data[c * 333:(c + 1) * 333])), np.int16))
            states.append((dec.index, dec.prev))
        streams.append((data, np.concatenate(out), np.array(states, np.int32)))
    np.savez_compressed(os.path.join(OUT, "adpcm.npz"), data=np.stack([d for d, _, _ in streams]),
                        pcm=np.stack([p for _, p, _ in streams]), states=np.stack([t for _, _, t in streams]))
    return len(streams)


def write_kiwi_iq_wav(path, blocks, fs=12000, t0=1234567.25):
    """A Kiwi IQ WAV file as kiwirecorder writes it: RIFF/WAVE, 16-byte fmt chunk (PCM, 2 channels, 16 bit), then a
    10-byte 'kiwi' GNSS chunk (<BBII) before every 'data' chunk of interleaved little-endian int16 I/Q."""
    import struct
    body = b"WAVE" + b"fmt " + struct.pack("<L", 16) + struct.pack("<HHLLHH", 1, 2, fs, fs * 4, 4, 16)
    t = t0
    for b in blocks:
        sec = int(t)
        body += b"kiwi" + struct.pack("<L", 10) + struct.pack("<BBII", 1, 0, sec, int(round((t - sec) * 1e9)))
        raw = np.ascontiguousarray(b, dtype="<i2").tobytes()
        body += b"data" + struct.pack("<L", len(raw)) + raw
        t += (len(raw) // 4) / (fs * (1 + 3e-5))          # a slightly fast sample clock, as real Kiwis have
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<L", len(body)) + body)


def gen_kiwi_wav():
    """kiwi/wavreader.py run on a synthetic recording -> tests/golden/kiwi_iq.wav + kiwi_iq_wav.npz."""
    sys.path.insert(0, ref_import.REFERENCE_DIR)
    from kiwi import wavreader
    rng = np.random.default_rng(11)
    blocks = [rng.integers(-20000, 20000, (512, 2)).astype(np.int16) for _ in range(6)]
    path = os.path.join(OUT, "kiwi_iq.wav")
    write_kiwi_iq_wav(path, blocks)
    t, z = wavreader.read_kiwi_iq_wav(path)
    per = [(tt, zz) for tt, zz in wavreader.KiwiIQWavReader(path)]
    np.savez_compressed(os.path.join(OUT, "kiwi_iq_wav.npz"), t=t, z=z, n_blocks=len(per),
                        none_blocks=np.array([tt is None for tt, _ in per]), last_t=per[-1][0], last_z=per[-1][1])
    print("kiwi_iq.wav", os.path.getsize(path))


def gen_tier_u():
    # waterfall: BASELINE config 1 (single 1024-pt frame) + one 16384-pt, 2-frame channel
    x1 = tier_u.synth_iq(1024, seed=1234)
    r1 = c_oracle.wf_rows(x1[None], zoom=0)
    x2 = tier_u.synth_batch(2, 2, 16384, seed=99)
    r2 = c_oracle.wf_rows(x2, zoom=3)
    np.savez_compressed(os.path.join(OUT, "tier_u_waterfall.npz"), iq1=x1, bytes1=c_oracle.wf_frame_bytes(x1[0]),
                        pixels1=r1["pixels"], spectrum1=r1["spectrum"], iq2=x2, pixels2=r2["pixels"],
                        sums2=r2["sums"], scalars2=r2["scalars"])
    # demod: one short stream per mode
    d = {}
    for mode in tier_u.MODES:
        p = tier_u.DemodParams(mode, decay=1000 if mode == "cw" else 4000, hang=(mode == "cw"))
        x = tier_u.synth_demod_iq(mode, 512 * 8, seed=5)
        st = tier_u.DemodState()
        pcm, rssi = tier_u.demod(x, p, st)
        d["iq_" + mode], d["pcm_" + mode], d["rssi_" + mode] = x, pcm.astype(np.float32), rssi.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "tier_u_demod.npz"), **d)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    m = ref_import.load()
    print("tier_p waterfall cases:", gen_tier_p_waterfall(m))
    print("tier_p audio blocks:", gen_tier_p_audio(m))
    print("tier_p audio resample blocks:", gen_tier_p_audio_resample(m))
    gen_tier_u()
    gen_kiwi_wav()
    print("palette:", gen_palette(m))
    print("adpcm streams:", gen_adpcm())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
