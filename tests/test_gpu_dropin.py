"""GPU: the drop-in classes keep the reference's call surface and produce what the
reference pipeline would produce from the same server-side lines / PCM."""
import os
import types

import numpy as np
import pytest

from oracle import c_oracle, tier_p, tier_u

pytestmark = pytest.mark.gpu


import fake_kiwi
import fake_ref


class IQSource:
    """Headless IQ source for the FFT path of the waterfall mixin (``iq_source.read_wf_frame``)."""

    def __init__(self, wf_frames):
        self.wf = list(wf_frames)

    def read_wf_frame(self):
        return self.wf.pop(0) if self.wf else None


def test_waterfall_dropin_wire_frames(ssdr):
    """The bound kiwi_waterfall fed W/F WIRE FRAMES (utils_supersdr.py:782-784) by a fake Kiwi: the product's own header
    parse, the GPU mean + spectrum_db2col, the ring image -- against the reference's arithmetic on the same lines."""
    K = ssdr.bind(fake_ref)
    disp = types.SimpleNamespace(DISPLAY_WIDTH=1024, WF_HEIGHT=40)
    rng = np.random.default_rng(3)
    lines = np.clip(rng.normal(110, 9, (24, 1024)), 0, 255).astype(np.uint8)
    msgs = [fake_kiwi.wf_frame(lines[i], x_bin=7, zoom=3, seq=i) for i in range(24)]
    msgs.insert(5, b"MSG some=thing else=1")                  # a non-W/F message in between is skipped (utils_supersdr.py:782)
    stream = fake_kiwi.FakeKiwiStream(msgs)
    wf = K.kiwi_waterfall(stream, 3, disp)
    assert isinstance(wf, fake_ref.kiwi_waterfall) and wf.wf_data.shape == (40, 1024) and not wf.wf_data.any()
    st = tier_p.ColourState()
    st.zoom = 3
    rows = []
    wf.averaging_n = 1
    for it in range(8):                                       # single lines (the 6th receive is the MSG: row recoloured)
        assert wf.run_once()
        if it == 5:
            continue
        k = it if it < 5 else it - 1
        spec, col, px = tier_p.waterfall_line(lines[k], st)
        assert np.array_equal(wf.spectrum, spec) and np.array_equal(wf.wf_color, col) and np.array_equal(wf.wf_pixels, px)
        assert np.float32(wf.wf_min_db) == st.wf_min_db and np.float32(wf.wf_max_db) == st.wf_max_db
        assert wf.kiwi_wf_seq == k
    wf.averaging_n = 4                                        # time-binning mean, utils_supersdr.py:881-888
    for it in range(4):
        assert wf.run_once()
        blk = lines[7 + 4 * it: 11 + 4 * it]
        spec, col, px = tier_p.waterfall_line(blk, st)
        assert np.array_equal(wf.spectrum, spec) and np.array_equal(wf.wf_color, col)
        rows.append(col)
    # ring image == the reference's scroll (utils_supersdr.py:893-897): newest on top, three lines of delay
    img = wf.wf_data
    assert img.shape == (40, 1024) and img.dtype == np.float64
    # 12 lines pushed: deque(maxlen=3).appendleft + pop shows line k - 2 on top at line k (the first line is dropped)
    assert np.array_equal(img[0], rows[1].astype(np.float64)) and np.array_equal(img[1], rows[0].astype(np.float64))
    wf.set_white_flag()
    assert np.all(wf.wf_data[0] == 255)
    assert stream.sent.count("SET keepalive") == 23


def test_waterfall_dropin_iq_source(ssdr):
    """The FFT path of the same mixin: raw IQ frames instead of finished lines."""
    K = ssdr.bind(fake_ref)
    disp = types.SimpleNamespace(DISPLAY_WIDTH=1024, WF_HEIGHT=40)
    frames = tier_u.synth_iq(1024, seed=9, frames=24)
    wf = K.kiwi_waterfall(fake_kiwi.FakeKiwiStream([]), 3, disp)
    wf.iq_source = IQSource(frames)
    wf.averaging_n = 4
    st = tier_p.ColourState()
    st.zoom = 3
    rows = []
    for it in range(6):
        assert wf.run_once()
        lines = np.stack([c_oracle.wf_frame_bytes(frames[it * 4 + k]) for k in range(4)])
        spec, col, px = tier_p.waterfall_line(lines, st)       # what the reference would compute from those lines
        assert np.array_equal(wf.spectrum, spec) and np.array_equal(wf.wf_color, col)
        rows.append(col)
    # a 3-deep delay deque (maxlen 3), scrolling starts with the 4th line, newest on top (utils_supersdr.py:893-897)
    assert np.array_equal(wf.wf_data[0], rows[3].astype(np.float64)) and np.array_equal(wf.wf_data[1], rows[2])
    assert np.array_equal(wf.wf_data[2], rows[1]) and not wf.wf_data[3].any()
    assert not wf.run_once() and wf.terminate            # stream ended


def test_sound_dropin_wire_frames(ssdr):
    """The bound kiwi_sound fed SND WIRE FRAMES: (a) int16 PCM frames as a stock Kiwi sends them (utils_supersdr.py:
    1065-1072) -> header fields + GPU interpolator; (b) IQ frames (SET mod=iq, kiwi/client.py:443-454) -> unpack +
    demodulator + interpolator on the GPU."""
    K = ssdr.bind(fake_ref)
    wfobj = types.SimpleNamespace(terminate=False)
    rng = np.random.default_rng(5)
    pcm = rng.integers(-20000, 20000, (6, 512)).astype(np.int16)
    msgs = [fake_kiwi.snd_frame(pcm[i], rssi_dbm=-73.5 - i, flags=2 * (i == 2), seq=i) for i in range(6)]
    snd = K.kiwi_sound(fake_kiwi.FakeKiwiStream(msgs), "USB", 30, 3000, wfobj, 4)
    st = tier_p.InterpState()
    snd.volume, snd.audio_balance = 80, 0.25
    snd.audio_rec.recording_flag = True
    for i in range(6):
        got = snd.process_audio_stream()
        assert got.dtype == np.int16 and np.array_equal(got, pcm[i])
        assert abs(snd.rssi - (-73.5 - i)) < 1e-9 and snd.adc_overflow_flag == (i == 2)
        snd.audio_buffer.put(got)
        out = np.zeros((2048, 2), np.int16)
        snd.play_buffer(out, 2048, None, None)
        mono, o2 = tier_p.play_buffer(pcm[i], st, 80, 0.25)
        assert np.abs(out.astype(int) - o2.astype(int)).max() <= 1
        assert np.abs(snd.audio_rec.audio_buffer[-1].astype(int) - mono.astype(np.int16).astype(int)).max() <= 1
    with pytest.raises(EOFError):
        snd.process_audio_stream()
    assert snd.terminate and wfobj.terminate
    # (b) IQ frames
    n_frames = 8
    iq = tier_u.synth_demod_iq("usb", 512 * n_frames, seed=2)
    iq = (np.rint(iq.real) + 1j * np.rint(iq.imag)).astype(np.complex64)          # the wire carries int16 counts
    msgs = [fake_kiwi.iq_frame(iq[i * 512:(i + 1) * 512], seq=i, gpssec=i) for i in range(n_frames)]
    wfobj = types.SimpleNamespace(terminate=False)
    snd = K.kiwi_sound(fake_kiwi.FakeKiwiStream(msgs), "USB", 30, 3000, wfobj, 4)
    snd.iq_mode = True
    ref, rssi = tier_u.demod(iq, tier_u.DemodParams("usb", 30, 3000), tier_u.DemodState())
    for k in range(n_frames):
        got = snd.process_audio_stream()
        assert got.dtype == np.int16 and got.shape == (512,)
        assert np.abs(got.astype(int) - tier_u.pcm_to_i16(ref[k * 512:(k + 1) * 512]).astype(int)).max() <= 1
        assert abs(snd.rssi - rssi[k]) < 1e-3
    # the reference's SET senders still run (control plane inherited) and the kernel parameters follow them
    snd.radio_mode, snd.lc, snd.hc, snd.decay = "CW", 400, 800, 1000
    snd.set_mode_freq_pb()
    snd.set_agc_params()
    assert snd.stream.sent[-2].startswith("SET mod=cw low_cut=400 high_cut=800") and snd.stream.sent[-1].startswith("SET agc=1")
    # TX mute, utils_supersdr.py:1141-1147
    snd.rssi = -10
    snd.audio_buffer.put(np.full(512, 1000, np.int16))
    out = np.ones((2048, 2), np.int16)
    snd.play_buffer(out, 2048, None, None)
    assert snd.mute_counter == 15 and not out.any()


def test_concurrent_threads_separate_handles(ssdr):
    """The reference calls its classes from three threads (W/F thread supersdr.py:121-122, SND thread utils_supersdr.py:
    1198-1199, PortAudio callback :1211-1213); ctypes drops the GIL, so the shim must be re-entrant across handles and
    its error slot thread-local."""
    import threading
    iq = tier_u.synth_batch(3, 2, 4096, seed=5)
    x = tier_u.synth_demod_iq("usb", 512 * 8, seed=2)[None]
    pcm = np.random.default_rng(1).integers(-20000, 20000, (1, 512)).astype(np.int16)
    wf, dm, ib = ssdr.WaterfallBank(4096, 3, 2), ssdr.DemodBank(1, 512 * 8), ssdr.InterpBank(1, 4, max_samples=512)
    ref_px = wf.process(iq)["pixels"].copy()
    ref_pcm = dm.process(x)["pcm_i16"].copy()
    ref_st = ib.process(pcm, 80, 0.1).copy()
    errors, results = [], {"wf": [], "dm": [], "ib": []}

    def run(kind):
        try:
            for _ in range(20):
                if kind == "wf":
                    results[kind].append(np.array_equal(wf.process(iq)["pixels"], ref_px))
                elif kind == "dm":
                    dm.reset()
                    results[kind].append(np.array_equal(dm.process(x)["pcm_i16"], ref_pcm))
                else:
                    ib.reset()
                    results[kind].append(np.array_equal(ib.process(pcm, 80, 0.1), ref_st))
                    try:                                     # an error raised in this thread ...
                        ib.process(np.zeros((1, 4096), np.int16))
                    except ssdr.SsdrError as e:
                        assert "outside" in str(e)           # ... carries this thread's message
        except Exception as e:                               # noqa: BLE001
            errors.append((kind, repr(e)))

    threads = [threading.Thread(target=run, args=(k,)) for k in ("wf", "dm", "ib")]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    assert all(len(v) == 20 and all(v) for v in results.values())
    wf.close(); dm.close(); ib.close()


def test_peer_ingest_across_processes(ssdr):
    """Root-ingest without a staging copy: another process maps this process's device buffer (CUDA IPC,
    ssdr_ipc_export / ssdr_ipc_open) and its waterfall kernel reads the IQ in place; the rows equal the local run."""
    import subprocess
    import sys
    from supersdr_b200 import sharding
    B, n, N = 6, 2, 16384
    iq = ssdr.DeviceBuffer(B * n * N * 8)
    px = ssdr.DeviceBuffer(B * N)
    ssdr._lib.check(ssdr.lib.ssdr_synth_iq_dev(iq.ptr, ssdr.SSDR_IQ_CF32, B, n, N, 4321))
    bank = ssdr.WaterfallBank(N, B, n)
    bank.process_dev(iq.ptr, ssdr.SSDR_IQ_CF32, px.ptr)
    bank.sync()
    want = px.download(np.uint8, (B, N))
    handle = sharding.export_device_buffer(iq.ptr.value)
    child = (
        "import sys, numpy as np\n"
        "import supersdr_b200 as S\n"
        "from supersdr_b200 import sharding\n"
        "S.init(0)\n"
        "B, n, N = %d, %d, %d\n"
        "p = sharding.open_peer_buffer(bytes.fromhex(sys.argv[1]))\n"
        "px = S.DeviceBuffer(B * N)\n"
        "bank = S.WaterfallBank(N, B, n)\n"
        "bank.set_remote_input(True)\n"
        "bank.process_dev(p, S.SSDR_IQ_CF32, px.ptr); bank.sync()\n"
        "print('CHECKSUM', int(px.download(np.uint8, (B, N)).astype(np.uint64).sum()), flush=True)\n"
        "sys.stdout.buffer.write(px.download(np.uint8, (B, N)).tobytes())\n"
        "sharding.close_peer_buffer(p)\n" % (B, n, N))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", child, handle.hex()], cwd=root, capture_output=True, timeout=300)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    head, _, rest = out.stdout.partition(b"\n")
    assert head.startswith(b"CHECKSUM")
    got = np.frombuffer(rest[:B * N], np.uint8).reshape(B, N)
    assert np.array_equal(got, want)
    bank.close(); iq.free(); px.free()


def test_handles_work_from_other_threads_on_any_device(ssdr):
    """the CUDA device is a per-thread setting.  A handle created in the main thread is used from a fresh
    thread (the reference runs kiwi_waterfall.run, kiwi_sound.run and the PortAudio callback on their own threads,
    utils_supersdr.py:879,1106,1150) -- on device 0 and, when the box has a second GPU, on device 1, where the new
    thread's default device (0) is the wrong one.  Stateless entry points follow the selected device as well."""
    import threading
    devices = [0]
    try:
        ssdr._lib.check(ssdr.lib.ssdr_init(1))
        devices.append(1)
    except ssdr.SsdrError:
        pass
    iq = tier_u.synth_batch(3, 2, 1024, seed=77)
    ref = c_oracle.wf_rows(iq)["pixels"]
    x = np.random.default_rng(1).normal(size=400)
    try:
        for dev in devices:
            ssdr._lib.check(ssdr.lib.ssdr_init(dev))
            bank = ssdr.WaterfallBank(1024, 3, 2)             # lives on `dev`
            db = ssdr.DemodBank(2, 1024)
            got, errs = {}, []

            def worker():
                try:                                          # this thread never called ssdr_init: runtime default = device 0
                    got["px"] = bank.process(iq)["pixels"]
                    got["pcm"] = db.process(np.zeros((2, 1024), np.complex64))["pcm_f32"]
                    got["fir"] = ssdr.filtering(6000, 48000).lowpass(x)
                except Exception as e:                        # noqa: BLE001
                    errs.append(e)
            t = threading.Thread(target=worker)
            t.start(); t.join()
            assert not errs, (dev, errs)
            assert np.array_equal(got["px"], ref), dev
            assert np.all(got["pcm"] == 0)
            assert np.allclose(got["fir"], np.convolve(x, tier_p.fir_design(6000, 48000), "valid"), rtol=0, atol=1e-12)
            bank.close(); db.close()
    finally:
        ssdr._lib.check(ssdr.lib.ssdr_init(0))


def test_nccl_scatter_gather_two_gpus(ssdr, tmp_path):
    """ssdr_nccl_scatter / ssdr_nccl_gather (grouped ncclSend/ncclRecv through the C ABI, no PyTorch): rank 0 scatters a
    batch from its HBM, every rank runs the waterfall kernel on its shard, the pixel rows are gathered on rank 0 and
    equal a single-GPU run.  Needs two GPUs (skipped on the one-GPU boxes)."""
    import subprocess
    import sys
    import textwrap
    try:
        ssdr._lib.check(ssdr.lib.ssdr_init(1))
    except ssdr.SsdrError:
        pytest.skip("needs two GPUs")
    finally:
        ssdr._lib.check(ssdr.lib.ssdr_init(0))
    if not ssdr.lib.ssdr_nccl_available():
        pytest.skip("NCCL not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w.py"
    script.write_text("import os, sys\nsys.path.insert(0, %r)\n" % root + textwrap.dedent("""
        import numpy as np
        import supersdr_b200 as S
        from supersdr_b200 import sharding
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        S.init(rank)
        comm = sharding.Comm(rank, world)
        total, n, N = 7, 2, 1024
        per = n * N * 8
        first, count = sharding.channel_shard(total, rank, world)
        root_buf = None
        if rank == 0:
            root_buf = S.DeviceBuffer(total * per)
            S._lib.check(S.lib.ssdr_synth_iq_dev(root_buf.ptr, S.SSDR_IQ_CF32, total, n, N, 5))
        mine = S.DeviceBuffer(count * per)
        got = comm.scatter_from_root(root_buf.ptr.value if rank == 0 else None, total, per, mine.ptr.value)
        assert got == count * per
        bank = S.WaterfallBank(N, count, n)
        px = S.DeviceBuffer(count * N)
        bank.process_dev(mine.ptr, S.SSDR_IQ_CF32, px.ptr)
        bank.sync()
        allpx = S.DeviceBuffer(total * N) if rank == 0 else None
        comm.gather_rows_to_root(px.ptr.value, total, N, allpx.ptr.value if rank == 0 else None)
        assert comm.max_over_ranks(float(rank)) == float(world - 1)
        comm.barrier()
        if rank == 0:
            one = S.WaterfallBank(N, total, n)
            ref = S.DeviceBuffer(total * N)
            one.process_dev(root_buf.ptr, S.SSDR_IQ_CF32, ref.ptr)
            one.sync()
            assert np.array_equal(allpx.download(np.uint8, (total, N)), ref.download(np.uint8, (total, N)))
            print("OK")
        comm.close()
    """))
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]
