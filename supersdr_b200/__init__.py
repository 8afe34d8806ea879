"""supersdr_b200 -- B200-native (sm_100a) IQ-sample DSP behind SuperSDR's waterfall and audio classes.

Python host code over the C ABI of ``libssdr_b200.so`` (include/ssdr_b200.h); hand-written CUDA
kernels, no PyTorch / Triton / cuFFT on the compute path and no CPU fallback.  Importing the package
requires the built shared library; computing requires a B200.
"""
from ._lib import (SsdrError, init, last_error, DeviceBuffer, PinnedArray, lib, EXPORTS, LIB_PATH,
                   SSDR_IQ_CF32, SSDR_IQ_S16BE, FS, WF_CAL_DB, KIWI_RATE, FRAME, FIR_TAPS)
from .waterfall import WaterfallBank, WaterfallImage, percentile_index, create_cm, palette_u8
from .sound import (DemodBank, InterpBank, ResampleLine, ImaAdpcmDecoder, filtering, demod_params, demod_plan,
                    design_lowpass, default_passband, unpack_iq)
from .dropin import bind, WaterfallHotPath, SoundHotPath, parse_wf_frame, parse_snd_frame, HOT_PATH_OVERRIDES
from .wavreader import KiwiIQWavReader, KiwiIQWavError, WavIQSource, read_kiwi_iq_wav

__all__ = ["ImaAdpcmDecoder", "WaterfallImage", "create_cm", "palette_u8", "SsdrError", "init", "last_error", "DeviceBuffer", "PinnedArray", "WaterfallBank", "bind", "WaterfallHotPath", "SoundHotPath", "parse_wf_frame", "parse_snd_frame",
           "percentile_index", "DemodBank", "InterpBank", "ResampleLine", "filtering",
           "demod_params", "demod_plan", "design_lowpass", "default_passband", "unpack_iq", "KiwiIQWavReader", "KiwiIQWavError",
           "WavIQSource", "read_kiwi_iq_wav"]
