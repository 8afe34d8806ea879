"""ctypes binding of libssdr_b200.so (include/ssdr_b200.h).

There is no CPU fallback: if the shared library has not been built, importing this module raises;
if no CUDA device is usable, every compute call raises ``SsdrError``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SSDR_B200_LIB: developer override used by scripts/exp_variants.sh to time experimental builds of the same ABI
LIB_PATH = os.environ.get("SSDR_B200_LIB") or os.path.join(_HERE, "libssdr_b200.so")

SSDR_IQ_CF32, SSDR_IQ_S16BE = 0, 1
SSDR_DEMOD_ENGINE_FFMA, SSDR_DEMOD_ENGINE_TCGEN05, SSDR_DEMOD_ENGINE_AUTO = 0, 1, 2
MODE_AM, MODE_USB, MODE_LSB, MODE_CW, MODE_NBFM = 0, 1, 2, 3, 4
FS = 32768.0
WF_CAL_DB = -10.0
KIWI_RATE = 12000
FRAME = 512
FIR_TAPS = 127
INTERP_TAPS_MAX = 64


class SsdrError(RuntimeError):
    pass


class WfDisplay(C.Structure):
    _fields_ = [("zoom", C.c_int32), ("auto_scale", C.c_int32), ("delta_low_db", C.c_int32),
                ("delta_high_db", C.c_int32), ("low_clip_db", C.c_float), ("dynamic_range", C.c_float)]


class WfScalars(C.Structure):
    _fields_ = [("low_clip_db", C.c_float), ("high_clip_db", C.c_float), ("dynamic_range", C.c_float),
                ("wf_min_db", C.c_float), ("wf_max_db", C.c_float)]


class DemodParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("low_cut_hz", C.c_float), ("high_cut_hz", C.c_float),
                ("freq_offset_hz", C.c_float), ("agc_on", C.c_int32), ("agc_hang", C.c_int32),
                ("agc_thresh_dbm", C.c_float), ("agc_slope_db", C.c_float), ("agc_decay_ms", C.c_float),
                ("agc_man_gain_db", C.c_float), ("taps", C.c_float * FIR_TAPS)]


SCALARS_DTYPE = np.dtype([("low_clip_db", "<f4"), ("high_clip_db", "<f4"), ("dynamic_range", "<f4"),
                          ("wf_min_db", "<f4"), ("wf_max_db", "<f4")])

_vp, _i, _f, _d, _sz, _u32 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_uint32
_pvp = C.POINTER(C.c_void_p)

_PROTOS = {
    "ssdr_abi_version": (C.c_int, []),
    "ssdr_last_error": (C.c_char_p, []),
    "ssdr_init": (_i, [_i]),
    "ssdr_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz), C.c_char_p, _i]),
    "ssdr_device_pci_bus_id": (_i, [C.c_char_p, _i]),
    "ssdr_launch_count": (C.c_uint64, []),
    "ssdr_ipc_export": (_i, [_vp, _vp]),
    "ssdr_ipc_open": (_i, [_vp, _pvp]),
    "ssdr_ipc_close": (_i, [_vp]),
    "ssdr_dev_alloc": (_i, [_pvp, _sz]),
    "ssdr_dev_free": (_i, [_vp]),
    "ssdr_host_alloc": (_i, [_pvp, _sz]),
    "ssdr_host_free": (_i, [_vp]),
    "ssdr_memcpy_h2d": (_i, [_vp, _vp, _sz]),
    "ssdr_memcpy_d2h": (_i, [_vp, _vp, _sz]),
    "ssdr_dev_memset": (_i, [_vp, _i, _sz]),
    "ssdr_device_sync": (_i, []),
    "ssdr_synth_iq_dev": (_i, [_vp, _i, _i, _i, _i, _u32]),
    "ssdr_nccl_available": (_i, []),
    "ssdr_nccl_unique_id": (_i, [_vp]),
    "ssdr_nccl_init": (_i, [_pvp, _vp, _i, _i]),
    "ssdr_nccl_destroy": (_i, [_vp]),
    "ssdr_nccl_scatter": (_i, [_vp, _vp, _vp, _vp, _vp, _i]),
    "ssdr_nccl_gather": (_i, [_vp, _vp, _vp, _vp, _vp, _i]),
    "ssdr_nccl_allreduce_max_f64": (_i, [_vp, C.POINTER(_d)]),
    "ssdr_nccl_barrier": (_i, [_vp]),
    "ssdr_nccl_sync": (_i, [_vp]),
    "ssdr_wf_create": (_i, [_pvp, _i, _i, _i, _i, _d, _i, _f]),
    "ssdr_wf_destroy": (_i, [_vp]),
    "ssdr_wf_set_display": (_i, [_vp, _i, _i, C.POINTER(WfDisplay)]),
    "ssdr_wf_set_remote_input": (_i, [_vp, _i]),
    "ssdr_wf_get_tables": (_i, [_vp, _vp, _vp, C.POINTER(_i)]),
    "ssdr_wf_get_window": (_i, [_vp, _vp]),
    "ssdr_wf_process": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "ssdr_wf_process_dev": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "ssdr_wf_colorrow_u8": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "ssdr_wf_colorrow_u8_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "ssdr_wf_sync": (_i, [_vp]),
    "ssdr_wf_time_dev": (_i, [_vp, _vp, _i, _vp, _i, C.POINTER(_f)]),
    "ssdr_demod_create": (_i, [_pvp, _i, _i]),
    "ssdr_demod_destroy": (_i, [_vp]),
    "ssdr_demod_set": (_i, [_vp, _i, _i, C.POINTER(DemodParams)]),
    "ssdr_demod_reset": (_i, [_vp]),
    "ssdr_demod_set_engine": (_i, [_vp, _i]),
    "ssdr_demod_plan": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_f)]),
    "ssdr_demod_process": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "ssdr_demod_process_dev": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "ssdr_demod_sync": (_i, [_vp]),
    "ssdr_demod_time_dev": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, C.POINTER(_f)]),
    "ssdr_interp_create": (_i, [_pvp, _i, _i, _vp, _i, _i]),
    "ssdr_interp_destroy": (_i, [_vp]),
    "ssdr_interp_reset": (_i, [_vp]),
    "ssdr_interp_process": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "ssdr_interp_process_dev": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "ssdr_interp_sync": (_i, [_vp]),
    "ssdr_wf_image_create": (_i, [_pvp, _i, _i, _i, _vp]),
    "ssdr_wf_image_destroy": (_i, [_vp]),
    "ssdr_wf_image_push": (_i, [_vp, _vp]),
    "ssdr_wf_image_push_dev": (_i, [_vp, _vp]),
    "ssdr_wf_image_white": (_i, [_vp]),
    "ssdr_wf_image_get": (_i, [_vp, _vp, _vp]),
    "ssdr_wf_image_trace": (_i, [_vp, _i, _i, _vp, _vp]),
    "ssdr_adpcm_decode": (_i, [_vp, _i, _i, _vp, _vp]),
    "ssdr_resample_line": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "ssdr_fir_valid_f64": (_i, [_vp, _sz, _vp, _i, _vp]),
    "ssdr_unpack_iq_s16be": (_i, [_vp, _vp, _sz]),
    "ssdr_unpack_iq_s16be_dev": (_i, [_vp, _vp, _sz]),
}
EXPORTS = sorted(_PROTOS)

if not os.path.isfile(LIB_PATH):
    raise ImportError("libssdr_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "or `make -C supersdr_b200/csrc`); supersdr_b200 has no CPU fallback")

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)          # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args

_initialised = False


def last_error():
    return lib.ssdr_last_error().decode("utf-8", "replace")


def check(rc):
    if rc < 0:
        raise SsdrError("libssdr_b200 error %d: %s" % (rc, last_error()))
    return rc


def init(device=None):
    """Select the CUDA device for this process (default: LOCAL_RANK or 0). Idempotent."""
    global _initialised
    if device is None:
        if _initialised:
            return
        device = int(os.environ.get("LOCAL_RANK", "0"))
    check(lib.ssdr_init(int(device)))
    _initialised = True


def numa_bind():
    """Pin this process's threads to the CPUs of the selected GPU's NUMA node, so that pinned host buffers allocated
    afterwards (first touch) and the threads that fill them sit next to the GPU's PCIe root complex.  With several
    ranks on one host this keeps each rank's H2D stream off the inter-socket link.  Returns the node (or None when the
    platform does not say / has one node)."""
    buf = C.create_string_buffer(32)
    check(lib.ssdr_device_pci_bus_id(buf, 32))
    bus = buf.value.decode().lower()
    try:
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError):
        return None


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p)


class DeviceBuffer:
    """Device memory owned by Python (bench / resident-data callers)."""

    def __init__(self, nbytes):
        init()
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib.ssdr_dev_alloc(C.byref(p), self.nbytes))
        self.ptr = p

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(lib.ssdr_memcpy_h2d(self.ptr, ptr(arr), arr.nbytes))
        return self

    def download(self, dtype, shape, offset_bytes=0):
        out = np.empty(shape, dtype=dtype)
        assert offset_bytes + out.nbytes <= self.nbytes
        check(lib.ssdr_memcpy_d2h(ptr(out), C.c_void_p(self.ptr.value + offset_bytes), out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib.ssdr_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """A numpy array backed by pinned host memory (fast, truly asynchronous H2D/D2H)."""

    def __init__(self, shape, dtype):
        init()
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib.ssdr_host_alloc(C.byref(p), max(self.nbytes, 1)))
        self._p = p
        buf = (C.c_char * self.nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self._p:
            self.array = None
            lib.ssdr_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
