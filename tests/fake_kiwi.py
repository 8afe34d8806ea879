"""TEST INFRASTRUCTURE: a fake KiwiSDR that speaks the wire formats the reference pins.

``FakeKiwiStream`` stands in for the mod_pywebsocket ``Stream`` the reference classes hold (``wf_stream`` / ``stream``):
``receive_message()`` returns the next message as BYTES in the exact frame layouts of

  * W/F frame:  b"W/F" + 1 byte + <III (x_bin, flags|zoom, seq) + uint8[WF_BINS]      utils_supersdr.py:782-784, kiwi/client.py:367-368,470-472
  * SND frame:  b"SND" + <B flags + <I seq + >H smeter + >h[512]                      utils_supersdr.py:1065-1072, kiwi/client.py:385-388
  * IQ frame:   the same 10-byte prefix + <BBII GPS header + >h interleaved I,Q       kiwi/client.py:443-454

and ``send_message()`` records the ``SET ...`` commands the client sends (utils_supersdr.py:741-742,976-980,1022-1029).
"""
import struct

import numpy as np


def wf_frame(line_u8, x_bin=0, zoom=0, seq=0):
    line = np.asarray(line_u8, dtype=np.uint8)
    return b"W/F" + b"\x00" + struct.pack("<III", x_bin, zoom, seq) + line.tobytes()


def snd_frame(pcm_i16, rssi_dbm=-80.0, flags=0, seq=0):
    smeter = int(round((rssi_dbm + 127.0) * 10.0))
    return b"SND" + struct.pack("<BI", flags, seq) + struct.pack(">H", smeter) + np.asarray(pcm_i16).astype(">i2").tobytes()


def iq_frame(iq_c64, rssi_dbm=-80.0, flags=0, seq=0, gpssec=0, gpsnsec=0):
    iq = np.asarray(iq_c64)
    inter = np.empty(2 * iq.size, ">i2")
    inter[0::2] = np.rint(iq.real)
    inter[1::2] = np.rint(iq.imag)
    smeter = int(round((rssi_dbm + 127.0) * 10.0))
    return (b"SND" + struct.pack("<BI", flags, seq) + struct.pack(">H", smeter) + struct.pack("<BBII", 0, 0, gpssec, gpsnsec)
            + inter.tobytes())


class FakeKiwiStream:
    def __init__(self, messages):
        self.messages = list(messages)
        self.sent = []

    def receive_message(self):
        return self.messages.pop(0) if self.messages else None

    def send_message(self, msg):
        self.sent.append(msg)
