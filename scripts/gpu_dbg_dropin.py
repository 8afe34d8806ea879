import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supersdr_b200 as S
from oracle import c_oracle, tier_p, tier_u
S.init()
frames = tier_u.synth_iq(1024, seed=9, frames=24)
T = c_oracle.thresholds(1024, -10.0)
for (B, n) in ((24, 1), (1, 1), (6, 4), (1, 4)):
    bank = S.WaterfallBank(1024, B, n)
    tot = 0
    for s in range(0, 24, B * n):
        iq = frames[s:s + B * n].reshape(B, n, 1024)
        res = bank.process(iq)
        ref = c_oracle.wf_rows(iq)
        d = res["spectrum"] != ref["spectrum"]
        tot += d.sum()
        if d.sum() and n == 1:
            for (b, k) in np.argwhere(d)[:4]:
                fr = iq[b, 0]
                by, spec = c_oracle.wf_frame_bytes(fr, want_spectrum=True)
                kk = (k + 512) % 1024
                X = spec[kk]
                P = np.float32(np.float32(X.imag) * np.float32(X.imag))
                P = np.float32(np.float64(np.float32(X.real)) * np.float64(np.float32(X.real)) + np.float64(P))
                kb = int(by[k])
                print("  B,n", B, n, "frame", s + b, "bin", k, "gpu", res["spectrum"][b, k], "ref", ref["spectrum"][b, k], "P", P, "T[k]", T[kb], "T[k+1]", T[kb + 1],
                      "rel dist", (P - T[kb]) / T[kb], (T[kb + 1] - P) / T[kb + 1])
    print("B", B, "n", n, "mismatches", tot)
    bank.close()
