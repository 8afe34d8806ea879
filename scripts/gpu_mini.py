import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supersdr_b200 as S
from oracle import c_oracle, tier_u
S.init()
for N, B, n in [(int(a), int(b), int(c)) for a, b, c in (x.split(",") for x in sys.argv[1:])]:
    iq = tier_u.synth_batch(B, n, N, seed=N + B)
    wb = S.WaterfallBank(N, B, n)
    try:
        res = wb.process(iq)
        ref = c_oracle.wf_rows(iq, threads=8)
        print(N, B, n, "spectrum", np.array_equal(res["spectrum"], ref["spectrum"]), "pixels", np.array_equal(res["pixels"], ref["pixels"]), flush=True)
    except Exception as e:
        print(N, B, n, "FAILED", e, flush=True)
        break
