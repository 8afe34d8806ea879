import numpy as np, sys
sys.path.insert(0, ".")
import supersdr_b200 as S
from oracle import c_oracle, tier_u
S.init(0)
for N in (256, 512, 1024, 2048, 4096, 8192, 16384):
    for window in (False, True):
        iq = tier_u.synth_batch(2, 1, N, seed=N)
        bank = S.WaterfallBank(N, 2, 1, window=window)
        res = bank.process(iq)
        ref = c_oracle.wf_rows(iq, window=window)
        d = res["spectrum"].astype(int) - ref["spectrum"].astype(int)
        print(N, window, "mismatch bins", int((d != 0).sum()), "max |d|", int(np.abs(d).max()), "first", np.nonzero(d[0])[0][:6])
        bank.close()
