import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
NSTEP = int(sys.argv[2]) if len(sys.argv) > 2 else 60
r, vr, vt, m = models.plummer(N, seed=77)
c = Cluster(N, n_star_dim=N + 64)
c.load(r, vr, vt, m * N)
c.calc_potential_new(); c.compute_energy(); c.upload()
bad = None
for step in range(NSTEP):
    c.evolve(t_end=1e9, max_steps=1, seed=11, tidally_strip=0, GAMMA=0.1)
    a, sg = c.dev_calc_sigma_r(20)
    st = dict(c.sigma_stats)
    n = int(c.End[0] - c.Start[0] + 1)
    print(step, "N_MAX", c.clus.N_MAX, "N_MAX_NEW", c.clus.N_MAX_NEW, "n_local", n, st, "nan sigma", int(np.isnan(sg[1:]).sum()), flush=True)
    if (st["broken"] or st["rerun"]) and bad is None and N <= 1_000_000:
        bad = step
        c.download()
        vr_, vt_ = c.field("vr")[1:n + 1], c.field("vt")[1:n + 1]
        t = c.star_m[1:n + 1] / N * (vr_ ** 2 + vt_ ** 2)
        w = np.convolve(t, np.ones(40), mode="valid")
        print("  terms: min", t.min(), "zeros", int((t == 0).sum()), "nan", int(np.isnan(t).sum()), "window min", w.min(), "at", int(np.argmin(w)), "window max", w.max(), "prefix", t.sum())
        print("  smallest terms at", np.argsort(t)[:10], t[np.argsort(t)[:10]])
        print("  tail r", c.star_r[n - 3:n + 2], "vr tail", vr_[-5:], "vt tail", vt_[-5:])
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gpurun_out', 'sigma_bad_state.npz'), r=c.star_r[:n + 2], m=c.star_m[:n + 2], vr=vr_, vt=vt_, N=N)
        break
