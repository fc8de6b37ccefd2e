"""Developer probe: how many medoids could the triangle inequality exclude from the re-assignment
of a proposal's ambiguous frames?  Medoid j can only win frame x if d(prop, m_j) < 2 d(x, prop)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from enspara_b200 import synth
from enspara_b200.cluster import util, kcenters as kc, _ops

n, A, k = 1_000_000, 500, 1000
data = synth.device_trajectory(n, A, seed=0)
res, eng = kc.kcenters(data, "rmsd", n_clusters=k, _return_engine=True)
ctr = torch.as_tensor([int(c) for c in res.center_indices], device="cuda")
medoids = data.gather(ctr)
rs = np.random.RandomState(0)
fr_all, fr_med, sizes = [], [], []
for cid in range(0, k, 10):
    members = torch.nonzero(eng.assign == cid).view(-1)
    p = members[rs.randint(len(members))]
    prop = data.gather(p.view(1))
    cc = _ops.one_to_all_device(util.RMSD, medoids, prop)
    d = _ops.one_to_all_device(util.RMSD, data, prop)
    amb = (eng.assign == cid) & (d > eng.dist)
    if int(amb.sum()) == 0:
        continue
    dm = d[amb]
    fr_all.append(float((cc < 2 * dm.max()).float().mean()))
    fr_med.append(float((cc < 2 * dm.median()).float().mean()))
    sizes.append(int(amb.sum()))
sizes = np.array(sizes); fa = np.array(fr_all); fm = np.array(fr_med)
print("proposals %d; ambiguous frames mean %.0f max %d" % (len(sizes), sizes.mean(), sizes.max()))
print("medoids with cc < 2*Dmax: mean %.3f, weighted by subset size %.3f" % (fa.mean(), (fa * sizes).sum() / sizes.sum()))
print("medoids with cc < 2*Dmedian: mean %.3f" % fm.mean())
print("cc quantiles of last proposal:", np.quantile(cc.cpu().numpy(), [0.01, 0.1, 0.5, 0.9]), "Dmax", float(dm.max()), "Dmed", float(dm.median()))
