"""Developer probe for ncu: k-centers with use_triangle_inequality, then a few PAM proposals
(pruned full pass), on 1M x 500."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from enspara_b200 import synth
from enspara_b200.cluster import util, kcenters as kc
from enspara_b200.cluster._pam import PamEngine
from enspara_b200.cluster.kcenters import _SingleComm
n, A, k = 1_000_000, 500, 300
data = synth.device_trajectory(n, A, seed=0)
res, eng = kc.kcenters(data, "rmsd", n_clusters=k, use_triangle_inequality=True, _return_engine=True)
pam = PamEngine(data, util.RMSD, _SingleComm(), eng.dist, eng.assign,
                [int(c) for c in res.center_indices])
pam.sweep(random_state=0, max_proposals=6)
torch.cuda.synchronize()
