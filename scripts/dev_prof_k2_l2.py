import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from enspara_b200 import synth
from enspara_b200.cluster._engine import KCentersEngine
from enspara_b200.cluster.kcenters import _SingleComm
X = synth.device_features(1_000_000, 64, seed=0)
eng = KCentersEngine(X, "euclidean", _SingleComm())
eng.run(97, 0.0)
torch.cuda.synchronize()
