import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from enspara_b200 import synth
from enspara_b200.cluster import _ops, util
data = synth.device_trajectory(262144, 500, seed=0)
cen = data.gather(torch.arange(0, 262144, 256, device="cuda")[:1024])
_ops.assign_device_tc(util.RMSD, data, cen)
torch.cuda.synchronize()
_ops.assign_device_tc(util.RMSD, data, cen)
torch.cuda.synchronize()
