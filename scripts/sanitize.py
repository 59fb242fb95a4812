"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): a few eager iterations of the
fused engine on the tiny workload plus one raster forward/backward through the autograd ops."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from homan_b200 import synth
from homan_b200.engine import FitEngine
from homan_b200.workload import make_workload

asset = synth.make_mano_asset(0, "right")
batch, lw = make_workload("tiny", mano_asset=asset)
eng = FitEngine(batch, lw, mano_asset=asset, use_graph=False)
for _ in range(2):
    eng.step()
torch.cuda.synchronize()
print("total", eng.total.cpu().tolist())
