"""64 views (with AugMix) of one 500x375 image between cudaProfilerStart/Stop -- for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rlcf_b200 import datautils as D
g = np.random.RandomState(3)
img = torch.from_numpy(g.randint(0, 256, size=(375, 500, 3)).astype(np.uint8))
torch.manual_seed(0); np.random.seed(0)
aug = D.AugMixAugmenter(n_views=63, augmix=True, device="cuda:0")
aug.views(img)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
aug.views(img)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
