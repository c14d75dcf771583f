"""Dev diagnostic: a few filter + render launches on the 2000 x 2000 x 5 grid (for an ncu capture of k_render*)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_semantic_segmentation_b200 import synthetic as syn
from vision_semantic_segmentation_b200.renderer import filter_and_render

mh, c = 2000, 5
g = torch.zeros((mh, mh, c), dtype=torch.float64, device="cuda")
g[mh // 4: 3 * mh // 4, mh // 4: 3 * mh // 4] = torch.randint(0, 9, (mh // 2, mh // 2, c), device="cuda").double()
for _ in range(4):
    filter_and_render(g, syn.COLORS_19[:c])
torch.cuda.synchronize()
