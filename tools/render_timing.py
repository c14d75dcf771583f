"""Dev diagnostic: filter + render kernel time (CUDA events) for the library named by SMAP_LIB_PATH."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_semantic_segmentation_b200 import synthetic as syn
from vision_semantic_segmentation_b200.renderer import filter_and_render, render_bev_map

peak = 6547.8
out = []
for mh, c in ((2000, 5), (2000, 19), (10000, 5)):
    g = torch.zeros((mh, mh, c), dtype=torch.float64, device="cuda")
    g[mh // 4: 3 * mh // 4, mh // 4: 3 * mh // 4] = torch.randint(0, 9, (mh // 2, mh // 2, c), device="cuda").double()
    colors = syn.COLORS_19[:c]
    for name, fn in (("filter+render", lambda: filter_and_render(g, colors)), ("render", lambda: render_bev_map(g, colors))):
        fn(); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        for i in range(7):
            ev[i].record(); fn()
        ev[7].record(); torch.cuda.synchronize()
        ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(7)]))
        b = mh * mh * c * 8.0 + mh * mh * 3.0
        out.append("%dx%dx%d %s %.1f us (%.0f%%)" % (mh, mh, c, name, ms * 1e3, 100 * b / (ms * 1e-3) / 1e9 / peak))
    del g
print(os.path.basename(os.environ.get("SMAP_LIB_PATH", "default")), " | ".join(out))
