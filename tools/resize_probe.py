"""Launches the resize kernel a few times for an ncu capture: aerial 750x750 u8 -> 256x256 (default) or, with `pano`,
panorama 224x1232 u8 -> 128x512; antialias, normalised."""
import sys
import torch
sys.path.insert(0, '.')
import witw_b200 as W
from witw_b200 import ops
gen = torch.Generator(device='cuda').manual_seed(0)
if len(sys.argv) > 1 and sys.argv[1] == "pano":
    raw, size = torch.randint(0, 256, (1024, 3, 224, 1232), device='cuda', dtype=torch.uint8, generator=gen), (128, 512)
else:
    raw, size = torch.randint(0, 256, (512, 3, 750, 750), device='cuda', dtype=torch.uint8, generator=gen), (256, 256)
for _ in range(4):
    W.resize_normalize(raw, size[0], size[1], True, mean=ops.IMG_MEAN, std=ops.IMG_STD)
torch.cuda.synchronize()
