"""A handful of resize / normalisation launches for compute-sanitizer (memcheck): odd sizes, both source types, both resize
variants, panorama windows, a tensor that ends right at its allocation's last bytes."""
import sys
import torch
sys.path.insert(0, '.')
import witw_b200 as W
from witw_b200 import ops
gen = torch.Generator().manual_seed(0)
for (ih, iw, oh, ow) in [(75, 75, 256, 256), (750, 750, 256, 256), (97, 411, 128, 512), (37, 53, 40, 70), (1, 1, 5, 3), (256, 256, 256, 256), (5, 7, 5, 7)]:
    for dt in (torch.uint8, torch.float32):
        img = torch.randint(0, 256, (2, 3, ih, iw), generator=gen).to(dt).cuda()
        for aa in (True, False):
            W.resize_normalize(img, oh, ow, aa)
            W.resize_normalize(img, oh, ow, aa, mean=ops.IMG_MEAN, std=ops.IMG_STD, col_start=ow // 2, col_count=max(1, ow // 3))
    torch.cuda.synchronize()
d = W.prepare_pair(torch.randint(0, 256, (3, 3, 224, 1232), generator=gen, dtype=torch.uint8).cuda(),
                   torch.randint(0, 256, (3, 3, 750, 750), generator=gen, dtype=torch.uint8).cuda(), fov=90, start=500)
torch.cuda.synchronize()
print("resize sanitize run done", tuple(d["polar"].shape))
