import torch, sys
sys.path.insert(0, '.')
import witw_b200 as W
gen = torch.Generator(device='cuda').manual_seed(0)
d = torch.rand(10000, 10000, device='cuda', generator=gen)
tiles = torch.randn(1024, 3, 256, 256, device='cuda', generator=gen)
for _ in range(3):
    W.rank_from_distances(d)
    W.topk_from_distances(d, 10)
    W.polar_transform(tiles)
torch.cuda.synchronize()
