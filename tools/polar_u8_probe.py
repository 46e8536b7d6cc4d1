import torch, sys
sys.path.insert(0, '.')
import witw_b200 as W
gen = torch.Generator(device='cuda').manual_seed(0)
t8 = torch.randint(0, 256, (1024, 3, 256, 256), device='cuda', dtype=torch.uint8, generator=gen)
for _ in range(4):
    W.normalized_polar(t8)
torch.cuda.synchronize()
