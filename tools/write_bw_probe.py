import torch
x = torch.empty(1024*3*128*512, device='cuda')           # 805 MB, the polar output of 1024 x 3 planes
y = torch.empty_like(x)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.fill_(1.0)); print('fill  %.3f ms  %.0f GB/s written' % (ms, x.numel()*4/ms/1e6))
ms = t(lambda: x.zero_()); print('zero  %.3f ms  %.0f GB/s written' % (ms, x.numel()*4/ms/1e6))
ms = t(lambda: y.copy_(x)); print('copy  %.3f ms  %.0f GB/s read+written' % (ms, 2*x.numel()*4/ms/1e6))
ms = t(lambda: x.sum()); print('sum   %.3f ms  %.0f GB/s read' % (ms, x.numel()*4/ms/1e6))
