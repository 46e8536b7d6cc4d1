#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "topk or rank" > gpurun_out/tests_topk.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/tests_topk.log | cut -c1-300
python - <<'PY'
import torch, sys, json
sys.path.insert(0,'.')
import witw_b200 as W
gen = torch.Generator(device='cuda').manual_seed(0)
d = torch.rand(10000, 10000, device='cuda', generator=gen)
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/iters
for k in (1,10,32):
    ms = timeit(lambda: W.topk_from_distances(d, k))
    print(json.dumps({"kernel":"topk_from_distances k=%d 10k x 10k (sample pass + thresholded pass + merge)"%k,"ms":ms,"achieved":4e8/ms/1e6,"unit":"GB/s"}))
PY
