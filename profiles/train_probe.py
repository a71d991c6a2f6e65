"""One training-style forward + backward of the hot path at KITTI size (module path, autograd): kernel list target."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda:0")
hp = bench.make_hot_path().to(dev).train()
L, R = bench.make_inputs(1, 1, dev)[0]
L = [t.clone().requires_grad_() for t in L]
R = [t.clone().requires_grad_() for t in R]
for it in range(3):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    out = hp(L, R)
    loss = sum(o.mean() for o in out)
    e1.record()
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    print("iter %d  forward %.2f ms  backward %.2f ms" % (it, e0.elapsed_time(e1), e1.elapsed_time(e2)))
