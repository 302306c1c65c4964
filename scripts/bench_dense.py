import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from efg_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
m = 70688
for cin, cout in [(256, 256), (256, 1024), (1024, 256)]:
    x = torch.randn(m, cin, device=dev); go = torch.randn(m, cout, device=dev)
    w = torch.randn(cout, 1, cin, device=dev) * 0.05
    fl = 2 * m * cin * cout
    t1 = timeit(lambda: ops.spconv_tc(x, w, None, None, 0))
    t2 = timeit(lambda: ops.spconv_tc(go, w, None, None, 1))
    t3 = timeit(lambda: ops.spconv_tc_wgrad(x, go, None, 1, cin, cout))
    t4 = timeit(lambda: torch.nn.functional.linear(x, w.view(cout, cin)))
    print("%4d->%4d fwd %7.1f us (%5.1f TF)  dgrad %7.1f (%5.1f)  wgrad %7.1f (%5.1f)   cublas fp32 fwd %7.1f (%5.1f)" % (cin, cout, t1, fl/t1/1e6, t2, fl/t2/1e6, t3, fl/t3/1e6, t4, fl/t4/1e6))
