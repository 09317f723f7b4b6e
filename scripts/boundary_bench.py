"""Cost of one dependent-kernel boundary: a chain of trivial fused-sampler-step launches."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ns2vc_b200 import _lib
L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn(n, device="cuda"); o = torch.randn(n, device="cuda"); m = torch.empty_like(x); y = torch.empty_like(x)
c = _lib.DpmCoef(0.5, 0.8, 0.9, -0.1, -0.05, 1.0, 1)
def chain(k):
    s = torch.cuda.current_stream().cuda_stream
    a, b = x, y
    for _ in range(k):
        _lib.check(L.ns2vc_dpm_step(a.data_ptr(), o.data_ptr(), m.data_ptr(), C.byref(c), m.data_ptr(), b.data_ptr(), n, None, s))
        a, b = b, a
K = 500
chain(K); torch.cuda.synchronize()
for mode in ("eager", "graph"):
    if mode == "graph":
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            chain(K)
        run = g.replay
    else:
        run = lambda: chain(K)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    print(f"n={n} {mode} PDL={os.environ.get('NS2VC_PDL','1')}: {e0.elapsed_time(e1)*1e3/K:.2f} us per dependent launch")
