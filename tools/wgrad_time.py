"""Graph-replay timing of pfo_wgrad_tf32 on the shapes of the training step (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfotgnrec_b200 import _lib
from pfotgnrec_b200._lib import ptr

_lib.load()
M = 49152
# (name, N (cols of G), K (cols of A), ldg, lda, bias)
shapes = [("gW2   G=dOUT[.,64]  A=H1[.,64]", 64, 64, 64, 64, 1), ("gWc1T G=CAT[.,328] A=dH1[.,64]", 328, 64, 328, 64, 0),
          ("gWqk  G=dQK[.,264] A=hq[.,64]", 264, 64, 264, 328, 1), ("gW_ih G=dGI[.,192] A=XG[.,193]", 192, 193, 192, 196, 1),
          ("gW_hh G=dGH[.,192] A=HG[.,64]", 192, 64, 192, 64, 1)]
for passes in (3, 1):
    tot = 0.0
    for name, N, K, ldg, lda, wb in shapes:
        G = torch.randn(M, ldg, device="cuda"); A = torch.randn(M, lda, device="cuda")
        dW = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda")
        ws = torch.empty(int(_lib.query("pfo_wgrad_tf32_workspace_floats", M, N, K, wb)), device="cuda")
        def run():
            _lib.call("pfo_wgrad_tf32", ptr(G), ldg, ptr(A), lda, None, M, None, N, K, ptr(dW), K, ptr(db) if wb else None, 0,
                      ptr(ws), passes)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                run()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 20
        tot += us
        print(f"{name} passes={passes}: {us:.1f} us/launch, {M * (N + K) * 4 / us / 1e3:.0f} GB/s algorithmic")
    print(f"  total passes={passes}: {tot:.1f} us")
