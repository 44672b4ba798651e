"""Timeline (clock64) of CTA (0,0) of pfo_linear_tf32 for the shapes of the training step (diagnostic)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfotgnrec_b200 import _lib
from pfotgnrec_b200._lib import ptr

lib = _lib.load()
lib.pfo_debug_set_linear_trace.argtypes = [ctypes.c_void_p]
lib.pfo_debug_set_linear_trace.restype = None
M = 49152
lib.pfo_debug_set_linear_min_stages(int(os.environ.get("PFO_MIN_STAGES", "2")))
lib.pfo_debug_set_linear_dbg(int(os.environ.get("PFO_LIN_DBG", "0")))
WT = int(os.environ.get("PFO_WT", "0"))
PASSES = tuple(int(x) for x in os.environ.get("PFO_PASSES", "3,1").split(","))
shapes = [("OUT 64->64", 64, 64, 64, 64), ("QK 64->264", 64, 264, 328, 264), ("H1 328->64", 328, 64, 328, 64),
          ("dCAT 64->328", 64, 328, 64, 328), ("dhq 264->64", 264, 64, 264, 64), ("GI 193->192", 193, 192, 196, 192),
          ("GH 64->192", 64, 192, 64, 192)]
for passes in PASSES:
    for name, K, N, lda, ldc in shapes:
        A = torch.randn(M, lda, device="cuda")
        W = torch.randn(N, K, device="cuda")
        b = torch.randn(N, device="cuda")
        C = torch.empty(M, ldc, device="cuda")
        tr = torch.zeros(128 + 2 * 4096, dtype=torch.int64, device="cuda")
        def run():
            _lib.call("pfo_linear_tf32", ptr(A), lda, None, ptr(W), N if WT else K, WT, ptr(b), None, 0, ptr(C), ldc, M, None, N, K,
                      1.0, 0, None, None, 0, 0, passes)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()                # replayed like the training step: no host launch cost in the timing
        with torch.cuda.graph(g):
            for _ in range(20):
                run()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 20
        lib.pfo_debug_set_linear_trace(tr.data_ptr())
        run()
        torch.cuda.synchronize()
        lib.pfo_debug_set_linear_trace(None)
        t = tr.cpu().tolist()
        t0 = t[0]
        byts = M * (K + N) * 4
        print(f"\n{name} passes={passes}: {us:.1f} us/launch graph replay (warm L2), {byts / us / 1e3:.0f} GB/s algorithmic; "
              f"CTA0 cycles: init {t[3] - t0}, W loaded {t[4] - t0}, prologue {t[1] - t0}, total {t[2] - t0}")
        import numpy as np
        w = np.array(t[128:]).reshape(-1, 2)
        w = w[w[:, 0] > 0]
        b0 = w[:, 0].min()
        print(f"  {len(w)} CTAs: start spread {(w[:, 0].max() - b0) / 1e3:.2f} us, first end {(w[:, 1].min() - b0) / 1e3:.2f} us, "
              f"last end {(w[:, 1].max() - b0) / 1e3:.2f} us, mean lifetime {(w[:, 1] - w[:, 0]).mean() / 1e3:.2f} us")
        for i in range(8):
            r = t[8 + 8 * i: 8 + 8 * i + 6]
            if r[0] == 0:
                break
            print(f"  tile {i}: tma_issue {r[0] - t0:7d} split_saw_full {r[1] - t0 if r[1] else -1:7d} mma_saw_ready {r[2] - t0:7d} "
                  f"mma_commit {r[3] - t0:7d} epi_saw_tfull {r[4] - t0:7d} epi_done {r[5] - t0:7d}")
