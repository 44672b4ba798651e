"""One launch set of pfo_linear_tf32 on a step shape (for ncu):  python tools/linear_one.py K N lda ldc passes [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfotgnrec_b200 import _lib
from pfotgnrec_b200._lib import ptr
K, N, lda, ldc, passes = (int(x) for x in sys.argv[1:6])
M = int(sys.argv[6]) if len(sys.argv) > 6 else 49152
lib = _lib.load()
lib.pfo_debug_set_linear_min_stages(int(os.environ.get("PFO_MIN_STAGES", "2")))
A = torch.randn(M, lda, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
C = torch.empty(M, ldc, device="cuda")
for _ in range(3):
    _lib.call("pfo_linear_tf32", ptr(A), lda, None, ptr(W), K, 0, ptr(b), None, 0, ptr(C), ldc, M, None, N, K, 1.0, 0, None, None, 0, 0, passes)
torch.cuda.synchronize()
