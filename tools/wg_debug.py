import sys, torch
sys.path.insert(0, ".")
from pfotgnrec_b200 import _lib
from pfotgnrec_b200._lib import ptr
M, N, K = 64, 128, 64
for passes in (1, 3):
    G = torch.ones(M, N, device="cuda"); A = torch.ones(M, K, device="cuda")
    dW = torch.full((N, K), -1.0, device="cuda"); db = torch.full((N,), -1.0, device="cuda")
    ws = torch.full((_lib.query("pfo_wgrad_tf32_workspace_floats", M, N, K, 1),), -7.0, device="cuda")
    _lib.call("pfo_wgrad_tf32", ptr(G), N, ptr(A), K, None, M, None, N, K, ptr(dW), K, ptr(db), 0, ptr(ws), passes)
    torch.cuda.synchronize()
    print("passes", passes, "dW[:2,:6]", dW[:2, :6].tolist(), "db[:4]", db[:4].tolist(), "ws[:8]", ws[:8].tolist(), "ws numel", ws.numel())
