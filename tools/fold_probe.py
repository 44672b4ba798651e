"""Run the attention fold forward + adjoint a few times (target for ncu / timing; diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfotgnrec_b200.engine import ModelConfig, _FoldAttention
cfg = ModelConfig(d=64, n_edge_feat=1, n_heads=2)
d, E, Ek = 64, 128, 129
shapes = [(E, E), (E, Ek), (E, Ek), (3 * E,), (E, E), (E,), (d, E + d), (d,), (d,)]
w = [(torch.randn(*s, device="cuda") * 0.3).requires_grad_(True) for s in shapes]
for it in range(4):
    got = _FoldAttention.apply(cfg, *w)
    torch.autograd.backward(got, [torch.ones_like(t) for t in got])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(20):
    got = _FoldAttention.apply(cfg, *w)
    torch.autograd.backward(got, [torch.ones_like(t) for t in got])
e1.record(); torch.cuda.synchronize()
print("fold fwd+bwd us per iteration (incl. host launch gaps):", e0.elapsed_time(e1) * 50)
