"""First-GPU-session probes: host cores, fp32 noise floors CPU vs CUDA."""
import json
import os
import platform

import numpy as np
import torch

out = {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads(), "machine": platform.processor()}
try:
    for line in open("/proc/cpuinfo"):
        if line.startswith("model name"):
            out["cpu_model"] = line.split(":", 1)[1].strip()
            break
except OSError:
    pass
out["gpu"] = torch.cuda.get_device_name(0)
out["sm_count"] = torch.cuda.get_device_properties(0).multi_processor_count
out["allow_tf32"] = torch.backends.cuda.matmul.allow_tf32
torch.manual_seed(0)
# TimeEncode: nn.Linear(1, 64) then cos, CPU vs CUDA on NBG-scale arguments
w = torch.tensor(1 / 10 ** np.linspace(0, 9, 64), dtype=torch.float32).reshape(64, 1)
b = torch.zeros(64)
t = (torch.rand(4096, 1) * 2e13).float()
cpu = torch.cos(torch.nn.functional.linear(t, w, b))
gpu = torch.cos(torch.nn.functional.linear(t.cuda(), w.cuda(), b.cuda())).cpu()
arg_c = torch.nn.functional.linear(t, w, b)
arg_g = torch.nn.functional.linear(t.cuda(), w.cuda(), b.cuda()).cpu()
out["timeencode_arg_bitexact_cpu_vs_cuda"] = bool(torch.equal(arg_c, arg_g))
out["timeencode_cos_maxdiff_cpu_vs_cuda"] = float((cpu - gpu).abs().max())
out["cos_same_arg_maxdiff"] = float((torch.cos(arg_c) - torch.cos(arg_c.cuda()).cpu()).abs().max())
gru = torch.nn.GRUCell(193, 64)
x, h = torch.randn(2048, 193), torch.randn(2048, 64)
out["grucell_maxdiff_cpu_vs_cuda"] = float((gru(x, h) - gru.cuda()(x.cuda(), h.cuda()).cpu()).abs().max())
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
