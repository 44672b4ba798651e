"""Side by side on the GPU: the UNMODIFIED reference's own torch-CUDA path vs the drop-in overlay.

The committed goldens (tests/golden/tgn_*.npz) were produced by the reference on CPU in the build container.  Here
the staged reference (baseline/_ref on the GPU box, /root/reference in the build container) is run ON THE SAME GPU,
in its own process (it shares the module paths model.* / modules.* / utils.* with the overlay), through
`python -m oracle.make_golden --device cuda`: same synthetic stream, same torch seed (so the same initial weights),
same batches.  The overlay then has to reproduce the reference-on-CUDA vectors -- embeddings, BPR loss, every
parameter gradient, memory, last_update, pending messages -- under the same 1e-5 / 5e-5 contract.  This also answers
SURVEY Appendix B questions 2-3 (does torch-CUDA TimeEncode / GRUCell / MultiheadAttention stay within the contract
of the CPU goldens): the test records the distance between the CPU goldens and the CUDA vectors as well.
Skipped when no reference tree is available.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import load_golden, rel_err
from test_gpu_model import overlay, check_against_vectors   # noqa: F401  (fixture + checker)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
from stage_reference import ref_root   # noqa: E402

# the models main.py builds (--model_name ours / tgn / jodie / dyrep / tgat) plus NBG-format timestamps
TAGS = ["ours", "ours_nbg", "tgn", "jodie", "dyrep", "tgat2"]


@pytest.fixture(scope="module")
def cuda_vectors(tmp_path_factory):
    if ref_root() is None:
        pytest.skip("no reference tree (neither /root/reference nor baseline/_ref)")
    out = tmp_path_factory.mktemp("ref_cuda")
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-m", "oracle.make_golden", "--device", "cuda", "--out", str(out),
                        "--tags", ",".join(TAGS)], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    return out


@pytest.mark.parametrize("tag", TAGS)
def test_overlay_matches_reference_running_on_cuda(overlay, cuda_vectors, tag):   # noqa: F811
    tgn_mod, utils_mod = overlay
    z = dict(np.load(os.path.join(str(cuda_vectors), f"tgn_{tag}.npz"), allow_pickle=False))
    worst = check_against_vectors(tgn_mod, utils_mod, z, "fp32", tag + "@cuda")
    # the reference itself, CPU (committed golden) vs CUDA (this run): the noise floor the contract lives above
    g = load_golden(f"tgn_{tag}.npz")
    floor = max(rel_err(z[k], g[k]) for k in z if k.startswith("b") and "_emb_" in k)
    print(f"[side-by-side] {tag}: overlay vs reference-on-CUDA worst rel. err {worst}; "
          f"reference CPU vs reference CUDA embeddings {floor:.2e}")
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "side_by_side.log"), "a") as f:
            f.write(f"{tag}: overlay-vs-reference@cuda {worst}; reference cpu-vs-cuda emb {floor:.3e}\n")
