"""Run the reference's UNCHANGED main.py on top of the drop-in overlay, on a GPU.

    python tools/run_reference_main.py [--models ours,tgn,jodie,dyrep,tgat] [--log DIR] [-- extra main.py flags]

The reference tree is /root/reference in the build container and the staged copy baseline/_ref on the GPU box
(baseline/stage_reference.py); nothing of it is imported by the product.  Writes a synthetic stream in the
reference's on-disk format (BASELINE config 1: 2 000 users x 200 stocks x 20 000 events, NBG-format timestamps)
into a temp dir and runs `python -m main --model_name M ...` there with sys.path = [cwd, repo, overlay, reference]:
main.py and evaluation.py come from the reference byte for byte, model.* / modules.* / utils.utils resolve to the
overlay (namespace packages).  `--stock` runs the same command WITHOUT the overlay (the reference's own
torch-CUDA path) for the side-by-side wall time.  One summary line per model; full output under --log.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
from pfotgnrec_b200.synth import make_stream, write_reference_format   # noqa: E402
from stage_reference import ref_root                                     # noqa: E402


def run_one(ref, work, model, extra, overlay=True, log_dir=None):
    path = [ROOT] + ([os.path.join(ROOT, "pfotgnrec_b200", "overlay")] if overlay else []) + [ref]
    env = dict(os.environ, WANDB_MODE="disabled", PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=os.pathsep.join(path),
               PFO_TRACE_IMPORTS="1")
    # which file each hot-path module was imported from is printed by a -c prologue, then main runs as __main__
    prologue = ("import runpy,sys,importlib;"
                "m=importlib.import_module('model.tgn');u=importlib.import_module('utils.utils');"
                "print('[imports] model.tgn <-',m.__file__);print('[imports] utils.utils <-',u.__file__);"
                "sys.argv=['main']+sys.argv[1:];runpy.run_module('main',run_name='__main__')")
    cmd = [sys.executable, "-c", prologue, "--model_name", model] + extra
    t0 = time.time()
    p = subprocess.run(cmd, cwd=work, env=env, capture_output=True, text=True)
    dt = time.time() - t0
    text = p.stdout + "\n" + p.stderr
    tag = f"{model}_{'overlay' if overlay else 'stock'}"
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
        clean = re.sub(r"\r[^\n]*", "", text)            # drop tqdm carriage-return frames
        open(os.path.join(log_dir, f"main_{tag}.log"), "w").write(clean[-20000:])
    imports = [l for l in text.split("\n") if l.startswith("[imports]")]
    tail = [l for l in text.strip().split("\n") if l.strip() and "\r" not in l][-3:]
    return {"model": model, "overlay": overlay, "rc": p.returncode, "wall_s": round(dt, 1), "imports": imports,
            "tail": tail if p.returncode else []}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="ours")
    ap.add_argument("--log", default=None)
    ap.add_argument("--stock", action="store_true", help="also run the reference's own modules (no overlay) on the GPU")
    ap.add_argument("--users", type=int, default=2000)
    ap.add_argument("--items", type=int, default=200)
    ap.add_argument("--events", type=int, default=20000)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    ref = ref_root()
    if ref is None:
        print("reference tree not available: skipped")
        return 0
    extra = [x for x in a.rest if x != "--"] or ["--bs", "128", "--epoch", "1", "--drop_out", "0.0", "--test_run"]
    work = tempfile.mkdtemp(prefix="pfo_main_")
    st = make_stream(n_users=a.users, n_items=a.items, n_events=a.events, n_days=200, seed=0, ts_mode="nbg")
    write_reference_format(st, work, period="30")
    rc = 0
    for model in a.models.split(","):
        for overlay in ([True, False] if a.stock else [True]):
            r = run_one(ref, work, model, extra, overlay=overlay, log_dir=a.log)
            print(json.dumps(r), flush=True)
            if overlay:
                rc |= r["rc"]
    print(f"[run_reference_main] reference={ref} args={' '.join(extra)} rc={rc}")
    return rc


if __name__ == "__main__":
    sys.exit(main())
