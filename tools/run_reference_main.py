"""Run the reference's UNCHANGED main.py on top of the drop-in overlay.

The reference tree is looked up at /root/reference, else at baseline/_ref (git-ignored copy that
travels to the GPU box); nothing of it is imported by the product or the tests.  Writes a synthetic
stream in the reference's on-disk format (BASELINE config 1: 2 000 users x 200 stocks x 20 000
events) into a temp dir and runs `python -m main` there with the overlay first on PYTHONPATH.
"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pfotgnrec_b200.synth import make_stream, write_reference_format   # noqa: E402


def main():
    ref = "/root/reference" if os.path.isdir("/root/reference") else os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref, "main.py")):
        print("reference tree not available: skipped")
        return 0
    extra = sys.argv[1:] or ["--model_name", "ours", "--bs", "128", "--epoch", "1", "--drop_out", "0.0"]
    work = tempfile.mkdtemp(prefix="pfo_main_")
    st = make_stream(n_users=2000, n_items=200, n_events=20000, n_days=200, seed=0, ts_mode="nbg")
    write_reference_format(st, work, period="30")
    env = dict(os.environ, WANDB_MODE="disabled", PYTHONDONTWRITEBYTECODE="1",
               PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "pfotgnrec_b200", "overlay"), ref]))
    t0 = time.time()
    p = subprocess.run([sys.executable, "-m", "main"] + extra, cwd=work, env=env, capture_output=True, text=True)
    dt = time.time() - t0
    tail = "\n".join((p.stdout + p.stderr).strip().split("\n")[-12:])
    print(tail)
    print(f"[run_reference_main] rc={p.returncode} wall={dt:.1f}s args={' '.join(extra)}")
    return p.returncode


if __name__ == "__main__":
    sys.exit(main())
