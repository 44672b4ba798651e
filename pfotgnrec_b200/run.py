"""Command-line driver of the B200 path: the loop of reference main.py (train epochs, then validation and test
evaluation) with the MV selection, BPR and the evaluation metric block on the device.

    python -m pfotgnrec_b200.run --data-root DIR --period 30 --model_name ours --bs 512 --epoch 5

reads `DIR/data/period_{p}/` in the reference's on-disk format;  `--synthetic U I E D` writes a synthetic
stream of that shape there first.  Flag names follow reference main.py:15-38.  One JSON line per epoch on stdout.
"""
import argparse
import json
import sys
import time

import torch


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--data-root", default=".")
    ap.add_argument("--period", default="30")
    ap.add_argument("--model_name", default="ours", choices=["ours", "tgn", "jodie", "dyrep", "tgat"])
    ap.add_argument("--bs", type=int, default=512)
    ap.add_argument("--epoch", type=int, default=1)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--drop_out", type=float, default=0.1)
    ap.add_argument("--n_head", type=int, default=2)
    ap.add_argument("--n_degree", type=int, default=10)
    ap.add_argument("--num_negatives", type=int, default=20)
    ap.add_argument("--p_neg_num", type=int, default=3)
    ap.add_argument("--gamma", type=float, default=2.0)
    ap.add_argument("--lambda_mv", type=float, default=0.5)
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "tf32", "bf16", "simt"])
    ap.add_argument("--test_run", action="store_true", help="two training batches per epoch (reference --test_run)")
    ap.add_argument("--synthetic", type=int, nargs=4, metavar=("USERS", "ITEMS", "EVENTS", "DAYS"), default=None)
    a = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("pfotgnrec_b200.run needs a CUDA device: the product path has no CPU fallback")
    from . import _lib
    from .synth import make_stream, read_reference_format, write_reference_format
    from .trainer import PfoTrainer, TrainConfig
    _lib.load()
    _lib.use_device(torch.device("cuda", a.gpu))
    if a.synthetic is not None:
        u, i, e, d = a.synthetic
        write_reference_format(make_stream(n_users=u, n_items=i, n_events=e, n_days=d, seed=0, ts_mode="nbg"),
                               a.data_root, period=a.period)
    st = read_reference_format(a.data_root, period=a.period)
    tc = TrainConfig(model=a.model_name, bs=a.bs, n_neighbors=a.n_degree, n_heads=a.n_head, dropout=a.drop_out,
                     lr=a.lr, num_negatives=a.num_negatives, p_neg_num=a.p_neg_num, gamma=a.gamma,
                     lambda_mv=a.lambda_mv, gemm_mode=a.gemm)
    tr = PfoTrainer(st, tc, device=torch.device("cuda", a.gpu))
    t0 = [time.time()]

    def log(out):
        out["seconds"] = time.time() - t0[0]
        t0[0] = time.time()
        print(json.dumps(out), flush=True)

    tr.fit(epochs=a.epoch, bs=a.bs, max_batches=2 if a.test_run else None, log=log)
    return 0


if __name__ == "__main__":
    sys.exit(main())
