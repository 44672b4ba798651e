#!/bin/bash
# GPU smoke of the command-line driver on BASELINE config 1 (2 000 users x 200 stocks x 20 000 events, synthetic files
# in the reference's on-disk format): one epoch of training + validation + test evaluation, one JSON line.
set -e
D=$(mktemp -d)
python -m pfotgnrec_b200.run --data-root "$D" --period 30 --synthetic 2000 200 20000 200 --model_name ours --bs 128 --epoch 1 \
  | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cli ok: loss %.4f valid_recall@5 %.4f test_ndcg@5 %.4f test_sharpe_avg_5_ %.4f (%d keys, %.1f s)' % (d['loss'], d['valid_recall_avg_5'], d['test_ndcg_avg_5'], d['test_sharpe_avg_5_'], len(d), d['seconds']))"
