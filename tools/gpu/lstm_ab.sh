#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/lstm_ab.log
: > $O
run() { echo "== $*" >> $O; timeout 180 env "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
run AVVAD_LSTM_PAIR=0 python tools/micro/lstm_ab.py 256 317 --save /tmp/ref256.pt
for v in 0 2 6 14; do
run AVVAD_LSTM_PAIR=1 AVVAD_LSTM_VARIANT=$v python tools/micro/lstm_ab.py 256 317 --cmp /tmp/ref256.pt
done
run AVVAD_LSTM_PAIR=1 AVVAD_LSTM_EPI_WARPS=16 AVVAD_LSTM_VARIANT=2 python tools/micro/lstm_ab.py 256 317 --cmp /tmp/ref256.pt
run AVVAD_LSTM_PAIR=1 AVVAD_LSTM_VARIANT=2 python tools/micro/lstm_ab.py 256 317 --cmp /tmp/ref256.pt --trace
run AVVAD_LSTM_PAIR=0 python tools/micro/lstm_ab.py 256 317 --train --save /tmp/ref256t.pt
run AVVAD_LSTM_PAIR=1 AVVAD_LSTM_VARIANT=2 python tools/micro/lstm_ab.py 256 317 --train --cmp /tmp/ref256t.pt
run AVVAD_LSTM_PAIR=0 python tools/micro/lstm_ab.py 128 317
run AVVAD_LSTM_PAIR=0 python tools/micro/lstm_ab.py 32 317
grep -v "^rc=0" $O
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x -k "lstm or model or strong or config or edge or pipeline or train" 2>&1 | tail -8
