#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/lstm_ab.log
: > $O
run() { echo "== $*" >> $O; timeout 180 env "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
run AVVAD_LSTM_PAIR=0 AVVAD_LSTM_CHUNKS=1 python tools/micro/lstm_ab.py 256 317 --save /tmp/ref256.pt
for c in 1 2 4 6 8 12; do
run AVVAD_LSTM_CHUNKS=$c python tools/micro/lstm_ab.py 256 317 --cmp /tmp/ref256.pt
done
run AVVAD_LSTM_CHUNKS=1 python tools/micro/lstm_ab.py 256 317 --train --save /tmp/ref256t.pt
run AVVAD_LSTM_CHUNKS=6 python tools/micro/lstm_ab.py 256 317 --train --cmp /tmp/ref256t.pt
run AVVAD_LSTM_CHUNKS=1 python tools/micro/lstm_ab.py 32 317 --save /tmp/ref32.pt
run AVVAD_LSTM_CHUNKS=6 python tools/micro/lstm_ab.py 32 317 --cmp /tmp/ref32.pt
run AVVAD_LSTM_CHUNKS=1 python tools/micro/lstm_ab.py 300 317 --save /tmp/ref300.pt
run AVVAD_LSTM_CHUNKS=6 python tools/micro/lstm_ab.py 300 317 --cmp /tmp/ref300.pt
grep -v "^rc=0" $O
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x -k "lstm or model or strong or config or edge or pipeline or train" 2>&1 | tail -4
