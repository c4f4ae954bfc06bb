#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/lstm_ab.log
: > $O
run() { echo "== $*" >> $O; timeout 180 env "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
run AVVAD_LSTM_PAIR=0 python tools/micro/lstm_ab.py 256 317 --save /tmp/ref256.pt
run AVVAD_LSTM_NP=128 python tools/micro/lstm_ab.py 256 317 --cmp /tmp/ref256.pt
run AVVAD_LSTM_NP=128 AVVAD_LSTM_VARIANT=512 python tools/micro/lstm_ab.py 256 317 --cmp /tmp/ref256.pt --trace
grep -v "^rc=0" $O
BENCH=1 bash tools/gpu/r2_suite.sh
