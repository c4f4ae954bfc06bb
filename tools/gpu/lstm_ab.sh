#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/lstm_ab.log
: > $O
run() { echo "== $*" >> $O; timeout 180 env "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
for B in 32 64 128; do
run AVVAD_LSTM_PAIR_MIN=129 python tools/micro/lstm_ab.py $B 317 --save /tmp/ref$B.pt
run AVVAD_LSTM_PAIR_MIN=1 python tools/micro/lstm_ab.py $B 317 --cmp /tmp/ref$B.pt
run AVVAD_LSTM_PAIR_MIN=129 python tools/micro/lstm_ab.py $B 317 --train --save /tmp/reft$B.pt
run AVVAD_LSTM_PAIR_MIN=1 python tools/micro/lstm_ab.py $B 317 --train --cmp /tmp/reft$B.pt
done
grep -v "^rc=0" $O
