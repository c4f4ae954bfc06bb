#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
O=gpurun_out/block_ab.log
: > $O
run() { echo "== $*" >> $O; timeout 300 env "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
run AVVAD_BLOCK17=0 python tools/micro/trunk_ab.py 301 --save /tmp/t301.pt
run AVVAD_BLOCK17=1 python tools/micro/trunk_ab.py 301 --cmp /tmp/t301.pt
run AVVAD_BLOCK17=0 python tools/micro/trunk_ab.py 20288 --save /tmp/t20k.pt
run AVVAD_BLOCK17=1 python tools/micro/trunk_ab.py 20288 --cmp /tmp/t20k.pt
grep -v "^rc=0" $O
