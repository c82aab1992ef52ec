#!/bin/bash
# FIRST GPU call of round 2 (under gpurun, one GPU): validate and time the variants round 1 left unrun:
#  - tcn2_mac_kernel<32,*,1,4> (four slots per item) and its shared-memory staged whole-sector stores (CRCNN_TCN2_NS=4 [CRCNN_TCN2_STAGE=1])
#  - the BEHZ kernels with the special-prime three-fold reduction (ab/libF.so = build with -DCRCNN_FOLD128, cross-compiled beforehand)
# Everything runs under `timeout`: a pipeline bug in a tcgen05 kernel hangs rather than fails.
#   gpurun --timeout 900 -- 'bash tools/round2_first.sh r02a'
tag=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader; nproc
CRCNN_TCN2_NS=4 timeout 200 python -m pytest tests/test_gpu_tcn.py -q -x > gpurun_out/${tag}_ns4_tests.log 2>&1
echo "ns4 tests rc=$?"; tail -3 gpurun_out/${tag}_ns4_tests.log
CRCNN_TCN2_NS=4 CRCNN_TCN2_STAGE=1 timeout 200 python -m pytest tests/test_gpu_tcn.py -q -x > gpurun_out/${tag}_ns4s_tests.log 2>&1
echo "ns4 staged tests rc=$?"; tail -3 gpurun_out/${tag}_ns4s_tests.log
for v in "" "CRCNN_TCN2_NS=4" "CRCNN_TCN2_NS=4 CRCNN_TCN2_STAGE=1" "CRCNN_TCN2_NS=4 CRCNN_TCN2_STAGE=1 CRCNN_TCN_FOLD=1"; do
  echo "== conv1 only: $v"; env $v timeout 40 python tools/quick_layers.py --first 0 --last 1 2>&1 | tail -1
done
CRCNN_B200_LIB=$PWD/ab/libF.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -x -k "square or golden or chain or seal" > gpurun_out/${tag}_fold128_tests.log 2>&1
echo "fold128 tests rc=$?"; tail -3 gpurun_out/${tag}_fold128_tests.log
echo "== square layer, default"; timeout 60 python tools/quick_layers.py --first 4 --last 5 2>&1 | tail -1
echo "== square layer, fold128"; CRCNN_B200_LIB=$PWD/ab/libF.so timeout 60 python tools/quick_layers.py --first 4 --last 5 2>&1 | tail -1
