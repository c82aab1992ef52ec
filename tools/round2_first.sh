#!/bin/bash
# FIRST GPU call of round 2 (under gpurun, one GPU): validate and time the experimental four-slots-per-item variant of the column-major
# limb-split kernel (tcn2_mac_kernel<32, *, 1, 4>; round 1 only ran six of its parity tests and one timing; DESIGN.md section 6 "Correction", section 9 item 1a).
# Everything runs under `timeout`: a pipeline bug in a tcgen05 kernel hangs rather than fails.
#   gpurun --timeout 600 -- 'bash tools/round2_first.sh r02a'
tag=${1:-r02a}
mkdir -p gpurun_out
CRCNN_TCN2_NS=4 timeout 240 python -m pytest tests/test_gpu_tcn.py tests/test_gpu_golden.py tests/test_gpu_builder.py -q > gpurun_out/${tag}_ns4_tests.log 2>&1
echo "ns4 tests rc=$?"; tail -5 gpurun_out/${tag}_ns4_tests.log
# whole-sector stores from shared-memory staging (STG.256): never run on hardware in round 1
CRCNN_TCN2_NS=4 CRCNN_TCN2_STAGE=1 timeout 240 python -m pytest tests/test_gpu_tcn.py tests/test_gpu_golden.py tests/test_gpu_builder.py -q > gpurun_out/${tag}_ns4s_tests.log 2>&1
echo "ns4 staged tests rc=$?"; tail -5 gpurun_out/${tag}_ns4s_tests.log
for v in "" "CRCNN_TCN2_NS=4" "CRCNN_TCN2_NS=4 CRCNN_TCN2_STAGE=1" "CRCNN_TCN2_NS=4 CRCNN_TCN2_STAGE=1 CRCNN_TCN_FOLD=1"; do
  echo "== conv1 only: $v"; env $v timeout 40 python tools/quick_layers.py --first 0 --last 1 2>&1 | tail -1
done
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_ns1.json 2> gpurun_out/${tag}_bench_ns1.err; echo "ns1 rc=$?"
CRCNN_TCN2_NS=4 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_ns4.json 2> gpurun_out/${tag}_bench_ns4.err; echo "ns4 rc=$?"
CRCNN_TCN2_NS=4 CRCNN_TCN2_STAGE=1 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_ns4_fold.json 2> gpurun_out/${tag}_bench_ns4_fold.err; echo "ns4 staged rc=$?"
python - <<P
import json
for f in ("ns1", "ns4", "ns4_fold"):   # ns4_fold.json holds the STAGED run
    try:
        d = json.load(open("gpurun_out/${tag}_bench_%s.json" % f))
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in d["per_layer_ms"].items() if "conv" in k or "fc4" in k})
    except Exception as e:
        print(f, "ERR", e)
P

# BEHZ kernels with the three-fold special-prime reduction (-DCRCNN_FOLD128; built here: nvcc is on the box): parity, then the square layer alone
make -s -C crcnn_b200/csrc EXTRA=-DCRCNN_FOLD128 OUT=../../ab/libF.so OBJDIR=../../build/objF > gpurun_out/${tag}_fold128_build.log 2>&1; echo "fold128 build rc=$?"
CRCNN_B200_LIB=$PWD/ab/libF.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -k "square or golden or chain or seal" > gpurun_out/${tag}_fold128_tests.log 2>&1
echo "fold128 tests rc=$?"; tail -4 gpurun_out/${tag}_fold128_tests.log
echo "== square layer, default"; timeout 60 python tools/quick_layers.py --first 4 --last 5 2>&1 | tail -1
echo "== square layer, fold128"; CRCNN_B200_LIB=$PWD/ab/libF.so timeout 60 python tools/quick_layers.py --first 4 --last 5 2>&1 | tail -1
