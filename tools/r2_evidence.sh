#!/bin/bash
# Round-2 evidence pass (one B200, under gpurun): GPU test suite, the bench arms, sub-path timings, CUPTI timelines, the
# ncu launch lists / DRAM sample / --set full captures that profiles/README.md cites.  Outputs land in gpurun_out/.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2_gpu_tests.log
python bench.py --steps 5 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
python bench.py --config train --steps 10 --warmup 3 > $O/r2_bench_train_n1.json 2> $O/r2_bench_train_n1.err
python tools/bench_misc.py > $O/r2_bench_misc.log 2>&1
python tools/vq_bench.py 2>&1 | grep -v -i warn > $O/r2_vq_bench.log
python tools/decode_curve.py > $O/r2_decode_curve.log 2>&1
python tools/trace_decode.py 24 0 > $O/r2_trace_decode_ctx0.log 2>&1
python tools/trace_decode.py 24 200 > $O/r2_trace_decode_ctx200.log 2>&1
python tools/trace_train.py 8 > $O/r2_trace_train.log 2>&1
# launch list of the first 12 decode positions (122 kernels each) and the DRAM traffic of two positions at context 133
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_decode_launches_ctx12.csv \
    python tools/profile_targets.py decode 14 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 16230 -c 250 --csv \
    --log-file $O/r2_decode_dram_ctx133.csv python tools/profile_targets.py decode 137 > /dev/null 2>&1
# --set full: fold GEMMs + attention of a decode position, the quantiser pair
ncu --set full --clock-control none --import-source on -k "regex:gemm_decode_fold|attn_decode|gemm_tc_kernel" --launch-skip 6000 -c 6 -f \
    -o $O/r2_prof_decode_chain python tools/profile_targets.py decode 60 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:vq_ --launch-skip 6 -c 2 -f -o $O/r2_prof_vq python tools/vq_bench.py > /dev/null 2>&1
ls -la $O | tail -30
