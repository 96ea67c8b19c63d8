#!/bin/bash
# Evidence pass of the FINAL round-2 build (one B200, under gpurun): GPU test suite, both bench arms with the driver's
# arguments, sub-path timings, CUPTI timelines of the VQVAE decoder / teacher-forced forward, the phase trace of the
# tcgen05 prefill attention, ncu launch lists (decode positions, VQVAE decoder) and --set full captures of an 80x848
# convolution pair and of the prefill attention, sanitizer runs over the kernels that changed.  Outputs: gpurun_out/r2f_*.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2f_gpu_tests.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2f_bench_reference_arm.json 2> $O/r2f_bench_reference_arm.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err
python tools/bench_misc.py > $O/r2f_bench_misc.log 2>&1
python tools/trace_vqvae.py 64 decode 2>&1 | grep -v -i warn > $O/r2f_trace_vqvae_decode.log
python tools/trace_vqvae.py 64 encode 2>&1 | grep -v -i warn > $O/r2f_trace_vqvae_encode.log
python tools/trace_forward.py 2>&1 | grep -v -i warn > $O/r2f_trace_forward.log
python tools/attn_trace.py 64 265 > $O/r2f_attn_trace.log 2>&1
python tools/decode_curve.py > $O/r2f_decode_curve.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2f_decode_launches_ctx12.csv \
    python tools/profile_targets.py decode 14 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_vqvae_decode_launches_b64.csv \
    python tools/profile_targets.py vqvae 64 > /dev/null 2>&1
# --set full: the last two 80x848 128->128 convolutions of a B=64 decode (conv2 of the last ResnetBlock carries the
# residual), and one launch of the tcgen05 prefill attention at bs=64
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persist --launch-skip 55 -c 2 -f -o $O/r2f_prof_conv_80x848 \
    python tools/profile_targets.py vqvae 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_prefill_tc --launch-skip 3 -c 1 -f -o $O/r2f_prof_attn_prefill_tc \
    python tools/profile_targets.py prefill > /dev/null 2>&1
compute-sanitizer --tool memcheck python tools/sanitize_targets.py gpt vqvae > $O/r2f_sanitizer_memcheck.log 2>&1
compute-sanitizer --tool racecheck python tools/sanitize_targets.py gpt vqvae > $O/r2f_sanitizer_racecheck.log 2>&1
ls -la $O | grep r2f_
