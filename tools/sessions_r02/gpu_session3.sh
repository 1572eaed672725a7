# round 2, GPU session 3: full GPU suite with the f4 tests, sanitizers on the new kernels, launch list of the bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3_tests.log
tail -5 gpurun_out/s3_tests.log
S=gpurun_out/s3_sanitizer.txt; : > $S
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2" >> $S
  timeout 600 compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2 2>&1 | grep -E "^OK|SUMMARY|DIVERG|Error|hazard" | head -8 >> $S
  echo "== compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2 --clean --fs 8000" >> $S
  timeout 600 compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2 --clean --fs 8000 2>&1 | grep -E "^OK|SUMMARY|DIVERG|Error|hazard" | head -8 >> $S
  echo "== compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 70 --seconds 0.6   (AECM_B200_MAX_CTAS=1: one CTA walks three rows per warp, the last one partial)" >> $S
  AECM_B200_MAX_CTAS=1 timeout 900 compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 70 --seconds 0.6 2>&1 | grep -E "^OK|SUMMARY|DIVERG|Error|hazard" | head -8 >> $S
done
echo "== compute-sanitizer --tool memcheck pytest -k 'index or fused or ragged or pipelined or delay_estimator_options_fused or fft'" >> $S
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "index or fused or ragged or pipelined or delay_estimator_options_fused or fft" 2>&1 | grep -E "passed|failed|SUMMARY" | head -5 >> $S
echo "== compute-sanitizer --tool racecheck pytest -k 'delay_estimator_options_fused'" >> $S
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "delay_estimator_options_fused" 2>&1 | grep -E "passed|failed|SUMMARY|hazard" | head -5 >> $S
cat $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aecm_ --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-realtime > gpurun_out/s3_launches_bench.log 2>&1
grep -c aecm_call gpurun_out/s3_launches.csv
