# round 2, session 13: the committed tree -- GPU suite (default grid, one-CTA grid), sanitizers on the row loop, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s13_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s13_tests.log; tail -3 gpurun_out/s13_tests.log
AECM_B200_MAX_CTAS=1 timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/s13_tests_one_cta.log 2>&1; echo "tests (AECM_B200_MAX_CTAS=1) rc=$?" >> gpurun_out/s13_tests_one_cta.log; tail -3 gpurun_out/s13_tests_one_cta.log
S=gpurun_out/s13_sanitizer.txt; : > $S
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2" >> $S
  timeout 600 compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2 2>&1 | grep -E "^OK|SUMMARY|DIVERG|Error|hazard" | head -8 >> $S
  echo "== compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2 --clean --fs 8000" >> $S
  timeout 600 compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2 --clean --fs 8000 2>&1 | grep -E "^OK|SUMMARY|DIVERG|Error|hazard" | head -8 >> $S
  echo "== AECM_B200_MAX_CTAS=1 compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 70 --seconds 0.6   (one CTA walks three rows per warp, the last one partial)" >> $S
  AECM_B200_MAX_CTAS=1 timeout 900 compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 70 --seconds 0.6 2>&1 | grep -E "^OK|SUMMARY|DIVERG|Error|hazard" | head -8 >> $S
done
echo "== compute-sanitizer --tool memcheck pytest -k 'index or fused or ragged or pipelined or delay_estimator_options_fused or fft or touch'" >> $S
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "index or fused or ragged or pipelined or delay_estimator_options_fused or fft or touch" 2>&1 | grep -E "passed|failed|SUMMARY" | head -5 >> $S
echo "== AECM_B200_MAX_CTAS=1 compute-sanitizer --tool synccheck pytest -k 'index or ragged'  (rows with a bad instanceIndex and warps without a row pass the rendezvous on their own paths)" >> $S
AECM_B200_MAX_CTAS=1 timeout 1200 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -q -x -k "index or ragged" 2>&1 | grep -E "passed|failed|SUMMARY" | head -5 >> $S
cat $S
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s13_bench_ref.json 2> gpurun_out/s13_err.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s13_bench.json 2>> gpurun_out/s13_err.txt; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/s13_bench.json').read().strip().splitlines()[-1]); print('value %.4g e2e %.4g parity %s rt %s issue %s'%(d['value'], d['e2e']['value'], d['parity_checked'], d['realtime']['latency_ms'], d['roofline']['issue']['frac']))
d=json.loads(open('gpurun_out/s13_bench_ref.json').read().strip().splitlines()[-1]); print('ref %.4g'%d['value'])"
