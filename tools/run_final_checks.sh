# Round-end evidence on the GPU box: full GPU suite, compute-sanitizer on the hot kernel, both bench arms.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/final_tests.log
S=gpurun_out/final_sanitizer.txt; : > $S
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2" >> $S
  timeout 600 compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2 2>&1 | grep -E "^OK|SUMMARY|DIVERG|Error|hazard" | head -8 >> $S
  echo "== compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2 --clean --fs 8000" >> $S
  timeout 600 compute-sanitizer --tool $tool python tools/gpu_debug_parity.py --n 3 --seconds 1.2 --clean --fs 8000 2>&1 | grep -E "^OK|SUMMARY|DIVERG|Error|hazard" | head -8 >> $S
done
echo "== compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -k 'index or fused or ragged or pipelined'" >> $S
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "index or fused or ragged or pipelined" 2>&1 | grep -E "passed|failed|SUMMARY" | head -5 >> $S
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench.err
python bench.py > gpurun_out/final_bench.json 2>> gpurun_out/final_bench.err
tail -2 gpurun_out/final_tests.log; cat $S; cut -c1-250 gpurun_out/final_bench_reference.json; cut -c1-250 gpurun_out/final_bench.json
