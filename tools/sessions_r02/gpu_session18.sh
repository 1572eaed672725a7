# round 2, session 18: the final tree -- GPU suite (with the mid-stream control-call scripts), then both bench arms as the driver runs them
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s18_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s18_tests.log; tail -6 gpurun_out/s18_tests.log
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/s18_bench.json
