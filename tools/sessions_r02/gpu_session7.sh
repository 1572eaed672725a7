# round 2, GPU session 7: verification of the committed tree -- GPU suite (default grid and one-CTA grid: every batch walks
# many rows per warp), smoke, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s7_tests.log; tail -3 gpurun_out/s7_tests.log
AECM_B200_MAX_CTAS=1 timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/s7_tests_one_cta.log 2>&1; echo "tests (AECM_B200_MAX_CTAS=1) rc=$?" >> gpurun_out/s7_tests_one_cta.log; tail -3 gpurun_out/s7_tests_one_cta.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s7_smoke.log 2>&1; tail -2 gpurun_out/s7_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s7_bench_ref.json 2> gpurun_out/s7_err.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s7_bench.json 2>> gpurun_out/s7_err.txt; echo "bench rc=$?"
cut -c1-300 gpurun_out/s7_bench.json; cut -c1-200 gpurun_out/s7_bench_ref.json; tail -3 gpurun_out/s7_err.txt
