# round 2, session 17: GPU suite with the randomised API scripts
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s17_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s17_tests.log; tail -6 gpurun_out/s17_tests.log
