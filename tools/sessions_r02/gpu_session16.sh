# round 2, session 16: the new f4 GPU cases on the committed tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s16_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s16_tests.log; tail -4 gpurun_out/s16_tests.log
