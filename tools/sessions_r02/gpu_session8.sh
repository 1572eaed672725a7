mkdir -p gpurun_out
gcc -O2 -I include examples/delay_estimator_options.c -L webrtc_aecm_b200 -laecm_b200 -Wl,-rpath,$PWD/webrtc_aecm_b200 -o /tmp/deo && /tmp/deo > gpurun_out/s8_example.txt 2>&1; echo "example rc=$?" >> gpurun_out/s8_example.txt; cat gpurun_out/s8_example.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s8_tests.log; tail -4 gpurun_out/s8_tests.log
AECM_B200_MAX_CTAS=1 timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/s8_tests_one_cta.log 2>&1; echo "tests (AECM_B200_MAX_CTAS=1) rc=$?" >> gpurun_out/s8_tests_one_cta.log; tail -4 gpurun_out/s8_tests_one_cta.log
