# round 2, session 20 (the last GPU seconds of the round): smoke() and a handful of parity tests on the final tree
mkdir -p gpurun_out
(python -c "import __graft_entry__ as g; g.smoke()"; timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_calls or ragged or device_fft or random_control") > gpurun_out/s20_final.log 2>&1
tail -5 gpurun_out/s20_final.log
