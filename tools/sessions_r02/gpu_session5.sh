# round 2, GPU session 5: finer rendezvous points (per frame / per potential block) vs per row only
mkdir -p gpurun_out
: > gpurun_out/s5_variants.txt
for cfg in "--batch 262144 --steps 24" "--batch 32768 --steps 100" "--batch 4096 --steps 300" "--batch 65536 --fs 8000 --steps 60"; do
  for lib in default variants/v8_syncframe.so variants/v8_syncblock.so; do
    if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$PWD/$lib; fi
    echo "== $lib $cfg" >> gpurun_out/s5_variants.txt
    timeout 600 python bench.py $cfg --no-e2e --no-cpu-baseline --no-parity 2>>gpurun_out/s5_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> gpurun_out/s5_variants.txt 2>&1
  done
done
cat gpurun_out/s5_variants.txt
for lib in variants/v8_syncframe.so variants/v8_syncblock.so; do
  AECM_B200_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scalar_api.py -m gpu -x -q -k "cuda_matches_oracle or ragged or index or fused or pipelined or threads or error_codes" > gpurun_out/s5_tests_$(basename $lib .so).log 2>&1
  tail -2 gpurun_out/s5_tests_$(basename $lib .so).log
done
