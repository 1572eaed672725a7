# round 2, session 12: issue the row's bulk loads before the CTA rendezvous (v8_loadfirst) vs after it (default)
mkdir -p gpurun_out
: > gpurun_out/s12_variants.txt
for cfg in "--batch 262144 --steps 24" "--batch 32768 --steps 100" "--batch 65536 --fs 8000 --steps 60"; do
  for lib in default variants/v8_loadfirst.so default variants/v8_loadfirst.so; do
    if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$PWD/$lib; fi
    echo "== $lib $cfg" >> gpurun_out/s12_variants.txt
    timeout 600 python bench.py $cfg --no-e2e --no-cpu-baseline --no-parity 2>>gpurun_out/s12_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> gpurun_out/s12_variants.txt 2>&1
  done
done
cat gpurun_out/s12_variants.txt
AECM_B200_LIB=$PWD/variants/v8_loadfirst.so AECM_B200_MAX_CTAS=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scalar_api.py -m gpu -x -q -k "cuda_matches_oracle or fused or index or ragged or 4096 or threads" > gpurun_out/s12_tests_loadfirst.log 2>&1; tail -2 gpurun_out/s12_tests_loadfirst.log
