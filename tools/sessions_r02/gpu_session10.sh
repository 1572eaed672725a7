# round 2, session 10: constant tables in shared memory (one copy per CTA) vs global/L1
mkdir -p gpurun_out
: > gpurun_out/s10_variants.txt
for cfg in "--batch 262144 --steps 24" "--batch 32768 --steps 100" "--batch 4096 --steps 300" "--batch 65536 --fs 8000 --steps 60"; do
  for lib in default variants/v8_smemtab.so default variants/v8_smemtab.so; do
    if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$PWD/$lib; fi
    echo "== $lib $cfg" >> gpurun_out/s10_variants.txt
    timeout 600 python bench.py $cfg --no-e2e --no-cpu-baseline --no-parity 2>>gpurun_out/s10_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> gpurun_out/s10_variants.txt 2>&1
  done
done
cat gpurun_out/s10_variants.txt
AECM_B200_LIB=$PWD/variants/v8_smemtab.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cuda_matches_oracle or fft or fused or index" > gpurun_out/s10_tests_smemtab.log 2>&1; tail -2 gpurun_out/s10_tests_smemtab.log
