# A/B of kernel build variants (variants/*.so, selected through AECM_B200_LIB) on the configurations given below.
mkdir -p gpurun_out
: > gpurun_out/variants.txt
for lib in default $VARIANTS; do
  for cfg in "--batch 4096" "--batch 32768 --steps 200" "--batch 65536 --fs 8000 --steps 100"; do
    if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$lib; fi
    echo "== $lib $cfg" >> gpurun_out/variants.txt
    python bench.py $cfg --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> gpurun_out/variants.txt 2>&1
  done
done
cat gpurun_out/variants.txt
