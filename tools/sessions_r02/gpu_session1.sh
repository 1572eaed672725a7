# round 2, GPU session 1: does the persistent kernel pass the GPU suite, and how do the variants compare?
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/s1_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s1_tests.log
tail -3 gpurun_out/s1_tests.log
: > gpurun_out/s1_variants.txt
for cfg in "--batch 4096 --steps 300" "--batch 32768 --steps 100" "--batch 131072 --steps 40" "--batch 262144 --steps 24" "--batch 65536 --fs 8000 --steps 60"; do
  for lib in default variants/v7.so variants/v8_resync.so variants/v8_nobin64.so; do
    if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$PWD/$lib; fi
    echo "== $lib $cfg" >> gpurun_out/s1_variants.txt
    timeout 600 python bench.py $cfg --no-e2e --no-cpu-baseline 2>>gpurun_out/s1_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> gpurun_out/s1_variants.txt 2>&1
  done
done
unset AECM_B200_LIB
cat gpurun_out/s1_variants.txt
# ncu: launch list + one full capture of the hot kernel at the north-star batch (skip the warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aecm_call --launch-skip 60 --launch-count 2 -o gpurun_out/s1_v8_b262144 -f python bench.py --batch 262144 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s1_ncu.log 2>&1
tail -2 gpurun_out/s1_ncu.log
ls -la gpurun_out/
