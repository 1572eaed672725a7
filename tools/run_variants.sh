mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "scalar" > gpurun_out/t10.log 2>&1; echo "tests rc=$?" >> gpurun_out/t10.log
for t in 1 16 64 256; do ./webrtc_aecm_b200/scalar_api_bench $t 400; done > gpurun_out/scalar_bench2.jsonl 2>&1
: > gpurun_out/variants.txt
for lib in default variants/lib_w14.so variants/lib_w7.so; do
  for cfg in "--batch 4096" "--batch 32768 --steps 200" "--batch 65536 --fs 8000 --steps 100"; do
    if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$lib; fi
    echo "== $lib $cfg" >> gpurun_out/variants.txt
    python bench.py $cfg --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> gpurun_out/variants.txt 2>&1
  done
done
tail -3 gpurun_out/t10.log; cat gpurun_out/scalar_bench2.jsonl | cut -c1-330; cat gpurun_out/variants.txt
