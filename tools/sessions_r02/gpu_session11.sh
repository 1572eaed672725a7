# does polling NVML every millisecond from a bench thread disturb the timed kernels?  A/B against the nvidia-smi sampler
mkdir -p gpurun_out
: > gpurun_out/s11_sampler.txt
for rep in 1 2; do
for cfg in "--batch 131072 --steps 40" "--batch 32768 --steps 100" "--batch 4096 --steps 300"; do
  for smp in nvml smi; do
    echo "== sampler=$smp $cfg" >> gpurun_out/s11_sampler.txt
    timeout 600 python bench.py $cfg --clock-sampler $smp --no-e2e --no-cpu-baseline --no-parity 2>>gpurun_out/s11_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step  clock samples %d'%(d['value'], d['ms_per_step']*1e3, d['clocks']['samples']))" >> gpurun_out/s11_sampler.txt 2>&1
  done
done
done
cat gpurun_out/s11_sampler.txt
