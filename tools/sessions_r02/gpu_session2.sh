# round 2, GPU session 2: new tests, new bench.py, variants of the persistent kernel, ncu captures for profiles/traffic.json
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2_tests.log
tail -5 gpurun_out/s2_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s2_bench.err; cut -c1-600 gpurun_out/s2_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s2_bench_ref.json 2>> gpurun_out/s2_bench.err; cut -c1-300 gpurun_out/s2_bench_ref.json
: > gpurun_out/s2_variants.txt
for cfg in "--batch 4096 --steps 300" "--batch 32768 --steps 100" "--batch 131072 --steps 40" "--batch 262144 --steps 24" "--batch 65536 --fs 8000 --steps 60"; do
  for lib in default variants/v7.so variants/v8_free.so variants/v8_cta14.so variants/v8_nobin64.so; do
    if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$PWD/$lib; fi
    echo "== $lib $cfg" >> gpurun_out/s2_variants.txt
    timeout 600 python bench.py $cfg --no-e2e --no-cpu-baseline --no-parity 2>>gpurun_out/s2_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> gpurun_out/s2_variants.txt 2>&1
  done
done
unset AECM_B200_LIB
cat gpurun_out/s2_variants.txt
# ncu: full capture at the north-star batch; light captures (traffic + instructions) at the other sizes
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aecm_call --launch-skip 60 --launch-count 2 -o gpurun_out/s2_v8_b262144 -f python bench.py --batch 262144 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/s2_ncu.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active
timeout 600 ncu --metrics $M --clock-control none -k regex:aecm_call --launch-skip 60 --launch-count 4 --csv --log-file gpurun_out/s2_ncu_b131072.csv python bench.py --batch 131072 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity >> gpurun_out/s2_ncu.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:aecm_call --launch-skip 15 --launch-count 4 --csv --log-file gpurun_out/s2_ncu_b32768.csv python bench.py --batch 32768 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity >> gpurun_out/s2_ncu.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:aecm_call --launch-skip 15 --launch-count 4 --csv --log-file gpurun_out/s2_ncu_b4096.csv python bench.py --batch 4096 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity >> gpurun_out/s2_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aecm_call --launch-skip 115 --launch-count 3 -o gpurun_out/s2_v8_b65536_8k -f python bench.py --batch 65536 --fs 8000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity >> gpurun_out/s2_ncu.log 2>&1
# launch list of the default bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-realtime > gpurun_out/s2_launches_bench.log 2>&1
ls -la gpurun_out/ | tail -20
