# round 2, GPU session 4: full suite again, bench with sub-batched realtime ticks, rendezvous-period variants, traffic at 65536@16k
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s4_tests.log
tail -3 gpurun_out/s4_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s4_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/s4_bench.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g parity %s realtime %s" % (d["value"], d["e2e"]["value"], d["parity_checked"], {k: d["realtime"][k] for k in ("deadline_misses","latency_ms","sub_batches_per_tick")}))
print("roofline", {k: d["roofline"][k] for k in ("frac","traffic_frac","issue")})
PY
: > gpurun_out/s4_variants.txt
for cfg in "--batch 262144 --steps 24" "--batch 32768 --steps 100" "--batch 65536 --fs 8000 --steps 60"; do
  for lib in default variants/v8_resync2.so variants/v8_resync4.so; do
    if [ "$lib" = default ]; then unset AECM_B200_LIB; else export AECM_B200_LIB=$PWD/$lib; fi
    echo "== $lib $cfg" >> gpurun_out/s4_variants.txt
    timeout 600 python bench.py $cfg --no-e2e --no-cpu-baseline --no-parity 2>>gpurun_out/s4_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g blocks/s  %.2f us/step'%(d['value'], d['ms_per_step']*1e3))" >> gpurun_out/s4_variants.txt 2>&1
  done
done
unset AECM_B200_LIB
cat gpurun_out/s4_variants.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:aecm_call --launch-skip 56 --launch-count 4 --csv --log-file gpurun_out/s4_ncu_b65536.csv python bench.py --batch 65536 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/s4_ncu.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:aecm_call --launch-skip 56 --launch-count 4 --csv --log-file gpurun_out/s4_ncu_b262144.csv python bench.py --batch 262144 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-parity >> gpurun_out/s4_ncu.log 2>&1
grep -c aecm_call gpurun_out/s4_ncu_b65536.csv gpurun_out/s4_ncu_b262144.csv
