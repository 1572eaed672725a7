# round 2, GPU session 6: bench lines at the BASELINE configs (NVML clock sampler), reference arm
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s6_bench_c4.json 2> gpurun_out/s6_err.txt; echo "c4 rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s6_bench_ref.json 2>> gpurun_out/s6_err.txt
timeout 900 python bench.py --batch 4096 --steps 300 --warmup 10 --realtime-seconds 3 > gpurun_out/s6_bench_c2.json 2>> gpurun_out/s6_err.txt; echo "c2 rc=$?"
timeout 900 python bench.py --batch 65536 --fs 8000 --steps 60 --warmup 5 --realtime-seconds 3 --no-cpu-baseline > gpurun_out/s6_bench_c3.json 2>> gpurun_out/s6_err.txt; echo "c3 rc=$?"
timeout 900 python bench.py --batch 131072 --steps 40 --warmup 5 --realtime-seconds 3 --no-cpu-baseline > gpurun_out/s6_bench_b131072.json 2>> gpurun_out/s6_err.txt; echo "b131072 rc=$?"
python - <<PY
import json
for f in ("c4","c2","c3","b131072"):
    try:
        d=json.loads(open(f"gpurun_out/s6_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g parity %s clocks %s rt %s issue %s traffic_frac %s frac %.3f" % (d["value"], d["e2e"]["value"], d["parity_checked"], d["clocks"], {k: d["realtime"][k] for k in ("deadline_misses","latency_ms")}, (d["roofline"]["issue"] or {}).get("frac"), d["roofline"]["traffic_frac"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/s6_err.txt
