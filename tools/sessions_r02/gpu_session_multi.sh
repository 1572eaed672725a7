# round 2, multi-GPU session (gpurun --gpus N): the product's own multi-GPU entry points on distinct devices, the strong-scaled
# bench under torchrun, the host-link ceiling, the single-process group bench, and (N=8) the 1 M-channel paced run.
N=${1:-2}
mkdir -p gpurun_out
P="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1
(cat /proc/self/status | grep -i "allowed_list"; ls /sys/devices/system/node/ | grep node; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc) > gpurun_out/m${N}_host.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "group" > gpurun_out/m${N}_group_tests.log 2>&1; tail -3 gpurun_out/m${N}_group_tests.log
timeout 900 $P tools/hostlink_probe.py > gpurun_out/m${N}_hostlink.json 2> gpurun_out/m${N}_hostlink.err; cut -c1-900 gpurun_out/m${N}_hostlink.json
timeout 1200 $P bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err; tail -2 gpurun_out/m${N}_bench.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/m${N}_bench.json").read().strip().splitlines()[-1])
    print("N=${N} value %.4g  e2e %.4g  ms/step %.3f  e2e ms/step %.3f  link %s  realtime %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["host_link_gbs"], {k: d["realtime"][k] for k in ("channels","deadline_misses","latency_ms")}))
    print("parity", d["parity_checked"], "host", d["host"])
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 900 python bench.py --group --gpus $N --steps 20 > gpurun_out/m${N}_group.json 2> gpurun_out/m${N}_group.err; cut -c1-400 gpurun_out/m${N}_group.json
timeout 900 python bench.py --group --group-threads 1 --gpus $N --steps 20 > gpurun_out/m${N}_group_threads.json 2>> gpurun_out/m${N}_group.err; cut -c1-400 gpurun_out/m${N}_group_threads.json
if [ "$N" = "8" ]; then
  # the north star's operating point: 131072 channels per GPU x 8 = 1 048 576 concurrent real-time channels, paced
  timeout 1200 $P bench.py --gpus 8 --total-batch 1048576 --steps 20 --warmup 5 > gpurun_out/m8_bench_1m.json 2> gpurun_out/m8_bench_1m.err; tail -2 gpurun_out/m8_bench_1m.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/m8_bench_1m.json").read().strip().splitlines()[-1])
    print("1M: value %.4g e2e %.4g realtime %s" % (d["value"], d["e2e"]["value"], d["realtime"]))
except Exception as e:
    print("1M bench line unreadable:", e)
PY
  for n in 2 4; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $n --steps 20 --warmup 5 --no-realtime > gpurun_out/m8_bench_n$n.json 2> gpurun_out/m8_bench_n$n.err
    python -c "
import json
d=json.loads(open('gpurun_out/m8_bench_n$n.json').read().strip().splitlines()[-1]); print('N=$n value %.4g e2e %.4g link %s'%(d['value'], d['e2e']['value'], d['e2e']['host_link_gbs']))"
  done
fi
ls gpurun_out | grep "^m${N}"
