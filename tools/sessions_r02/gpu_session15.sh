# round 2, session 15 (8 GPUs): the committed tree at N=8 -- strong-scaled bench, then the 1 048 576-channel paced run
mkdir -p gpurun_out
P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 400 $P bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/s15_bench_n8.json 2> gpurun_out/s15_err.txt; echo "n8 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/s15_bench_n8.json').read().strip().splitlines()[-1]); print('N=8 value %.4g e2e %.4g link %s parity %s clocks %s rt %s'%(d['value'], d['e2e']['value'], d['e2e']['host_link_gbs'], d['parity_checked'], d['clocks'], d['realtime']['latency_ms']))"
timeout 400 $P bench.py --gpus 8 --total-batch 1048576 --steps 20 --warmup 5 > gpurun_out/s15_bench_1m.json 2>> gpurun_out/s15_err.txt; echo "1m rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/s15_bench_1m.json').read().strip().splitlines()[-1]); print('1M value %.4g e2e %.4g rt %s'%(d['value'], d['e2e']['value'], {k: d['realtime'][k] for k in ('channels','deadline_misses','late_releases','latency_ms')}))"
