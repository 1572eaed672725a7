# round 2, session 9 (2 GPUs): the whole GPU suite where two devices are visible, and the strong-scaled bench under torchrun
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s9_tests.log; tail -4 gpurun_out/s9_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s9_bench_n2.json 2> gpurun_out/s9_err.txt; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/s9_bench_n2.json').read().strip().splitlines()[-1]); print('N=2 value %.4g e2e %.4g parity %s clocks %s rt %s'%(d['value'], d['e2e']['value'], d['parity_checked'], d['clocks'], d['realtime']['latency_ms']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/s9_bench_ref_n2.json 2>> gpurun_out/s9_err.txt; cut -c1-200 gpurun_out/s9_bench_ref_n2.json
tail -3 gpurun_out/s9_err.txt
