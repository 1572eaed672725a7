"""Host-side multi-rank logic on CPU: instance sharding over ranks (gloo, world_size 2) reproduces the
unsharded result exactly and needs no data-path collective -- only the checksum gather used for reporting."""
import os
import subprocess
import sys

import numpy as np

from webrtc_aecm_b200.shard import shard_range, stream_checksum

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 262144, 1000003):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


WORKER = r'''
import os, sys, json
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["AECM_ROOT"]); sys.path.insert(0, os.path.join(os.environ["AECM_ROOT"], "tests"))
import oracle_lib as ol
from webrtc_aecm_b200.shard import shard_range, stream_checksum
from webrtc_aecm_b200.synth import synth_streams
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, fs, calls = 11, 16000, 40
first, count = shard_range(N, rank, world)
# every rank generates only ITS slice of the one global stream family (first_instance) ...
far, near = synth_streams(count, calls * 160, fs, seed=606, delay_ms=40, first_instance=first)
out, ret, _ = ol.run_streams(far, near, fs=fs, impl="oracle", threads=2)   # ... and processes it independently
sums = [None] * world
dist.all_gather_object(sums, (first, count, stream_checksum(out), out.tobytes()))
if rank == 0:
    whole = np.concatenate([np.frombuffer(b, dtype=np.int16).reshape(c, -1) for _, c, _, b in sorted(sums)])
    print("RESULT " + json.dumps({"parts": [(f, c, s) for f, c, s, _ in sorted(sums)], "whole": stream_checksum(whole)}))
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding_equals_unsharded(tmp_path):
    import json

    import oracle_lib as ol
    from webrtc_aecm_b200.synth import synth_streams

    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, AECM_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    got = json.loads(line[len("RESULT "):])
    far, near = synth_streams(11, 40 * 160, 16000, seed=606, delay_ms=40)
    want, _, _ = ol.run_streams(far, near, fs=16000, impl="oracle")
    assert got["whole"] == stream_checksum(want)
    assert [(f, c) for f, c, _ in got["parts"]] == [shard_range(11, r, 2) for r in range(2)]
    for f, c, s in got["parts"]:
        assert s == stream_checksum(want[f:f + c])


def test_synth_is_shard_and_chunk_invariant():
    from webrtc_aecm_b200.synth import synth_streams

    far, near = synth_streams(6, 4000, 16000, seed=5, delay_ms=40)
    f2, n2 = synth_streams(2, 4000, 16000, seed=5, delay_ms=40, first_instance=3)
    np.testing.assert_array_equal(far[3:5], f2)
    np.testing.assert_array_equal(near[3:5], n2)
    f3, n3 = synth_streams(6, 1500, 16000, seed=5, delay_ms=40, first_sample=2500)
    np.testing.assert_array_equal(far[:, 2500:], f3)
    np.testing.assert_array_equal(near[:, 2500:], n3)
