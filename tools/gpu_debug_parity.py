#!/usr/bin/env python3
"""Step a few instances call by call on the GPU and in the oracle; report the first divergence in output
or in live state (field names).  Debugging aid, not part of the test-suite."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402
from webrtc_aecm_b200 import AecmBatch  # noqa: E402
from webrtc_aecm_b200.synth import synth_streams  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--fs", type=int, default=16000)
    ap.add_argument("--seconds", type=float, default=6.0)
    ap.add_argument("--kind", default="echo")
    ap.add_argument("--clean", action="store_true")
    ap.add_argument("--echo-mode", type=int, default=3)
    ap.add_argument("--cng", type=int, default=1)
    ap.add_argument("--max-report", type=int, default=3)
    args = ap.parse_args()
    fs, n = args.fs, args.n
    spc = 160 if fs == 16000 else 80
    calls = int(args.seconds * 100)
    far, near = synth_streams(n, calls * spc, fs, seed=11, delay_ms=np.arange(n) * 20 % 120, kind=args.kind)
    clean = synth_streams(n, calls * spc, fs, seed=12, delay_ms=10)[1] if args.clean else None

    b = AecmBatch(n, 0)
    assert b.init(fs) == 0
    rc, st = b.set_config(args.cng, args.echo_mode)
    assert rc == 0 and (st == 0).all(), (rc, st)
    orcs = [ol.OracleInstance(fs, args.cng, args.echo_mode) for _ in range(n)]
    # state right after init
    for i in range(n):
        bad = ol.compare_states(b.read_state(i), orcs[i].state(), args.clean)
        if bad:
            print(f"[init] inst {i}:", *bad[:10], sep="\n   ")
            return 1
    reported = 0
    t0 = time.time()
    for c in range(calls):
        sl = slice(c * spc, (c + 1) * spc)
        f = np.ascontiguousarray(far[:, sl])
        nr = np.ascontiguousarray(near[:, sl])
        cl = np.ascontiguousarray(clean[:, sl]) if clean is not None else None
        rc, out, status = b.process_stream(f, nr, cl, samples_per_call=spc)
        assert rc == 0, rc
        for i in range(n):
            orcs[i].buffer_farend(f[i])
            orc_rc, o = orcs[i].process(nr[i], None if cl is None else cl[i], 40)
            ok = (o == out[i]).all() and orc_rc == status[i]
            bad = ol.compare_states(b.read_state(i), orcs[i].state(), args.clean)
            if not ok or bad:
                print(f"[call {c}] inst {i}: out_equal={(o == out[i]).all()} first_bad_sample="
                      f"{int(np.argmax(o != out[i])) if not (o == out[i]).all() else -1} rc dev={status[i]} oracle={orc_rc}")
                for line in bad[:12]:
                    print("    ", line)
                reported += 1
                if reported >= args.max_report:
                    return 1
    print(f"OK: {n} instances x {calls} calls bit-exact in output and state ({time.time() - t0:.1f}s)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
