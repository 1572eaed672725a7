"""the WAV front-end as a multi-file batch.  Every output of aecm_wav_batch must be byte-identical to
the file the reference CLI (unmodified main.cc + reference library, oracle/_ref/aecm_run_ref) writes for the same
pair: mixed 8/16 kHz rates, ragged lengths (whole batch advances in lockstep, instances drop out as they end), a
trailing partial frame, an unsupported rate, and a far file shorter than the near file is the one stated difference."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from test_wav_batch_io import REF_BIN, TOOL, write_wav
from webrtc_aecm_b200.synth import synth_streams

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(TOOL)), reason="CLI binaries not built")]


def make_pairs(root, n, seed):
    rng = np.random.default_rng(seed)
    pairs = []
    for i in range(n):
        fs = 16000 if i % 3 else 8000
        n_samples = int(fs * rng.uniform(0.8, 4.0)) + int(rng.integers(0, fs // 100))  # ragged, with a partial last frame
        far, near = synth_streams(1, n_samples, fs, seed=seed * 100 + i, delay_ms=int(rng.integers(0, 100)))
        d = root / f"p{i:02d}"
        d.mkdir()
        write_wav(d / "far.wav", far[0], fs)
        write_wav(d / "near.wav", near[0], fs)
        pairs.append(d)
    return pairs


def reference_outputs(pairs, tmp_path):
    want = {}
    for d in pairs:
        r = tmp_path / "ref" / d.name
        r.mkdir(parents=True)
        shutil.copy(d / "far.wav", r / "far.wav")
        shutil.copy(d / "near.wav", r / "near.wav")
        subprocess.run([REF_BIN, str(r / "far.wav"), str(r / "near.wav")], stdin=subprocess.DEVNULL, capture_output=True, timeout=300, check=True)
        want[d.name] = open(r / "near_out.wav", "rb").read()
    return want


@pytest.mark.parametrize("extra", [[], ["--devices", "0,0,0", "--chunk-calls", "37"]], ids=["one_shard", "three_shards_odd_chunks"])
def test_batch_outputs_are_byte_identical_to_the_reference_cli(tmp_path, extra):
    pairs = make_pairs(tmp_path, 14, seed=31)
    # an unsupported rate (written back unchanged by both) rides along
    odd = tmp_path / "odd"
    odd.mkdir()
    x = (np.arange(22050) * 13 % 5000 - 2500).astype(np.int16)
    write_wav(odd / "far.wav", x, 22050)
    write_wav(odd / "near.wav", x[::-1], 22050)
    pairs.append(odd)
    want = reference_outputs(pairs, tmp_path)
    args = []
    for d in pairs:
        args += [str(d / "far.wav"), str(d / "near.wav")]
    res = subprocess.run([TOOL, *extra, *args], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr + res.stdout
    changed = 0
    for d in pairs:
        got = open(d / "near_out.wav", "rb").read()
        assert got == want[d.name], d.name
        changed += got[44:] != open(d / "near.wav", "rb").read()[44:]
    assert changed == len(pairs) - 1  # every supported pair was actually processed
    assert "15 pairs written" in res.stdout


def test_echo_mode_and_ms_options_reach_the_instances(tmp_path):
    """--echo-mode 3 --ms 20 must equal the oracle run with that configuration (main.cc hard-codes 1 and 40)."""
    import oracle_lib as ol

    fs, n = 16000, 16000 * 2
    far, near = synth_streams(3, n, fs, seed=77, delay_ms=30)
    args = []
    for i in range(3):
        write_wav(tmp_path / f"far{i}.wav", far[i], fs)
        write_wav(tmp_path / f"near{i}.wav", near[i], fs)
        args += [str(tmp_path / f"far{i}.wav"), str(tmp_path / f"near{i}.wav")]
    res = subprocess.run([TOOL, "--echo-mode", "3", "--ms", "20", "--cng", "0", *args], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    want, _, _ = ol.run_streams(far, near, fs=fs, impl="oracle", echo_mode=3, cng_mode=0, ms=20)
    for i in range(3):
        got = np.frombuffer(open(tmp_path / f"near{i}_out.wav", "rb").read()[44:], dtype="<i2")
        np.testing.assert_array_equal(got, want[i])


def test_one_unreadable_pair_costs_only_its_own_output(tmp_path):
    """Per-pair semantics, as in the reference CLI (which simply skips a pair it cannot read): a junk file among good
    pairs must not withhold the good pairs' outputs.  The tool still exits 1 and names what was lost."""
    pairs = make_pairs(tmp_path, 4, seed=9)
    want = reference_outputs(pairs, tmp_path)
    (tmp_path / "junk_far.wav").write_bytes(b"this is not a RIFF file")
    (tmp_path / "junk_near.wav").write_bytes(b"RIFF\x00\x00\x00\x00nope")
    args = []
    for d in pairs[:2]:
        args += [str(d / "far.wav"), str(d / "near.wav")]
    args += [str(tmp_path / "junk_far.wav"), str(tmp_path / "junk_near.wav")]
    for d in pairs[2:]:
        args += [str(d / "far.wav"), str(d / "near.wav")]
    res = subprocess.run([TOOL, *args], capture_output=True, text=True, timeout=300)
    assert res.returncode == 1, (res.stdout, res.stderr)
    assert "1 pair(s) unreadable" in res.stderr and "4 pairs written" in res.stdout
    assert not (tmp_path / "junk_near_out.wav").exists()
    for d in pairs:
        assert open(d / "near_out.wav", "rb").read() == want[d.name], d.name
