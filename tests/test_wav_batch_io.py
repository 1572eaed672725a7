"""aecm_wav_batch (the reference CLI as a multi-file batch): file handling that needs no GPU.  --passthrough reads
and rewrites without processing, which pins the WAV parser/writer and the main.cc path conventions on CPU."""
import os
import struct
import subprocess
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "webrtc_aecm_b200", "aecm_wav_batch")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "aecm_run_ref")

pytestmark = pytest.mark.skipif(not os.path.exists(TOOL), reason="aecm_wav_batch not built")


def write_wav(path, samples, fs, width=2, channels=1):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(width)
        w.setframerate(fs)
        s = np.asarray(samples, dtype=np.int64)
        if width == 2:
            raw = s.astype("<i2").tobytes()
        elif width == 1:
            raw = ((s >> 8) + 128).astype(np.uint8).tobytes()
        elif width == 3:
            raw = b"".join(struct.pack("<i", int(v) << 8)[:3] for v in s)
        else:
            raw = (s << 16).astype("<i4").tobytes()
        w.writeframes(raw)


def run(*args):
    return subprocess.run([TOOL, *map(str, args)], capture_output=True, text=True, timeout=120)


def test_passthrough_roundtrip_all_integer_widths(tmp_path):
    rng = np.random.default_rng(5)
    x = rng.integers(-32768, 32768, 16000 + 37).astype(np.int16)
    for width in (1, 2, 3, 4):
        write_wav(tmp_path / f"near{width}.wav", x, 16000, width)
        write_wav(tmp_path / f"far{width}.wav", x[::-1], 16000, width)
    args = []
    for width in (1, 2, 3, 4):
        args += [tmp_path / f"far{width}.wav", tmp_path / f"near{width}.wav"]
    r = run("--passthrough", *args)
    assert r.returncode == 0, r.stderr
    for width in (1, 2, 3, 4):
        raw = open(tmp_path / f"near{width}_out.wav", "rb").read()
        # canonical 44-byte header: RIFF size, fmt 16, PCM, mono, rate, byte rate, align 2, 16 bit, data size
        assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
        assert struct.unpack("<IHHIIHH", raw[16:36]) == (16, 1, 1, 16000, 32000, 2, 16)
        assert struct.unpack("<I", raw[4:8])[0] == len(raw) - 8 and struct.unpack("<I", raw[40:44])[0] == 2 * len(x)
        got = np.frombuffer(raw[44:], dtype="<i2")
        want = x if width > 1 else ((x.astype(np.int32) >> 8) << 8).astype(np.int16)  # 8-bit files keep the top byte
        np.testing.assert_array_equal(got, want)


def test_output_name_list_file_and_explicit_output(tmp_path):
    x = np.arange(800, dtype=np.int16)
    d = tmp_path / "dir.with.dots"
    d.mkdir()
    write_wav(d / "near.take2.wav", x, 8000)
    write_wav(d / "noext", x, 8000)
    write_wav(tmp_path / "far.wav", x, 8000)
    lst = tmp_path / "pairs.txt"
    lst.write_text(f"# far near [out]\n{tmp_path / 'far.wav'} {d / 'noext'}\n{tmp_path / 'far.wav'} {d / 'near.take2.wav'} {tmp_path / 'explicit.wav'}\n")
    r = run("--passthrough", "--list", lst, tmp_path / "far.wav", d / "near.take2.wav")
    assert r.returncode == 0, r.stderr
    # main.cc:189-190: "_out" goes before the extension of the last path component only
    assert (d / "near.take2_out.wav").exists() and (d / "noext_out").exists() and (tmp_path / "explicit.wav").exists()


def test_stereo_is_skipped_and_other_rates_are_written_unchanged(tmp_path):
    x = (np.arange(4410) * 7 % 2000).astype(np.int16)
    write_wav(tmp_path / "near44.wav", x, 44100)
    write_wav(tmp_path / "far44.wav", x, 44100)
    write_wav(tmp_path / "near_st.wav", np.repeat(x, 2), 16000, channels=2)
    write_wav(tmp_path / "far_st.wav", x, 16000)
    r = run("--passthrough", tmp_path / "far44.wav", tmp_path / "near44.wav", tmp_path / "far_st.wav", tmp_path / "near_st.wav")
    assert r.returncode == 0, r.stderr
    assert not (tmp_path / "near_st_out.wav").exists()  # main.cc:46-51, 158-162
    with wave.open(str(tmp_path / "near44_out.wav"), "rb") as w:  # main.cc:109-113 then :169
        assert w.getframerate() == 44100 and w.readframes(w.getnframes()) == x.astype("<i2").tobytes()


def test_bad_usage_and_unreadable_files(tmp_path):
    assert run().returncode == 2
    assert run("--passthrough", tmp_path / "only_far.wav").returncode == 2
    assert run("--passthrough", "--echo-mode", 7, "a", "b").returncode == 2
    (tmp_path / "junk.wav").write_bytes(b"not a wav file at all")
    write_wav(tmp_path / "far.wav", np.zeros(160, np.int16), 16000)
    r = run("--passthrough", tmp_path / "far.wav", tmp_path / "junk.wav")
    assert r.returncode == 1 and "not a RIFF/WAVE file" in r.stderr


def test_no_cuda_device_is_a_hard_error(tmp_path):
    """The tool has no CPU path: on a machine without a GPU it must refuse instead of writing unprocessed audio."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    write_wav(tmp_path / "far.wav", np.zeros(1600, np.int16), 16000)
    write_wav(tmp_path / "near.wav", np.zeros(1600, np.int16), 16000)
    r = run(tmp_path / "far.wav", tmp_path / "near.wav")
    assert r.returncode == 3 and "no CUDA device" in r.stderr
    assert not (tmp_path / "near_out.wav").exists()


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference CLI not built (needs /root/reference at build time)")
def test_header_and_unprocessed_tail_match_reference_cli(tmp_path):
    """The reference CLI (CPU) and the batch tool in --passthrough agree on everything that is not signal processing:
    the header bytes and the trailing partial frame that main.cc leaves untouched."""
    rng = np.random.default_rng(11)
    x = rng.integers(-3000, 3000, 1600 + 77).astype(np.int16)
    for d in ("ref", "ours"):
        (tmp_path / d).mkdir()
        write_wav(tmp_path / d / "near.wav", x, 16000)
        write_wav(tmp_path / d / "far.wav", x, 16000)
    subprocess.run([REF_BIN, str(tmp_path / "ref" / "far.wav"), str(tmp_path / "ref" / "near.wav")], stdin=subprocess.DEVNULL,
                   capture_output=True, timeout=120, check=True)
    assert run("--passthrough", tmp_path / "ours" / "far.wav", tmp_path / "ours" / "near.wav").returncode == 0
    ref = open(tmp_path / "ref" / "near_out.wav", "rb").read()
    ours = open(tmp_path / "ours" / "near_out.wav", "rb").read()
    assert len(ref) == len(ours) and ref[:44] == ours[:44]
    assert ref[44 + 1600 * 2:] == ours[44 + 1600 * 2:] == x[1600:].astype("<i2").tobytes()
