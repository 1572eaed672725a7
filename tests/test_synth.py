"""The synthetic stream generator exists twice: numpy (tests, CPU reference arm) and torch (bench.py fills HBM with it).
Both must produce the same bits, or the bench's in-run parity replay would compare different streams."""
import numpy as np
import pytest

from webrtc_aecm_b200.synth import synth_streams, synth_streams_torch


@pytest.mark.parametrize("kind", ["echo", "stress"])
@pytest.mark.parametrize("fs", [8000, 16000])
def test_torch_twin_equals_numpy_generator(kind, fs):
    n, T = 9, 3 * fs // 4 + 37
    delay = (np.arange(n) * 13) % 121
    f, m = synth_streams(n, T, fs, 0xAEC2, delay, kind, first_instance=1000, first_sample=777)
    ft, mt = synth_streams_torch(n, T, fs, 0xAEC2, delay, kind, first_instance=1000, first_sample=777, device="cpu")
    np.testing.assert_array_equal(f, ft.numpy())
    np.testing.assert_array_equal(m, mt.numpy())
    assert f.dtype == np.int16 and np.abs(f.astype(np.int32)).max() > 1000  # not silence


def test_slices_of_the_global_stream_agree():
    """Shards (first_instance), chunks (first_sample) and picked rows (instances=) are views of ONE global stream."""
    fs, T = 16000, 6000
    f, m = synth_streams(8, T, fs, 7, 40, "echo")
    f2, m2 = synth_streams(3, T, fs, 7, 40, "echo", first_instance=5)
    np.testing.assert_array_equal(f[5:], f2)
    np.testing.assert_array_equal(m[5:], m2)
    f3, m3 = synth_streams(8, 2000, fs, 7, 40, "echo", first_sample=4000)
    np.testing.assert_array_equal(f[:, 4000:], f3)
    np.testing.assert_array_equal(m[:, 4000:], m3)
    f4, m4 = synth_streams(3, T, fs, 7, 40, "echo", instances=[6, 0, 3])
    np.testing.assert_array_equal(f[[6, 0, 3]], f4)
    np.testing.assert_array_equal(m[[6, 0, 3]], m4)
    f5, m5 = synth_streams_torch(3, 2000, fs, 7, 40, "echo", first_sample=4000, device="cpu", instances=[6, 0, 3])
    np.testing.assert_array_equal(f[[6, 0, 3], 4000:], f5.numpy())
    np.testing.assert_array_equal(m[[6, 0, 3], 4000:], m5.numpy())
