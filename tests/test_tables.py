"""The generated constant tables are current and equal the reference's tables (when the reference is here)."""
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_tables as g  # noqa: E402

REF = "/root/reference/aecm"


def test_generated_headers_are_current():
    assert g.main(check=True) == 0


def _ref_array(fname, name):
    src = open(os.path.join(REF, fname)).read()
    m = re.search(name + r"\[[^\]]*\]\s*(?:ALIGN8_END)?\s*=\s*\{([^}]*)\}", src, re.S)
    return [int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_tables_equal_reference():
    sin = _ref_array("complex_fft.c", "kSinTable1024")
    assert [g.sin1024(k) for k in range(1024)] == sin
    assert g.sqrt_hanning() == _ref_array("aecm_core_c.cc", "WebRtcAecm_kSqrtHanning")
    assert g.cos360() == _ref_array("aecm_core.cc", "WebRtcAecm_kCosTable")
    assert g.sin360() == _ref_array("aecm_core.cc", "WebRtcAecm_kSinTable")
    assert g.CHANNEL_8K == _ref_array("aecm_core.cc", "kChannelStored8kHz")
    assert g.CHANNEL_16K == _ref_array("aecm_core.cc", "kChannelStored16kHz")


def test_twiddle_packing():
    tw = [int(x.rstrip("u"), 16) for x in g.twiddle_pairs()]
    assert len(tw) == 64 and tw[0] == 32767  # cos(0) = 32767, sin(0) = 0
    assert tw[32] == (32767 << 16)           # t=32: j=256 -> cos = S[512] = 0, sin = S[256] = 32767
