"""Pin the oracle's integer/host parts against golden vectors generated from the REFERENCE's own
headers (tools/make_golden.py via oracle/ref_shim.cpp): SDK/cuda/random.h:30-67,
SDK/sutil/WorkDistribution.h:50-81, SDK/sutil/Camera.cpp:34-46."""
import json
import pathlib
import struct

import numpy as np

from oracle import pyoracle as orc

KAT = json.loads((pathlib.Path(__file__).parent / "golden" / "kat.json").read_text())


def bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def test_tea4_and_rnd_streams_bit_exact():
    for e in KAT["tea4_rnd"]:
        seed = orc.tea4(e["v0"], e["v1"])
        assert seed == e["seed"], (e["v0"], e["v1"])
        vals, state = orc.rnd_stream(seed, len(e["rnd_bits"]))
        assert [bits(v) for v in vals] == e["rnd_bits"]
        assert state == e["state_after"]


def test_survey_known_answers():
    # SURVEY.md §8(c) derived vectors
    assert orc.tea4(0, 0) == 0x5DF5F2BF
    assert orc.tea4(589823, 15) == 0xCE0B2619
    vals, state = orc.rnd_stream(0x5DF5F2BF, 3)
    assert state == 0xE6E5B614
    np.testing.assert_allclose(vals, [0.294449925, 0.695515215, 0.897309542], rtol=0, atol=1e-9)


def test_lcg_from_zero():
    import ctypes as C
    st = C.c_uint32(0)
    for ret, state in KAT["lcg_from_0"]:
        assert orc.lib().orc_lcg(C.byref(st)) == ret
        assert st.value == state


def test_static_work_distribution():
    for e in KAT["work_distribution"]:
        assert orc.wd_num_samples(e["w"], e["h"], e["ngpu"]) == e["num_samples"]
        for s, x, y in e["pixels"]:
            assert orc.wd_sample_pixel(e["w"], e["h"], e["ngpu"], e["gpu"], s) == (x, y)


def test_work_distribution_is_a_bijection():
    w, h = 768, 768
    for n in (1, 2, 4, 8):
        seen = np.zeros((h, w), np.int32)
        for gpu in range(n):
            ns = orc.wd_num_samples(w, h, n)
            for s in range(0, ns, 97):
                x, y = orc.wd_sample_pixel(w, h, n, gpu, s)
                assert 0 <= x < w and 0 <= y < h
                seen[y, x] += 1
        assert seen.max() == 1


def test_camera_uvw_bit_exact():
    for e in KAT["camera_uvw"]:
        asp = struct.unpack("<f", struct.pack("<I", e["aspect_bits"]))[0]
        U, V, W = orc.camera_uvw(e["eye"], e["lookat"], e["up"], e["fov_y"], asp)
        got = [bits(x) for x in list(U) + list(V) + list(W)]
        assert got == e["uvw_bits"]


def test_deterministic_sincos_accuracy():
    import ctypes as C
    phis = np.linspace(0, 2 * np.pi, 20001).astype(np.float32)
    s, c = C.c_float(), C.c_float()
    err = 0.0
    for p in phis:
        orc.lib().orc_sincos(float(p), C.byref(s), C.byref(c))
        err = max(err, abs(s.value - np.sin(np.float64(p))), abs(c.value - np.cos(np.float64(p))))
    assert err < 3e-7, err
