"""Host-side steps either side of the path: sutil::saveImage's PPM / PNG conventions and imgui_test's
load_assimp layout (SDK/sutil/sutil.cpp:542-709, SDK/imgui_test/triangle_gas.cpp:78-168).  No GPU."""
import numpy as np
import pytest

from optix_raytracer_b200 import host


def test_save_image_ppm_flips_rows_and_drops_alpha(tmp_path):
    frame = np.zeros((3, 2, 4), np.uint8)
    frame[0, 0] = (10, 20, 30, 255)   # launch index (0, 0): bottom-left of the picture
    frame[2, 1] = (200, 210, 220, 7)
    p = tmp_path / "out.ppm"
    host.save_image(p, frame)
    raw = p.read_bytes()
    assert raw.startswith(b"P6\n2 3\n255\n")
    pix = np.frombuffer(raw[len(b"P6\n2 3\n255\n"):], np.uint8).reshape(3, 2, 3)
    assert tuple(pix[2, 0]) == (10, 20, 30) and tuple(pix[0, 1]) == (200, 210, 220)


def test_save_image_float_buffers_are_srgb_converted_like_the_reference(tmp_path):
    acc = np.zeros((1, 4, 4), np.float32)
    acc[0, :, 0] = [0.0, 0.001, 0.5, 2.0]
    p = tmp_path / "acc.ppm"
    host.save_image(p, acc)
    pix = np.frombuffer(p.read_bytes().split(b"255\n", 1)[1], np.uint8).reshape(1, 4, 3)
    # int(256 * toSRGB(f)) clamped: 0 -> 0, 0.001 -> int(256 * 12.92e-3) = 3, 0.5 -> int(256 * 0.7354) = 188, 2.0 -> 255
    assert list(pix[0, :, 0]) == [0, 3, 188, 255]
    host.save_image(tmp_path / "lin.ppm", acc, disable_srgb_conversion=True)
    lin = np.frombuffer((tmp_path / "lin.ppm").read_bytes().split(b"255\n", 1)[1], np.uint8).reshape(1, 4, 3)
    assert list(lin[0, :, 0]) == [0, 0, 128, 255]


def test_save_image_png_round_trip_and_errors(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    frame = (np.arange(2 * 3 * 4) % 256).astype(np.uint8).reshape(2, 3, 4)
    host.save_image(tmp_path / "f.png", frame)
    back = np.array(Image.open(tmp_path / "f.png"))
    assert np.array_equal(back, frame[::-1])
    with pytest.raises(ValueError, match="uchar4 images to EXR"):
        host.save_image(tmp_path / "f.exr", frame)
    with pytest.raises(ValueError):
        host.save_image(tmp_path / "g.ppm", np.zeros((2, 2), np.uint8))


def test_obj_loader_has_the_layout_and_the_floor_of_load_assimp(tmp_path):
    obj = tmp_path / "quad.obj"
    obj.write_text("v 0 0.5 0\nv 1 0.5 0\nv 1 1.5 0\nv 0 1.5 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\n")
    v, n, m = host.load_obj_like_assimp(obj)
    # the quad is fanned into two triangles, unindexed; then 20 x 20 x 2 floor triangles at the model's lowest y with material 26
    assert v.shape == (6 + 2400, 3) and n.shape == (6 + 2400, 3) and m.shape == (2 + 800,)
    assert np.array_equal(v[:6], np.array([[0, .5, 0], [1, .5, 0], [1, 1.5, 0], [0, .5, 0], [1, 1.5, 0], [0, 1.5, 0]], np.float32))
    assert (m[:2] == 0).all() and (m[2:] == 26).all()
    assert np.allclose(v[6:, 1], 0.5) and np.array_equal(n[6:], np.tile(np.array([[0, 1, 0]], np.float32), (2400, 1)))
    assert np.allclose(v[6], [-1.0, 0.5, -1.0]) and np.allclose(v[-1], [1.0, 0.5, 1.0])


@pytest.mark.parametrize("shape", [(5, 7, 3), (40, 33, 4), (16, 16, 4)])
def test_save_image_exr_round_trip(tmp_path, shape):
    """EXR branch of sutil::saveImage (SDK/sutil/sutil.cpp:660-702): fp16 channels in alphabetical order, ZIP blocks of 16 lines from
    16 x 16 on, rows in buffer order, values linear."""
    rng = np.random.default_rng(1)
    acc = (rng.random(shape, dtype=np.float32) * 4).astype(np.float32)
    acc[0, 0, :3] = (0.0, 1.0, 65504.0)
    p = tmp_path / "a.exr"
    host.save_image(p, acc)
    raw = p.read_bytes()
    assert raw[:8] == (20000630).to_bytes(4, "little") + (2).to_bytes(4, "little")
    names = [b"A", b"B", b"G", b"R"] if shape[2] == 4 else [b"B", b"G", b"R"]
    pos = [raw.index(b"channels\0chlist\0") + 20]
    assert raw[pos[0]:pos[0] + 1] == names[0] and b"compression\0compression\0\x01\0\0\0" + bytes([3 if min(shape[:2]) >= 16 else 0]) in raw
    back = host.load_exr(p)
    assert sorted(back) == sorted(n.decode() for n in names)
    for k, c in enumerate("RGBA"[:shape[2]]):
        assert np.array_equal(back[c], acc[..., k].astype(np.float16).astype(np.float32)), c


def test_nbt_model_round_trip_and_layout(tmp_path):
    """imgui_test's load_nbt (SDK/imgui_test/triangle_gas.cpp:16-76): meshes in file order, little-endian float triplets inside big-endian NBT
    byte arrays, one material index per vertex; gzip-wrapped or raw."""
    rng = np.random.default_rng(2)
    a_v, a_n = rng.random((6, 3), dtype=np.float32), rng.random((6, 3), dtype=np.float32)
    b_v, b_n = rng.random((3, 3), dtype=np.float32) - 2, rng.random((3, 3), dtype=np.float32)
    for compress in (True, False):
        p = tmp_path / f"m{int(compress)}.nbt"
        host.save_nbt(p, {"blob": (a_v, a_n), "floor": (b_v, b_n)}, compress=compress)
        assert (p.read_bytes()[:2] == b"\x1f\x8b") == compress
        v, n, m = host.load_nbt(p)
        assert np.array_equal(v, np.concatenate([a_v, b_v])) and np.array_equal(n, np.concatenate([a_n, b_n]))
        assert m.shape == (9,) and m.dtype == np.int32 and not m.any()
    raw = (tmp_path / "m0.nbt").read_bytes()
    # root compound, empty name; first child: compound "blob"; its first entry: byte array "vertices" of 72 bytes, big-endian length
    assert raw[:3] == b"\x0a\x00\x00" and raw[3:10] == b"\x0a\x00\x04blob" and raw[10:25] == b"\x07\x00\x08vertices\x00\x00\x00\x48"
    assert np.array_equal(np.frombuffer(raw[25:37], "<f4"), a_v[0])
    with pytest.raises(RuntimeError, match="can't find file"):
        host.load_nbt(tmp_path / "missing.nbt")
    bad = tmp_path / "bad.nbt"
    bad.write_bytes(b"\x01\x00\x00\x05")
    with pytest.raises(ValueError):
        host.load_nbt(bad)


def test_exr_is_read_by_an_independent_decoder(tmp_path, monkeypatch):
    """The file save_exr writes, decoded by OpenCV's OpenEXR reader (not by this repo's load_exr)."""
    import os
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    for shape in [(40, 33, 4), (5, 7, 3)]:
        acc = (rng.random(shape, dtype=np.float32) * 4).astype(np.float32)
        p = str(tmp_path / "x.exr")
        host.save_image(p, acc)
        try:
            img = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        except cv2.error as e:  # codec compiled out or disabled in this process
            pytest.skip(f"OpenCV cannot decode EXR here: {e}")
        if img is None:
            pytest.skip("this OpenCV build has no OpenEXR codec")
        got = img[..., [2, 1, 0, 3]] if shape[2] == 4 else img[..., ::-1]
        assert np.array_equal(got, acc.astype(np.float16).astype(np.float32))


def test_load_gltf_reads_the_duck_as_the_committed_fixture_has_it():
    """host.load_gltf on the reference's Duck.gltf (baseline/_ref/SDK/data, where baseline/Makefile put the assets) against the
    fixture tests/golden/duck_mesh.npz that tools/make_duck_fixture.py made from the same file; the comparison with sutil::loadScene
    itself runs on the GPU box (tests/test_gpu_reference_samples.py)."""
    import pathlib
    import pytest
    from optix_raytracer_b200 import host
    from tests import common
    path = pathlib.Path(__file__).resolve().parents[1] / "baseline" / "_ref" / "SDK" / "data" / "Duck" / "Duck.gltf"
    if not path.exists():
        pytest.skip("baseline/_ref/SDK/data/Duck not present (make -C baseline needs /root/reference)")
    sc, fx = host.load_gltf(path), common.duck_scene()
    p, q = sc["meshes"][0]["primitives"][0], fx["meshes"][0]["primitives"][0]
    for k in ("positions", "normals"):
        assert np.array_equal(p[k].view(np.uint32), q[k].view(np.uint32))
    assert np.array_equal(p["indices"], q["indices"]) and np.array_equal(p["texcoords"][0], q["texcoords"][0])
    assert np.allclose(sc["instances"][0]["transform"], fx["instances"][0]["transform"], rtol=0, atol=1e-7)
    assert p["views"]["positions"][3] == 12 and p["views"]["indices"][4] == 2 and len(sc["buffers"]) == 1
    assert sc["materials"][0]["base_color_tex"] is not None and p["colors"] is None
