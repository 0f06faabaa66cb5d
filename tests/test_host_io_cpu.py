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
    with pytest.raises(ValueError):
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
