"""The C-ABI library loads without a GPU and exports every symbol include/euclider_b200.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "euclider_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eucl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    names = declared_symbols()
    assert len(names) >= 19
    for name in names:
        assert hasattr(built_lib, name), f"{name} is declared in the header but not exported"


def test_python_prototypes_cover_the_header(built_lib):
    from euclider_b200 import _capi

    assert sorted(_capi.PROTOTYPES) == declared_symbols()


def test_struct_sizes_match_the_header(built_lib):
    """ctypes mirrors must have the C layout (a mismatch would corrupt every call)."""
    from euclider_b200 import _capi

    assert C.sizeof(_capi.EuclPrim) == 8 + 8 * 4 * 2 + 16
    assert C.sizeof(_capi.EuclNode) == 16 and C.sizeof(_capi.EuclEntity) == 16
    assert C.sizeof(_capi.EuclColorOp) == 8 + 12 * 8
    assert C.sizeof(_capi.EuclCamera) == 16 + 4 * 32
    assert C.sizeof(_capi.EuclRenderOpts) == 48
    assert C.sizeof(_capi.EuclStats) == 24 + 64 * 8 + 16 + 20 + 4  # five floats + graph_replays


def test_no_gpu_means_loud_failure(built_lib):
    """Without a CUDA device the render path must fail, not fall back to a CPU renderer."""
    import euclider_b200 as eb

    if built_lib.eucl_device_count() > 0:
        pytest.skip("a GPU is present")
    env = eb.Parser.default().parse(MINIMAL_SCENE, load_textures=False)
    env.set_texture(0, 1, 1, bytes([1, 2, 3, 255]))
    with pytest.raises(eb.EuclError) as err:
        env.render((8, 8))
    assert err.value.status == -32  # EUCL_ERR_NO_DEVICE


def test_version_and_error_strings(built_lib):
    assert b"sm_100a" in built_lib.eucl_version()
    assert built_lib.eucl_scene_parse(b"{", C.byref(C.c_void_p())) == -11
    assert b"Invalid JSON" in built_lib.eucl_last_error()


MINIMAL_SCENE = """
{"Universe3": {"camera": {"PitchYawCamera3": []},
  "entities": [{"Void3::new_with_vacuum": []}],
  "background": {"MappedTextureImpl3::new": [{"uv_sphere_3": [{"Point3::new": [0, 0, 0]}]},
                                              {"texture_image_linear": ["./none.png"]}]}}}
"""


def test_write_ppm_flips_rows(built_lib, tmp_path):
    import numpy as np

    rgb = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(3, 2, 3)  # 3 rows (bottom-up) of 2 pixels
    path = tmp_path / "t.ppm"
    assert built_lib.eucl_write_ppm(str(path).encode(), 2, 3, rgb.ctypes.data) == 0
    data = path.read_bytes()
    assert data.startswith(b"P6\n2 3\n255\n")
    assert data[len(b"P6\n2 3\n255\n"):] == rgb[::-1].tobytes()
    assert built_lib.eucl_write_ppm(b"/nonexistent-dir/x.ppm", 2, 3, rgb.ctypes.data) == -1
