"""The C-ABI library loads and exports every entry point include/gpis_b200.h declares (no compute
calls: this runs without a GPU), and the host library exports the class wrappers."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from gpismap_b200 import build
    build.build_all()
    return build


def declared():
    src = open(os.path.join(ROOT, "include", "gpis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpis_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    L = ctypes.CDLL(built.LIB_CUDA)
    names = declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    from gpismap_b200 import cabi
    assert set(cabi.EXPORTS) <= set(names)


def test_config_defaults_match_reference_params(built):
    """params.h:27-110 and the constants hard-coded in the orchestrators."""
    from gpismap_b200 import cabi
    c3 = cabi.default_config(3)
    assert (c3.map_scale, c3.map_noise, c3.var_thre) == (ctypes.c_float(0.04).value, ctypes.c_float(5e-3).value, 0.5)
    assert c3.cluster_half == ctypes.c_float(0.025).value
    assert c3.search_half == ctypes.c_float(ctypes.c_float(0.025).value * 3.0).value   # AABB3(x, C_leng*3.0) takes a float
    c2 = cabi.default_config(2)
    assert c2.cluster_half == ctypes.c_float(0.8).value and abs(c2.search_half - 4.8) < 1e-6
    assert c2.var_thre == ctypes.c_float(0.4).value
    assert cabi.lib().gpis_config_default(None, 3) < 0 and cabi.lib().gpis_config_default(ctypes.byref(c2), 4) < 0


def test_no_cpu_fallback(built):
    """Without a CUDA device the context cannot be created; nothing silently computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpismap_b200 import cabi
    with pytest.raises(RuntimeError):
        cabi.Ctx(3)


def test_host_library_exports(built):
    L = ctypes.CDLL(built.LIB_HOST)
    for n in ["gm3_create", "gm3_update", "gm3_test", "gm3_reset", "gm3_get_all_points", "gm3_set_cam",
              "gm2_create", "gm2_update", "gm2_test", "gm2_reset"]:
        assert hasattr(L, n), n
