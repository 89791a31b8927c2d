"""Source-level drop-in check of the mex command protocol: the reference's own, unmodified MATLAB
gateways (mex/mexGPisMap3.cpp, mex/mexGPisMap.cpp) must compile against this repo's class headers
(include/gpismap/GPisMap3.h, GPisMap.h) with a stub mex.h, and link against the drop-in host library with only the
mx* symbols unresolved — every GPisMap / GPisMap3 member the gateways call exists with the reference's signature.
Needs /root/reference (this container); skipped where it is absent."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MEX = "/root/reference/mex"


@pytest.mark.parametrize("gateway", ["mexGPisMap3.cpp", "mexGPisMap.cpp"])
def test_reference_gateway_compiles_and_links_against_the_dropin(gateway, tmp_path):
    src = os.path.join(REF_MEX, gateway)
    if not os.path.exists(src):
        pytest.skip("reference sources not present")
    from gpismap_b200 import build
    host = build.build_host()
    obj = str(tmp_path / (gateway + ".o"))
    subprocess.check_call(["g++", "-std=c++17", "-fPIC", "-c", "-w", "-I" + os.path.join(ROOT, "tests", "mex_stub"),
                           "-I" + os.path.join(ROOT, "include", "gpismap"), "-I" + os.path.join(ROOT, "include"), src, "-o", obj])
    # link as a shared object against the drop-in; the only undefined symbols left must be MATLAB's
    so = str(tmp_path / (gateway + ".so"))
    subprocess.check_call(["g++", "-shared", "-o", so, obj, host, "-Wl,-rpath," + os.path.dirname(host)])
    und = subprocess.check_output(["nm", "-D", "--undefined-only", so], text=True)
    missing = [ln.split()[-1] for ln in und.splitlines() if "GPisMap" in ln]
    resolved = subprocess.check_output(["nm", "-D", "--defined-only", host], text=True)
    for sym in missing:
        assert sym in resolved, f"the gateway needs {sym}, which the drop-in library does not export"
    assert any(s.startswith("mx") for s in (ln.split()[-1] for ln in und.splitlines())), "stub symbols expected to stay unresolved"
