"""TEST SCAFFOLDING: builds tests/_mock/{libgpis_mock.so, libgpismap_host_mock.so} — the host classes
linked against a CPU stand-in of the C ABI that is backed by the oracle — so host logic can be
checked without a GPU. Never used by the product path."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_mock")


def build():
    os.makedirs(OUT, exist_ok=True)
    mock = os.path.join(OUT, "libgpis_mock.so")   # its own soname: must never alias the real library
    host = os.path.join(OUT, "libgpismap_host_mock.so")
    srcs_mock = [os.path.join(HERE, "mock_cabi", "mock_gpis.cpp"), os.path.join(ROOT, "oracle", "gpis_oracle.c")]
    hostdir = os.path.join(ROOT, "gpismap_b200", "host")
    srcs_host = [os.path.join(hostdir, f) for f in sorted(os.listdir(hostdir)) if f.endswith(".cpp")]
    deps = srcs_mock + srcs_host + [os.path.join(hostdir, f) for f in os.listdir(hostdir)] + \
        [os.path.join(ROOT, "include", "gpis_b200.h")]
    newest = max(os.path.getmtime(p) for p in deps)
    if os.path.exists(mock) and os.path.exists(host) and min(os.path.getmtime(mock), os.path.getmtime(host)) > newest:
        return host
    obj = os.path.join(OUT, "gpis_oracle.o")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-ffp-contract=off", "-DREAL=float", "-c",
                           os.path.join(ROOT, "oracle", "gpis_oracle.c"), "-o", obj])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                           "-o", mock, srcs_mock[0], obj, "-lm"])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
                           "-I" + os.path.join(ROOT, "include"), "-I" + hostdir, "-o", host] + srcs_host +
                          ["-L" + OUT, "-lgpis_mock", "-Wl,-rpath,$ORIGIN"])
    return host


if __name__ == "__main__":
    print(build())
