"""Invariants of the host-side containers behind the serial passes of update() (tests/host_unit/structs_check.cpp):
the list-based active set against std::set, the per-cell count array and the cluster-level index of PRTree under
random insert / remove sequences in 2-D and 3-D. CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_active_set_counts_and_cluster_index(tmp_path):
    exe = str(tmp_path / "structs_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "gpismap_b200", "host"),
                           os.path.join(ROOT, "tests", "host_unit", "structs_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-2000:] + out.stderr[-2000:]
