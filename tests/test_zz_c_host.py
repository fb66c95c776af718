"""examples/q6_host.c: a host in plain C over the C ABI (no Python, no torch in the process) -- the shape of a compiled
host such as the reference's Haskell binary.  Without a GPU it must build, link against libvdl_cuda.so and FAIL LOUDLY (there
is no CPU fallback); on a GPU it runs the reference's Q6 program and checks the answer against its own scalar loop."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def q6_host(tmp_path_factory):
    sys.path.insert(0, ROOT)
    from mplan2vdl_b200.build import build_library
    build_library()
    exe = str(tmp_path_factory.mktemp("c_host") / "q6_host")
    libdir = os.path.join(ROOT, "mplan2vdl_b200")
    subprocess.check_call(["cc", "-std=c11", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "q6_host.c"),
                           "-o", exe, "-L" + libdir, "-lvdl_cuda", "-Wl,-rpath," + libdir])
    return exe


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(has_gpu(), reason="this is the no-GPU behaviour")
def test_c_host_builds_links_and_fails_loudly_without_a_gpu(q6_host):
    r = subprocess.run([q6_host, os.path.join(ROOT, "plans", "q06.vdl"), "1000"], capture_output=True, text=True)
    assert r.returncode == 2 and "vdl_ctx_create" in r.stderr and r.stdout == ""


def test_c_host_usage_and_missing_plan(q6_host):
    assert subprocess.run([q6_host], capture_output=True).returncode == 64
    assert subprocess.run([q6_host, "/nonexistent.vdl"], capture_output=True).returncode == 66


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [1, 4097, 3_000_000])
def test_c_host_runs_q6_and_agrees_with_its_own_scalar_loop(q6_host, rows):
    import json
    r = subprocess.run([q6_host, os.path.join(ROOT, "plans", "q06.vdl"), str(rows)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout)
    assert line["revenue"] == line["expected"] and line["fused_scans"] == 1 and line["rows"] == rows
