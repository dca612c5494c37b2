"""The emulated kernels and the C-ABI layer under AddressSanitizer (compute-sanitizer is not available on the GPU pool):
skipped where no g++ with libasan is installed."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASAN_CXX = os.environ.get("ASAN_CXX", "/usr/bin/g++")


def _libasan():
    try:
        p = subprocess.run([ASAN_CXX, "-print-file-name=libasan.so"], capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception:
        return None
    return p if os.path.isabs(p) and os.path.exists(p) else None


def test_emulated_kernels_run_clean_under_asan():
    libasan = _libasan()
    if not libasan:
        pytest.skip("no g++ with libasan here")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu"), "asan", "ASAN_CXX=" + ASAN_CXX], check=True)
    lib = os.path.join(ROOT, "tests", "emu", "_build", "libyaik_b200_emu_asan.so")
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests")]))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "asan_cases.py"), lib], capture_output=True, text=True, timeout=900, env=env, cwd=os.path.join(ROOT, "tests"))
    assert r.returncode == 0 and "asan cases ok" in r.stdout, (r.stdout[-1000:], r.stderr[-3000:])
