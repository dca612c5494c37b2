"""Host logic of the multi-GPU strip path (yaik_b200/strips.py) on the CPU: the partition, the two halo exchanges and the
merge, driven through the CPU-emulated build of the kernels (tests/emu — test infrastructure) and checked against the
oracle run on the whole image.  The world_size-2 case runs one strip per gloo rank with the host-staged transport."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from strips_check import check_against_oracle
from yaik_b200 import capi, strips

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "_build", "libyaik_b200_emu.so")


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    return capi.load_library(EMU)


def test_strip_rows_partition():
    for h in (64, 72, 192, 1024, 16384, 4100):
        for n in (1, 2, 3, 4, 8):
            rows = strips.strip_rows(h, n)
            assert rows[0][0] == 0 and sum(r[1] for r in rows) == h
            assert all(y0 % 64 == 0 for y0, _ in rows)
            assert all(sh % 64 == 0 for _, sh in rows[:-1]) and all(sh > 0 for _, sh in rows)
            assert all(rows[i][0] + rows[i][1] == rows[i + 1][0] for i in range(len(rows) - 1))
            assert len(rows) == min(n, (h + 63) // 64)


@pytest.mark.parametrize("name,w,h,n", [("patchy", 128, 192, 2), ("patchy", 128, 192, 3), ("ramp", 64, 136, 2), ("synth", 128, 256, 4)])
def test_strips_local_transport_matches_whole_image(emu_lib, name, w, h, n):
    if name == "patchy":
        planes = cases._patchy(w, h, 41, 4, 3)
    elif name == "ramp":
        planes = cases._smooth_noisy(w, h, 42, 2)
    else:
        from yaik_b200.synth import make_image, SEED_BASE
        planes = make_image(w, h, 3, SEED_BASE + 3)
    ctxs = [capi.Context(w, h, planes=3, slots=1, lib=emu_lib) for _ in range(n)]
    try:
        merged = strips.LocalTransport(ctxs).run(planes, n_strips=n)
        check_against_oracle(merged, planes)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("w,h,n,seed", [(128, 128, 2, 16), (256, 256, 4, 12), (128, 256, 3, 19)])
def test_strips_rgba_alpha_stage_of_the_strip_set(emu_lib, w, h, n, seed):
    """RGBA image in strips: the alpha-zero tile rejection of the strips merges into the image's bitmap / bound / remaining
    pixels (pow2 squares are the reference's domain; the rectangle checks the gradient / range streams only)."""
    from yaik_b200.synth import make_image, SEED_BASE
    planes = make_image(w, h, 4, SEED_BASE + seed)
    ctxs = [capi.Context(w, h, planes=4, slots=1, lib=emu_lib) for _ in range(n)]
    try:
        merged = strips.LocalTransport(ctxs).run(planes, n_strips=n)
        if w == h:
            check_against_oracle(merged, planes)
        else:
            check_against_oracle(dict(merged, alpha=None), planes[:3])
            assert merged["alpha"] is not None
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("i", range(6))
def test_strips_random_shapes(emu_lib, i):
    """Random widths (any multiple of 8, odd region counts included), heights and strip counts, both upload formats."""
    from yaik_b200.synth import _rand
    r = _rand(4242, i, np.arange(6, dtype=np.uint64))
    w = 8 * (2 + int(r[0] % np.uint64(22)))              # 16 .. 184
    h = 8 * (9 + int(r[1] % np.uint64(24)))              # 72 .. 256 (at least two 64-row blocks)
    n = 2 + int(r[2] % np.uint64(3))
    planes = cases._patchy(w, h, 500 + i, 4, 1 + int(r[3] % np.uint64(4)))
    ctxs = [capi.Context(w, h, planes=3, slots=1, lib=emu_lib) for _ in range(n)]
    try:
        for c in ctxs:
            c.set_upload_format(i % 2 == 0)
        merged = strips.LocalTransport(ctxs).run(planes, n_strips=n)
        check_against_oracle(merged, planes)
    finally:
        for c in ctxs:
            c.close()


def test_strips_two_gloo_ranks_host_transport(emu_lib, tmp_path):
    """world_size 2, one strip per rank, exchanges staged through host memory and sent point to point over gloo."""
    script = os.path.join(ROOT, "tests", "strips_rank.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + os.getpid() % 500), YK_STRIPS_LIB=EMU)
    procs = [subprocess.Popen([sys.executable, script, "--rank", str(r), "--world", "2", "--backend", "gloo", "--mode", "host",
                               "--w", "128", "--h", "192"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "strips ok" in outs[0]
