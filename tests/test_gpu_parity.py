"""GPU parity tests proper (-m gpu): the CUDA library through the C ABI against the C oracle, bit-exact, on the
seeded cases (all edge cases the oracle was pinned on), on BASELINE.json config sizes, and through
size-independent properties at the full sizes.  Nothing here reads /root/reference."""
import numpy as np
import pytest

import cases
from parity import check_chroma, check_image
from yaik_b200 import capi
from yaik_b200.synth import make_image, mip_chain, SEED_BASE

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return capi.load_library()        # raises if libyaik_b200.so is missing: no fallback


@pytest.fixture(scope="module")
def ctx(lib):
    c = capi.Context(1024, 1024, planes=4, slots=2, lib=lib)
    yield c
    c.close()


@pytest.mark.parametrize("name", list(cases.SMALL_CASES))
def test_small_cases_fused(ctx, name):
    planes, stages = cases.SMALL_CASES[name]()
    check_image(ctx, planes, stages, fused=True)


@pytest.mark.parametrize("name", ["patchy128", "patchy_192x136", "alpha_island128", "ramp64_a2"])
def test_small_cases_stage_by_stage(ctx, name):
    planes, stages = cases.SMALL_CASES[name]()
    check_image(ctx, planes, stages, fused=False)


@pytest.mark.parametrize("name", __import__("golden_check").fixtures())
def test_cuda_matches_golden_vectors_of_the_reference(ctx, name):
    """The CUDA path against vectors produced by the reference itself (tests/golden), without going through the oracle."""
    import golden_check
    g, planes, stages = golden_check.load(name)
    ctx.set_image(planes)                             # the signed R1 fixture falls back to the int32 upload by itself
    if "grad" in stages:
        st = (capi.STAGE_ALPHA if ("alpha" in stages and planes.shape[0] == 4) else 0) | capi.STAGE_GRADIENT | (capi.STAGE_RANGE1D if "r2" in stages else 0)
        ctx.analyze(st)
    golden_check.check(g, stages, alpha=ctx.alpha_reject, gradient_pass=ctx.gradient_pass, range1d=ctx.range1d,
                       range_dyn=lambda n, m3: ctx.range_dyn(n, mode3=m3, want_dst=True), state=ctx.download_state,
                       chroma=lambda cfg, modes: ctx.chroma(cfg, modes))


@pytest.mark.parametrize("name,pre,cfg,modes", cases.CHROMA_CASES)
def test_chroma_front_end(ctx, name, pre, cfg, modes):
    """RGB -> YCoCg, SampleDown, DynamicTileEncode of Y / reduced Co / Cg against the oracle."""
    planes, _ = cases.SMALL_CASES[name]()
    check_chroma(ctx, planes, pre, cfg, modes)


@pytest.mark.parametrize("i", range(16))
def test_chroma_front_end_random_configurations(ctx, i):
    """The configurations tests/test_oracle_vs_ref.py pins the oracle on against the compiled reference."""
    planes, pre, cfg, modes = cases.random_chroma_case(i)
    check_chroma(ctx, planes, pre, cfg, modes)


def test_chroma_front_end_2048_rgba(lib):
    """The bench texture through the CLI's chroma configuration (half-width Co and Cg, box average)."""
    c = capi.Context(2048, 2048, planes=4, slots=1, lib=lib)
    try:
        check_chroma(c, make_image(2048, 2048, 4, SEED_BASE + 1), ("alpha", "grad"), (1, 0, 1, 0), (2, 2))
    finally:
        c.close()


def test_config0_512_rgb(ctx):
    """BASELINE.json configs[0]: 512x512 RGB illustration-like image, gradient + range stages."""
    check_image(ctx, make_image(512, 512, 3, SEED_BASE + 0), ("grad", "r2", "r1"))


def test_1024_rgba_full_compare(ctx):
    """One texture of BASELINE.json configs[2] (256 x 1024x1024 RGBA): full comparison with the oracle."""
    check_image(ctx, make_image(1024, 1024, 4, SEED_BASE + 2), ("alpha", "grad", "r2"))


def test_batch_equals_single(lib):
    """configs[2] shape: a batch launch over slots gives exactly what one-image launches give."""
    imgs = [make_image(256, 256, 4, SEED_BASE + 100 + i) for i in range(4)]
    cb = capi.Context(256, 256, planes=4, slots=4, lib=lib)
    cs = capi.Context(256, 256, planes=4, slots=1, lib=lib)
    try:
        for i, im in enumerate(imgs):
            cb.set_image(im, i)
        cb.analyze(capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D, slot0=0, n_slots=4)
        first = [cb.gradient_pass(4, 4, slot=i) for i in range(4)]
        cb.reset_states(0, 4)                                    # a second batch on the same slots after one batched reset
        cb.analyze(capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D, slot0=0, n_slots=4)
        for i in range(4):
            again = cb.gradient_pass(4, 4, slot=i)
            assert np.array_equal(again["bitmap"], first[i]["bitmap"]) and np.array_equal(again["rgb"], first[i]["rgb"])
        cb.reset_states(0, 4)
        cb.analyze(capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D, slot0=0, n_slots=4)
        for i, im in enumerate(imgs):
            cs.set_image(im, 0)
            cs.analyze(capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D)
            for sx, sy in capi.PASS_ORDER:
                a, b = cb.gradient_pass(sx, sy, slot=i), cs.gradient_pass(sx, sy)
                assert a["tiledone"] == b["tiledone"] and a["bbox"] == b["bbox"]
                assert np.array_equal(a["bitmap"], b["bitmap"]) and np.array_equal(a["rgb"], b["rgb"])
            for p in range(3):
                a, b = cb.range1d(p, slot=i), cs.range1d(p)
                assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["type"], b["type"])
    finally:
        cb.close(); cs.close()


def test_config1_2048_rgba_properties(lib):
    """BASELINE.json configs[1] at full size: oracle comparison of every stream plus size-independent properties
    (bitmap popcount == TileDone, claimed cells == union of accepted tiles, R2 length == 16 bytes per unclaimed
    cell, idempotence of a second run)."""
    planes = make_image(2048, 2048, 4, SEED_BASE + 1)
    c = capi.Context(2048, 2048, planes=4, slots=1, lib=lib)
    try:
        check_image(c, planes, ("alpha", "grad", "r2"), state=False)
        c.set_image(planes)
        c.analyze(capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D)
        first = [c.gradient_pass(sx, sy) for sx, sy in capi.PASS_ORDER]
        for g in first:
            assert int(np.unpackbits(g["bitmap"]).sum()) == g["tiledone"]
        st = c.download_state(recon=False)
        claimed = (st["smoothMap"] != 0)
        cells = claimed.reshape(512, 4, 512, 4)
        assert (cells.all(axis=(1, 3)) | ~cells.any(axis=(1, 3))).all()          # claims are whole 4x4 cells
        r = c.range1d(0)
        assert r["idx"].size == int((~claimed).sum())
        assert r["idx"].max() <= 16
        c.set_image(planes)
        c.analyze(capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D)
        for g, (sx, sy) in zip(first, capi.PASS_ORDER):
            h = c.gradient_pass(sx, sy)
            assert np.array_equal(g["bitmap"], h["bitmap"]) and np.array_equal(g["rgb"], h["rgb"])
    finally:
        c.close()


def _load_device_planes(c, planes, slot):
    """int32 planes the caller put in HBM, handed over with yk_set_image_device (what bench.py's `value` runs on)."""
    ch, h, w = planes.shape
    ptrs = []
    for i in range(ch):
        ptrs.append(c.device_plane(slot, i))
        c.copy_from_host(ptrs[-1], np.ascontiguousarray(planes[i], dtype=np.int32))
    c.set_image_device(ptrs, ch, w, h, slot)


@pytest.mark.parametrize("how", ["int32_upload", "device_planes"])
def test_config1_2048_rgba_int32_kernel_full_compare(lib, how):
    """BASELINE.json configs[1] through the int32 variant of the analysis kernel (yk_k_analyze, the one bench.py times):
    yk_set_upload_format(0) + yk_set_image, and yk_set_image_device on resident planes; every stream against the oracle."""
    planes = make_image(2048, 2048, 4, SEED_BASE + 1)
    c = capi.Context(2048, 2048, planes=4, slots=1, lib=lib)
    try:
        c.set_upload_format(False)
        check_image(c, planes, ("alpha", "grad", "r2"), state=False, load=_load_device_planes if how == "device_planes" else None)
    finally:
        c.close()


def test_bench_configuration_matches_oracle(lib):
    """The configuration bench.py times: 8 contexts on 8 CUDA streams, analysis launches of a quarter of the SMs, int32 planes
    resident in HBM, 8 distinct 2048x2048 RGBA textures issued round-robin without synchronising in between.  Every
    context's result must equal the digest of a single-context, full-SM run that was itself compared with the oracle."""
    import torch
    from parity import results_digest
    st = capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D
    imgs = [make_image(2048, 2048, 4, SEED_BASE + 1 + i) for i in range(8)]
    c0 = capi.Context(2048, 2048, planes=4, slots=1, lib=lib)
    want = []
    try:
        c0.set_upload_format(False)
        for k, im in enumerate(imgs):
            if k < 2:
                check_image(c0, im, ("alpha", "grad", "r2"), state=False)      # oracle-verified
            else:
                c0.set_image(im); c0.analyze(st)
            want.append(results_digest(c0.fetch_all(0)))
    finally:
        c0.close()
    ctxs = [capi.Context(2048, 2048, planes=4, slots=1, lib=lib) for _ in range(8)]
    streams = [torch.cuda.Stream() for _ in range(8)]
    try:
        for c, s, im in zip(ctxs, streams, imgs):
            c.set_stream(s.cuda_stream); c.set_upload_format(False); c.set_analysis_ctas(c.sm_count() // 4)
            _load_device_planes(c, im, 0)
        for rep in range(6):
            for c in ctxs:
                c.reset_state(0); c.analyze(st)
        got = [results_digest(c.fetch_all(0)) for c in ctxs]
        assert got == want
    finally:
        for c in ctxs:
            c.close()


def test_config4_mip_chain_4096(lib):
    """BASELINE.json configs[4] at the size it is quoted on: RGBA mip chain 4096 -> 4, every level against the oracle."""
    chain = mip_chain(4096, SEED_BASE + 4)
    c = capi.Context(4096, 4096, planes=4, slots=1, lib=lib)
    try:
        for lvl in chain:
            s = lvl.shape[1]
            stages = ["grad"]
            if s >= 16 and lvl[3].any():
                stages.append("alpha")
            if s >= 8:
                stages.append("r2")
            check_image(c, lvl, tuple(stages), state=(s <= 512))
    finally:
        c.close()


def test_config4_mip_chain(lib):
    """BASELINE.json configs[4]: RGBA mip chain down to 4x4 (small-tile and alpha-rejection paths); levels run the
    stages the reference itself supports at that size (SURVEY.md hazards 7, 11)."""
    chain = mip_chain(512, SEED_BASE + 4)
    c = capi.Context(512, 512, planes=4, slots=1, lib=lib)
    try:
        for lvl in chain:
            s = lvl.shape[1]
            stages = ["grad"]
            if s >= 16 and lvl[3].any():
                stages.append("alpha")
            if s >= 8:
                stages.append("r2")
            if s >= 32:
                stages.append("r1")
            check_image(c, lvl, tuple(stages))
    finally:
        c.close()


def test_out_of_range_sample_is_reported(ctx):
    planes = make_image(64, 64, 3, SEED_BASE + 7).copy()
    planes[1, 10, 10] = 300
    ctx.set_image(planes)
    ctx.analyze(capi.STAGE_GRADIENT)
    with pytest.raises(capi.YaikError) as e:
        ctx.gradient_pass(4, 4)
    assert e.value.code == -4


def test_alternative_data_paths(lib):
    """Upload formats (packed bytes / int32), device-resident int32 planes (what bench.py's `value` runs on), yk_fetch_all."""
    import paths_check
    c = capi.Context(256, 256, planes=4, slots=1, lib=lib)
    try:
        paths_check.check_upload_formats(c)
        paths_check.check_out_of_range(c)
        paths_check.check_fetch_all(c)

        def to_device(planes):
            ptrs = []
            for i in range(planes.shape[0]):
                ptrs.append(c.device_plane(0, i))
                c.copy_from_host(ptrs[-1], np.ascontiguousarray(planes[i], dtype=np.int32))
            return ptrs
        paths_check.check_device_resident_planes(c, to_device)
    finally:
        c.close()


@pytest.mark.parametrize("i", range(64))
def test_random_shapes_match_oracle(ctx, i):
    """Randomised sizes (any multiple of 4 up to 192 x 160) and contents, both upload formats, fused and stage by stage."""
    from test_random_shapes_emulated import _case
    planes, stages = _case(i)
    ctx.set_upload_format(i % 2 == 0)
    try:
        check_image(ctx, planes, stages, fused=(i % 3 != 0))
    finally:
        ctx.set_upload_format(True)


@pytest.mark.parametrize("w,h,ch,seed", [(2048, 2048, 4, SEED_BASE + 1), (4096, 4096, 3, SEED_BASE + 3)])
def test_streams_decode_with_the_reference_decoders_corner_rule(lib, w, h, ch, seed):
    """Size-independent property at BASELINE sizes (configs[1]; a 4096x4096 crop-equivalent of configs[3]): replaying the
    reference decoder's corner consumption (decoder/YAIK_Gradient.cpp) over the emitted bitmaps reads every rgbStream to
    its last byte, gives every touched lattice point the source pixel's colour, and touches exactly the points the encoder
    marked in mappedRGB."""
    import decoder_walk
    planes = make_image(w, h, ch, seed)
    c = capi.Context(w, h, planes=ch, slots=1, lib=lib)
    try:
        c.set_image(planes)
        c.analyze((capi.STAGE_ALPHA if ch == 4 else 0) | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D)
        passes = [c.gradient_pass(sx, sy) for sx, sy in capi.PASS_ORDER]
        st = c.download_state(recon=False)
        n = decoder_walk.check(planes, passes, st["mappedRGB"][0])
        assert n > 1000
        # the range stage's streams through Decompress1D's consumption rule (decoder/YAIK_3DTile.cpp:24-240)
        claimed = st["smoothMap"][::4, ::4] != 0
        for pl in range(3):
            r = c.range1d(pl)
            assert decoder_walk.check_range1d(planes[pl], claimed, r["idx"], r["type"]) > 100
    finally:
        c.close()


def test_repeatability_under_concurrency(lib):
    """The persistent kernel's queue / barrier protocol under load: four contexts on four host threads analyse the same
    1024x1024 textures over and over, with full and half-size analysis launches; every run must give the same streams."""
    import threading
    from parity import results_digest
    imgs = [make_image(1024, 1024, 4, SEED_BASE + 40 + i) for i in range(2)]
    st = capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D

    def digest(c):
        return results_digest(c.fetch_all(0, copy=True))

    ref = []
    c0 = capi.Context(1024, 1024, planes=4, slots=1, lib=lib)
    for im in imgs:
        c0.set_image(im); c0.analyze(st); ref.append(digest(c0))
    c0.close()
    errors = []

    def worker(t):
        c = capi.Context(1024, 1024, planes=4, slots=1, lib=lib)
        try:
            c.set_analysis_ctas(0 if t % 2 == 0 else c.sm_count() // 2)
            c.set_upload_format(t < 2)
            for it in range(40):
                k = (it + t) % 2
                c.set_image(imgs[k]); c.analyze(st)
                if digest(c) != ref[k]:
                    errors.append((t, it))
        finally:
            c.close()
    ths = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errors, errors


# ---- tile-row strips of one large image (BASELINE.json configs[3]) --------------------------------
@pytest.mark.parametrize("w,h,n", [(256, 512, 2), (256, 512, 4), (192, 328, 3), (2048, 1024, 4)])
def test_strips_on_one_gpu_match_whole_image(lib, w, h, n):
    """The strip protocol (pixel-row halo, boundary touch words, merge) with every strip in its own context on this GPU,
    halo copies device to device: the merged streams must be what the oracle gives for the whole image."""
    from strips_check import check_against_oracle
    from yaik_b200 import strips
    planes = make_image(w, h, 3, SEED_BASE + 3) if w > 256 else cases._patchy(w, h, 43, 4, 3)
    ctxs = [capi.Context(w, h, planes=3, slots=1, lib=lib) for _ in range(n)]
    try:
        merged = strips.LocalTransport(ctxs).run(planes, n_strips=n)
        check_against_oracle(merged, planes)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("w,h,n,ch", [(256, 512, 2, 3), (512, 512, 4, 4), (2048, 1024, 4, 3), (192, 328, 3, 3)])
def test_strips_without_the_host_in_the_loop(lib, w, h, n, ch):
    """yk_strips_link / yk_strips_run: every strip enqueues whole images on its stream, the halo exchanges are ordered by
    epoch flags in (peer-mapped) halo memory; three images back to back, the last one's streams against the oracle."""
    from strips_check import check_against_oracle
    from yaik_b200 import strips
    planes = make_image(w, h, ch, SEED_BASE + 3) if w >= 512 else cases._patchy(w, h, 43, 4, 3)
    ctxs = [capi.Context(w, h, planes=ch, slots=1, lib=lib) for _ in range(n)]
    try:
        merged = strips.LocalTransport(ctxs).run_device_flags(planes, n_strips=n, images=3)
        check_against_oracle(merged, planes)
    finally:
        for c in ctxs:
            c.close()


def test_strips_4096x16384_match_oracle(lib):
    """configs[3] at a quarter of its width and its full height (67 Mpixel RGB, 4 strips of 4096 rows): the merged streams
    against the oracle on the whole image."""
    from strips_check import check_against_oracle
    from yaik_b200 import strips
    base = make_image(2048, 2048, 3, SEED_BASE + 3)
    planes = np.ascontiguousarray(np.tile(base, (1, 8, 2)))
    assert planes.shape == (3, 16384, 4096)
    ctxs = [capi.Context(4096, 4096, planes=3, slots=1, lib=lib) for _ in range(4)]
    try:
        merged = strips.LocalTransport(ctxs).run(planes, n_strips=4)
        check_against_oracle(merged, planes)
    finally:
        for c in ctxs:
            c.close()


def test_strips_two_ranks_ipc_peer_copies():
    """Two processes (two GPUs when the box has them, else both on GPU 0), halo exchange by CUDA IPC + peer copies."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29600 + os.getpid() % 300))
    procs = [subprocess.Popen([sys.executable, os.path.join(root, "tests", "strips_rank.py"), "--rank", str(r), "--world", "2",
                               "--backend", "gloo", "--mode", "ipc", "--w", "512", "--h", "1024"],
                              env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "strips ok" in outs[0]
