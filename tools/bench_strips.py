#!/usr/bin/env python
"""BASELINE.json configs[3]: one 16384x16384 synthetic RGB image split into tile-row strips over the ranks of a torchrun
launch (one GPU each), halo exchange by CUDA IPC peer copies (NVLink P2P), no collective on the data path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_strips.py [--side 16384] [--steps 5]

Timed region per step (planes resident in HBM as int32, like bench.py's `value`): reset + exchange 1 + phase 0 + exchange 2
+ phase 1, bracketed by barriers; time = max over ranks.  The image is a seeded 2048x2048 RGB texture tiled over the
plane (generating 268 Mpixel with the cell generator takes minutes); rank 0 prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from yaik_b200 import capi, strips
    from yaik_b200.synth import make_image, SEED_BASE
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # control plane only: IPC handles, barriers
    side = args.side
    rows = strips.strip_rows(side, world)
    y0, sh = rows[rank]
    base = make_image(2048, 2048, 3, SEED_BASE + 3)
    reps = (side + 2047) // 2048
    mine = np.ascontiguousarray(np.tile(base, (1, reps, reps))[:, y0:y0 + sh, :side])
    lib = capi.load_library()
    ctx = capi.Context(side, sh, planes=3, slots=1, device=local, lib=lib)
    ctx.set_upload_format(False)
    ctx.set_image(mine, 0)
    ctx.strip_config(side, y0)
    halo = ctx.strip_halo()
    handles = [None] * world
    dist.all_gather_object(handles, ctx.ipc_export(halo.haloIn))
    above = rank - 1 if rank > 0 else None
    below = rank + 1 if rank + 1 < world else None
    peers = {r: ctx.ipc_open(handles[r]) for r in (above, below) if r is not None}

    def step():
        ctx.reset_state(0)
        ctx.strip_config(side, y0)
        if above is not None:
            for p in range(3):
                ctx.copy_async(peers[above] + halo.pixelRowInOffset + p * halo.pixelRowStride, halo.pixelRowOut[p], halo.planeRowBytes)
        ctx.sync(); dist.barrier()
        ctx.strip_phase(0)
        if above is not None:
            ctx.copy_async(peers[above] + halo.touchInBottomOffset, halo.touchOutTop, halo.touchBytes)
        if below is not None:
            ctx.copy_async(peers[below] + halo.touchInTopOffset, halo.touchOutBottom, halo.touchBytes)
        ctx.sync(); dist.barrier()
        ctx.strip_phase(1)
        ctx.sync()

    for _ in range(args.warmup):
        step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dist.barrier()
    dt = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([dt], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rb = ctx.result_bytes(0)
    if rank == 0:
        mp = side * side / 1e6
        print(json.dumps({"metric": "encoded megapixels/sec (gradient+range stages)", "workload": f"one {side}x{side} synthetic RGB image in {len(rows)} tile-row strips, one GPU each, NVLink P2P halo exchange (CUDA IPC), no collective",
                          "value": round(mp / float(t.item()), 1), "unit": "MP/s", "n_gpus": world, "ms_per_image": round(1e3 * float(t.item()), 3),
                          "steps": args.steps, "strip_rows": rows, "halo_bytes_per_boundary": int(3 * halo.planeRowBytes + 2 * halo.touchBytes),
                          "result_bytes_rank0": rb}), flush=True)
    for ptr in peers.values():
        ctx.ipc_close(ptr)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
