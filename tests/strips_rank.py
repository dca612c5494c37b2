"""One rank of a strip run (used by tests/test_strips_host.py on gloo and by the 2-GPU test / bench on NCCL-capable boxes):
rank r analyses strip r of a seeded image; rank 0 checks the merged result against the oracle on the whole image."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, default=int(os.environ.get("RANK", "0")))
    ap.add_argument("--world", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--mode", default="host")
    ap.add_argument("--w", type=int, default=128)
    ap.add_argument("--h", type=int, default=192)
    args = ap.parse_args()
    import torch.distributed as dist
    import cases
    from yaik_b200 import capi, strips
    from strips_check import check_against_oracle

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group(args.backend, rank=args.rank, world_size=args.world)
    lib = capi.load_library(os.environ.get("YK_STRIPS_LIB") or None)
    device = 0
    if args.mode == "ipc":
        import torch
        device = args.rank % max(1, torch.cuda.device_count())
    planes = cases._patchy(args.w, args.h, 41, 4, 3)
    rows = strips.strip_rows(args.h, args.world)
    mine = planes[:, rows[args.rank][0]:rows[args.rank][0] + rows[args.rank][1]] if args.rank < len(rows) else None
    ctx = capi.Context(args.w, args.h, planes=3, slots=1, device=device, lib=lib)
    try:
        merged = strips.DistTransport(ctx, dist, mode=args.mode).run(mine, args.h, rows)
        if args.rank == 0:
            check_against_oracle(merged, planes)
            print("strips ok:", len(rows), "strips over", args.world, "ranks, transport", args.mode, flush=True)
    finally:
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
