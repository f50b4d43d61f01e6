"""tools/peer_check.py -- run under torchrun on N GPUs of one box: the in-library exchange of the vector
pull (gb200_peerbuf_*: peer stores over NVLink + device-side flags) against the single-GPU result.

Every rank owns a block of A's vectors, computes its block of w = A min.+ d (the SSSP step), publishes it
into all ranks' dense copies and waits; after every one of `--rounds` epochs every rank's copy must hold
exactly the w that rank 0 computes alone from the whole A, bit for bit (MIN_PLUS_FP64 is exact).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/peer_check.py --scale 18
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=18)
    ap.add_argument("--rounds", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    os.environ["GB200_DEVICE"] = str(local)
    import bench
    import graphblas_b200 as gb
    gb.init(local)
    a = argparse.Namespace(workload="sssp", scale=args.scale, ef=16, bfs_dir="push")
    w = bench.make_workload(gb, a, f"cuda:{local}")
    A, d = w["A"], w["B"]
    n = A.vdim
    bounds = gb.partition_by_flops(A.p, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    mine = gb.DMatrix(bench.slice_vectors(gb, A, lo, hi))
    whole = gb.DMatrix(A) if rank == 0 else None

    def allgather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    pb = gb.PeerBuf(n, "FP64", rank, world, allgather_bytes)
    ok = True
    x = d.x.copy()
    for r in range(args.rounds):
        dv = gb.Matrix(n, 1, np.array([0, n]), np.arange(n), x, None, "FP64")
        dd = gb.DMatrix(dv)
        rh, info = gb.axb_device_keep(None, False, mine, dd, w["semiring"], True)
        pb.publish(rh)
        gb.free_result(rh)
        pb.wait()
        vals, pres = pb.read()
        if rank == 0:
            ref = gb.axb_device(None, False, whole, dd, w["semiring"], True).matrix
            want_p = np.zeros(n, dtype=bool)
            want_p[ref.i] = True
            same = np.array_equal(pres, want_p) and np.array_equal(vals[ref.i].view(np.uint64), ref.x.view(np.uint64))
            print(f"round {r}: nnz(w) {ref.nnz}, identical {same}", flush=True)
            ok = ok and same
        # the next round relaxes from d' = min (d, w) (computed on the host here: this is a check, not a bench)
        x = np.where(pres, np.minimum(x, vals), x)
        dd.free()
        flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    pb.free()
    if rank == 0:
        print("peer_check:", "ok" if ok else "FAILED", f"({world} GPUs, n={n}, {args.rounds} rounds)")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
