"""Two-GPU check of Sbdart.run_sharded over NCCL (run under torchrun, one rank per GPU):
the sharded records must equal the single-GPU records byte for byte.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/sharded_run_gpu.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbdart_b200 as sb                      # noqa: E402
from sbdart_b200.frontend import Sbdart      # noqa: E402
from solvers import make_solve_cuda          # noqa: E402

NAMELISTS = [
    "&INPUT idatm=2, wlinf=.25, wlsup=4.0, wlinc=.005, nstr=16, iout=1 /",
    "&INPUT idatm=2, nstr=8, iaer=1, vis=23, wlinf=.5, wlsup=.7, wlinc=.01, sza=30, iout=20, uzen=0,40,80,140, phi=0,90,180 /",
]


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    s = sb.Solver(local)
    solve = make_solve_cuda(s)
    ok = True
    for nl in NAMELISTS:
        sharded = Sbdart(nl).run_sharded(solve, dist)
        single = Sbdart(nl).run(solve)
        same = sharded == single
        ok = ok and same
        if dist.get_rank() == 0:
            print(f"rank0: {len(single.splitlines())} record lines, sharded == single: {same}", flush=True)
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if dist.get_rank() == 0:
        print("ALL RANKS IDENTICAL" if int(flag.item()) else "MISMATCH", flush=True)
    s.close()
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
