#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): one sharded fv3jedi_lm dynamics step
(NL, TL, AD) over NCCL must reproduce the single-GPU result, and the distributed dot-product test must hold.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py [--nonhydro]
  ... tools/multigpu_check.py --module tracer_2d    (q_split = 0: the level maxima of the Courant numbers cross the ranks)
  ... tools/multigpu_check.py --module c2l_ord4     (cubed_to_latlon: D-grid halo update)
"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "fv3-jedi-linearmodel_b200")):
    sys.path.insert(0, p)


def main():
    nonhydro = "--nonhydro" in sys.argv
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import fv3lm
    from common import metrics
    from test_multirank import _inputs, _run_all, _MODULE_INPUTS
    module = sys.argv[sys.argv.index("--module") + 1] if "--module" in sys.argv else "step"
    N, K, ak, bk, f, act, p, dx, y = _MODULE_INPUTS[module]() if module != "step" else _inputs(nonhydro)
    h = fv3lm.FV3LM(fv3lm.default_config(N, K, rank=rank, nranks=world), ak, bk)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(h.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    h.comm_init_nccl(bytes(idt.cpu().numpy().tobytes()))
    h.set_metrics(metrics(N))
    res = _run_all(h, N, K, f, act, p, dx, y, module)
    tot = {}
    for k in sorted(res):
        t = torch.from_numpy(res[k]).cuda()
        dist.all_reduce(t)
        tot[k] = t.cpu().numpy()
    ok = True
    if rank == 0:
        h1 = fv3lm.FV3LM(fv3lm.default_config(N, K), ak, bk)
        h1.set_metrics(metrics(N))
        ref = _run_all(h1, N, K, f, act, p, dx, y, module)
        worst = 0.0
        for k in ref:
            e = np.abs(tot[k] - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-300)
            worst = max(worst, e)
            if not e < 1e-11:
                ok = False; print("MISMATCH", k, e)
        lhs = sum((tot["tl." + o] * y[o]).sum() for o in y)
        rhs = sum((dx[k] * tot["ad." + k]).sum() for k in act)
        dot = abs(lhs - rhs) / max(abs(lhs), abs(rhs))
        ok = ok and dot <= 1e-10
        nex, nb = h.comm_stats()
        print("multigpu_check module=%s world=%d layout=%dx%d nsub/rank=%d nonhydro=%s: max rel err vs single GPU %.2e, dot-product err %.2e, %d exchanges %.2f MB sent/rank -> %s"
              % (module, world, h.lx, h.ly, h.nsub, nonhydro, worst, dot, nex, nb / 1e6, "OK" if ok else "FAIL"))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
