"""torchrun worker for test_two_gpu_sharded_nccl: sharded top-2 over NCCL == single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_match_b200 import backend, sharded, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
q, t = synth.make_pair(30000, 40001, seed=1239)
qd = torch.from_numpy(q).to(dev)
lo, hi = sharded.shard_range(len(t), rank, world)
keys, d2, idx = sharded.sharded_top2(qd, torch.from_numpy(t[lo:hi]).to(dev), lo)
fd2, fidx = backend.top2(qd, torch.from_numpy(t).to(dev))
ok = torch.equal(d2, fd2) and torch.equal(idx, fidx)
# the all-to-all exchange (every rank merges its slice of the queries) + ratio test + re-assembly
sidx, sd2, sratio, smask = sharded.ratio_match_sharded(qd, torch.from_numpy(t[lo:hi]).to(dev), lo, 0.7, full=True)
_, _, fratio, fmask = backend.ratio_match(qd, torch.from_numpy(t).to(dev), 0.7, want_ratio=True)
ok = ok and torch.equal(sd2, fd2) and torch.equal(sidx, fidx) and torch.equal(smask, fmask) and torch.equal(sratio, fratio)
q_lo, q_hi = sharded.shard_range(len(q), rank, world)
pidx, pd2, _, pmask = sharded.ratio_match_sharded(qd, torch.from_numpy(t[lo:hi]).to(dev), lo, 0.7, want_ratio=False)
ok = ok and torch.equal(pidx, fidx[q_lo:q_hi]) and torch.equal(pmask, fmask[q_lo:q_hi])
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0 and int(flag.item()) == 1:
    print("SHARDED_OK")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
