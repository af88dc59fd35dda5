"""Small end-to-end run of every entry point, meant to be executed under compute-sanitizer."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import oracle
from fast_match_b200 import backend, synth
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
q, t = synth.make_pair(700, 1300, seed=3)
od2, oidx = oracle.c_top2(q, t)
for algo in (1, 2):
    d2, idx, r, m = backend.ratio_match(d(q), d(t), 0.7, algo=algo, want_ratio=True)
    assert np.array_equal(d2.cpu().numpy().view(np.uint32), od2) and np.array_equal(idx.cpu().numpy(), oidx)
qp, qo, tp, to = synth.make_groups(40, 1, 300, seed=5)
gd2, gidx, gt2q = oracle.c_grouped_mutual(qp, qo, tp, to)
for algo in (1, 2):
    o = backend.grouped_mutual(d(qp), d(qo), d(tp), d(to), algo=algo)
    assert np.array_equal(o[0].cpu().numpy().view(np.uint32), gd2) and np.array_equal(o[2].cpu().numpy(), gt2q)
keys = torch.stack([backend.top2(d(q), d(t[:600]), want_keys=True)[2], backend.top2(d(q), d(t[600:]), t_index_base=600, want_keys=True)[2]])
_, md2, midx = backend.merge_top2(keys)
assert np.array_equal(md2.cpu().numpy().view(np.uint32), od2) and np.array_equal(midx.cpu().numpy(), oidx)
h = backend.top2_host(q, t, want_dist=True, tau=0.7)
assert np.array_equal(h[0], od2)
torch.cuda.synchronize()
print("sanitize_small ok")
