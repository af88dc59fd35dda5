"""Worker of test_gpu_pair_modes.py: runs with FM_TC_PAIR forced to 0 or 1 (the library reads it once
per process) and checks the dense tcgen05 kernel against the C oracle on boundary shapes."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import oracle
from fast_match_b200 import backend, synth

SHAPES = [(1, 1), (5, 300), (255, 255), (256, 256), (257, 513), (511, 64), (512, 1000), (513, 255),
          (1023, 2049), (1500, 700), (3000, 9000), (20000, 600)]


def main():
    dev = torch.device("cuda:0")
    for M, N in SHAPES:
        q, t = synth.make_pair(M, N, seed=1000 + M + N)
        if M >= 300:
            q[7] = q[3]
            t[min(N - 1, 11)] = t[2]                       # exact duplicates: ties go to the lowest index
        sel = np.unique(np.linspace(0, M - 1, min(M, 400)).astype(np.int64))
        od2, oidx = oracle.c_top2(q[sel], t, 17)
        d2, idx, keys = backend.top2(torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev),
                                     t_index_base=17, want_keys=True, algo=backend.FM_ALGO_TCGEN05)
        assert np.array_equal(d2.cpu().numpy().view(np.uint32)[sel], od2), (M, N)
        assert np.array_equal(idx.cpu().numpy()[sel], oidx), (M, N)
        assert np.array_equal(keys.cpu().numpy().view(np.uint64)[sel], oracle.pack_keys(od2, oidx)), (M, N)
    print("pair-mode worker ok")


if __name__ == "__main__":
    main()
