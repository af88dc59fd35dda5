import sys, torch
sys.path.insert(0, ".")
import bench
from fast_match_b200 import backend
class A: groups = 10000
print(bench.grouped_leg(A, torch.device("cuda:0"), {}, backend, False)["ms"])
