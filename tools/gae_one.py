"""One GAE launch at the >=1 GB shape (for an ncu --set full capture of dram traffic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from openrl_b200 import lib
L = lib.load()
T, B = 128, 1 << 21
dev = torch.device("cuda:0")
r = torch.randn(T, B, device=dev); vp = torch.randn(T + 1, B, device=dev)
m = (torch.rand(T + 1, B, device=dev) > 0.01).float(); act = torch.ones(T + 1, B, device=dev)
vn = torch.tensor([0.3, 2.0, 0.5], device=dev); ret = torch.empty(T + 1, B, device=dev)
adv = torch.empty(T, B, device=dev); st = torch.empty(8, dtype=torch.float64, device=dev)
for _ in range(3):
    lib.check(L.orl_gae(lib.ptr(r), lib.ptr(vp), lib.ptr(m), None, lib.ptr(act), lib.ptr(vp[T]), lib.ptr(vn), lib.ptr(ret),
                        lib.ptr(adv), lib.ptr(st), T, B, 0.99, 0.95, 5, torch.cuda.current_stream().cuda_stream), "gae")
torch.cuda.synchronize()
