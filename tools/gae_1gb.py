"""The GAE kernel at the >= 1 GB shape (T=128, B=2^21, the 24 B/element variant the driver launches) for ncu captures:
    ncu --set full --clock-control none -k regex:gae -s 3 -c 1 -o gpurun_out/r2_gae1g python tools/gae_1gb.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openrl_b200 import lib  # noqa: E402

L = lib.load()
dev = torch.device("cuda:0")
Tn, B = 128, 1 << 21
r = torch.randn(Tn, B, device=dev); vp = torch.randn(Tn + 1, B, device=dev)
m = (torch.rand(Tn + 1, B, device=dev) > 0.01).float(); act = torch.ones(Tn + 1, B, device=dev)
vn = torch.tensor([0.3, 2.0, 0.5], device=dev); ret = torch.empty(Tn + 1, B, device=dev)
adv = torch.empty(Tn, B, device=dev); st = torch.empty(8, dtype=torch.float64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    flush.zero_()
    lib.check(L.orl_gae(lib.ptr(r), lib.ptr(vp), lib.ptr(m), None, lib.ptr(act), lib.ptr(vp[Tn]), lib.ptr(vn), lib.ptr(ret), lib.ptr(adv),
                        lib.ptr(st), Tn, B, 0.99, 0.95, 5, s), "gae")
torch.cuda.synchronize()
print("ok")
