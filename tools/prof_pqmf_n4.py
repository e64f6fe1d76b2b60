#!/usr/bin/env python
"""Run PQMF N = 4 analysis + synthesis a few times (for ncu): python tools/prof_pqmf_n4.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harness  # noqa: E402,F401
import ias_b200  # noqa: E402

dev = torch.device("cuda:0")
x = torch.rand((1024, 1, 176400), device=dev) * 2 - 1
m = ias_b200.PQMF(N=4).to(dev)
for _ in range(3):
    z = m.analysis(x)
    y = m.synthesis(z)
torch.cuda.synchronize()
