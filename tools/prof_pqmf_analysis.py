#!/usr/bin/env python
"""Run PQMF analysis (plain and pooled) once per shape (for ncu): python tools/prof_pqmf_analysis.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harness  # noqa: E402,F401
import ias_b200  # noqa: E402

dev = torch.device("cuda:0")
x = torch.rand((1024, 1, 176400), device=dev) * 2 - 1
for N in (16, 3):
    m = ias_b200.PQMF(N=N).to(dev)
    for _ in range(2):
        z = m.analysis(x)
    for _ in range(2):
        z, f = m.analysis_pooled(x, 256)
torch.cuda.synchronize()
