#!/usr/bin/env python3
"""Device-resident step time: speculation on/off"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import *
eng = seqalign.BatchAligner(0, seqalign.Scoring.sw_cli_default())
bs = []
for k in range(6):
    a, oa, b, ob = synthetic_batch(2 + 1000 * k, 100000, 150, 150)
    bs.append([torch.from_numpy(x).cuda() for x in (a, oa, b, ob)])
ds = torch.zeros(100000, dtype=torch.int32, device="cuda")
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
def run(k):
    a, oa, b, ob = bs[k % 6]
    eng.run_device(seqalign.SW, a.data_ptr(), oa.data_ptr(), b.data_ptr(), ob.data_ptr(), 100000, ds.data_ptr(), 0, 0, s.cuda_stream)
for k in range(5): run(k)
torch.cuda.synchronize(); t = time.perf_counter()
for k in range(50): run(k)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 50
print("run_device: %.3f ms/step -> %.0f GCUPS; kernel %.3f ms; spec hits/misses %s; kernel %s" % (dt * 1e3, 2250 / dt / 1e3, eng.last_kernel_ms, eng.speculation_stats(), eng.last_kernel))
