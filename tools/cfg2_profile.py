#!/usr/bin/env python
"""configs[1] shape (4096 x 128x128 Life, float32 actions, no sums) as a 200-step CUDA graph -- the
workload behind extras.cfg2_4096x128x128; run under ncu to capture step_stream_kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
wl = bench.StepWorkload(torch, dev, 4096, 128, 32, "B3/S23", False, 512)
print("us/step", 1e3 * bench.graph_rate(torch, wl, dev, False, 200, repeats=5, tail=False))
