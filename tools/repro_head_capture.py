import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
try:
    print(bench.fused_head_record(dev, 256, 16000, 10, 6452.8))
except Exception:
    traceback.print_exc()
