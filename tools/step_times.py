"""Streaming step-kernel timings (B=65536, Philox and injected noise) for one build of the library (DAD_LIB_PATH)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from dynamics_aware_diffusion_b200 import _native as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
r = bench.stream_step_roofline(torch.device("cuda", 0), bench.peaks(), B=B)
print(os.path.basename(N.LIB_PATH), "B=%d" % B, "philox %.1f us %.0f GB/s (%.1f%%) | injected %.1f us %.0f GB/s (%.1f%%)" % (
    r["philox"]["avg_launch_ms"] * 1e3, r["philox"]["achieved"], 100 * r["philox"]["frac"],
    r["injected_noise"]["avg_launch_ms"] * 1e3, r["injected_noise"]["achieved"], 100 * r["injected_noise"]["frac"]))
