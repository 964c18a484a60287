"""Run bench.py's control flow (single- and multi-rank arms) on the CPU emulation of the CUDA sources, with a tiny
workload: checks the JSON contract and the rank plumbing in the no-GPU container.  TEST INFRASTRUCTURE; the numbers it
prints are meaningless as performance.  Usage (optionally under torchrun):  python tests/emu_bench_check.py [bench args]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_emu"))
import ddcmd_b200 as dd  # noqa: E402
import build_emu  # noqa: E402

dd._lib = dd._declare(ctypes.CDLL(build_emu.build()))
import bench  # noqa: E402
import torch  # noqa: E402

torch.Tensor.pin_memory = lambda self, *a, **k: self   # no driver here: page-locked staging is a GPU-box concern

if __name__ == "__main__":
    sys.argv = ["bench.py"] + sys.argv[1:]
    bench.main()
