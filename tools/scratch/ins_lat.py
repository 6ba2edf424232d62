import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import slam_constructor_b200 as sg, latency_bench
ctx = sg.Context(0)
m = latency_bench.measure(ctx)
print({k: {kk: vv for kk, vv in v.items() if "append" in kk or "cells" in kk} for k, v in m.items()})
