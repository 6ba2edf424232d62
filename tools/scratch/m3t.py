import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
os.environ["SLAMGPU_M3_TIMING"] = "1"
import slam_constructor_b200 as sg, latency_bench
ctx = sg.Context(0)
print(latency_bench.pyramid_numbers(ctx))
