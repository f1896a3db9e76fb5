import sys, os
sys.path.insert(0, os.getcwd())
import bench, argparse
a = argparse.Namespace(gpus=1, steps=2, warmup=2, impl="ours", agents=10_000_000, subdiv=80, ref_subdiv=80, ref_steps=1, ref_warmup=1, no_cpu_baseline=True)
print(bench.run_ours(a, 0, 1)["roofline"])
