import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.dev_bench_other import bench_k2
if __name__ == "__main__":
    bench_k2()
    bench_k2(n=4_000_000, F=64, k=100)
    bench_k2(n=1_000_000, F=256, k=100)
