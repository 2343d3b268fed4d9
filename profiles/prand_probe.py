import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mazu_b200 as mz
gib = int(sys.argv[1]) if len(sys.argv) > 1 else 8
print(mz.measure_random_gather(gib << 30, 1 << 28, iters=2))
