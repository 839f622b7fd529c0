import os, sys
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT,"tools")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"finiteelementcontainers.jl_b200"))
import bench_configs as bc
print(bc.j2(64))
