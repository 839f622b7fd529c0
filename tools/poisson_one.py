"""Poisson hex8 n^3: a few residual / stiffness / action assemblies (driver for ncu captures of config 2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))
import bench_configs as bc  # noqa: E402

print(bc.poisson(int(sys.argv[1]) if len(sys.argv) > 1 else 128))
