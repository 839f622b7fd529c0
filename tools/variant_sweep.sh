# usage: bash tools/variant_sweep.sh "" _g2 ...   -- bench each lib/libfecb200<suffix>.so (built with FECB200_VARIANT)
for v in "$@"; do
  FECB200_LIB=$PWD/finiteelementcontainers.jl_b200/lib/libfecb200$v.so timeout 300 python bench.py --no-cpu --steps 10 2>gpurun_out/sweep_err$v.log | tail -1 > gpurun_out/sweep$v.json
  python -c "import json,sys; d=json.loads(open('gpurun_out/sweep$v.json').read()); print('variant[$v]', d['ms_per_step'], d['roofline']['kernel_ms'], d['ops']['tangent_ms'], d['ops']['residual_ms'])" || tail -5 gpurun_out/sweep_err$v.log
done
