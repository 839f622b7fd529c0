# usage: bash tools/variant_sweep.sh "" _g2 ...   -- bench each lib/libfecb200<suffix>.so (built with FECB200_VARIANT)
for v in "$@"; do
  FECB200_LIB=$PWD/finiteelementcontainers.jl_b200/lib/libfecb200$v.so timeout 300 python bench.py --no-cpu --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant[$v]', d['ms_per_step'], d['roofline']['kernel_ms'], d['ops']['tangent_ms'])"
done
