# A/B: library variants vs the in-tree build on the same box: headline timings
for v in "$@"; do
  if [ $v = cur ]; then unset ZFISTA_B200_LIB; else export ZFISTA_B200_LIB=$PWD/profiles/variants/libzf_$v.so; fi
  echo "=== $v"
  for w in fds jos1 jos1_l1; do
    python bench.py --workload $w --no-extras --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w ms_per_step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
  done
done
