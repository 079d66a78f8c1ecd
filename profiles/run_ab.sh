# A/B: round-1 library vs the in-tree build on the same box: digest + headline timings
for v in r01 cur; do
  if [ $v = cur ]; then unset ZFISTA_B200_LIB; else export ZFISTA_B200_LIB=$PWD/profiles/variants/libzf_$v.so; fi
  echo "=== $v"
  python profiles/check_digest.py 2>&1 | grep -v identical
  for w in fds jos1 jos1_l1; do
    python bench.py --workload $w --no-extras --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w ms_per_step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
  done
done
