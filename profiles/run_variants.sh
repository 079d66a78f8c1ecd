for v in t1 t2 t4; do
  export ZFISTA_B200_LIB=$PWD/profiles/variants/libzf_$v.so
  echo "=== $v"; python profiles/check_digest.py 2>&1 | tail -14
  python bench.py --no-extras --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fds ms_per_step', d['ms_per_step'], 'value', d['value'], 'nit_max', d['nit_max_rank0'])"
  python bench.py --workload jos1 --no-extras --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('jos1 ms_per_step', d['ms_per_step'], 'value', d['value'])"
  python bench.py --workload jos1_l1 --no-extras --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('jos1_l1 ms_per_step', d['ms_per_step'], 'value', d['value'])"
done
