for cfg in "1 8" "1 12" "1 16" "x x"; do set -- $cfg
  echo "== pack_pinned=$1 threads=$2"
  if [ "$1" != x ]; then export LDPC535_PACK_PINNED=$1 LDPC535_PACK_THREADS=$2; else unset LDPC535_PACK_PINNED LDPC535_PACK_THREADS; fi; python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('dev %.3f Gbit/s  e2e %.3f Gbit/s  %.2f ms  match=%s'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['matches_device_path']))"
done
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA" 
