for f in gr-ldpc_ece535a_b200/build/variants/lib_*.so; do
  echo "== $f"
  LDPC535_LIB=$PWD/$f python bench.py --codewords 4000000 --steps 3 --warmup 3 --no-cpu --kernel c4-thread 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['governing']['edge_iterations_per_s'])"
done
