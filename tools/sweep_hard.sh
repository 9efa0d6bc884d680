for f in gr-ldpc_ece535a_b200/build/variants/lib_*.so; do
  echo "== $f"
  LDPC535_LIB=$PWD/$f python tools/profile_kernels.py --which methods --c4-codewords 10000000 2>&1 | grep "method 3.*early=0"
done
