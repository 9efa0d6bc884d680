for f in gr-ldpc_ece535a_b200/build/variants/lib_*.so; do
  echo "== $f"
  LDPC535_LIB=$PWD/$f python tools/profile_kernels.py --which encode_generic 2>&1 | grep encode_generic
done
