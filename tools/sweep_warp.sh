for f in gr-ldpc_ece535a_b200/build/variants/lib_*.so; do
  echo "== $f"
  LDPC535_LIB=$PWD/$f python tools/early_stop_sweep.py 2>&1 | grep -E "2 dB max_iters  5|4 dB|8 dB max_iters  5" | sed 's/c4-thread.*| //'
  LDPC535_LIB=$PWD/$f python tools/profile_kernels.py --which decode_warp,methods 2>&1 | grep -E "decode_warp.*iters=50|method 0.*iters=5 "
done
