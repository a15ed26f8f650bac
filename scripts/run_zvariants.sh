for v in "" _zf_p2t64 _zf_w1 _zf_w4; do
  if [ -z "$v" ]; then unset DBSPM_B200_LIB; else export DBSPM_B200_LIB=$PWD/variants/libdbspm_b200$v.so; fi
  python scripts/time_corr.py pyridine large_slab 2>&1 | tail -2
done
