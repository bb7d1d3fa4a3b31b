# A/B of score-kernel build variants on bench-like data (bits must agree between variants that only move data differently).
#   tools/build_variant.sh NAME "-DFLAG=..."   builds tools/variants/NAME/libskb.so from the working tree
#   VARIANTS="NAME1 NAME2 default" bash tools/ab_pq.sh     ("default" = the in-tree library; the first variant sets the reference bits)
set -x
for v in ${VARIANTS:-default}; do
  if [ $v = default ]; then unset SKB_LIB_PATH; else export SKB_LIB_PATH=$PWD/tools/variants/$v/libskb.so; fi
  python tools/ab_pq.py 8,12,16 1000000 $v 2>&1 | grep -v Warning
done > gpurun_out/ab_pq.txt 2>&1
rm -f gpurun_out/ab_ref_*.npy
cat gpurun_out/ab_pq.txt
