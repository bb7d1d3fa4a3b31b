for v in default refexact; do
  if [ $v = default ]; then unset SKB_LIB_PATH; else export SKB_LIB_PATH=$PWD/tools/variants/$v/libskb.so; fi
  python tools/scan_new.py 8 1000000 2>&1 | grep -v Warning
done > gpurun_out/ab_ref.txt 2>&1
cat gpurun_out/ab_ref.txt
