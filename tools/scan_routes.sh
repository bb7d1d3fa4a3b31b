for k in 1 2; do python tools/scan_new.py 4,5,6,7,8 1000000 $k 2>&1 | grep -v Warning | sed "s/^/kernel=$k /"; done > gpurun_out/scan_routes.txt 2>&1
cat gpurun_out/scan_routes.txt
