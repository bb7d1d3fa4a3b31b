python tools/perf_scan.py 1,2,3,4,5,6,8,12,16 1000000 > gpurun_out/r04i_perf_scan_dims.txt 2>/dev/null
python tools/perf_scan.py 1,2,3,4 100000 >> gpurun_out/r04i_perf_scan_dims.txt 2>/dev/null
python tools/run_configs.py --c4-queries 20000000 > gpurun_out/r04i_configs_fullsize.json 2> gpurun_out/r04i_configs.err
python tools/profile_concentric.py > gpurun_out/r04i_concentric_4shells_profile.txt 2>&1
cat gpurun_out/r04i_perf_scan_dims.txt; head -30 gpurun_out/r04i_concentric_4shells_profile.txt
