# end-of-round measurement set on one box: GPU tests, bench (both arms), all configs, target shape, aux kernels
T=${1:-r04g}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${T}_gputests.txt
python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err
python tools/run_configs.py > gpurun_out/${T}_configs_fullsize.json 2> gpurun_out/${T}_configs.err
python tools/run_target.py > gpurun_out/${T}_target_8d_1e6x1e8_1gpu.json 2> gpurun_out/${T}_target.err
python tools/bench_aux.py > gpurun_out/${T}_bench_aux_hbm_kernels.json 2> gpurun_out/${T}_aux.err
python tools/perf_scan.py 1,2,3,4,5,6,8,12,16 1000000 > gpurun_out/${T}_perf_scan_dims.txt 2> gpurun_out/${T}_scan.err
cat gpurun_out/${T}_gputests.txt; head -c 300 gpurun_out/${T}_bench_1gpu.json; echo; tail -3 gpurun_out/${T}_perf_scan_dims.txt
