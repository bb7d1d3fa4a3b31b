# reduced end-of-round set: GPU tests, both bench arms, target shape, ncu evidence of the bench command
T=${1:-r04o}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${T}_gputests.txt
python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err
python tools/run_target.py > gpurun_out/${T}_target_8d_1e6x1e8_1gpu.json 2> gpurun_out/${T}_target.err
cat gpurun_out/${T}_gputests.txt; head -c 300 gpurun_out/${T}_bench_1gpu.json; echo
bash tools/round_end_ncu.sh ${T}
