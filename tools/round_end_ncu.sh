# ncu evidence of the bench command: launch list, then one --set full capture of the dominant kernel (after a clean run)
T=${1:-r04k}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plugin-leg > gpurun_out/${T}_b.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plugin-leg > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:skb_score_pq -s 4 -c 1 -f -o gpurun_out/${T}_score_d8_bench python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plugin-leg > gpurun_out/${T}_ncu2.log 2>&1
ls -la gpurun_out/${T}_*
