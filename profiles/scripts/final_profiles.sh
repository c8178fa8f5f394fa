# The command list that produced the final round-2 captures on the GPU box (gpurun -- 'bash profiles/scripts/final_profiles.sh')
set -x
python profiles/run_op.py fwdbwd 8 4096 32 > gpurun_out/runop.log 2>&1 || exit 1
python profiles/bench_tmix.py > gpurun_out/tmix4.json 2>&1
python profiles/bench_bi.py > gpurun_out/bench_bi4b.json 2>&1
python profiles/bench_waves.py > gpurun_out/waves2.jsonl 2>&1
# launch list of the op step
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2g_launches.csv python profiles/run_op.py fwdbwd 8 4096 32 > gpurun_out/ncu_l3.log 2>&1
# full capture of the two kernels
ncu --set full --clock-control none --import-source on -k regex:wkv6_tc3 -s 4 -c 2 -o gpurun_out/r2g_tc3 -f python profiles/run_op.py fwdbwd 8 4096 32 > gpurun_out/ncu_full3.log 2>&1
tail -1 gpurun_out/tmix4.json; tail -1 gpurun_out/bench_bi4b.json; tail -6 gpurun_out/waves2.jsonl
