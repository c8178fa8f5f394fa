set -x
python profiles/run_op.py fwdbwd 8 4096 32 > gpurun_out/runop.log 2>&1 || exit 1
python profiles/bench_tmix.py > gpurun_out/tmix3.json 2>&1
python profiles/bench_bi.py > gpurun_out/bench_bi3b.json 2>&1
python profiles/bench_elementwise.py > gpurun_out/elementwise3.jsonl 2>&1
python profiles/bench_add_ln.py > gpurun_out/add_ln3.json 2>&1
# launch list of the op step
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2f_launches.csv python profiles/run_op.py fwdbwd 8 4096 32 > gpurun_out/ncu_l2.log 2>&1
# full capture of the two kernels (the third launch of each)
ncu --set full --clock-control none --import-source on -k regex:wkv6_tc3 -s 4 -c 2 -o gpurun_out/r2f_tc3 -f python profiles/run_op.py fwdbwd 8 4096 32 > gpurun_out/ncu_full2.log 2>&1
# elementwise kernels: add_ln fwd, gn_gate, ce_fwd/bwd
ncu --set full --clock-control none --import-source on -k regex:add_ln_fwd -s 2 -c 1 -o gpurun_out/r2f_add_ln -f python profiles/bench_add_ln.py > gpurun_out/ncu_addln.log 2>&1
tail -2 gpurun_out/tmix3.json; tail -2 gpurun_out/bench_bi3b.json; tail -1 gpurun_out/add_ln3.json; ls -la gpurun_out/r2f_*
