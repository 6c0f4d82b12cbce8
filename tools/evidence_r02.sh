set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02s2_n1.json 2> gpurun_out/bench_r02s2_n1.err; tail -c 300 gpurun_out/bench_r02s2_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02s2_reference.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r02s2_reference.json
python tools/ncu_target.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02s2_ncu_launches.csv python tools/ncu_target.py > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
timeout 400 ncu --set full --clock-control none -k regex:"conv1d_tc_kernel" -s 140 -c 8 -o gpurun_out/r02s2_conv1d_tc_c64 python tools/ncu_target.py > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log
timeout 400 ncu --set full --clock-control none -k regex:"conv1d_tc_kernel" -s 160 -c 8 -o gpurun_out/r02s2_conv1d_tc_c128 python tools/ncu_target.py > gpurun_out/ncu_e.log 2>&1; tail -1 gpurun_out/ncu_e.log
timeout 400 ncu --set full --clock-control none -k regex:"bn_bwd_lean|conv1d_wgrad_tc_kernel" -s 40 -c 10 -o gpurun_out/r02s2_bn_wgrad python tools/ncu_target.py > gpurun_out/ncu_f.log 2>&1; tail -1 gpurun_out/ncu_f.log
timeout 200 python tools/profile_step.py > gpurun_out/r02s2_step_profile.txt 2>&1
timeout 120 python tools/bench_bn.py > gpurun_out/r02s2_bench_bn.txt 2>&1
timeout 120 python tools/bench_wgrad.py > gpurun_out/r02s2_bench_wgrad.txt 2>&1
timeout 120 python tools/timeline_conv1d.py > gpurun_out/r02s2_timeline_conv1d.txt 2>&1
timeout 400 python tools/bench_configs.py > gpurun_out/bench_configs_r02s2.json 2> gpurun_out/bench_configs_r02s2.err; tail -c 400 gpurun_out/bench_configs_r02s2.json
ls -la gpurun_out/*.ncu-rep
