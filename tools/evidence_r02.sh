set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 200 gpurun_out/bench_final_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_final_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_final_reference.json
python tools/ncu_target.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_ncu_launches.csv python tools/ncu_target.py > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
timeout 200 python tools/profile_step.py > gpurun_out/final_step_profile.txt 2>&1
timeout 200 python tools/trace_step_conv1d.py > gpurun_out/final_trace_step_conv1d.txt 2>&1; tail -3 gpurun_out/final_trace_step_conv1d.txt
timeout 200 python tools/trace_step_wgrad.py > gpurun_out/final_trace_step_wgrad.txt 2>&1; tail -3 gpurun_out/final_trace_step_wgrad.txt
timeout 400 python tools/bench_configs.py > gpurun_out/bench_configs_final.json 2> gpurun_out/bench_configs_final.err; tail -c 200 gpurun_out/bench_configs_final.json
