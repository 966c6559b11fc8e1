set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py 2>gpurun_out/bench_r01b.err > gpurun_out/bench_r01b_n1.json; tail -c 600 gpurun_out/bench_r01b_n1.json
timeout 600 python bench.py --impl reference 2>/dev/null > gpurun_out/bench_r01b_ref.json
timeout 600 python tools/stage_report.py > gpurun_out/stages_r01c.json 2>gpurun_out/stages_r01c.err; tail -n 2 gpurun_out/stages_r01c.err
timeout 600 python tools/kernel_report.py > gpurun_out/kernel_report_r01.json 2>gpurun_out/kernel_report_r01.err
timeout 300 python tools/hist_bench.py > gpurun_out/hist_bench_r01.json 2>/dev/null
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_r01c.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r01c.log 2>&1
echo done
