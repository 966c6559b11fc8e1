# Round-2 measurement run on ONE B200 (gpurun): tests, smoke, bench (both arms), kernel report, ncu launch list, one ncu
# capture of the single-sample kernel.  Everything lands in gpurun_out/; tools/make_profiles_r02.py turns it into profiles/.
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_gputest.log; tail -2 gpurun_out/r02_gputest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02_bench_n1.err > gpurun_out/r02_bench_n1.json; tail -c 300 gpurun_out/r02_bench_n1.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2> gpurun_out/r02_ref.err > gpurun_out/r02_bench_reference_arm.json; tail -c 300 gpurun_out/r02_bench_reference_arm.json
timeout 600 python tools/kernel_report.py > gpurun_out/r02_kernel_report.json 2> gpurun_out/r02_kernel_report.err
SLU_NO_PACKED=1 timeout 300 python tools/kernel_report.py evidential fused Dirichlet > gpurun_out/r02_kernel_report_scalar.json 2>/dev/null
ARGS="--steps 3 --warmup 3 --windows 2 --no-cpu-baseline --no-e2e --no-ncu-traffic --skip config1,config3,config4,config5"
timeout 300 python bench.py $ARGS > gpurun_out/r02_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/r02_ncu_launches.log 2>&1
timeout 120 python tools/run_single_pass_once.py > gpurun_out/r02_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:reduce_single -s 3 -c 1 -o gpurun_out/r02_single_prof python tools/run_single_pass_once.py > gpurun_out/r02_ncu_single.log 2>&1
echo done
