set -x
python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/b_small.json 2> gpurun_out/b_small.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench_tensor.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_l.log 2>&1
python scripts/tc_one.py 0.1 && ncu --set full --clock-control none --import-source on -k regex:csr_tc_kernel -s 1 -c 1 -o gpurun_out/tc_final_d10 -f python scripts/tc_one.py 0.1 > gpurun_out/ncu_tc.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
