set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_plain.json 2> gpurun_out/r01_bench_plain.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r01_bench_plain.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r01_bench_short_plain.json 2>/dev/null; echo "short rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r01_bench_under_ncu.json 2>/dev/null; echo "ncu list rc=$?"
XG_OVERLAP=0 python tools/prof_basefc.py 3e8 10000 60000 1 > /dev/null 2>&1; echo "prof plain rc=$?"
XG_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:k_basefc_count -s 1 -c 1 -o gpurun_out/r01_k_basefc_count -f python tools/prof_basefc.py 3e8 10000 60000 1 > gpurun_out/ncu_basefc.log 2>&1; echo "ncu full rc=$?"
