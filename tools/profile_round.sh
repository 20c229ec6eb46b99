# Round profile, run on the GPU box (gpurun): plain runs first, each profiler pass only after the same command exited 0.
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r02'
# then, in the build container:
#   python tools/summarize_ncu.py launches gpurun_out/rNN_launches.csv > profiles/rNN_launches_bench.txt
#   python tools/summarize_ncu.py report gpurun_out/rNN_basefc.ncu-rep > profiles/rNN_k_basefc_ncu_full.txt   (count + finalize)
#   python tools/ncu_lines.py gpurun_out/rNN_basefc.ncu-rep k_basefc_count > profiles/rNN_k_basefc_count_lines.txt
#   python tools/make_traffic.py gpurun_out/rNN_basefc.ncu-rep <reads of the captured launch> gpurun_out/rNN_baf.ncu-rep 50000000
R=${1:-r02}
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${R}_bench_n1.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${R}_bench_short_plain.json 2>/dev/null; echo "short rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${R}_bench_under_ncu.json 2>/dev/null; echo "ncu list rc=$?"
python tools/prof_basefc.py 3e8 10000 60000 1 > /dev/null 2>&1; echo "prof plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_basefc_count|k_basefc_finalize_segs' -s 2 -c 3 \
    -o gpurun_out/${R}_basefc -f python tools/prof_basefc.py 3e8 10000 60000 1 > gpurun_out/${R}_ncu_basefc.log 2>&1; echo "ncu basefc rc=$?"
python tools/prof_baf_host.py 5e7 > /dev/null 2>&1; echo "baf plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_baf_scan -s 1 -c 1 \
    -o gpurun_out/${R}_baf -f python tools/prof_baf_host.py 5e7 > gpurun_out/${R}_ncu_baf.log 2>&1; echo "ncu baf rc=$?"
