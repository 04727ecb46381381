set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; tail -3 gpurun_out/r2_pytest_final.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
P="python scripts/profile_search.py --bench-data 1 --d1 12"
export TRAFFIC_JSON=gpurun_out/r2_traffic.json
$P --ef 64 > gpurun_out/r2_prof_plain_ef64.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:beam_kernel -c 1 -f -o /tmp/r2_search_ef64 $P --ef 64 > gpurun_out/r2_prof_ncu_ef64.log 2>&1
python scripts/ncu_summary.py /tmp/r2_search_ef64.ncu-rep 40 > gpurun_out/r2_search_ef64_summary.txt 2>&1
python scripts/make_traffic_json.py sift:64=/tmp/r2_search_ef64.ncu-rep > /dev/null 2> gpurun_out/r2_traffic.err
$P --ef 256 > gpurun_out/r2_prof_plain_ef256.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:beam_kernel -c 1 -f -o /tmp/r2_search_ef256 $P --ef 256 > gpurun_out/r2_prof_ncu_ef256.log 2>&1
python scripts/ncu_summary.py /tmp/r2_search_ef256.ncu-rep 40 > gpurun_out/r2_search_ef256_summary.txt 2>&1
python scripts/make_traffic_json.py sift:256=/tmp/r2_search_ef256.ncu-rep > /dev/null 2>> gpurun_out/r2_traffic.err
G="python scripts/profile_search.py --bench-data 1 --d1 16 --d 960"
$G --ef 64 > gpurun_out/r2_prof_plain_gist.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:beam_kernel -c 1 -f -o /tmp/r2_search_gist $G --ef 64 > gpurun_out/r2_prof_ncu_gist.log 2>&1
python scripts/ncu_summary.py /tmp/r2_search_gist.ncu-rep 40 > gpurun_out/r2_search_gist_ef64_summary.txt 2>&1
python scripts/make_traffic_json.py gist:64=/tmp/r2_search_gist.ncu-rep > /dev/null 2>> gpurun_out/r2_traffic.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_n1.csv $B > gpurun_out/r2_bench_underncu.json 2> gpurun_out/r2_bench_underncu.err
cat gpurun_out/r2_traffic.json | head -12
