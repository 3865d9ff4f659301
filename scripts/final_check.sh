cd /root/repo
python -m pytest tests -m gpu -q -x > gpurun_out/final_gpu_tests.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
cat gpurun_out/final_bench.json
python bench.py --steps 2 --warmup 3 --stream-frames 4 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --stream-frames 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches_bench.csv
