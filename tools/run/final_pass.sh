cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/prof /tmp/rep
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; tail -c 200 gpurun_out/final_n1.err
timeout 1400 python bench_ops.py --out gpurun_out/ops_r02.jsonl > gpurun_out/ops_r02.log 2>&1; wc -l gpurun_out/ops_r02.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/prof/launches_r02_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rrt_fused_kernel -s 1 -c 1 -f -o /tmp/rep/rrt python tools/prof_rrt.py 8 > gpurun_out/ncu_rrt.log 2>&1
python tools/ncu_lines.py /tmp/rep/rrt.ncu-rep 40 > gpurun_out/prof/r02_rrt_fused_lines.txt 2>&1
python tools/ncu_summary.py gpurun_out/prof/launches_r02_final.csv /tmp/rep/rrt.ncu-rep r02f_rrt_fused r02f_rrt_fused > gpurun_out/sum_rrt.log 2>&1
cp profiles/r02f_rrt_fused* gpurun_out/prof/
ls gpurun_out/prof
