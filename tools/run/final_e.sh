cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/prof /tmp/rep
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; tail -c 200 gpurun_out/final_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_stream_kernel -s 20 -c 1 -f -o /tmp/rep/nn_stream python tools/prof_nn_single.py 1000000 > gpurun_out/ncu_nn_stream.log 2>&1
python tools/ncu_summary.py - /tmp/rep/nn_stream.ncu-rep r02f_nn_stream r02f_nn_stream > gpurun_out/sum_nn_stream.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/prof/launches_r02_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_summary.py gpurun_out/prof/launches_r02_final.csv /tmp/rep/nn_stream.ncu-rep r02f_bench r02f_nn_stream > gpurun_out/sum_launches.log 2>&1
cp profiles/r02f_* gpurun_out/prof/
ls gpurun_out/prof | head
