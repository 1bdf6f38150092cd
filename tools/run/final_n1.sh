set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 300 gpurun_out/final_ref.json
python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; tail -c 300 gpurun_out/final_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:validity_persistent_kernel -s 4 -c 1 -f -o gpurun_out/r2f_validity python bench.py --steps 3 --warmup 3 --no-sections --no-cpu-baseline > gpurun_out/ncu_val.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:validity_scene_kernel -s 3 -c 1 -f -o gpurun_out/r2f_scene python tools/prof_scene.py 256 > gpurun_out/ncu_scene.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_stream_kernel -s 20 -c 1 -f -o gpurun_out/r2f_nn_stream python tools/prof_nn_single.py > gpurun_out/ncu_nn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_check_kernel -s 10 -c 1 -f -o gpurun_out/r2f_edge python tools/prof_edge.py > gpurun_out/ncu_edge.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
