# ncu captures of the final build, summarised on the box (the reports themselves are too large to travel back)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/prof /tmp/rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/prof/launches_r02_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
cap() {  # name regex skip cmd...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/rep/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py - /tmp/rep/$name.ncu-rep r02f_$name r02f_$name > gpurun_out/sum_$name.log 2>&1
}
cap validity_persistent validity_persistent_kernel 4 python bench.py --steps 3 --warmup 3 --no-sections --no-cpu-baseline
cap validity_scene validity_scene_kernel 3 python tools/prof_scene.py 256
cap nn_stream nn_stream_kernel 20 python tools/prof_nn_single.py
cap edge_check edge_check_kernel 10 python tools/prof_edge.py
cap sample_tail sample_tail_kernel 1 python tools/bench_sampler_steer.py
python tools/ncu_summary.py gpurun_out/prof/launches_r02_final.csv /tmp/rep/validity_persistent.ncu-rep r02f_bench validity > gpurun_out/sum_launches.log 2>&1
cp profiles/r02f_* profiles/validity_ncu_summary.json gpurun_out/prof/ 2>/dev/null
ls -la gpurun_out/prof | head -40
