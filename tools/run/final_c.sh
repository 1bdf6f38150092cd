cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/prof /tmp/rep
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; tail -c 200 gpurun_out/final_n1.err
python bench.py --no-sections --no-cpu-baseline > gpurun_out/final_n1_b.json 2>/dev/null
cap() {
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/rep/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py - /tmp/rep/$name.ncu-rep r02f_$name r02f_$name > gpurun_out/sum_$name.log 2>&1
}
cap nn_stream nn_stream_kernel 20 python tools/prof_nn_single.py 1000000
cap edge_check edge_check_kernel 10 python tools/prof_edge.py
cp profiles/r02f_* gpurun_out/prof/ 2>/dev/null
ls gpurun_out/prof
