cd $GRAFT_REPO_ROOT
mkdir -p /tmp/rep gpurun_out/prof
python tools/prof_rrt.py 8 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rrt_fused_kernel -s 1 -c 1 -f -o /tmp/rep/rrt python tools/prof_rrt.py 8 > gpurun_out/ncu_rrt.log 2>&1
python tools/ncu_lines.py /tmp/rep/rrt.ncu-rep 70 > gpurun_out/prof/r02_rrt_fused_lines.txt 2>&1
python tools/ncu_summary.py - /tmp/rep/rrt.ncu-rep r02f_rrt_fused r02f_rrt_fused > gpurun_out/sum_rrt.log 2>&1
cp profiles/r02f_rrt_fused* gpurun_out/prof/
head -60 gpurun_out/prof/r02_rrt_fused_lines.txt
