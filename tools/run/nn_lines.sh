cd $GRAFT_REPO_ROOT
mkdir -p /tmp/rep gpurun_out/prof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_stream_kernel -s 20 -c 1 -f -o /tmp/rep/nn_stream python tools/prof_nn_single.py 1000000 > gpurun_out/ncu_nn_stream.log 2>&1
python tools/ncu_lines.py /tmp/rep/nn_stream.ncu-rep 40 > gpurun_out/prof/r02_nn_stream_lines.txt 2>&1
ncu -i /tmp/rep/nn_stream.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]
for m in ['sm__cycles_active.avg','sm__cycles_elapsed.max','gpu__time_duration.sum','smsp__warps_launched.sum','sm__ctas_launched.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum']:
    if m in h: print(m, r[2][h.index(m)], r[1][h.index(m)])
" >> gpurun_out/prof/r02_nn_stream_lines.txt
cat gpurun_out/prof/r02_nn_stream_lines.txt
