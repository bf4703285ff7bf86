mkdir -p gpurun_out
B="python bench.py --profile-step --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1d.csv $B > gpurun_out/p1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:gemm_tc --csv --log-file gpurun_out/gemm_dram_r1d.csv $B > gpurun_out/p2.log 2>&1
for k in conv_dw_kernel conv_norm_silu_kernel ctc_prefix_beam_kernel conv1_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -o gpurun_out/prof_$k -f $B > gpurun_out/p_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/launches_r1d.csv gpurun_out/gemm_dram_r1d.csv
