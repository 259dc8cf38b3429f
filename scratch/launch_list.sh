#!/bin/bash
# ncu launch list (time, instructions, occupancy, DRAM bytes) of the backward kernels of one encoder layer:
#   scratch/launch_list.sh OUT.csv [bench args]
out=$1; shift
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_red.sum --clock-control none -k regex:"win|fill" -s 3 -c 6 --csv --log-file $out python bench.py --workload encoder1 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra "$@" > /dev/null 2>&1
