#!/bin/bash
# usage: [NCU_SKIP=6 NCU_COUNT=1] bash tools/ncu_blend.sh TAG KERNEL_REGEX [ENV...]  -> gpurun_out/TAG/<name>.ncu-rep + raw csv of the key metrics
TAG=$1; KRE=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
name=$(echo "${*:-default}" | tr ' =' '__')
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" --launch-skip ${NCU_SKIP:-6} --launch-count ${NCU_COUNT:-1} \
  -o $OUT/full_$name -f python bench.py --no-cpu-baseline --no-other-workloads --steps 2 --warmup 3 > $OUT/ncu_$name.log 2>&1
ncu -i $OUT/full_$name.ncu-rep --page raw --csv > $OUT/raw_$name.csv 2>/dev/null
python - "$OUT/raw_$name.csv" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("==", name[:90])
    for k in keys:
        if k in hdr:
            print(f"  {k:95s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
PY
