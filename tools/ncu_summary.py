"""Summarise raw-page CSV exports of ncu captures (tools/ncu_capture.sh) into the few numbers DESIGN.md / profiles/ quote."""
import csv, sys, os
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe cycles %"),
    ("sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active", "dmma pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
def load(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}
for p in sys.argv[1:]:
    d = load(p)
    print(f"== {os.path.basename(p)}  kernel {d.get('Kernel Name', ('?',))[0]}")
    for k, name in KEYS:
        hit = [h for h in d if h == k]
        if hit:
            v, u = d[hit[0]]
            print(f"   {name:28s} {v} {u}")
