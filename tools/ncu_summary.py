"""Summarise ncu exports (`--page raw --csv`) and launch lists into the text files kept under profiles/."""
import csv, re, sys, collections

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "pipe_fmaheavy_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefronts_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "lsu_wavefronts_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__average_warp_latency_issue_stalled_barrier.pct", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
]


def short(n):
    n = re.sub(r"void <unnamed>::", "", n)
    n = re.sub(r"<unnamed>::", "", n)
    return re.sub(r"\(.*", "", n)


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")])}
        for k, name in KEYS:
            if k in hdr:
                d[name] = (r[hdr.index(k)], units[hdr.index(k)])
        out.append(d)
    return out


def launches(path, step_marker="ssv_kernel"):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = [(short(r[ki]), float(r[vi].replace(",", "")), r[ui]) for r in rows[1:]]
    return seq


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "raw":
        for d in raw(path):
            print(d["kernel"])
            for k, v in d.items():
                if k != "kernel":
                    print("    %-22s %s %s" % (k, v[0], v[1]))
    else:
        seq = launches(path)
        print(len(seq), "launches")
