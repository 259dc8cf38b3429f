"""Summarises an `ncu --set full` report (.ncu-rep) into the few numbers DESIGN.md / bench.py cite.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.md

Runs here (no GPU needed): it only reads the report through `ncu -i ... --page raw --csv`.
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> L1 read bytes"),
    ("l1tex__m_l1tex2xbar_write_bytes.sum", "L1 -> L2 write bytes (stores + reductions)"),
    ("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "L1 -> L2 request path busy %"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 reduction sectors (REDG)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data pipe busy %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps / issue)"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{path}`\n")
    print("Captured with `ncu --set full --clock-control none --import-source on` (cold caches, serialised replays:")
    print("compare shares and ratios, not absolute times with bench.py).\n")
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("msda::", "") for r in data]
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in METRICS:
        if key not in idx:
            continue
        u = units[idx[key]]
        print(f"| {label} ({u}) | " + " | ".join(r[idx[key]] for r in data) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
