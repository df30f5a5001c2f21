"""Turn `ncu --set full` reports into the two tracked summaries of this directory.

    python profiles/ncu_extract.py gpurun_out/r01_full.ncu-rep [more.ncu-rep ...]

writes profiles/r01_ncu_key_metrics.csv (one row per captured launch, the counters the notes quote) and
profiles/r01_ncu_summary.json (the per-kernel numbers bench.py reads: DRAM bytes per unit for `roofline.traffic`).
Needs the `ncu` binary (reads reports only: no GPU).
"""
import csv, io, json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def rows_of(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rd[2:]:
        d = {"Kernel Name": r[col["Kernel Name"]], "Block Size": r[col["Block Size"]], "Grid Size": r[col["Grid Size"]]}
        for k in KEYS:
            if k in col and r[col[k]] != "":
                d[k] = float(r[col[k]].replace(",", "")) * SCALE.get(units[col[k]], 1.0)   # bytes in bytes, times in ms
        yield d


def main(reports):
    rows = [r for rep in reports for r in rows_of(rep)]
    with open(os.path.join(HERE, "r01_ncu_key_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name", "Block Size", "Grid Size"] + KEYS)
        w.writerow(["", "", "", ""] + ["ms" if k.startswith("gpu__time") else "byte" if "bytes" in k else "" for k in KEYS])
        for i, r in enumerate(rows):
            w.writerow([i, r["Kernel Name"], r["Block Size"], r["Grid Size"]] + [r.get(k, "") for k in KEYS])

    def biggest(name):
        c = [r for r in rows if name in r["Kernel Name"]]
        return max(c, key=lambda r: r["gpu__time_duration.sum"]) if c else None

    summ = {"source": "ncu --set full --clock-control none of profiles/prof_run.py (ct_mul over 1024 fresh pairs, faithful enc_value over 128 values, "
                      "ct_add over 2^14 synthetic pairs), B200, round 1 (final kernels); the longest captured launch of each kernel; "
                      "made by profiles/ncu_extract.py"}
    sg = biggest("sigma_fused_kernel")
    if sg:
        # every warp of the launch writes exactly one 1 KiB row per edge: edges = DRAM bytes written / 1 024, rounded by the
        # launch's known size when prof_run.py printed it (PVACB_NCU_EDGES), else estimated from the bytes
        edges = int(os.environ.get("PVACB_NCU_EDGES", "0")) or round(sg["dram__bytes_write.sum"] / 1026.8)
        alu_pct = sg["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]
        summ["sigma_fused_kernel"] = {
            "edges": edges, "gpu_time_ms": sg["gpu__time_duration.sum"],
            "dram_read_bytes": sg["dram__bytes_read.sum"], "dram_write_bytes": sg["dram__bytes_write.sum"],
            "dram_bytes_per_edge": (sg["dram__bytes_read.sum"] + sg["dram__bytes_write.sum"]) / edges,
            "warp_instructions_per_edge": sg["smsp__inst_executed.sum"] / edges,
            # ALU-pipe warp instructions per edge: the pipe issues one warp instruction per 2 cycles per SM sub-partition
            "alu_pipe_warp_instructions_per_edge": alu_pct / 100.0 * sg["sm__cycles_elapsed.avg"] * 0.5 * 4 * 148 / edges,
            "alu_pipe_pct": alu_pct, "fma_pipe_pct": sg["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"],
            "lts_throughput_pct": sg["lts__throughput.avg.pct_of_peak_sustained_elapsed"],
            "issue_active_pct": sg["smsp__issue_active.avg.pct_of_peak_sustained_active"],
            "warps_active_pct": sg["sm__warps_active.avg.pct_of_peak_sustained_active"],
            "no_instruction_stall_per_issue": sg.get("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
            "registers": int(sg["launch__registers_per_thread"]),
        }
    pr = biggest("prf_lpn_kernel")
    if pr:
        summ["prf_lpn_kernel"] = {
            "gpu_time_ms": pr["gpu__time_duration.sum"], "lsu_pipe_pct": pr["sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"],
            "alu_pipe_pct": pr["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"],
            "shared_wavefronts": pr.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "shared_bank_conflicts": pr.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
            "dram_read_bytes": pr["dram__bytes_read.sum"], "dram_write_bytes": pr["dram__bytes_write.sum"],
            "registers": int(pr["launch__registers_per_thread"]),
        }
    cc = biggest("concat_kernel")
    if cc:
        pairs = 1 << 14
        summ["concat_kernel"] = {
            "pairs": pairs, "gpu_time_ms": cc["gpu__time_duration.sum"],
            "dram_read_bytes": cc["dram__bytes_read.sum"], "dram_write_bytes": cc["dram__bytes_write.sum"],
            "dram_throughput_pct": cc["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"],
            "algorithmic_bytes": pairs * (2 * (40 * 1052 + 58) + 80 * 1052 + 108),
        }
    with open(os.path.join(HERE, "r01_ncu_summary.json"), "w") as f:
        json.dump(summ, f, indent=1)
    print(json.dumps(summ, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
