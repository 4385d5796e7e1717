"""Summarise ncu reports (gpurun_out/*.ncu-rep) into profiles/ncu_summary_<tag>.{json,md}.

    python tools/summarize_ncu.py r01 gpurun_out/r01_*.ncu-rep
"""
import csv, io, json, os, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "regs",
    "launch__occupancy_limit_registers": "occ_limit_regs_ctas",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_ctas",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [_one(hdr, units, vals) for vals in rows[2:] if vals]


def _one(hdr, units, vals):
    d = {"kernel": vals[hdr.index("Kernel Name")]}
    for i, h in enumerate(hdr):
        if h in WANT:
            v = vals[i].replace(",", "")
            try:
                v = float(v)
            except ValueError:
                pass
            u = units[i]
            if isinstance(v, float):
                if u == "ns": v, u = v / 1e3, "us"
                if u == "ms": v, u = v * 1e3, "us"
                if u == "Kbyte": v, u = v * 1e3, "byte"
                if u == "Mbyte": v, u = v * 1e6, "byte"
                if u == "Gbyte": v, u = v * 1e9, "byte"
            d[WANT[h]] = v
            d[WANT[h] + "_unit"] = u
    if "dram_read" in d and "dram_write" in d:
        d["dram_bytes"] = d["dram_read"] + d["dram_write"]
        d["dram_gbs"] = d["dram_bytes"] / (d["time_us"] * 1e-6) / 1e9
    return d


def main():
    tag = sys.argv[1]
    res = [d for p in sys.argv[2:] for d in read(p)]
    os.makedirs("profiles", exist_ok=True)
    json.dump(res, open(f"profiles/ncu_summary_{tag}.json", "w"), indent=1)
    with open(f"profiles/ncu_summary_{tag}.md", "w") as f:
        f.write(f"# ncu --set full summaries ({tag}); 131072 Gaussians, 512x512, one view, B200, --clock-control none\n\n")
        f.write("| kernel | time us | DRAM bytes | DRAM GB/s | DRAM % peak | issue active % | occupancy % | regs | warp instr | FMA % | ALU % | XU % | FP64 % | LSU % |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for d in res:
            f.write("| {k} | {t:.1f} | {b:.3g} | {g:.0f} | {p:.2f} | {ia:.1f} | {oc:.1f} | {r:.0f} | {wi:.3g} | {fma:.1f} | {alu:.1f} | {xu:.1f} | {f64:.1f} | {lsu:.1f} |\n".format(
                k=d["kernel"].split("(")[0], t=d["time_us"], b=d.get("dram_bytes", 0), g=d.get("dram_gbs", 0), p=d.get("dram_pct_of_peak", 0),
                ia=d.get("issue_active_pct", 0), oc=d.get("achieved_occupancy_pct", 0), r=d.get("regs", 0), wi=d.get("warp_instructions", 0),
                fma=d.get("pipe_fma_pct", 0), alu=d.get("pipe_alu_pct", 0), xu=d.get("pipe_xu_pct", 0), f64=d.get("pipe_fp64_pct", 0), lsu=d.get("pipe_lsu_pct", 0)))
    print(open(f"profiles/ncu_summary_{tag}.md").read())


if __name__ == "__main__":
    main()
