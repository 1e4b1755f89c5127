"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) as a markdown
table + per-kernel DRAM traffic JSON.   python tools/ncu_summary.py raw.csv out.md [traffic.json] [last_n_launches]"""
import csv, json, sys, collections

COLS = [("gpu__time_duration.sum", "us", 1.0), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1),
        ("launch__registers_per_thread", "regs", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma-pipe%", 1), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 1),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%", 1),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%", 1),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%", 1), ("lts__t_sector_hit_rate.pct", "l2hit%", 1),
        ("dram__bytes_read.sum", "dram rd MB", 1), ("dram__bytes_write.sum", "dram wr MB", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("smsp__inst_executed.sum", "warp inst", 1), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst", 1)]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["| kernel | " + " | ".join(c[1] for c in COLS) + " |", "|---|" + "---|" * len(COLS)]
    traffic = collections.OrderedDict()
    body = [r for r in rows[2:] if len(r) >= len(hdr)]
    if len(sys.argv) > 4:
        body = body[-int(sys.argv[4]):]   # the last N launches = one steady-state scan
    for r in body:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("gm::", "")
        cells = []
        for key, label, _ in COLS:
            v = num(r[idx[key]]) if key in idx else None
            u = units[idx[key]] if key in idx else ""
            if v is None:
                cells.append("-")
                continue
            if label == "us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            if label.startswith("dram") and label.endswith("MB"):
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                v *= scale
            cells.append(f"{v:.2f}" if abs(v) < 1e4 else f"{v:.3g}")
        out.append(f"| {name} | " + " | ".join(cells) + " |")
        rd = num(r[idx["dram__bytes_read.sum"]]); wr = num(r[idx["dram__bytes_write.sum"]])
        sc = lambda v, u: v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        if rd is not None and wr is not None:
            b = sc(rd, units[idx["dram__bytes_read.sum"]]) + sc(wr, units[idx["dram__bytes_write.sum"]])
            traffic.setdefault(name, []).append(b)
    open(sys.argv[2], "a").write("\n".join(out) + "\n")
    if len(sys.argv) > 3 and sys.argv[3] != "-":
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
