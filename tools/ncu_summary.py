"""Summarise an .ncu-rep (raw page) and a launch-list csv into profiles/ text files (CPU box)."""
import csv, io, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__cluster", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield dict(zip(hdr, r)), dict(zip(hdr, units))


def launches(path):
    agg = collections.OrderedDict()
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    for r in csv.DictReader(io.StringIO(txt)):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit.startswith("u") else v if unit.startswith("m") else v * 1e3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v_ms
    return agg


if __name__ == "__main__":
    rep, lst, out = sys.argv[1], sys.argv[2], sys.argv[3]
    with open(out, "w") as f:
        f.write("# ncu summary (%s)\n\n## launch list (%s): per-kernel launches and total device time\n" % (rep, lst))
        agg = launches(lst)
        tot = sum(v[1] for v in agg.values())
        for k, (n, ms) in agg.items():
            f.write("%-70s launches %3d  total %10.3f ms  share %6.2f %%\n" % (k[:70], n, ms, 100 * ms / tot))
        f.write("\n## --set full capture of the dominant kernel (raw page, selected metrics)\n")
        for vals, units in raw(rep):
            f.write("kernel: %s grid %s block %s\n" % (vals.get("Kernel Name"), vals.get("Grid Size"), vals.get("Block Size")))
            for k in vals:
                if any(k == kk or k.startswith(kk) for kk in KEYS) and vals[k] != "" and ".max" not in k and ".min" not in k and ".sum.p" not in k:
                    f.write("  %-90s %s %s\n" % (k, vals[k], units.get(k, "")))
    print(open(out).read())
