"""Summaries of ncu outputs for profiles/: `launches` aggregates a gpu__time_duration launch list per kernel,
`full` extracts the roofline-relevant counters of every launch in a --set full report (needs `ncu` on PATH)."""
import collections, csv, re, subprocess, sys

def launches(fn):
    lines = [l for l in open(fn) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"kernel,launches,total_us,avg_us,share")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"\"{k}\",{v[0]},{v[1]:.1f},{v[1] / v[0]:.2f},{v[1] / tot:.4f}")
    print(f"\"TOTAL\",{sum(v[0] for v in agg.values())},{tot:.1f},,1.0")

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def full(fn):
    out = subprocess.run(["ncu", "-i", fn, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    print(",".join(f"\"{w} [{units[i]}]\"" for w, i in idx))
    for row in r[2:]:
        print(",".join(f"\"{row[i]}\"" for _, i in idx))

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
