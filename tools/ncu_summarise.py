"""Summarise ncu outputs into profiles/ (tracked).
   python tools/ncu_summarise.py launches <launches.csv> <out.txt>
   python tools/ncu_summarise.py full <report.ncu-rep> <out.txt> [gemm_traffic.json]"""
import collections, csv, json, subprocess, sys


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "usecond": v, "nsecond": v / 1e3, "msecond": v * 1e3}.get(row["Metric Unit"], v)
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("gava::", "")
        agg[name][0] += 1; agg[name][1] += v; n += 1
    tot = sum(v for _, v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {n} consecutive launches (~1 forward step), "
                f"cold-cache + serialised: compare SHARES\n")
        f.write(f"{'kernel':50s} {'n':>4s} {'total_us':>10s} {'share':>7s} {'avg_us':>9s}\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:50]:50s} {c:4d} {v:10.1f} {v / tot:7.3f} {v / c:9.1f}\n")
        f.write(f"{'TOTAL':50s} {n:4d} {tot:10.1f}\n")
    print(open(out).read())


KEYS = ["launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg"]


def full(rep, out, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    gemm = []
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch per line ({rep.split('/')[-1]})\n")
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(CUtensorMap")[0].split("(const")[0].replace("void ", "").replace("gava::", "")[:44]
            f.write(name + "\n")
            for k in KEYS:
                if k in ix:
                    f.write(f"    {k} [{units[ix[k]]}] = {r[ix[k]]}\n")
            if "gemm2" in name:
                to_b = lambda v, u: float(v) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
                rd = to_b(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
                wr = to_b(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
                gemm.append((name, rd + wr, float(r[ix["gpu__time_duration.sum"]])))
    print(open(out).read())
    if traffic_json and gemm:
        json.dump({"source": rep.split("/")[-1], "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch of gemm2_tcgen05_kernel (the CTA-pair kernel: the "
                   "large GEMMs QKV, out-proj, fc1, fc2 of a transformer layer): mean per epilogue kind, then over the kinds",
                   # one figure per epilogue kind first (a capture window may hold two fc1 and one QKV), then their mean
                   "dram_bytes_per_launch": (lambda by: sum(sum(v) / len(v) for v in by.values()) / len(by))(
                       {n: [g[1] for g in gemm if g[0] == n] for n in {g[0] for g in gemm}}),
                   "per_launch": [{"kernel": g[0], "dram_bytes": g[1], "us": g[2]} for g in gemm]}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
