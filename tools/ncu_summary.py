"""Here (no GPU needed): turn the .ncu-rep captures and the launch list brought back from the GPU box into the tracked
summaries under profiles/.  Usage: python tools/ncu_summary.py <round-prefix> <launch-list.csv> <name=file.ncu-rep> ..."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    prefix, launch_csv = sys.argv[1], sys.argv[2]
    out_dir = os.path.join(ROOT, "profiles")
    lines, traffic = [], {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full "
                                      "--clock-control none` captures summarised in %s_ncu_full_summary.txt (B=2, "
                                      "64x128x128 unless the name says otherwise); bench.py copies the entry of its "
                                      "dominant kernel's largest launch into roofline.traffic" % prefix}
    for arg in sys.argv[3:]:
        name, path = arg.split("=", 1)
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        kernels = rows[2:]
        names = name.split("|")                       # "a|b=file": successive kernels of the report are named a, b
        for ki, vals in enumerate(kernels):
            kname = vals[hdr.index('Kernel Name')]
            short = kname.split('(')[0].split('<')[0].replace('void ', '')
            if len(names) > 1:
                key = names[ki] if ki < len(names) else f"{names[-1]}:{short}"
            else:
                key = name if len(kernels) == 1 else f"{name}:{short}"
            if key in traffic:
                continue
            lines.append(f"== {key}   ({os.path.basename(path)})")
            lines.append(f"{'Kernel Name':74s} {kname[:110]}")
            tot = 0.0
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append(f"{w:74s} {vals[i]} {units[i]}")
                    if w.startswith("dram__bytes"):
                        tot += to_bytes(vals[i], units[i])
            traffic[key] = int(tot)
    open(os.path.join(out_dir, f"{prefix}_ncu_full_summary.txt"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(out_dir, f"{prefix}_traffic.json"), "w"), indent=1)
    # launch list
    with open(launch_csv) as f:
        body = [l for l in f if not l.startswith("==")]
    rd = csv.reader(body)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0.0, 0])
    for row in rd:
        if len(row) <= iv:
            continue
        try:
            v = float(row[iv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row[iu], 1.0)
        a = agg[row[ik].split("(")[0]]
        a[0] += v
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    n = sum(v[1] for v in agg.values())
    out = [f"# ncu launch list ({os.path.basename(launch_csv)}): {n} launches, {tot:.1f} us in total",
           "# per-launch times are cold-cache and serialised: compare SHARES with bench.py's roofline.by_kernel, not absolutes",
           f"{'total_us':>10s} {'launches':>8s} {'avg_us':>8s} {'share':>6s}  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append(f"{v[0]:10.1f} {v[1]:8d} {v[0] / v[1]:8.2f} {100 * v[0] / tot:5.1f}%  {k}")
    open(os.path.join(out_dir, f"{prefix}_ncu_launch_list_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
