"""Condense ncu exports into the text summaries committed under profiles/.
  python tools/ncu_summary.py launches <launches.csv>            -> compact per-launch table + per-kernel share
  python tools/ncu_summary.py full <report.ncu-rep>              -> key metrics per kernel + hot instructions"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def short(name):
    m = re.search(r"(attn_\w+|temporal_attn_kernel|\w+_kernel)\s*<([^>]*)>", name)
    if m:
        return f"{m.group(1)}<{m.group(2)}>"
    m = re.search(r"(\w+)\s*\(", name)
    return (m.group(1) if m else name)[:60]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    tot = defaultdict(lambda: [0, 0.0])
    print("id,kernel,grid,block,duration_us")
    for r in rows:
        ns = float(r[-1])
        k = short(r[4])
        tot[k][0] += 1
        tot[k][1] += ns
        print(f"{r[0]},{k},{r[8].replace(',', 'x').replace(' ', '')},{r[7].replace(',', 'x').replace(' ', '')},{ns / 1e3:.1f}")
    total = sum(v[1] for v in tot.values())
    print("\n# share of profiled device time by kernel (cold-cache, serialised launches: compare shares, not absolutes)")
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"# {100 * ns / total:6.2f}%  {n:4d} launches  avg {ns / n / 1e3:10.1f} us  {k}")


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpc__cycles_elapsed.max", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"]


def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        print(f"== {name}")
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = (r[i], units[i])
                print(f"   {k} = {r[i]} {units[i]}")
        try:  # achieved DRAM bandwidth of this launch (under ncu: cold caches, serialised — for the bytes, not the time)
            to_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            to_s = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
            nbytes = sum(float(vals[k][0].replace(",", "")) * to_b[vals[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            secs = float(vals["gpu__time_duration.sum"][0].replace(",", "")) * to_s[vals["gpu__time_duration.sum"][1]]
            print(f"   => DRAM read+write {nbytes / 1e6:.1f} MB in {secs * 1e6:.1f} us = {nbytes / secs / 1e9:.0f} GB/s")
        except Exception:  # noqa: BLE001
            pass
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + name.split("<")[0]],
                             capture_output=True, text=True).stdout
        with open("/tmp/_ncu_src.csv", "w") as fh:
            fh.write(src)
        hot = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_hot.py"), "/tmp/_ncu_src.csv", "12"],
                             capture_output=True, text=True).stdout
        print("   -- warp-stall sampling (top instructions) --")
        for line in hot.splitlines():
            print("   " + line)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
