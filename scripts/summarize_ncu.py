"""Summarise ncu artefacts from gpurun_out/ into profiles/ (tracked).  Usage:
   python scripts/summarize_ncu.py launches <launches.csv> <out.txt>
   python scripts/summarize_ncu.py rep <file.ncu-rep> <out.txt>"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active.avg.pct", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct", "sm__pipe_fp64_cycles_active.avg.pct", "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "dram__cycles_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__sass_inst_executed_op_shared"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki], [0, 0.0, r[gi], r[bi]])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path}); cold-cache serialised launches: compare SHARES\n")
        f.write(f"# total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
        for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t / tot:7.3%}  n={n:4d}  total={t / 1e6:10.3f} ms  avg={t / n / 1e3:10.2f} us  grid={g} block={b}  {k[:110]}\n")


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({path})\n")
        for r in rows[2:]:
            f.write(f"\n== {r[ni][:120]}\n")
            for i, h in enumerate(hdr):
                if any(k in h for k in KEYS):
                    f.write(f"{h:95s} {r[i]:>18s} {units[i]}\n")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    body = [r for r in rows if len(r) > 5 and r[0].startswith("0x")]
    if body:
        tot = sum(int(r[2]) for r in body)
        agg = collections.Counter()
        for r in body:
            t = r[1].split()
            agg[t[1] if t[0].startswith("@") else t[0]] += int(r[2])
        with open(out, "a") as f:
            f.write(f"\n== warp-stall samples by SASS opcode (last kernel in the report; {tot} samples)\n")
            for k, v in agg.most_common(14):
                f.write(f"{k:24s} {v:9d} {v / tot:7.3%}\n")
            f.write("== SASS evidence: " + ", ".join(f"{k}={sum(1 for r in body if k in r[1])}" for k in
                    ("DMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "UTCIMMA", "UTCHMMA", "LDTM", "SYNCS")) + "\n")


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2], sys.argv[3])
