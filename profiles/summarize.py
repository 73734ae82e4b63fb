"""Turns ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/r4_launches.csv            > profiles/r1_launches_bench.txt
    python profiles/summarize.py kernel   gpurun_out/r4_prof_fwd.ncu-rep [idx]  > profiles/r1_fwd_c2_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[row["Metric Unit"]]
        agg[row["Kernel Name"].split("(")[0][:90]][0] += 1
        agg[row["Kernel Name"].split("(")[0][:90]][1] += v
        tot += v
    print(f"# ncu --metrics gpu__time_duration.sum launch list: {path}  (cold-cache, serialised: compare SHARES)")
    print(f"# total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:12.1f} us {n:5d}x avg {t / n:10.2f} us {100 * t / tot:6.2f}%  {k}")


def kernel(path, idx=0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    r = data[idx]
    name = r[hdr.index("Kernel Name")]
    print(f"# ncu --set full --clock-control none: {path}, launch #{idx}: {name}")
    for k in KEYS:
        for i, h in enumerate(hdr):
            if h == k:
                print(f"{k:80s} {r[i]:>18s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    heads = [i for i, rr in enumerate(rows) if rr and rr[0] == "Address"]
    if not heads:
        return
    h = rows[heads[min(idx, len(heads) - 1)]]
    end = heads[idx + 1] - 1 if idx + 1 < len(heads) else len(rows)
    body = rows[heads[min(idx, len(heads) - 1)] + 1:end]
    ci = {n: i for i, n in enumerate(h)}
    S = ci["# Samples"]
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(rr[S] or 0) for rr in body) or 1
    print(f"\n# warp-state samples: {tot} over {len(body)} SASS instructions")
    for n, v in sorted(((n, sum(int(rr[ci[n]] or 0) for rr in body)) for n in stalls), key=lambda kv: -kv[1])[:8]:
        print(f"{n:28s} {v:8d} {100 * v / tot:5.1f}%")
    print("# top stalled instructions (samples, SASS, dominant stall)")
    for rr in sorted(body, key=lambda rr: -int(rr[S] or 0))[:16]:
        st = max(((int(rr[ci[n]] or 0), n) for n in stalls))
        print(f"{int(rr[S]):7d}  {rr[ci['Source']][:72]:72s} {st[1]}")
    ops = collections.Counter(rr[ci["Source"]].replace("@P0", "").replace("@!P0", "").split()[0].split(".")[0] for rr in body if rr[ci["Source"]].strip())
    native = {k: v for k, v in ops.items() if k.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "SYNCS", "HMMA", "MUFU"))}
    print("# Blackwell-native SASS mnemonics in this kernel:", dict(sorted(native.items())))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
