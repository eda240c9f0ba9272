"""Turn the ncu outputs of tools/gpu_ncu.sh (brought back under gpurun_out/) into the committed evidence:
profiles/r02_ncu_summary.md (launch list shares + per-launch figures of the full captures) and profiles/r02_traffic.json
(dram bytes per launch of the dominant kernel, read by bench.py's `roofline.traffic`).

    python tools/ncu_summary.py          # here, no GPU: ncu -i ... --page raw --csv does the reading
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
PROF = os.path.join(REPO, "profiles")


def family(name):
    m = re.search(r"gemm_tc_kernel<(\d+), (\d), (\d), (\d)>", name)
    if m:
        mode = {"0": "row-major", "1": "3x3 conv", "2": "7x7 stem"}[m.group(3)]
        return f"gemm_tc<{m.group(1)}{', LN' if m.group(2) == '1' else ''}, {mode}{', DLN' if m.group(4) == '1' else ''}>"
    for key in ("attention_tc_kernel", "layernorm256_twice", "layernorm256", "ln_partials", "stem_canvas", "maxpool", "query_encode", "gemm_simt",
                "attention_simt", "f32_to_split16", "split16_to_f32", "exchange_push", "exchange_wait"):
        if key in name:
            return key
    return re.sub(r"\(.*", "", name.split("::")[-1])[:40]


def launches():
    path = os.path.join(OUT, "r02_ncu_launches_bench_n1.csv")
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3))
    # the graph-replayed steady state: drop everything before the last model forward pattern repeats - simply use the
    # kernels of this library (the flush / copy kernels of torch are listed separately)
    mine = [(family(n), us) for n, us in rows if "cotr::" in n or "unnamed>::" in n]
    other = sum(us for n, us in rows if not ("cotr::" in n or "unnamed>::" in n))
    agg = collections.OrderedDict()
    for fam, us in mine:
        a = agg.setdefault(fam, [0, 0.0])
        a[0] += 1; a[1] += us
    total = sum(v[1] for v in agg.values())
    return agg, total, other, len(rows)


def raw(rep):
    out = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        def g(name, scale=None):
            v = r[col[name]].replace(",", "")
            u = units[col[name]]
            x = float(v) if v not in ("", "no data") else float("nan")
            if scale == "bytes":
                x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            if scale == "us":
                x *= {"ns": 1e-3, "us": 1, "ms": 1e3}[u]
            return x
        res.append({
            "kernel": family(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
            "us": g("gpu__time_duration.sum", "us"),
            "dram_read": g("dram__bytes_read.sum", "bytes"), "dram_write": g("dram__bytes_write.sum", "bytes"),
            "tensor_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "tensor_pct_elapsed": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "dram_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "regs": g("launch__registers_per_thread"),
        })
    return res


def main():
    agg, total, other, n = launches()
    gemm = raw("r02_ncu_gemm.ncu-rep")
    attn = raw("r02_ncu_attn.ncu-rep")
    md = ["# Round 2 - ncu evidence (B200, one GPU, `tools/gpu_ncu.sh`; summarised by `tools/ncu_summary.py`)", "",
          "Commands: `ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv python bench.py --quick --steps 2 --warmup 3`",
          "(launch list, `profiles/r02_ncu_launches_bench_n1.csv`; graph kernel nodes are profiled one by one: cold-cache and",
          "serialised, so compare SHARES, not absolutes) and `ncu --set full --clock-control none --import-source on -k regex:<kernel>`",
          "for the two dominant kernels.  `bench.py --quick` = the headline steps only.", "",
          f"## Launch list: {n} profiled launches, {total / 1e3:.2f} ms in this library's kernels (+ {other / 1e3:.2f} ms torch fill / copy kernels)", "",
          "| kernel | launches | us (ncu) | share |", "|---|---|---|---|"]
    for fam, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| {fam} | {cnt} | {us:.0f} | {100 * us / total:.1f}% |")
    gemm_share = sum(us for f, (c, us) in agg.items() if f.startswith("gemm_tc")) / total
    md += ["", f"All `gemm_tc` instantiations together: {100 * gemm_share:.1f}% of the kernel time (bench.py's event-timed share: see the bench line).", "",
           "## Full captures", "",
           "| launch | grid | duration | tensor pipe active (while SM active / of elapsed) | DRAM read / write | DRAM throughput | regs |", "|---|---|---|---|---|---|---|"]
    for r in gemm + attn:
        md.append(f"| {r['kernel']} | {r['grid']} x {r['block']} | {r['us']:.1f} us | {r['tensor_pct']:.1f}% / {r['tensor_pct_elapsed']:.1f}% | {r['dram_read'] / 1e6:.2f} MB / {r['dram_write'] / 1e6:.2f} MB | {r['dram_pct']:.1f}% | {r['regs']:.0f} |")
    per_launch = [r["dram_read"] + r["dram_write"] for r in gemm]
    traffic = {"gemm_tc": {"dram_bytes_per_launch": sum(per_launch) / len(per_launch), "launches": len(per_launch),
                           "per_launch": per_launch},
               "attention_tc": {"dram_bytes_per_launch": sum(r["dram_read"] + r["dram_write"] for r in attn) / max(1, len(attn))},
               "source": "profiles/r02_ncu_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches)"}
    os.makedirs(PROF, exist_ok=True)
    with open(os.path.join(PROF, "r02_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    with open(os.path.join(PROF, "r02_ncu_summary.md"), "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
