#!/usr/bin/env python3
"""Summarise gpurun_out/ artefacts (ncu launch lists, ncu --set full reports, bench JSON lines, microbench log)
into tracked files under profiles/.  Run in the build container after a gpurun call."""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
G = ROOT / "gpurun_out"
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"


def launch_table(csv_path: Path):
    rows = list(csv.reader(open(csv_path, errors="ignore")))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(d["Metric Unit"], 1)
            name = re.sub(r"\(.*", "", d["Kernel Name"])
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ns
    tot = sum(v[1] for v in agg.values()) or 1
    lines = ["| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:80]}` | {v[0]} | {v[1]/1e6:.2f} | {v[1]/tot*100:.1f}% | {v[1]/v[0]/1e3:.1f} |")
    return "\n".join(lines), tot / 1e6


def ncu_raw(rep: Path):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        item = {"kernel": re.sub(r"\(.*", "", d.get("Kernel Name", ""))}
        for k in keep:
            if k in d:
                item[k] = f"{d[k]} {units[hdr.index(k)]}"
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(v) > 0.1:
                        item["stall:" + k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = v
                except ValueError:
                    pass
        res.append(item)
    return res


def ncu_raw_csv(path: Path):
    """The same extract from a `ncu -i report --page raw --csv` export made on the GPU box (the reports themselves exceed the
    64 MiB that travels back).  Returns (items, total dram bytes over the launches)."""
    rows = list(csv.reader(open(path, errors="ignore")))
    if len(rows) < 3:
        return [], 0.0
    hdr, units = rows[0], rows[1]
    keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    res, total = [], 0.0
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        item = {"kernel": re.sub(r"\(.*", "", d.get("Kernel Name", ""))}
        for k in keep:
            if k in d:
                item[k] = f"{d[k]} {units[hdr.index(k)]}"
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in d:
                try:
                    total += float(d[k].replace(",", "")) * scale.get(units[hdr.index(k)], 1.0)
                except ValueError:
                    pass
        st = []
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(v) > 0.3:
                        st.append((float(v), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        item["stalls (cycles per issued instruction)"] = ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7])
        res.append(item)
    return res, total


def main():
    OUT.mkdir(exist_ok=True)
    md = [f"# profiles ({TAG}) - generated by tools/make_profiles.py from gpurun_out/", ""]
    for name in sorted(G.glob("launches_*.csv")):
        table, tot = launch_table(name)
        md += [f"## ncu launch list: `{name.name}`", "",
               "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares, not absolutes); "
               f"total {tot:.1f} ms", "", table, ""]
        (OUT / f"{TAG}_{name.name}").write_bytes(name.read_bytes()[:2_000_000])
        import gzip
        (OUT / f"{TAG}_{name.name}.gz").write_bytes(gzip.compress(name.read_bytes()))
        (OUT / f"{TAG}_{name.name}").unlink()
    traffic = {}
    for raw in sorted(G.glob(f"{TAG}_*_raw.csv")):
        items, total = ncu_raw_csv(raw)
        active = [it for it in items if "gpu__time_duration.sum" in it and not it["gpu__time_duration.sum"].startswith("0.00")]
        md += [f"## ncu --set full (raw page exported on the box): `{raw.name}`", "",
               f"{len(items)} launches captured ({len(active)} doing work, the rest are rounds that returned at once); "
               f"dram__bytes_read + dram__bytes_write summed over all of them: **{total / 1e9:.2f} GB**", ""]
        for item in active:
            md.append("```")
            md += [f"{k}: {v}" for k, v in item.items()]
            md.append("```")
        md.append("")
        traffic[raw.name] = total
        (OUT / raw.name).write_bytes(raw.read_bytes())
    if f"{TAG}_acc_g1_raw.csv" in traffic:
        (OUT / f"{TAG}_traffic.json").write_text(json.dumps({
            "accumulate_g1_2_21_bytes": traffic[f"{TAG}_acc_g1_raw.csv"],
            "accumulate_g2_2_20_bytes": traffic.get(f"{TAG}_acc_g2_raw.csv"),
            "source": f"profiles/{TAG}_acc_g1_raw.csv: ncu --set full --clock-control none of every k_bat_round<Fq> / k_bat_finish<Fq> launch of ONE "
                      "2^21-term G1 accumulation (tools/r2_evidence.sh), dram__bytes_read.sum + dram__bytes_write.sum summed over the launches"}, indent=1))
    for rep in sorted(G.glob("*.ncu-rep")):
        md += [f"## ncu --set full: `{rep.name}`", ""]
        for item in ncu_raw(rep):
            md.append("```")
            md += [f"{k}: {v}" for k, v in item.items()]
            md.append("```")
        md.append("")
    for b in sorted(G.glob(f"{TAG}_bench_*.json")):
        try:
            line = [l for l in b.read_text().splitlines() if l.startswith("{")][0]
            d = json.loads(line)
        except Exception:
            continue
        md += [f"## bench line: `{b.name}`", "", "```json", json.dumps({k: d[k] for k in d if k not in ("times_ms",)}, indent=1), "```", ""]
        (OUT / b.name).write_text(line + "\n")
    for extra in ("mb.log", "sweep_parity.json", f"{TAG}_pytest_gpu.log", f"{TAG}_ntt_once.log", f"{TAG}_ab_bingcd.log", f"{TAG}_g2ab.log",
                  f"{TAG}_step5.log", f"{TAG}_step7.log", f"{TAG}_step8.log", f"{TAG}_step9.log", f"{TAG}_mp2.log", f"{TAG}_mp8.log",
                  f"{TAG}_step11.log", f"{TAG}_step12.log", f"{TAG}_step13.log", f"{TAG}_pytest_mixed.log", f"{TAG}_smoke.log",
                  f"{TAG}_mp2c.log", f"{TAG}_mp4b.log", f"{TAG}_mp8b.log", f"{TAG}_mp_spdz_2.log", f"{TAG}_mp_spdz_4.log", f"{TAG}_mp_gsz_2.log",
                  f"{TAG}_mp_gsz_4.log", f"{TAG}_mp_spdz_8.log", f"{TAG}_mp_gsz_8.log", "mp_pytest_additive_2.log", "mp_pytest_spdz_2.log", "mp_pytest_gsz_2.log"):
        p = G / extra
        if p.exists():
            md += [f"## `{extra}`", "", "```", p.read_text()[:6000], "```", ""]
    (OUT / f"{TAG}_summary.md").write_text("\n".join(md))
    print("wrote", OUT / f"{TAG}_summary.md")


if __name__ == "__main__":
    main()
