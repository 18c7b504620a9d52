"""Turn gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep into small committed summaries under profiles/."""
import collections
import csv
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel share of the step ----
lines = [l for l in open(os.path.join(root, "gpurun_out", f"launches_{tag}.csv")) if l.startswith('"')]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0][:90]
    val = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    val = val / 1000 if u == "ns" else (val * 1000 if u == "ms" else val)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += val
    tot += val
with open(os.path.join(out_dir, f"launches_{tag}.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}): `python bench.py --steps 2 --warmup 3 --no_cpu_baseline`\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- cold-cache, serialised launches: compare SHARES, not absolutes.\n")
    f.write("Covers warm-up + timed graph replays + the eager stage-timing steps + the loss roofline probe + input generation.\n\n")
    f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |\n")
    f.write(f"\ntotal {tot:.0f} us over {sum(n for n, _ in agg.values())} launches\n")

# ---- full capture: key metrics per kernel ----
rep = os.path.join(root, "gpurun_out", f"prof_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
seen = collections.OrderedDict()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0][:80]
    seen.setdefault(name, r)  # first capture of each kernel
with open(os.path.join(out_dir, f"ncu_full_{tag}.md"), "w") as f:
    f.write(f"# ncu --set full summary ({tag})\n\n`ncu --set full --clock-control none --import-source on` on `python bench.py --steps 2 --warmup 3`; one launch per kernel shown.\n")
    f.write("`traffic` for bench.py's roofline = dram read + write below.\n\n")
    for name, r in seen.items():
        f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w, i in idx[1:]:
            f.write(f"| {w} | {r[i]} | {units[i]} |\n")
        f.write("\n")
# ---- machine-readable per-launch DRAM traffic for bench.py's roofline.traffic ----
import json
traffic = {}
ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
def _to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def _opt(r, h, name):
    return float(r[h.index(name)].replace(",", "")) if name in h and r[h.index(name)] not in ("", "n/a") else None
for name, r in seen.items():
    key = name.replace("void ", "").replace("csn::", "")
    traffic[key] = {"dram_read_bytes": _to_bytes(r[ir], units[ir]), "dram_write_bytes": _to_bytes(r[iw], units[iw]),
                    "duration_us_under_ncu": float(r[it].replace(",", "")) * ({"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[it], 1)),
                    "tensor_pipe_active_pct": _opt(r, hdr, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "tc_pipe_active_pct": _opt(r, hdr, "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active")}
# the cfg3-shape DINO loss kernel is captured separately (scripts/gpu_ncu_loss.sh -> gpurun_out/prof_loss.ncu-rep)
loss_rep = os.path.join(root, "gpurun_out", "prof_loss.ncu-rep")
if os.path.isfile(loss_rep):
    lraw = subprocess.run(["ncu", "-i", loss_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lrows = list(csv.reader(lraw.splitlines()))
    lh, lu = lrows[0], lrows[1]
    r = lrows[2]
    name = r[lh.index("Kernel Name")].split("(")[0].replace("void ", "").replace("csn::", "")
    traffic[name + " [cfg3 shape, scripts/loss_bench.py]"] = {
        "dram_read_bytes": _to_bytes(r[lh.index("dram__bytes_read.sum")], lu[lh.index("dram__bytes_read.sum")]),
        "dram_write_bytes": _to_bytes(r[lh.index("dram__bytes_write.sum")], lu[lh.index("dram__bytes_write.sum")]),
        "duration_us_under_ncu": float(r[lh.index("gpu__time_duration.sum")].replace(",", ""))}
    with open(os.path.join(out_dir, f"ncu_full_{tag}.md"), "a") as f:
        f.write(f"## `{name}` (cfg3 shape: 6 student + 2 teacher views, 64 trials, K = 65536; `scripts/gpu_ncu_loss.sh`)\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in want[1:] + ["launch__cluster_max_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]:
            if w in lh:
                f.write(f"| {w} | {r[lh.index(w)]} | {lu[lh.index(w)]} |\n")
        f.write("\n")

# ---- cfg4 (H = 512 cluster recurrence): its own launch list and full capture (scripts/gpu_ncu_cfg4.sh) ----
c4_csv = os.path.join(root, "gpurun_out", f"launches_cfg4_{tag}.csv")
if os.path.isfile(c4_csv):
    lines4 = [l for l in open(c4_csv, errors="ignore") if l.startswith('"')]
    agg4, tot4 = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines4):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0][:90]
        val = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        val = val / 1000 if u == "ns" else (val * 1000 if u == "ms" else val)
        a = agg4.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
        tot4 += val
    with open(os.path.join(out_dir, f"launches_cfg4_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list, cfg4 step ({tag}): `CSN_LSTM_NO_OVERLAP=1 NSTEPS=1 python scripts/bench_cfg4.py` (3 warm-up + 1 timed step, first 400 launches)\n\n")
        f.write("L2 / H512 / 128 trials.  ncu serialises kernels, so the GEMMs that normally run BESIDE the cluster recurrence (and feed / follow it\n"
                "through flags) are put back in sequence for the capture: the shares below are the sequential ones.  Cold-cache: compare SHARES.\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg4.items(), key=lambda kv: -kv[1][1])[:16]:
            f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot4:.1f}% |\n")
        f.write(f"\ntotal {tot4:.0f} us over {sum(n for n, _ in agg4.values())} launches\n")
c4_rep = os.path.join(root, "gpurun_out", f"prof_cfg4_{tag}.ncu-rep")
if os.path.isfile(c4_rep):
    craw = subprocess.run(["ncu", "-i", c4_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    crows = list(csv.reader(craw.splitlines()))
    ch, cu = crows[0], crows[1]
    cseen = collections.OrderedDict()
    for r in crows[2:]:
        cseen.setdefault(r[ch.index("Kernel Name")].split("(")[0][:80], r)
    with open(os.path.join(out_dir, f"ncu_full_{tag}.md"), "a") as f:
        for name, r in cseen.items():
            f.write(f"## `{name}` (cfg4: T = 440, 128 trials, H = 512, two trial groups per 16-CTA cluster; `scripts/gpu_ncu_cfg4.sh`)\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in want[1:] + ["launch__cluster_max_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]:
                if w in ch:
                    f.write(f"| {w} | {r[ch.index(w)]} | {cu[ch.index(w)]} |\n")
            f.write("\n")
            key = name.replace("void ", "").replace("csn::", "")
            traffic[key + " [cfg4, scripts/bench_cfg4.py]"] = {
                "dram_read_bytes": _to_bytes(r[ch.index("dram__bytes_read.sum")], cu[ch.index("dram__bytes_read.sum")]),
                "dram_write_bytes": _to_bytes(r[ch.index("dram__bytes_write.sum")], cu[ch.index("dram__bytes_write.sum")]),
                "duration_us_under_ncu": float(r[ch.index("gpu__time_duration.sum")].replace(",", "")) * ({"ns": 1e-3, "us": 1, "ms": 1e3}.get(cu[ch.index("gpu__time_duration.sum")], 1)),
                "tensor_pipe_active_pct": _opt(r, ch, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}

json.dump({"tag": tag, "source": f"ncu --set full, profiles/ncu_full_{tag}.md", "kernels": traffic},
          open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
print(open(os.path.join(out_dir, f"launches_{tag}.md")).read()[:2500])
print(open(os.path.join(out_dir, f"ncu_full_{tag}.md")).read()[:6000])
