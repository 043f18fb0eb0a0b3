"""Turn the raw artefacts a GPU run leaves in gpurun_out/ into the tracked summaries under profiles/.
usage: python scripts/summarise_profiles.py <round tag, e.g. r01>"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)


def launch_table(path, title, cmd):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
        name = r[kn].split("(")[0]
        name = name.replace("pb::<unnamed>::", "").replace("void ", "")[-70:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["# %s" % title, "", "`%s`" % cmd, "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, n, t, t / n, 100 * t / tot))
    return "\n".join(out) + "\n"


def launch_table_metric(path, title, cmd, metric="gpu__time_duration.sum"):
    """Like launch_table, for CSVs that carry several metrics per launch (keeps `metric`)."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mn] != metric:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] in ("ns", "nsecond") else (v * 1000 if r[mu] in ("ms", "msecond") else v)
        name = r[kn].split("(")[0].replace("pb::<unnamed>::", "").replace("void ", "")[-70:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["# %s" % title, "", "`%s`" % cmd, "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, n, t, t / n, 100 * t / tot))
    return "\n".join(out) + "\n"


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__cluster_size",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
            "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum.per_second",
            "lts__t_sector_hit_rate.pct", "smsp__warps_active.avg.per_cycle_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
    d = collections.OrderedDict()
    d["kernel"] = r[hdr.index("Kernel Name")][:120]
    for k in want:
        if k in hdr:
            d[k] = "%s %s" % (r[hdr.index(k)], units[hdr.index(k)])
    return d


if __name__ == "__main__":
    if os.path.exists(os.path.join(GO, "launches.csv")):
        open(os.path.join(PR, "%s_launches_bench.md" % tag), "w").write(launch_table(
            os.path.join(GO, "launches.csv"), "ncu launch list of the bench (scoring steps)",
            "ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 3 --warmup 3 --no-cpu --skip-em"))
    if os.path.exists(os.path.join(GO, "fit_launches.csv")):
        open(os.path.join(PR, "%s_launches_fit.md" % tag), "w").write(launch_table(
            os.path.join(GO, "fit_launches.csv"), "ncu launch list of PLDA.fit (C2: 100k x 200, 1k speakers, 5 EM iters, run twice)",
            "ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/fit_once.py 200 1000 100 5"))
    for name in ("prof_gemm", "prof_gemm_d512", "prof_prep"):
        rep = os.path.join(GO, name + ".ncu-rep")
        if os.path.exists(rep):
            d = ncu_raw(rep)
            with open(os.path.join(PR, "%s_ncu_%s.json" % (tag, name)), "w") as f:
                json.dump(d, f, indent=1)
    # round 2: every <tag>_prof_<kernel>.ncu-rep in gpurun_out/ -> profiles/<tag>_ncu_<kernel>.json
    for fn in sorted(os.listdir(GO)):
        if fn.startswith(tag + "_prof_") and fn.endswith(".ncu-rep"):
            d = ncu_raw(os.path.join(GO, fn))
            with open(os.path.join(PR, "%s_ncu_%s.json" % (tag, fn[len(tag) + 6:-8])), "w") as f:
                json.dump(d, f, indent=1)
    for fn, title, cmd in ((tag + "_fit_launches.csv", "ncu launch list of PLDA.fit (C2: 100k x 200, 1k speakers, 10 EM iters, run twice)",
                            "ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/fit_once.py 200 1000 100 10"),
                           (tag + "_stats_launches.csv", "ncu launch list of PLDA.fit (2M x 512 fp32, 20k speakers, 1 EM iter, run twice)",
                            "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv python scripts/r2_stats_probe.py 2000000 512 20000 1 f32"),
                           (tag + "_ragged_launches.csv", "ncu launch list of score_grid with ragged enrol counts (10k x 10k, d = 160, counts 1..5), then uniform counts",
                            "ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/r2_sink_probe.py ragged"),
                           (tag + "_bench_launches.csv", "ncu launch list of the bench (scoring steps)",
                            "ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 3 --warmup 3 --no-cpu --skip-em --headline-only")):
        path = os.path.join(GO, fn)
        if os.path.exists(path):
            open(os.path.join(PR, fn.replace(".csv", ".md").replace("_launches", "").replace(tag + "_", tag + "_launches_")), "w").write(
                launch_table_metric(path, title, cmd))
    for name in ("bench.json", "bench_ref.json"):
        p = os.path.join(GO, name)
        if os.path.exists(p) and os.path.getsize(p) > 10:
            open(os.path.join(PR, "%s_%s" % (tag, name)), "w").write(open(p).read())
    print("profiles written:", sorted(os.listdir(PR)))
