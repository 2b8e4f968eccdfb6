"""Turns the raw ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py r1 gpurun_out/launches_r1.csv gpurun_out/prof_r1_implicit.ncu-rep gpurun_out/prof_r1_explicit.ncu-rep
"""
import collections, csv, io, json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
try:
    COMMIT = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=HERE).stdout.strip()
except Exception:
    COMMIT = "unknown"
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]

out = ["# ncu summary, round %s" % tag, "",
       "Source: `ncu --set full --clock-control none --import-source on` on one B200 (profiles/run_profile.py; tet10 55^3x6 LinearElastic CSR",
       "assembly and hex27 128^3 NeoHookean explicit force), and the launch list of `bench.py` (`--metrics gpu__time_duration.sum`).",
       "Numbers under ncu are cold-cache and serialised: use them for shares and pipe utilisation, not as bench values.", ""]

# ---- launch list
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("void ", "")[:90]
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    tot[name] += v; cnt[name] += 1
unit_ns = max(tot.values()) > 1e6
with open(os.path.join(HERE, "%s_launches.csv" % tag), "w") as f:
    f.write("kernel,launches,total_ms,ms_per_launch\n")
    for n, v in tot.most_common():
        ms = v / 1e6 if unit_ns else v
        f.write('"%s",%d,%.4f,%.4f\n' % (n, cnt[n], ms, ms / cnt[n]))
ours = {n: v for n, v in tot.items() if n.startswith("fl::")}
s_our = sum(ours.values())
out += ["## Launch list of `bench.py --steps 2 --warmup 3 --explicit-steps 3` (all kernels: `%s_launches.csv`)" % tag, "",
        "| kernel | launches | ms/launch | share of our kernels |", "|---|---|---|---|"]
for n, v in sorted(ours.items(), key=lambda kv: -kv[1])[:10]:
    ms = v / 1e6 if unit_ns else v
    out.append("| `%s` | %d | %.3f | %.1f%% |" % (n, cnt[n], ms / cnt[n], 100 * v / s_our))
out.append("")

traffic = {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h, units = rr[0], rr[1]
    seen = set()
    for r in rr[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
        if name in seen:          # the drivers repeat every assembly a few times: one capture per kernel is summarised
            continue
        seen.add(name)
        out += ["## `%s`  (%s)" % (name, os.path.basename(rep)), "", "| metric | value | unit |", "|---|---|---|"]
        vals = {}
        # fp64 / tensor pipe counters: every column ncu collected for them (instruction counts by pipe and pipe-active cycles), so
        # that tensor-pipe utilisation is a counter and not arithmetic
        pipe = [c for c in h if ("dmma" in c or "pipe_fp64" in c or "pipe_tensor" in c) and c not in METRICS
                and c.endswith("avg.pct_of_peak_sustained_active") and "hmma" not in c and "imma" not in c]
        for m in METRICS + pipe:
            if m in h:
                vals[m] = r[h.index(m)]
                out.append("| %s | %s | %s |" % (m, r[h.index(m)], units[h.index(m)]))
        out.append("")
        def tobytes(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        try:
            rd = tobytes(vals["dram__bytes_read.sum"], units[h.index("dram__bytes_read.sum")])
            wr = tobytes(vals["dram__bytes_write.sum"], units[h.index("dram__bytes_write.sum")])
            key = name.split("<")[0].replace("fl::", "") + "_bytes_per_launch"
            traffic.setdefault(key, rd + wr)
        except Exception:
            pass
traffic["_source"] = "profiles/%s_summary.md: ncu --set full --clock-control none captures (%s) at commit %s; per launch, cold cache" % (
    tag, ", ".join(os.path.basename(r) for r in reps), COMMIT)
json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
open(os.path.join(HERE, "%s_summary.md" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
print(traffic)
