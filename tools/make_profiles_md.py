#!/usr/bin/env python
"""Turn the JSON / CSV artefacts a GPU run left in gpurun_out/ into the markdown tables kept under profiles/.
CPU only.   python tools/make_profiles_md.py"""
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def table(rows, cols):
    out = ["| " + " | ".join(c for c, _ in cols) + " |", "|" + "---|" * len(cols)]
    for r in rows:
        out.append("| " + " | ".join(str(f(r)) for _, f in cols) + " |")
    return "\n".join(out)


def stage_md(src, dst_json, dst_md, title, how):
    d = json.load(open(os.path.join(G, src)))
    shutil.copy(os.path.join(G, src), os.path.join(P, dst_json))
    key = [k for k in d["rows"][0] if k.endswith("_per_s")][0]
    cols = [("stage", lambda r: r["stage"]), ("ms", lambda r: r["ms"]), ("algorithmic MB", lambda r: r["algorithmic_MB"]),
            ("GB/s", lambda r: r["GBps"]), ("of measured peak", lambda r: r["frac_of_measured_peak"]), ("scans/s", lambda r: r[key]),
            ("note", lambda r: r.get("note", ""))]
    with open(os.path.join(P, dst_md), "w") as f:
        f.write(f"# {title}\n\n{how}\nAlgorithmic bytes per SURVEY.md 8d; peak = {d['peak_GBps']} GB/s (measured copy, MEASURED_PEAKS.json).\n\n")
        f.write(table(d["rows"], cols) + "\n")


def launches_md(src, dst_csv, dst_md, bench_json):
    rows = [r for r in csv.reader(open(os.path.join(G, src))) if len(r) > 10]
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        agg.setdefault(name, []).append(float(r[iv].replace(",", "")))
    shutil.copy(os.path.join(G, src), os.path.join(P, dst_csv))
    ours = {k: v for k, v in agg.items() if "slu::" in k}
    n_steps = min(len(v) for v in ours.values())
    step = sum(sum(v) / len(v) for v in ours.values())
    red = [k for k in ours if "reduce_staged" in k][0]
    b = json.load(open(os.path.join(P, bench_json)))
    with open(os.path.join(P, dst_md), "w") as f:
        f.write("# ncu launch list, `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e` (B200, round 1)\n\n"
                "Per-launch device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised:\n"
                f"compare SHARES, not absolutes).  Raw rows: profiles/{dst_csv}.\n\n| kernel | launches | mean ns |\n|---|---|---|\n")
        for k, v in agg.items():
            f.write(f"| `{k[:72]}` | {len(v)} | {sum(v) / len(v):.0f} |\n")
        f.write(f"\nOne step = {len(ours)} libslu launches = {step:.0f} ns under ncu; the fused reduction kernel is "
                f"{sum(ours[red]) / len(ours[red]):.0f} ns = {100 * sum(ours[red]) / len(ours[red]) / step:.1f}% of it.\n"
                f"bench.py, un-profiled CUDA events, same run configuration: kernel {b['roofline']['kernel_ms']} ms of {b['ms_per_step']} ms per step = "
                f"{b['roofline']['kernel_share_of_step']} (profiles/{bench_json}).\n"
                "The `at::` rows are torch.randn / fill kernels that build the synthetic inputs before the timed region.\n")


if __name__ == "__main__":
    shutil.copy(os.path.join(G, "bench_r01b_n1.json"), os.path.join(P, "bench_r01_n1.json"))
    shutil.copy(os.path.join(G, "bench_r01b_ref.json"), os.path.join(P, "bench_r01_reference_arm.json"))
    shutil.copy(os.path.join(G, "hist_bench_r01.json"), os.path.join(P, "hist_bench_r01.json"))
    stage_md("stages_r01c.json", "stages_r01.json", "stages_r01.md", "Per-stage timings through the Python wrappers, one B200, round 1 (`tools/stage_report.py`)",
             "CUDA events around one wrapper call, median of 20 after 3 warm-ups (includes the Python wrapper and, for the loss rows, autograd);\nL2 flushed between runs wherever the working set would fit in it.")
    stage_md("kernel_report_r01.json", "kernel_report_r01.json", "kernel_report_r01.md", "GPU-side stage timings, one B200, round 1 (`tools/kernel_report.py`)",
             "Each stage is captured into a CUDA graph once and the graph replayed between CUDA events (median of 20, L2 flushed before each replay):\nno Python or launch overhead inside the timed region.")
    launches_md("launches_r01c.csv", "launches_r01.csv", "launches_r01.md", "bench_r01_n1.json")
    print("ok")
