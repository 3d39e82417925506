"""Copies the last gpurun results into profiles/ (tracked): bench lines, launch list of the last step, ncu summary."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02_v1"
for src, dst in (("bench.log", "%s_bench.json"), ("bench_ref.log", "%s_bench_reference.json"), ("configs.jsonl", "%s_configs_3_4.jsonl"),
                 ("ubench.log", "%s_ubench.log"), ("bench_2gpu.log", "%s_bench_2gpu.json"), ("widths.jsonl", "%s_line_widths.jsonl"),
                 ("sparse.jsonl", "%s_sparse.jsonl"), ("files.jsonl", "%s_files_to_kf.jsonl"), ("ncu_vl.txt", "%s_ncu_virtual_lines.txt"),
                 ("ncu_sparse.txt", "%s_ncu_sparse.txt"), ("launches_sparse.txt", "%s_launches_sparse.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst % tag))
rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hdr, out = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            out.append((d["Kernel Name"], float(d["Metric Value"].replace(",", "")), d["Metric Unit"]))
idx = max(i for i, o in enumerate(out) if "probe_line_width" in o[0])
step = out[idx:]
tot = sum(o[1] for o in step if "kf::" in o[0])
with open(os.path.join(P, "%s_launches_bench.txt" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline\n")
    f.write("# kernels of the LAST step (ns; cold-cache and serialised under ncu: shares matter, not absolutes)\n")
    for o in step:
        f.write("%-110s %12.0f %s\n" % (o[0][:110], o[1], o[2]))
    f.write("# share of count_fasta_lines_kernel among the library's kernels of the step: %.1f %%\n" %
            (100 * sum(o[1] for o in step if "count_fasta_lines_kernel" in o[0]) / tot))
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, "prof_ln_bench.ncu-rep"), "--traffic-json", "1000", "5000000", "7"],
                     capture_output=True, text=True).stdout
open(os.path.join(P, "%s_ncu_count_fasta_lines.txt" % tag), "w").write(
    "# ncu --set full --import-source on --clock-control none -k regex:count_fasta_lines_kernel -s 3 -c 1: python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline\n" + txt)
print(open(os.path.join(P, "%s_launches_bench.txt" % tag)).read())
print(txt)
