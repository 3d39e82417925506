"""Print the metrics we track from an .ncu-rep (run where ncu is installed): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("KERNEL", d["Kernel Name"][:90], "grid", d["Grid Size"], "block", d["Block Size"])
    for k in KEYS:
        if k in d and d[k] != "":
            print("  %-78s %s" % (k, d[k]))
    st = [(k, float(d[k])) for k in hdr if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and d[k] not in ("", "n/a")]
    tot = sum(v for _, v in st) or 1
    print("  -- warp-state samples (%d total)" % tot)
    for k, v in sorted(st, key=lambda x: -x[1])[:10]:
        print("     %-40s %6.1f %%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot))
    ops = [(k, float(d[k])) for k in hdr if k.startswith("sass__inst_executed_per_opcode") or k.startswith("smsp__sass_inst_executed_op_")]

if len(sys.argv) >= 6 and sys.argv[2] == "--traffic-json":
    # python tools/ncu_summary.py rep --traffic-json GENOMES BASES K  -> profiles/traffic.json for bench.py's roofline.traffic
    import json, os
    G, NB, K = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    best = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "count_fasta_lines_kernel<" in d["Kernel Name"]:
            units = dict(zip(hdr, rows[1]))
            def to_bytes(key):
                v, u = float(d[key]), units[key]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
            best = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    out = {"genomes": G, "bases": NB, "k": K, "dram_bytes_per_launch": best,
           "source": "ncu --set full --clock-control none, count_fasta_lines_kernel, dram__bytes_read.sum + dram__bytes_write.sum (%s)" % os.path.basename(rep)}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
    print(out)
