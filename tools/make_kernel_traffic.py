"""Dev tool: profiles/kernel_traffic.json from an `ncu --set full` capture of one frame (bench.py reads it for
`roofline.traffic` and the FusedBounce compute roofline).  Per kernel family: the family's LONGEST captured launch.

  python tools/make_kernel_traffic.py <rep.ncu-rep> <workload> <samples of the frame> [out.json]
"""
import csv, io, json, re, subprocess, sys

rep, workload, samples = sys.argv[1], sys.argv[2], float(sys.argv[3])
out_path = sys.argv[4] if len(sys.argv) > 4 else "profiles/kernel_traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
units = dict(zip(hdr, rows[1]))


def fam(name):
    for key, f in (("FusedBounceT", "FusedBounce"), ("ShadeT", "Shade"), ("ShadowResolveT", "ShadowResolve"), ("k_mesh_prefilter", "k_mesh_prefilter"),
                   ("k_prefilter_bounds", "k_prefilter_bounds"), ("k_gate_write", "k_gate_write"), ("k_gate_flags", "k_gate_flags"),
                   ("ShadowGate", "k_gate_flags (ShadowGate)"), ("k_finalize", "Finalize"), ("k_path_warp", "PathTail"), ("Refine", "Refine"),
                   ("Verify1", "Verify1+Verify2"), ("Verify2", "Verify2")):
        if key in name:
            return f
    return None


def val(d, k, scale=1.0):
    try:
        v = float(d[ix[k]].replace(",", ""))
    except Exception:
        return None
    u = units.get(k, "")
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)
    return v * mult * scale


best = {}
for d in data:
    f = fam(d[ix["Kernel Name"]])
    if not f:
        continue
    ms = val(d, "gpu__time_duration.sum")
    if f in best and best[f]["launch_ms"] >= ms:
        continue
    best[f] = {
        "dram_bytes_per_launch": (val(d, "dram__bytes_read.sum") or 0) + (val(d, "dram__bytes_write.sum") or 0),
        "launch_ms": ms, "ncu_kernel": d[ix["Kernel Name"]][:110],
        "registers": val(d, "launch__registers_per_thread"),
        "warp_inst_per_launch": val(d, "smsp__inst_executed.sum"),
        "issue_active_frac": (val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0) / 100.0,
        "fp64_pipe_frac": (val(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or 0) / 100.0,
        "fma_pipe_frac": (val(d, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") or 0) / 100.0,
        "alu_pipe_frac": (val(d, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") or 0) / 100.0,
        "warps_active_frac": (val(d, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0) / 100.0,
        "threads_per_inst": val(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
    }
# float64 thread-instructions of the FIRST captured kernel (FusedBounce's bounce-0 launch) from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', src)
if len(blocks) > 1 and "FusedBounce" in blocks[1].split("\n", 1)[0]:
    t = list(csv.reader(io.StringIO(blocks[1].split("\n", 1)[1])))
    h = {k: i for i, k in enumerate(t[0])}
    f64 = 0
    for r in t[1:]:
        if len(r) <= h["Thread Instructions Executed"]:
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[h["Source"]])
        if m and m.group(1).split(".")[0] in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX", "MUFU"):
            if m.group(1).startswith("MUFU") and "64" not in m.group(1):
                continue
            f64 += int(r[h["Thread Instructions Executed"]] or 0)
    best["FusedBounce"]["fp64_thread_inst_per_launch"] = f64
    best["FusedBounce"]["samples_per_launch"] = samples
try:
    tab = json.load(open(out_path))
except Exception:
    tab = {}
tab[workload] = best
json.dump(tab, open(out_path, "w"), indent=1)
for k, v in best.items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "ncu_kernel"})
