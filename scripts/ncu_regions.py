"""Where the step kernel's time goes, from the SASS page of an `ncu --set full --import-source on` capture: instructions are classed by how often
they execute (once per launch = prologue/epilogue, once per substep = straight-line body, once per PGS sweep = sweep) and the warp-stall samples
are summed per class.   python scripts/ncu_regions.py gpurun_out/prof_step_r02.ncu-rep > profiles/step_kernel_r02_regions.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hdr_i], rows[hdr_i + 1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ex = []
for r in data:
    try:
        ex.append(int(r[ix["Instructions Executed"]]))
    except Exception:
        ex.append(0)
base = min(e for e in ex if e > 0)                       # executed once per warp
cls = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r, e in zip(data, ex):
    try:
        smp = int(r[ix["# Samples"]])
    except Exception:
        continue
    key = "pgs_sweep (x13 substeps x23 sweeps)" if e > 100 * base else ("substep_body (x13)" if e >= 10 * base else "prologue_epilogue (x1)")
    c = cls[key]; c[0] += 1; c[1] += smp
    for s in stalls:
        try:
            c[2][s] += int(r[ix[s]])
        except Exception:
            pass
    tot += smp
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "")
print("region, static instructions, stall samples, share of all samples, CPI estimate (samples with stall_selected = issued)")
for k, (n, s, st) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    sel = st.get("stall_selected", 0)
    print("%-42s %5d %6d  %5.1f%%  cycles/issue %.2f" % (k, n, s, 100.0 * s / tot, s / max(sel, 1)))
    print("      " + ", ".join("%s %.1f%%" % (a.replace("stall_", ""), 100.0 * b / s) for a, b in st.most_common(7)))
