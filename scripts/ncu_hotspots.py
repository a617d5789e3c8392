"""Source-level summary of an `ncu --set full --import-source on` report (no GPU needed): executed-instruction mix by
opcode, share of the warp-stall samples per opcode, the most sampled instructions and the stall reasons of one kernel.

    python scripts/ncu_hotspots.py gpurun_out/prof_X.ncu-rep KERNEL_REGEX > profiles/r01_X_hotspots.txt"""
import collections
import csv
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if not heads:
    sys.exit("no source page for %s in %s" % (regex, rep))
name = next((r[1] for r in rows[:heads[0]] if r and r[0] == "Kernel Name"), regex)
h = rows[heads[0]]
end = heads[1] - 1 if len(heads) > 1 else len(rows)
data = [r for r in rows[heads[0] + 1:end] if len(r) > 10 and r[0].startswith("0x")]
i_src, i_exec, i_smp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot = sum(int(r[i_exec]) for r in data)
smp = sum(int(r[i_smp]) for r in data)
print("report   %s" % rep)
print("kernel   %s" % name[:120])
print("SASS instructions %d, executed warp instructions %d, stall samples %d" % (len(data), tot, smp))
ex, st = collections.Counter(), collections.Counter()
for r in data:
    w = r[i_src].strip().split()
    op = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
    ex[op] += int(r[i_exec])
    st[op] += int(r[i_smp])
print("\nopcode          executed   stall samples")
for op, c in ex.most_common(16):
    print("%-14s %7.1f %%  %7.1f %%" % (op, 100.0 * c / tot, 100.0 * st[op] / smp))
print("\nmost sampled instructions (share of stall samples, times executed)")
for r in sorted(data, key=lambda r: -int(r[i_smp]))[:12]:
    print("%5.1f %%  %12s  %s" % (100.0 * int(r[i_smp]) / smp, r[i_exec], r[i_src].strip()[:90]))
cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.Counter()
for r in data:
    for c in cols:
        try:
            agg[c] += int(r[h.index(c)])
        except ValueError:
            pass
t = sum(agg.values()) or 1
print("\nstall reasons: " + ", ".join("%s %.1f %%" % (k[6:], 100.0 * v / t) for k, v in agg.most_common(8)))
