"""Top stall sites of one kernel from an ncu report captured with --import-source on.
usage: python tools/ncu_hot_sass.py REPORT.ncu-rep KERNEL_REGEX [N [SKIP]]  -> markdown table (SASS view) of the
(SKIP+1)-th captured launch whose base name matches"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
name = next(csv.reader([lines[0]]))[1]
end = next((k for k in range(1, len(lines)) if lines[k].startswith('"Kernel Name"')), len(lines))   # first launch only
rows = list(csv.reader(io.StringIO("\n".join(lines[1:end]))))
hdr, data = rows[0], [r for r in rows[1:] if r and r[0].startswith("0x")]
i_src, i_smp = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
i_exc = hdr.index("L2 Theoretical Sectors Global Excessive") if "L2 Theoretical Sectors Global Excessive" in hdr else None
i_conf = hdr.index("L1 Conflicts Shared N-Way") if "L1 Conflicts Shared N-Way" in hdr else None
tot = sum(int(r[i_smp] or 0) for r in data)
print(f"`{name[:110]}` — {tot} stall samples; where they wait (the instruction shown is the one that cannot issue, i.e. the consumer of the slow result):\n")
print("| share | samples | SASS | L2 excess sectors | smem N-way |\n|---:|---:|---|---:|---:|")
for r in sorted(data, key=lambda r: -int(r[i_smp] or 0))[:top]:
    s = int(r[i_smp] or 0)
    print(f"| {100 * s / tot:.1f}% | {s} | `{' '.join(r[i_src].split())}` | {r[i_exc] if i_exc is not None else ''} | {r[i_conf] if i_conf is not None else ''} |")
