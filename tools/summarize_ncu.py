"""Turn ncu CSV exports into the markdown summaries kept under profiles/.

  launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv <cmd>
                python tools/summarize_ncu.py launches L.csv "<cmd>" > profiles/rNN_launches_summary.md
  full capture: ncu -i R.ncu-rep --page raw --csv > R.csv
                python tools/summarize_ncu.py full R.csv "<cmd>" > profiles/rNN_ncu_full_summary.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name).strip()
    name = name.replace("void ", "").replace("cfem::", "")
    return name


def launches(path, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    mu = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        if r[mu] in ("nsecond", "ns"):
            v /= 1e3
        elif r[mu] in ("msecond", "ms"):
            v *= 1e3
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"Command: `{cmd}` (run plain first, exit 0), then the same under `ncu --metrics gpu__time_duration.sum --clock-control none --csv`.")
    print("Per-launch times are cold-cache and serialised (programmatic dependent launch does not overlap under ncu): compare shares, not absolutes.\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    def share(pred):
        return 100 * sum(t for k, (n, t) in agg.items() if pred(k)) / tot
    print(f"\nGroups: persistent BiCGStab (k_bicg_persist) {share(lambda k: 'k_bicg_persist' in k):.1f}%; "
          f"Chebyshev mass sweeps (k_tile_t16<Ep16Cheb..>, k_cheb_stream) {share(lambda k: 'Cheb' in k or 'k_cheb' in k):.1f}%; "
          f"other SpMV (k_tile_t16<Ep16Spmv/BiT>, k_spmv_stream) {share(lambda k: ('k_tile_t16' in k and 'Cheb' not in k) or 'k_spmv_stream' in k):.1f}%; "
          f"assembly (k_tile_assemble) {share(lambda k: 'k_tile_assemble' in k):.1f}%; "
          f"epsilon/statistics (k_epsilon*, k_stats*) {share(lambda k: 'k_epsilon' in k or 'k_stats' in k):.1f}%.")


FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def full(path, cmd):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    cols = [(m, hdr.index(m)) for m in FULL if m in hdr]
    agg = OrderedDict()
    for r in data:
        a = agg.setdefault(short(r[kn]), [0, [0.0] * len(cols)])
        a[0] += 1
        for k, (m, i) in enumerate(cols):
            try:
                a[1][k] += float(r[i].replace(",", ""))
            except ValueError:
                pass
    print(f"Command: `{cmd}` (after the same command exited 0 without ncu).  Mean over the captured launches.\n")
    print("| kernel | n | " + " | ".join(f"{m} [{units[i]}]" for m, i in cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for k, (n, v) in agg.items():
        print(f"| `{k}` | {n} | " + " | ".join(f"{x / n:.2f}" for x in v) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
