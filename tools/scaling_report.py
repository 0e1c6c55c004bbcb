"""Builds profiles/r02_scaling.md and profiles/r02_kpp_scaling.md from the JSON lines of tools/gpu_scale_session.sh
(gpurun_out/r02s{1,2,4,8}_*.json).  Nothing is measured here; the numbers are the bench lines' own."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def load(tag, name):
    p = os.path.join(OUT, f"{tag}_{name}.json")
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        return None


def its(c):
    return c.get("krylov_its_per_step"), (c.get("mass_its_per_step") or c.get("mass_pcg_its_per_step"))


def table(name, base_override=None):
    rows = []
    for n in (1, 2, 4, 8):
        d = load(f"r02s{n}", name)
        if d is None and n == 1 and base_override:
            d = base_override
        if d:
            rows.append((n, d))
    return rows


def main():
    base_b = load("r02h", "n1") or load("r02i", "default")
    out = []
    out.append("# Round 2 - multi-GPU scaling of the default workload (Burgers RV P1, 1024 x 1024 cells = 1,050,625 dofs PER GPU)\n")
    out.append("Command per point: `tools/gpu_scale_session.sh N r02sN burgers` = the driver's own launch "
               "(`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N --steps 20 --warmup 3`), one box per N; "
               "N = 1 from `bench.py --sweep` on a one-GPU box.  Hilbert-range partition, lin_rtol = mass_rtol = 1e-11 "
               "(row-equilibrated), persistent BiCGStab + Chebyshev chain, low-latency halo words.\n")
    out.append("| GPUs | ms / step | G DoF-updates/s | weak efficiency | Krylov / mass its per step | in-run parity (rel L2 vs oracle) |")
    out.append("|---:|---:|---:|---:|---|---|")
    rows = table("burgers", base_b)
    t1 = rows[0][1]["ms_per_step"] if rows and rows[0][0] == 1 else None
    for n, d in rows:
        c = d["config"]
        par = d.get("parity_rel_l2") or {}
        ptxt = ", ".join(f"{k.split('_')[0]} {v:.1e}" for k, v in par.items() if isinstance(v, float)) or "-"
        eff = f"{t1 / d['ms_per_step']:.3f}" if t1 else "-"
        k, m = its(c)
        out.append(f"| {n} | {d['ms_per_step']:.3f} | {d['value'] / 1e9:.3f} | {eff} | {k:.1f} / {m:.1f} | {ptxt} |")
    out.append("\n## Where a step's extra time goes (device-side accounting of the UN-profiled timed run, rank 0)\n")
    out.append("`cfem_comm_timers`: SM-clock stamps taken inside the kernels around (a) every poll of a low-latency halo word "
               "by a boundary tile, (b) the cross-rank part of every in-kernel all-reduce (last CTA: push tagged words to all "
               "ranks, wait for theirs), (c) every grid barrier of the persistent BiCGStab as seen by its first worker CTA "
               "(waiting for the slowest CTA of the phase + the reduction + (b)).  No profiler, no per-launch events: "
               "the ranks run exactly as in the timed region.\n")
    out.append("| GPUs | step - step(1 GPU) [ms] | all-reduces / step | cross-rank all-reduce wait [ms / step] (mean / max us each) | halo-word polls: mean wait [us] (max) | solver barriers / step | barrier time of worker 0 [ms / step] |")
    out.append("|---:|---:|---:|---|---|---:|---:|")
    for n, d in rows:
        w = d["config"].get("comm_wait")
        if not w:
            continue
        out.append(f"| {n} | {d['ms_per_step'] - t1:.3f} | {w['allreduces_per_step']:.1f} | {w['allreduce_us_per_step'] / 1e3:.3f} "
                   f"({w['allreduce_us_per_step'] / max(w['allreduces_per_step'], 1):.1f} / {w['allreduce_us_max']:.0f}) | "
                   f"{w['halo_wait_us_per_cta_wait']:.2f} ({w['halo_wait_us_max']:.1f}) | {w['solver_barriers_per_step']:.1f} | "
                   f"{w['solver_barrier_us_per_step_worker0'] / 1e3:.3f} |")
    out.append("\nReading: halo waits are gone (a boundary tile finds its ghost words present; what is left is the load latency). "
               "The cross-rank all-reduces cost 3 us each at two ranks and 6 us at eight, three per BiCGStab iteration; the "
               "rest of the gap is the distributed kernel variants themselves (tile_order indirection, one CTA given to the "
               "halo push, ghost checks: +0.13 ms at one GPU with `CFEM_FORCE_GHOST=1`), the per-step exchanges that are still "
               "separate launches (four halo exchanges of state fields, the min/max/sum statistics), and rank skew at the "
               "~50 global synchronisation points of a 2.5 ms step.\n")
    open(os.path.join(ROOT, "profiles", "r02_scaling.md"), "w").write("\n".join(out) + "\n")

    out = []
    out.append("# Round 2 - KPP rotating wave, scaling sweep on 32 M - 128 M cells (BASELINE.json configs[4])\n")
    out.append("Mesh: jittered + randomly renumbered triangulation (SURVEY section 8d variant B, `cfem_b200.meshes.jittered`), "
               "`dt = 0.64 h` (the reference's ratio, KPP_exact.py:38,75), Cvel 0.5, CRV 4, BDF2 residual, Newton rtol 1e-4, "
               "lin_rtol = mass_rtol = 1e-11 on the row-equilibrated residual; 10 timed steps after 3 warm-up steps from the "
               "initial condition, device-timed (max over ranks).  `tools/gpu_scale_session.sh N r02sN kweak kstrong` "
               "(`bench.py --workload kpp --sweep`), one box per N.  Hilbert-range partition.\n")
    for title, name, kind in (("Weak scaling: 2828^2 x 2 = 16.0 M cells (8.0 M dofs) per GPU", "kpp_weak", "weak"),
                              ("Strong scaling: 4000^2 x 2 = 32.0 M cells (16.0 M dofs) in total", "kpp_strong", "strong")):
        out.append(f"## {title}\n")
        out.append("| GPUs | cells | ms / step | M DoF-updates/s | efficiency | Krylov its / step | ms per Krylov iteration-step | efficiency per iteration | context set-up (rank 0) [s] | ghosts (rank 0) |")
        out.append("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
        rows = table(name)
        if not rows:
            continue
        t1 = rows[0][1]["ms_per_step"]
        k1 = rows[0][1]["config"]["krylov_its_per_step"]
        for n, d in rows:
            c = d["config"]
            eff = (t1 / d["ms_per_step"]) if kind == "weak" else (t1 / d["ms_per_step"] / n)
            per = d["ms_per_step"] / c["krylov_its_per_step"]
            effi = (t1 / k1) / per if kind == "weak" else (t1 / k1) / per / n
            out.append(f"| {n} | {c['cells']:,} | {d['ms_per_step']:.2f} | {d['value'] / 1e6:.1f} | {eff:.3f} | {c['krylov_its_per_step']:.1f} | "
                       f"{per:.3f} | {effi:.3f} | {c['context_setup_s_rank0']:.1f} | {c.get('n_ghosts_rank0', 0):,} |")
        out.append("")
    out.append("`efficiency` is the plain ratio of step times.  The Krylov iteration count grows with the global mesh in the weak "
               "sweep (49 -> 53 per step: Jacobi preconditioning on a larger domain at the same CFL), which is an algorithmic "
               "cost, not a parallel one; `efficiency per iteration` divides it out (step time / Krylov iterations).\n")
    out.append("Set-up: every rank orders the global nodes (one sort) and flags the boundary with one pass over the global cells, "
               "then builds adjacency, CSR pattern, tiles and halo lists for ITS part only (`csrc/setup.cpp`, stages A / B); "
               "the column above is `Context(...)` + `nodal_h()` on rank 0 and includes the NCCL / CUDA-IPC bring-up.\n")
    open(os.path.join(ROOT, "profiles", "r02_kpp_scaling.md"), "w").write("\n".join(out) + "\n")
    print(open(os.path.join(ROOT, "profiles", "r02_scaling.md")).read())
    print(open(os.path.join(ROOT, "profiles", "r02_kpp_scaling.md")).read())


if __name__ == "__main__":
    main()
