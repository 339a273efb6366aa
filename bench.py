#!/usr/bin/env python
"""Benchmark of the Kiwi source-inversion hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5|small] [--batch B]
    python bench.py --impl reference ...      # the CPU restatement of the reference on the host cores

A "step" is one batched evaluation (kiwi_eval_sources) of B candidate bilateral sources: source
discretisation -> synthesis at all receivers -> scaling -> misfits.  Metric: source evaluations per
second, whole job over all ranks.  One JSON line is printed by rank 0.

Default workload: c5, the configuration BASELINE.json's metric ("... @1/2/4/8 B200; HBM GB/s") is quoted on -- the dense-array
sweep of bilateral candidates over 2000 receivers, sharded over the GPUs (configs[4]); it is also the largest configuration
that fits one GPU.  c3 is the same source and sweep on 200 receivers, c2 the tensor-core moment-tensor grid search, c4 the
eikonal / amplitude-spectrum case.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "source evals/sec (synth+misfit, all receivers)"
UNIT = "evals/s"

WORKLOADS = {
    # Izmit bilateral rupture, ~1e4 sub-sources, 200 x 3 components,
    # bench-L analytical full-space database (HBM resident, > L2), time-domain L2 misfit
    "c3": dict(db="bench-L", nx=2000, nz=150, dx=100.0, dz=200.0, nrcv=200, effective_dt=0.35, dmin=45e3, dmax=150e3,
               norm="l2norm", batch=32, cpu_sample=2),
    # point moment-tensor grid search, (north, east, depth) lattice x unit
    # tensors on a Fibonacci sphere, rise time 1 s, effective_dt 0.5 -> 3 centroids per candidate
    "c2": dict(db="bench-L", nx=2000, nz=150, dx=100.0, dz=200.0, nrcv=100, effective_dt=0.5, dmin=45e3, dmax=150e3,
               norm="l2norm", batch=100000, cpu_sample=2000, source="moment_tensor"),
    # eikonal rupture-front source (rise time -> fold kernel), cosine taper, band-pass filter,
    # amplitude-spectrum L1 misfit (shared-memory FFT kernel); bord radius 5-12 km so that every centroid stays inside bench-L.
    # 4096 candidates per step (round 1: 32): the fast-marching solves are sequential per candidate and differ 6-fold in size; the
    # device solves them one warp each, 26 to an SM (3848 at a time), while the host threads solve the largest grids: the wave lasts
    # as long as its largest grid, so the rate grows with the batch (137 evals/s at 32, 490 at 1024, 1040 at 4096)
    "c4": dict(db="bench-L", nx=2000, nz=150, dx=100.0, dz=200.0, nrcv=200, effective_dt=0.5, dmin=45e3, dmax=150e3,
               norm="ampspec_l1norm", batch=4096, cpu_sample=1, source="eikonal",
               taper=([2.0, 6.0, 70.0, 80.0], [0, 1, 1, 0]), filter=([0.01, 0.02, 0.1, 0.2], [0, 1, 1, 0])),
    # the dense-array sweep -- the C3 source and candidate grid on 2000 receivers
    "c5": dict(db="bench-L", nx=2000, nz=150, dx=100.0, dz=200.0, nrcv=2000, effective_dt=0.35, dmin=45e3, dmax=150e3,
               norm="l2norm", batch=8, cpu_sample=1),
    # quick functional run (kiwibench-size pieces)
    "small": dict(db="bench-L/8", nx=1000, nz=60, dx=100.0, dz=400.0, nrcv=24, effective_dt=0.5, dmin=45e3, dmax=55e3,
                  norm="l2norm", batch=8, cpu_sample=2),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # samples that arrived inside the timed region; a short region is padded with the samples of the
        # warm-up steps right before it (same kernels, same load) so that there are at least three
        inside = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= (t1 or t) + 0.15)]
        before = [r for (t, r) in self.rows if t0 is not None and t < t0]
        padded = len(inside) < 3
        rows = (before[-(3 - len(inside)):] if padded else []) + inside
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(inside), "padded_with_warmup_samples": bool(padded)}


def make_db(w, rank, world, barrier):
    """rank 0 builds the analytical database once and shares it through a KGF1 file."""
    from kiwi_b200 import Gfdb, synthetic
    if world == 1:
        return synthetic.bench_l_db(w["nx"], w["nz"], w["dx"], w["dz"])
    path = "/tmp/kiwi_bench_%d_%d_%d.kgf1" % (w["nx"], w["nz"], os.getppid())
    if rank == 0:
        db = synthetic.bench_l_db(w["nx"], w["nz"], w["dx"], w["dz"])
        db.write(path)
    barrier()
    if rank != 0:
        db = Gfdb.read(path)
    barrier()
    if rank == 0:
        try:
            os.remove(path)
        except OSError:
            pass
    return db


def configure(eng, db, w, rlat, rlon, rdep):
    from kiwi_b200 import synthetic  # noqa: F401
    eng.set_database(db)
    eng.set_local_interpolation("bilinear")
    eng.set_receivers(rlat, rlon, rdep, ["ned"] * len(rlat))
    eng.set_source_location(30.0, 70.0, 0.0)
    eng.set_effective_dt(w["effective_dt"])
    eng.set_misfit_method(w["norm"])
    if w.get("taper"):
        for ir in range(1, len(rlat) + 1):
            eng.set_misfit_taper(ir, *w["taper"])
    if w.get("filter"):
        eng.set_misfit_filter(*w["filter"])


def set_references(src, engines, nrcv, dt, scale=1.07):
    """reference = synthetics of the base source with +7 % moment."""
    # (all traces are fetched before the first one is set: setting a reference on `src` itself invalidates its evaluation)
    REFS.clear()
    for ir in range(1, nrcv + 1):
        for ic in range(1, 4):
            first, data = src.get_seismogram(ir, ic, 1)
            REFS[(ir, ic)] = (first, data * np.float32(scale))
    for (ir, ic), (first, data) in REFS.items():
        for e in engines:
            e.set_ref_seismogram(ir, ic, (first - 1) * dt, data)


REFS = {}


def copy_references(src_unused, dst, nrcv, dt):
    """give `dst` exactly the reference traces that set_references() handed out last"""
    for (ir, ic), (first, data) in REFS.items():
        dst.set_ref_seismogram(ir, ic, (first - 1) * dt, data)


def run_reference_arm(args, w, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The Fortran
    cannot be built here (no Fortran compiler, HDF5, FFTW: SURVEY.md 8c), so this is the line-by-line
    C++ restatement (oracle/), OpenMP over receivers exactly where the reference has its only
    OpenMP loop (minimizer_engine.f90:893-903)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import OracleEngine, lib as olib
    from kiwi_b200 import synthetic
    ncores = os.cpu_count() or 1
    db = synthetic.bench_l_db(w["nx"], w["nz"], w["dx"], w["dz"])
    rlat, rlon, rdep = synthetic.receivers(w["nrcv"], (30.0, 70.0), w["dmin"], w["dmax"])
    o = OracleEngine(threads=ncores)
    configure(o, db, w, rlat, rlon, rdep)
    stype, allc, base = candidates(w, max(args.batch, 32))
    o.eval_sources(stype, base)
    set_references(o, [o], w["nrcv"], db.meta()["dt"])
    sample = max(1, min(args.batch, w["cpu_sample"]))
    cands = allc[:sample]
    for _ in range(min(args.warmup, 1)):
        o.time_eval(stype, cands[:1])
    t = 0.0
    for _ in range(args.steps):
        t += o.time_eval(stype, cands)
    value = sample * args.steps / t
    threads = olib().oracle_max_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(w, args, sample),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d candidate(s) per step of the same workload; restated CPU path (oracle/), not the Fortran binary" % sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def candidates(w, n):
    """candidate list of the workload and the base source the references come from"""
    from kiwi_b200 import synthetic
    if w.get("source") == "moment_tensor":
        side = max(1, int(round((n / 100.0) ** (1.0 / 3.0))))
        p = synthetic.moment_tensor_sweep(side, 100)
        reps = -(-n // p.shape[0])
        return "moment_tensor", np.tile(p, (reps, 1))[:n], p[p.shape[0] // 2 + 7]
    if w.get("source") == "eikonal":
        # time north east depth moment strike dip rake bord-x bord-y bord-radius nukl-x nukl-y rel-rupture-velocity rise-time
        # (rise time 0: with a rise time the end of the folded strip -- and with it the padded FFT length of the amplitude-spectrum norm --
        #  is decided by fp32 noise in the reference itself, DESIGN.md section 2, so the CPU check below would compare two coin flips)
        base = np.array([0, 0, 0, 15000, 2e20, 91, 87, 164, 0, 0, 9000, 0, 0, 0.8, 0.0], np.float32)
        p = np.tile(base, (n, 1))
        i = np.arange(n)
        nrad = 4 if n <= 300 else -(-n // 75)                             # (more radii where a larger batch is asked for: all candidates distinct)
        p[:, 10] = np.linspace(5000.0, 12000.0, nrad)[i % nrad]           # bord radius
        p[:, 13] = np.linspace(0.7, 0.9, 3)[(i // nrad) % 3]              # relative rupture velocity
        g5 = np.linspace(-3000.0, 3000.0, 5)
        p[:, 11] = g5[(i // (3 * nrad)) % 5]; p[:, 12] = g5[(i // (15 * nrad)) % 5]   # nucleation point on a 5 x 5 grid
        return "eikonal", p.astype(np.float32), base
    return "bilateral", synthetic.bilateral_sweep(n), synthetic.IZMIT


def workload_config(w, args, batch):
    if w.get("source") == "eikonal":
        return {"workload": "C4 eikonal rupture-front source (bord radius 5-12 km, rupture velocity 0.7-0.9 vs, nucleation on a 5x5 grid) "
                            "x %d receivers x ned, %s GFDB %dx%dx10, cosine taper + band-pass 0.02-0.1 Hz, %s, bilinear"
                            % (w["nrcv"], w["db"], w["nx"], w["nz"], w["norm"]),
                "name": args.workload, "candidates_per_step": batch, "receivers": w["nrcv"], "effective_dt": w["effective_dt"],
                "cache": "database %s exceeds L2 (no L2 flush needed)" % w["db"]}
    if w.get("source") == "moment_tensor":
        return {"workload": "C2 point moment-tensor grid search: (north, east, depth) lattice x 100 unit tensors, 3 centroids each, x %d receivers "
                            "x ned, %s GFDB %dx%dx10, %s, bilinear" % (w["nrcv"], w["db"], w["nx"], w["nz"], w["norm"]),
                "name": args.workload, "candidates_per_step": batch, "receivers": w["nrcv"], "effective_dt": w["effective_dt"],
                "cache": "database %s exceeds L2 (no L2 flush needed)" % w["db"]}
    return {"workload": "%s bilateral (Izmit, minimizer.f90:1632) ~1e4 sub-sources x %d receivers x ned, %s GFDB %dx%dx10, %s, "
                        "bilinear, candidates = lattice points of the strike/dip/rake/depth/length sweep of SURVEY.md 8d"
                        % ("C5 dense-array sweep:" if w["nrcv"] >= 2000 else "C3", w["nrcv"], w["db"], w["nx"], w["nz"], w["norm"]),
            "name": args.workload, "candidates_per_step": batch, "receivers": w["nrcv"], "effective_dt": w["effective_dt"],
            "cache": "database %s exceeds L2; every candidate streams its own node set (no L2 flush needed)" % w["db"]}


def roofline(wname, w, B, steps, stage, nsynth, synth_ms, b_alg, b_log):
    """roofline object of the dominant kernel of a workload"""
    peak, peak_kind = peaks()
    if w.get("source") == "moment_tensor" and stage[3] > stage[2]:
        # grid-search path: the fused synthesis + tensor-core contraction + misfit epilogue (k_mt_fused).  Algorithmic flops:
        # 2 * Ncand * 6 * (samples of all traces) (GF components x 6 MT components x N candidates, six non-zero
        # coefficients per candidate and component); executed: K = 32 per sample (7 rows -- six GF components and the reference --
        # in four hi/lo split products, padded).  Peak: measured dense bf16 / 2 (TF32 runs at half the bf16 rate).
        sum_t = sum(d.size for (f, d) in REFS.values())
        flops = 2.0 * B * 6.0 * sum_t
        ms_launch = stage[3] / max(steps, 1)
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1400.0) / 2.0 if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 700.0
        ach = flops / (ms_launch * 1e-3) / 1e12
        fused = stage[2] == 0
        roof = {"bound": "tensor", "kernel": "k_mt_fused" if fused else "k_mt_contract", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
                "peak_kind": "measured bf16 sustained / 2 (tf32)", "traffic": None, "algorithmic_flops_per_step": flops,
                "executed_flops_per_step": (32.0 / 6.0 if fused else 4.0) * flops, "ms_per_launch_group": ms_launch,
                "note": "K = 6 contraction: a (location, receiver) pair is ~600 samples x 100 candidates x K = 32, so the kernel is bound by what "
                        "surrounds the MMAs (gather + tap filter of the GF components, operand tiles, TMEM -> registers -> fp64 norm), "
                        "not by the tensor pipe (ncu: tensor pipe ~10 % active, issue slots ~35 %)"}
    else:
        # one "launch" = the depth-band launches of k_synth that together synthesise a sub-chunk of candidates
        ms_per_step = synth_ms / max(steps, 1)
        achieved = b_alg * B / (ms_per_step * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_synth", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_kind": peak_kind, "traffic": None, "algorithmic_bytes_per_eval": b_alg, "logical_bytes_per_eval": b_log,
                "evals_per_launch": float(B), "launches_per_step": nsynth / max(steps, 1), "ms_per_launch": ms_per_step,
                "note": "achieved = SURVEY.md 8(d) algorithmic bytes (no reuse assumed across receivers or candidates) over the "
                        "kernel time: with the depth-band launches the gather is served from L2, so this is the rate at which "
                        "node blocks reach the SMs, not DRAM traffic (dram_frac is)"}
    if roof["bound"] == "hbm" and roof["frac"] > 1.2:
        # candidates of the step share most of their sub-fault positions (config C4: 5 x 5 nucleation points and three rupture velocities
        # per rupture area): B_alg, which assumes no reuse across candidates, over-counts what has to reach the SMs, and the quotient is
        # not a bandwidth.  Reported for what it is, not as a fraction of the HBM peak.
        roof["b_alg_rate_over_hbm_peak"] = roof["frac"]
        roof["b_alg_rate"] = roof["achieved"]
        roof["achieved"] = None
        roof["frac"] = None
        roof["note"] += "; candidates share node sets here, so B_alg x evaluations / time exceeds any bandwidth: see b_alg_rate"
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            ent = json.load(open(tpath)).get(wname + ":" + roof["kernel"])
            if ent:   # ncu DRAM bytes per evaluation (one --set full capture per depth band) x evaluations per step
                roof["traffic"] = ent["bytes_per_eval"] * B
                roof["traffic_source"] = ent.get("source")
                if roof["bound"] == "hbm":
                    roof["dram_frac"] = roof["traffic"] / (roof["ms_per_launch"] * 1e-3) / 1e9 / peak
                    roof["traffic_over_algorithmic"] = ent["bytes_per_eval"] / b_alg
                    if "l2_hit_rate" in ent:
                        roof["l2_hit_rate"] = ent["l2_hit_rate"]
        except Exception:
            pass
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kiwi_b200", choices=["kiwi_b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="candidates per step and per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short C3 / C4 / C2 legs appended to the default C5 line")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.batch <= 0:
        args.batch = w["batch"]
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference_arm(args, w, rank)
        return

    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)          # anything a library prints on fd 1 (e.g. the NCCL version banner) goes to stderr
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import kiwi_b200
    from kiwi_b200 import synthetic
    db = make_db(w, rank, world, barrier)
    dt = db.meta()["dt"]
    rlat, rlon, rdep = synthetic.receivers(w["nrcv"], (30.0, 70.0), w["dmin"], w["dmax"])
    eng = kiwi_b200.Engine(local)
    configure(eng, db, w, rlat, rlon, rdep)
    B = args.batch
    # candidates are partitioned over the ranks; weak scaling: B per GPU
    stype, allc, base = candidates(w, max(B * world, 32))
    eng.set_source_params(stype, base)
    set_references(eng, [eng], w["nrcv"], dt)
    nm = eng.nmisfits
    # the product's sharded evaluator (kiwi_b200.sharding.ShardedEngine): equal counts and equal summed fault area per rank
    # (balanced_partition) -- the sweep's candidates differ in fault length, i.e. in the number of sub-sources, and the step time is
    # the slowest rank's -- then one all_gather of the misfit block over NCCL; the gathered block stays on the device
    from kiwi_b200.sharding import ShardedEngine, balanced_partition
    sharded = ShardedEngine(eng)
    pool = np.ascontiguousarray(allc[:B * world])
    cost = (pool[:, 9] + pool[:, 10]) * pool[:, 11] if stype == "bilateral" else np.ones(pool.shape[0])
    shares = balanced_partition(cost, world)
    mine = np.ascontiguousarray(pool[shares[rank]])

    def step_device():
        d_mis, d_st = sharded.eval_sources_device(stype, pool, partition=shares)
        return d_st, eng.last_timing()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        st, _t = step_device()
    assert not bool(st.any()), "candidate evaluation failed: %s" % st
    b_alg, b_log, nsamp, nskip = eng.last_batch_bytes(4)
    assert nskip == 0, "%d centroids fell outside the database" % nskip
    if w.get("source") == "moment_tensor":
        # the grid path synthesises 6 basis sources per location; bytes per *candidate* evaluation
        nloc = len({tuple(r) for r in np.concatenate([mine[:, :4], mine[:, 10:11]], 1).tolist()})
        scale = 6.0 * nloc / B if eng.last_timing()["launches"][3] and nloc * 8 <= B else 1.0
        b_alg, b_log = b_alg * scale, b_log * scale

    # ---- timed region 1: device-resident results (value) -----------------------------------------------
    barrier()
    t0 = time.perf_counter()
    t0_wall = time.time()
    dev_ms = synth_ms = 0.0
    launches = nsynth = 0
    stage = np.zeros(4)
    for _ in range(args.steps):
        st, t = step_device()
        dev_ms += t["total_ms"]; synth_ms += t["synthesis_ms"]; launches += sum(t["launches"]); nsynth += t["launches"][2]
        stage += [t["discretise_ms"], t["geometry_ms"], t["synthesis_ms"], t["misfit_ms"]]
    barrier()
    wall = time.perf_counter() - t0
    t1_wall = time.time()
    tt = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = [float(v) for v in tt.cpu()]
    # whole-job throughput, K steps timed with CUDA events on the engine's stream around each call
    # (device work + per-step table uploads), max over ranks
    value = world * B * args.steps / (dev_ms_max * 1e-3)

    # ---- timed region 2: end to end through the C ABI with host buffers (e2e) ----------------------------
    def step_e2e():
        if world == 1:
            return eng.eval_sources(stype, mine, pinned=True)        # page-locked result buffer (kiwi_host_alloc)
        return sharded.eval_sources(stype, pool, partition=shares)   # host buffers in, the complete answer out on every rank
    step_e2e()                                                       # untimed: allocates the page-locked result buffer
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mis, st = step_e2e()
    barrier()
    e2e_wall = time.perf_counter() - t0
    te = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(te.cpu()[0])
    if world > 1:
        mis = mis[shares[rank]]
    clocks = sampler.stop(t0_wall, t1_wall) if rank == 0 else None
    h2d = int(mine.nbytes + B * (36 + 100) + 7 * 8 * B)     # params + CandDev/BilatCand tables + STF taps
    d2h = int(mis.nbytes + st.nbytes + 4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    roof = roofline(args.workload, w, B, args.steps, stage, nsynth, synth_ms, b_alg, b_log)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(w, args, B),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roof,
            "stage_ms_per_step": {k: float(v) / args.steps for k, v in zip(["discretise", "geometry", "synthesis", "misfit"], stage)},
            "wall_ms_per_step": wall_ms_max / args.steps}

    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(w, db, rlat, rlon, rdep, dt, stype, mine, mis, eng)
    if world == 1 and args.workload == "c5" and not args.no_secondary:
        # the other configurations of BASELINE.json on the same database, a few steps each (the headline fields above are C5's)
        eng.close()
        line["secondary"] = {}
        for name in ("c3", "c4", "c2"):
            try:
                line["secondary"][name] = secondary_leg(name, db, local, args.no_cpu_baseline)
            except Exception as exc:      # a failing secondary leg must not take the headline line with it
                line["secondary"][name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(w, db, rlat, rlon, rdep, dt, stype, cands, gpu_misfits, geng=None):
    """The oracle (restated CPU path) timed on the host cores on a bounded sample of the same
    workload, and used as the checker of the batch just measured."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import OracleEngine, lib as olib
    ncores = os.cpu_count() or 1
    o = OracleEngine(threads=ncores)
    configure(o, db, w, rlat, rlon, rdep)
    copy_references(None, o, w["nrcv"], dt)
    n = max(1, min(w["cpu_sample"], cands.shape[0]))
    t0 = time.perf_counter()
    mo, so = o.eval_sources(stype, cands[:n])
    t = time.perf_counter() - t0
    # the same sample with the strip arithmetic of the restatement carried in double (oracle -DKO_WIDE):
    # separates the GPU's deviation from the fp32 reference path's own accumulation noise
    ow = OracleEngine(threads=ncores, wide=True)
    configure(ow, db, w, rlat, rlon, rdep)
    copy_references(None, ow, w["nrcv"], dt)
    mw, _ = ow.eval_sources(stype, cands[:n])

    def dev(a, b):
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 0.1 * np.abs(b[..., 1:2]))))
    d32, dw, d32w = dev(gpu_misfits[:n], mo), dev(gpu_misfits[:n], mw), dev(mo, mw)
    parity = {"gpu_vs_fp32_path": d32, "gpu_vs_double_accumulation": dw, "fp32_path_vs_double_accumulation": d32w, "tol": 1e-5}
    if geng is not None:
        # seismograms of the last candidate of the sample against the fp32 path, trace by trace on up to 100 receivers spread over
        # the array: max |gpu - oracle| / peak (the north_star bar; the misfits are differences of these traces)
        geng.set_source_params(stype, cands[n - 1])
        worst = {"seis_vs_fp32_path": 0.0, "seis_vs_double_accumulation": 0.0, "seis_fp32_path_vs_double_accumulation": 0.0}
        ntr = 0
        for ir in np.unique(np.linspace(1, w["nrcv"], min(w["nrcv"], 100)).astype(int)):
            for ic in range(1, 4):
                tg, to, tw = geng.get_seismogram(int(ir), ic), o.get_seismogram(int(ir), ic), ow.get_seismogram(int(ir), ic)
                for key, (fa, da), (fb, dbb) in (("seis_vs_fp32_path", tg, to), ("seis_vs_double_accumulation", tg, tw),
                                                 ("seis_fp32_path_vs_double_accumulation", to, tw)):
                    peak = float(np.abs(dbb).max()) if dbb.size else 0.0
                    if (fa, da.size) != (fb, dbb.size):
                        worst[key] = float("inf")      # spans must be equal
                    elif peak > 0:
                        worst[key] = max(worst[key], float(np.abs(da - dbb).max()) / peak)
                ntr += 1
        parity.update(worst)
        parity["seis_traces_compared"] = ntr
    if geng is not None:
        # the same sample in the reference's order of operations (kiwi_set_accumulation, csrc/synth_exact.cu): held against the fp32
        # restatement as it stands, at the north_star bar (1e-5 relative for seismograms and misfits)
        try:
            geng.set_accumulation(True)
            geng.eval_sources(stype, cands[:n])          # (untimed: the mode's buffers and the host worker threads come into being)
            mr, sr = geng.eval_sources(stype, cands[:n])
            tr_ = geng.last_timing()
            ro = {"misfit_vs_fp32_path": float(np.max(np.abs(mr - mo) / np.maximum(np.abs(mo), 1e-300))), "unit": UNIT,
                  "value": n / (tr_["total_ms"] * 1e-3) if tr_["total_ms"] > 0 else None}
            geng.set_source_params(stype, cands[n - 1])
            worst_ro = 0.0
            for ir in np.unique(np.linspace(1, w["nrcv"], min(w["nrcv"], 100)).astype(int)):
                for ic in range(1, 4):
                    (fa, da), (fb, dbb) = geng.get_seismogram(int(ir), ic), o.get_seismogram(int(ir), ic)
                    peak = float(np.abs(dbb).max()) if dbb.size else 0.0
                    if (fa, da.size) != (fb, dbb.size):
                        worst_ro = float("inf")
                    elif peak > 0:
                        worst_ro = max(worst_ro, float(np.abs(da - dbb).max()) / peak)
            ro["seis_vs_fp32_path"] = worst_ro
            ro["ok"] = bool(worst_ro <= 1e-5 and ro["misfit_vs_fp32_path"] <= 1e-5)
            ro["note"] = ("reference-order synthesis (every operation of make_seismogram per output sample in the reference's order) against "
                          "the fp32 restatement as it stands: max |a - b| / trace peak, max relative misfit deviation; value = its own rate")
            parity["reference_order"] = ro
        except Exception as exc:
            parity["reference_order"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        finally:
            geng.set_accumulation(False)
    sn = parity.get("seis_fp32_path_vs_double_accumulation", 0.0)
    parity["ok"] = bool(parity.get("seis_vs_double_accumulation", 0.0) <= 1e-5 and parity.get("seis_vs_fp32_path", 0.0) <= 1e-5 + 1.5 * sn and
                        dw <= 1e-5 and d32 <= max(1e-5, 2.0 * d32w))
    parity["note"] = ("seis_*: max |a - b| / trace peak over the sampled traces (spans equal), last candidate of the sample; gpu_*: "
                      "relative misfit deviation, relative to max(misfit, 0.1 x norm factor).  Bar: 1e-5 against the restatement with "
                      "the strips carried in double, and 1e-5 + 1.5 x the fp32 restatement's own distance from it against the fp32 "
                      "restatement: at ~1e4 sub-sources the reference's sequential fp32 accumulation is itself ~5e-5 of the trace peak "
                      "away from the exact sum of the same terms (tests/test_fullsize_parity_gpu.py, DESIGN.md section 5)")
    return {"value": n / t, "unit": UNIT, "cores": int(olib().oracle_max_threads()), "kind": "port",
            "sample": "first %d candidate(s) of the step, %.1f s; restated CPU path (oracle/), OpenMP over receivers" % (n, t),
            "parity": parity}


def secondary_leg(name, db, device, no_cpu, steps=4, warmup=3):
    """one of the other BASELINE.json configurations, a few steps on one GPU: value (device, CUDA events), e2e (host buffers through
    the C ABI), roofline, CPU baseline + parity on a bounded sample"""
    import kiwi_b200
    from kiwi_b200 import synthetic

    class A:
        workload = name
    w = dict(WORKLOADS[name])
    B = w["batch"]
    dt = db.meta()["dt"]
    rlat, rlon, rdep = synthetic.receivers(w["nrcv"], (30.0, 70.0), w["dmin"], w["dmax"])
    eng = kiwi_b200.Engine(device)
    try:
        configure(eng, db, w, rlat, rlon, rdep)
        stype, allc, base = candidates(w, max(B, 32))
        eng.set_source_params(stype, base)
        set_references(eng, [eng], w["nrcv"], dt)
        mine = np.ascontiguousarray(allc[:B])
        for _ in range(warmup):
            st = eng.eval_sources_on_device(stype, mine)
        assert not st.any(), "candidate evaluation failed: %s" % st
        b_alg, b_log, nsamp, nskip = eng.last_batch_bytes(4)
        if w.get("source") == "moment_tensor":
            nloc = len({tuple(r) for r in np.concatenate([mine[:, :4], mine[:, 10:11]], 1).tolist()})
            scale = 6.0 * nloc / B if eng.last_timing()["launches"][3] and nloc * 8 <= B else 1.0
            b_alg, b_log = b_alg * scale, b_log * scale
        dev_ms = synth_ms = 0.0
        nsynth = launches = 0
        stage = np.zeros(4)
        for _ in range(steps):
            eng.eval_sources_on_device(stype, mine)
            t = eng.last_timing()
            dev_ms += t["total_ms"]; synth_ms += t["synthesis_ms"]; nsynth += t["launches"][2]; launches += sum(t["launches"])
            stage += [t["discretise_ms"], t["geometry_ms"], t["synthesis_ms"], t["misfit_ms"]]
        # end to end: host parameters in, host results out.  The moment-tensor grid search returns what its driver
        # (kiwi_b200.grid_search.MisfitGrid) reads: the global misfit of every candidate and the best one, reduced on the device
        grid = w.get("source") == "moment_tensor"

        def e2e_step():
            if grid:
                eng.eval_sources_on_device(stype, mine)
                return eng.outer_misfits(B, outer_norm="l2norm")
            return eng.eval_sources(stype, mine, pinned=True)
        e2e_step()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = e2e_step()
        e2e_s = time.perf_counter() - t0
        d2h = int(res[0].nbytes + res[1].nbytes + (res[2].nbytes if grid else 0))
        out = {"config": workload_config(w, A, B), "value": B * steps / (dev_ms * 1e-3), "unit": UNIT, "steps": steps, "warmup": warmup,
               "ms_per_step": dev_ms / steps,
               "e2e": {"value": B * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(mine.nbytes), "d2h_bytes_per_step": d2h,
                       "result": "global misfit per candidate + best candidate (outer l2norm on the device)" if grid else "misfit block"},
               "gpu_launches": int(launches), "roofline": roofline(name, w, B, steps, stage, nsynth, synth_ms, b_alg, b_log),
               "stage_ms_per_step": {k: float(v) / steps for k, v in zip(["discretise", "geometry", "synthesis", "misfit"], stage)}}
        if not no_cpu:
            mis, st = eng.eval_sources(stype, mine[:max(1, w["cpu_sample"])])
            out["cpu_baseline"] = cpu_baseline(w, db, rlat, rlon, rdep, dt, stype, mine, mis, eng)
        return out
    finally:
        eng.close()


if __name__ == "__main__":
    main()
