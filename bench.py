#!/usr/bin/env python
"""Benchmark of the particle-filter hot path (BASELINE.json metric).

A "step" is one whole filter evaluation -- what ``particle_filter$run(pars)`` does on
the compiled-compare path (R/particle_filter_state.R:138-162): reset the particles to
the model's initial state, then 100 data intervals of 4 model steps each with compare,
log-mean-exp, systematic resampling and state gather -- on the SIR model at
2^20 particles (config C2).  At N > 1 GPUs every rank runs its own filter (independent
chains, the pmcmc_parallel layout: R/pmcmc_parallel.R:8-20), so scaling is weak and the
headline has no collective; ranks only meet at the timing barrier.  The configurations
that DO exchange data between GPUs -- C3 pmcmc chains, C4 nested populations, C5 smc2 --
are timed after the headline in the same launch (bench_configs.py) and reported under
"configs" in the same JSON line (MCB_BENCH_CONFIGS=0 skips them).

  python bench.py --gpus N --steps K --warmup W          (this engine)
  python bench.py --impl reference ...                   (CPU arm: the oracle port on
                                                          all host cores; dust/R absent)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PARTICLES = 1 << 20
RATE = 4
METRIC = "SIR particle-steps/sec (particle_filter$run, 2^20 particles per GPU)"
UNIT = "particle-steps/s"
# Algorithmic bytes per particle per data interval.  SURVEY.md 8(d) counts 264 B for the
# unfused pipeline (K1 152 + K2 8 + K3 20 + K4 84).  The engine fuses K3+K4 into K1's load,
# so the bytes its kernels must move are fewer; the per-kernel roofline uses THESE (the
# conservative figure), the whole-interval fraction uses the survey's 264 B.
#   k_run_compare : ancestor slot 4 + state in 40 (gathered) + rng in 32
#                   + state out 40 + rng out 32 + logw out 8                      = 156
#   k_weights_scan: logw in 8 + cumulative weight out 8 + ancestor slot emptied 4 = 20
#   k_ancestors   : cumulative weight in 8 + ancestor slot out 4 (one per
#                   particle with offspring, at most)                             = 12
#   k_finish_state: (once per filter call) slot 4 + state in 40 + state out 40    = 84
BYTES_K1, BYTES_K2, BYTES_K3, BYTES_FINISH = 156, 20, 12, 84
BYTES_ENGINE_INTERVAL = BYTES_K1 + BYTES_K2 + BYTES_K3      # what this engine moves: 188
BYTES_INTERVAL = 264                                         # SURVEY.md 8(d), unfused


def load_data():
    d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "sir_incidence.csv"), delimiter=",",
                   skiprows=1)
    return d[:, 1].astype(np.uint64) * RATE, d[:, 0]


def ncu_traffic_bytes(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the
    committed summary of the last `ncu --set full` capture (profiles/); None if absent."""
    import csv
    import glob

    import re

    def version(path):   # ..._r01_v26.csv -> (1, 26): the latest capture, not the last name
        return tuple(int(x) for x in re.findall(r"\d+", os.path.basename(path)))

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_full_summary_r*.csv")),
                   key=version)
    if not files:
        return None, None
    try:
        rows = list(csv.reader(open(files[-1])))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        vals = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
                for r in rows[2:] if kernel in r[0]]
        return (sum(vals) / len(vals) if vals else None), os.path.basename(files[-1])
    except Exception:
        return None, None


def ncu_issue_counters(kernel):
    """Issue-slot and pipe utilisation of `kernel` (percent of peak, sustained, active
    cycles) from the same committed `ncu --set full` summary as the traffic: the update
    kernel is bound by instruction issue, not by HBM, so this is the roofline it is
    read against (BASELINE north_star: "a fraction of the HBM or issue roofline").
    None if the summary is absent or does not hold the kernel."""
    import csv
    import glob
    import re

    def version(path):
        return tuple(int(x) for x in re.findall(r"\d+", os.path.basename(path)))

    try:
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_full_summary_r*.csv")),
                       key=version)
        rows = list(csv.reader(open(files[-1])))
        hdr = rows[0]
        keys = {"issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                "warp_instructions": "smsp__inst_executed.sum"}
        mine = [r for r in rows[2:] if kernel in r[0]]
        if not mine:
            return None
        out = {k: sum(float(r[hdr.index(col)]) for r in mine) / len(mine)
               for k, col in keys.items()}
        out["source"] = os.path.basename(files[-1])
        return out
    except Exception:
        return None


def ncu_launch_list_us(kernel):
    """Average duration (us) of `kernel` in the committed ncu launch list of the latest
    capture (profiles/ncu_launches_r*.csv: `gpu__time_duration.sum` per launch, serialised,
    cold caches) -- the kernel's own execution time, beside the CUDA-event figure, which for
    a short kernel also holds the launch.  (None, None) if absent."""
    import csv
    import glob
    import re

    def version(path):
        return tuple(int(x) for x in re.findall(r"\d+", os.path.basename(path)))

    try:
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_launches_r*.csv")), key=version)
        rows = list(csv.reader(open(files[-1])))
        hdr = next(r for r in rows if "Kernel Name" in r and "Metric Value" in r)
        kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        vals = [float(r[mv].replace(",", "")) for r in rows[rows.index(hdr) + 1:]
                if len(r) > mv and ("::" + kernel) in r[kn]]
        return (sum(vals) / len(vals) / 1e3 if vals else None), os.path.basename(files[-1])
    except Exception:
        return None, None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, active=True):
        # `active = False` (every rank but 0): a no-op.  One sampler per box is enough, and
        # nvidia-smi queries from eight processes at once contend for the driver with the
        # very launches and synchronisations being timed.
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.active = active

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.active:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names)
                   if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows),
                "power_w_max": max(float(r[2]) for r in self.rows)}


def cpu_baseline(n_particles, n_threads, reps=1):
    """The oracle port (OpenMP over particles, serial scan) on the host cores."""
    from oracle import oracle as O

    t, x = load_data()
    m = O.OracleModel(n_particles=n_particles, seed=1, n_threads=n_threads)
    m.set_data(t, x)
    best = None
    ll = None
    for _ in range(reps + 1):  # first is warm-up
        m.update_state(pars=O.default_pars(), time=0)
        t0 = time.perf_counter()
        ll = m.filter()["log_likelihood"][0]
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_particles * int(t[-1]) / best, best, ll


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = N_PARTICLES   # the configuration the metric is quoted on, not a fraction
    t, _ = load_data()
    from oracle import oracle as O

    O.build()
    times = []
    m = O.OracleModel(n_particles=n_sample, seed=1, n_threads=cores)
    tt, x = load_data()
    m.set_data(tt, x)
    for i in range(args.warmup + args.steps):
        m.update_state(pars=O.default_pars(), time=0)
        t0 = time.perf_counter()
        m.filter()
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n_sample * int(t[-1]) * len(times) / total
    sample = ("one whole filter evaluation per step: %d particles x 100 intervals x 4 steps, "
              "the workload itself" % n_sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (reference's inst/sir_incidence.csv, 100 days, rate 4)",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "mcstate is pure R and its native dependency dust is absent from the image, so "
                "the reference arm is the C/OpenMP oracle port of dust's filter loop",
    }
    emit(line)


def config_dict(n_gpus):
    return {"workload": "SIR particle_filter$run, 2^20 particles x 100 data intervals x 4 steps "
                        "(BASELINE config C2), one independent filter per GPU",
            "n_particles_per_gpu": N_PARTICLES, "n_intervals": 100, "rate": RATE,
            "parallelism": "chains x%d" % n_gpus,
            "l2": "live set 138 MB per GPU exceeds the 126 MB L2; a 256 MB buffer is also "
                  "written between steps"}


def run_engine(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mcstate_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    run_configs = os.environ.get("MCB_BENCH_CONFIGS", "1") != "0"
    configs_error = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif run_configs:
        # one rank: the sharded callers of bench_configs.py still take a process group
        import socket

        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        try:
            dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0,
                                    world_size=1, device_id=torch.device("cuda", local))
        except Exception as exc:   # the headline does not need the group: say so and go on
            configs_error = "process group: %r" % (exc,)
            run_configs = False

    from mcstate_b200 import dust
    from mcstate_b200.particle_filter import particle_filter
    from mcstate_b200.particle_filter_data import particle_filter_data

    t, x = load_data()
    n_steps_model = int(t[-1])
    gen = dust.dust_example("sir")
    # chain k on rank k: seeds by long jump so placement does not matter (R/pmcmc_parallel.R:133)
    seed = dust.dust_rng_distributed_state(1, 1, world)[rank]
    data = particle_filter_data({"day": (t // RATE).astype(np.int64), "incidence": x}, "day",
                                RATE, 0)
    # the user-facing object: particle_filter$new(data, model, n_particles, compare = NULL)
    pf = particle_filter(data, gen, N_PARTICLES, None, seed=seed, gpu_config=local)
    pars = {"beta": 0.2, "gamma": 0.1}
    pf.run(pars)                       # creates the model object and uploads the data once
    model = pf._last_model[0]
    # everything of the timed region -- the L2 flush, the engine's kernels, the events --
    # is ordered on ONE stream torch knows about (not torch's default stream: its handle
    # is 0, which the engine reads as "the handle's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    model.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        flush.zero_()
        model.update_state(pars=pars, time=0)
        model.filter_enqueue()
        return model.filter_collect()

    def step_e2e():
        # particle_filter$run(pars): host parameters in, host log-likelihood out
        flush.zero_()
        return np.atleast_1d(pf.run(pars))

    for _ in range(args.warmup):
        step_resident()
    # ---- timed region: K filter evaluations, state resident in HBM ------------------
    launches0 = model.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    # clocks / throttle reasons are sampled from here to the end of the e2e loop: three
    # back-to-back passes over the same workload (the headline pass alone lasts ~0.1 s,
    # one nvidia-smi query)
    clocks = ClockSampler(local, active=rank == 0)
    clocks.__enter__()
    ev0.record(stream)
    for _ in range(args.steps):
        ll = step_resident()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = model.launch_count() - launches0

    # ---- the same K evaluations again with CUDA events recorded around every kernel
    # inside the graph (per-kernel durations for the roofline; the records serialise
    # the launches a little, so this pass is not the headline number)
    model.set_kernel_timing(True)
    step_resident()  # builds the instrumented graph
    model.kernel_timing()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev2.record(stream)
    for _ in range(args.steps):
        step_resident()
    ev3.record(stream)
    barrier()
    ms_instrumented = ev2.elapsed_time(ev3)
    k_ms, k_n = model.kernel_timing()
    model.set_kernel_timing(False)

    # ---- end to end through the host API (host pars in, host log-likelihood out) ---
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks.__exit__(None, None, None)

    # ---- not the headline: BASELINE config C3's filter (one pMCMC chain of 2^16 particles
    # per GPU) on the single-launch path (csrc/persistent.cuh) and on the per-interval path
    small = None
    if rank == 0:
        small = {"n_particles": 1 << 16, "unit": "ms per filter evaluation (100 intervals)"}
        for name, on in (("single_launch", True), ("two_launch", False)):
            ms = gen.new(pars, 0, 1 << 16, seed=1, gpu_config=local)
            ms.set_persistent(on)
            ms.set_data(dust.dust_data({"time_end": t, "incidence": x}), False)
            for _ in range(3):
                ms.update_state(pars=pars, time=0)
                ms.filter()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                ms.update_state(pars=pars, time=0)
                ms.filter()
            torch.cuda.synchronize()
            small[name] = (time.perf_counter() - t0) / 20 * 1e3

    tmax = torch.tensor([ms_total, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(tmax[0]), float(tmax[1])
    # ---- the configurations with a collective (C3, C4, C5), same launch, own clocks
    configs = None
    if run_configs:
        import bench_configs

        del pf, model
        try:
            configs = bench_configs.run_configs(local,
                                                lambda idx: ClockSampler(idx, active=rank == 0))
        except Exception as exc:   # the headline line is printed whatever happens here
            import traceback

            traceback.print_exc()
            configs_error = repr(exc)
    if configs is None and configs_error is not None:
        configs = {"error": configs_error}

    work = world * N_PARTICLES * n_steps_model * args.steps
    value = work / (ms_total * 1e-3)
    e2e_value = work / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        names = ["k_run_compare", "k_weights_scan", "k_ancestors", "k_finish_state"]
        per_launch_bytes = [BYTES_K1 * N_PARTICLES, BYTES_K2 * N_PARTICLES,
                            BYTES_K3 * N_PARTICLES, BYTES_FINISH * N_PARTICLES]
        kernels = {}
        busy_ms = float(sum(k_ms[k] for k in range(len(names)) if k_n[k]))
        for k, name in enumerate(names):
            if k_n[k]:
                avg_ms = k_ms[k] / float(k_n[k])
                # share_of_step: of the time the step's kernels run (what an ncu launch
                # list can be compared with); share_of_wall: of the instrumented pass,
                # which also holds the event records and launch gaps
                ncu_us, ncu_src = ncu_launch_list_us(name)
                kernels[name] = {"avg_us": 1e3 * avg_ms, "launches": int(k_n[k]),
                                 "ncu_launch_list_us": ncu_us, "ncu_launch_list": ncu_src,
                                 "share_of_step": k_ms[k] / busy_ms,
                                 "share_of_wall": k_ms[k] / ms_instrumented,
                                 "achieved_gbs": per_launch_bytes[k] / (avg_ms * 1e-3) / 1e9}
        dom = max(kernels, key=lambda n: kernels[n]["share_of_step"]) if kernels else None
        roofline = None
        if dom:
            a = kernels[dom]["achieved_gbs"]
            traffic, traffic_src = ncu_traffic_bytes(dom)
            roofline = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s",
                        "frac": a / peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_kind,
                        "algorithmic_bytes_per_launch": per_launch_bytes[names.index(dom)],
                        # the whole pipeline on the bytes this engine moves per particle and
                        # interval (188 B: the gather is fused into K1's load) -- and, for
                        # comparison only, on SURVEY.md 8(d)'s 264 B of an unfused pipeline
                        "whole_interval_frac": (BYTES_ENGINE_INTERVAL * N_PARTICLES * 100 *
                                                args.steps / (ms_total * 1e-3) / 1e9) / peak,
                        "whole_interval_frac_unfused_264B": (
                            BYTES_INTERVAL * N_PARTICLES * 100 * args.steps /
                            (ms_total * 1e-3) / 1e9) / peak,
                        "instrumented_ms_per_step": ms_instrumented / args.steps,
                        "issue": ncu_issue_counters(dom),
                        "kernels": kernels}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1:
            n_cpu = N_PARTICLES
            v, secs, ll_cpu = cpu_baseline(n_cpu, cores, reps=2)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "oracle filter at %d particles x 100 intervals (one GPU's whole "
                             "workload), best of 2 after a warm-up run, %.2f s each" % (n_cpu, secs)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (reference's inst/sir_incidence.csv, 100 days, rate 4)",
            "config": config_dict(world),
            "filter_evals_per_s": world * args.steps / (ms_total * 1e-3),
            "log_likelihood": float(ll[0]),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * 8,
                    "d2h_bytes_per_step": 8 + 8,
                    "api": "particle_filter(data, model, 2^20, compare=None).run(pars)"},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks.summary(),
            "small_filter_c3": small,
            "configs": configs,
        }
        emit(line)
    if dist.is_initialized():
        dist.destroy_process_group()


_REAL_STDOUT = None


def own_stdout():
    """stdout carries ONE JSON line and nothing else: whatever a library prints there (NCCL's
    version banner, for one) is sent to stderr for the whole run, and the line is written to
    the real stdout at the end."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    args = ap.parse_args()
    own_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
