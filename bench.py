#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json): closed-loop env-steps/s of the batched rollout
and HJB-residual states/s of the vhjb pass, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic input (one rollout launch over all
environments of this rank, or one vhjb residual+gradient+Adam step over this rank's sampled states).
Rank 0 prints ONE JSON line.  N > 1: launched under torchrun, one rank per GPU; environments / states are sharded
by rank (weak scaling: the per-GPU batch is fixed), no data-path collective for rollouts.

``--impl reference`` times the reference's own CPU implementation of the same workload on the host cores: the UNMODIFIED
reference loop (``u = ctl.get_control_efforts(x); x = dyn.simulate(x, u)``, scripts/test_vhjb_policy.py:146-151) from the
travelling copy under baseline/_ref (tools/install_reference_baseline.py; /root/reference in the build container), one
process per core on a bounded sample — ``cpu_baseline.kind = "reference"``.  The vectorised NumPy port
(oracle/rollout_oracle.py, fp64 and fp32) is reported beside it; it is the fallback when no reference copy is present.

What this file takes from ``oracle/`` (test infrastructure): the CHECKER of the untimed ``parity`` block, the ``cpu_baseline``
legs and the reference arm.  Everything on the measured arm — dynamics, controllers, value-net weights, synthetic batches,
initial states — comes from the package (``q_learning_with_hjb_b200/workloads.py``, ``hjb_sample_states``).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ---------------------------------------------------------------------------------------------------------------
# workloads.  flops = "FLOP model v1" per env-step; bytes = HBM bytes per env-step when recording.
# ---------------------------------------------------------------------------------------------------------------
ROLLOUTS = {
    # name: system, controller, envs per GPU, horizon, flops/step (euler, rk4), config label
    "quad2d_hover": dict(sys="quad2d", ctl="quad2d_hover", envs=1 << 24, T=1000, flops={"euler": 65, "rk4": 158},
                         label="C4: 2-D drone hovering, quadrotors model-based controller, 16M envs x 1000 steps"),
    "cartpole_lqr": dict(sys="cartpole", ctl="cartpole_lqr", envs=4096, T=500, flops={"euler": 50, "rk4": 151},
                         label="C1: cartpole LQR balancing, 4096 initial states x 500 steps"),
    "cartpole_lqr_big": dict(sys="cartpole", ctl="cartpole_lqr", envs=1 << 24, T=500, flops={"euler": 50, "rk4": 151},
                             label="cartpole LQR balancing, 16M initial states x 500 steps"),
    "acrobot_es": dict(sys="acrobot", ctl="acrobot_es", envs=1 << 22, T=2000, flops={"euler": 156, "rk4": 302}, parity_T=25,
                       label="C3: acrobot energy-shaping swing-up, 4M envs x 2000 steps"),
    "quad10d_hover": dict(sys="quad10d", ctl="quad10d_hover", envs=1 << 23, T=1000, flops={"euler": 124, "rk4": 258},
                          label="10-D quadcopter hover LQR, 8M envs x 1000 steps"),
}
DEFAULT_WORKLOAD = "quad2d_hover"


X0_SPEC = {"quad2d": ([0] * 6, [1] * 6), "cartpole": ([0, 3.14, 0, 0], [2.4, 0.05, 1, 0.05]),
           "acrobot": ([0] * 4, [0.1] * 4), "quad10d": ([0] * 10, [1, 1, 1, .5, .5, 1, 1, 1, .5, .5]),
           "linear": ([0, 0], [1, 1])}


def synthetic_x0_host(sysname, count, seed, first=0):
    """Synthetic initial states: the system's own x0 distribution wrap(U(-std, std) + mean), fp32, from the
    counter-based stream (Philox4x32-10 keyed by the seed, counted by the environment index).  This is the NumPy twin
    (oracle/x0_stream.py) of the device generator hjb_sample_states: the same numbers, bit for bit — the timed runs
    generate them on the device (synthetic_x0_device), the CPU legs and the parity check here."""
    from oracle import x0_stream as X
    mean, std = X0_SPEC[sysname]
    return X.sample_states(sysname, np.float32(mean), np.float32(std), int(seed), int(first), int(count))


def synthetic_x0_device(dyn, sysname, count, seed, first=0):
    mean, std = X0_SPEC[sysname]
    return dyn.sample_initial_states(count, seed=seed, first=first, mean=np.float32(mean), std=np.float32(std))


def rollout_parity(w, integ, fast=True, envs=4096, seed=1234, record_stride=0):
    """Untimed parity leg of the timed plan: the SAME kernel instantiation (fast / accurate trigonometry, final state +
    Q = I, R = I cost, record stride) on the first `envs` environments of the SAME synthetic x0, against the oracle
    (oracle/rollout_oracle.py, NumPy fp64) over the workload's horizon — the acrobot (chaotic) over its stated 50 steps.
    Errors are |a - b| / max(1, |b|), angles modulo 2 pi (tests/helpers.py::rel_err).  Bounds: final states 1e-5 (the
    north star's trajectory bound); the accumulated cost 1e-4 (the bound of the notebook known-answer tests: l contains
    u = -K dx, which amplifies the state error by |K| — 34 for the cart-pole)."""
    import torch
    from oracle import rollout_oracle as O
    from tests.helpers import WRAP_IDX, make_controller, make_dynamics, rel_err
    from q_learning_with_hjb_b200.rollout import BatchedRollout, RunningCost

    T = w.get("parity_T", w["T"])
    dyn = make_dynamics(w["sys"])
    dyn.fast_trig = bool(fast)
    ctl = make_controller(w["ctl"], dyn)
    n, m = dyn.get_dimension()
    xf = getattr(ctl, "xf", np.zeros(n)); uf = getattr(ctl, "uf", np.zeros(m))
    plan = BatchedRollout(dyn, ctl, envs, T, integrator=integ, record_stride=record_stride,
                          cost=RunningCost(np.eye(n), np.eye(m), xf, uf))
    x0 = synthetic_x0_host(w["sys"], envs, seed)
    res = plan.launch(torch.as_tensor(x0).cuda())
    torch.cuda.synchronize()
    osys = O.std_system(w["sys"])
    octl = O.std_controller(w["ctl"], osys)
    ocost = O.OracleCost(np.eye(n), np.eye(m), np.asarray(xf, dtype=np.float64), np.asarray(uf, dtype=np.float64))
    _, _, xfo, Jo = O.rollout(osys, octl, x0.astype(np.float64), T, integ, record_stride=0, cost=ocost)
    ex = rel_err(res.x_final.cpu().numpy(), xfo, WRAP_IDX[w["sys"]])
    ec = rel_err(res.cost.cpu().numpy(), Jo)
    return {"against": "oracle/rollout_oracle.py (NumPy fp64 restatement of the reference loop)", "envs": envs, "horizon": T,
            "integrator": integ, "kernel_variant": plan.kernel_variant(), "max_rel_err_x_final": ex,
            "max_rel_err_cost": ec, "max_rel_err": ex, "tolerance": 1e-5, "tolerance_cost": 1e-4,
            "ok": bool(ex <= 1e-5 and ec <= 1e-4),
            "checksum_mean_cost": float(res.cost.double().mean()), "oracle_mean_cost": float(Jo.mean()),
            "error_measure": "|a - b| / max(1, |b|), angles modulo 2 pi"}


# ---------------------------------------------------------------------------------------------------------------
# CPU side (cpu_baseline and --impl reference): the unmodified reference loop, and the vectorised oracle port
# ---------------------------------------------------------------------------------------------------------------
def _cast_floats(obj, dt):
    for k, v in list(vars(obj).items()):
        if isinstance(v, np.ndarray) and v.dtype.kind == "f":
            setattr(obj, k, v.astype(dt))
        elif isinstance(v, dict):
            for kk, vv in v.items():
                if isinstance(vv, np.ndarray) and vv.dtype.kind == "f":
                    v[kk] = vv.astype(dt)


def _cpu_rollout_worker(args):
    sysname, ctlname, envs, T, integ, seed, dtype = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import rollout_oracle as O
    dt = np.dtype(dtype).type
    O.set_dtype(dt)
    osys = O.std_system(sysname)
    octl = O.std_controller(ctlname, osys)
    ocost = O.OracleCost(np.eye(osys.n), np.eye(osys.m), octl.xf if octl.xf is not None else np.zeros(osys.n),
                         octl.uf if octl.uf is not None else np.zeros(osys.m))
    for o in (osys, octl, ocost):
        _cast_floats(o, dt)
    x0 = synthetic_x0_host(sysname, envs, seed).astype(dt)
    t0 = time.perf_counter()
    O.rollout(osys, octl, x0, T, integ, record_stride=0, cost=ocost)
    return time.perf_counter() - t0


def cpu_rollout_throughput(w, integ, cores, envs_per_core, T, dtype="float64"):
    """env-steps/s of the oracle port with one process per core (each integrates envs_per_core x T)."""
    ctx = mp.get_context("spawn")
    jobs = [(w["sys"], w["ctl"], envs_per_core, T, integ, 99 + i, dtype) for i in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_rollout_worker, jobs)
    wall = time.perf_counter() - t0
    return cores * envs_per_core * T / wall, wall


def reference_available():
    from oracle import ref_loader as R
    return R.available()


def _make_reference_pair(sysname, ctlname, cached_riccati=True):
    """The reference's OWN classes (unmodified, through the import stubs) for a bench workload."""
    import importlib
    from oracle import ref_loader as R
    if sysname == "quad2d":
        dyn = R.make_quad2d()
        ctl = importlib.import_module("controller.quadrotors_model_based_controller").Quadrotors2DHoveringController(
            dyn, np.zeros(6), np.eye(6), np.eye(2))
    elif sysname == "quad10d":
        dyn = R.make_quad10d()
        ctl = importlib.import_module("controller.quadrotors_model_based_controller").NearHoverQuadcopterHoveringController(
            dyn, np.zeros(10), np.eye(10), np.eye(3))
    elif sysname == "cartpole":
        dyn = R.make_cartpole()
        es = importlib.import_module("controller.cartpole_energy_shaping").CartpoleEnergyShapingController(dyn)
        if ctlname == "cartpole_es":
            ctl = R.CachedLqrTerm(es) if cached_riccati else es
        else:                                       # the notebook's LQR about xf (cartpole_balancing.ipynb cell 4:24-25)
            K, _ = es.get_lqr_term()
            xf = np.array([0, 3.1415926, 0, 0])

            class NotebookLqr:
                def get_control_efforts(self, x):
                    return -K @ dyn.states_wrap(x - xf)
            ctl = NotebookLqr()
    elif sysname == "acrobot":
        dyn = R.make_acrobot()
        es = importlib.import_module("controller.acrobot_energy_shaping").AcrobotEnergyShapingController(dyn)
        ctl = R.CachedLqrTerm(es) if cached_riccati else es
    else:
        raise ValueError(sysname)
    return dyn, ctl


def _cpu_reference_worker(args):
    """The reference's loop, one environment at a time (scripts/test_vhjb_policy.py:146-151), on this process's core."""
    sysname, ctlname, envs, T, seed, cached = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    dyn, ctl = _make_reference_pair(sysname, ctlname, cached)
    x0s = synthetic_x0_host(sysname, envs, seed).astype(np.float64)
    t0 = time.perf_counter()
    for e in range(envs):
        x = x0s[e]
        for _ in range(T):
            u = ctl.get_control_efforts(x)
            x = dyn.simulate(x, u)
    return time.perf_counter() - t0


def cpu_reference_throughput(w, cores, envs_per_core, T, cached=True):
    """env-steps/s of the UNMODIFIED reference loop, one process per core."""
    ctx = mp.get_context("spawn")
    jobs = [(w["sys"], w["ctl"], envs_per_core, T, 99 + i, cached) for i in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_reference_worker, jobs)
    wall = time.perf_counter() - t0
    return cores * envs_per_core * T / wall, wall


def cpu_baseline_rollout(w, integ, T, literal_envs=64, port_envs=None):
    """The `cpu_baseline` object of a rollout line: the reference-literal loop when a copy of the
    reference is present (Euler: the reference has no other integrator), the vectorised port (fp64 and fp32) beside it."""
    cores = os.cpu_count() or 1
    port_envs = port_envs or max(256, int(4e6 // T))
    v64, wall64 = cpu_rollout_throughput(w, integ, cores, port_envs, T, "float64")
    v32, wall32 = cpu_rollout_throughput(w, integ, cores, port_envs, T, "float32")
    port = {"fp64": v64, "fp32": v32, "unit": "env-steps/s",
            "sample": f"{cores} procs x {port_envs} envs x {T} {integ} steps, oracle/rollout_oracle.py (NumPy, vectorised "
                      f"over environments), {wall64:.1f} s + {wall32:.1f} s"}
    if reference_available() and integ == "euler":
        Tl = min(T, 1000)
        v, wall = cpu_reference_throughput(w, cores, literal_envs, Tl, cached=True)
        out = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "reference",
               "sample": f"{cores} procs x {literal_envs} envs x {Tl} Euler steps of the UNMODIFIED reference loop "
                         f"(scripts/test_vhjb_policy.py:146-151; NumPy fp64, one environment per call"
                         + ("; Riccati solution cached" if w["ctl"].endswith("_es") else "") + f"), {wall:.1f} s",
               "port": port}
        if w["ctl"].endswith("_es"):   # as written, the energy-shaping controllers re-solve the Riccati equation every step
            vu, wallu = cpu_reference_throughput(w, cores, 2, min(Tl, 200), cached=False)
            out["as_written_riccati_per_step"] = {"value": vu, "sample": f"{cores} procs x 2 envs x {min(Tl, 200)} steps, {wallu:.1f} s"}
        return out
    return {"value": v64, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": port["sample"], "port": port}


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples that arrived inside [t_begin, t_end] (host wall clock around the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        rows = [r for (ts, r) in self.rows if t_begin is None or (t_begin - 0.02 <= ts <= t_end + 0.08)]
        if not rows:
            rows = [r for (_, r) in self.rows]
        sm, smax, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples in the upper half of the observed range (idle samples bracket the timed region)
        load = [v for v in sm if v >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_reference_arm(args, w, integ, rank, world):
    """The reference's own CPU implementation of the workload on the host cores.  Rank 0 only.  Each bench step is a
    bounded sample: 64 environments per core through the UNMODIFIED reference loop (when a copy of the reference is
    present: baseline/_ref on the GPU box), else the vectorised port."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    T = w["T"]
    literal = reference_available() and integ == "euler"
    if literal:
        Tl, epc = min(T, 1000), 64

        def step(envs=epc):
            return cpu_reference_throughput(w, cores, envs, Tl, cached=True)
        kind = "reference"
        sample = (f"{cores} procs x {epc} envs x {Tl} Euler steps per bench step of the UNMODIFIED reference loop "
                  "(scripts/test_vhjb_policy.py:146-151, one environment per call, NumPy fp64"
                  + ("; Riccati solution cached" if w["ctl"].endswith("_es") else "") + ")")
    else:
        epc = max(256, min(w["envs"] // cores, int(6e6 // T)))

        def step(envs=epc):
            return cpu_rollout_throughput(w, integ, cores, envs, T)
        kind = "port"
        sample = f"{cores} procs x {epc} envs x {T} {integ} steps per bench step (oracle/rollout_oracle.py, NumPy fp64)"
    vals, t_all = [], 0.0
    for _ in range(args.warmup):
        step(max(4, epc // 8))
    for _ in range(args.steps):
        v, wall = step()
        vals.append(v); t_all += wall
    value = float(np.mean(vals))
    cpu = {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample}
    if literal:   # the vectorised port beside it (the best-effort CPU comparator), fp64 and fp32
        pe = max(256, int(4e6 // T))
        cpu["port"] = {"fp64": cpu_rollout_throughput(w, integ, cores, pe, T, "float64")[0],
                       "fp32": cpu_rollout_throughput(w, integ, cores, pe, T, "float32")[0], "unit": "env-steps/s",
                       "sample": f"{cores} procs x {pe} envs x {T} steps, oracle/rollout_oracle.py"}
    line = {
        "impl": "reference", "metric": "closed-loop env-steps/s", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the b200 arm's config keys (same workload), plus what the bounded sample of one step was
        "config": {"workload": w["label"], "envs_per_gpu": args.envs or w["envs"], "horizon": T, "integrator": integ,
                   "record": "per-env cost" if args.record_stride == 0 else f"every {args.record_stride} steps",
                   "trig": "libm (NumPy fp64)", "parallelism": f"{cores} host processes", "seed": "1234 + rank",
                   "sample": sample},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def ncu_traffic(tag, family):
    """DRAM bytes per launch of the dominant kernel from this round's ncu capture (profiles/traffic.json, written by
    profiles/summarize_ncu.py), quoted ONLY when the capture was made on the kernel sources being timed (the entry's
    kernel_source_hash equals the hash of the current sources); otherwise (None, why)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(tag)
    except (OSError, ValueError):
        t = None
    if not t:
        return None, "no ncu capture of this configuration"
    from q_learning_with_hjb_b200.build import source_hash
    if t.get("kernel_source_hash") != source_hash(family):
        return None, f"stale: {tag} was captured on other kernel sources ({t.get('kernel_source_hash')})"
    return t["dram_bytes_per_launch"], f"profiles/{tag}.md, kernel sources {t['kernel_source_hash']}"


def rollout_roofline(w, integ, n, m, envs, T, kernel_ms, record_stride, fma_peak, peaks, sm_max_mhz, traffic_tag=None):
    """FP32 CUDA-core bound (FLOP model v1) or, for recorded trajectories, the HBM write-back bound —
    whichever fraction is larger is the binding roofline."""
    fpe = w["flops"][integ if integ in w["flops"] else "euler"]
    per_gpu = float(envs) * T / (kernel_ms * 1e-3)
    achieved = per_gpu * fpe / 1e12
    nominal = 148 * 128 * 2 * (sm_max_mhz or peaks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    rec_bytes = (n + m) * 4.0 / record_stride if record_stride > 0 else 0.0
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic, traffic_src = ncu_traffic(traffic_tag, "rollout") if traffic_tag else (None, "not profiled")
    cost_flops = 2 * n + 3 * m       # Q = I, R = I form: one FFMA per state component, FADD + FFMA per input
    roof = {"bound": "fp32", "achieved": achieved, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved / fma_peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "measured: hjb_fma_peak_probe FFMA-only kernel in this run (FP32 CUDA-core peak is not in "
                           "MEASURED_PEAKS.json)",
            "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
            "flops_per_env_step": fpe, "flop_model": "SURVEY.md 8d v1 (sin / cos / tan = 1 flop each, cost not counted)",
            # FLOP model v1 leaves the cost accumulation out ("add 2n + 2m + ... when the cost output is requested"); the
            # timed kernel does accumulate it: the fraction with those flops counted is given beside the headline one
            "cost_flops_per_env_step": cost_flops,
            "frac_with_cost_flops": achieved * (fpe + cost_flops) / fpe / fma_peak,
            # algorithmic HBM bytes: x0 read + x_final / cost write per env, plus the recorded trajectory
            "hbm": {"achieved_gbs": (per_gpu * rec_bytes + envs * (2 * n + 1) * 4 / (kernel_ms * 1e-3)) / 1e9,
                    "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}
    if roof["hbm"]["achieved_gbs"] / hbm_peak > roof["frac"]:
        roof = {"bound": "hbm", "achieved": roof["hbm"]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": roof["hbm"]["achieved_gbs"] / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": roof["hbm"]["peak_source"], "bytes_per_env_step": rec_bytes,
                "fp32": {k: roof[k] for k in ("achieved", "peak", "frac", "flops_per_env_step")}}
    return roof


def measure_rollout(wname, integ, envs, T, record_stride, steps, warmup, rank, world, fma_peak, peaks, fast=True,
                    first=0, seed=None, traffic_tag=None, want_parity=False):
    """One rollout configuration, inputs resident in HBM (generated there): aggregate env-steps/s over all ranks (no
    collective on the data path), the kernel's mean launch time on this rank, its roofline.  Used for the secondary lines."""
    import torch
    import torch.distributed as dist
    from q_learning_with_hjb_b200.workloads import make_controller, make_dynamics
    from q_learning_with_hjb_b200.rollout import BatchedRollout, RunningCost
    w = ROLLOUTS[wname]
    dyn = make_dynamics(w["sys"])
    dyn.fast_trig = fast
    ctl = make_controller(w["ctl"], dyn)
    n, m = dyn.get_dimension()
    cost = RunningCost(np.eye(n), np.eye(m), getattr(ctl, "xf", np.zeros(n)), getattr(ctl, "uf", np.zeros(m)))
    plan = BatchedRollout(dyn, ctl, envs, T, integrator=integ, record_stride=record_stride, cost=cost)
    x0 = synthetic_x0_device(dyn, w["sys"], envs, (1234 + rank) if seed is None else seed, first)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    # warm-up: at least 3 launches AND ~50 ms of GPU work (the previous configuration's oracle check ran on the CPU for
    # seconds: the clocks of an idle GPU ramp up again over the first milliseconds — visible on the 50 us launches of C1)
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record()
    for _ in range(max(warmup, 3)):
        plan.launch(x0)
    w1.record()
    torch.cuda.synchronize()
    per = max(w0.elapsed_time(w1) / max(warmup, 3), 1e-3)
    for _ in range(int(min(2000, 50.0 / per))):
        plan.launch(x0)
    if envs <= (1 << 16):
        steps = max(steps, 20)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for a, b in evs:
        a.record(); plan.launch(x0); b.record()
    t1.record()
    barrier()
    ms = torch.tensor([t0.elapsed_time(t1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    d = {"workload": w["label"], "integrator": integ, "envs_per_gpu": envs, "horizon": T,
         "record": "final state + per-env cost" if record_stride == 0 else f"every {record_stride} steps",
         "value": float(envs) * T * steps * world / (float(ms.item()) * 1e-3), "unit": "env-steps/s", "steps": steps,
         "kernel_ms": kernel_ms, "kernel_variant": plan.kernel_variant(),
         "roofline": rollout_roofline(w, integ, n, m, envs, T, kernel_ms, record_stride, fma_peak, peaks, None, traffic_tag)}
    if want_parity and rank == 0:
        d["parity"] = rollout_parity(w, integ, fast=fast, envs=min(4096, envs), seed=1234, record_stride=0)
    del plan, x0
    torch.cuda.empty_cache()
    return d


def verify_rollout_sharding(rank, world):
    """N-GPU == 1-GPU for rollouts: every rank integrates its shard of ONE global batch (environment i = sample i of the
    stream keyed by one seed); rank 0 then integrates the whole batch alone and compares bit for bit."""
    import torch
    import torch.distributed as dist
    from q_learning_with_hjb_b200.workloads import make_controller, make_dynamics
    from q_learning_with_hjb_b200 import parallel
    N, T = 1 << 18, 200
    dyn = make_dynamics("quad2d")
    ctl = make_controller("quad2d_hover", dyn)
    lo, hi = parallel.shard_bounds(N, rank, world)
    mine = dyn.rollout(ctl, synthetic_x0_device(dyn, "quad2d", hi - lo, 4321, lo), T, record_stride=0).x_final
    parts = [torch.empty((parallel.shard_bounds(N, r, world)[1] - parallel.shard_bounds(N, r, world)[0], 6), device="cuda")
             for r in range(world)]
    dist.all_gather(parts, mine.contiguous())
    if rank != 0:
        return None
    whole = dyn.rollout(ctl, synthetic_x0_device(dyn, "quad2d", N, 4321, 0), T, record_stride=0).x_final
    return {"what": f"{N} envs x {T} steps of C4's system: shards of one seeded global batch on {world} GPUs against the "
                    "whole batch on GPU 0", "bit_identical": bool(torch.equal(torch.cat(parts), whole))}


def run_rollout(args, w, integ):
    rank, world, local = dist_setup(args.gpus)
    if args.impl == "reference":
        run_reference_arm(args, w, integ, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path is CUDA-only)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from q_learning_with_hjb_b200.workloads import make_controller, make_dynamics
    from q_learning_with_hjb_b200.rollout import BatchedRollout, RunningCost

    envs = args.envs or w["envs"]
    T = args.horizon or w["T"]
    fast = not args.accurate_trig
    dyn = make_dynamics(w["sys"])
    dyn.fast_trig = fast
    ctl = make_controller(w["ctl"], dyn)
    n, m = dyn.get_dimension()
    cost = RunningCost(np.eye(n), np.eye(m), getattr(ctl, "xf", np.zeros(n)), getattr(ctl, "uf", np.zeros(m)))
    plan = BatchedRollout(dyn, ctl, envs, T, integrator=integ, record_stride=args.record_stride, cost=cost)

    # synthetic inputs: generated ON THE DEVICE by the counter-based generator, seed = 1234 + rank; resident
    # in HBM for `value`.  The end-to-end legs start from the same numbers in pinned host memory.
    x0_dev = synthetic_x0_device(dyn, w["sys"], envs, 1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = load_peaks()
    # ---- FP32 FMA peak probe (roofline denominator for the CUDA-core bound) ----
    fma_peak_tflops = measure_fma_peak()

    # ---- device-resident timing (`value`) ----
    clocks = ClockSampler(local); clocks.start()
    for _ in range(max(args.warmup, 3)):
        plan.launch(x0_dev)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_begin = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    t_begin.record()
    for a, b in evs:
        a.record(); plan.launch(x0_dev); b.record()
    t_end.record()
    barrier()
    clk = clocks.stop(wall0, time.time())
    total_ms = t_begin.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))

    # ---- end to end through the public API with HOST buffers (`e2e`): pinned x0 -> H2D -> rollout -> D2H of the per-
    # environment cost (the result of the closed-loop evaluation: what scripts/test_vhjb_policy.py:146-151 accumulates).
    # Cost-only plan: the final states stay on the device (round 1 brought them back too: 403 MB more per step). ----
    e2e_plan = plan if args.record_stride > 0 else BatchedRollout(dyn, ctl, envs, T, integrator=integ, record_stride=0,
                                                                  cost=cost, want_final=False)
    pinned = e2e_plan.pinned_x0()
    pinned.copy_(x0_dev)
    torch.cuda.synchronize()

    def timed_e2e(fn):
        for _ in range(2):
            fn()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            out = fn()
        t1.record()
        barrier()
        return t0.elapsed_time(t1), out
    e2e_ms, (_, cost_host) = timed_e2e(lambda: e2e_plan.run_host(pinned, copy_in=False))
    checksum = float(cost_host.double().mean())
    h2d_bytes, d2h_bytes = e2e_plan.h2d_bytes(), e2e_plan.d2h_bytes()
    # the same call with a SEED as its input: x0 generated on the device range by range (nothing crosses PCIe on the way in)
    seeded_ms, (_, cost_seeded) = timed_e2e(lambda: e2e_plan.run_seeded(1234 + rank))
    seeded_same = bool(torch.equal(cost_seeded, cost_host))

    times = torch.tensor([total_ms, e2e_ms, seeded_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, seeded_ms = times.tolist()

    # ---- untimed: the timed instantiation against the oracle on the first 4096 environments of the same x0 ----
    parity = None
    if rank == 0 and not args.no_parity:
        parity = rollout_parity(w, integ, fast=fast, envs=min(4096, envs), seed=1234 + rank,
                                record_stride=args.record_stride)
        parity["checksum_e2e_mean_cost_all_envs"] = checksum
    verify = verify_rollout_sharding(rank, world) if (world > 1 and not args.no_verify) else None
    del e2e_plan, plan, x0_dev, pinned
    torch.cuda.empty_cache()

    # ---- secondary configurations, each with its own roofline (driver-visible: C4 RK4, C3, C1, recorded trajectories) ----
    extra = None
    if args.workload == DEFAULT_WORKLOAD and not args.no_extra and integ == "euler" and args.record_stride == 0 \
            and args.envs == 0 and args.horizon == 0:
        es = min(args.steps, 3)
        mk = lambda *a, **k: measure_rollout(*a, steps=es, warmup=3, rank=rank, world=world, fma_peak=fma_peak_tflops,
                                             peaks=peaks, fast=fast, **k)
        extra = {
            "C4_rk4": mk("quad2d_hover", "rk4", 1 << 24, 1000, 0, traffic_tag="r02_rollout_quad2d_rk4", want_parity=True),
            "C3_acrobot_euler": mk("acrobot_es", "euler", 1 << 22, 2000, 0, traffic_tag="r02_rollout_acrobot_euler",
                                   want_parity=True),
            "C1_cartpole_euler": mk("cartpole_lqr", "euler", 4096, 500, 0, want_parity=True),
            "C1_cartpole_rk4": mk("cartpole_lqr", "rk4", 4096, 500, 0, want_parity=True),
            "cartpole_16M_euler": mk("cartpole_lqr_big", "euler", 1 << 24, 500, 0),
            "quad10d_hover_euler": mk("quad10d_hover", "euler", 1 << 23, 1000, 0, want_parity=True),
            # recorded trajectories: the HBM write-back is the bound (quad-10D rows go through the staged TMA bulk stores)
            "quad10d_record_every_step": mk("quad10d_hover", "euler", 1 << 18, 500, 1),
            "C4_record_every_step": mk("quad2d_hover", "euler", 1 << 19, 500, 1),
        }
        if world > 1:   # strong scaling: C3's 4M environments split over the ranks (shards of one seeded global batch)
            per = (1 << 22) // world
            extra["strong_C3_4M_envs_total"] = mk("acrobot_es", "euler", per, 2000, 0, seed=1234, first=rank * per)
            extra["strong_C3_4M_envs_total"]["scaling"] = "strong"

    # ---- the metric's second half: HJB-residual states/s (C5), measured in the same run ----
    vhjb = None
    if args.workload == DEFAULT_WORKLOAD and not args.no_vhjb:
        vhjb = measure_vhjb(VHJB["vhjb_quad10d"], args.steps, args.warmup, rank, world, local,
                            want_cpu=(world == 1 and not args.no_cpu_baseline), extras=not args.no_extra,
                            verify=not args.no_verify)
        if vhjb is not None and vhjb["roofline"]["bound"] == "fp32":
            vhjb["roofline"]["peak"] = fma_peak_tflops
            vhjb["roofline"]["frac"] = vhjb["roofline"]["achieved"] / fma_peak_tflops
    units = float(envs) * T * args.steps * world
    value = units / (total_ms * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)

    if rank == 0:
        roof = rollout_roofline(w, integ, n, m, envs, T, kernel_ms, args.record_stride, fma_peak_tflops, peaks,
                                clk.get("sm_max_mhz"),
                                traffic_tag=(f"r02_rollout_quad2d_{integ}" if (args.workload == DEFAULT_WORKLOAD and args.envs == 0
                                                                              and args.horizon == 0 and args.record_stride == 0)
                                             else None))
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_rollout(w, integ, T)
        line = {
            "metric": "closed-loop env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["label"], "envs_per_gpu": envs, "horizon": T, "integrator": integ,
                       "record": "final state + per-env cost" if args.record_stride == 0 else f"every {args.record_stride} steps",
                       "trig": "libdevice" if args.accurate_trig else "table + Taylor sin/cos (7e-8) + MUFU.RCP (fast_trig, the default)",
                       "parallelism": f"env-shard x{world}",
                       "l2": "inputs larger than L2 (x0 >= 400 MB per launch)" if envs * n * 4 > 126e6 else "small workload",
                       "seed": "1234 + rank (device-side Philox4x32-10 stream, hjb_sample_states)"},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms / args.steps, "checksum_mean_cost": checksum,
                    "result": "per-environment closed-loop cost (pinned host memory)",
                    "seeded": {"value": units / (seeded_ms * 1e-3), "ms_per_step": seeded_ms / args.steps,
                               "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h_bytes,
                               "what": "the same call with a seed as input: x0 generated on the device",
                               "same_costs_as_host_x0": seeded_same}},
            "gpu_launches": args.steps,
            "kernel_ms": kernel_ms,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clk, "parity": parity, "verify": verify,
            "extra": extra,
            "vhjb": vhjb,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
# vhjb: HJB-residual + value-net training step on sampled states
# ---------------------------------------------------------------------------------------------------------------
VHJB = {
    "vhjb_quad10d": dict(problem="quad10d", states=1 << 20,
                         label="C5: 10-D quadcopter vhjb value learning, 1,048,576 states per GPU per iteration "
                               "(8M/iter on 8 GPUs), residual + loss-gradient + grad all-reduce + Adam"),
    "vhjb_di": dict(problem="di_mintime", states=1 << 20,
                    label="C2: double integrator minimum-time vhjb (sin net, bang-bang), 1M sampled states, "
                          "residual + loss-gradient + Adam"),
}
# logical fp32 flops per state: 17 GEMMs of the full pass / 6 GEMMs of the residual-only pass
VHJB_FLOPS_FULL = lambda n: 1280 * n + 294912
VHJB_FLOPS_RES = lambda n: 512 * n + 98304


def _cpu_vhjb_worker(args):
    name, B, threads, full = args
    import torch
    torch.set_num_threads(threads)
    from oracle import vhjb_oracle as V
    from tests.helpers_vhjb import problem, sample_batch
    V.set_dtype(torch.float32)          # the reference's JAX code computes in float32
    p = problem(name)
    orc = V.VhjbOracle(p, V.init_weights(p.sys.n, seed=0))
    xs, dones, costs = sample_batch(name, B, seed=7)
    t0 = time.perf_counter()
    if full:
        orc.loss_and_grad(xs, dones, costs, 1e-5)
    else:
        orc.losses(xs, dones, costs)
    return time.perf_counter() - t0


def cpu_vhjb_throughput(name, B, full=True):
    """states/s of the torch-CPU restatement (oracle/vhjb_oracle.py, float32, all host threads)."""
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(1) as pool:
        pool.map(_cpu_vhjb_worker, [(name, max(1024, B // 8), cores, full)])      # warm-up
        dt = pool.map(_cpu_vhjb_worker, [(name, B, cores, full)])[0]
    return B / dt, dt, cores


def verify_vhjb_sharding(rank, world):
    """N-GPU == 1-GPU on hardware (plan item 5): three train steps on ONE global batch split over the
    ranks (NCCL all-reduces, the next step's done-counts riding on the gradient all-reduce) against three steps of the
    single-GPU path on the concatenated batch; every rank's weights must be identical, and equal to the single-GPU
    weights up to the order of the fp32 partial sums."""
    import torch
    import torch.distributed as dist
    from q_learning_with_hjb_b200 import parallel
    from q_learning_with_hjb_b200 import workloads as WL
    from q_learning_with_hjb_b200.controller.vhjb import AdamState
    B = 64 * 148 * 8 + 1000                      # ragged: shard sizes differ
    xs, dones, costs = (torch.as_tensor(a).cuda() for a in WL.sample_vhjb_batch("quad10d", B, seed=77))
    lo, hi = parallel.shard_bounds(B, rank, world)
    W0 = torch.as_tensor(WL.flat_params(WL.init_weights(10, seed=3))).cuda()

    def run(local_only):
        k, _ = WL.make_vhjb_kernels("quad10d")
        w = W0.clone()
        opt = AdamState(0, torch.zeros_like(w), torch.zeros_like(w))
        sl = slice(0, B) if local_only else slice(lo, hi)
        x, d, c = xs[sl].contiguous(), dones[sl].contiguous(), costs[sl].contiguous()
        losses = []
        for i in range(3):
            sums, norm = k.train_step(w, opt, x, d, c, 0.25, 1e-3, local=local_only, next_dones=None if local_only else d)
            losses.append((sums / norm).tolist())
        return w, losses
    w_dist, l_dist = run(False)
    gathered = [torch.empty_like(w_dist) for _ in range(world)]
    dist.all_gather(gathered, w_dist)
    cross = max(float((g - gathered[0]).abs().max()) for g in gathered)
    w_one, l_one = run(True)
    rel = float((w_dist - w_one).abs().max() / w_one.abs().max())
    moved = float((w_one - W0).abs().max())
    lrel = max(abs(a - b) / abs(b) for la, lb in zip(l_dist, l_one) for a, b in zip(la, lb))
    t = torch.tensor([cross, rel, lrel], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cross, rel, lrel = t.tolist()
    return {"what": f"3 train steps, {B} states (C5's problem) split over {world} GPUs against the single-GPU path on the "
                    "whole batch", "max_abs_diff_between_ranks": cross, "max_rel_diff_vs_single_gpu_weights": rel,
            "max_rel_diff_vs_single_gpu_losses": lrel, "weights_moved_by": moved, "tolerance": 1e-6,
            "ok": bool(cross == 0.0 and rel <= 1e-6 and lrel <= 1e-5)}


def measure_vhjb(w, steps, warmup, rank, world, local, want_cpu, extras=False, verify=False):
    """Returns the dict describing the vhjb measurement (rank 0) — used for the vhjb workloads' own JSON line and as
    the secondary measurement attached to the default line."""
    import torch
    import torch.distributed as dist
    from q_learning_with_hjb_b200 import workloads as WL       # (product-side: nothing of oracle/ or tests/ on this arm)
    from q_learning_with_hjb_b200.controller.vhjb import AdamState

    B = w["states"]
    k, p = WL.make_vhjb_kernels(w["problem"])
    n = len(p.xf)
    W = WL.init_weights(n, seed=0)
    params = torch.as_tensor(WL.flat_params(W)).cuda()
    opt = AdamState(0, torch.zeros_like(params), torch.zeros_like(params))
    xs, dones, costs = WL.sample_vhjb_batch(w["problem"], B, seed=1234 + rank)
    host = [torch.as_tensor(a).pin_memory() for a in (xs, dones, costs)]
    dev = [h.cuda() for h in host]
    out_host = torch.empty(4, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def train():
        # (several GPUs: the next step's batch is known — its done-counts ride on this step's gradient all-reduce)
        k.train_step(params, opt, dev[0], dev[1], dev[2], 1e-5, 1e-3, next_dones=dev[1] if world > 1 else None)

    def residual_only():
        k.residual(params, dev[0], dev[1], dev[2], want=())

    def e2e():
        # public host-batch call: H2D of the pinned batch pipelined under the kernel (VhjbKernels.train_step_host)
        sums, norm = k.train_step_host(params, opt, host[0], host[1], host[2], 1e-5, 1e-3)
        out_host[:2].copy_(sums, non_blocking=True)
        out_host[2:].copy_(norm, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def grad_kernel_only():
        k.loss_grad(params, dev[0], dev[1], dev[2], 1e-5)

    tensor_path = os.environ.get("HJB_VHJB_IMPL", "") != "simt"
    for _ in range(max(warmup, 3)):
        train(); residual_only()
    # clocks: the K timed steps last a few tens of ms, too short for nvidia-smi's sampling; the same train step is
    # therefore also run for ~1 s with the sampler on (untimed), and the timed steps follow immediately
    # The number of load-loop steps is agreed between the ranks BEFORE the loop (every train step holds two
    # all-reduces: a per-rank time-based loop would issue different numbers of collectives and deadlock).
    est_ms = timed(train, 3) / 3.0                                                             # max over ranks
    n_load = 0 if os.environ.get("HJB_BENCH_NO_CLOCK_LOOP") else int(min(2000, max(20, 1000.0 / max(est_ms, 1e-3))))
    clocks = ClockSampler(local); clocks.start()
    wall0 = time.time()
    for i in range(n_load):                                                                    # (skipped under ncu)
        train()
        if i % 20 == 19:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    train_ms = timed(train, steps)
    clk = clocks.stop(wall0, time.time())
    res_ms = timed(residual_only, steps)
    for _ in range(2):           # (first call: the three reduction kernels of this entry point are loaded lazily)
        grad_kernel_only()
    grad_ms = timed(grad_kernel_only, steps)
    for _ in range(2):
        e2e()
    e2e_ms = timed(e2e, steps)
    k.check_streams()
    h2d_bytes = int(sum(h.numel() * 4 for h in host))
    ver = verify_vhjb_sharding(rank, world) if (verify and world > 1) else None
    sub = None
    if extras:
        del dev, host, params, opt
        torch.cuda.empty_cache()
        sub = {}
        if w["problem"] == "quad10d":
            # C2 (double integrator, sin net, bang-bang, min-time residual) and strong scaling of C5's 8M states per iteration
            sub["C2_double_integrator"] = measure_vhjb(VHJB["vhjb_di"], min(steps, 5), warmup, rank, world, local, False)
            total = 1 << 23
            sub["strong_C5_8M_states_total"] = measure_vhjb(dict(VHJB["vhjb_quad10d"], states=total // world,
                                                                 label=f"C5 strong scaling: {total} states per iteration "
                                                                       f"split over {world} GPU(s)"),
                                                            min(steps, 5), warmup, rank, world, local, False)
            if sub["strong_C5_8M_states_total"] is not None:
                sub["strong_C5_8M_states_total"]["scaling"] = "strong"
    if rank != 0:
        return None
    peaks = load_peaks()
    train_rate = B * world * steps / (train_ms * 1e-3)
    res_rate = B * world * steps / (res_ms * 1e-3)
    e2e_rate = B * world * steps / (e2e_ms * 1e-3)
    kern_rate = B * steps / (grad_ms * 1e-3)          # per GPU, the fused loss+gradient kernel (+ its 3 tiny reduces)
    logical_tflops = kern_rate * VHJB_FLOPS_FULL(n) / 1e12
    peer = world > 1 and getattr(k, "_peer", None) is not None
    exchange = ("none (single GPU)" if world == 1 else
                "peer memory over NVLink inside the reduce kernel (hjb_vhjb_train_step_peer): reduce + exchange + Adam in one "
                "launch, no NCCL call in the step" if peer else
                "NCCL: one all-reduce of gradient + loss sums + next normalisers per step")
    d = {
        "metric": "HJB-residual states/s (residual + loss-gradient + Adam train step)", "value": train_rate,
        "unit": "states/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": train_ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["label"], "states_per_gpu": B, "value_net": [n, 128, 128, 64],
                   "activation": p.act, "parallelism": f"state-shard x{world}, grad all-reduce" if world > 1 else "single GPU",
                   "exchange": exchange,
                   "l2": "states (>= 40 MB) streamed once per step; weights resident in shared memory", "seed": "1234 + rank",
                   "kernel": "tcgen05 fp16x3 (vhjb_tc.cuh)" if tensor_path else "CUDA-core fp32 (vhjb_simt.cuh)"},
        "residual_only_states_per_s": res_rate,
        "e2e": {"value": e2e_rate, "unit": "states/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 16, "ms_per_step": e2e_ms / steps},
        # kernels of this library per step (profiles/r02_vhjb_quad10d_tc.md lists them).  Single process: count x2, fused
        # tensor pass, deferred fp32 pass, reduce-and-Adam (hjb_vhjb_train_step); multi-GPU over peer memory: count x2 of the
        # next batch, the two passes, reduce-exchange-Adam; multi-GPU over NCCL: the two passes, reduce x3, count x2 of the
        # next batch, Adam (the NCCL all-reduce — ONE per step — is not counted)
        "gpu_launches": steps * (5 if world == 1 or peer else 8),
        "collectives_per_step": 0 if world == 1 or peer else 1,
        "verify": ver, "extra": sub,
        "kernel_ms": grad_ms / steps,
        "clocks": clk,
    }
    if tensor_path:
        # the 12 GEMMs without an n-dimension are 294 912 logical flops/state (full pass), 98 304
        # (residual only); each is EXECUTED as 3 fp16 products (hi hi + lo hi + hi lo).  The n-wide GEMMs (padded to
        # 16) also run on the tensor pipe but are not counted.
        executed = 3 * kern_rate * 294912 / 1e12
        executed_res = 3 * (B * steps / (res_ms * 1e-3)) * 98304 / 1e12
        peak = peaks.get("bf16_tflops", 1590.0)
        d["roofline"] = {"bound": "tensor", "achieved": executed, "unit": "TFLOP/s", "peak": peak, "frac": executed / peak,
                         "traffic": (ncu_traffic({"quad10d": "r02_vhjb_quad10d_tc", "di_mintime": "r02_vhjb_di_tc_sin"}.get(w["problem"], ""), "vhjb")[0]
                                     if B == 1 << 20 else None),
                         "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst; kind::f16 fp16/bf16 share the rate)"
                                         if "bf16_tflops" in peaks else "fallback 1590 TFLOP/s (B200_PROFILING.md)"),
                         "peak_sustained": peaks.get("bf16_tflops_sustained"),
                         "frac_of_sustained": (executed / peaks["bf16_tflops_sustained"]) if peaks.get("bf16_tflops_sustained") else None,
                         "kernel": "vhjb_tc_kernel<GRAD> (fused loss + gradient), timed alone with CUDA events",
                         "split_products": 3, "tensor_flops_per_state_logical": 294912,
                         "logical_tflops_all_17_gemms": logical_tflops, "flops_per_state": VHJB_FLOPS_FULL(n),
                         "residual_only": {"achieved": executed_res, "frac": executed_res / peak,
                                           "tensor_flops_per_state_logical": 98304},
                         "note": "SS-mode tcgen05.mma with M=128, N=64 costs 48 cycles (32 for the 4 KB A tile + 16 for B from "
                                 "shared memory) against a 32-cycle math floor: 67 % is the ceiling of this tile shape"}
    else:
        d["roofline"] = {"bound": "fp32", "achieved": logical_tflops, "unit": "TFLOP/s", "peak": None, "frac": None,
                         "traffic": None, "flops_per_state": VHJB_FLOPS_FULL(n),
                         "kernel": "vhjb_kernel (CUDA-core fp32 version)"}
    if want_cpu:   # B = 1,048,576, full train step and residual only
        v, dt, cores = cpu_vhjb_throughput(w["problem"], 1 << 20, full=True)
        vr, dtr, _ = cpu_vhjb_throughput(w["problem"], 1 << 20, full=False)
        d["cpu_baseline"] = {"value": v, "unit": "states/s", "cores": cores, "kind": "port",
                             "residual_only": vr,
                             "sample": f"1048576 states, torch-CPU float32 restatement of controller/vhjb.py "
                                       f"(oracle/vhjb_oracle.py; JAX is not installed, the reference's own vhjb code cannot "
                                       f"run), {cores} threads, {dt:.1f} s (full step) + {dtr:.1f} s (residual only)"}
    return d


def run_vhjb(args, w):
    rank, world, local = dist_setup(args.gpus)
    if args.impl == "reference":
        if rank != 0:
            return
        vals, t_all = [], 0.0
        Bs = 1 << 20
        for _ in range(min(args.steps, 3)):
            v, dt, cores = cpu_vhjb_throughput(w["problem"], Bs, full=True)
            vals.append(v); t_all += dt
        value = float(np.mean(vals))
        sample = f"{Bs} states per bench step, torch-CPU float32 restatement of controller/vhjb.py (oracle/vhjb_oracle.py), {cores} threads"
        print(json.dumps({
            "impl": "reference", "metric": "HJB-residual states/s (residual + loss-gradient + Adam train step)",
            "value": value, "unit": "states/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the b200 arm's config keys (same workload), plus what the bounded sample of one step was
            "config": {"workload": w["label"], "states_per_gpu": w["states"], "value_net": "[n, 128, 128, 64]",
                       "parallelism": f"{cores} host threads (torch)", "seed": "1234 + rank", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "states/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), flush=True)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path is CUDA-only)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = measure_vhjb(w, args.steps, args.warmup, rank, world, local, want_cpu=(world == 1 and not args.no_cpu_baseline),
                     verify=not args.no_verify)
    if rank == 0:
        if d["roofline"]["bound"] == "fp32":
            fma = measure_fma_peak()
            d["roofline"]["peak"] = fma
            d["roofline"]["frac"] = d["roofline"]["achieved"] / fma
            d["roofline"]["peak_source"] = "measured: hjb_fma_peak_probe FFMA-only kernel in this run"
        print(json.dumps(d), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_fma_peak():
    """FP32 FMA peak (TFLOP/s) of this GPU, measured with the library's FFMA-only probe kernel."""
    import ctypes as C
    import torch
    from q_learning_with_hjb_b200 import _lib as L
    sink = torch.empty(148 * 8 * 256 * 2, device="cuda", dtype=torch.float32)
    flops = C.c_double(0)
    # a PEAK: the best of several probe launches after a warm-up long enough for the clocks to ramp (a single cold
    # launch was seen to read 61 instead of 71 TFLOP/s, which flatters every fraction reported against it)
    for _ in range(6):
        L.check(L.lib().hjb_fma_peak_probe(L.ptr(sink), sink.numel(), 20000, C.byref(flops), L.stream_ptr()))
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib().hjb_fma_peak_probe(L.ptr(sink), sink.numel(), 20000, C.byref(flops), L.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--integrator", default="euler", choices=["euler", "rk4"])
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (default: the workload's)")
    ap.add_argument("--horizon", type=int, default=0)
    ap.add_argument("--record-stride", type=int, default=0)
    ap.add_argument("--accurate-trig", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed oracle check of the timed instantiation")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configurations of the default line")
    ap.add_argument("--no-verify", action="store_true", help="skip the N-GPU == 1-GPU checks (N > 1)")
    ap.add_argument("--no-vhjb", action="store_true", help="skip the secondary vhjb measurement of the default line")
    ap.add_argument("--watchdog", type=int, default=900,
                    help="seconds after which a hung run dumps every thread's stack to stderr and exits (0 = off)")
    args = ap.parse_args()
    if args.watchdog > 0:   # a deadlocked collective must not hang the box until the driver's own limit
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    if args.workload in ROLLOUTS:
        run_rollout(args, ROLLOUTS[args.workload], args.integrator)
    elif args.workload in VHJB:
        run_vhjb(args, VHJB[args.workload])
    else:
        raise SystemExit(f"unknown workload {args.workload}; choose from {sorted(ROLLOUTS) + sorted(VHJB)}")


if __name__ == "__main__":
    # The contract is ONE JSON line on stdout.  Libraries write banners to file descriptor 1 (NCCL prints its version
    # there): keep a private copy of the real stdout for Python's print and point fd 1 at stderr for everything else.
    _real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real, "w", buffering=1)
    main()
