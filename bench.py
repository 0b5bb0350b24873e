#!/usr/bin/env python
"""Headline benchmark: ellipsoid-reachability rollouts/sec (B x H onestep calls) on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3                 # product arm (libsegp.so, CUDA)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1  # CPU arm: the reference algorithm's port
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus 8 --steps 5 --warmup 3

Workload (BASELINE.json): config C4 = cart-pole, GP N=5000, H=20, n_s=4, B=65536 candidate sequences -- the
configuration the metric is quoted on.  It fits one GPU (8 workspace chunks of 8192), so the default N=1 line runs ALL
65536 candidates on the one GPU and N GPUs shard them (`--scaling strong`, B_g = 65536 / N: at N=8 exactly BASELINE
C4).  `--scaling weak` keeps the 8192-candidate shard per GPU instead (B = 8192 x N; round 1's lines).
A "step" is one pass of the hot path over the rank's batch: B_g independent H-step rollouts
(multistep_reachability) = B_g x H one-step calls.  `value` times the device-resident path (inputs in HBM,
CUDA events); `e2e` times the same call through the host-buffer C-ABI entry (segp_multistep_host: pinned host
inputs, H2D + D2H inside the timed region) over the same number of steps.  Every rank prints nothing except rank 0's
single JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ellipsoid_reachability_rollouts_per_sec"
UNIT = "rollouts/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        d["_source"] = "MEASURED_PEAKS.json (of measured)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "_source": "B200_PROFILING.md fallback (of fallback)"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

    def summary(self, t0, t1):
        sm, smax, power, reasons = [], [], [], set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "reasons": sorted(reasons), "samples": len(sm)}


def _oracle_model(w):
    from oracle import gp_oracle
    noise = [h["noise"] + 1e-5 + 1e-8 for h in w.hyp]
    if any(k.startswith("lin_") for k in w.kern_types):
        ls, var, pl, lin = gp_oracle.vectors_from_reference_hyp(w.kern_types, w.hyp, w.n_s + w.n_u, semantics="casadi")
        return gp_oracle.GPOracle(w.x_train, w.y_train, w.kern_types, ls, var, noise, prod_linear=pl, linear=lin)
    return gp_oracle.GPOracle(w.x_train, w.y_train, w.kern_types, np.stack([h["lengthscale"] for h in w.hyp]),
                              [h["variance"] for h in w.hyp], noise)


def cpu_baseline(w, seconds_target=15.0):
    """The vectorised float64 oracle (oracle/reach_oracle.multistep_batch: Cholesky + solve_triangular on N x B
    panels, all host cores through BLAS) on a bounded sample of the same workload."""
    from oracle import reach_oracle
    ora = _oracle_model(w)
    cores = os.cpu_count() or 1
    sample = 8
    t_used = 0.0
    best = None
    while True:
        k_ff = w.k_ff[:sample]
        t0 = time.perf_counter()
        out = reach_oracle.multistep_batch(w.p0, ora, w.k_fb, k_ff, w.l_mu, w.l_sigma, None, w.c_safety, w.a, w.b)
        dt = time.perf_counter() - t0
        t_used += dt
        best = (sample, dt)
        if t_used > seconds_target or sample * 4 > w.k_ff.shape[0] or dt * 4 > seconds_target:
            break
        sample *= 4
    return {"value": best[0] / best[1], "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "{} of the workload's rollouts (H={}), vectorised float64 NumPy/SciPy oracle "
                      "(Cholesky form), {:.2f} s".format(best[0], w.horizon, best[1])}, out


def parity_against_oracle(res, oracle_out):
    """The oracle rollouts of the cpu_baseline leg are the FIRST candidates of rank 0's shard: compare the GPU result
    of the timed arm with them at the workload's full model size (rtol gate of BASELINE.json: 1e-4)."""
    p_o, q_o, v_o = oracle_out
    n = p_o.shape[0]

    def err(got, want, atol_scale):
        got = got[:n].cpu().numpy()
        scale = atol_scale * max(1.0, float(np.max(np.abs(want))))
        return float(np.max(np.abs(got - want) / (np.abs(want) + scale / 1e-4)))

    pairs = (("p_all", res.p_all, p_o), ("q_all", res.q_all, q_o), ("var_all", res.var_all, v_o))
    # strict: absolute floor 1e-9 x scale (round 1's figure: entries down to 1e-5 of the largest count fully);
    # gate: SURVEY.md section 8d, np.allclose(rtol = 1e-4, atol = 1e-6 x scale)
    e = {k: err(g, o, 1e-9) for k, g, o in pairs}
    e_gate = {k: err(g, o, 1e-6) for k, g, o in pairs}
    return {"checked_rollouts": int(n), "against": "float64 oracle (oracle/reach_oracle.multistep_batch) at full model size",
            "max_rel_err": e, "max_rel_err_gate_atol_1e-6_scale": e_gate, "rtol_gate": 1e-4,
            "ok": bool(max(e_gate.values()) < 1e-4), "ok_strict": bool(max(e.values()) < 1e-4)}


def _set_blas_threads(n):
    """torchrun exports OMP_NUM_THREADS=1 to its workers: give the CPU arm the host cores it is entitled to."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:   # pragma: no cover
        pass
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(n)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, rank 0 only (the
    other ranks exit without work).  With oracle/_ref in place (oracle/build_ref.py: the reference's unmodified
    gp_reachability.py / utils.py / utils_ellipsoid.py / gp_models_utils_casadi.py) the timed code IS the reference:
    multistep_reachability (gp_reachability.py:159-212) over SimpleGPModel.__call__'s arithmetic (gp_pred,
    gp_models_utils_casadi.py:177-197), one trajectory at a time as the reference runs it; only the posterior state
    (GPy) and the mean Jacobian (CasADi AD) are restated (oracle/ref_ssm.py).  Without it: the oracle port."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _set_blas_threads(cores)
    from oracle import reach_oracle, ref_loader
    from safe_exploration_b200 import workloads
    per_step = max(1, args.ref_rollouts)
    w = workloads.make(args.config, batch=per_step * (args.steps + args.warmup), kern=args.kern or None)
    ora = _oracle_model(w)
    ora._ensure_inv()
    kind = "port"
    multistep = reach_oracle.multistep_reachability
    ssm = ora
    how = "oracle port of multistep_reachability + SimpleGPModel.__call__ (explicit-inverse form)"
    if ref_loader.available():
        from oracle.ref_ssm import ReferenceGP
        ref_reach, _, _ = ref_loader.load()
        multistep = ref_reach.multistep_reachability
        ssm = ReferenceGP(ora, w.hyp, w.kern_types)
        kind = "reference"
        how = ("the reference's own multistep_reachability + gp_pred, unmodified files from {} (posterior state and "
               "mean Jacobian restated: GPy / CasADi absent)".format(ref_loader.origin()))

    def one_step(i):
        for j in range(per_step):
            multistep(w.p0[:, None], ssm, w.k_fb, w.k_ff[i * per_step + j], w.l_mu, w.l_sigma, None, w.c_safety, 0,
                      w.a, w.b)

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")     # the reference leaks complex128 from scipy.linalg.eig (utils.py:133-141)
        for i in range(args.warmup):
            one_step(i)
        t0 = time.perf_counter()
        for i in range(args.steps):
            one_step(args.warmup + i)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = "{} rollouts per step (H={}), one trajectory at a time, float64 NumPy on {} BLAS threads: {}".format(
        per_step, w.horizon, cores, how)
    cfg = _config_dict(args, w, per_step, 1, "host")
    cfg["workload"] = ("{}: GP N={} H={} n_s={} n_u={} kernel={}; CPU arm: {} rollouts per step on the host cores of "
                       "rank 0 (launched with --gpus {}; no GPU is used)".format(
                           w.name, w.n_train, w.horizon, w.n_s, w.n_u, w.kern_types[0], per_step, world))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _config_dict(args, w, b_per_gpu, world, where):
    return {"workload": "{}: {} GP N={} H={} n_s={} n_u={} kernel={}; B={} candidate sequences per GPU x {} GPU(s) "
                        "= {} (BASELINE {} is B={})".format(
                            w.name, "cart-pole" if w.n_s == 4 else ("pendulum" if w.n_s == 2 else "synthetic 10-D"),
                            w.n_train, w.horizon, w.n_s, w.n_u, w.kern_types[0], b_per_gpu, world, b_per_gpu * world,
                            w.name, _cfg_batch(w.name)),
            "n_train": w.n_train, "horizon": w.horizon, "n_s": w.n_s, "n_u": w.n_u,
            "batch_per_gpu": b_per_gpu, "global_batch": b_per_gpu * world, "parallelism": "dp{}".format(world),
            "inputs": where,
            "l2_policy": _l2_policy(w, b_per_gpu)}


def _cfg_batch(name):
    from safe_exploration_b200 import workloads
    return workloads.CONFIGS[name][5]


def _l2_policy(w, b_per_gpu):
    """Working set of one H-step call against the 126 MB L2: the int8 digit planes of the factor (5 B per entry of
    the lower block triangle) and of the K* block of one chunk (5 B per kernel value), both re-read every step."""
    n_pad = -(-w.n_train // 128) * 128
    factor_mb = w.n_s * (n_pad // 128) * (n_pad // 128 + 1) * 5 * 8192 / 1e6
    kstar_mb = w.n_s * n_pad * 5 * min(b_per_gpu, 8192) / 1e6
    if factor_mb + kstar_mb > 2 * 126:
        return ("inputs larger than L2: every step streams the packed factor ({:.0f} MB) and a K* block of {:.0f} MB "
                "against the 126 MB L2".format(factor_mb, kstar_mb))
    return ("no flush: factor ({:.0f} MB) + K* block ({:.0f} MB) fit the 126 MB L2 at this configuration; the steps of a "
            "call are dependent (step t reads what step t-1 wrote), so a flush between calls would not change what "
            "the H-step loop sees".format(factor_mb, kstar_mb))


def run_product(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import safe_exploration_b200 as se
    from safe_exploration_b200 import distributed as sd
    from safe_exploration_b200 import workloads

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the product has no CPU path)")
    se.ssm.DEFAULT_TRI_MODE = args.tri_mode
    se.ssm.DEFAULT_I8_DIGITS = args.i8_digits
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = workloads.CONFIGS[args.config]
    if args.batch_per_gpu:
        b_per_gpu = args.batch_per_gpu
    elif args.scaling == "strong":
        b_per_gpu = max(1, cfg[5] // world)          # the quoted configuration's B, whatever the GPU count
    else:
        b_per_gpu = max(1, cfg[5] // cfg[6])         # the per-GPU shard of the quoted configuration
    w = workloads.make(args.config, batch=b_per_gpu * world, n_train=args.n_train or None, kern=args.kern or None)
    s0, s1 = sd.shard_range(b_per_gpu * world, rank, world)
    k_ff_shard = np.ascontiguousarray(w.k_ff[s0:s1])

    t_setup0 = time.perf_counter()
    gp = sd.build_replicated_model(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, w.kern_types, w.hyp, rank, world,
                                   src=0, device=local_rank, redundant=args.redundant_factor)
    torch.cuda.synchronize(dev)
    t_setup = time.perf_counter() - t_setup0
    if args.ksplit > 0:
        gp.set_option("ksplit", args.ksplit)
    if args.i8_panel_group > 0:
        gp.set_option("i8_panel_group", args.i8_panel_group)
    if args.i8_cluster > 0:
        gp.set_option("i8_cluster", args.i8_cluster)

    # ---------------- device-resident inputs (the `value` arm)
    p0_d = torch.as_tensor(w.p0, device=dev)
    kff_d = torch.as_tensor(k_ff_shard, device=dev)
    kfb_d = torch.as_tensor(w.k_fb, device=dev)
    rargs = (w.l_mu, w.l_sigma, None, None, w.c_safety, w.a, w.b)

    bsz_d, hor_d = k_ff_shard.shape[0], k_ff_shard.shape[1]
    out_d = se.RolloutResult(torch.empty((bsz_d, hor_d, w.n_s), dtype=torch.float64, device=dev),
                             torch.empty((bsz_d, hor_d, w.n_s, w.n_s), dtype=torch.float64, device=dev),
                             torch.empty((bsz_d, hor_d, w.n_s), dtype=torch.float64, device=dev),
                             torch.empty((bsz_d,), dtype=torch.int32, device=dev))
    if args.no_graph:
        gp.set_option("graph", 0)
    if args.guard_kappa > 0:
        gp.set_param("guard_kappa", args.guard_kappa)
    if args.substreams >= 0:
        gp.set_option("substreams", args.substreams)

    def step_device():
        # the result buffers are re-used, as a sampling-MPC loop would: from the second call on the library replays the
        # launches of a chunk as a CUDA graph
        return se.rollout(gp, p0_d, kff_d, kfb_d, *rargs, out=out_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        res = step_device()
    barrier()
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if os.environ.get("SEGP_I8_PROF"):
        gp.set_option("i8_prof", 1)
    launches0 = gp.get_option("launches")
    gp.set_option("time_tri", 1)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        res = step_device()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    tri_ns = gp.get_option("tri_ns")
    tri_count = gp.get_option("tri_launches")
    gp.set_option("time_tri", 0)
    launches = gp.get_option("launches") - launches0
    ms_max = sd.max_across_ranks(ms)
    status_bad = int((res.status != 0).sum().item())
    finite = bool(torch.isfinite(res.q_all).all().item())

    # ---------------- end-to-end arm: host buffers through segp_multistep_host
    kff_pin = torch.as_tensor(k_ff_shard).pin_memory()
    kff_host = kff_pin.numpy()
    bsz, hor = k_ff_shard.shape[0], k_ff_shard.shape[1]
    h2d = kff_host.nbytes + w.p0.nbytes + w.k_fb.nbytes
    d2h = 8 * bsz * hor * (w.n_s + w.n_s * w.n_s + w.n_s) + 4 * bsz

    out_pin = se.pinned_result(gp, bsz, hor)   # page-locked result buffers, allocated once, as the input is

    def step_host():
        return se.rollout(gp, w.p0, kff_host, w.k_fb, *rargs, out=out_pin)

    e2e_steps = max(1, args.e2e_steps or args.steps)      # the same number of steps as the device arm by default
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res_h = step_host()
    torch.cuda.synchronize(dev)
    t_e2e = time.perf_counter() - t0
    t_e2e_max = sd.max_across_ranks(t_e2e)
    if rank == 0:
        time.sleep(0.2)
        sampler.stop()
    same = bool(np.array_equal(res_h.q_all, res.q_all.cpu().numpy()))

    if os.environ.get("SEGP_I8_PROF") and rank == 0:
        from safe_exploration_b200.ssm import _tensor_from_ptr
        torch.cuda.synchronize(dev)
        view = _tensor_from_ptr(torch, gp.get_option("i8_prof_ptr"), 128 * 8 * 8, gp.device)
        prof = view.cpu().numpy().view(np.int64).reshape(128, 8)
        prof = prof[prof[:, 4] > 0]
        sys.stderr.write("i8_prof (last launch; MMA thread of each cluster): clusters %d\n" % len(prof))
        for name, col in (("total", 0), ("wait_full", 1), ("wait_peer", 2), ("wait_tmem_empty", 3), ("tiles", 4),
                          ("stages", 5)):
            c = prof[:, col]
            sys.stderr.write("  %-16s min %10d  median %10d  max %10d\n" % (name, c.min(), np.median(c), c.max()))
    if rank != 0:
        return
    peaks = _peaks()
    total_b = b_per_gpu * world
    value = total_b * args.steps / (ms_max * 1e-3)
    # roofline of the dominant kernel (tri_sumsq): algorithmic flop per launch = n_s * N^2 * columns of the launch
    # (with the pipelined driver a chunk is contracted as two half-chunk launches: the per-launch figures are the
    # timed region's totals divided by the number of launches actually timed)
    flop_launch = float(w.n_s) * w.n_train ** 2 * b_per_gpu * w.horizon * args.steps / max(tri_count, 1)
    tri_avg_s = (tri_ns / max(tri_count, 1)) * 1e-9
    achieved = flop_launch / tri_avg_s / 1e12 if tri_avg_s > 0 else None
    bf16_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    share = (tri_ns * 1e-6) / ms if ms > 0 else None
    mode = gp.get_option("tri_mode_effective")
    digits = gp.get_option("i8_digits_effective")
    if mode == 4 and gp.get_option("tri_persistent"):
        mode = 5   # automatic mode launched the persistent variant of the same contraction (short-tile models)
    rep = gp.precision_report()
    if mode in (1, 4, 5):
        # int8 digit planes: `products` int8 products per algorithmic multiply-add, exact int32 accumulation in TMEM.
        # `achieved` is ALGORITHMIC flop/s (n_s N^2 B per launch); `peak` is the measured bf16 figure of
        # MEASURED_PEAKS.json, so `frac` is the algorithmic fraction of the bf16 tensor peak: error-free splitting
        # caps it at 2/products (int8 runs at twice the bf16 rate).  `pipe_*` say how busy the int8 pipe actually
        # was: executed int8 op/s (padding and all products counted) against the int8 rate measured in this run by
        # segp_i8_peak (same instruction shape, no loads).
        products = 10 if digits == 4 else 15
        chunk = gp.get_option("chunk")
        n_pad = gp.get_option("n_train_padded")
        nblk = n_pad // 128
        # 96-trajectory panels contracted per launch, averaged the same way (padding of the last panel counted)
        panels = sum(-(-min(chunk, b_per_gpu - c0) // 96) for c0 in range(0, b_per_gpu, chunk)) \
            * w.horizon * args.steps / max(tri_count, 1)
        kblocks = nblk * (nblk + 1) // 2
        executed = 2.0 * products * w.n_s * (128 * 128 * kblocks) * panels * 96
        i8_96, i8_256 = _i8_peak(gp, local_rank, 96), _i8_peak(gp, local_rank, 256)
        dmma = _dmma_peak(gp, local_rank)   # the only native pipe that produces float64 products (mma.sync m8n8k4.f64)
        pipe_tops = executed / tri_avg_s / 1e12 if tri_avg_s > 0 else None
        roofline = {"bound": "tensor", "kernel": {1: "tri_i8_kernel", 4: "tri_i8m_kernel", 5: "tri_i8mp_kernel"}[mode]
                    + ("<split>" if digits == 4 else "<classic>"),
                    "pipe": "int8 tcgen05 (tcgen05.mma kind::i8, M=128 N=96/192 K=32, int32 accumulators in TMEM); "
                            "float64-grade result from balanced base-256 digit planes, {} products ({})".format(
                                products, "diagonal-split set, 4 x 4 digits + the diagonal's leading digit; guarded by "
                                "the a-posteriori error estimate, flagged panels recomputed with 15 products"
                                if digits == 4 else "5 x 5 digits"),
                    "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": (achieved / bf16_peak) if achieved else None,
                    "peak_source": peaks["_source"] + " bf16_tflops_sustained; algorithmic flop against the bf16 "
                                   "peak, ceiling 2/{} = {:.3f} for the {}-product int8 splitting".format(
                                       products, 2.0 / products, products),
                    "frac_of_splitting_ceiling": (achieved / (bf16_peak * 2.0 / products)) if achieved else None,
                    "int8_products": products,
                    "pipe_executed_tops": pipe_tops, "pipe_peak_tops_n96": i8_96, "pipe_peak_tops_n256": i8_256,
                    "pipe_frac_of_n96_peak": (pipe_tops / i8_96) if (pipe_tops and i8_96) else None,
                    "pipe_frac_of_n256_peak": (pipe_tops / i8_256) if (pipe_tops and i8_256) else None,
                    "pipe_frac_of_2x_bf16_peak": (pipe_tops / (2.0 * bf16_peak)) if pipe_tops else None,
                    "fp64_dmma_peak_tflops": dmma,
                    "achieved_over_fp64_dmma_peak": (achieved / dmma) if (achieved and dmma) else None,
                    "avg_launch_ms": tri_avg_s * 1e3, "launches_timed": tri_count, "share_of_step": share,
                    "algorithmic_flop_per_launch": flop_launch,
                    "traffic": _traffic(args.config, mode, digits),
                    "traffic_source": "profiles/traffic.json: dram bytes of the committed ncu --set full capture of this "
                                      "kernel and configuration (static; refreshed with the kernel), not measured in "
                                      "this run"}
    else:
        dmma = _dmma_peak(gp, local_rank)
        roofline = {"bound": "tensor", "kernel": "tri_sumsq_kernel", "pipe": "fp64 DMMA (mma.sync m8n8k4.f64)",
                    "achieved": achieved, "peak": dmma, "unit": "TFLOP/s",
                    "frac": (achieved / dmma) if (achieved and dmma) else None,
                    "peak_source": "segp_dmma_peak measured in this run (fp64 DMMA pipe; tri_mode=0); "
                                   + peaks["_source"] + " holds only bf16",
                    "bf16_peak": bf16_peak, "frac_of_bf16_peak": (achieved / bf16_peak) if achieved else None,
                    "avg_launch_ms": tri_avg_s * 1e3, "launches_timed": tri_count, "share_of_step": share,
                    "algorithmic_flop_per_launch": flop_launch, "traffic": _traffic(args.config, mode, 0)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "tri_mode": {0: "fp64 DMMA", 1: "int8 digit planes on tcgen05 (single CTA), float64 recombination",
                         2: "int8 digit planes on tcgen05 (CTA pairs), float64 recombination",
                         3: "int8 digit planes on tcgen05 (persistent CTA pairs), float64 recombination",
                         4: "int8 digit planes on tcgen05 (single-CTA MMAs, W multicast over CTA pairs), float64 "
                            "recombination",
                         5: "int8 digit planes on tcgen05 (persistent clusters over folded tiles, single-CTA MMAs, W "
                            "multicast over CTA pairs), float64 recombination"}[mode],
            "config": _config_dict(args, w, b_per_gpu, world, "device-resident"),
            "onestep_calls_per_sec": value * w.horizon,
            "algorithmic_tflops": value * w.horizon * workloads.flop_per_step(w.n_s, w.n_u, w.n_train) / 1e12,
            "e2e": {"value": total_b * e2e_steps / t_e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "api": "safe_exploration_b200.rollout(out=pinned_result(...)) -> segp_multistep_host (pinned host k_ff and result buffers)",
                    "bit_identical_to_device_arm": same},
            "gpu_launches": int(launches), "roofline": roofline,
            "precision": {"digit_set_products": (10 if digits == 4 else 15) if mode != 0 else None,
                          "guard_rtol": rep["guard_rtol"], "guard_kappa": rep["guard_kappa"],
                          "probe": {k: rep[k] for k in rep if k.startswith("probe_")},
                          "panels_recomputed_with_15_products": int(rep["fallback_panels"]),
                          "low_precision_flags": int((res.status & 8 != 0).sum().item())},
            "cuda_graph": {"enabled": bool(gp.get_option("graph")), "graphs_cached": int(gp.get_option("graphs_cached"))},
            "schedule": "serial per chunk ({} sub-batch chain(s)), replayed as a CUDA graph".format(
                2 if (gp.get_option("substreams") != 0 and gp.get_option("n_train_padded") <= 1024) else 1),
            "clocks": sampler.summary(t_wall0, t_wall1),
            "setup_s": t_setup, "factorize_gemms_on_tcgen05": bool(gp.get_option("fact_i8_effective")),
            "bad_status": status_bad, "all_finite": finite}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], oracle_out = cpu_baseline(w, args.cpu_seconds)
        line["parity"] = parity_against_oracle(res, oracle_out)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)


def _traffic(config, mode, digits):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum), or None for configurations that were not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        d = json.load(f)
    return d.get("{}:tri_mode{}:digits{}".format(config, mode, digits), d.get("{}:tri_mode{}".format(config, mode)) if digits != 4 else None)


def _i8_peak(gp, device, umma_n):
    import ctypes
    out = ctypes.c_double()
    rc = gp._lib.segp_i8_peak(device, umma_n, 20000, ctypes.byref(out))
    return out.value if rc == 0 else None


def _dmma_peak(gp, device):
    import ctypes
    out = ctypes.c_double()
    rc = gp._lib.segp_dmma_peak(device, 2000, ctypes.byref(out))
    return out.value if rc == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end arm (0 = --steps)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the quoted configuration's B on every GPU count (B_g = B / N); weak: its "
                         "per-GPU shard on every GPU (B = B_g x N)")
    ap.add_argument("--i8-digits", type=int, default=0, choices=[0, 4, 5],
                    help="digit set of the int8 contraction: 0 automatic (factorize-time probe), 4 = 10 products, "
                         "5 = 15 products")
    ap.add_argument("--kern", default="", choices=["", "rbf", "mat52", "lin_rbf", "lin_mat52"],
                    help="swap the configuration's kernel (lin_*: the composite kernels of the reference's journal "
                         "configs; the line is then not a BASELINE configuration)")
    ap.add_argument("--no-graph", action="store_true", help="direct launches instead of CUDA-graph replay")
    ap.add_argument("--substreams", type=int, default=-1, help="sub-batch streams of small models (-1 = library default)")
    ap.add_argument("--guard-kappa", type=float, default=0.0, help="override the precision guard's kappa (0 = library default)")
    ap.add_argument("--ref-rollouts", type=int, default=2, help="rollouts per step of the reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tri-mode", type=int, default=-1, choices=[-1, 0, 1, 4, 5],
                    help="variance contraction pipe: -1 auto (int8 tcgen05 when possible), 0 fp64 DMMA, 1 int8 tcgen05 "
                         "reference kernel (one CTA per tile), 4 single-CTA MMAs over merged K* planes, W multicast "
                         "over a CTA pair, 5 the same as a persistent kernel over folded (equal-length) tiles")
    ap.add_argument("--i8-panel-group", type=int, default=0,
                    help="panels per L2 group of the tcgen05 contraction (even; 0 = automatic); tuning experiments")
    ap.add_argument("--n-train", type=int, default=0,
                    help="override the configuration's number of training points (tuning experiments only: the line "
                         "is then not a BASELINE configuration)")
    ap.add_argument("--i8-cluster", type=int, default=0, choices=[0, 2, 4],
                    help="CTAs per cluster of tri_i8m sharing one W stage by multicast (0 = library default)")
    ap.add_argument("--ksplit", type=int, default=0,
                    help="splits of the training points in the K* kernel (0 = automatic); tuning experiments")
    ap.add_argument("--redundant-factor", action="store_true",
                    help="factorise on every rank instead of broadcasting the factor")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    # stdout carries exactly ONE line (rank 0's JSON): anything a library prints on file descriptor 1 meanwhile
    # (NCCL's version banner, for one) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    from safe_exploration_b200 import distributed as sd
    rank, world, local_rank = sd.init_from_env()
    try:
        run_product(args, rank, world, local_rank)
    finally:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
