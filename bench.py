#!/usr/bin/env python
"""bench.py -- benchmarks of the grail-rs waveform path on B200 (contract: see the task statement).

A "step" is one pass of the hot path (Sequencer -> Jitter -> Synthesize) over one batch of synthetic phoneme
sequences.  `--config` picks the BASELINE.json workload (the headline, and the default, is config 2):

  2  batch of 1 024 utterances x 10 phonemes (~5 s), default voice, 44.1 kHz per GPU ("scaling": "weak": utterances
     are independent, every rank owns one such batch, no data-path collective)
  3  one 10-minute utterance (26 457 161 samples): the parallel-in-time path; does not shard (N > 1: replicas)
  4  65 536 short utterances with per-utterance random voices, ONE fixed batch split over the ranks by LPT on the exact
     sample counts ("scaling": "strong"); optional NCCL gather of the outputs into batch order (--gather 1), per-rank
     load imbalance and a parity block against the oracle
  5  the config-2 shape at 16 / 22.05 / 44.1 / 48 kHz, one line with a table: GPU samples/s and host-CPU samples/s per rate

  value  = samples/s, whole job, inputs (phoneme tables, schedules, work items) already resident in HBM
  e2e    = the same metric through the C-ABI call a reference user would make (grail_cuda_synthesize_batch):
           host phoneme records in, host (pinned) f32 samples out, H2D + D2H inside the timed region
  roofline / cpu_baseline: see DESIGN.md section 6

  --impl reference  times the reference's own CPU algorithm (the oracle restatement, all host threads) on the same
                    workload, whole batch per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "audio samples/sec (batched utterances)"
UNIT = "samples/s"
SAMPLE_RATE = 44100.0
N_UTTS = 1024
N_PHONEMES = 10
N_UTTS_CONFIG4 = 65536
RATES_CONFIG5 = (16000.0, 22050.0, 44100.0, 48000.0)
# algorithmic f32 flops per sample, as written in the reference: whole path, and the share the
# dominant kernel (k_formant) covers = everything but the scalar frequency/phase lane (26 flops)
FLOPS_PER_SAMPLE = 762
FLOPS_PER_SAMPLE_FORMANT = 736
# the same count when the formants whose amplitude is zero in every phoneme are left out, as the reference's result
# allows (about 400 for the default voice's 4 active formants, minus the 26 of the frequency lane)
FLOPS_PER_SAMPLE_FORMANT_ACTIVE = 374
# What the kernels EXECUTE per config-2 launch (thread-level FP32 operations from the SASS opcode histogram of an
# `ncu --set full --import-source on` capture, FFMA = 2 flops, FFMA2 = 4, FADD2 / FMUL2 = 2) and the DRAM bytes ncu saw:
# written by scripts/roofline_from_ncu.py from the round's own capture, never typed in by hand.
ROOFLINE_FILE = os.path.join(ROOT, "profiles", "r2_roofline_inputs.json")


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_config2(rank: int, n_utts: int = N_UTTS, sample_rate: float = SAMPLE_RATE):
    from grail_rs_b200 import workloads as W
    elems, offs, vp = W.config2(n_utts, N_PHONEMES, sample_rate)
    vp = vp.copy()
    vp["jitter_seed"] = (np.arange(n_utts) + rank * n_utts).astype(np.uint32)
    return elems, offs, vp


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            inside = t0 - 0.05 <= ts <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(f[1]))
                    power.append(float(f[3]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            if inside:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def roofline_inputs() -> dict:
    try:
        return json.load(open(ROOFLINE_FILE))
    except Exception:
        return {}


def bind_to_gpu_cpus(device_index: int):
    """pin the calling process to the CPUs NVML reports as local to the GPU; returns the mask size or None"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


_COUNT_CACHE: dict = {}


def cpu_run(elems, offs, vp, threads: int):
    """the reference's CPU algorithm (oracle restatement, -O3 strict f32), one utterance per task on `threads` host
    threads; returns (samples, seconds)"""
    from oracle import oracle as O
    # the oracle's own sample counts (nothing of the product on this path), memoised per (phoneme lengths, rate)
    counts = np.empty(len(offs) - 1, np.uint64)
    for u in range(len(offs) - 1):
        e = elems[offs[u]:offs[u + 1]]
        key = (e["length"].tobytes(), float(vp[u]["sample_rate"]))
        if key not in _COUNT_CACHE:
            _COUNT_CACHE[key] = O.count_samples(e, float(vp[u]["sample_rate"]))
        counts[u] = _COUNT_CACHE[key]
    oo = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    out = np.empty(int(oo[-1]), np.float32)
    t0 = time.perf_counter()
    O.synthesize_batch(elems, offs, vp, out_offsets=oo, n_threads=threads, o3=True, out=out)
    return int(oo[-1]), time.perf_counter() - t0


def config_workload(config: int, rank: int, world: int):
    """(elems, offs, vp, description, scaling, extra) of this rank's share of the workload"""
    from grail_rs_b200 import workloads as W
    if config == 2:
        e, o, v = workload_config2(rank)
        return e, o, v, (f"config2: {N_UTTS} utterances x {N_PHONEMES} phonemes (~5 s) per GPU, default voice, 44.1 kHz, "
                         "jitter_seed = utterance index"), "weak", {}
    if config == 3:
        e, o, v = W.config3()
        return e, o, v, "config3: one 10-minute utterance (1 200 phonemes, 26 457 161 samples), default voice, 44.1 kHz; replicas at N > 1", "weak", {}
    if config == 4:
        import grail_rs_b200 as g
        from grail_rs_b200 import sharding
        e, o, v = W.config4(N_UTTS_CONFIG4)
        counts = g.count_samples(e, o, v)
        assign = sharding.lpt_assign(counts, world)
        se, so, sv = sharding.shard_batch(e, o, v, assign[rank])
        loads = [int(counts[a].sum()) for a in assign]
        extra = {"assign": assign, "counts": counts, "full": (e, o, v), "loads": loads}
        return se, so, sv, (f"config4: {N_UTTS_CONFIG4} utterances x 2-4 phonemes, per-utterance random voice (8 live formants, "
                            f"pitch, jitter), 44.1 kHz, ONE batch split over {world} GPU(s) by LPT on exact sample counts"), "strong", extra
    raise ValueError(config)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    if args.config == 5:
        return run_reference_config5(args, cores)
    elems, offs, vp, desc, scaling, extra = config_workload(args.config, 0, 1)
    if args.config == 4:
        # bounded sample of the 65 536 utterances: every 16th (4 096 utterances, 2.7e8 samples) per step
        from grail_rs_b200 import sharding
        e, o, v = extra["full"]
        elems, offs, vp = sharding.shard_batch(e, o, v, np.arange(0, N_UTTS_CONFIG4, 16))
        sample = "every 16th of the 65 536 config-4 utterances per step"
    elif args.config == 3:
        cores = 1
        sample = "the whole utterance, one core (a single chain does not parallelise on the CPU)"
    else:
        sample = f"all {N_UTTS} config-2 utterances per step"
    if args.warmup and len(offs) - 1 > cores:          # page the library and the thread pool in on a few utterances
        from grail_rs_b200 import sharding
        cpu_run(*sharding.shard_batch(elems, offs, vp, np.arange(cores)), cores)
    total, secs = 0, 0.0
    for _ in range(args.steps):
        n, dt = cpu_run(elems, offs, vp, cores)
        total += n
        secs += dt
    v = total / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} ({total // args.steps} samples), oracle -O3 strict f32, one utterance per task"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rtf": v / SAMPLE_RATE,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_reference_config5(args, cores):
    rows, total, secs = [], 0, 0.0
    for rate in RATES_CONFIG5:
        e, o, v = workload_config2(0, N_UTTS, rate)
        n, dt = cpu_run(e, o, v, cores)
        rows.append({"sample_rate": rate, "cpu_samples_per_s": n / dt, "samples": n})
        total += n
        secs += dt
    v = total / secs
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
                      "warmup": 0, "ms_per_step": 1e3 * secs, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "config5: the config-2 shape at 16 / 22.05 / 44.1 / 48 kHz", "rates": rows},
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                       "sample": "all 1 024 utterances at each of the four rates, once"},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
    return 0


class Dist:
    """rank / world plumbing: NCCL only for barriers, the max-over-ranks timing and the optional output gather"""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
        torch.cuda.set_device(self.local)
        # NUMA: keep this rank's threads (and therefore its pinned host buffers, which are placed where they are first
        # touched) on the CPUs next to its GPU; the CPU baseline leg widens the mask again.
        self.all_cpus = os.sched_getaffinity(0)
        self.numa = bind_to_gpu_cpus(self.local) if self.world > 1 else None
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_resident(D: Dist, ctx, plan, d_out, steps: int, warmup: int, sampler=None):
    """W warm-up launches, then exactly `steps` launches of the resident plan between two events on the library's
    stream, barrier + synchronize on both sides; returns (ms max over ranks, wall t0, wall t1)"""
    torch = D.torch
    stream = torch.cuda.ExternalStream(ctx.stream_handle, device=torch.device("cuda", D.local))
    for _ in range(max(warmup, 3)):
        plan.launch(d_out)
    plan.join()
    ctx.synchronize()
    if sampler is not None:
        sampler.start()
        time.sleep(0.25)
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(steps):
        plan.launch(d_out)
    plan.join()                      # the main stream waits for pipelined launches before the end event
    e1.record(stream)
    e1.synchronize()
    D.barrier()
    t1 = time.time()
    return D.max(e0.elapsed_time(e1)), t0, t1


def kernel_split(ctx, plan, d_out, reps: int):
    """per-kernel CUDA-event times of single launches (events inside the library, same stream), averaged"""
    kern = {"frequency_ms": 0.0, "phase_ms": 0.0, "formant_ms": 0.0, "schedule_ms": 0.0, "total_ms": 0.0}
    for _ in range(reps):
        plan.launch(d_out)
        ctx.synchronize()
        t = plan.timings()
        for k in kern:
            kern[k] += t[k] / reps
    return kern


def roofline_block(ctx, n_samples: int, kern: dict, step_ms: float, config: int) -> dict:
    """the dominant kernel (k_formant) against the FP32 pipe: executed flops from the round's ncu capture (config 2 only,
    where the capture was taken), the as-written / live-formant algorithmic counts as separate named fields"""
    probe = ctx.probe_fp32_peak()
    peaks = measured_peaks()
    ri = roofline_inputs() if config == 2 else {}
    peak_tf = probe["ffma_flops"] / 1e12
    f_ms = kern["formant_ms"]
    as_written = FLOPS_PER_SAMPLE_FORMANT * n_samples / (f_ms * 1e-3) / 1e12
    live = FLOPS_PER_SAMPLE_FORMANT_ACTIVE * n_samples / (f_ms * 1e-3) / 1e12
    kf = ri.get("k_formant", {})
    executed = kf.get("executed_flops_per_launch")
    achieved = executed / (f_ms * 1e-3) / 1e12 if executed else None
    hbm_bytes = 4 * n_samples * 2          # saw read + f32 samples written (algorithmic bytes of k_formant)
    step_exec = ri.get("step_executed_flops")
    out = {
        "kernel": "k_formant", "bound": "fp32",
        "achieved": achieved if achieved is not None else live, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": (achieved if achieved is not None else live) / peak_tf,
        "achieved_is": ("FP32 operations the kernel EXECUTES per launch (SASS opcode histogram of the committed ncu capture, "
                        f"{ROOFLINE_FILE.replace(ROOT + '/', '')}) / launch time measured here") if achieved is not None
                       else "algorithmic flops of the live formants (no ncu capture for this workload)",
        "traffic": kf.get("dram_bytes_per_launch"),
        "peak_source": "FFMA issue-rate probe measured live in this run (MEASURED_PEAKS.json has no fp32 figure)",
        "launch_ms": f_ms,
        "as_written": {"flops_per_sample": FLOPS_PER_SAMPLE_FORMANT, "tflops": as_written, "frac": as_written / peak_tf,
                       "note": "SURVEY 8d count of the reference's source; exceeds 1 because zero-amplitude formants are skipped and "
                               "coefficients are interpolated over 16-sample blocks: NOT a utilisation figure"},
        "live_formants": {"flops_per_sample": FLOPS_PER_SAMPLE_FORMANT_ACTIVE, "tflops": live, "frac": live / peak_tf},
        "whole_step": {"ms": step_ms, "as_written_frac": FLOPS_PER_SAMPLE * n_samples / (step_ms * 1e-3) / 1e12 / peak_tf,
                       "executed_frac": (step_exec / (step_ms * 1e-3) / 1e12 / peak_tf) if step_exec else None},
        "issue_slots_busy_ncu": kf.get("issue_slots_busy_pct"), "fma_pipe_busy_ncu": kf.get("fma_pipe_busy_pct"),
        "mufu_peak_per_s": probe["mufu_ops"],
        "hbm": {"achieved": hbm_bytes / (f_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": (hbm_bytes / (f_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                "peak_source": "MEASURED_PEAKS.json (measured)" if peaks.get("hbm_gbs") else "absent"},
    }
    return out


def parity_block(ctx, plan, out_t, elems, offs, vp, n_check: int) -> dict:
    """n_check utterances of this rank's shard against the oracle (audio within tolerance); counts exact for ALL"""
    from oracle import oracle as O
    from grail_rs_b200 import sharding
    from grail_rs_b200 import workloads as W
    import grail_rs_b200 as g
    n = len(offs) - 1
    oo = plan.out_offsets
    counts = g.count_samples(elems, offs, vp)
    counts_ok = bool(np.array_equal(np.diff(oo.astype(np.int64)), counts.astype(np.int64)))
    pick = np.unique(np.linspace(0, n - 1, min(n, n_check)).astype(np.int64))
    se, so, sv = sharding.shard_batch(elems, offs, vp, pick)
    want, woo, wc = O.synthesize_batch(se, so, sv, n_threads=host_cores())
    counts_ok &= bool(np.array_equal(wc.astype(np.int64), counts[pick].astype(np.int64)))
    worst = {"max_abs": 0.0, "snr_db": float("inf")}
    for k, u in enumerate(pick.tolist()):
        got = out_t[int(oo[u]):int(oo[u + 1])].cpu().numpy()
        st = W.parity_stats(got, want[int(woo[k]):int(woo[k + 1])])
        worst["max_abs"] = max(worst["max_abs"], st["max_abs"])
        worst["snr_db"] = min(worst["snr_db"], st["snr_db"])
    return {"utterances_checked": int(len(pick)), "counts_bit_exact_all": counts_ok, "max_abs": worst["max_abs"],
            "snr_db": worst["snr_db"], "within_tolerance": bool(worst["max_abs"] <= 1e-4 and worst["snr_db"] >= 90.0),
            "phase_unproven_utterances": plan.phase_stats()["unproven_utterances"]}


def run_ours(args):
    if args.config == 5:
        return run_ours_config5(args)
    D = Dist()
    torch = D.torch
    import grail_rs_b200 as g
    rank, world = D.rank, D.world
    ctx = g.Context(D.local)
    elems, offs, vp, desc, scaling, extra = config_workload(args.config, rank, world)
    # consecutive launches of the resident plan overlap: step i+1's frequency / phase kernels run on a second stream
    # and scratch set under step i's formant kernel (library option "pipeline"; same bits)
    ctx.set_option("pipeline", 1 if args.pipeline else 0)
    plan = ctx.plan(elems, offs, vp)
    ctx.set_option("pipeline", 0)
    n_samples = plan.total_samples
    out = torch.empty(n_samples, dtype=torch.float32, device="cuda")
    d_out = out.data_ptr()

    # ---------------- device-resident throughput ("value") ----------------
    sampler = ClockSampler(D.local) if rank == 0 else None
    ms, t_wall0, t_wall1 = timed_resident(D, ctx, plan, d_out, args.steps, args.warmup, sampler)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    total_samples = D.sum(float(n_samples))          # all ranks (config 2: N x the batch; config 4: the one batch)
    value = total_samples * args.steps / (ms * 1e-3)
    # the same K steps launched in order (no overlap between consecutive steps), for the record: `value` above uses the
    # plan option "pipeline" unless --pipeline 0
    in_order = None
    if args.pipeline and args.config in (2, 3):
        plan_io = ctx.plan(elems, offs, vp)          # (the ctx option is back to 0 here)
        ms_io, _, _ = timed_resident(D, ctx, plan_io, d_out, args.steps, args.warmup)
        in_order = {"value": total_samples * args.steps / (ms_io * 1e-3), "unit": UNIT, "ms_per_step": ms_io / args.steps,
                    "note": "the same plan without the pipeline option: every step's kernels in stream order"}
        plan_io.close()
    reps = max(3, min(args.steps, 10))
    kern = kernel_split(ctx, plan, d_out, reps)
    launches = plan.timings()["n_launches"] * args.steps
    pstats = plan.phase_stats()

    # ---------------- config 4: imbalance, parity block, optional NCCL gather ----------------
    cfg4 = None
    if args.config == 4:
        from grail_rs_b200 import sharding
        loads = extra["loads"]
        par = parity_block(ctx, plan, out, elems, offs, vp, args.parity_utts)
        par_all = {"within_tolerance": bool(D.sum(0.0 if par["within_tolerance"] and par["counts_bit_exact_all"] else 1.0) == 0.0),
                   "max_abs": D.max(par["max_abs"]), "snr_db": -D.max(-par["snr_db"]),
                   "utterances_checked": int(D.sum(par["utterances_checked"])), "per_rank": par["utterances_checked"]}
        cfg4 = {"load_imbalance_max_over_mean": max(loads) / (sum(loads) / len(loads)), "samples_per_rank": loads,
                "parity": par_all}
        if args.gather and world > 1:
            # (the first all_gather of a process pays NCCL's channel set-up and buffer registration -- 0.7 s was seen at
            #  N = 4 --, so one small warm-up collective goes first and the full gather is timed on its second run)
            warm = torch.zeros(1 << 20, dtype=torch.float32, device="cuda")
            warm_all = torch.empty(world << 20, dtype=torch.float32, device="cuda")
            D.dist.all_gather_into_tensor(warm_all, warm)
            del warm, warm_all
            full = sharding.gather_outputs(out, extra["counts"][extra["assign"][rank]], extra["assign"], extra["counts"], ctx=ctx)
            del full
            D.barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            full = sharding.gather_outputs(out, extra["counts"][extra["assign"][rank]], extra["assign"], extra["counts"], ctx=ctx)
            g1.record()
            torch.cuda.synchronize()
            gather_ms = D.max(g0.elapsed_time(g1))
            # every rank's piece sits where batch order says, bit for bit
            all_off = np.concatenate([[0], np.cumsum(extra["counts"].astype(np.int64))])
            mine = extra["assign"][rank]
            oo = plan.out_offsets
            ok = True
            for k in np.unique(np.linspace(0, len(mine) - 1, 64).astype(np.int64)).tolist():
                u = int(mine[k])
                ok &= bool(torch.equal(full[all_off[u]:all_off[u + 1]], out[int(oo[k]):int(oo[k + 1])]))
            cfg4["gather"] = {"ms": gather_ms, "bytes_gathered_per_rank": int(full.numel() * 4),
                              "gb_per_s_per_rank": full.numel() * 4 / (gather_ms * 1e-3) / 1e9,
                              "own_pieces_bit_equal": bool(D.sum(0.0 if ok else 1.0) == 0.0),
                              "checksum": float(full[:: max(1, full.numel() // (1 << 20))].double().abs().sum().item()),
                              "how": "one NCCL all_gather_into_tensor of padded shards + one grail_cuda_copy_segments launch"}
            del full

    # ---------------- end to end through the C ABI with host buffers ("e2e") ----------------
    e2e = None
    e2e_i16 = None
    if n_samples * 4 <= args.e2e_max_gb * (1 << 30):
        host_out = ctx.pinned_empty(n_samples, np.float32)
        oo = plan.out_offsets.copy()
        h2d = elems.nbytes + offs.nbytes + vp.nbytes
        e2e_steps = max(2, min(args.steps, 5))
        ctx.synthesize_batch(elems, offs, vp, out=host_out, out_offsets=oo)   # warm the buffer pool
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.synthesize_batch(elems, offs, vp, out=host_out, out_offsets=oo)
        D.barrier()
        e2e_s = D.max(time.perf_counter() - t0)
        e2e = {"value": total_samples * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(n_samples * 4), "steps": e2e_steps,
               "api": "grail_cuda_synthesize_batch (host records in, pinned f32 out)", "cpus_bound_per_rank": D.numa}
        checksum = float(np.abs(host_out[: min(n_samples, 220476)]).sum())
        # the same call with i16 PCM out (what the reference's WAV path keeps, examples/cli.rs:49-51): half the D2H bytes
        host_pcm = ctx.pinned_empty(n_samples, np.int16)
        ctx.synthesize_batch(elems, offs, vp, out=host_pcm, out_offsets=oo, fmt=g.I16)
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.synthesize_batch(elems, offs, vp, out=host_pcm, out_offsets=oo, fmt=g.I16)
        D.barrier()
        e2e_i16 = {"value": total_samples * e2e_steps / D.max(time.perf_counter() - t0), "unit": UNIT,
                   "d2h_bytes_per_step": int(n_samples * 2),
                   "api": "grail_cuda_synthesize_batch_i16 (the reference's WAV sample format, pinned i16 out)"}
    else:
        checksum = float(out[: min(n_samples, 220476)].abs().sum().item())

    line = None
    if rank == 0:
        roofline = roofline_block(ctx, n_samples, kern, ms / args.steps, args.config)
        # ---------------- CPU baseline on this box's host cores (bounded sample) ----------------
        os.sched_setaffinity(0, D.all_cpus)
        cores = host_cores()
        if args.config == 3:
            se, so, sv, cores_used, sample = elems, offs, vp, 1, "the whole 10-minute utterance once, one core"
        else:
            from grail_rs_b200 import sharding
            n = len(offs) - 1
            nb = min(n, max(cores, 32 * cores if args.config == 2 else 256 * cores))
            se, so, sv = sharding.shard_batch(elems, offs, vp, np.arange(nb))
            cores_used, sample = cores, f"the first {nb} of this rank's {n} utterances"
        n_cpu, dt_cpu = 0, 0.0
        while dt_cpu < 2.0:                      # bounded: ~10-30 core-seconds of CPU work
            n_i, dt_i = cpu_run(se, so, sv, cores_used)
            n_cpu += n_i
            dt_cpu += dt_i
        cpu = {"value": n_cpu / dt_cpu, "unit": UNIT, "cores": cores_used, "kind": "port",
               "sample": f"{sample}, repeated to {n_cpu} samples ({dt_cpu:.2f} s wall), oracle -O3 strict f32, one utterance per task"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "samples_per_step_rank0": n_samples, "samples_per_step_all_ranks": int(total_samples),
                       "l2": f"inputs larger than L2 ({12 * n_samples / 1e9:.1f} GB touched per step)",
                       "parallelism": f"utterance-sharded x{world}, no data-path collective",
                       "pipeline": "consecutive steps overlap on two streams (plan option)" if args.pipeline else "off"},
            "e2e": e2e if e2e is not None else {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                                "skipped": f"output of {n_samples * 4 / 2**30:.1f} GiB per rank exceeds --e2e-max-gb"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "kernels_ms": kern, "phase_stats": pstats, "rtf_per_gpu": value / world / float(vp[0]["sample_rate"]),
            "checksum": checksum,
        }
        if e2e_i16 is not None:
            line["e2e_i16"] = e2e_i16
        if in_order is not None:
            line["in_order"] = in_order
        if cfg4 is not None:
            line["config4"] = cfg4
    plan.close()
    ctx.close()
    D.close()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def run_ours_config5(args):
    """the config-2 shape at four sample rates: per rate the device-resident samples/s over all ranks, the end-to-end
    figure, and (rank 0) the host CPU's samples/s on a bounded sample -- one JSON line with the table"""
    D = Dist()
    torch = D.torch
    import grail_rs_b200 as g
    from grail_rs_b200 import sharding
    ctx = g.Context(D.local)
    rows, tot_samples, tot_ms, launches = [], 0.0, 0.0, 0
    clocks = None
    for i, rate in enumerate(RATES_CONFIG5):
        elems, offs, vp = workload_config2(D.rank, N_UTTS, rate)
        ctx.set_option("pipeline", 1 if args.pipeline else 0)
        plan = ctx.plan(elems, offs, vp)
        ctx.set_option("pipeline", 0)
        n = plan.total_samples
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        sampler = ClockSampler(D.local) if (D.rank == 0 and i == 2) else None
        ms, t0, t1 = timed_resident(D, ctx, plan, out.data_ptr(), args.steps, args.warmup, sampler)
        if sampler is not None:
            clocks = sampler.stop(t0, t1)
        kern = kernel_split(ctx, plan, out.data_ptr(), 3)
        launches += plan.timings()["n_launches"] * args.steps
        total = D.sum(float(n))
        row = {"sample_rate": rate, "samples_per_utterance": int(n // N_UTTS), "gpu_samples_per_s": total * args.steps / (ms * 1e-3),
               "ms_per_step": ms / args.steps, "kernels_ms": kern, "rtf_per_gpu": total * args.steps / (ms * 1e-3) / D.world / rate}
        if D.rank == 0:
            os.sched_setaffinity(0, D.all_cpus)
            cores = host_cores()
            nb = min(N_UTTS, 32 * cores)
            se, so, sv = sharding.shard_batch(elems, offs, vp, np.arange(nb))
            n_cpu, dt_cpu = cpu_run(se, so, sv, cores)
            row["cpu_samples_per_s"] = n_cpu / dt_cpu
            row["cpu_cores"] = cores
            row["cpu_sample"] = f"first {nb} utterances once"
            row["gpu_over_cpu"] = row["gpu_samples_per_s"] / row["cpu_samples_per_s"]
        rows.append(row)
        tot_samples += total * args.steps
        tot_ms += ms
        plan.close()
        del out
    if D.rank == 0:
        line = {"metric": METRIC, "value": tot_samples / (tot_ms * 1e-3), "unit": UNIT, "n_gpus": D.world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "config5: the config-2 shape (1 024 utterances x 10 phonemes per GPU, default voice rebuilt per "
                                       "rate) at 16 / 22.05 / 44.1 / 48 kHz; a step = the four rates one after another",
                           "l2": "inputs larger than L2", "rates": rows},
                "e2e": {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                        "skipped": "see --config 2 for the end-to-end figure (same shape)"},
                "gpu_launches": int(launches), "clocks": clocks,
                "cpu_baseline": {"value": float(np.mean([r["cpu_samples_per_s"] for r in rows])), "unit": UNIT,
                                 "cores": rows[0]["cpu_cores"], "kind": "port", "sample": "per rate: " + rows[0]["cpu_sample"]}}
        print(json.dumps(line), flush=True)
    ctx.close()
    D.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json workload (2 = headline)")
    ap.add_argument("--gather", type=int, default=1, help="config 4, N > 1: gather the outputs into batch order over NCCL")
    ap.add_argument("--parity-utts", type=int, default=64, help="config 4: utterances per rank checked against the oracle")
    ap.add_argument("--e2e-max-gb", type=float, default=4.0, help="skip the end-to-end leg when a rank's f32 output is larger")
    ap.add_argument("--pipeline", type=int, default=1, help="overlap consecutive steps of the resident plan (0 = launch in order)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
