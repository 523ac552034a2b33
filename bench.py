#!/usr/bin/env python
"""bench.py -- pairings/sec of the batched pairing hot path on N B200s (BASELINE.json metric).

  python bench.py --gpus 1 --steps 5 --warmup 3                  # our arm (CUDA, libpbc_b200.so)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W  # N > 1: one rank per GPU
  python bench.py --impl reference --gpus 1 --steps K --warmup W # the reference CPU path

One "step" = one pass of the hot path over one batch: 2^20 Type A (param/a.param) pairings per
GPU (BASELINE.json configs[1]), synthetic seeded inputs in the reference wire format.  The batch
shards embarrassingly: every rank owns its own 2^20 pairs (weak scaling), no collective on the
data path; timing is CUDA events on the launching stream, max over ranks.

`value`  : whole-job pairings/s with inputs already resident in HBM.
`e2e`    : same metric through the host-buffer C-ABI call (pbc_b200_pairings_apply) with pinned
           HOST buffers: H2D of both inputs and D2H of the outputs inside the timed region.
`roofline`: integer-multiplier pipe: algorithmic IMAD.WIDE.U32 unit operations of the
           dominant kernel (Miller loop) / its CUDA-event duration, against the IMAD.WIDE.U32
           issue rate measured live by a microkernel; plus the (non-binding) HBM view.
`cpu_baseline`: the unmodified reference (oracle/_ref) on the host cores, bounded sample, whose
           outputs are also compared byte-for-byte with the GPU's.
"""
from __future__ import annotations
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pbc_b200.params import PARAMS  # noqa: E402
from pbc_b200 import synth  # noqa: E402

GRID = 4096                    # inputs are drawn from a GRID x GRID grid of distinct points
SEED = 20260922
# --- algorithmic work, SURVEY.md 8(d) / DESIGN.md "work per unit" ---------------------------
UNIT_OPS_PER_MULMOD = {"a": 528}          # 2(2t)^2 + 2t IMAD.WIDE.U32 for t = 8 64-bit limbs
REF_MULMODS = {"a": 4394}                 # reference algorithm, whole pairing
REF_MULMODS_MAIN = {"a": 4394 - 719}      # ... of which Miller loop (final exp = 719)
# what k_a_miller executes per pairing: 2093 multiplications (528 unit ops) + 961 squarings
# (408 unit ops: product-scanning squaring) -- DESIGN.md "work per unit"
EXEC_MULMODS_MAIN = {"a": 2093 + 961}
EXEC_UNIT_OPS_MAIN = {"a": 2093 * 528 + 961 * 408}
WIRE_BYTES = {"a": (128, 128, 128)}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------
# CPU reference leg (oracle/_ref = the unmodified reference; else the Python port)
# ------------------------------------------------------------------------------------------
def _ref_worker(args):
    name, Pb, Qb, n = args
    from oracle import ref as R
    rp = R.RefPairing(PARAMS[name])
    t0 = time.perf_counter()
    out = rp.pairing(Pb, Qb, n)
    return out, time.perf_counter() - t0


def _port_worker(args):
    name, Pb, Qb, n = args
    from oracle import pbc_oracle as O
    pr = O.pairing_from_param(PARAMS[name])
    t0 = time.perf_counter()
    out = O.pairing_batch(pr, Pb, Qb, n)
    return out, time.perf_counter() - t0


def cpu_kind():
    try:
        from oracle import ref as R
        if R.available():
            return "reference"
    except Exception:
        pass
    return "port"


def cpu_pairings(name, P, Q, n, cores, pool=None):
    """n pairings on `cores` processes (each with its own pairing_t: the library is not
    thread-safe).  Returns (bytes, wall seconds)."""
    import multiprocessing as mp
    g1, g2, _ = WIRE_BYTES[name]
    kind = cpu_kind()
    worker = _ref_worker if kind == "reference" else _port_worker
    per = (n + cores - 1) // cores
    jobs = []
    for c in range(cores):
        lo, hi = c * per, min(n, (c + 1) * per)
        if lo < hi:
            jobs.append((name, bytes(P[lo * g1:hi * g1]), bytes(Q[lo * g2:hi * g2]), hi - lo))
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(len(jobs))
    t0 = time.perf_counter()
    res = pool.map(worker, jobs)
    wall = time.perf_counter() - t0
    if own:
        pool.close()
        pool.join()
    return b"".join(r[0] for r in res), wall


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_rate_guess(name, kind):
    return {"reference": {"a": 1100.0}, "port": {"a": 110.0}}[kind][name]


# ------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.rows = []
        self.proc = None
        self.dev = dev

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    name = args.workload
    prm = synth.parse_param(PARAMS[name])
    Pb, Qb = synth.type_a_points(prm, GRID, SEED)
    cores = host_cores()
    kind = cpu_kind()
    per_step = int(min(args.n, max(cores * 64, args.ref_seconds * cpu_rate_guess(name, kind) * cores)))
    import multiprocessing as mp
    P, Q = synth.build_batch(Pb, Qb, per_step, 0)
    P, Q = P.tobytes(), Q.tobytes()
    pool = mp.get_context("fork").Pool(cores)
    for _ in range(args.warmup):
        cpu_pairings(name, P, Q, min(per_step, cores * 16), cores, pool)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pairings(name, P, Q, per_step, cores, pool)
    wall = time.perf_counter() - t0
    pool.close()
    pool.join()
    val = per_step * args.steps / wall
    line = {
        "impl": "reference", "metric": "pairings/sec", "value": val, "unit": "pairings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (512-bit F_q)" if kind == "port" else "u64 limbs (GMP mpn)",
        "data": "synthetic",
        "config": {"workload": workload_name(name), "batch_per_step": per_step, "param": name + ".param",
                   "note": "bounded sample of the 2^20 workload per step; CPU only"},
        "cpu_baseline": {"value": val, "unit": "pairings/s", "cores": cores, "kind": kind,
                         "sample": "%d pairings per step x %d steps, %d processes" % (per_step, args.steps, cores)},
        "e2e": {"value": val, "unit": "pairings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_name(name):
    return {"a": "type A (param/a.param) element_pairing, batch 2^20 (P,Q) pairs per GPU, 512-bit F_q"}[name]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="a", choices=["a"])
    ap.add_argument("--n", type=int, default=1 << 20, help="pairings per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget (rank 0, N=1)")
    ap.add_argument("--ref-seconds", type=float, default=3.0, help="--impl reference: seconds per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    name = args.workload
    n = args.n
    g1, g2, gt = WIRE_BYTES[name]

    # ---- synthetic inputs (host) ----
    prm = synth.parse_param(PARAMS[name])
    Pb, Qb = synth.type_a_points(prm, GRID, SEED)
    Ph, Qh = synth.build_batch(Pb, Qb, n, offset=rank * n)

    # ---- CPU baseline first (forks before CUDA is initialised) ----
    cpu = None
    cpu_out = None
    sample_n = 0
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        kind = cpu_kind()
        sample_n = int(min(n, max(cores * 32, args.cpu_seconds * cpu_rate_guess(name, kind) * cores)))
        cpu_out, wall = cpu_pairings(name, Ph[:sample_n * g1].tobytes(), Qh[:sample_n * g2].tobytes(),
                                     sample_n, cores)
        cpu = {"value": sample_n / wall, "unit": "pairings/s", "cores": cores, "kind": kind,
               "sample": "first %d pairs of the step's batch, %d processes, %.1f s wall" % (sample_n, cores, wall)}

    import torch
    import torch.distributed as dist
    from pbc_b200.pairing import Pairing, kernel_launches, bench_imad

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    pr = Pairing(PARAMS[name])
    Pp = torch.from_numpy(Ph.copy()).pin_memory()
    Qp = torch.from_numpy(Qh.copy()).pin_memory()
    Op = torch.empty(n * gt, dtype=torch.uint8).pin_memory()
    dP, dQ = Pp.to(dev), Qp.to(dev)
    dO = torch.empty(n * gt, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()

    def step():
        pr.apply_device(dO.data_ptr(), dP.data_ptr(), dQ.data_ptr(), n, st.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    pr.set_stage_profiling(True)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = [0.0, 0.0, 0.0]
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = kernel_launches() - launches0
    barrier()
    # per-kernel durations: one more (untimed-for-value) pass per stage sample, events on the stream
    for _ in range(min(3, args.steps)):
        step()
        torch.cuda.synchronize()
        tms = pr.stage_times()
        stage = [a + b for a, b in zip(stage, tms)]
    stage = [x / min(3, args.steps) for x in stage]
    clocks = sampler.stop()
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- parity against the CPU baseline sample ----
    parity = None
    if cpu_out is not None:
        got = dO[:sample_n * gt].cpu().numpy().tobytes()
        parity = {"checked": sample_n, "bit_exact": got == cpu_out}

    # ---- end to end: host buffers through the C ABI ----
    for _ in range(2):
        pr.apply_into(Op, Pp, Qp, n)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pr.apply_into(Op, Pp, Qp, n)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t2 = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_val = world * n * args.steps / float(t2.item())
    e2e_same = bool((Op[:4096 * gt] == dO[:4096 * gt].cpu()).all().item())

    if rank == 0:
        # ---- roofline of the dominant kernel (Miller loop) ----
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        iters = 3000
        ims = bench_imad(sms * 8, 256, iters, 3)
        peak = sms * 8 * 256 * 32 * iters / (ims * 1e-3)          # IMAD.WIDE.U32 / s, measured live
        unit = UNIT_OPS_PER_MULMOD[name]
        main_ms = stage[0]
        ach_ref = n * REF_MULMODS_MAIN[name] * unit / (main_ms * 1e-3)
        ach_exec = n * EXEC_UNIT_OPS_MAIN[name] / (main_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_summary.json")))["k_a_miller"]["dram_bytes_per_launch"]
        except Exception:
            pass
        hbm_ach = n * sum(WIRE_BYTES[name]) / (sum(stage) * 1e-3) / 1e9
        line = {
            "metric": "pairings/sec", "value": value, "unit": "pairings/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (16 x 32-bit = 512-bit F_q, exact integer)", "data": "synthetic",
            "config": {"workload": workload_name(name), "param": name + ".param", "batch_per_gpu": n,
                       "global_batch": world * n, "parallelism": "shard%d (independent pairs, no collective)" % world,
                       "inputs": "%dx%d grid of seeded subgroup points, all pairs distinct" % (GRID, GRID),
                       "l2": "inputs+workspace %.0f MB per step > 126 MB L2" % ((n * (g1 + g2 + gt + 576)) / 1e6)},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "pairings/s", "h2d_bytes_per_step": n * (g1 + g2),
                    "d2h_bytes_per_step": n * gt, "timer": "perf_counter around the blocking C-ABI call, pinned host buffers",
                    "matches_device_resident_output": e2e_same},
            "gpu_launches": launches,
            "stage_ms": {"miller": stage[0], "batch_invert": stage[1], "final_exp": stage[2]},
            "roofline": {"bound": "int-mul pipe (IMAD.WIDE.U32 issue; SURVEY 8d)", "kernel": "k_a_miller",
                         "achieved": ach_ref / 1e12, "peak": peak / 1e12, "unit": "T IMAD.WIDE.U32/s",
                         "frac": ach_ref / peak,
                         "achieved_executed": ach_exec / 1e12, "frac_executed": ach_exec / peak,
                         "peak_source": "live microkernel k_imad_peak (MEASURED_PEAKS.json has no integer peak)",
                         "work": "reference-equivalent %d mulmods x %d unit ops per pairing; executed %d mul+sqr = %d unit ops"
                                 % (REF_MULMODS_MAIN[name], unit, EXEC_MULMODS_MAIN[name], EXEC_UNIT_OPS_MAIN[name]),
                         "traffic": traffic,
                         "hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_ach / hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                                 "note": "384 wire bytes per pairing: does not bound the path"}},
            "cpu_baseline": cpu,
            "parity": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
