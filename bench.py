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
# --- workloads: BASELINE.json configs[1] (default), [2], [3], [4] ---------------------------------
# unit = one 32x32->64 multiply-accumulate (IMAD.WIDE.U32); a t-limb(64-bit) Montgomery mulmod is
# 2(2t)^2 + 2t of them: 528 for t = 8 (type a), 78 for t = 3 (types f, d)  -- SURVEY.md 8(d).
# ref_mulmods = reference-algorithm mulmods per output (SURVEY 8d probes); ref_main = the part the
# first kernel (Miller loop) stands for.  exec_unit_ops_main = what OUR first kernel executes
# (type A, weight-(1,2) Miller loop: 1616 multiplications of 528 units + 1121 squarings of 408).
WORKLOADS = {
    "a": dict(param="a", mode="single", k=1, n=1 << 20, unit=528, ref_mulmods=4394, ref_main=4394 - 719,
              exec_unit_ops_main=1616 * 528 + 1121 * 408, cpu_rate=1100.0, port_rate=110.0,
              name="type A (param/a.param) element_pairing, batch 2^20 (P,Q) pairs per GPU, 512-bit F_q",
              dtype="u32 limbs (16 x 32-bit = 512-bit F_q, exact integer)", kernels=("k_a_miller", "k_batch_invert", "k_a_finalexp")),
    "f": dict(param="f", mode="single", k=1, n=1 << 20, unit=78, ref_mulmods=98183, ref_main=None,
              exec_unit_ops_main=None, exec_unit_ops_all=883715, cpu_rate=70.0, port_rate=3.0,
              name="type F (param/f.param, BN k=12) element_pairing, batch 2^20 pairs per GPU, 158-bit F_q",
              dtype="u32 limbs (5 x 32-bit = 160-bit F_q, exact integer)", kernels=("k_f_miller", "-", "k_f_finalexp")),
    "d": dict(param="d159", mode="single", k=1, n=1 << 18, unit=78, ref_mulmods=23039, ref_main=None,
              exec_unit_ops_main=None, exec_unit_ops_all=872910, cpu_rate=350.0, port_rate=10.0,
              name="type D (param/d159.param, MNT k=6) element_pairing, batch 2^18 pairs per GPU, 159-bit F_q",
              dtype="u32 limbs (5 x 32-bit = 160-bit F_q, exact integer)", kernels=("k_d_miller", "-", "k_d_finalexp")),
    "g": dict(param="g149", mode="single", k=1, n=1 << 18, unit=78, ref_mulmods=None, ref_main=None,
              exec_unit_ops_main=None, exec_unit_ops_all=2933830, cpu_rate=110.0, port_rate=3.0,
              name="type G (param/g149.param, Freeman k=10) element_pairing, batch 2^18 pairs per GPU, 149-bit F_q",
              dtype="u32 limbs (5 x 32-bit = 160-bit F_q, exact integer)", kernels=("k_g_miller", "-", "k_g_finalexp")),
    "prod16": dict(param="a", mode="prod", k=16, n=1 << 16, unit=528, ref_mulmods=41536, ref_main=41536 - 719,
                   exec_unit_ops_main=16 * (1616 * 528 + 1121 * 408), cpu_rate=130.0, port_rate=8.0,
                   name="type A element_prod_pairing n=16, 2^16 outputs (2^20 Miller loops) over all GPUs",
                   dtype="u32 limbs (16 x 32-bit = 512-bit F_q, exact integer)",
                   kernels=("k_a_miller+k_a_prod", "k_batch_invert", "k_a_finalexp")),
    "pp": dict(param="a", mode="pp", k=1, n=1 << 20, unit=528, ref_mulmods=1838 + 719, ref_main=1838,
               exec_unit_ops_main=(160 * 7 + 5) * 528 + 4 * 408, cpu_rate=400.0, port_rate=110.0,
               name="type A pairing_pp_init + pairing_pp_apply: one fixed first argument, 2^20 second arguments per GPU",
               dtype="u32 limbs (16 x 32-bit = 512-bit F_q, exact integer)",
               kernels=("k_a_pp_init+k_a_pp_apply", "k_batch_invert", "k_a_finalexp")),
}
def _a1_exec_ops():
    """unit ops k_a1_miller executes per pairing: 13 M + 6 S per bit of n, 16 M + 3 S per chord;
    M = 2*34^2 + 34 (operand scanning), S = 34*35/2 + 34^2 + 34 (product scanning) IMAD.WIDE.U32"""
    n = synth.parse_param(PARAMS["a1"])["n"]
    steps = n.bit_length() - 1
    chords = bin(n >> 1).count("1") - 1
    M, S = 2 * 34 * 34 + 34, 34 * 35 // 2 + 34 * 34 + 34
    return steps * (13 * M + 6 * S) + chords * (16 * M + 3 * S)


WORKLOADS["a1"] = dict(
    param="a1", mode="single", k=1, n=148 * 96 * 2, unit=2 * 34 * 34 + 34, ref_mulmods=None, ref_main=None,
    exec_unit_ops_main=_a1_exec_ops(), cpu_rate=48.0, port_rate=6.0,
    name="type A1 (param/a1.param, 1033-bit p, 1022-bit composite-order-capable n) element_pairing, "
         "batch 2 x 148 x 96 pairs per GPU",
    dtype="u32 limbs (34 x 32-bit, 1033-bit F_p, exact integer)",
    kernels=("k_a1_miller", "k_batch_invert", "k_a1_finalexp"))
WIRE = {"a1": (260, 260, 260), "a": (128, 128, 128), "f": (40, 80, 240), "d159": (40, 120, 120), "g149": (38, 190, 190)}


def make_inputs(w, n_out, offset_out=0):
    """(P, Q) numpy uint8 arrays for outputs offset_out .. offset_out+n_out-1 of workload w."""
    prm = synth.parse_param(PARAMS[w["param"]])
    if w["param"] in ("a", "a1"):
        Pb, Qb = synth.type_a_points(prm, GRID if w["param"] == "a" else 512, SEED)
    else:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", w["param"] + ".json")))["pairing"]
        Pb, Qb = synth.type_fd_points(prm, [bytes.fromhex(x) for x in g["P"][:2]],
                                      [bytes.fromhex(x) for x in g["Q"][:2]], GRID)
    return synth.build_batch(Pb, Qb, n_out * w["k"], offset_out * w["k"])


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------
# CPU reference leg (oracle/_ref = the unmodified reference; else the Python port)
# ------------------------------------------------------------------------------------------
def _ref_worker(args):
    name, mode, k, Pb, Qb, n = args
    from oracle import ref as R
    rp = R.RefPairing(PARAMS[name])
    t0, c0 = time.perf_counter(), time.process_time()
    if mode == "pp":
        out = rp.pp_pairing(Pb, Qb, n)
    else:
        out = rp.pairing(Pb, Qb, n) if mode == "single" else rp.prod_pairing(Pb, Qb, k, n)
    return out, time.perf_counter() - t0, time.process_time() - c0


def _port_worker(args):
    name, mode, k, Pb, Qb, n = args
    from oracle import pbc_oracle as O
    pr = O.pairing_from_param(PARAMS[name])
    t0, c0 = time.perf_counter(), time.process_time()
    if mode == "pp":
        out = O.pairing_batch(pr, Pb[:pr.g1_len] * n, Qb, n)
    elif mode == "single":
        out = O.pairing_batch(pr, Pb, Qb, n)
    else:
        a, b = pr.g1_len, pr.g2_len
        out = b"".join(O.prod_pairing_bytes(pr, [Pb[(i * k + j) * a:(i * k + j + 1) * a] for j in range(k)],
                                            [Qb[(i * k + j) * b:(i * k + j + 1) * b] for j in range(k)])
                       for i in range(n))
    return out, time.perf_counter() - t0, time.process_time() - c0


def cpu_kind():
    try:
        from oracle import ref as R
        if R.available():
            return "reference"
    except Exception:
        pass
    return "port"


def cpu_pairings(w, P, Q, n, cores, pool=None):
    """n outputs of workload w on `cores` processes (each with its own pairing_t: the library is
    not thread-safe).  Returns (bytes, wall seconds, CPU seconds summed over the
    workers): CPU seconds / wall is how many cores the box actually delivered."""
    import multiprocessing as mp
    g1, g2, _ = WIRE[w["param"]]
    k = w["k"]
    kind = cpu_kind()
    worker = _ref_worker if kind == "reference" else _port_worker
    per = (n + cores - 1) // cores
    jobs = []
    for c in range(cores):
        lo, hi = c * per, min(n, (c + 1) * per)
        if lo < hi:
            pslice = bytes(P[:g1]) if w["mode"] == "pp" else bytes(P[lo * k * g1:hi * k * g1])
            jobs.append((w["param"], w["mode"], k, pslice, bytes(Q[lo * k * g2:hi * k * g2]), hi - lo))
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(len(jobs))
    t0 = time.perf_counter()
    res = pool.map(worker, jobs)
    wall = time.perf_counter() - t0
    if own:
        pool.close()
        pool.join()
    return b"".join(r[0] for r in res), wall, sum(r[2] for r in res)


def _cgroup_cpu_quota():
    """CPUs the container may use according to its cgroup (v2 cpu.max, v1 cfs quota), or None"""
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max" and float(p) > 0:
            return float(q) / float(p)
    except Exception:
        pass
    try:
        q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0 and p > 0:
            return q / p
    except Exception:
        pass
    return None


def host_cores():
    """worker processes for the CPU legs: the CPUs this process may run on, capped by the
    container's CPU quota when it has one (more runnable processes than quota only thrash)"""
    try:
        cores = max(1, len(os.sched_getaffinity(0)))
    except Exception:
        cores = max(1, os.cpu_count() or 1)
    quota = _cgroup_cpu_quota()
    if quota:
        cores = max(1, min(cores, int(quota + 0.999)))
    return cores


def cpu_rate_guess(w, kind):
    """outputs per second per core, to size the bounded CPU sample"""
    return w["cpu_rate"] if kind == "reference" else w["port_rate"]


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
    w = WORKLOADS[args.workload]
    n_full = args.n or w["n"]
    cores = host_cores()
    kind = cpu_kind()
    per_step = int(min(n_full, max(cores * 4, args.ref_seconds * cpu_rate_guess(w, kind) * cores)))
    import multiprocessing as mp
    P, Q = make_inputs(w, per_step)
    P, Q = P.tobytes(), Q.tobytes()
    pool = mp.get_context("fork").Pool(cores)
    for _ in range(args.warmup):
        cpu_pairings(w, P, Q, min(per_step, cores * 2), cores, pool)   # warm-up
    t0 = time.perf_counter()
    cpu_s = 0.0
    for _ in range(args.steps):
        cpu_s += cpu_pairings(w, P, Q, per_step, cores, pool)[2]
    wall = time.perf_counter() - t0
    pool.close()
    pool.join()
    val = per_step * args.steps / wall
    unit = "pairings/s" if w["mode"] in ("single", "pp") else "outputs/s"
    line = {
        "impl": "reference", "metric": "pairings/sec", "value": val, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "weak" if w["mode"] in ("single", "pp") else "strong",
        "vs_baseline": None, "dtype": w["dtype"] if kind == "port" else "u64 limbs (GMP mpn)",
        "data": "synthetic",
        "config": {"workload": w["name"], "batch_per_step": per_step, "param": w["param"] + ".param",
                   "note": "bounded sample of the workload per step; CPU only"},
        "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": kind,
                         "sample": "%d outputs per step x %d steps, %d processes" % (per_step, args.steps, cores),
                         "cpu_seconds": cpu_s, "effective_cores": cpu_s / max(wall, 1e-9)},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="a", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="outputs per GPU per step (default: the workload's)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget (rank 0, N=1)")
    ap.add_argument("--ref-seconds", type=float, default=3.0, help="--impl reference: seconds per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    w = WORKLOADS[args.workload]
    k, single = w["k"], w["mode"] in ("single", "pp")
    # single pairings: every rank owns its own full batch (weak scaling); the product config is a
    # fixed 2^16 outputs split by output across the ranks (strong scaling)
    n = args.n or (w["n"] if single else max(1, w["n"] // world))
    g1, g2, gt = WIRE[w["param"]]
    unit_name = "pairings/s" if single else "outputs/s"

    # ---- synthetic inputs (host): this rank's shard ----
    Ph, Qh = make_inputs(w, n, offset_out=rank * n)

    # ---- CPU baseline first (forks before CUDA is initialised) ----
    cpu = None
    cpu_out = None
    sample_n = 0
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        kind = cpu_kind()
        sample_n = int(min(n, max(cores * 2, args.cpu_seconds * cpu_rate_guess(w, kind) * cores)))
        cpu_out, wall, cpu_s = cpu_pairings(w, Ph[:sample_n * k * g1].tobytes(), Qh[:sample_n * k * g2].tobytes(),
                                            sample_n, cores)
        cpu = {"value": sample_n / wall, "unit": unit_name, "cores": cores, "kind": kind,
               "sample": "first %d outputs of the step's batch, %d processes, %.1f s wall" % (sample_n, cores, wall),
               "cpu_seconds": cpu_s, "effective_cores": cpu_s / max(wall, 1e-9)}

    import torch
    import torch.distributed as dist
    from pbc_b200.pairing import Pairing, kernel_launches, bench_imad
    from pbc_b200 import _lib as _pbc_lib
    if _pbc_lib.IS_SIMULATOR:
        raise SystemExit("bench.py measures the CUDA library; PBC_B200_LIB points at the test suite's CPU simulator")

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)     # any level >= VERSION prints "NCCL version ..." on stdout; keep it to the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    pr = Pairing(PARAMS[w["param"]])
    Pp = torch.from_numpy(Ph.copy()).pin_memory()
    Qp = torch.from_numpy(Qh.copy()).pin_memory()
    Op = torch.empty(n * gt, dtype=torch.uint8).pin_memory()
    dP, dQ = Pp.to(dev), Qp.to(dev)
    dO = torch.empty(n * gt, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()

    def step():
        if w["mode"] == "pp":
            pr.pp_apply_device(dO.data_ptr(), dP.data_ptr(), dQ.data_ptr(), n, st.cuda_stream)
        elif single:
            pr.apply_device(dO.data_ptr(), dP.data_ptr(), dQ.data_ptr(), n, st.cuda_stream)
        else:
            pr.prod_apply_device(dO.data_ptr(), dP.data_ptr(), dQ.data_ptr(), k, n, st.cuda_stream)

    def host_step():
        if w["mode"] == "pp":
            pr.pp_apply_into(Op, Pp, Qp, n)
        elif single:
            pr.apply_into(Op, Pp, Qp, n)
        else:
            pr.prod_apply_into(Op, Pp, Qp, k, n)

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
    # per-kernel durations: CUDA events recorded by the library on the launching stream
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
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t2 = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_val = world * n * args.steps / float(t2.item())
    chk = min(n, 4096) * gt
    e2e_same = bool((Op[:chk] == dO[:chk].cpu()).all().item())

    if rank == 0:
        # ---- roofline of the dominant kernel ----
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        iters = 3000
        ims = bench_imad(sms * 8, 256, iters, 3)
        peak = sms * 8 * 256 * 32 * iters / (ims * 1e-3)          # IMAD.WIDE.U32 / s, measured live
        unit = w["unit"]
        dom = max(range(3), key=lambda i: stage[i])
        if w["ref_main"] is not None:
            # type A: the Miller kernel, reference-equivalent and executed work both known
            kern, kms = w["kernels"][0], stage[0]
            ach_ref = n * w["ref_main"] * unit / (kms * 1e-3)
            ach_exec = n * w["exec_unit_ops_main"] / (kms * 1e-3) if w["exec_unit_ops_main"] else None
            work = ("reference-equivalent %d mulmods x %d unit ops per output in this kernel; executed %s unit ops"
                    % (w["ref_main"], unit, w["exec_unit_ops_main"]))
        elif w["exec_unit_ops_main"]:
            # no reference probe for this type: the roofline is the work the Miller kernel executes
            kern, kms = w["kernels"][0], stage[0]
            ach_ref = ach_exec = n * w["exec_unit_ops_main"] / (kms * 1e-3)
            work = "executed %d unit ops per output in this kernel (no reference mulmod probe for this type)" % w["exec_unit_ops_main"]
        else:
            # types f, d: SURVEY 8(d) gives the reference's mulmod count for the whole pairing only,
            # so the roofline is taken over the whole kernel sequence (Miller + final exponentiation)
            kern, kms = "+".join(x for x in w["kernels"] if x != "-"), sum(stage)
            ach_ref = n * (w["ref_mulmods"] or 0) * unit / (kms * 1e-3)
            # executed work of the same sequence: 32x32 products per pairing counted by the CPU
            # simulator of the library while it runs these kernels (tests/test_kernels_on_cpu_sim.py)
            ach_exec = n * w["exec_unit_ops_all"] / (kms * 1e-3) if w.get("exec_unit_ops_all") else None
            work = ("reference-equivalent %s mulmods x %d unit ops per pairing over the whole kernel sequence "
                    "(None: SURVEY has no probe for this type, frac is 0), executed %s 32x32 products per pairing; "
                    "dominant kernel %s = %.0f%% of the step"
                    % (w["ref_mulmods"], unit, w.get("exec_unit_ops_all"), w["kernels"][dom],
                       100 * stage[dom] / max(sum(stage), 1e-9)))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = traffic_src = None
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_summary.json")))[w["kernels"][0].split("+")[0]]
            # one ncu --set full capture of this kernel (dram__bytes_read.sum + dram__bytes_write.sum);
            # every thread moves the same bytes, so the capture's per-pairing figure scales to this launch
            traffic = t["dram_bytes_per_pairing"] * n * k
            traffic_src = "%s (n = %d), scaled to this launch's %d Miller loops" % (t["capture"], t["capture_n"], n * k)
        except Exception:
            pass
        hbm_ach = n * (k * (g1 + g2) + gt) / (sum(stage) * 1e-3) / 1e9
        roof = {"bound": "int-mul pipe (IMAD.WIDE.U32 issue; SURVEY 8d)", "kernel": kern,
                "achieved": ach_ref / 1e12, "peak": peak / 1e12, "unit": "T IMAD.WIDE.U32/s",
                "frac": ach_ref / peak,
                "peak_source": "live microkernel k_imad_peak (MEASURED_PEAKS.json has no integer peak)",
                "work": work, "traffic": traffic, "traffic_source": traffic_src,
                "hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_ach / hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "note": "%d wire bytes per output: does not bound the path" % (k * (g1 + g2) + gt)}}
        if ach_exec is not None:
            roof["achieved_executed"] = ach_exec / 1e12
            roof["frac_executed"] = ach_exec / peak
        ws_per = 576 * k if w["param"] == "a" else {"f": 61 * 4, "d159": 31 * 4, "g149": 51 * 4, "a1": 6 * 136}[w["param"]]
        line = {
            "metric": "pairings/sec", "value": value, "unit": unit_name, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak" if single else "strong", "vs_baseline": None,
            "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": w["name"], "param": w["param"] + ".param", "batch_per_gpu": n,
                       "global_batch": world * n, "pairings_per_output": k,
                       "parallelism": "shard%d (independent outputs, no collective)" % world,
                       "inputs": "%dx%d grid of seeded subgroup points, all pairs distinct" % ((GRID, GRID) if w["param"] != "a1" else (512, 512)),
                       "l2": "inputs+outputs+workspace %.0f MB per step vs 126 MB L2"
                             % ((n * (k * (g1 + g2) + gt + ws_per)) / 1e6)},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": unit_name, "h2d_bytes_per_step": n * k * (g1 + g2),
                    "d2h_bytes_per_step": n * gt,
                    "timer": "perf_counter around the blocking C-ABI call, pinned host buffers",
                    "matches_device_resident_output": e2e_same},
            "gpu_launches": launches,
            "stage_ms": dict(zip(("main", "mid", "final_exp"), stage)),
            "stage_kernels": list(w["kernels"]),
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
