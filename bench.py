"""Headline benchmark: one n=40 complex128 permanent (BASELINE.json configs[4]),
Gray-code terms/s and wall time, at 1..8 GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --steps K --warmup W    # the reference C++ on host cores

For N > 1 the driver launches one rank per GPU with torchrun; run by hand with
``--gpus N`` and no torchrun this script re-launches itself that way.

A "step" is one complete permanent of the 40x40 Haar-random unitary: its
2^39-term Gray-code space is split over the N ranks (contiguous segment
ranges), every rank walks its share with the sm_100a kernels, and ONE NCCL
all-gather of four doubles per rank (summed error-free on every rank) combines
the partials (strong scaling: the total
work is fixed).  Rank 0 prints ONE JSON line.

* ``value``   -- terms/s, device-timed (CUDA events on the launching stream
  around exactly K steps, max over ranks), inputs resident in HBM
  (``pq_perm_job_*``), the all-gather inside the timed region.
* ``e2e``     -- the same K steps through the public host-buffer API
  (``piquasso_b200.distributed.permanent_allgather`` == ``permanent`` at N=1):
  host planning, H2D of the matrix, kernels, all-gather, D2H of the result,
  wall clock, max over ranks.
* ``roofline``-- FP64 pipe: algorithmic flops (8n+2 per term) of
  rank 0's walk kernel over its CUDA-event duration, against the DFMA
  throughput measured in this run (MEASURED_PEAKS.json has no FP64 entry).
* ``cpu_baseline`` (N=1) -- the reference's own C++ (oracle/_ref) on this
  box's host cores, on a bounded sample of the same workload.
* ``secondary`` -- BASELINE configs[0..3] measured in the same run
  (bench_secondary.py): n=20 / n=30 permanents and the 60-mode n-ary cases with
  the reference timed on the SAME input (N=1), and the 10^4-shot sampler, whose
  shots are sharded over the ranks at N>1.
"""

from __future__ import annotations

import argparse
import json
import os
import socket
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "n40_c128_permanent_gray_code_terms_per_s"
UNIT = "terms/s"
NOMINAL_FP64_TFLOPS = 37.2  # 148 SM x 64 lanes x 2 flop x 1.965 GHz
# `ncu --set full` capture of the headline instantiation (committed under profiles/):
# roofline.traffic is read from it, never typed in.
TRAFFIC_NCU_CSV = os.path.join(ROOT, "profiles", "r02_ncu_perm_walk_binary_n40_raw.csv")


def traffic_from_ncu(path=TRAFFIC_NCU_CSV):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes, one launch) of the
    walk kernel from the committed ncu raw-page CSV; None when it is absent."""
    import csv
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        with open(path, newline="") as fh:
            rows = list(csv.reader(fh))
        hdr, units, vals = rows[0], rows[1], rows[2]
        total = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            total += float(vals[i].replace(",", "")) * scale[units[i]]
        return total
    except (OSError, ValueError, IndexError, KeyError):
        return None


def haar_matrix(n, seed):
    from scipy.stats import unitary_group
    return np.ascontiguousarray(unitary_group.rvs(n, random_state=seed), dtype=np.complex128)


def workload_config(n):
    return {
        "workload": "BASELINE configs[4]: n=%d complex128 Haar-random unitary permanent "
                    "(scipy unitary_group.rvs(%d, random_state=%d)), all multiplicities 1, "
                    "2^%d Glynn/Gray-code terms" % (n, n, n, n - 1),
        "n": n,
        "terms_per_step": 2 ** (n - 1),
        "flops_per_term": 8 * n + 2,
    }


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own C++ on the host cores
# ---------------------------------------------------------------------------
def reference_sample(n, seed, digits):
    """Bounded sample of the n-column workload the reference can finish (and is
    valid for: idx_max < 2^31): the first `digits`+1 rows of the same matrix,
    `digits` of them with multiplicity 1 and one carrying the remaining
    photons, all n columns with multiplicity 1 -- the same 8n+2 flops per term
    and the same hot loop (src/permanent.cpp:218-250), 2^(digits-1) *
    (n - digits + 1) terms."""
    a = haar_matrix(n, seed)[: digits + 1]
    rows = np.ones(digits + 1, dtype=np.int32)
    rows[-1] = n - digits
    cols = np.ones(n, dtype=np.int32)
    terms = 2 ** (digits - 1) * (n - digits + 1)
    return np.ascontiguousarray(a), rows, cols, terms


def run_reference_steps(n, seed, digits, steps, warmup):
    import oracle
    a, rows, cols, terms = reference_sample(n, seed, digits)
    if oracle.ref_available():
        fn, kind = oracle.ref_permanent, "reference"
        threads = 4 * (os.cpu_count() or 1)   # src/permanent.cpp:145-147
    else:
        nthreads = oracle.num_threads()
        fn = lambda m, r, c: oracle.permanent(m, r, c, njobs=4 * nthreads)  # noqa: E731
        kind, threads = "port", 4 * nthreads
    small = reference_sample(n, seed, 12)
    for _ in range(max(warmup, 1)):
        fn(*small[:3])  # page in the library / spin up the OpenMP pool
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn(a, rows, cols)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {
        "value": terms * steps / total,
        "seconds_per_step": total / steps,
        "kind": kind,
        "cores": os.cpu_count() or 1,
        "threads": threads,
        "sample": "%dx%d slice of the same Haar matrix, rows=[1]*%d+[%d], cols=[1]*%d: "
                  "%d Gray-code terms per step (same 8n+2 flops/term; the full 2^%d-term "
                  "n=%d permanent is outside the reference's valid range, idx_max > 2^31)"
                  % (digits + 1, n, digits, n - digits, n, terms, n - 1, n),
        "terms_per_step": terms,
    }


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    res = run_reference_steps(args.n, args.seed, args.ref_digits, args.steps, args.warmup)
    cfg = workload_config(args.n)
    cfg["sample"] = res["sample"]
    line = {
        "impl": "reference",
        "metric": METRIC, "value": res["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"],
                         "threads": res["threads"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(self.NAMES, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax),
                "power_w_max": max(power), "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# this implementation
# ---------------------------------------------------------------------------
def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def main_arm(args):
    import ctypes

    import torch
    import torch.distributed as dist

    from piquasso_b200 import _lib
    from piquasso_b200.distributed import combine, finish, permanent_allgather

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback "
                         "(use --impl reference for the host-core arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nccl_log_dir = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own log is the evidence of the communicator (ranks, transport): keep
        # it visible.  Whatever the launcher set is left alone; otherwise the INIT
        # lines go to stderr, so that stdout stays the one JSON line.
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            # unset, or a level that says nothing about the communicator (the image
            # presets VERSION): raise it to INFO for the INIT subsystem only
            import tempfile
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
            if "NCCL_DEBUG_FILE" not in os.environ:
                # one file per rank (NCCL re-opens whatever it is given, so it must not
                # be /dev/stderr); rank 0 replays all of them at the end
                nccl_log_dir = os.path.join(tempfile.gettempdir(),
                                            "pq_nccl_%s" % os.environ.get("MASTER_PORT", "0"))
                os.makedirs(nccl_log_dir, exist_ok=True)
                os.environ["NCCL_DEBUG_FILE"] = os.path.join(nccl_log_dir, "rank%d.log" % rank)
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    n, K, W = args.n, args.steps, args.warmup
    a = haar_matrix(n, args.seed)
    ones = np.ones(n, dtype=np.int32)
    terms_total = 2 ** (n - 1)

    peak = lib.pq_fp64_peak_tflops(local_rank, 1 << 17)

    job = ctypes.c_void_p()
    status = ctypes.c_int(0)
    triv = np.zeros(2)
    _lib.check(lib.pq_perm_job_create_c128(
        a.ctypes.data_as(_lib.c_double_p), n, n, ones.ctypes.data_as(_lib.c_int32_p),
        ones.ctypes.data_as(_lib.c_int32_p), rank, world, local_rank, ctypes.byref(job),
        ctypes.byref(status), triv.ctypes.data_as(_lib.c_double_p)))
    info = _lib.PlanInfo()
    _lib.check(lib.pq_perm_job_info(job, ctypes.byref(info)))
    my_terms = lib.pq_perm_job_terms(job)

    partial = torch.zeros(4, dtype=torch.float64, device=dev)
    gathered = torch.zeros(4 * world, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def step():
        flush.zero_()
        _lib.check(lib.pq_perm_job_launch(job, ctypes.c_void_p(stream.cuda_stream),
                                          ctypes.c_void_p(partial.data_ptr())))
        if world > 1:
            dist.all_gather_into_tensor(gathered, partial)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = lib.pq_launch_count()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(K):
        step()
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = lib.pq_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    result = finish(combine(gathered.cpu().numpy()) if world > 1 else partial.cpu().numpy(), n)

    hist = np.zeros(max(K, 1))
    nh = lib.pq_kernel_ms_history(local_rank, hist.ctypes.data_as(_lib.c_double_p), K)
    kernel_ms = float(np.mean(hist[:nh])) if nh else float("nan")

    t = torch.tensor([elapsed_ms, kernel_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms = float(tmax[0])
        launches_total = int(tsum[2])
    else:
        launches_total = int(launches)

    # ---- end to end through the public host-buffer API ------------------------
    permanent_allgather(a, ones, ones, device_index=local_rank)  # one untimed call
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_result = permanent_allgather(a, ones, ones, device_index=local_rank)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    h2d = int((info.active_rows + 1) * info.cols_padded * 16) * world
    d2h = 32 * world

    # ---- the other BASELINE configs, same run (all ranks: the sampler is sharded) ----
    secondary = None
    if not args.no_secondary:
        import bench_secondary
        t0 = time.perf_counter()
        secondary = bench_secondary.run(lib, peak, rank, world, local_rank,
                                        with_reference=(world == 1 and rank == 0),
                                        sampler_shots=args.sampler_shots)
        if secondary is not None:
            secondary["seconds_spent"] = time.perf_counter() - t0

    line = None
    if rank == 0:
        ms_per_step = elapsed_ms / K
        value = terms_total * K / (elapsed_ms * 1e-3)
        flops_launch = float(info.flops_per_term) * float(my_terms)
        achieved = flops_launch / (kernel_ms * 1e-3) / 1e12
        cfg = workload_config(n)
        cfg.update({
            "partition": "%d segments of %d terms, contiguous 1/%d share per rank; one NCCL "
                         "all-gather (4 x f64 per rank) + error-free sum per step" % (info.nseg, info.seg_len, world),
            "kernel": {1: "generic n-ary walk", 2: "binary constant-bank walk"}[info.kernel],
            "l2": "256 MiB memset between steps (inside the timed region; the path's working "
                  "set is the %d-byte matrix, not HBM-resident data)" % (h2d // world),
        })
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_per_step, "wall_time_s_per_permanent": ms_per_step / 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "result": [result.real, result.imag],
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak > 0 else None,
                "traffic": traffic_from_ncu(),
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one "
                                  "perm_walk_binary_pm<40,3,64> launch, ncu --set full, "
                                  "profiles/" + os.path.basename(TRAFFIC_NCU_CSV),
                "peak_source": "measured in this run on rank 0's GPU: 16 independent DFMA "
                               "chains per thread with distinct operands, 2 flop per FMA "
                               "(pq_fp64_peak_tflops; 62-63 of the 64 FMA/clk/SM); "
                               "MEASURED_PEAKS.json holds no FP64 figure",
                "nominal_peak": NOMINAL_FP64_TFLOPS,
                "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
                "kernel_ms": kernel_ms,
                "flops_per_launch": flops_launch,
                "note": "per GPU, rank 0's walk kernel; an FP64 instruction mix of 2 DFMA-class "
                        "adds + 2 DMUL + 2 DFMA per column caps algorithmic flops at 67.6% of "
                        "DFMA peak",
            },
            "e2e": {"value": terms_total * K / e2e_s, "unit": UNIT,
                    "ms_per_step": e2e_s / K * 1e3,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "result": [complex(e2e_result).real, complex(e2e_result).imag]},
            "gpu_launches": launches_total,
            "clocks": clocks,
        }
        if world == 1:
            cb = run_reference_steps(n, args.seed, args.ref_digits, 1, 1)
            line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"],
                                    "threads": cb["threads"], "kind": cb["kind"],
                                    "sample": cb["sample"],
                                    "seconds": cb["seconds_per_step"]}
        else:
            line["cpu_baseline"] = None
        line["secondary"] = secondary
    lib.pq_perm_job_destroy(job)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        if world > 1 and nccl_log_dir:
            # NCCL's INIT log of every rank (communicator size, transports) on both
            # streams, BEFORE the result: the JSON line stays the last line of stdout
            time.sleep(0.5)
            for r in range(world):
                try:
                    with open(os.path.join(nccl_log_dir, "rank%d.log" % r)) as fh:
                        text = fh.read()
                except OSError:
                    continue
                sys.stderr.write(text)
                sys.stdout.write(text)
            sys.stderr.flush()
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=40, help="matrix size (headline: 40)")
    ap.add_argument("--seed", type=int, default=40)
    ap.add_argument("--ref-digits", type=int, default=25,
                    help="binary Gray digits of the reference's bounded sample")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the other BASELINE configs (the `secondary` object)")
    ap.add_argument("--sampler-shots", type=int, default=10000,
                    help="shots of the configs[3] sampler run in `secondary`")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(free_port()), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return main_arm(args)


if __name__ == "__main__":
    sys.exit(main())
