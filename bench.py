#!/usr/bin/env python
"""bench.py -- Stark252 coset-LDE + Merkle commit throughput on B200 (BASELINE.json metric).

One step = the commitment phase of one Cairo fib-70k-shaped proof (BASELINE.json configs[1],
SURVEY.md section 8 config C2): N = 2^19 rows, blowup 4 (M = 2^21), coset offset 3:
    interpolate_and_commit(main trace, 34 columns)          prover.rs:126-159
    interpolate_and_commit(aux trace, 18 columns)           prover.rs:208
    round-2 LDE + commit of H1, H2 (2 columns)              prover.rs:254-276
    fri_commit_phase(19 layers from 2^21) + grinding(20)    fri/mod.rs:20-72, grinding.rs:40-48
`elems` = field elements that end up under a Merkle root in that step
        = M*(34+18+2) LDE values + sum of the 19 FRI layer sizes.

`value`  : elems/s with the inputs already resident in HBM (S252_DEVICE buffers).
`e2e`    : the same step through the reference-facing C ABI with HOST buffers (S252_HOST): every call gets
           a pinned host table and uploads it itself inside the timed region (column groups streamed over
           PCIe under the transforms of the previous group); roots / last value / nonce are read back.
           `e2e.prefetch_pipeline` is the extra a streaming prover can have on top (next step's inputs
           prefetched on the copy stream under the current step).
Timing   : CUDA events on the library's stream, W >= 3 warm-up steps, max over ranks; every step
           streams ~8 GB through HBM (inputs/outputs far larger than the 126 MB L2).
N > 1    : one process per GPU, each proving an independent trace (weak scaling, no data-path
           collective) -- see DESIGN.md "multi-GPU".
--impl reference : the CPU restatement of the reference (oracle/, threads over columns as under the
           reference's `parallel` feature) at the SAME size (N = 2^19), as many steps as fit the time budget.
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

LOG_N, BLOWUP, OFFSET, GRIND = 19, 4, 3, 20
COLS_MAIN, COLS_AUX, COLS_COMP = 34, 18, 2
METRIC = "stark252_lde_merkle_commit_elems_per_s"
UNIT = "elems/s"


def workload_config(log_n):
    n = 1 << log_n
    m = n * BLOWUP
    fri_elems = sum(m >> k for k in range(log_n))
    elems = m * (COLS_MAIN + COLS_AUX + COLS_COMP) + fri_elems
    return {
        "workload": "C2 cairo-fib-70k commit phase: N=2^%d rows, blowup %d, 34 main + 18 aux + 2 composition columns, "
                    "%d FRI layers + grinding %d" % (log_n, BLOWUP, log_n, GRIND),
        "trace_rows": n, "lde_rows": m, "columns": [COLS_MAIN, COLS_AUX, COLS_COMP], "blowup": BLOWUP,
        "coset_offset": OFFSET, "fri_layers": log_n, "grinding_factor": GRIND, "elems_per_step": elems,
        "l2_policy": "inputs and outputs of every kernel exceed L2 (126 MB); no flush needed",
        "parallelism": "one independent trace per GPU",
    }


def splitmix_felts(seed, count):
    """count field elements in the reference's LW layout, seeded."""
    idx = np.arange(1, 4 * count + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    z = z.reshape(count, 4)
    z[:, 0] &= np.uint64((1 << 59) - 1)     # < 2^251 < p
    return z


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.stop_flag = device, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ GPU arm
PEAK_NOTE = ("IMAD.WIDE.U32 (32x32->64) products with data-dependent multiplicands: 32 lanes/clk/SM on B200, i.e. one warp "
             "instruction per 4 cycles per SM sub-partition.  Round 1 quoted 18.3 T: that probe multiplied loop-invariant registers, "
             "ptxas hoisted the products out of the loop and the loop timed IADD3 pairs (DESIGN.md 3.1)")


def integer_peaks(ctx):
    """The roofline denominators MEASURED_PEAKS.json does not carry: measured live on this GPU (about 0.2 s, outside every timed
    region) with s252_microbench_int_pipes; profiles/int_peaks.json (an earlier run of tools/microbench.py on this pool) is the
    fallback."""
    import ctypes as C
    from lambdaworks_cairo_prover_b200 import _native as N
    peaks, src = {}, "profiles/int_peaks.json (tools/microbench.py on this pool's B200)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "int_peaks.json")))
    except Exception:
        pass
    try:
        arr = (C.c_double * 8)()
        ctx.check(N.lib().s252_microbench_int_pipes(ctx.handle, arr))
        if arr[0] > 0 and arr[1] > 0:
            peaks = dict(peaks, imad_wide_gops=arr[0], lop3_gops=arr[1], shf_gops=arr[2], imad_wide_plus_lop3_gops=arr[4],
                         imad_wide_carry_rows_gops=arr[5], imad_lo_gops=arr[6])
            src = "measured live in this run (s252_microbench_int_pipes, before the timed region)"
    except Exception:
        pass
    return float(peaks.get("imad_wide_gops", 9250.0)), float(peaks.get("lop3_gops", 18500.0)), src


_LAUNCH_AFFINITY = None


def restore_launch_affinity():
    """The CPU baselines run on every core the process was launched with, not on the GPU-local subset."""
    if _LAUNCH_AFFINITY:
        try:
            os.sched_setaffinity(0, _LAUNCH_AFFINITY)
        except Exception:
            pass


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU before any host table is allocated: pinned host memory then
    sits on the GPU's NUMA node and the uploads of several ranks do not share one socket's memory controllers and inter-socket
    link.  Returns the number of CPUs of the set (0: left as launched)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        global _LAUNCH_AFFINITY
        _LAUNCH_AFFINITY = os.sched_getaffinity(0)
        cpus &= _LAUNCH_AFFINITY
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import lambdaworks_cairo_prover_b200 as P
    from lambdaworks_cairo_prover_b200 import _native as N
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    numa_cpus = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    log_n = args.log_n
    cfg = workload_config(log_n)
    n, m = cfg["trace_rows"], cfg["lde_rows"]
    ctx = P.Context(local_rank)
    L = N.lib()
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    # synthetic inputs (seed 0xB200 + config index 2; every rank its own trace)
    seed = 0xB200 + 2 + 1000 * rank
    host = {
        "main": splitmix_felts(seed, n * COLS_MAIN),
        "aux": splitmix_felts(seed + 1, n * COLS_AUX),
        "comp": splitmix_felts(seed + 2, n * COLS_COMP),
        "p0": splitmix_felts(seed + 3, n),
    }
    pinned = {}
    for k, v in host.items():
        t = torch.empty(v.nbytes, dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = v.reshape(-1).view(np.uint8)
        pinned[k] = t
    dev = {}
    for k, v in host.items():
        p = ctx.device_alloc(v.nbytes)
        ctx.to_device(p, v)
        dev[k] = p
    h2d_bytes = sum(v.nbytes for v in host.values())
    from lambdaworks_cairo_prover_b200 import felt
    offset_fe = felt.from_int(OFFSET)

    def step(mem):
        src = dev if mem == N.DEVICE else {k: t.data_ptr() for k, t in pinned.items()}   # noqa: F821
        tr = P.DefaultTranscript()
        root = np.empty(32, dtype=np.uint8)
        handles = []
        d2h = 0
        for key, cols in (("main", COLS_MAIN), ("aux", COLS_AUX)):
            h = C.c_void_p()
            ctx.check(L.s252_interpolate_and_commit(ctx.handle, C.c_void_p(src[key]), n, cols, BLOWUP, OFFSET, mem,
                                                    C.byref(h), N.ptr(root)))
            handles.append(h)
            tr.append(root.tobytes())
            d2h += 32
        h = C.c_void_p()
        ctx.check(L.s252_lde_and_commit(ctx.handle, C.c_void_p(src["comp"]), n, COLS_COMP, n, BLOWUP, OFFSET, mem,
                                        C.byref(h), N.ptr(root)))
        handles.append(h)
        tr.append(root.tobytes())
        d2h += 32
        fh = C.c_void_p()
        last = np.empty(4, dtype=np.uint64)
        roots = np.empty((log_n, 32), dtype=np.uint8)
        ctx.check(L.s252_fri_commit_phase(ctx.handle, log_n, C.c_void_p(src["p0"]), n, tr.handle, N.ptr(offset_fe), m, mem,
                                          C.byref(fh), N.ptr(last), N.ptr(roots)))
        d2h += 32 * log_n + 32 * BLOWUP
        nonce = P.generate_nonce_with_grinding(tr.challenge(), GRIND, ctx)
        d2h += 8
        for hh in handles:
            L.s252_commit_destroy(hh)
        L.s252_fri_destroy(fh)
        return root.tobytes(), last.copy(), nonce, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(mem, steps, warmup, profile):
        for _ in range(warmup):
            step(mem)
        barrier()
        if profile:
            ctx.profile(True, reset=True)
        launches0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()            # one sampler per job: N nvidia-smi pollers would perturb the run
        e0.record(stream)
        out = None
        for _ in range(steps):
            out = step(mem)
        e1.record(stream)
        ctx.synchronize()
        barrier()
        sampler.stop_flag = True
        ms = e0.elapsed_time(e1)
        prof = ctx.profile_read() if profile else None
        if profile:
            ctx.profile(False)
        if rank == 0:
            sampler.join(timeout=2)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, prof, ctx.launch_count - launches0, sampler.summary()

    ms_dev, out_dev, prof, launches, clocks = timed(N.DEVICE, args.steps, args.warmup, True)

    # e2e: every step's inputs are copied from pinned host memory inside the timed region.  The
    # copies of step i+1 are queued on the library's copy stream while step i computes
    # (double-buffered device staging), the way a prover that streams traces would run.
    staging = [{k: ctx.device_alloc(v.nbytes) for k, v in host.items()} for _ in range(2)]

    def prefetch(slot):
        for k, t in pinned.items():
            ctx.to_device_async(staging[slot][k], t.data_ptr(), t.numel())

    def timed_e2e(steps, warmup):
        nonlocal dev
        saved = dev
        prefetch(0)
        for i in range(warmup):
            ctx.copy_stream_wait()
            prefetch((i + 1) % 2)
            dev = staging[i % 2]
            step(N.DEVICE)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        # slot (warmup % 2) already holds a prefetched copy issued during warm-up: re-issue it inside
        # the timed region so that exactly `steps` uploads are timed
        prefetch(warmup % 2)
        out = None
        for i in range(steps):
            slot = (warmup + i) % 2
            ctx.copy_stream_wait()
            if i + 1 < steps:
                prefetch(1 - slot)
            dev = staging[slot]
            out = step(N.DEVICE)
        e1.record(stream)
        ctx.synchronize()
        barrier()
        dev = saved
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    ms_e2e, out_e2e = timed_e2e(args.steps, max(args.warmup, 1))
    ms_sync, out_sync, _, _, _ = timed(N.HOST, max(1, min(args.steps, 3)), 3, False)
    assert out_sync[0] == out_dev[0] and out_sync[2] == out_dev[2], "host-buffer path disagrees"
    assert out_dev[0] == out_e2e[0] and out_dev[2] == out_e2e[2], "device-resident and host-buffer paths disagree"

    cairo_line = None if args.no_cairo else cairo_prove_bench(ctx, args, world, rank, barrier, dist if world > 1 else None)

    elems = cfg["elems_per_step"] * world
    value = elems * args.steps / (ms_dev * 1e-3)
    e2e_value = elems * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        tname, tstat = top
        achieved = tstat["bytes"] / (tstat["ms"] * 1e-3) / 1e9 if tstat["ms"] > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(tname)
        except Exception:
            pass
        total_kernel_ms = sum(v["ms"] for v in prof.values())
        imad_peak, lop_peak, int_peak_src = integer_peaks(ctx)
        kernels = {}
        for name, st in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            sec = st["ms"] * 1e-3
            kernels[name] = {
                "launches_per_step": st["launches"] / args.steps, "ms_per_step": st["ms"] / args.steps,
                "share": st["ms"] / total_kernel_ms if total_kernel_ms else 0.0,
                "hbm_gbs": st["bytes"] / sec / 1e9 if sec else 0.0,
                "imad_frac": (st["muls"] * 80 / sec / 1e9) / imad_peak if sec else 0.0,
                "keccak_gperms": st["perms"] / sec / 1e9 if sec else 0.0,
                # Keccak-f[1600] here is 24 x (122 LOP3 + 58 SHF) = 4320 ALU-pipe lane-operations
                "alu_frac": (st["perms"] * 4320 / sec / 1e9) / lop_peak if sec else 0.0,
            }
        # time-weighted fraction of the binding integer pipe over the whole step: IMAD.WIDE issue for the field
        # kernels, LOP3/SHF issue for the Keccak kernels (whichever is larger for the kernel)
        step_frac = sum(k["share"] * max(k["imad_frac"], k["alu_frac"]) for k in kernels.values())
        sync_steps = max(1, min(args.steps, 3))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (256-bit Montgomery) / u64 Keccak lanes", "data": "synthetic (splitmix64-seeded field elements)",
            "config": cfg,
            "e2e": {"value": elems * sync_steps / (ms_sync * 1e-3), "unit": UNIT, "ms_per_step": ms_sync / sync_steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": out_e2e[3], "steps": sync_steps,
                    "how": "the plugin calls with HOST buffers: s252_interpolate_and_commit / s252_lde_and_commit / s252_fri_commit_phase "
                           "(S252_HOST) on pinned host tables; each call streams its own table over PCIe in column groups under the "
                           "transforms of the previous group; roots/last value/nonce read back",
                    "prefetch_pipeline": {"value": e2e_value, "ms_per_step": ms_e2e / args.steps,
                                          "how": "extra: the NEXT step's inputs prefetched with s252_copy_to_device_async (copy stream, "
                                                 "double-buffered) under the current step's S252_DEVICE calls"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "host_cpus_bound": numa_cpus,
            "roofline": {"bound": "int_issue", "kernel": tname, "achieved": tstat["muls"] * 80 / (tstat["ms"] * 1e-3) / 1e9 if tstat["ms"] else 0.0,
                         "peak": imad_peak, "unit": "G lane-op/s (IMAD.WIDE.U32)",
                         "frac": kernels[tname]["imad_frac"], "traffic": traffic,
                         "peak_source": int_peak_src + "; not in MEASURED_PEAKS.json, which holds HBM and bf16 peaks only",
                         "peak_note": PEAK_NOTE,
                         "work": "80 wide multiply-adds per field multiplication (SURVEY 8d) x the multiplications of the launch (DESIGN.md section 4)",
                         "hbm": {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src}},
            "int_roofline": {"bound": "integer issue (IMAD.WIDE for field kernels, LOP3/SHF for Keccak kernels)",
                             "step_frac": step_frac,
                             "step_frac_how": "sum over kernels of (share of step time) x max(imad_frac, alu_frac); peaks are the "
                                              "isolated-instruction rates below",
                             "imad_wide_peak_gops": imad_peak, "lop3_peak_gops": lop_peak,
                             "peak_source": int_peak_src,
                             "kernels": kernels},
            "result": {"last_root": out_dev[0].hex(), "nonce": out_dev[2]},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_sample(args.cpu_log_n or 17)
        if cairo_line is not None:
            line["cairo_prove"] = cairo_line
            if not args.no_cpu_baseline and world == 1:
                restore_launch_affinity()
                line["cairo_prove"]["cpu_baseline"] = cpu_cairo_prove_sample(args.cpu_fib_n)
    for p in list(dev.values()) + [q for s in staging for q in s.values()]:
        ctx.device_free(p)
    if rank == 0:
        if world == 1 and not args.no_c4:
            # the strong-scaling workload of the N > 1 runs (ONE C4 commit) on this one GPU: the N = 1 point of that curve
            ctx.trim()
            try:
                res = subprocess.run([sys.executable, os.path.abspath(__file__), "--mode", "sharded", "--steps", "3", "--warmup", "3", "--no-cairo"],
                                     capture_output=True, text=True, timeout=900)
                c4 = json.loads(res.stdout.strip().splitlines()[-1])
                line["c4_one_gpu"] = {k: c4[k] for k in ("value", "unit", "ms_per_step", "scaling", "parity_ok", "e2e", "stages_ms", "config", "result")}
                # the N = 1 points of the other two one-object-on-N-GPUs measurements of the N > 1 line
                for src, dst in (("c3_one_column", "c3_one_gpu"), ("c5_fri", "c5_one_gpu")):
                    if src in c4:
                        line[dst] = c4[src]
            except Exception as e:
                line["c4_one_gpu"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cairo_prove_bench(ctx, args, world, rank, barrier, dist):
    """The second half of BASELINE.json's metric: "Cairo fib prove time".  generate_cairo_proof on the
    regenerated fib(1,1,70000) trace (benches/criterion_prover_70k.rs: ProofOptions::new_secure(
    Provable80Bits, 3) = blowup 4, 80 queries, grinding 20), host table in, serialized proof out;
    wall clock around the C ABI call (host orchestration, the 570 MB upload and the read-backs included).
    With default_test_options the proof must be byte-identical to the reference's golden file."""
    import torch

    import lambdaworks_cairo_prover_b200 as P
    from lambdaworks_cairo_prover_b200 import _native as N, cairo
    fib_n = args.fib_n
    t0 = time.perf_counter()
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(fib_n))
    trace = cairo.build_main_trace(regs, mem, size)
    front_s = time.perf_counter() - t0
    N.lib().s252_cairo_trace_pin(trace.handle)
    identical = None
    golden = os.path.join(ROOT, "tests", "golden", "reference_proofs", "fibonacci_%d.proof" % fib_n)
    if fib_n == 70000 and os.path.exists(golden):
        raw = open(golden, "rb").read()
        ln = int.from_bytes(raw[:8], "big")
        identical = cairo.generate_cairo_proof(trace, P.ProofOptions.default_test_options(), ctx) == raw[8:8 + ln]
    opts = P.ProofOptions.new_secure("Provable80Bits", 3)
    for _ in range(max(args.warmup, 1)):
        proof = cairo.generate_cairo_proof(trace, opts, ctx)
    barrier()
    times = []
    for _ in range(args.steps):
        a = time.perf_counter()
        proof = cairo.generate_cairo_proof(trace, opts, ctx)
        ctx.synchronize()
        times.append((time.perf_counter() - a) * 1e3)
    stages = json.loads(N.lib().s252_cairo_last_prove_stages().decode())
    barrier()
    ms = float(np.median(times))
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"metric": "cairo_fib_prove_time", "unit": "ms", "higher_is_better": False, "value": ms,
            "proofs_per_s": world * 1e3 / ms, "n_gpus": world, "program": "cairo0 fibonacci_%d" % fib_n, "trace_rows": trace.n_rows(),
            "columns": [trace.n_cols, 18, 2], "options": {"blowup": opts.blowup_factor, "fri_queries": opts.fri_number_of_queries,
                                                          "coset_offset": opts.coset_offset, "grinding": opts.grinding_factor},
            "proof_bytes": len(proof), "ms_all": [round(x, 2) for x in times], "stages_ms": stages,
            "h2d_bytes": int(trace.n_rows() * trace.n_cols * 32), "front_end_s": round(front_s, 2),
            "golden_proof_byte_identical": identical,
            "how": "wall clock around s252_cairo_prove (C ABI): pinned host trace table in, StarkProof::serialize bytes out; "
                   "N>1: one independent proof per GPU, max over ranks"}


def fib_trace_table(fib_n):
    """The main trace of fibonacci_<n> as (table[n_rows, n_cols, 4], public inputs), built by the library's host front-end in a
    CHILD process: the process that times the CPU arm never loads the product library."""
    import pickle
    import tempfile
    code = ("import sys, pickle, numpy as np; sys.path.insert(0, %r)\n"
            "from lambdaworks_cairo_prover_b200 import cairo\n"
            "regs, mem, size = cairo.run_program(cairo.fibonacci_program(%d))\n"
            "t = cairo.build_main_trace(regs, mem, size)\n"
            "table = np.array(t.table).reshape(t.n_rows(), t.n_cols, 4)\n"
            "p = t.pub_inputs\n"
            "pub = {k: getattr(p, k) for k in ('pc_init', 'ap_init', 'fp_init', 'pc_final', 'ap_final', 'num_steps', 'range_check_min', "
            "'range_check_max', 'public_memory')}\n"
            "pickle.dump((table, pub), open(sys.argv[1], 'wb'), protocol=4)\n" % (ROOT, fib_n))
    from types import SimpleNamespace
    with tempfile.NamedTemporaryFile(suffix=".pkl") as f:
        subprocess.run([sys.executable, "-c", code, f.name], check=True)
        table, pub = pickle.load(open(f.name, "rb"))
    return table, SimpleNamespace(**pub)


def cpu_cairo_prove_sample(fib_n):
    """The CPU restatement of prove::<CairoAIR> (oracle/, pinned byte-for-byte on the reference's golden
    proof) on a bounded sample: a shorter fibonacci program, same options."""
    from types import SimpleNamespace
    from oracle.cairo_prover import cairo_prove
    table, pub = fib_trace_table(fib_n)
    # ProofOptions::new_secure(SecurityLevel::Provable80Bits, 3), src/starks/proof/options.rs:67-73
    opts = SimpleNamespace(blowup_factor=4, fri_number_of_queries=80, coset_offset=3, grinding_factor=20)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    cairo_prove(table, pub, opts, threads=cores)
    dt = time.perf_counter() - t0
    rows = table.shape[0]
    return {"value": dt * 1e3, "unit": "ms", "cores": cores, "kind": "port", "trace_rows": rows,
            "sample": "fibonacci_%d (2^%d rows instead of 2^19), same options; oracle/ restatement of the reference prover "
                      "(C kernels, python round structure; LDE and constraint evaluation threaded)" % (fib_n, rows.bit_length() - 1)}


C4_LOG_N, C4_COLS, C4_BLOWUP, C4_SEED = 22, 33, 8, 0xB2040000


def c4_config(log_n, world):
    n = 1 << log_n
    m = n * C4_BLOWUP
    return {
        "workload": "C4 (BASELINE configs[3]): ONE interpolate_and_commit of a synthetic 2^%d-row x %d-column Cairo-layout trace, blowup %d, "
                    "columns sharded over %d GPU(s)" % (log_n, C4_COLS, C4_BLOWUP, world),
        "trace_rows": n, "lde_rows": m, "columns": [C4_COLS], "blowup": C4_BLOWUP, "coset_offset": OFFSET, "elems_per_step": m * C4_COLS,
        "l2_policy": "inputs and outputs of every kernel exceed L2 (126 MB); no flush needed",
        "parallelism": "one trace: columns sharded for iNTT + coset LDE (no communication), all-to-all (grouped NCCL point-to-point chunks, "
                       "pipelined per column group under the LDE of the next group) to row blocks, per-GPU leaf hashing + subtree, subtree roots "
                       "all-gathered, top levels on every rank",
    }


def golden_case(name):
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "synthetic", "big_configs.json")))["cases"].get(name)
    except Exception:
        return None


def run_gpu_sharded(args):
    """N > 1 (default): ONE trace on N GPUs, strong scaling -- the C4 commit (north_star's 2^22-row target), then ONE Cairo
    proof sharded over the same GPUs.  `value`: shards resident in HBM when the timed region starts; `e2e`: pinned host shards, upload
    inside.  The root must equal the one the CPU oracle pinned offline (tests/golden/synthetic/big_configs.json)."""
    import torch
    import torch.distributed as dist

    import lambdaworks_cairo_prover_b200 as P
    from lambdaworks_cairo_prover_b200 import distributed as D
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import random_felts

    world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % (29400 + os.getpid() % 500), rank=0, world_size=1,
                                device_id=torch.device("cuda", local_rank))
    log_n = args.c4_log_n
    cfg = c4_config(log_n, world)
    n, m = cfg["trace_rows"], cfg["lde_rows"]
    ctx = P.Context(local_rank)
    backend = D.GpuBackend(ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    a, b = D.column_shards(C4_COLS, world)[rank]
    groups = max(1, min(args.pipeline_groups, b - a))
    ranges = D.group_ranges(b - a, groups)
    host, dev = [], []
    for lo, hi in ranges:                                     # TraceTable::get_cols per pipeline group (set-up, not timed)
        t = torch.empty((n, hi - lo, 4), dtype=torch.int64, pin_memory=True)
        v = t.numpy().view(np.uint64)
        for j in range(lo, hi):
            v[:, j - lo, :] = random_felts(C4_SEED + a + j, n)
        host.append(t)
        dev.append(t.to(torch.device("cuda", local_rank)))

    def step(tables, timings=None):
        tr = P.DefaultTranscript()
        shards = D.column_shards(C4_COLS, world)
        all_ranges = [D.group_ranges(hi - lo, groups) for lo, hi in shards]
        with D.backend_scope(backend):
            sc = D.exchange_and_commit(backend.lde_pipeline(tables, n, C4_BLOWUP, OFFSET), all_ranges, shards, m, C4_COLS, tr, backend,
                                       exchange="p2p", timings=timings)
        root = sc.root
        sc.free()
        return root

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(tables, steps, warmup, profile, step=step):
        for _ in range(warmup):
            step(tables)
        barrier()
        if profile:
            ctx.profile(True, reset=True)
        launches0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank)
        if rank == 0 and profile:
            sampler.start()
        t0 = time.perf_counter()
        e0.record(stream)
        root = None
        for _ in range(steps):
            root = step(tables)
        e1.record(stream)
        ctx.synchronize()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        sampler.stop_flag = True
        ms = e0.elapsed_time(e1)
        prof = ctx.profile_read() if profile else None
        if profile:
            ctx.profile(False)
        tt = torch.tensor([ms, wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0 and profile:
            sampler.join(timeout=2)
        return float(tt[0].item()), float(tt[1].item()), root, prof, ctx.launch_count - launches0, sampler.summary()

    ms_dev, wall_dev, root, prof, launches, clocks = timed(dev, args.steps, args.warmup, True)
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e, wall_e2e, root_e2e, _, _, _ = timed(host, e2e_steps, 2, False)
    stages = {}
    step(dev, stages)                                          # one more, instrumented (host waits at every mark)
    stage_t = torch.tensor([stages.get(k, 0.0) for k in ("lde", "exchange", "hash", "roots")], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stage_t, op=dist.ReduceOp.MAX)
    golden = golden_case("c4" if log_n == C4_LOG_N else "c4_small" if log_n == 14 else "")
    parity_ok = None if golden is None else (root.hex() == golden["root"] and root_e2e == root)
    # ---- the same commit through the library's own collective entry point (s252_interpolate_and_commit_sharded: NCCL called from
    # C++, no torch.distributed on the data path).  When it works it is the headline: it is the call a non-Python host makes.
    torch_path = None
    native_err = None
    torch.cuda.empty_cache()                                   # the row-block buffers of the path above (torch's allocator) go back to the driver
    try:
        from lambdaworks_cairo_prover_b200 import sharded as S
        uid = [S.unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(uid, src=0)
        comm = S.Communicator(ctx, uid[0], rank, world)

        def native_step(tables, timings=None):
            if tables[0].is_cuda:
                sc = S.interpolate_and_commit_sharded(None, n, C4_COLS, C4_BLOWUP, OFFSET, comm, device_pointers=[(t.data_ptr(), t.shape[1]) for t in tables])
            else:
                sc = S.interpolate_and_commit_sharded(tables, n, C4_COLS, C4_BLOWUP, OFFSET, comm)
            r = sc.root
            sc.free()
            return r
        n_ms_dev, n_wall_dev, n_root, n_prof, n_launches, n_clocks = timed(dev, args.steps, args.warmup, True, native_step)
        n_ms_e2e, _, n_root_e2e, _, _, _ = timed(host, e2e_steps, 2, False, native_step)
        comm.close()
        native_ok = n_root == n_root_e2e and (golden is None or n_root.hex() == golden["root"])
        flag = torch.tensor([1 if native_ok else 0], device="cuda", dtype=torch.int64)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()):
            torch_path = {"ms_per_step": ms_dev / args.steps, "e2e_ms_per_step": ms_e2e / e2e_steps, "wall_ms_per_step": wall_dev / args.steps,
                          "gpu_launches": launches, "root": root.hex(),
                          "how": "the same commit orchestrated from Python: torch.distributed batch_isend_irecv between C-ABI building blocks "
                                 "(distributed.py; the path the gloo tests cover)"}
            ms_dev, wall_dev, root, prof, launches, clocks, ms_e2e, root_e2e = n_ms_dev, n_wall_dev, n_root, n_prof, n_launches, n_clocks, n_ms_e2e, n_root_e2e
        else:
            native_err = "root mismatch on the C-ABI path: %s / %s" % (n_root.hex(), n_root_e2e.hex())
    except Exception as e:                                        # keep the line: the Python-orchestrated numbers stand
        native_err = "%s: %s" % (type(e).__name__, e)
    parity_ok = None if golden is None else (root.hex() == golden["root"] and root_e2e == root)
    c3_line = None
    if args.c3_log_n:
        try:
            c3_line = c3_sharded_bench(ctx, args, world, rank, barrier, dist)
        except Exception as e:
            c3_line = {"error": "%s: %s" % (type(e).__name__, e)}
    c5_line = None
    if args.c5_log_n:
        try:
            c5_line = c5_sharded_bench(ctx, args, world, rank, barrier, dist)
        except Exception as e:
            c5_line = {"error": "%s: %s" % (type(e).__name__, e)}
    cairo_line = None
    if not args.no_cairo:
        try:
            cairo_line = cairo_prove_sharded_bench(ctx, args, world, rank, barrier, dist)
        except Exception as e:                                # the commit line must survive a failure of the extra
            cairo_line = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        elems = cfg["elems_per_step"]
        imad_peak, lop_peak, int_peak_src = integer_peaks(ctx)
        total_kernel_ms = sum(v["ms"] for v in prof.values()) or 1.0
        kernels = {}
        for name, st in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            sec = st["ms"] * 1e-3
            kernels[name] = {"launches_per_step": st["launches"] / args.steps, "ms_per_step": st["ms"] / args.steps, "share": st["ms"] / total_kernel_ms,
                             "imad_frac": (st["muls"] * 80 / sec / 1e9) / imad_peak if sec else 0.0,
                             "alu_frac": (st["perms"] * 4320 / sec / 1e9) / lop_peak if sec else 0.0}
        tname = next(iter(kernels))
        tstat = prof[tname]
        line = {
            "metric": METRIC, "value": elems * args.steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32x8 (256-bit Montgomery) / u64 Keccak lanes", "data": "synthetic (splitmix64-seeded field elements, one stream per column)",
            "config": cfg,
            "parity_ok": parity_ok,
            "result": {"root": root.hex(), "golden_root": None if golden is None else golden["root"],
                       "golden_source": "tests/golden/synthetic/big_configs.json (CPU oracle, offline)"},
            "e2e": {"value": elems * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "h2d_bytes_per_step": int(n * C4_COLS * 32), "d2h_bytes_per_step": 32 * world,
                    "how": "every rank's column shard in pinned host memory, uploaded group by group on the copy stream under the previous "
                           "group's transforms; subtree roots read back"},
            "wall_ms_per_step": wall_dev / args.steps,
            "stages_ms": dict(zip(("lde", "exchange_exposed", "hash", "roots"), [round(float(x), 3) for x in stage_t.tolist()])),
            "stages_how": "one extra step with a host wait at every mark (max over ranks): `exchange_exposed` is what is left of the all-to-all "
                          "after the LDE of the later column groups has hidden the earlier groups' transfers",
            "pipeline_groups": groups, "gpu_launches": launches, "clocks": clocks, "host_cpus_bound": numa_cpus,
            "call": ("s252_interpolate_and_commit_sharded (C ABI; NCCL called from the library, one process per GPU)" if torch_path is not None
                     else "distributed.exchange_and_commit (torch.distributed between C-ABI building blocks)"),
            "roofline": {"bound": "int_issue", "kernel": tname, "achieved": tstat["muls"] * 80 / (tstat["ms"] * 1e-3) / 1e9 if tstat["ms"] else 0.0,
                         "peak": imad_peak, "unit": "G lane-op/s (IMAD.WIDE.U32)", "frac": kernels[tname]["imad_frac"], "traffic": None,
                         "peak_source": int_peak_src, "peak_note": PEAK_NOTE,
                         "scope": "rank 0's launches"},
            "kernels_rank0": kernels,
        }
        if torch_path is not None:
            line["torch_distributed_path"] = torch_path
        if native_err is not None:
            line["c_abi_path_error"] = native_err
        if c3_line is not None:
            line["c3_one_column"] = c3_line
        if c5_line is not None:
            line["c5_fri"] = c5_line
        if cairo_line is not None:
            line["cairo_prove"] = cairo_line
        print(json.dumps(line))
    for t in dev:
        del t
    dist.destroy_process_group()
    if parity_ok is False:
        raise SystemExit("bench.py: the sharded commit's root differs from the pinned oracle root")


def c3_sharded_bench(ctx, args, world, rank, barrier, dist):
    """BASELINE config C3 on N GPUs: interpolate_and_commit of ONE column (blowup 4) whose transform is shared by the GPUs
    (four-step NTT with one all-to-all, column_distributed.py); root checked against the pinned oracle root."""
    import torch

    import lambdaworks_cairo_prover_b200 as P
    from lambdaworks_cairo_prover_b200 import column_distributed as CD
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import random_felts
    log_n, blowup = args.c3_log_n, 4
    golden = golden_case("c3_%d" % log_n)
    n = 1 << log_n
    seed = golden["seed"] if golden else 0xB203 + log_n
    col = torch.from_numpy(random_felts(seed, n).view(np.int64)).pin_memory()
    be = CD.GpuColumnBackend(ctx)
    times, root = [], None
    for it in range(2 + 3):
        barrier()
        t0 = time.perf_counter()
        sc = CD.interpolate_and_commit_column_sharded(col, log_n, blowup, OFFSET, P.DefaultTranscript(), be)
        root = sc.root
        ctx.synchronize()
        barrier()
        if it >= 2:
            times.append((time.perf_counter() - t0) * 1e3)
        sc.free()
    ms = float(np.median(times))
    tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    stages = {}
    CD.interpolate_and_commit_column_sharded(col, log_n, blowup, OFFSET, P.DefaultTranscript(), be, timings=stages).free()
    CD._PLANS.clear()
    ctx.trim()
    torch.cuda.empty_cache()
    return {"workload": "C3: ONE column of 2^%d rows, blowup %d: interpolate_and_commit with the transform shared by %d GPU(s) (four-step, one "
                        "all-to-all per transform; slab upload inside the timed region)" % (log_n, blowup, world),
            "ms": ms, "ms_all": [round(x, 2) for x in times], "elems_per_s": n * blowup / (ms * 1e-3), "root": root.hex(),
            "stages_ms_rank0": {k: round(v, 3) for k, v in stages.items()},
            "parity_ok": None if golden is None else root.hex() == golden["root"]}


def c5_sharded_bench(ctx, args, world, rank, barrier, dist):
    """BASELINE config C5 on N GPUs: fri_commit_phase from the 2^(log_n+2) coset evaluations of a seeded polynomial (row blocks,
    pairwise fold exchange, collapse to one GPU at 2^19: fri_distributed.py) + grinding split over the GPUs; every root, the last
    value and the nonce are checked against the answers the CPU oracle pinned offline."""
    import torch

    import lambdaworks_cairo_prover_b200 as P
    from lambdaworks_cairo_prover_b200 import _native as N, felt, distributed as D, fri_distributed as F
    from lambdaworks_cairo_prover_b200.cairo_distributed import GpuCairoBackend, _dev_tensor
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import random_felts
    log_n = args.c5_log_n
    golden = golden_case("c5" if log_n == 22 else "c5_small" if log_n == 16 else "")
    blowup, grind = (golden["blowup"], golden["grinding_factor"]) if golden else (4, 20)
    seed = golden["seed"] if golden else 0xC500
    n = 1 << log_n
    m = n * blowup
    # layer 0 (not timed): the LDE of the polynomial on every rank, of which the rank keeps its block of rows
    commit, _ = P.lde_and_commit([P.Polynomial(random_felts(seed, n))], P.Domain(n, P.ProofOptions(blowup, 3, OFFSET, 1)), ctx)
    rows = m // world
    lde = _dev_tensor(N.lib().s252_commit_device_lde(commit.handle), m * 4, torch.device("cuda", torch.cuda.current_device())).view(m, 4)
    p0_block = lde[rank * rows:(rank + 1) * rows].clone()
    ctx.synchronize()
    commit.free()
    ctx.trim()
    be = GpuCairoBackend(ctx)
    times, result = [], None
    with D.backend_scope(be):
        for it in range(2 + 3):
            t = P.DefaultTranscript()
            t.append(bytes(32))
            blk = p0_block.clone()
            ctx.synchronize()
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            fri = F.fri_commit_phase_sharded(blk, m, log_n, t, OFFSET, be, None)
            nonce = F.generate_nonce_with_grinding_sharded(t.challenge(), grind, be, None)
            ctx.synchronize()
            barrier()
            if it >= 2:
                times.append((time.perf_counter() - t0) * 1e3)
            result = ([r.hex() for r in fri.roots], felt.to_bytes_be(fri.last_value).hex(), int(nonce), fri.tail_first)
            fri.free(be)
    ms = float(np.median(times))
    tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    del p0_block, lde
    ctx.trim()
    torch.cuda.empty_cache()
    ok = None
    if golden is not None:
        ok = result[0] == golden["roots"] and result[1] == golden["last_value"] and result[2] == golden["nonce"]
    return {"workload": "C5: fri_commit_phase from 2^%d evaluations (%d layers, blowup %d) + grinding %d, ONE phase on %d GPU(s): row-block "
                        "layer trees, pairwise half-layer exchange per fold, collapse to one GPU at 2^19" % (log_n + 2, log_n, blowup, grind, world),
            "ms": ms, "ms_all": [round(x, 2) for x in times], "sharded_layers": result[3], "last_root": result[0][-1], "nonce": result[2],
            "parity_ok": ok}


def cairo_prove_sharded_bench(ctx, args, world, rank, barrier, dist):
    """ONE Cairo proof on N GPUs: fibonacci_70000 under Provable80Bits (the options of benches/criterion_prover_70k.rs) and a 4x longer
    trace, through s252_cairo_prove_sharded -- one collective C-ABI call per rank, the whole orchestration and NCCL inside the library
    (csrc/cairo_sharded.cuh) -- with the torch.distributed-orchestrated prover (cairo_distributed.py) beside it; the proof must have
    the digest pinned by the CPU oracle."""
    import hashlib

    import lambdaworks_cairo_prover_b200 as P
    from lambdaworks_cairo_prover_b200 import _native as N, cairo, sharded as S
    from lambdaworks_cairo_prover_b200.cairo_distributed import generate_cairo_proof_sharded
    out = {"metric": "cairo_fib_prove_time", "unit": "ms", "higher_is_better": False, "n_gpus": world,
           "call": "s252_cairo_prove_sharded (C ABI: one collective call per rank, NCCL called from the library)",
           "how": "wall clock, barrier to barrier, max over ranks: the host trace table of every rank in, StarkProof::serialize bytes out on "
                  "rank 0; ONE proof sharded over the GPUs"}
    opts = P.ProofOptions.new_secure("Provable80Bits", 3)
    uid = [S.unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    comm = S.Communicator(ctx, uid[0], rank, world)
    for fib_n, key in ((args.fib_n, "fib"), (args.fib_n_large, "fib_large")):
        if not fib_n:
            continue
        regs, mem, size = cairo.run_program(cairo.fibonacci_program(fib_n))
        trace = cairo.build_main_trace(regs, mem, size)
        N.lib().s252_cairo_trace_pin(trace.handle)
        times, proof = [], None
        for it in range(2 + min(args.steps, 5)):
            barrier()
            t0 = time.perf_counter()
            proof = S.generate_cairo_proof_sharded(trace, opts, comm)
            barrier()
            if it >= 2:
                times.append((time.perf_counter() - t0) * 1e3)
        native_stages = json.loads(N.lib().s252_cairo_last_prove_stages().decode() or "{}")
        py_times, py_proof, stages = [], None, {}
        for it in range(2 + 3):
            barrier()
            t0 = time.perf_counter()
            py_proof = generate_cairo_proof_sharded(trace, opts, ctx)
            barrier()
            if it >= 2:
                py_times.append((time.perf_counter() - t0) * 1e3)
        generate_cairo_proof_sharded(trace, opts, ctx, timings=stages)
        golden = golden_case("fib%d_80bits" % fib_n)
        entry = {"program": "cairo0 fibonacci_%d" % fib_n, "trace_rows": trace.n_rows(), "value": float(np.median(times)),
                 "ms_all": [round(x, 2) for x in times],
                 "stages_ms_rank0_host_marks": {k: round(v, 3) for k, v in native_stages.items() if not isinstance(v, dict)},
                 "torch_distributed_path": {"value": float(np.median(py_times)), "ms_all": [round(x, 2) for x in py_times],
                                            "stages_ms_synchronised": {k: round(v, 3) for k, v in stages.items() if not isinstance(v, dict)},
                                            "commit_detail_ms": {k: round(v, 3) for k, v in stages.get("commit_detail", {}).items()}}}
        if rank == 0:
            entry["proof_bytes"] = len(proof)
            entry["proof_sha256"] = hashlib.sha256(proof).hexdigest()
            entry["same_bytes_both_paths"] = proof == py_proof
            entry["parity_ok"] = None if golden is None else (entry["proof_sha256"] == golden["sha256"] and proof == py_proof)
        out[key] = entry
        trace.free()
    comm.close()
    if "fib" in out:
        out["value"] = out["fib"]["value"]
    return out


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_commit_sample(log_n, threads):
    """The same step on the CPU restatement, at a reduced trace length (bounded sample)."""
    from oracle import pyoracle as O
    n = 1 << log_n
    m = n * BLOWUP
    t0 = time.perf_counter()
    tr = O.Transcript()
    for seed, cols in ((11, COLS_MAIN), (12, COLS_AUX)):
        trace = splitmix_felts(seed, n * cols).reshape(n, cols, 4)
        r = O.interpolate_and_commit(trace, BLOWUP, OFFSET, threads=threads, want_lde=False, want_nodes=False)
        tr.append(r["root"])
    comp = splitmix_felts(13, n * COLS_COMP).reshape(COLS_COMP, n, 4)
    lde = np.stack([O.evaluate_polynomial_on_lde_domain(comp[j], BLOWUP, n, O.fe_from_u64(OFFSET)) for j in range(COLS_COMP)])
    _, root = O.commit_columns(lde)
    tr.append(root)
    p0 = splitmix_felts(14, n)
    O.fri_commit_phase(log_n, p0, tr, O.fe_from_u64(OFFSET), m, keep=False)
    O.generate_nonce_with_grinding(tr.challenge(), GRIND)
    dt = time.perf_counter() - t0
    elems = workload_config(log_n)["elems_per_step"]
    return elems / dt, dt


def cpu_baseline_sample(log_n):
    restore_launch_affinity()
    cores = os.cpu_count() or 1
    value, dt = cpu_commit_sample(log_n, cores)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "trace_rows": 1 << log_n,
            "sample": "the same step at N=2^%d rows instead of 2^%d (all 54 columns, %d FRI layers, grinding %d): %.1f s of CPU work; "
                      "C restatement of the reference (the Rust reference cannot be built here: no cargo), LDE threaded over columns "
                      "like the reference's `parallel` feature (prover.rs:169-183), Merkle trees and FRI sequential as in the reference, "
                      "so most of the run uses ONE core" % (log_n, LOG_N, log_n, GRIND, dt)}


def run_reference(args):
    """The reference arm: the CPU restatement at the SAME configuration as the GPU arm (N = 2^19 by default).  One such step is
    ~45 s of CPU work, so the arm runs as many full-size steps as fit --cpu-budget-s (at least one) and says how many."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_gpus = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1")))
    if n_gpus > 1 and args.mode != "traces":
        return run_reference_c4(args, cores, n_gpus)
    log_n = args.cpu_log_n if args.cpu_log_n else args.log_n
    cpu_commit_sample(8, cores)                      # loads the checker library, touches every code path once
    t0 = time.perf_counter()
    total_elems, steps_done, last = 0, 0, 0.0
    while steps_done < args.steps and (steps_done == 0 or time.perf_counter() - t0 + last < args.cpu_budget_s):
        _, last = cpu_commit_sample(log_n, cores)
        total_elems += workload_config(log_n)["elems_per_step"]
        steps_done += 1
    wall = time.perf_counter() - t0
    value = total_elems / wall
    cfg = workload_config(log_n)
    cfg["parallelism"] = "host cores only"
    sample = ("%d step(s) of the C2 commit phase at N=2^%d rows (54 columns, blowup %d, %d FRI layers, grinding %d), %.0f s each; oracle/ C "
              "restatement (the Rust reference cannot be built here: no cargo, un-vendored git dependencies), LDE threaded over columns as "
              "under `parallel`, Merkle trees and FRI sequential as in the reference" % (steps_done, log_n, BLOWUP, log_n, GRIND, wall / steps_done))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps_done, "steps_requested": args.steps, "warmup": 0, "ms_per_step": wall * 1e3 / steps_done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 (256-bit Montgomery)", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_cairo:
        line["cairo_prove"] = cpu_cairo_prove_sample(args.cpu_fib_n)
    print(json.dumps(line))


def run_reference_c4(args, cores, n_gpus):
    """The reference arm beside the sharded arm (N > 1): the same C4-shaped commit (33 columns, blowup 8) on the host cores.  The full
    2^22-row table is ~9 minutes of CPU work per step, so each step is a 2^17-row sample of it (elems/s is size-normalised; the workload
    string says so)."""
    from oracle import pyoracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import random_felts
    log_n = args.cpu_log_n or 17
    n = 1 << log_n
    trace = np.empty((n, C4_COLS, 4), dtype=np.uint64)
    for j in range(C4_COLS):
        trace[:, j, :] = random_felts(C4_SEED + j, n)
    O.interpolate_and_commit(trace[:256], C4_BLOWUP, OFFSET, threads=cores, want_lde=False, want_nodes=False)
    t0 = time.perf_counter()
    steps_done, last = 0, 0.0
    while steps_done < args.steps and (steps_done == 0 or time.perf_counter() - t0 + last < args.cpu_budget_s):
        a = time.perf_counter()
        O.interpolate_and_commit(trace, C4_BLOWUP, OFFSET, threads=cores, want_lde=False, want_nodes=False)
        last = time.perf_counter() - a
        steps_done += 1
    wall = time.perf_counter() - t0
    cfg = c4_config(log_n, 0)
    cfg["workload"] = ("bounded CPU sample of C4: ONE interpolate_and_commit of a 2^%d-row x %d-column trace, blowup %d (the GPU arm commits 2^%d "
                       "rows; a full-size CPU step is ~9 minutes)" % (log_n, C4_COLS, C4_BLOWUP, C4_LOG_N))
    cfg["parallelism"] = "host cores only"
    value = cfg["elems_per_step"] * steps_done / wall
    sample = ("%d step(s), %.0f s each; oracle/ C restatement (no cargo here), LDE threaded over columns as under the reference's `parallel` "
              "feature, leaf hashing and tree sequential as in the reference" % (steps_done, wall / steps_done))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps_done, "steps_requested": args.steps,
        "warmup": 0, "ms_per_step": wall * 1e3 / steps_done, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64x4 (256-bit Montgomery)", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    # stdout carries exactly one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL_DEBUG=INFO lines (comm sizes, transports) belong on stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=LOG_N, help="trace length exponent (default: the C2 size)")
    ap.add_argument("--cpu-log-n", type=int, default=0,
                    help="trace length exponent of the CPU runs (default: --impl reference runs the full size, the cpu_baseline "
                         "object of the GPU line a 2^17-row sample)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="--impl reference: stop starting new full-size steps after this long")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cairo", action="store_true", help="skip the Cairo fib prove-time measurement")
    ap.add_argument("--no-c4", action="store_true", help="N=1: skip the one-GPU point of the sharded C4 commit")
    ap.add_argument("--fib-n", type=int, default=70000, help="fibonacci program of the prove-time measurement")
    ap.add_argument("--cpu-fib-n", type=int, default=4000, help="fibonacci program of the bounded CPU prove sample")
    ap.add_argument("--mode", default=None, choices=["traces", "sharded"],
                    help="'sharded' (default for N>1) = ONE C4 trace, columns sharded over the GPUs with an all-to-all before leaf "
                         "hashing, and ONE Cairo proof on all GPUs (strong scaling); 'traces' (default for N=1) = the C2 commit phase, "
                         "one independent trace per GPU (weak scaling)")
    ap.add_argument("--c4-log-n", type=int, default=C4_LOG_N, help="trace length exponent of the sharded C4 commit")
    ap.add_argument("--pipeline-groups", type=int, default=4, help="sharded mode: column groups per rank of the LDE -> exchange pipeline")
    ap.add_argument("--c5-log-n", type=int, default=22, help="sharded mode: FRI commit phase from 2^(this+2) evaluations (0 = skip; 22 and 16 are pinned)")
    ap.add_argument("--c3-log-n", type=int, default=24, help="sharded mode: rows (log2) of the one-column C3 commit (0 = skip; 24 and 26 are pinned)")
    ap.add_argument("--fib-n-large", type=int, default=280000, help="sharded mode: the longer fibonacci program (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "sharded" or (args.mode is None and int(os.environ.get("WORLD_SIZE", "1")) > 1):
        run_gpu_sharded(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
