#!/usr/bin/env python
"""bench.py — the hot path on synthetic input of BASELINE.json's shapes; prints ONE JSON line (rank 0).

  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...     # the CPU restatement of the reference on the host cores

BASELINE.json's metric has two halves and the line carries both:
  * top level  — scanned bp/s on configs[3], the scan sweep: 500 PWMs (len 8-40) over 10M x 200 bp (2 Gbp), forward +
    reverse strands, fused threshold, per-motif occurrence counts.  A "step" is one pass of the scan over all sequences.
    At N>1 the sequences are sharded over ranks (strong scaling: total work fixed), no data-path collective; the per-motif
    counts are summed once per step with one NCCL all_reduce.  Inputs exceed L2 (0.5 GB packed), no flush needed.
  * "training" — training sequences/s on configs[1] (20k x 100 bp, batch 6 per rank, AdaBelief): optimiser steps of the
    unrolled CSC network (forward + hand-derived reverse pass + update), one all_reduce of the 30 433 gradients per step
    at N>1 (weak scaling: every rank adds a batch of 6).
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="both", choices=["both", "scan", "train"])
    ap.add_argument("--nseq", type=int, default=10_000_000)
    ap.add_argument("--seqlen", type=int, default=200)
    ap.add_argument("--motifs", type=int, default=500)
    ap.add_argument("--train-nseq", type=int, default=20_000)
    ap.add_argument("--train-seqlen", type=int, default=100)
    ap.add_argument("--train-steps", type=int, default=1500, help="optimiser steps per timed bench step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of each cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------------------------
# scan half
# ------------------------------------------------------------------------------------------------------------------
def make_motifs(K, seed=4):
    from motifs_jl_b200 import synth
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(K, 8, 40, seed))
    return ms, synth.stated_thresholds(ms, 0.7)


def cells_per_seq(lens, Lb):
    lens = np.asarray(lens, np.int64)
    return int(2 * ((Lb - lens + 1).clip(min=0) * lens).sum())


def scan_config(args, **extra):
    c = {"workload": f"scan sweep: {args.motifs} PWMs (len 8-40) x {args.nseq} seqs x {args.seqlen} bp, fwd+rc, fused threshold, counts",
         "baseline_config_index": 3, "n_seqs": args.nseq, "seq_len": args.seqlen, "n_motifs": args.motifs,
         "thresholds": "0.7 x max score (Float16)", "l2": "inputs larger than L2 (no flush needed)", "parallelism": f"seq-shard x{args.gpus}"}
    c.update(extra)
    return c


def train_config(args, world, **extra):
    c = {"workload": f"CSC training: {args.train_nseq} seqs x {args.train_seqlen} bp (planted gapped motif), batch 6 per rank, "
                     f"M=50 K=24 h=12 q=32, 6 XYZ + 3 DF passes, AdaBelief",
         "baseline_config_index": 1, "n_seqs": args.train_nseq, "seq_len": args.train_seqlen, "global_batch": 6 * world,
         "optimizer_steps_per_bench_step": args.train_steps, "parallelism": f"dp{world}", "l2": "working set (<20 MB) is L2 resident by design"}
    c.update(extra)
    return c


def cpu_scan_rate(pw, lens, thr, ascii_rows, target_s):
    """oracle (CPU port) scan+filter+counts on a bounded sample, all host threads; returns (bp/s, sample, cores, seconds)."""
    from oracle import scan_oracle as so
    cores = os.cpu_count() or 1
    codes = so.ascii_to_codes(ascii_rows)
    probe = min(len(codes), 16 * cores)
    t0 = time.perf_counter()
    so.scan(pw, lens, codes[:probe], thr, want_hits=False)
    dt = time.perf_counter() - t0
    n = int(min(len(codes), max(probe, probe * target_s / max(dt, 1e-6))))
    t0 = time.perf_counter()
    so.scan(pw, lens, codes[:n], thr, want_hits=False)
    dt = time.perf_counter() - t0
    return n * codes.shape[1] / dt, n, cores, dt


def cpu_train_rate(ascii_rows, flat, target_s):
    """oracle (PyTorch-CPU position-space restatement + autograd) loss/gradient/AdaBelief steps on the host cores."""
    import torch
    from oracle import csc_oracle as co, scan_oracle as so
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = co.Hyperparam()
    codes = so.ascii_to_codes(ascii_rows)
    n = co.n_params(hp)
    p = flat.astype(np.float32).copy()
    mt, st = np.zeros(n), np.zeros(n)
    rng = np.random.default_rng(0)
    steps, t0 = 0, time.perf_counter()
    while True:
        idx = rng.permutation(len(codes))[:hp.batch_size]
        _, g, _ = co.loss_and_grad(codes[idx], p, hp)
        pn, mt, st = co.adabelief_step(p[:n].astype(np.float64), g[:n].astype(np.float64), mt, st, t=steps + 1)
        p[:n] = pn.astype(np.float32)
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= target_s or steps >= 400:
            break
    return hp.batch_size * steps / dt, steps, cores, dt


def bench_scan(args, ctx, torch, dist, world, rank, local, dev, stream):
    import motifs_jl_b200 as mb  # noqa: F401
    from oracle import scan_oracle as so                                     # cpu_baseline leg + pwm layout helper only
    K, Lb = args.motifs, args.seqlen
    n_lo, n_hi = args.nseq * rank // world, args.nseq * (rank + 1) // world
    n_local = n_hi - n_lo
    ms, thr = make_motifs(K)
    pw, lens = so.pack_pwms(ms.pwms)
    g = torch.Generator(device=dev)
    g.manual_seed(4 + 1000 * rank)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_dev = torch.empty((n_local, Lb), dtype=torch.uint8, device=dev)
    chunk = 1 << 20
    for s in range(0, n_local, chunk):
        e = min(n_local, s + chunk)
        ascii_dev[s:e] = lut[torch.randint(0, 4, (e - s, Lb), device=dev, generator=g, dtype=torch.int64)]
    ascii_host = torch.empty((n_local, Lb), dtype=torch.uint8, pin_memory=True)
    ascii_host.copy_(ascii_dev)
    torch.cuda.synchronize()
    seqs = ctx.seqs_from_device_ptr(ascii_dev.data_ptr(), n_local, Lb)
    del ascii_dev
    counts_dev = torch.zeros((K, 4), dtype=torch.int64, device=dev)

    def reduce_counts(c):
        if world > 1:
            counts_dev.copy_(torch.from_numpy(c))
            dist.all_reduce(counts_dev)
            return counts_dev.cpu().numpy()
        return c

    def step_resident():
        _, c = ctx.scan(seqs, pw, lens, thr, want_hits=False, want_counts=True)
        return reduce_counts(c)

    def step_e2e():
        s2 = ctx.seqs_from_host_ptr(ascii_host.data_ptr(), n_local, Lb, wait=False)     # H2D of this step's input + pack, overlapping the scan
        _, c = ctx.scan(s2, pw, lens, thr, want_hits=False, want_counts=True)
        s2.free()
        return reduce_counts(c)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_scan = t_cnt = 0.0
        n_launch = nk = 0
        timed.t_verify = 0.0
        e0.record(stream)
        for _ in range(steps):
            res = fn()
            t, l = ctx.last_timing()
            t_scan += t["scan"]; t_cnt += t["count"]; nk += l["scan"]; timed.t_verify += t["emit"]
            n_launch += sum(l.values())
        e1.record(stream)
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / steps, res, t_scan, t_cnt, nk, n_launch

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, counts, t_scan, t_cnt, n_scan_launch, n_launch = timed(step_resident, args.steps, args.warmup)
    t_verify = timed.t_verify
    from motifs_jl_b200 import _lib as _mblib
    scan_path = _mblib.scan_last_path(ctx)                                    # 1: tcgen05 pre-filter + exact re-scoring, 0: SIMT scan_kernel
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, counts2, *_ = timed(step_e2e, max(1, args.steps), 1)
    assert np.array_equal(counts, counts2)
    total_bp = args.nseq * Lb
    out = None
    if rank == 0:
        peaks = load_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel: scan_kernel.  Algorithmic bytes per launch = packed sequence bytes of the batch (0.25 B/bp) + the PWM
        # tables once; cells = table look-up-adds (the binding resource, 2 B of shared memory each).
        launches_per_step = n_scan_launch / args.steps
        bytes_per_launch = (n_local * Lb / 4.0) / launches_per_step + float(2 * 4 * lens.sum() * 2)
        ms_per_launch = t_scan / max(1, n_scan_launch)
        ach = bytes_per_launch / (ms_per_launch * 1e-3) / 1e9
        cells_rate = cells_per_seq(lens, Lb) * n_local / ((t_scan / args.steps) * 1e-3)
        sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        alu_peak = 148 * 4 * (512.0 / 28.0) * sm_mhz * 1e6             # cells/s when the ALU pipe is saturated
        out = {"metric": "scanned_bp_per_sec", "value": total_bp / (ms_step / 1e3), "unit": "bp/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f16", "data": "synthetic", "config": scan_config(args),
               "e2e": {"value": total_bp / (ms_e2e / 1e3), "unit": "bp/s", "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": int(n_local * Lb + pw.nbytes + lens.nbytes + 2 * K), "d2h_bytes_per_step": int(K * 4 * 8)},
               "gpu_launches": int(n_launch), "clocks": clocks,
               "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                            # ncu --set full, profiles/r01_scan_v2_ncu_full_summary.csv: dram read+write = 30.76 KB per sequence of a launch
                            # (hit-mask writes: 1 bit per motif, strand, position); algorithmic bytes are 50 B per sequence
                            "traffic": 30760.0 * (n_local / launches_per_step) if (Lb == 200 and K == 500) else None,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "kernel": "scan_kernel",
                            "ms_per_launch": ms_per_launch, "kernel_share_of_step": (t_scan / args.steps) / ms_step,
                            "count_kernel_share_of_step": (t_cnt / args.steps) / ms_step,
                            "note": "2-bit packing makes the scan table-lookup bound, not HBM bound (SURVEY §8d): see binding",
                            "binding": {"bound": "alu-issue", "what": "per PWM cell: 1/2 PRMT (entry select) + 1/2 HADD2/HFMA2 (one sequential Float16 add); the 8 PRMT + ~4 HADD2 per column "
                                                "share the ALU pipe (1 instr / 2 clk / SM sub-partition): ~28 clk per 512 cells",
                                        "cells_per_s": cells_rate, "achieved": cells_rate / 1e12, "peak": alu_peak / 1e12, "unit": "Tcell/s",
                                        "frac": cells_rate / alu_peak, "peak_source": "148 SM x 4 sub-partitions x 512 cells / 28 clk x SM clock under load"}},
               "checks": {"counts_sum": [int(x) for x in counts.sum(axis=0)]}}
        if scan_path == 1:
            # dominant kernel: k_scan_tc (csrc/scan_tc.cuh), a one-hot GEMM on the tensor cores.  Algorithmic flops: 8 per PWM
            # cell = 2 (multiply-add) x 4 (one-hot bases), cells = strands x N x sum_k (Lb - len_k + 1) len_k; the launch also computes the
            # zero padding of K to the longest motif of a 256-slot block and the windows that straddle two sequences, which are not counted.
            tf_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1413.0)))
            flops_per_launch = 8.0 * cells_per_seq(lens, Lb) * n_local / launches_per_step
            tf = flops_per_launch / (ms_per_launch * 1e-3) / 1e12
            hbm = dict(out["roofline"]); hbm.pop("binding", None); hbm.pop("traffic", None); hbm["kernel"] = "k_scan_tc"
            hbm["note"] = "algorithmic HBM bytes (0.25 B/bp + tables) over the same kernel time: the scan is not HBM bound (SURVEY §8d)"
            out["roofline"] = {"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak,
                               # ncu --set full of one launch (profiles/r01_scan_tc_ncu_full_summary.csv): dram read + write = 15.96 MB for
                               # 291 271 sequences = 54.8 B per sequence (algorithmic: 50 B packed sequence + the tables once)
                               "traffic": 54.8 * (n_local / launches_per_step) if (Lb == 200 and K == 500) else None,
                               "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback",
                               "kernel": "k_scan_tc", "ms_per_launch": ms_per_launch,
                               "kernel_share_of_step": (t_scan / args.steps) / ms_step,
                               "verify_kernel_share_of_step": (t_verify / args.steps) / ms_step,
                               "count_kernel_share_of_step": (t_cnt / args.steps) / ms_step,
                               "algorithmic_flops_per_launch": flops_per_launch,
                               "note": "FP16 one-hot GEMM pre-filter (FP32 accumulate in TMEM), candidates re-scored with the reference's sequential Float16 adds: "
                                       "hit sets identical to the SIMT kernel; cells/s below counts useful PWM cells only",
                               "cells_per_s": cells_rate, "hbm": hbm}
        if world == 1 and not args.no_cpu_baseline:
            sample = ascii_host[: min(n_local, 20000)].numpy()
            rate, n, cores, dt = cpu_scan_rate(pw, lens, thr, sample, args.cpu_seconds)
            out["cpu_baseline"] = {"value": rate, "unit": "bp/s", "cores": cores, "kind": "port", "seconds": dt,
                                   "sample": f"first {n} of {args.nseq} sequences x {Lb} bp, all {K} PWMs, both strands (oracle/scan_oracle.c, OpenMP)"}
            _, oc = so.scan(pw, lens, so.ascii_to_codes(sample[:n]), thr, want_hits=False)
            s3 = ctx.seqs_from_ascii(sample[:n])
            _, gc = ctx.scan(s3, pw, lens, thr, want_hits=False)
            out["checks"]["sample_counts_match_oracle"] = bool(np.array_equal(oc, gc))
            s3.free()
    seqs.free()
    del ascii_host
    return out


# ------------------------------------------------------------------------------------------------------------------
# training half
# ------------------------------------------------------------------------------------------------------------------
def bench_train(args, ctx, torch, dist, world, rank, local, dev):
    from motifs_jl_b200 import model as mdl, parallel, synth
    from motifs_jl_b200._lib import CscModel
    hp = mdl.Hyperparam()
    N, Lb = args.train_nseq, args.train_seqlen
    a = synth.planted_gapped(N, Lb, 2)
    n_train = N - int(np.floor((1 - 0.9) * N))                            # loadfasta/helpers.jl:144
    a = a[:n_train]
    pinned = torch.from_numpy(a).pin_memory()
    cdl = mdl.ucdl(hp, np.random.default_rng(2))
    side = torch.cuda.Stream(device=dev)                                  # library stream; NCCL orders itself against it
    ctx.set_stream(side.cuda_stream)
    out = None
    with torch.cuda.stream(side):
        seqs = ctx.seqs_from_host_ptr(pinned.data_ptr(), n_train, Lb)
        model = CscModel(ctx, hp, Lb, n_groups=1)
        model.set_params(cdl.flat)
        grad_view = None
        if world > 1:
            _, gptr = model.device_ptrs()
            grad_view = torch.as_tensor(mdl._DevArray(gptr, model.n_total), device=dev)
        rng = np.random.default_rng(1234)                                 # same permutation stream on every rank
        per_step = hp.batch_size * world
        state = {"perm": rng.permutation(n_train), "pos": 0, "loss": None, "l1": None, "launches": 0}

        def opt_step():
            if state["pos"] + per_step > n_train:
                state["perm"], state["pos"] = rng.permutation(n_train), 0
            lo = state["pos"] + rank * hp.batch_size
            idx = state["perm"][lo: lo + hp.batch_size]
            state["pos"] += per_step
            model.step_begin(seqs, idx)                                   # H2D: 6 sequence indices; graph replay of fwd + reverse pass
            if world > 1:
                parallel.all_reduce_mean_(grad_view)                      # one NCCL all-reduce of the gradient vector
            state["loss"], state["l1"] = model.adabelief_step()           # update + D2H of loss and l1(F)

        def bench_step():
            for _ in range(args.train_steps):
                opt_step()

        for _ in range(min(args.warmup, 3) * 50):                         # >= 3 warm-up iterations of the step (graph capture included)
            opt_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.last_timing()[1]["csc"]
        e0.record(side)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            bench_step()
        e1.record(side)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_bench_step = float(tt.item()) / args.steps
        n_opt = args.steps * args.train_steps

        # end-to-end leg: every optimiser step ships ITS batch from host memory (the reference's `S |> gpu`, train.jl:41):
        # 6 ASCII rows -> 2 bit/base -> one H2D copy -> step -> D2H of loss and l1
        def opt_step_host():
            if state["pos"] + per_step > n_train:
                state["perm"], state["pos"] = rng.permutation(n_train), 0
            lo = state["pos"] + rank * hp.batch_size
            idx = state["perm"][lo: lo + hp.batch_size]
            state["pos"] += per_step
            model.step_begin_host(a[idx])
            if world > 1:
                parallel.all_reduce_mean_(grad_view)
            state["loss"], state["l1"] = model.adabelief_step()

        for _ in range(50):
            opt_step_host()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(side)
        for _ in range(args.train_steps):
            opt_step_host()
        f1.record(side)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_e2e_opt = float(t2.item()) / args.train_steps
        if rank == 0:
            seq_s = hp.batch_size * world * args.train_steps / (ms_bench_step / 1e3)
            out = {"metric": "training_sequences_per_sec", "value": seq_s, "unit": "seq/s", "n_gpus": world, "steps": args.steps,
                   "ms_per_step": ms_bench_step, "ms_per_optimizer_step": ms_bench_step / args.train_steps, "higher_is_better": True,
                   "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": train_config(args, world),
                   # value: sequences resident in HBM, a step ships 6 indices.  e2e: every step ships its 6 sequences from the host.
                   "e2e": {"value": hp.batch_size * world / (ms_e2e_opt / 1e3), "unit": "seq/s", "ms_per_optimizer_step": ms_e2e_opt,
                           "h2d_bytes_per_step": int(args.train_steps * hp.batch_size * ((Lb + 15) // 16) * 4),
                           "d2h_bytes_per_step": int(args.train_steps * 8),
                           "note": "per optimiser step: 6 ASCII rows packed to 2 bit/base on the host, one H2D copy, loss + l1 read back"},
                   "gpu_launches": int(ctx.last_timing()[1]["csc"] - l0) if ctx.last_timing()[1]["csc"] >= l0 else None,
                   "kernels_per_optimizer_step": None, "final_loss": state["loss"], "final_l1_F": state["l1"],
                   "wall_s": wall, "optimizer_steps_timed": n_opt}
            if world == 1 and not args.no_cpu_baseline:
                rate, steps, cores, dt = cpu_train_rate(a[:2000], cdl.flat, args.cpu_seconds)
                out["cpu_baseline"] = {"value": rate, "unit": "seq/s", "cores": cores, "kind": "port", "seconds": dt,
                                       "sample": f"{steps} optimiser steps of batch 6 (oracle/csc_oracle.py, PyTorch-CPU fp32 + autograd, {cores} threads)"}
        model.free()
        # extra data point (NOT the headline): 16 independent batches of 6 per optimiser step on one GPU — the same arithmetic as
        # 16 data-parallel ranks (global batch 96, gradients averaged), i.e. a different optimisation trajectory than the reference's
        if rank == 0 and world == 1:
            G = 16
            mg = CscModel(ctx, hp, Lb, n_groups=G)
            mg.set_params(cdl.flat)
            r2 = np.random.default_rng(7)
            for _ in range(20):
                mg.step_begin(seqs, r2.permutation(n_train)[:6 * G]); mg.adabelief_step()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nst = 200
            g0.record(side)
            for _ in range(nst):
                mg.step_begin(seqs, r2.permutation(n_train)[:6 * G]); mg.adabelief_step()
            g1.record(side)
            torch.cuda.synchronize()
            msg = g0.elapsed_time(g1) / nst
            out["batched_groups"] = {"groups_per_step": G, "global_batch": 6 * G, "value": 6 * G / (msg / 1e3), "unit": "seq/s",
                                     "ms_per_optimizer_step": msg,
                                     "note": "each group is a reference batch (own median mask, own D/F updates); gradients averaged over groups"}
            mg.free()
        seqs.free()
    ctx.set_stream(None)
    return out


def run_reference(args):
    """--impl reference: the reference cannot run here (Julia absent, CuArray-typed code), so this arm times the oracle's CPU
    restatement of the same path on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from motifs_jl_b200 import model as mdl, synth
    from oracle import scan_oracle as so
    ms, thr = make_motifs(args.motifs)
    pw, lens = so.pack_pwms(ms.pwms)
    per_step_target = max(2.0, min(args.cpu_seconds, 60.0 / max(1, args.steps + args.warmup)))
    sample = synth.random_ascii(min(args.nseq, 20000), args.seqlen, 4)
    _, n, cores, _ = cpu_scan_rate(pw, lens, thr, sample, per_step_target)
    codes = so.ascii_to_codes(sample[:n])
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        so.scan(pw, lens, codes, thr, want_hits=False)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms_step = 1e3 * float(np.mean(times))
    value = n * args.seqlen / (ms_step / 1e3)
    out = {"impl": "reference", "metric": "scanned_bp_per_sec", "value": value, "unit": "bp/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": scan_config(args, sample_seqs=n),
           "cpu_baseline": {"value": value, "unit": "bp/s", "cores": cores, "kind": "port",
                            "sample": f"{n} of {args.nseq} sequences x {args.seqlen} bp, all {args.motifs} PWMs, both strands, per step"},
           "e2e": {"value": value, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    if args.workload in ("both", "train"):
        a = synth.planted_gapped(2000, args.train_seqlen, 2)
        cdl = mdl.ucdl(mdl.Hyperparam(), np.random.default_rng(2))
        rate, steps, cores, dt = cpu_train_rate(a, cdl.flat, min(args.cpu_seconds, 20.0))
        out["training"] = {"impl": "reference", "metric": "training_sequences_per_sec", "value": rate, "unit": "seq/s",
                           "config": train_config(args, 1), "higher_is_better": True,
                           "cpu_baseline": {"value": rate, "unit": "seq/s", "cores": cores, "kind": "port",
                                            "sample": f"{steps} optimiser steps of batch 6 in {dt:.1f} s (oracle/csc_oracle.py)"},
                           "e2e": {"value": rate, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import motifs_jl_b200 as mb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = mb.Context(local)
    stream = torch.cuda.current_stream()
    out = None
    if args.workload in ("both", "scan"):
        out = bench_scan(args, ctx, torch, dist, world, rank, local, dev, stream)
    if args.workload in ("both", "train"):
        tr = bench_train(args, ctx, torch, dist, world, rank, local, dev)
        if rank == 0:
            if out is None:
                out = tr
                out["warmup"] = args.warmup
            else:
                out["training"] = tr
                out["gpu_launches"] = int(out["gpu_launches"]) + int(tr.get("gpu_launches") or 0)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
