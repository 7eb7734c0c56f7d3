#!/usr/bin/env python
"""bench.py — the hot path on synthetic input of BASELINE.json's shapes; prints ONE JSON line (rank 0).

  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...     # the CPU restatement of the reference on the host cores

Workload (config.workload): BASELINE.json configs[3], the scan sweep — 500 PWMs (len 8-40) over 10M x 200 bp
(2 Gbp), forward + reverse strands, fused threshold, per-motif occurrence counts.  A "step" is one pass of the
scan over all sequences.  At N>1 the 10M sequences are sharded over ranks (strong scaling: total work fixed),
no data-path collective; the per-motif counts are summed once per step with one NCCL all_reduce.
Inputs are larger than L2 (0.5 GB packed, 4 KB of masks per 32 positions), so no explicit L2 flush is needed.
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
    ap.add_argument("--nseq", type=int, default=10_000_000)
    ap.add_argument("--seqlen", type=int, default=200)
    ap.add_argument("--motifs", type=int, default=500)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
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


def make_motifs(K, seed=4):
    from motifs_jl_b200 import synth
    ms = synth.motifs_from_count_matrices(synth.random_count_matrices(K, 8, 40, seed))
    return ms, synth.stated_thresholds(ms, 0.7)


def cells_per_seq(lens, Lb):
    lens = np.asarray(lens, np.int64)
    return int(2 * ((Lb - lens + 1).clip(min=0) * lens).sum())


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


def run_reference(args):
    """--impl reference: the reference cannot run here (Julia absent, CuArray-typed), so this arm times the
    oracle's CPU restatement of the same path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from motifs_jl_b200 import synth
    from oracle import scan_oracle as so
    ms, thr = make_motifs(args.motifs)
    pw, lens = so.pack_pwms(ms.pwms)
    per_step_target = max(2.0, min(args.cpu_seconds, 60.0 / max(1, args.steps + args.warmup)))
    sample = synth.random_ascii(min(args.nseq, 20000), args.seqlen, 4)
    rate0, n, cores, dt = cpu_scan_rate(pw, lens, thr, sample, per_step_target)
    times = []
    codes = so.ascii_to_codes(sample[:n])
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        so.scan(pw, lens, codes, thr, want_hits=False)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms_step = 1e3 * float(np.mean(times))
    value = n * args.seqlen / (ms_step / 1e3)
    out = {"impl": "reference", "metric": "scanned_bp_per_sec", "value": value, "unit": "bp/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f16", "data": "synthetic",
           "config": workload_config(args, sample_seqs=n),
           "cpu_baseline": {"value": value, "unit": "bp/s", "cores": cores, "kind": "port",
                            "sample": f"{n} of {args.nseq} sequences x {args.seqlen} bp, all {args.motifs} PWMs, both strands, per step"},
           "e2e": {"value": value, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args, **extra):
    c = {"workload": f"scan sweep: {args.motifs} PWMs (len 8-40) x {args.nseq} seqs x {args.seqlen} bp, fwd+rc, fused threshold, counts",
         "baseline_config_index": 3, "n_seqs": args.nseq, "seq_len": args.seqlen, "n_motifs": args.motifs,
         "thresholds": "0.7 x max score (Float16)", "l2": "inputs larger than L2 (no flush needed)", "parallelism": f"seq-shard x{args.gpus}"}
    c.update(extra)
    return c


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import motifs_jl_b200 as mb
    from oracle import scan_oracle as so   # only for the cpu_baseline leg (rank 0, N=1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = mb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    K, Lb = args.motifs, args.seqlen
    n_lo = args.nseq * rank // world
    n_hi = args.nseq * (rank + 1) // world
    n_local = n_hi - n_lo
    ms, thr = make_motifs(K)
    pw, lens = so.pack_pwms(ms.pwms)        # layout helper only (numpy)

    # synthetic sequences: iid uniform bases generated on the device, then an ASCII copy in pinned host memory
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

    def step_resident():
        _, c = ctx.scan(seqs, pw, lens, thr, want_hits=False, want_counts=True)
        if world > 1:
            counts_dev.copy_(torch.from_numpy(c))
            dist.all_reduce(counts_dev)
            return counts_dev.cpu().numpy()
        return c

    def step_e2e():
        s2 = ctx.seqs_from_host_ptr(ascii_host.data_ptr(), n_local, Lb)     # H2D of this step's input + pack
        _, c = ctx.scan(s2, pw, lens, thr, want_hits=False, want_counts=True)
        s2.free()
        if world > 1:
            counts_dev.copy_(torch.from_numpy(c))
            dist.all_reduce(counts_dev)
            return counts_dev.cpu().numpy()
        return c

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
        n_launch = 0
        nk = 0
        e0.record(stream)
        for _ in range(steps):
            res = fn()
            t, l = ctx.last_timing()
            t_scan += t["scan"]; t_cnt += t["count"]; nk += l["scan"]
            n_launch += sum(l.values())
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / steps, res, t_scan, t_cnt, nk, n_launch

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, counts, t_scan, t_cnt, n_scan_launch, n_launch = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, counts2, *_ = timed(step_e2e, max(1, args.steps), 1)
    assert np.array_equal(counts, counts2)

    total_bp = args.nseq * Lb
    value = total_bp / (ms_step / 1e3)
    e2e = total_bp / (ms_e2e / 1e3)

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel: scan_kernel.  Algorithmic bytes per launch = packed sequence bytes of the batch (0.25 B/bp)
        # + the PWM tables once; cells = table look-up-adds (the binding resource, 2 B of smem each).
        launches_per_step = n_scan_launch / args.steps
        bytes_per_launch = (n_local * Lb / 4.0) / launches_per_step + float(2 * 4 * lens.sum() * 2)
        ms_per_launch = t_scan / max(1, n_scan_launch)
        ach = bytes_per_launch / (ms_per_launch * 1e-3) / 1e9
        cells_step = cells_per_seq(lens, Lb) * n_local
        cells_rate = cells_step / ((t_scan / args.steps) * 1e-3)
        sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        smem_peak = 148 * 128 * sm_mhz * 1e6                                 # B/s of shared-memory read bandwidth
        out = {"metric": "scanned_bp_per_sec", "value": value, "unit": "bp/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f16", "data": "synthetic", "config": workload_config(args),
               "e2e": {"value": e2e, "unit": "bp/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(n_local * Lb + pw.nbytes + lens.nbytes + 2 * K),
                       "d2h_bytes_per_step": int(K * 4 * 8)},
               "gpu_launches": int(n_launch),
               "clocks": clocks,
               "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                            "kernel": "scan_kernel", "ms_per_launch": ms_per_launch, "kernel_share_of_step": (t_scan / args.steps) / ms_step,
                            "count_kernel_share_of_step": (t_cnt / args.steps) / ms_step,
                            "binding": {"bound": "smem-gather", "what": "2 B shared-memory table read + 1 Float16 add per PWM cell",
                                        "cells_per_s": cells_rate, "achieved": cells_rate * 2 / 1e9, "peak": smem_peak / 1e9, "unit": "GB/s",
                                        "frac": cells_rate * 2 / smem_peak, "peak_source": "148 SM x 128 B/clk x measured SM clock"}},
               "checks": {"counts_sum": [int(x) for x in counts.sum(axis=0)]}}
        if world == 1 and not args.no_cpu_baseline:
            sample = ascii_host[: min(n_local, 20000)].numpy()
            rate, n, cores, dt = cpu_scan_rate(pw, lens, thr, sample, args.cpu_seconds)
            out["cpu_baseline"] = {"value": rate, "unit": "bp/s", "cores": cores, "kind": "port", "seconds": dt,
                                   "sample": f"first {n} of {args.nseq} sequences x {Lb} bp, all {K} PWMs, both strands (oracle/scan_oracle.c, OpenMP)"}
            # the sample is also a parity check of the full-size run's first sequences
            _, oc = so.scan(pw, lens, so.ascii_to_codes(sample[:n]), thr, want_hits=False)
            s3 = ctx.seqs_from_ascii(sample[:n])
            _, gc = ctx.scan(s3, pw, lens, thr, want_hits=False)
            out["checks"]["sample_counts_match_oracle"] = bool(np.array_equal(oc, gc))
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
