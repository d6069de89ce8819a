#!/usr/bin/env python
"""bench.py -- FSE (tANS) encode/decode throughput on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3few|c3uni|c4|c5]

A "step" is one pass of the hot path over one batch: compress the resident input (histogram +
normalise + header + table build + encode + offset scan + gather) and decompress it again (header
parse + table build + decode).  `value` is uncompressed GB per second of that round trip with the
input resident in HBM; `encode_GBps` / `decode_GBps` give the two directions on their own.
`e2e` is the same round trip through the host-buffer entry points (fse_b200_compress_host /
fse_b200_decompress_host) with pinned host buffers, copies inside the timed region.

N > 1 (launched under torchrun): every rank owns a contiguous block range of the logical stream
(its own 256 MiB slice), no collective on the data path; one all-gather of the per-rank compressed
totals per step places the output.  Weak scaling.

--impl reference times the CPU path: the C restatement of the reference crate's fse_compress2 /
fse_decompress2 loops (oracle/, "port" -- the crate is Rust and cannot be built in this image),
multithreaded over blocks on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

WORKLOADS = {
    # name: (generator kind, seed, bytes per GPU, block size, table_log, table_mode, BASELINE.json config)
    # c1 is the reference's own CPU case (benches/fse_benchmark.rs): one stream, two states; only --impl reference runs it
    "c1": ("geo", 0xC0FFEE01, 1 << 20, 1 << 20, 11, 0, "1 MiB synthetic skewed bytes (geometric 0.2), one stream, table_log 11, fse_compress2 + fse_decompress2"),
    "c2": ("text", 0xC0FFEE02, 256 << 20, 65536, 0, 0, "256 MiB synthetic text-like bytes, 64 KiB blocks, per-block tables"),
    "c3few": ("few", 0xC0FFEE03, 1 << 30, 65536, 11, 0, "1 GiB low-entropy (few-symbol) bytes, 64 KiB blocks"),
    "c3uni": ("uniform", 0xC0FFEE03, 1 << 30, 65536, 11, 0, "1 GiB near-uniform random bytes, 64 KiB blocks"),
    "c4": ("geo", 0xC0FFEE04, 1 << 30, 131072, 0, 0, "skewed (geometric 0.2) bytes, 128 KiB blocks, block-range sharded, 1 GiB per GPU"),
    "c5": ("geo", 0xC0FFEE05, 1 << 30, 131072, 11, 1, "skewed bytes, 128 KiB blocks, one global table via histogram all-reduce, 1 GiB per GPU"),
}
N_STATES = 128


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_throughput(kind, seed, block_size, sample_bytes, threads, reps=1):
    """The reference's CPU path (C restatement, 2 interleaved states, 64-bit accumulator), threads over
    blocks.  Returns (round-trip GB/s, encode GB/s, decode GB/s, compressed/uncompressed)."""
    import ctypes as C
    import numpy as np
    import oracle_lib as O
    L = O.lib()
    src = O.generate(kind, seed, sample_bytes)
    nb = (src.size + block_size - 1) // block_size
    stride = L.fse_or_compress_bound(block_size) + 64
    scratch = np.zeros((nb, stride), dtype=np.uint8)          # allocated and touched outside the timed region
    sizes = np.zeros(nb, dtype=np.uint64)
    status = np.zeros(nb, dtype=np.int32)
    out = np.zeros(src.size, dtype=np.uint8)
    p = O.BlockParams(block_size, 0, 2, threads, 1)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    best_e, best_d = 1e9, 1e9
    for _ in range(reps + 1):                                  # first pass warms caches / page tables
        t0 = time.perf_counter()
        L.fse_or_compress_blocks(ptr(src), src.size, C.byref(p), ptr(scratch), stride, ptr(sizes), ptr(status))
        t1 = time.perf_counter()
        L.fse_or_decompress_blocks(ptr(scratch), stride, ptr(sizes), nb, C.byref(p), ptr(out), src.size, ptr(status))
        t2 = time.perf_counter()
        assert not status.any() and np.array_equal(out, src)
        best_e, best_d = min(best_e, t1 - t0), min(best_d, t2 - t1)
    ratio = float(sizes.sum()) / src.size
    return sample_bytes / (best_e + best_d) / 1e9, sample_bytes / best_e / 1e9, sample_bytes / best_d / 1e9, ratio


def run_reference(args, wl):
    kind, seed, nbytes, bs, tlog, tmode, desc = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 64 << 20
    if wl == "c1":                                            # a single stream is serial: one core, the whole MiB
        threads, sample = 1, nbytes
    vals = []
    for _ in range(args.warmup):
        cpu_port_throughput(kind, seed, bs, sample, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_port_throughput(kind, seed, bs, sample, threads))
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    rt = sum(v[0] for v in vals) / len(vals)
    line = {
        "impl": "reference", "metric": "fse_roundtrip_GBps_uncompressed", "value": rt, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl + ": " + desc, "block_size": bs, "n_states": 2,
                   "sample": "%d MiB of the workload per step" % (sample >> 20)},
        "encode_GBps": sum(v[1] for v in vals) / len(vals), "decode_GBps": sum(v[2] for v in vals) / len(vals),
        "compressed_ratio": vals[-1][3],
        "cpu_baseline": {"value": rt, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": "%d MiB of %s, %d B blocks, fse_compress2+fse_decompress2 loops (C restatement; the "
                                   "reference is Rust, no toolchain in this image), %d threads over blocks" % (sample >> 20, kind, bs, threads)},
        "e2e": {"value": rt, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import entropy_coders_b200 as E

    kind, seed, nbytes, bs, tlog, tmode, desc = WORKLOADS[wl]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is made: send it to stderr so
        # that stdout carries exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device=torch.device("cuda", local)))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)               # kernels, NCCL and the timing events share this stream
    torch.cuda.set_stream(stream)
    ctx = E.Context(local, stream=stream.cuda_stream)

    # this rank's contiguous block range of the logical stream
    src = ctx.generate(kind, seed, nbytes, first_index=rank * nbytes)
    p = ctx.params(bs, tlog, N_STATES, tmode)
    nb = ctx.num_blocks(nbytes, bs)
    cap = ctx.bound(nbytes, p)
    dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    offsets = torch.empty(nb + 1, dtype=torch.int64, device=dev)
    status = torch.empty(nb, dtype=torch.int32, device=dev)
    status_d = torch.empty(nb, dtype=torch.int32, device=dev)
    out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    totals = torch.zeros(world, dtype=torch.int64, device=dev)

    from entropy_coders_b200 import sharding as S

    def global_table():
        counts = ctx.histogram_global(src)
        S.allreduce_histogram(counts)                    # NCCL: the only exchange of the global-table mode
        return ctx.set_global_table(counts, tlog)

    side = torch.cuda.Stream(device=dev)                 # the tiny exchange runs beside the decode kernel
    ev_c, ev_x = torch.cuda.Event(), torch.cuda.Event()

    def step():
        if tmode == 1:
            global_table()
        ctx.compress_blocks_async(src, p, dst, offsets, status)
        if world > 1:                                    # place the output: all-gather of the per-rank totals,
            ev_c.record(stream)                          # exclusive scan -> this rank's base offset
            with torch.cuda.stream(side):
                side.wait_event(ev_c)
                totals.copy_(S.gather_totals(offsets[nb:nb + 1], dev))
                S.base_offsets(totals)
                ev_x.record(side)
        ctx.decompress_blocks_async(dst, cap, offsets, nb, p, out, nbytes, status_d)
        if world > 1:
            stream.wait_event(ev_x)                      # the step is complete when both are

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    total = int(offsets[nb].item())
    assert not status.cpu().numpy().any() and not status_d.cpu().numpy().any(), "block failures"
    assert torch.equal(out, src), "round trip mismatch"

    sampler = ClockSampler(local)
    ctx.set_timing(True)
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    tm = ctx.get_timing()
    ctx.set_timing(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_step = ms / args.steps
    value = world * nbytes / (ms_step * 1e-3) / 1e9

    # per-direction device times (sum of the kernels of each direction, CUDA events on the launching stream)
    enc_ms = sum(tm[k][0] for k in ("hist", "encode", "scan", "gather")) / args.steps
    dec_ms = tm["decode"][0] / args.steps
    peak, peak_src = peaks()
    # roofline of the dominant kernel: algorithmic bytes (N + C per direction) / its duration
    dom = max(("encode", "decode"), key=lambda k: tm[k][0])
    dom_ms = tm[dom][0] / max(tm[dom][1], 1)
    alg_bytes = nbytes + total
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    # workload (profiles/r1_ncu_full_c2_n128.txt); only meaningful for the configuration that was profiled
    traffic = {"encode": 272.98e6 + 143.34e6, "decode": 179.75e6 + 219.48e6}[dom] if (wl == "c2" and N_STATES == 128) else None

    # e2e: host buffers through the host entry points, copies inside the timed region (rank-local data)
    e2e = None
    if not args.no_e2e:
        hsrc = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        hsrc.copy_(src)
        hdst = torch.empty(total + 4096, dtype=torch.uint8, pin_memory=True)
        hout = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        ctx2 = E.Context(local)
        if tmode == 1:
            hdr, _ = global_table()
            ctx2.set_global_table_from_header(hdr)

        def e2e_step():
            _, offs, st, tot = ctx2.compress_host(hsrc, bs, tlog, N_STATES, tmode, dst=hdst)
            o, st2 = ctx2.decompress_host(hdst, tot, offs, nbytes, bs, tlog, N_STATES, tmode, dst=hout)
            return tot, offs
        for _ in range(2):
            e2e_step()
        ksteps = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            tot, offs = e2e_step()
        t1 = time.perf_counter()
        assert torch.equal(hout, hsrc), "e2e round trip mismatch"
        te = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item()) / ksteps
        e2e = {"value": world * nbytes / e2e_s / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(nbytes + tot + (nb + 1) * 8),
               "d2h_bytes_per_step": int(tot + (nb + 1) * 8 + nb * 4 + nbytes + nb * 4),
               "steps": ksteps, "api": "fse_b200_compress_host + fse_b200_decompress_host, pinned host buffers"}
        ctx2.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sample = 64 << 20
        rt, ce, cd, _ = cpu_port_throughput(kind, seed, bs, sample, threads, reps=2)
        cpu = {"value": rt, "unit": "GB/s", "cores": threads, "kind": "port", "encode_GBps": ce, "decode_GBps": cd,
               "sample": "%d MiB of the same workload, %d B blocks, fse_compress2+fse_decompress2 loop structure "
                         "(C restatement of the Rust reference), %d threads over blocks, best of 2" % (sample >> 20, bs, threads)}

    if rank == 0:
        line = {
            "metric": "fse_roundtrip_GBps_uncompressed", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl + ": " + desc, "bytes_per_gpu": nbytes, "block_size": bs, "blocks_per_gpu": nb,
                       "n_states": N_STATES, "table_log": tlog or "optimal_log2 (11)", "table_mode": "global" if tmode else "per-block",
                       "l2": "inputs (%d MiB) larger than L2 (126 MB); no flush" % (nbytes >> 20)},
            "encode_GBps": nbytes / (enc_ms * 1e-3) / 1e9, "decode_GBps": nbytes / (dec_ms * 1e-3) / 1e9,
            "compressed_ratio": total / nbytes,
            "kernel_ms_per_step": {k: tm[k][0] / args.steps for k in tm},
            "roofline": {"bound": "hbm", "kernel": {"encode": "k_encode128_blocks", "decode": "k_decode128c_blocks"}[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "N + C per launch (uncompressed + compressed bytes of one rank) / mean launch time of the dominant kernel"},
            "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--table-log", type=int, default=None, help="override the workload's table_log (BASELINE config 3 sweeps 9/11/12)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.table_log is not None:
        w = list(WORKLOADS[args.workload])
        w[4] = args.table_log
        WORKLOADS[args.workload] = tuple(w)
    if args.impl == "reference":
        run_reference(args, args.workload)
    elif args.workload == "c1":
        sys.exit("c1 is the reference's single-stream CPU case: run it with --impl reference (the GPU path codes it "
                 "bit-exactly in tests/test_gpu_parity.py::test_crate_compress_roundtrip, one lane, not a throughput case)")
    else:
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()
