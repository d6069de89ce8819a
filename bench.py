#!/usr/bin/env python
"""bench.py -- FSE (tANS) encode/decode throughput on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c4|c5|c2|c3few|c3uni|c1] [--no-configs] [--no-e2e] [--no-cpu]

Default workload = the configuration BASELINE.json's metric is quoted on: c4, 8 GiB of skewed (geometric 0.2)
bytes in 128 KiB blocks with per-block tables, STRONG-scaled: rank g of N owns the contiguous block range
[g*B/N, (g+1)*B/N) of the one logical stream (N = 1 codes all 8 GiB), no collective on the data path.

A "step" is one pass of the hot path over the resident input: compress (histogram + normalise + header +
table build + encode + offset scan + dense placement) and decompress it again (header parse + table build +
decode).  `value` = uncompressed GB per second of that round trip, whole job, device resident, CUDA events,
max over ranks.  `e2e` = the same round trip through the host-buffer entry points (fse_b200_compress_host /
fse_b200_decompress_host) with pinned host buffers, host<->device copies inside the timed region.
`configs` carries short runs of the other BASELINE.json configurations (c2, the c3 table_log sweep, c5).

--impl reference times the reference's CPU path on the box's host cores: the C restatement of the crate's
fse_compress2 / fse_decompress2 loops (oracle/, kind "port": the crate is Rust and no Rust toolchain exists
in this image), pthreads over blocks, built -O3 -march=native on the box.  Rank 0 alone runs it.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

GIB = 1 << 30
WORKLOADS = {
    # name: generator kind, seed, bytes (total when strong-scaled, per GPU when weak), block size, table_log,
    #       table mode, scaling, description (BASELINE.json config)
    "c1": dict(kind="geo", seed=0xC0FFEE01, nbytes=1 << 20, bs=1 << 20, tlog=11, tmode=0, scaling="weak",
               desc="1 MiB synthetic skewed bytes (geometric 0.2), one stream, table_log 11, fse_compress2 + fse_decompress2 (CPU only)"),
    "c2": dict(kind="text", seed=0xC0FFEE02, nbytes=256 << 20, bs=65536, tlog=0, tmode=0, scaling="weak",
               desc="256 MiB synthetic text-like bytes, 64 KiB blocks, per-block tables"),
    "c3few": dict(kind="few", seed=0xC0FFEE03, nbytes=GIB, bs=65536, tlog=11, tmode=0, scaling="weak",
                  desc="1 GiB low-entropy (few-symbol) bytes, 64 KiB blocks, per-block tables"),
    "c3uni": dict(kind="uniform", seed=0xC0FFEE03, nbytes=GIB, bs=65536, tlog=11, tmode=0, scaling="weak",
                  desc="1 GiB near-uniform random bytes, 64 KiB blocks, per-block tables"),
    "c4": dict(kind="geo", seed=0xC0FFEE04, nbytes=8 * GIB, bs=131072, tlog=0, tmode=0, scaling="strong",
               desc="8 GiB skewed (geometric 0.2) bytes, 128 KiB blocks, per-block tables, block-range sharded"),
    "c5": dict(kind="geo", seed=0xC0FFEE05, nbytes=8 * GIB, bs=131072, tlog=11, tmode=1, scaling="strong",
               desc="8 GiB skewed bytes, 128 KiB blocks, one global table via histogram all-reduce, block-range sharded"),
}
N_STATES = 128
SEGMENT_SIZE = 0          # > 0: code every block as block_size / SEGMENT_SIZE independent 128-state streams (fse_shared_enc.cuh); measured slower


def kernel_names(tmode, seg, tlog=0, nblocks=0, num_sms=148, smem_per_sm=233472, smem_optin=232448):
    """the kernels behind the timed spans: the dispatch rules of fse_b200.cu (compress / decompress_blocks_async) restated"""
    n, tl = N_STATES, (tlog or 11)
    if n == 128 and tmode == 1 and tl <= 11:
        enc, dec = "k_encode_sh_global", "k_decode_sh_global"          # CTA-owned bank-replicated tables
    elif n == 128 and seg:
        enc, dec = "k_encode_sh_blocks", "k_decode_sh_blocks"
    elif n == 128:
        enc = "k_encode128_blocks"                                     # one private table set per warp
        half = smem_per_sm // 2 - 1024

        def warps(per_warp):
            w = min(16, (min(half, smem_optin) - 64) // per_warp)
            return (w, 2) if w >= 1 else (min(16, (smem_optin - 64) // per_warp), 1)
        size = 1 << max(tl, 9)
        (wc, cc), (ww, cw) = (warps(3 * size + 1056) if tl <= 12 else (0, 2)), warps(4 * size + 1040)
        wide = wc < 1 or ww * cw >= wc * cc or (tl <= 11 and nblocks >= 2 * num_sms * cw * max(ww, 1))
        dec = "k_decode128_blocks" if wide else "k_decode128c_blocks"  # 32-bit entries / compact tables
    elif n == 64:
        enc, dec = "k_encode64_blocks", ("k_decode64c_blocks" if tl <= 12 else "k_decode64_blocks")
    elif n <= 2 and tmode == 0 and tl <= 12:
        enc = "k_tps_prepare_enc + k_tps_encode_smem"                  # one thread per stream, tables in shared memory (fse_tps.cuh)
        dec = "k_tps_prepare_dec + k_tps_decode_smem"
    else:
        enc, dec = "k_encode_blocks", "k_decode_blocks"
    return {"hist": "k_hist_blocks16", "encode": enc, "decode": dec, "scan": "k_scan_sizes", "gather": "k_gather"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def mem_available_gib():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) / (1 << 20)
    except Exception:
        pass
    return 0.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  nvidia-smi needs ~0.1 s before its first
    row, so it is started before the warm-up and every row is stamped on arrival; stop(t0, t1) keeps the rows that arrived
    inside the timed region.  A region too short for three rows (strong-scaled shards) is followed by the same steps,
    untimed, until there are three: those rows are taken under the same load and counted separately."""
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
            self.rows.append((time.time(), line.strip()))

    def count(self, t0, t1):
        return sum(1 for t, _ in list(self.rows) if t0 <= t <= t1)

    def stop(self, t0, t1, t2=None):
        """rows of [t0, t1] (the timed region), plus those of (t1, t2] (the same load continued) when given"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, inside = [], [], set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if not (t0 <= t <= (t2 if t2 is not None else t1)):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            inside += t <= t1
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": inside}
        if t2 is not None:
            out["note"] = "timed region of %.0f ms: the rest of the rows were taken while the same steps went on, untimed" % ((t1 - t0) * 1e3)
        return out


# ---------------------------------------------------------------------------------------------- CPU legs (oracle/)

_native = None


def native_oracle():
    """The timing copy of the CPU port: oracle/fse_oracle.c built -O3 -march=native ON THIS HOST (the portable build
    that travels with the repo is x86-64-v2).  Falls back to the portable build when gcc is missing."""
    global _native
    if _native is not None:
        return _native
    import oracle_lib as O
    L, flags = O.lib(), "-O3 -march=x86-64-v2 (portable build)"
    out = os.path.join(ROOT, "oracle", "_build", "libfse_oracle_native.so")
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-o", out,
                               os.path.join(ROOT, "oracle", "fse_oracle.c"), "-lpthread"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        L = C.CDLL(out)
        L.fse_or_compress_bound.argtypes = [C.c_size_t]
        L.fse_or_compress_bound.restype = C.c_size_t
        L.fse_or_ref_compress2.restype = C.c_long
        L.fse_or_ref_decompress2.restype = C.c_long
        flags = "-O3 -march=native (built on this host)"
    except Exception:
        pass
    _native = (L, flags)
    return _native


class CpuPort:
    """The reference's CPU path over blocks (C restatement: 2 interleaved states, 64-bit accumulator, one flush per
    symbol pair), pthreads over blocks.  Buffers are allocated and touched once, outside every timed region."""

    def __init__(self, kind, seed, block_size, nbytes, threads):
        import numpy as np
        import oracle_lib as O
        self.L, self.flags = native_oracle()
        self.np, self.O = np, O
        self.src = O.generate(kind, seed, nbytes)
        self.nb = (nbytes + block_size - 1) // block_size
        self.stride = int(self.L.fse_or_compress_bound(block_size)) + 64
        self.scratch = np.zeros((self.nb, self.stride), dtype=np.uint8)
        self.sizes = np.zeros(self.nb, dtype=np.uint64)
        self.status = np.zeros(self.nb, dtype=np.int32)
        self.out = np.zeros(nbytes, dtype=np.uint8)
        self.p = O.BlockParams(block_size, 0, 2, threads, 1)
        self.nbytes, self.threads = nbytes, threads

    def roundtrip(self, check=False):
        """-> (encode seconds, decode seconds)"""
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        L = self.L
        t0 = time.perf_counter()
        L.fse_or_compress_blocks(ptr(self.src), C.c_size_t(self.nbytes), C.byref(self.p), ptr(self.scratch),
                                 C.c_size_t(self.stride), ptr(self.sizes), ptr(self.status))
        t1 = time.perf_counter()
        L.fse_or_decompress_blocks(ptr(self.scratch), C.c_size_t(self.stride), ptr(self.sizes), C.c_size_t(self.nb),
                                   C.byref(self.p), ptr(self.out), C.c_size_t(self.nbytes), ptr(self.status))
        t2 = time.perf_counter()
        if check:
            assert not self.status.any() and self.np.array_equal(self.out, self.src), "CPU port round trip failed"
        return t1 - t0, t2 - t1

    def ratio(self):
        return float(self.sizes.sum()) / self.nbytes


def cpu_single_stream(kind, seed, nbytes, reps=10, warm=3):
    """One stream, one thread: fse_compress2 + fse_decompress2 (the shape of benches/fse_benchmark.rs:30-52)."""
    import numpy as np
    import oracle_lib as O
    L, _ = native_oracle()
    src = O.generate(kind, seed, nbytes)
    cap = int(L.fse_or_compress_bound(nbytes)) + 64
    comp = np.zeros(cap, dtype=np.uint8)
    out = np.zeros(nbytes + 64, dtype=np.uint8)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    be = bd = 1e9
    for i in range(warm + reps):
        t0 = time.perf_counter()
        ln = L.fse_or_ref_compress2(ptr(src), C.c_size_t(nbytes), ptr(comp), C.c_size_t(cap))
        t1 = time.perf_counter()
        got = L.fse_or_ref_decompress2(ptr(comp), C.c_size_t(ln), ptr(out), C.c_size_t(nbytes + 64))
        t2 = time.perf_counter()
        assert got == nbytes and np.array_equal(out[:nbytes], src)
        if i >= warm:
            be, bd = min(be, t1 - t0), min(bd, t2 - t1)
    return {"bytes": nbytes, "threads": 1, "encode_GBps": nbytes / be / 1e9, "decode_GBps": nbytes / bd / 1e9,
            "roundtrip_GBps": nbytes / (be + bd) / 1e9, "compressed_ratio": ln / nbytes, "best_of": reps, "warmups": warm}


def cpu_baseline_leg(w):
    """cpu_baseline of the main line: a bounded sample of the same workload, all host threads, best of 10 after 3
    warm-ups, plus the reference bench's own 32 KiB shape and c1, one thread each."""
    threads = os.cpu_count() or 1
    sample = min(w["nbytes"], 512 << 20)
    port = CpuPort(w["kind"], w["seed"], w["bs"], sample, threads)
    be = bd = 1e9
    for i in range(13):
        e, d = port.roundtrip(check=(i == 0))
        if i >= 3:
            be, bd = min(be, e), min(bd, d)
    shapes = {}
    try:
        shapes["fse_benchmark_32KiB"] = cpu_single_stream("geo", 0xC0FFEE00, 1 << 15, reps=50, warm=10)
        shapes["c1_1MiB"] = cpu_single_stream("geo", 0xC0FFEE01, 1 << 20)
    except Exception as ex:                                        # never lose the main number to an extra
        shapes["error"] = repr(ex)
    return {"value": sample / (be + bd) / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
            "encode_GBps": sample / be / 1e9, "decode_GBps": sample / bd / 1e9, "compressed_ratio": port.ratio(),
            "build": port.flags,
            "sample": "%d MiB of the same workload (%s), %d B blocks, fse_compress2 + fse_decompress2 loop structure (C "
                      "restatement of the Rust reference: no Rust toolchain in this image), %d threads over blocks, best of 10 "
                      "after 3 warm-ups" % (sample >> 20, w["kind"], w["bs"], threads),
            "single_thread_shapes": shapes}


def run_reference(args, wl):
    w = WORKLOADS[wl]
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    nbytes = w["nbytes"]                                           # the whole job, whatever --gpus says (strong scaling)
    if w["scaling"] == "weak":
        nbytes *= max(args.gpus, 1)
    sample = nbytes
    # host memory: source + strided scratch + output ~ 3.1 x the sample
    if mem_available_gib() * GIB < 3.5 * sample:
        sample = max(64 << 20, min(nbytes, int(mem_available_gib() * GIB / 4) // w["bs"] * w["bs"]))
    if wl == "c1":                                                 # a single stream is serial: one core
        threads = 1
    port = CpuPort(w["kind"], w["seed"], w["bs"], sample, threads)
    for i in range(max(args.warmup, 1)):
        port.roundtrip(check=(i == 0))
    te = td = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e, d = port.roundtrip()
        te, td = te + e, td + d
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    rt = sample * args.steps / (te + td) / 1e9
    what = "the whole workload" if sample == nbytes else "%d MiB of the workload (host memory bound)" % (sample >> 20)
    line = {
        "impl": "reference", "metric": "fse_roundtrip_GBps_uncompressed", "value": rt, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl + ": " + w["desc"], "total_bytes": nbytes, "block_size": w["bs"], "n_states": 2,
                   "sample_bytes_per_step": sample},
        "encode_GBps": sample * args.steps / te / 1e9, "decode_GBps": sample * args.steps / td / 1e9,
        "compressed_ratio": port.ratio(),
        "cpu_baseline": {"value": rt, "unit": "GB/s", "cores": threads, "kind": "port", "build": port.flags,
                         "sample": "%s per step: %s, %d B blocks, fse_compress2 + fse_decompress2 loops (C restatement; the "
                                   "reference is Rust, no toolchain in this image), %d threads over blocks"
                                   % (what, w["kind"], w["bs"], threads)},
        "e2e": {"value": rt, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU legs

class Job:
    """One workload resident on this rank's GPU: buffers, the step, and its measurements."""

    def __init__(self, env, wl, tlog=None):
        import torch
        from entropy_coders_b200 import sharding as S
        self.env, self.wl, self.torch, self.S = env, wl, torch, S
        w = dict(WORKLOADS[wl])
        if tlog is not None:
            w["tlog"] = tlog
        self.w = w
        ctx, world, rank = env["ctx"], env["world"], env["rank"]
        bs = w["bs"]
        if w["scaling"] == "strong":
            self.total_bytes = w["nbytes"]
            _, _, first, nbytes = S.shard_blocks(self.total_bytes, bs, rank, world)
        else:
            self.total_bytes = w["nbytes"] * world
            first, nbytes = rank * w["nbytes"], w["nbytes"]
        self.nbytes, self.first = nbytes, first
        dev = env["dev"]
        self.src = ctx.generate(w["kind"], w["seed"], nbytes, first_index=first)
        self.seg = int(env.get("segment_size", 0)) if w["tmode"] == 0 else 0
        if self.seg and bs % self.seg:
            self.seg = 0
        self.p = ctx.params(bs, w["tlog"], N_STATES, w["tmode"], self.seg)
        self.nb = ctx.num_streams(nbytes, self.p)
        self.cap = ctx.bound(nbytes, self.p)
        self.dst = torch.empty(self.cap, dtype=torch.uint8, device=dev)
        self.offsets = torch.empty(self.nb + 1, dtype=torch.int64, device=dev)
        self.status = torch.empty(max(self.nb, 1), dtype=torch.int32, device=dev)
        self.status_d = torch.empty(max(self.nb, 1), dtype=torch.int32, device=dev)
        self.out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.totals = torch.zeros(world, dtype=torch.int64, device=dev)
        self.mine = torch.zeros(1, dtype=torch.int64, device=dev)
        self.ev_c = torch.cuda.Event()
        self.ev_x = torch.cuda.Event()
        self.pending_exchange = False

    def global_table(self):
        ctx = self.env["ctx"]
        counts = ctx.histogram_global(self.src)
        self.S.allreduce_histogram(counts)                # NCCL: the only exchange of the global-table mode
        return ctx.set_global_table(counts, self.w["tlog"])

    def step(self):
        env, torch = self.env, self.torch
        ctx, stream, side, world = env["ctx"], env["stream"], env["side"], env["world"]
        if self.w["tmode"] == 1:
            self.global_table()
        ctx.compress_blocks_async(self.src, self.p, self.dst, self.offsets, self.status)
        if world > 1:
            # Placement of this rank's output in the logical stream: all-gather of the per-rank totals on a side
            # stream.  Nothing on the data path consumes it, so the main stream never waits for it inside the
            # step; it is joined once, after the timed region (finish()).
            self.ev_c.record(stream)
            with torch.cuda.stream(side):
                side.wait_event(self.ev_c)
                self.mine.copy_(self.offsets[self.nb:self.nb + 1])
                torch.distributed.all_gather_into_tensor(self.totals, self.mine)
                self.ev_x.record(side)
            self.pending_exchange = True
        ctx.decompress_blocks_async(self.dst, self.cap, self.offsets, self.nb, self.p, self.out, self.nbytes, self.status_d)

    def finish(self):
        if self.pending_exchange:
            self.env["stream"].wait_event(self.ev_x)
            self.pending_exchange = False

    def check(self):
        torch = self.torch
        assert not self.status[:self.nb].cpu().numpy().any() and not self.status_d[:self.nb].cpu().numpy().any(), "block failures"
        assert torch.equal(self.out, self.src), "round trip mismatch"
        return int(self.offsets[self.nb].item())

    def base_offset(self):
        return int((torch_cumsum_exclusive(self.totals))[self.env["rank"]].item()) if self.env["world"] > 1 else 0

    def release(self):
        for k in ("src", "dst", "offsets", "status", "status_d", "out"):
            setattr(self, k, None)
        self.torch.cuda.empty_cache()


def torch_cumsum_exclusive(t):
    import torch
    return torch.cumsum(t, 0) - t


def barrier(env):
    import torch
    torch.cuda.synchronize()
    if env["world"] > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def measure(env, job, steps, warmup, sample_clocks=False):
    """warmup untimed steps, then exactly `steps` timed steps between barriers; max over ranks."""
    import torch
    ctx, stream, world, rank = env["ctx"], env["stream"], env["world"], env["rank"]
    for _ in range(max(warmup, 3)):
        job.step()
    job.finish()
    barrier(env)
    total = job.check()
    sampler = ClockSampler(env["local"]) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()                                      # before two more warm-up steps: nvidia-smi needs ~0.1 s to its first row
    if sample_clocks:                                        # on every rank (a step may carry the size exchange)
        for _ in range(2):
            job.step()
        job.finish()
    ctx.set_timing(True)
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(env)
    t0 = time.time()
    e0.record(stream)
    for _ in range(steps):
        job.step()
    job.finish()
    e1.record(stream)
    barrier(env)
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    tm = ctx.get_timing()
    ctx.set_timing(False)
    t = torch.tensor([ms, float(total)], dtype=torch.float64, device=env["dev"])
    tmax = t.clone()
    if world > 1:
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    ms_step = float(tmax[0].item()) / steps
    total_all = float(t[1].item()) if world > 1 else float(total)
    clocks = None
    if sample_clocks:
        # a timed region too short for three nvidia-smi rows (20 ms apart) is followed by the same steps, untimed, on every
        # rank (the decision depends on the all-reduced time only), and the rows of that stretch are taken as well
        t2 = None
        if ms_step * steps < 250.0:
            for _ in range(int(300.0 / ms_step) + 1):
                job.step()
            job.finish()
            barrier(env)
            t2 = time.time()
        if sampler:
            clocks = sampler.stop(t0, t1, t2)
    enc_ms = sum(tm[k][0] for k in ("hist", "encode", "scan", "gather")) / steps
    dec_ms = tm["decode"][0] / steps
    peak, peak_src = peaks()
    dom = max(("hist", "encode", "decode"), key=lambda k: tm[k][0])
    dom_ms = tm[dom][0] / max(tm[dom][1], 1)
    alg_bytes = job.nbytes + total if dom != "hist" else job.nbytes
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    return {
        "value": job.total_bytes / (ms_step * 1e-3) / 1e9, "ms_per_step": ms_step, "steps": steps,
        "encode_GBps": job.nbytes / (enc_ms * 1e-3) / 1e9, "decode_GBps": job.nbytes / (dec_ms * 1e-3) / 1e9,
        "compressed_ratio": total_all / job.total_bytes, "compressed_bytes_rank0": total,
        "kernel_ms_per_step": {k: tm[k][0] / steps for k in tm},
        "roofline": {"bound": "hbm", "kernel": kernel_names(job.w["tmode"], job.seg, job.w["tlog"], job.nb)[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms,
                     "direction_frac": {"encode": (job.nbytes + total) / (enc_ms * 1e-3) / 1e9 / peak,
                                        "decode": (job.nbytes + total) / (dec_ms * 1e-3) / 1e9 / peak},
                     "note": "N + C of one rank's launch (uncompressed + compressed bytes; N alone for the histogram) / mean "
                             "launch time of the kernel with the largest share of the step; direction_frac = the same bytes "
                             "over ALL kernels of the direction"},
        "gpu_launches": int(launches), "clocks": clocks, "dominant": dom,
    }


def e2e_leg(env, job, steps):
    """The round trip through the host-buffer entry points; pinned host buffers, copies inside the timed region."""
    import torch
    import entropy_coders_b200 as E
    w, nbytes, nb = job.w, job.nbytes, job.nb
    total = int(job.offsets[nb].item())
    sample = nbytes
    if mem_available_gib() * GIB < 4.0 * nbytes * max(env["world"], 1):    # pinned: source + compressed + output per rank
        sample = max(w["bs"], min(nbytes, int(mem_available_gib() * GIB / (5 * env["world"])) // w["bs"] * w["bs"]))
    hsrc = torch.empty(sample, dtype=torch.uint8, pin_memory=True)
    hsrc.copy_(job.src[:sample])
    hdst = torch.empty(int(total * (sample / nbytes) * 1.05) + (1 << 20), dtype=torch.uint8, pin_memory=True)
    hout = torch.empty(sample, dtype=torch.uint8, pin_memory=True)
    ctx2 = E.Context(env["local"])
    if w["tmode"] == 1:
        hdr, _ = job.global_table()
        ctx2.set_global_table_from_header(hdr)
    tot = offs = None

    phase = [0.0, 0.0]

    def e2e_step():
        ta = time.perf_counter()
        _, o, st, t = ctx2.compress_host(hsrc, w["bs"], w["tlog"], N_STATES, w["tmode"], dst=hdst, segment_size=job.seg)
        tb = time.perf_counter()
        ctx2.decompress_host(hdst, t, o, sample, w["bs"], w["tlog"], N_STATES, w["tmode"], dst=hout, segment_size=job.seg)
        phase[0] += tb - ta
        phase[1] += time.perf_counter() - tb
        return t, o
    for _ in range(2):
        e2e_step()
    ksteps = max(1, min(steps, 3 if sample > GIB else 5))
    barrier(env)
    phase[0] = phase[1] = 0.0
    t0 = time.perf_counter()
    for _ in range(ksteps):
        tot, offs = e2e_step()
    t1 = time.perf_counter()
    assert torch.equal(hout, hsrc), "e2e round trip mismatch"
    te = torch.tensor([t1 - t0], dtype=torch.float64, device=env["dev"])
    if env["world"] > 1:
        torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
    e2e_s = float(te.item()) / ksteps
    snb = (sample + (job.seg or w["bs"]) - 1) // (job.seg or w["bs"])
    # the raw link: the same pinned buffers copied each way on every rank at once (what bounds the e2e figure)
    link = None
    try:
        lb = min(sample, 1 << 30)
        dtmp = torch.empty(lb, dtype=torch.uint8, device=env["dev"])
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        barrier(env)
        ev[0].record()
        dtmp.copy_(hsrc[:lb], non_blocking=True)
        ev[1].record()
        hout[:lb].copy_(dtmp, non_blocking=True)
        ev[2].record()
        torch.cuda.synchronize()
        # both directions at once (what the chunked pipelines of the two calls actually see)
        dtmp2 = torch.empty(lb, dtype=torch.uint8, device=env["dev"])
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        dv = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        barrier(env)
        with torch.cuda.stream(s_up):
            dv[0].record()
            dtmp.copy_(hsrc[:lb], non_blocking=True)
            dv[1].record()
        with torch.cuda.stream(s_dn):
            dv[2].record()
            hout[:lb].copy_(dtmp2, non_blocking=True)
            dv[3].record()
        torch.cuda.synchronize()
        bw = torch.tensor([lb / ev[0].elapsed_time(ev[1]) / 1e6, lb / ev[1].elapsed_time(ev[2]) / 1e6,
                           lb / dv[0].elapsed_time(dv[1]) / 1e6, lb / dv[2].elapsed_time(dv[3]) / 1e6], dtype=torch.float64, device=env["dev"])
        if env["world"] > 1:
            torch.distributed.all_reduce(bw, op=torch.distributed.ReduceOp.MIN)
        h2d, d2h, h2d_dx, d2h_dx = (float(x) for x in bw.tolist())
        moved = float(sample + tot)                          # per direction and rank: N + C each way over a round trip
        # the two calls run one after the other (the reference's API shape): compress moves N up while C comes down,
        # decompress moves C up while N comes down; each call is bounded by its larger transfer at the duplex rate
        def two_calls(up, down):
            c_s = max(sample / up, tot / down) / 1e9
            d_s = max(tot / up, sample / down) / 1e9
            return env["world"] * sample / (c_s + d_s) / 1e9
        link = {"h2d_GBps_min_rank": h2d, "d2h_GBps_min_rank": d2h,
                "h2d_duplex_GBps_min_rank": h2d_dx, "d2h_duplex_GBps_min_rank": d2h_dx,
                "round_trip_bound_GBps": env["world"] * sample / (moved / (min(h2d, d2h) * 1e9)) / 1e9,
                "two_call_bound_GBps": two_calls(h2d, d2h),
                "two_call_duplex_GBps": two_calls(h2d_dx, d2h_dx),
                "note": "pinned copies of 1 GiB on all ranks at once, one direction at a time and both at once (duplex); "
                        "round_trip_bound = N / ((N + C) / slower direction) if the two calls could overlap each other; "
                        "two_call_* = N / (max(N/up, C/down) + max(C/up, N/down)): the calls as the API has them, back to back, "
                        "at the one-way rates (the bound) and at the rates with both directions saturated (the smaller "
                        "transfer of a call loads the other direction only part of the time: e2e lies between the two)"}
        del dtmp, dtmp2
    except Exception as ex:
        link = {"error": repr(ex)}
    res = {"value": env["world"] * sample / e2e_s / 1e9 if w["scaling"] == "weak" or sample != nbytes
           else job.total_bytes / e2e_s / 1e9,
           "unit": "GB/s", "h2d_bytes_per_step": int(sample + tot + (snb + 1) * 8),
           "d2h_bytes_per_step": int(tot + (snb + 1) * 8 + snb * 4 + sample + snb * 4),
           "steps": ksteps, "bytes_per_rank": sample,
           "compress_host_ms": phase[0] / ksteps * 1e3, "decompress_host_ms": phase[1] / ksteps * 1e3,
           "pcie": link, "numa_node": env.get("numa"),
           "api": "fse_b200_compress_host + fse_b200_decompress_host, pinned host buffers, rank-local data"}
    ctx2.close()
    del hsrc, hdst, hout
    return res


def traffic_probe(wl, tlog, dom_kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, measured now: a child process
    runs one step of the same workload under `ncu --metrics ...` (N = 1 only).  None when ncu is unavailable."""
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    log = os.path.join(ROOT, "gpurun_out", "traffic_probe.csv")
    os.makedirs(os.path.dirname(log), exist_ok=True)
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv",
           "--log-file", log, "-k", "regex:k_(hist|encode|decode)", "-c", "12",
           sys.executable, os.path.abspath(__file__), "--probe", "--workload", wl]
    if tlog is not None:
        cmd += ["--table-log", str(tlog)]
    cmd += ["--segment-size", str(SEGMENT_SIZE), "--n-states", str(N_STATES)]
    try:
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=420, check=True)
        import csv
        rd = wr = None
        with open(log) as f:
            rows = [r for r in csv.reader(f) if len(r) > 5]
        hdr = rows[0]
        ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[1:]:
            if r[ki].startswith(dom_kernel):
                v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
                if r[mi] == "dram__bytes_read.sum":
                    rd = v                                  # the last launch of the kernel wins (after the warm-up launch)
                elif r[mi] == "dram__bytes_write.sum":
                    wr = v
        if rd is None or wr is None:
            return None, "kernel not in the ncu log"
        return rd + wr, "ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch, measured in this run"
    except Exception as ex:
        return None, "probe failed: %r" % (ex,)


def bind_to_gpu_numa(local):
    """Pin this process to the CPUs that are local to its GPU (sysfs local_cpulist of the PCI device), so that the pinned
    host buffers of the e2e leg are first-touched on the GPU's own NUMA node.  Returns the node number or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return int(open(base + "/numa_node").read().strip())
    except Exception:
        return None


def make_env():
    import torch
    import torch.distributed as dist
    import entropy_coders_b200 as E
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is made: send it to stderr so
        # that stdout carries exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    stream = torch.cuda.Stream(device=dev)               # kernels, NCCL and the timing events share this stream
    torch.cuda.set_stream(stream)
    side = torch.cuda.Stream(device=dev)
    ctx = E.Context(local, stream=stream.cuda_stream)
    return {"segment_size": SEGMENT_SIZE, "numa": numa, "world": world, "rank": rank, "local": local, "dev": dev, "stream": stream, "side": side, "ctx": ctx}


def run_probe(args, wl):
    """child of traffic_probe(): one warm-up step and one profiled step, nothing printed"""
    env = make_env()
    job = Job(env, wl, args.table_log)
    job.step()
    job.step()
    barrier(env)
    job.check()


def run_ours(args, wl):
    global N_STATES
    import torch
    env = make_env()
    world, rank = env["world"], env["rank"]
    job = Job(env, wl, args.table_log)
    w = job.w
    m = measure(env, job, args.steps, args.warmup, sample_clocks=True)
    if rank == 0 and world == 1 and not args.no_traffic:
        m["roofline"]["traffic"], m["roofline"]["traffic_source"] = traffic_probe(wl, args.table_log, m["roofline"]["kernel"])

    e2e = None
    if not args.no_e2e:
        try:
            e2e = e2e_leg(env, job, args.steps)
        except Exception as ex:
            e2e = {"error": repr(ex)}
            barrier(env)
    job.release()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline_leg(w)
        except Exception as ex:
            cpu = {"error": repr(ex)}

    # the other BASELINE.json configurations, short runs (every rank takes part: c5 has a collective)
    configs = {}
    if not args.no_configs and wl == "c4":
        plan = [("c2", None), ("c5", None)] + [(k, t) for k in ("c3few", "c3uni") for t in (9, 11, 12)]
        for name, tl in plan:
            key = name if tl is None else "%s_tl%d" % (name, tl)
            try:
                j = Job(env, name, tl)
                r = measure(env, j, 5, 3)
                configs[key] = {"workload": name + ": " + j.w["desc"], "scaling": j.w["scaling"], "total_bytes": j.total_bytes,
                                "block_size": j.w["bs"], "table_log": tl if tl is not None else (j.w["tlog"] or "optimal_log2 (11)"),
                                "n_states": N_STATES, "value": r["value"], "unit": "GB/s", "ms_per_step": r["ms_per_step"],
                                "encode_GBps_per_gpu": r["encode_GBps"], "decode_GBps_per_gpu": r["decode_GBps"],
                                "compressed_ratio": r["compressed_ratio"], "kernel_ms_per_step": r["kernel_ms_per_step"],
                                "roofline_frac": r["roofline"]["frac"], "roofline_kernel": r["roofline"]["kernel"]}
                j.release()
            except Exception as ex:
                configs[key] = {"error": repr(ex)}
                barrier(env)
        # c4 in the reference's OWN stream formats (fse_compress2 / fse_compress: two states / one state per block, the
        # bytes a user of the crate has): one thread per stream, tables in shared memory (fse_tps.cuh)
        keep = N_STATES
        for ns in (2, 1):
            key = "c4_reference_format_%dstate" % ns
            try:
                N_STATES = ns
                j = Job(env, "c4")
                r = measure(env, j, 3, 3)
                configs[key] = {"workload": "c4: " + j.w["desc"] + "; block format = the crate's fse_compress%s" % ("2" if ns == 2 else ""),
                                "scaling": j.w["scaling"], "total_bytes": j.total_bytes, "block_size": j.w["bs"], "n_states": ns,
                                "value": r["value"], "unit": "GB/s", "ms_per_step": r["ms_per_step"],
                                "encode_GBps_per_gpu": r["encode_GBps"], "decode_GBps_per_gpu": r["decode_GBps"],
                                "compressed_ratio": r["compressed_ratio"], "kernel_ms_per_step": r["kernel_ms_per_step"],
                                "roofline_frac": r["roofline"]["frac"], "roofline_kernel": r["roofline"]["kernel"]}
                j.release()
            except Exception as ex:
                configs[key] = {"error": repr(ex)}
                barrier(env)
            finally:
                N_STATES = keep

    if rank == 0:
        line = {
            "metric": "fse_roundtrip_GBps_uncompressed", "value": m["value"], "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": w["scaling"], "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl + ": " + w["desc"], "total_bytes": job.total_bytes, "bytes_per_gpu": job.nbytes,
                       "block_size": w["bs"], "blocks_per_gpu": job.nb, "n_states": N_STATES, "segment_size": job.seg,
                       "table_log": w["tlog"] or "optimal_log2 (11)", "table_mode": "global" if w["tmode"] else "per-block",
                       "sharding": "rank g owns blocks [g*B/N, (g+1)*B/N) of the one logical stream",
                       "l2": "inputs (%d MiB per GPU) larger than L2 (126 MB); no flush" % (job.nbytes >> 20)},
            "encode_GBps": m["encode_GBps"], "decode_GBps": m["decode_GBps"], "compressed_ratio": m["compressed_ratio"],
            "kernel_ms_per_step": m["kernel_ms_per_step"], "roofline": m["roofline"], "e2e": e2e, "cpu_baseline": cpu,
            "gpu_launches": m["gpu_launches"], "clocks": m["clocks"], "configs": configs,
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--table-log", type=int, default=None, help="override the workload's table_log (BASELINE config 3 sweeps 9/11/12)")
    ap.add_argument("--n-states", type=int, default=None, help="interleaved states per block (default 128; 1 and 2 are the reference's own formats)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-traffic", action="store_true")
    ap.add_argument("--probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--segment-size", type=int, default=None,
                    help="bytes per independently coded segment of a block (per-block tables); 0 = one stream per block")
    args = ap.parse_args()
    global SEGMENT_SIZE, N_STATES
    if args.n_states is not None:
        N_STATES = args.n_states
    if args.segment_size is not None:
        SEGMENT_SIZE = args.segment_size
    if args.impl == "reference":
        run_reference(args, args.workload)
    elif args.workload == "c1":
        sys.exit("c1 is the reference's single-stream CPU case: run it with --impl reference (the GPU path codes it "
                 "bit-exactly in tests/test_gpu_parity.py::test_crate_compress_roundtrip, one lane, not a throughput case)")
    elif args.probe:
        run_probe(args, args.workload)
    else:
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()
