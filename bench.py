#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its named config.

  python bench.py --gpus N --steps K --warmup W            (ours; torchrun-launched for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path)

Workload (config.workload "text64m"): BASELINE config 3, one 64 MiB English-like text block
(SURVEY 8d C3 generator, seed 0x5EED0064 + rank) compressed as ONE whole-file BWT block and
decompressed again.  A step = compress + decompress of that block; the metric is uncompressed
input MB/s over the round trip (10^6 bytes / s), counted only when the compressed file equals the
reference's golden bytes and the round trip is bit exact (checked during warm-up, every rank).
  value : inputs resident in HBM (bzap_compress_device / bzap_decompress_device), CUDA events
  e2e   : same through the host-buffer C ABI (bzap_compress / bzap_decompress) from pinned host
          memory, H2D and D2H copies inside the timed region
N > 1: independent files, one per rank per step, no collective on the data path (weak scaling).
The Calgary batch (BASELINE config 2) is measured after the timed region and reported under
"calgary_batch" on the same line.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "compress+decompress round-trip throughput of uncompressed bytes (bit-exact, byte-identical file)"
UNIT = "MB/s"
N_TEXT = 1 << 26
BASE_SEED = 0x5EED0064


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 8:
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    # stdout carries exactly ONE line, the JSON result: anything a library prints there during the run
    # (NCCL announces its version on the first communicator) is sent to stderr instead
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import bwt_mtf_huffman_compressor_b200 as bz
    import workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.size
    data = W.synthetic_text(n, BASE_SEED + rank)
    ctx = bz.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    cap = bz.compress_bound(n)
    h_in = torch.from_numpy(data.copy()).pin_memory()
    h_file = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = h_in.cuda()
    d_file = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_back = torch.empty(n, dtype=torch.uint8, device="cuda")

    # ---- validity gate: byte-identical file + bit-exact round trip (every rank) -------------------
    flen = ctx.compress_ptr(d_in.data_ptr(), n, d_file.data_ptr(), cap, device=True)
    ctx.decompress_ptr(d_file.data_ptr(), flen, d_back.data_ptr(), n, device=True)
    torch.cuda.synchronize()
    if not torch.equal(d_back, d_in):
        raise SystemExit("bench.py: round trip is not bit exact")
    file_sha = hashlib.sha256(d_file[:flen].cpu().numpy().tobytes()).hexdigest()
    golden_ok = None
    if n == N_TEXT:
        # sha256 of the REAL reference's output for this rank's input (seed 0x5EED0064 + rank, ranks 0..7)
        G = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
        g = G["text"][str(n)] if rank == 0 else G.get("text_rank", {}).get(str(rank))
        if g is not None:
            golden_ok = file_sha == g["sha256"]
            if not golden_ok:
                raise SystemExit("bench.py: rank %d: compressed file differs from the reference golden" % rank)

    def step_device():
        fl = ctx.compress_ptr(d_in.data_ptr(), n, d_file.data_ptr(), cap, device=True)
        sc = ctx.stats()
        ctx.decompress_ptr(d_file.data_ptr(), fl, d_back.data_ptr(), n, device=True)
        sd = ctx.stats()
        return fl, sc, sd

    # e2e: EB blocks per step through the batch entry points (the reference's file loop, main.cpp:424-437):
    # pinned host buffers, 2 workers (streams) on this GPU, so that the H2D copy of one block and the
    # read-back of another overlap the kernels of a third; every copy is inside the timed region
    EB = args.e2e_blocks
    h_ins = [h_in] + [torch.from_numpy(data.copy()).pin_memory() for _ in range(EB - 1)]
    h_files = [h_file] + [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(EB - 1)]
    h_backs = [h_back] + [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(EB - 1)]

    def step_host():
        fls = bz.batch_ptrs(True, [t.data_ptr() for t in h_ins], [n] * EB, [t.data_ptr() for t in h_files], None, n_streams=args.e2e_streams)
        bz.batch_ptrs(False, [t.data_ptr() for t in h_files], fls, [t.data_ptr() for t in h_backs], [n] * EB, n_streams=args.e2e_streams)
        return fls[0]

    for _ in range(max(args.warmup, 3)):
        step_device()
    def measure():
        # ---- timed region: device-resident ---------------------------------------------------------------
        sampler = ClockSampler(local) if rank == 0 else None
        launches0 = ctx.stats().kernel_launches
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        agg = {"c_ms": 0.0, "d_ms": 0.0, "sort_ms": 0.0, "sort_bytes": 0, "passes": 0, "rounds": 0, "iters": 0,
               "c_bwt": 0.0, "c_mtf": 0.0, "c_huf": 0.0, "d_huf": 0.0, "d_mtf": 0.0, "d_bwt": 0.0, "walk_ms": 0.0,
               "walk_bytes": 0}
        e0.record(stream)
        for _ in range(args.steps):
            fl, sc, sd = step_device()
            agg["c_ms"] += sc.ms_total
            agg["d_ms"] += sd.ms_total
            agg["sort_ms"] += sc.ms_sort
            agg["sort_bytes"] += sc.sort_bytes
            agg["passes"] += sc.bwt_full_passes
            agg["rounds"] = sc.bwt_rounds
            agg["iters"] = sd.decode_sync_iters
            agg["c_bwt"] += sc.ms_bwt; agg["c_mtf"] += sc.ms_mtf; agg["c_huf"] += sc.ms_huffman
            agg["d_huf"] += sd.ms_huffman; agg["d_mtf"] += sd.ms_mtf; agg["d_bwt"] += sd.ms_bwt
            agg["walk_ms"] += sd.ms_walk; agg["walk_bytes"] += sd.walk_bytes
        e1.record(stream)
        barrier()
        dev_ms = e0.elapsed_time(e1)
        launches = ctx.stats().kernel_launches - launches0
        # ---- timed region: host buffers (e2e) ---------------------------------------------------------------
        for _ in range(2):
            step_host()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fl = step_host()
        f1.record(stream)
        barrier()
        host_wall_ms = (time.perf_counter() - t0) * 1e3
        host_ms = max(f0.elapsed_time(f1), host_wall_ms)
        clocks = sampler.stop() if sampler else None
        for hb in h_backs:
            if not np.array_equal(hb.numpy(), data):
                raise SystemExit("bench.py: host round trip is not bit exact")
        return dev_ms, host_ms, agg, launches, clocks, fl

    dev_ms, host_ms, agg, launches, clocks, fl = measure()
    # a run that saw a hardware / thermal slowdown is rejected and measured once more (all ranks together)
    bad = 0.0
    if clocks and any(r in clocks["reasons"] for r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")):
        bad = 1.0
    remeasured = False
    flag = torch.tensor([bad], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if float(flag[0]) > 0:
        remeasured = True
        dev_ms, host_ms, agg, launches, clocks, fl = measure()
    if clocks is not None:
        clocks["remeasured_after_throttle"] = remeasured

    cal = None if args.no_calgary else calgary_batch(bz, W, rank, world, dist)
    del d_in, d_file, d_back
    torch.cuda.empty_cache()
    blk = None if args.no_block1g else single_block(bz, W, ctx, rank, world, local, dist, args.block_size)

    t = torch.tensor([dev_ms, host_ms, 0.0 if golden_ok else 1.0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, host_ms = float(t[0]), float(t[1])
    if n == N_TEXT and world <= 8:
        golden_ok = float(t[2]) == 0.0                # every rank's file equals its reference golden
    total_bytes = float(n) * world * args.steps
    value = total_bytes / (dev_ms * 1e-3) / 1e6
    e2e = total_bytes * args.e2e_blocks / (host_ms * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = peaks()
        K = args.steps
        sort_gbs = agg["sort_bytes"] / (agg["sort_ms"] * 1e-3) / 1e9 if agg["sort_ms"] > 0 else 0.0
        walk_gbs = agg["walk_bytes"] / (agg["walk_ms"] * 1e-3) / 1e9 if agg["walk_ms"] > 0 else 0.0
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": round(dev_ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 symbols / u32 ranks / u64 sort keys (integer only)", "data": "synthetic",
            "config": {"workload": "text64m" if n == N_TEXT else "text%d" % n, "block_bytes": n, "blocks_per_step_per_gpu": 1,
                       "generator": "book1 words, splitmix64 seed 0x5EED0064+rank (SURVEY 8d C3)",
                       "l2": "working set (64 MiB input + 28 B/byte of sort state) larger than the 126 MB L2; no explicit flush",
                       "parallelism": "independent files, one per GPU per step, no data-path collective" if world > 1 else "1 GPU"},
            "compress_MBps": round(n * K / (agg["c_ms"] * 1e-3) / 1e6, 2),
            "decompress_MBps": round(n * K / (agg["d_ms"] * 1e-3) / 1e6, 2),
            "stage_ms_per_step": {"bwt": round(agg["c_bwt"] / K, 3), "mtf": round(agg["c_mtf"] / K, 3),
                                  "hist+huffman_encode": round(agg["c_huf"] / K, 3),
                                  "huffman_decode": round(agg["d_huf"] / K, 3), "imtf": round(agg["d_mtf"] / K, 3),
                                  "ibwt": round(agg["d_bwt"] / K, 3)},
            "compressed_bytes": int(fl), "file_sha256_matches_reference_golden": golden_ok,
            "bwt_rounds": agg["rounds"], "decode_sync_iters": agg["iters"],
            "e2e": {"value": round(e2e, 2), "unit": UNIT, "h2d_bytes_per_step": int(n + fl) * args.e2e_blocks,
                    "d2h_bytes_per_step": int(fl + n) * args.e2e_blocks, "ms_per_step": round(host_ms / K, 4),
                    "blocks_per_step": args.e2e_blocks, "ms_per_block": round(host_ms / K / args.e2e_blocks, 4),
                    "api": "bzap_compress_batch_gpus + bzap_decompress_batch_gpus, pinned host buffers, %d workers per GPU " % args.e2e_streams +
                           "(H2D / D2H of one block overlap the kernels of another)"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "onesweep_pass_kernel<u64 key, u32 payload> (BWT prefix-doubling sort pass over all N rotations)",
                         "bound": "hbm", "achieved": round(sort_gbs, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(sort_gbs / peak, 4), "traffic": None, "peak_source": peak_src,
                         "launches_per_step": agg["passes"] // K,
                         "avg_launch_ms": round(agg["sort_ms"] / max(agg["passes"], 1), 4),
                         "algorithmic_bytes_per_launch": int(agg["sort_bytes"] // max(agg["passes"], 1)),
                         "share_of_step": round(agg["sort_ms"] / dev_ms, 4)},
            "roofline_decompress": {
                "kernel": "ibwt_walk_len_kernel (inverse BWT: the one walk over T that also emits the output bytes)",
                "bound": "hbm", "achieved": round(walk_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(walk_gbs / peak, 4),
                "traffic": None, "peak_source": peak_src, "launches_per_step": 1,
                "avg_launch_ms": round(agg["walk_ms"] / K, 4), "algorithmic_bytes_per_launch": int(agg["walk_bytes"] // K),
                "share_of_step": round(agg["walk_ms"] / dev_ms, 4),
                "note": "N dependent random 4-byte reads of T: bound by random 32-byte-sector accesses, not by streaming bandwidth"},
            "clocks": clocks,
            "calgary_batch": cal,
            "single_block_1g": blk,
        }
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                tj = json.load(open(prof))
                # constants from the committed `ncu --set full` captures (profiles/), not measured in this run
                line["roofline"]["traffic"] = tj.get("onesweep_pass_u64_dram_bytes_per_launch")
                line["roofline"]["traffic_source"] = "ncu capture under profiles/ (constant, not re-measured per run)"
                line["roofline_decompress"]["traffic"] = tj.get("ibwt_walk_dram_bytes_per_launch")
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(data)
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    ctx.close()                      # collective when the ranks share a communicator: every rank gets here
    if dist is not None:
        dist.destroy_process_group()


def single_block(bz, W, ctx, rank, world, local, dist, n):
    """BASELINE config 5 (ii): ONE 1 GiB text block (seed 0x5EED1024) compressed by all N GPUs together --
    strong scaling.  N = 1: bzap_compress_device.  N > 1: bzap_compress_block_distributed (distributed
    prefix doubling, NCCL all-to-all over NVLink, csrc/dist_block.cu); the text is resident on every GPU
    and the file lands on rank 0.  Device-timed (CUDA events inside the library), max over ranks, mean
    of two runs after one warm-up; the file must equal the oracle golden (tests/golden/golden.json)."""
    import torch
    data = W.synthetic_text(n, 0x5EED1024)
    text = torch.from_numpy(data).cuda()
    del data
    cap = bz.compress_bound(n)
    out = torch.empty(cap if rank == 0 else 16, dtype=torch.uint8, device="cuda")
    if world > 1:
        box = [bz.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], world, rank)
    ms, rounds, ln = [], 0, 0
    for it in range(3):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        if world == 1:
            ln = ctx.compress_ptr(text.data_ptr(), n, out.data_ptr(), cap, device=True)
            st = ctx.stats()
            t, rounds = st.ms_total, st.bwt_rounds
        else:
            ln = ctx.compress_block_distributed(text.data_ptr(), n, out.data_ptr() if rank == 0 else 0, cap)
            st = ctx.dist_stats()
            t, rounds = st.ms_total, st.rounds
        if it > 0:
            ms.append(t)
    tt = torch.tensor(ms, dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    sec = float(tt.mean()) * 1e-3
    if rank != 0:
        return None
    sha = hashlib.sha256(out[:ln].cpu().numpy().tobytes()).hexdigest()
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["text"].get(str(n))
    return {"block_bytes": n, "n_gpus": world, "seconds": round(sec, 4), "MBps": round(n / sec / 1e6, 1), "rounds": int(rounds),
            "compressed_bytes": int(ln), "sha256": sha,
            "matches_golden": bool(g is not None and g.get("seed") == "0x5EED1024" and sha == g["sha256"]),
            "path": "bzap_compress_device" if world == 1 else "bzap_compress_block_distributed (NCCL, %d ranks)" % world,
            "scaling": "strong", "timing": "CUDA events inside the library, text resident in HBM on every rank, max over ranks"}


def calgary_batch(bz, W, rank, world, dist):
    """BASELINE config 2: the 14 Calgary files (3,141,622 B, one BWT block each) as a batch through
    the host-buffer ABI (bzap_compress_batch: 8 streams per GPU; files sharded over ranks when
    N > 1).  Latency bound on a B200; reported beside the headline, not part of `value`."""
    import torch
    from bwt_mtf_huffman_compressor_b200 import sharding
    cal = W.calgary()
    names = W.CALGARY_FILES
    sizes = [len(cal[n]) for n in names]
    mine = sharding.shard_files(sizes, world)[rank]
    datas = [np.frombuffer(cal[names[i]], dtype=np.uint8) for i in mine]
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["calgary"]
    ok = True
    blobs = []
    tcs, tds = [], []
    reps = 7
    for it in range(reps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        blobs = bz.compress_batch(datas, 8) if datas else []
        t1 = time.perf_counter()
        outs = bz.decompress_batch(blobs, 8) if blobs else []
        t2 = time.perf_counter()
        if it >= 2:
            tcs.append(t1 - t0)
            tds.append(t2 - t1)
        if it == 0:
            for i, b, o in zip(mine, blobs, outs):
                ok = ok and hashlib.sha256(b.tobytes()).hexdigest() == g[names[i]]["sha256"] and o.tobytes() == cal[names[i]]
    # median of the repetitions: the batch is launch bound (about 1500 driver calls in ~3 ms), so one
    # descheduled host thread on a shared box doubles a single repetition
    tc = sharding.max_over_ranks(float(np.median(tcs)), dist, "cuda")
    td = sharding.max_over_ranks(float(np.median(tds)), dist, "cuda")
    ok = sharding.max_over_ranks(0.0 if ok else 1.0, dist, "cuda") == 0.0
    total = float(sum(sizes))
    return {"files": len(names), "bytes": int(sum(sizes)), "streams_per_gpu": 8,
            "compress_MBps": round(total / tc / 1e6, 1), "decompress_MBps": round(total / td / 1e6, 1),
            "roundtrip_MBps": round(total / (tc + td) / 1e6, 1), "byte_identical_to_reference_goldens": bool(ok),
            "timing": "host wall clock incl. H2D/D2H, median of %d repetitions, max over ranks" % reps}


# ---------------------------------------------------------------------------------------------------
def ref_bins():
    d = os.path.join(ROOT, "oracle", "_ref")
    c, dd = os.path.join(d, "ref_compress"), os.path.join(d, "ref_decompress")
    return (c, dd) if os.path.exists(c) and os.path.exists(dd) else None


def time_reference_roundtrip(samples, workers):
    """Runs the reference's compress + decompress on each sample (one fresh process per file, the
    only mode whose bytes are defined -- SURVEY App. B), `workers` samples at a time.  Returns
    (seconds, kind)."""
    bins = ref_bins()
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i, s in enumerate(samples):
            p = os.path.join(tmp, "in%d" % i)
            np.ascontiguousarray(s).tofile(p)
            paths.append(p)
        if bins:
            def work(p):
                subprocess.run([bins[0], p, p + ".bz"], check=True, stdout=subprocess.DEVNULL)
                subprocess.run([bins[1], p + ".bz", p + ".out"], check=True, stdout=subprocess.DEVNULL)
            kind = "reference"
        else:
            import oracle_lib as O

            def work(p):
                d = np.fromfile(p, dtype=np.uint8)
                O.o_decompress(O.o_compress(d))
            kind = "port"
        t0 = time.perf_counter()
        if workers <= 1:
            for p in paths:
                work(p)
        else:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(workers) as ex:
                list(ex.map(work, paths))
        dt = time.perf_counter() - t0
        if bins:
            for p, s in zip(paths[:1], samples[:1]):
                if not np.array_equal(np.fromfile(p + ".out", dtype=np.uint8), s):
                    raise SystemExit("reference round trip failed")
    return dt, kind


def cpu_baseline(data):
    """The reference's own CPU path on this box's host cores, single thread (it has no threading):
    compress + decompress of the first 4, 8 and 16 MiB of the same block, one process each, run one
    after the other.  `value` is the measured 16 MiB figure; the reference's sort is super-linear, so
    the full 64 MiB block is slower still -- `full_block_extrapolated` fits t = a * n^b through the
    three points (bench.py --impl reference times the genuine 64 MiB block)."""
    pts = []
    kind = "reference"
    for mib in (4, 8, 16):
        sample = data[: mib << 20]
        dt, kind = time_reference_roundtrip([sample], 1)
        pts.append((sample.size, dt))
    ln = np.log(np.array(pts, dtype=np.float64))
    b, a = np.polyfit(ln[:, 0], ln[:, 1], 1)
    t_full = float(np.exp(a + b * np.log(float(data.size))))
    n16, t16 = pts[-1]
    return {"value": round(n16 / t16 / 1e6, 3), "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first 16 MiB of the text64m block, compress + decompress, one process, wall clock incl. file I/O",
            "points_MBps": {"%dMiB" % (n_ >> 20): round(n_ / t_ / 1e6, 3) for n_, t_ in pts},
            "full_block_extrapolated": {"MBps": round(data.size / t_full / 1e6, 3), "seconds": round(t_full, 1),
                                        "fit": "t = a * n^b, b = %.3f" % b}}


def run_reference(args):
    """The reference's CPU implementation on the SAME config as our arm: the genuine 64 MiB text
    block (seed 0x5EED0064) through the unmodified ref_compress + ref_decompress, one fresh process
    each (single thread: the reference has no threading and cannot split a block).  One round trip
    costs about 90 s, so it is timed ONCE whatever --steps says (steps_measured = 1, no warm-up).
    The all-cores figure (independent 2 MiB slices, one process per core) is kept as an extra key."""
    import workloads as W
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    n = args.size
    data = W.synthetic_text(n, BASE_SEED)
    dt, kind = time_reference_roundtrip([data], 1)
    value = n / dt / 1e6
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    per = 2 << 20
    slices = [W.synthetic_text(per, BASE_SEED + 1000 + i) for i in range(workers)]
    dts, _ = time_reference_roundtrip(slices, workers)
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "steps_measured": 1, "warmup": 0, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "text64m" if n == N_TEXT else "text%d" % n, "block_bytes": n, "blocks_per_step_per_gpu": 1,
                       "generator": "book1 words, splitmix64 seed 0x5EED0064 (SURVEY 8d C3)",
                       "note": "the full block, timed once: one reference round trip takes ~90 s"},
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": 1, "kind": kind,
                             "sample": "the whole %d-byte block, compress + decompress, one process each, wall clock incl. file I/O" % n},
            "all_cores_slices": {"value": round(per * workers / dts / 1e6, 3), "unit": UNIT, "cores": workers,
                                 "sample": "%d independent 2 MiB slices, one reference process per slice, all at once "
                                           "(not the benchmarked block: the reference cannot split one)" % workers},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=N_TEXT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-calgary", action="store_true", help="skip the Calgary batch (profiling runs)")
    ap.add_argument("--no-block1g", action="store_true", help="skip the single 1 GiB block (profiling runs)")
    ap.add_argument("--block-size", type=int, default=1 << 30)
    ap.add_argument("--e2e-blocks", type=int, default=8, help="blocks per e2e step (batch entry points)")
    ap.add_argument("--e2e-streams", type=int, default=4, help="workers (streams) per GPU for the e2e batch")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
