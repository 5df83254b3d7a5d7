#!/usr/bin/env python3
"""bench.py -- uncompressed-VCF GB/s of the genotype-column encode (and decode) hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is ONE pass of the encode path over one resident block of synthetic data lines
(BASELINE.json configs[1]: 1000 Genomes chr20-shaped, 2504 phased diploid samples; default
1.8 M lines ~= 18 GB per GPU, weak scaling: every rank encodes its own line shard, no collective).
`value` = uncompressed bytes of all ranks / max-over-ranks device time (CUDA events on the launch
stream, inputs already in HBM).  `e2e` = the same metric through the host-pointer C-ABI call
(pinned host buffers, H2D + kernels + D2H inside the timed region).  `roofline` is for the
dominant kernel, algorithmic bytes = N_in + N_out, peak = MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference` time the reference's own CPU compressor (oracle/_ref, built
from the unmodified sources) on the box's host cores on a bounded sample of the same workload.
Decode numbers ride along in "decode".  One JSON line on stdout (rank 0).
"""
import argparse
import importlib
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "uncompressed-VCF GB/s encode"
UNIT = "GB/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (pynvml; the recipe's nvidia-smi line, in-process)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = str(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self.nv:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.t:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no samples"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# the reference's CPU compressor on a bounded sample (oracle/_ref binary, else the C port)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(sample: bytes, header: bytes, n_proc: int, steps: int, warmup: int, decode: bool = False):
    """Shards `sample` (whole lines) over n_proc processes of the reference CLI; returns
    (GB/s of uncompressed VCF over all processes per step list, kind, decode GB/s list)."""
    import oraclelib as O
    use_bin = O.have_ref_binary()
    if use_bin:
        try:
            use_bin = subprocess.run([O.REF_BIN, "nonsense-verb"], capture_output=True, timeout=20).returncode in (0, 1)
        except Exception:  # noqa: BLE001
            use_bin = False
    kind = "reference" if use_bin else "port"
    wd = tempfile.mkdtemp(prefix="vcfc_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        lines = sample.split(b"\n")[:-1]
        per = max(1, (len(lines) + n_proc - 1) // n_proc)
        shards = []
        for p in range(n_proc):
            part = lines[p * per:(p + 1) * per]
            if not part:
                break
            fp = os.path.join(wd, f"s{p}.vcf")
            with open(fp, "wb") as f:
                f.write(header + b"\n".join(part) + b"\n")
            shards.append(fp)
        total = sum(len(x) + 1 for x in lines)

        def one_pass(verb):
            t0 = time.perf_counter()
            if use_bin:
                ps = []
                for fp in shards:
                    src, dst = (fp, fp + "c") if verb == "compress" else (fp + "c", fp + ".rt")
                    ps.append(subprocess.Popen([O.REF_BIN, verb, src, dst], stdout=subprocess.DEVNULL))
                for p in ps:
                    if p.wait() != 0:
                        raise RuntimeError("reference binary failed")
            else:
                fn = O.lib().vcfc_oracle_compress_file if verb == "compress" else O.lib().vcfc_oracle_decompress_file
                ts = []
                for fp in shards:
                    src, dst = (fp, fp + "c") if verb == "compress" else (fp + "c", fp + ".rt")
                    ts.append(threading.Thread(target=fn, args=(src.encode(), dst.encode())))   # ctypes drops the GIL
                [t.start() for t in ts]
                [t.join() for t in ts]
            return time.perf_counter() - t0

        enc, dec = [], []
        for i in range(warmup + steps):
            dt = one_pass("compress")
            if i >= warmup:
                enc.append(total / dt / 1e9)
        if decode:
            for i in range(1 + max(1, steps // 2)):
                dt = one_pass("decompress")
                if i >= 1:
                    dec.append(total / dt / 1e9)
        return enc, kind, dec, len(shards), total
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def host_sample(kind, n_lines, n_samples, seed):
    import torch
    import vcfsynth
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    b, _ = vcfsynth.generate(kind, n_lines, n_samples, seed=seed, device=dev)
    return bytes(b.cpu().numpy())



# ------------------------------------------------------------------------------------------------
# file -> file: the verbs a user of the reference runs (compress IN OUT / decompress IN OUT), on tmpfs
# ------------------------------------------------------------------------------------------------
def _best(fn, n=3):
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


def file_bench(pkg, codec, d_in, starts, n_lines, header: bytes, samples: int):
    """e2e_file: vcfc_compress_file / vcfc_decompress_file (pinned chunk ring, reader / worker / writer threads) and the
    `vcfc` CLI on a bounded prefix of the workload written to tmpfs, beside the reference CLI on a smaller prefix."""
    import shutil
    import oraclelib as O
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    wd = tempfile.mkdtemp(prefix="vcfc_file_", dir=shm)
    res = {"tmpfs": shm}
    try:
        free = shutil.disk_usage(shm).free
        target = min(float(starts[n_lines]), 4e9, free * 0.25)
        k = int((starts[:n_lines + 1] <= target).sum().item()) - 1
        k = max(1, k)
        nbytes = int(starts[k])
        ip, op, rp = (os.path.join(wd, x) for x in ("in.vcf", "out.vcfc", "rt.vcf"))
        with open(ip, "wb") as f:
            f.write(header)
            step = 256 << 20
            for a in range(0, nbytes, step):
                f.write(d_in[a:min(nbytes, a + step)].cpu().numpy().tobytes())
        fbytes = os.path.getsize(ip)
        res["file_bytes"] = fbytes
        res["lines"] = k
        rc = codec.compress(ip, op)                      # warm-up: pinned pool, device buffers
        if rc != 0:
            raise RuntimeError(f"compress rc {rc}")
        tc = _best(lambda: codec.compress(ip, op), 3)
        codec.decompress2_fd(op, rp)
        td = _best(lambda: codec.decompress2_fd(op, rp), 2)
        same = subprocess.run(["cmp", "-s", ip, rp]).returncode == 0
        res["lib"] = {"api": "vcfc_compress_file / vcfc_decompress_file (context warm)", "compress_gbs": fbytes / tc / 1e9,
                      "decompress_gbs": fbytes / td / 1e9, "round_trip_identical": bool(same),
                      "vcfc_bytes": os.path.getsize(op)}
        if os.path.exists(pkg.CLI_PATH):
            op2, rp2 = os.path.join(wd, "cli.vcfc"), os.path.join(wd, "cli.vcf")
            t = time.perf_counter()
            r1 = subprocess.run([pkg.CLI_PATH, "compress", ip, op2], capture_output=True)
            tcc = time.perf_counter() - t
            t = time.perf_counter()
            r2 = subprocess.run([pkg.CLI_PATH, "decompress", op2, rp2], capture_output=True)
            tdc = time.perf_counter() - t
            ok = r1.returncode == 0 and r2.returncode == 0 and subprocess.run(["cmp", "-s", op, op2]).returncode == 0
            res["cli"] = {"cmd": "vcfc compress / decompress (one process each: includes CUDA context creation and pinned allocation)",
                          "compress_gbs": fbytes / tcc / 1e9, "decompress_gbs": fbytes / tdc / 1e9, "compress_s": tcc,
                          "decompress_s": tdc, "same_bytes_as_lib": bool(ok)}
            for x in (op2, rp2):
                if os.path.exists(x):
                    os.remove(x)
        if O.have_ref_binary():
            kr = max(1, int((starts[:n_lines + 1] <= 150e6).sum().item()) - 1)
            nb = int(starts[kr])
            ipr, opr, rpr = (os.path.join(wd, x) for x in ("ref.vcf", "ref.vcfc", "ref.rt"))
            with open(ipr, "wb") as f:
                f.write(header)
                f.write(d_in[:nb].cpu().numpy().tobytes())
            t = time.perf_counter()
            a = subprocess.run([O.REF_BIN, "compress", ipr, opr], capture_output=True).returncode
            trc = time.perf_counter() - t
            t = time.perf_counter()
            b = subprocess.run([O.REF_BIN, "decompress", opr, rpr], capture_output=True).returncode
            trd = time.perf_counter() - t
            if a == 0 and b == 0:
                fb = os.path.getsize(ipr)
                res["reference_cli"] = {"cmd": "oracle/_ref/main_release compress / decompress (one process, one core)",
                                        "sample_bytes": fb, "compress_gbs": fb / trc / 1e9, "decompress_gbs": fb / trd / 1e9}
    except Exception as e:  # noqa: BLE001
        res["error"] = str(e)[:200]
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    return res


def index_query_bench(pkg, codec, dev):
    """The next rows of the scope table on the configuration BASELINE.md quotes them on (config 1 shape: 10k x 2504,
    random_vcf.py distribution): binned index build (separate pass and fused with compress), indexed range query, linear
    query -- library calls with a warm context, the `vcfc` CLI, and the reference binary on the same files."""
    import shutil
    import oraclelib as O
    import vcfsynth
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    wd = tempfile.mkdtemp(prefix="vcfc_idx_", dir=shm)
    res = {"workload": "random_vcf.py distribution, 10000 lines x 2504 samples (config 1 shape), bin size 150, query 1:12000-14000"}
    try:
        d, _ = vcfsynth.generate("random", 10000, 2504, seed=5, device=dev)
        vcf = vcfsynth.header(2504) + d.cpu().numpy().tobytes()
        ip, op = os.path.join(wd, "c1.vcf"), os.path.join(wd, "c1.vcfc")
        open(ip, "wb").write(vcf)
        null = os.open(os.devnull, os.O_WRONLY)
        codec.compress(ip, op)
        t_c = _best(lambda: codec.compress(ip, op))
        t_ci = _best(lambda: pkg.Codec.compress_index_multi([codec], ip, op, op + ".vcfci", 150))
        fused = open(op + ".vcfci", "rb").read()
        t_i = _best(lambda: codec.create_binned_index(op, op + ".vcfci", 150))
        same = fused == open(op + ".vcfci", "rb").read()
        t_qi = _best(lambda: codec.query_binned_index(op, "1:12000-14000", null))
        t_q = _best(lambda: codec.query(op, "1:12000-14000", null))
        res["lib"] = {"compress_s": t_c, "compress_plus_index_fused_s": t_ci, "index_fused_extra_s": max(0.0, t_ci - t_c),
                      "create_binned_index_s": t_i, "fused_index_same_bytes": bool(same),
                      "query_binned_index_s": t_qi, "query_s": t_q, "note": "library calls, context warm, files on tmpfs"}
        if os.path.exists(pkg.CLI_PATH):
            def cli(*a):
                return _best(lambda: subprocess.run([pkg.CLI_PATH, *a], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL), 2)
            res["cli"] = {"create_binned_index_s": cli("create-binned-index", "150", op), "query_binned_index_s": cli("query-binned-index", op, "1:12000-14000"),
                          "query_s": cli("query", op, "1:12000-14000"), "note": "one process per verb: includes CUDA context creation (~0.3-0.5 s)"}
        if O.have_ref_binary():
            def ref(*a):
                return _best(lambda: subprocess.run([O.REF_BIN, *a], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL), 2)
            shutil.copy(op, op + ".ref")
            res["reference_cli"] = {"create_binned_index_s": ref("create-binned-index", "150", op + ".ref"),
                                    "query_binned_index_s": ref("query-binned-index", op + ".ref", "1:12000-14000"),
                                    "query_s": ref("query", op + ".ref", "1:12000-14000"),
                                    "index_same_bytes": open(op + ".ref.vcfci", "rb").read() == fused if os.path.exists(op + ".ref.vcfci") else None}
        os.close(null)
    except Exception as e:  # noqa: BLE001
        res["error"] = str(e)[:200]
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    return res

# ------------------------------------------------------------------------------------------------
_RESULT_FD = None


def emit(obj):
    """The ONE JSON line, on the process's real stdout (see main: fd 1 points at stderr meanwhile)."""
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def main():
    # Libraries write to stdout behind Python's back (NCCL's version banner, also when it is switched on in a conf file): fd 1
    # points at stderr while the bench runs, and only the result line goes to the real stdout.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lines", type=int, default=int(os.environ.get("VCFC_BENCH_LINES", 1_800_000)))
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--kind", default="kg", choices=["kg", "random"])
    ap.add_argument("--seed", type=int, default=20)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--cpu-lines-per-proc", type=int, default=3000)
    args = ap.parse_args()
    if args.warmup < 3:
        log("note: timing rules ask for >= 3 warm-up steps")

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    workload = (f"{'1000G-chr20-shaped' if args.kind == 'kg' else 'random_vcf.py-distribution'} synthetic VCF, "
                f"{args.lines} lines x {args.samples} phased diploid samples per GPU")
    config = {"workload": workload, "lines_per_gpu": args.lines, "samples": args.samples, "generator": f"tests/vcfsynth.py:{args.kind}",
              "seed": args.seed, "sharding": f"line blocks, {world} rank(s), no collective",
              "l2": "input per step is far larger than the 126 MB L2 (no flush needed)"}

    # ---------------- reference arm: CPU only, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        n_proc = max(1, os.cpu_count() or 1)
        n_lines = min(args.lines * world, n_proc * args.cpu_lines_per_proc)
        sample = host_sample(args.kind, n_lines, args.samples, args.seed)
        import vcfsynth
        enc, kind, _, used, total = cpu_reference_run(sample, vcfsynth.header(args.samples), n_proc, args.steps, args.warmup)
        v = statistics.median(enc)
        desc = (f"first {n_lines} lines ({total / 1e6:.0f} MB) of the workload, sharded by lines over {used} processes of the "
                f"{'unmodified reference CLI (oracle/_ref/main_release compress)' if kind == 'reference' else 'oracle port'}, files in /dev/shm")
        emit(({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / v / 1e6,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": used, "kind": kind, "sample": desc},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # ---------------- B200 arm ----------------
    import torch
    import torch.distributed as dist
    import oraclelib as O
    import vcfsynth
    pkg = importlib.import_module("vcf-compression_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # the version banner goes to stdout, where one JSON line is expected
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    codec = pkg.Codec(local_rank)
    t0 = time.time()
    cap_in = args.lines * vcfsynth.max_line_bytes(args.kind, args.samples)
    d_buf = torch.empty(cap_in, dtype=torch.uint8, device=dev)
    d_in, lens = vcfsynth.generate(args.kind, args.lines, args.samples, seed=args.seed, first_line=rank * args.lines,
                                   device=dev, out=d_buf)
    n_in = d_in.numel()
    torch.cuda.synchronize()
    log(f"[rank {rank}] generated {n_in / 1e9:.2f} GB in {time.time() - t0:.1f}s")
    out_cap = int(n_in * (0.30 if args.kind == "kg" else 0.45)) + (1 << 20)
    d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(8, dtype=torch.int64, device=dev)
    d_offs = torch.empty(args.lines + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- parity before timing: oracle on line windows, full-size round trip on the device ----
    codec.encode_dev(d_in.data_ptr(), n_in, d_out.data_ptr(), out_cap, d_res.data_ptr(), stream, d_offs.data_ptr(), args.lines + 1)
    r = codec.fetch_result(d_res.data_ptr(), stream)
    if r.status != 0 or r.n_lines != args.lines:
        raise SystemExit(f"encode failed: status {r.status} ({pkg.strerror(r.status)}), lines {r.n_lines}")
    n_out = int(r.out_len)
    enc_path = codec.last_path
    starts = torch.zeros(args.lines + 1, dtype=torch.int64, device=dev)
    starts[1:] = torch.cumsum(lens, 0)
    offs = d_offs[:args.lines]
    win = min(args.lines, 1500)
    checked = 0
    for lo in sorted({0, max(0, args.lines // 2 - win // 2), args.lines - win}):
        a, b = int(starts[lo]), int(starts[lo + win])
        oa = int(offs[lo])
        ob = int(offs[lo + win]) if lo + win < args.lines else n_out
        orc, oout, onl, _ = O.compress_block(bytes(d_in[a:b].cpu().numpy()))
        if orc != 0 or onl != win or oout != bytes(d_out[oa:ob].cpu().numpy()):
            raise SystemExit(f"PARITY FAILURE vs oracle on lines [{lo}, {lo + win})")
        checked += win
    dec_path = None
    if not args.no_decode:
        d_txt = torch.empty(n_in + 64, dtype=torch.uint8, device=dev)
        codec.decode_dev(d_out.data_ptr(), n_out, args.samples, d_txt.data_ptr(), d_txt.numel(), d_res.data_ptr(), stream)
        r2 = codec.fetch_result(d_res.data_ptr(), stream)
        if r2.status != 0 or r2.out_len != n_in or not torch.equal(d_txt[:n_in], d_in):
            raise SystemExit(f"ROUND TRIP FAILURE: status {r2.status}, {r2.out_len} vs {n_in} bytes")
        dec_path = codec.last_path
    log(f"[rank {rank}] parity ok: {checked} lines vs oracle, full round trip on device; ratio {n_in / n_out:.2f}; "
        f"encode path {enc_path}, decode path {dec_path}")

    def timed(fn, which):
        """W warm-up + K timed steps; returns (ms per step max over ranks, kernel ms per step, launches, clocks)."""
        for _ in range(args.warmup):
            fn()
        codec.set_timing(True)
        barrier()
        l0 = codec.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = []
        with ClockSampler(local_rank) as cs:
            e0.record()
            for _ in range(args.steps):
                fn()
                # reading the kernel-only time costs a sync AFTER the step's own work has been queued;
                # the dev API has already synchronised the stream to read the device status by then
                kms.append(codec.last_kernel_ms(which))
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1) / args.steps
        codec.set_timing(False)
        return max_over_ranks(ms), kms, codec.launches - l0, cs.summary()

    enc_ms, enc_kms, enc_launches, clocks = timed(
        lambda: codec.encode_dev(d_in.data_ptr(), n_in, d_out.data_ptr(), out_cap, d_res.data_ptr(), stream), 0)
    tot_in = sum_over_ranks(float(n_in))
    tot_out = sum_over_ranks(float(n_out))
    value = tot_in / (enc_ms * 1e-3) / 1e9

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"

    # DRAM traffic per launch: measured once per round with `ncu --set full` on a 100k-line run of this same
    # workload (profiles/traffic_r1.json: dram__bytes_read.sum + dram__bytes_write.sum next to that run's
    # algorithmic bytes) and scaled by the algorithmic bytes of this launch; null when the file is absent.
    # (profiles/traffic_r2.json is this round's capture.)
    traffic_ref = {}
    for tf in ("traffic_r2.json", "traffic_r1.json"):          # the latest round's capture
        try:
            traffic_ref = json.load(open(os.path.join(ROOT, "profiles", tf)))
            break
        except Exception:  # noqa: BLE001
            pass

    def roofline(kms, alg_bytes, step_ms, name, which):
        k = [x for x in kms if x and x > 0]
        kern_ms = statistics.mean(k) if k else step_ms
        ach = alg_bytes / (kern_ms * 1e-3) / 1e9
        tr = traffic_ref.get(which)
        traffic = alg_bytes * tr["dram_bytes"] / tr["algorithmic_bytes"] if tr and name.startswith("k_") else None
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": (tr or {}).get("source") if traffic else None,
                "kernel": name, "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / step_ms if step_ms else None,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}

    enc_kernel = "k_encode_stream (single pass over the input)" if enc_path == pkg.PATH_FAST else "generic line-serial kernels (whole step)"
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": enc_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
           "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": enc_launches,
           "compression_ratio": tot_in / tot_out, "path": {"encode": enc_path, "decode": dec_path},
           "parity": f"{checked} lines byte-identical to the oracle; full-size device round trip identical",
           "roofline": roofline(enc_kms, n_in + n_out, enc_ms, enc_kernel, "encode")}

    if not args.no_decode:
        dec_ms, dec_kms, dec_launches, dclocks = timed(
            lambda: codec.decode_dev(d_out.data_ptr(), n_out, args.samples, d_txt.data_ptr(), d_txt.numel(), d_res.data_ptr(), stream), 2)
        dec_kernel = "k_dec_expand_grid (fill and patch; k_dec_expand when sample text is off the 4-byte grid)" if dec_path == pkg.PATH_FAST else "generic line-serial kernels (whole step)"
        out["decode"] = {"metric": "uncompressed-VCF GB/s decode", "value": tot_in / (dec_ms * 1e-3) / 1e9, "unit": UNIT,
                         "ms_per_step": dec_ms, "gpu_launches": dec_launches, "clocks": dclocks,
                         "roofline": roofline(dec_kms, n_in + n_out, dec_ms, dec_kernel, "decode")}
        del d_txt

    # ---- end to end through the host-pointer C ABI: pinned host buffers, H2D + kernels + D2H timed ----
    if not args.no_e2e:
        import psutil
        avail = psutil.virtual_memory().available
        e2e_in = n_in
        if n_in * 1.6 * max(1, min(world, 8)) > avail * 0.6:
            frac = (avail * 0.6 / max(1, min(world, 8))) / (n_in * 1.6)
            k = max(1, int(args.lines * frac))
            e2e_in = int(starts[k])
        try:
            h_in = torch.empty(e2e_in, dtype=torch.uint8, pin_memory=True)
            h_in.copy_(d_in[:e2e_in])
            e2e_cap = int(e2e_in * (0.30 if args.kind == "kg" else 0.45)) + (1 << 20)
            h_out = torch.empty(e2e_cap, dtype=torch.uint8, pin_memory=True)
            torch.cuda.synchronize()
            res = None

            def e2e_step():
                nonlocal res
                res = codec.encode_host_ptr(h_in.data_ptr(), e2e_in, h_out.data_ptr(), e2e_cap)
                if res[0] != 0:
                    raise SystemExit(f"e2e encode failed: {res}")

            for _ in range(args.warmup):
                e2e_step()
            barrier()
            t1 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            dt = (time.perf_counter() - t1) / args.steps
            barrier()
            dt = max_over_ranks(dt)
            e2e_out = res[1]
            ok = bytes(h_out[:min(e2e_out, 1 << 20)].numpy()) == bytes(d_out[:min(e2e_out, 1 << 20)].cpu().numpy())
            out["e2e"] = {"value": sum_over_ranks(float(e2e_in)) / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": e2e_in,
                          "d2h_bytes_per_step": e2e_out, "ms_per_step": dt * 1e3,
                          "api": "vcfc_encode_block (host pointers, pinned; chunked H2D/kernels/D2H on two streams)",
                          "bytes_match_device_run": bool(ok),
                          "workload_note": "full workload" if e2e_in == n_in else f"first {e2e_in} bytes (host RAM bound)"}
            # what plain copies of the same bytes cost on this box with all ranks copying at once (H2D of the text and D2H
            # of the .vcfc on two streams, pinned memory): the ceiling the end-to-end number is to be read against
            try:
                d_c = torch.empty(e2e_out + 64, dtype=torch.uint8, device=dev)
                h_c = torch.empty(e2e_out + 64, dtype=torch.uint8, pin_memory=True)
                s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()

                def copies():
                    with torch.cuda.stream(s_a):
                        d_buf[:e2e_in].copy_(h_in, non_blocking=True)
                    with torch.cuda.stream(s_b):
                        h_c[:e2e_out].copy_(d_c[:e2e_out], non_blocking=True)
                    s_a.synchronize()
                    s_b.synchronize()

                copies()
                barrier()
                t1 = time.perf_counter()
                for _ in range(3):
                    copies()
                ct = max_over_ranks((time.perf_counter() - t1) / 3)
                barrier()
                ceil_v = sum_over_ranks(float(e2e_in)) / ct / 1e9
                out["e2e"]["copy_ceiling"] = {"value": ceil_v, "unit": UNIT, "what": "pinned H2D of the input + D2H of the output bytes only, "
                                              "all ranks at once, two streams", "frac": out["e2e"]["value"] / ceil_v}
                del d_c, h_c
            except Exception as e:  # noqa: BLE001
                out["e2e"]["copy_ceiling"] = {"error": str(e)[:120]}
            if not args.no_decode:
                h_txt = torch.empty(e2e_in + 64, dtype=torch.uint8, pin_memory=True)
                dres = None

                def e2e_dec():
                    nonlocal dres
                    dres = codec.decode_host_ptr(h_out.data_ptr(), e2e_out, args.samples, h_txt.data_ptr(), h_txt.numel())
                    if dres[0] != 0:
                        raise SystemExit(f"e2e decode failed: {dres}")

                for _ in range(min(args.warmup, 2)):
                    e2e_dec()
                barrier()
                t1 = time.perf_counter()
                n_dec = max(1, args.steps // 2)
                for _ in range(n_dec):
                    e2e_dec()
                ddt = max_over_ranks((time.perf_counter() - t1) / n_dec)
                out["decode"]["e2e"] = {"value": sum_over_ranks(float(e2e_in)) / ddt / 1e9, "unit": UNIT,
                                        "h2d_bytes_per_step": e2e_out, "d2h_bytes_per_step": e2e_in, "ms_per_step": ddt * 1e3,
                                        "round_trip_ok": bool(torch.equal(h_txt[:e2e_in], h_in))}
                # the same bytes as plain copies in the decode direction (D2H of the text, H2D of the .vcfc)
                try:
                    s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()
                    d_c = torch.empty(e2e_out + 64, dtype=torch.uint8, device=dev)

                    def copies_dec():
                        with torch.cuda.stream(s_a):
                            h_txt[:e2e_in].copy_(d_buf[:e2e_in], non_blocking=True)
                        with torch.cuda.stream(s_b):
                            d_c[:e2e_out].copy_(h_out[:e2e_out], non_blocking=True)
                        s_a.synchronize()
                        s_b.synchronize()

                    copies_dec()
                    barrier()
                    t1 = time.perf_counter()
                    for _ in range(3):
                        copies_dec()
                    ct = max_over_ranks((time.perf_counter() - t1) / 3)
                    barrier()
                    ceil_d = sum_over_ranks(float(e2e_in)) / ct / 1e9
                    out["decode"]["e2e"]["copy_ceiling"] = {"value": ceil_d, "unit": UNIT, "what": "pinned D2H of the text + H2D of the .vcfc bytes "
                                                            "only, all ranks at once, two streams", "frac": out["decode"]["e2e"]["value"] / ceil_d}
                    del d_c
                except Exception as e:  # noqa: BLE001
                    out["decode"]["e2e"]["copy_ceiling"] = {"error": str(e)[:120]}
                del h_txt
            del h_in, h_out
        except RuntimeError as e:
            out["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(e)[:200]}

    # ---- file -> file and the index / query rows: N = 1 only ----
    if world == 1 and not args.no_e2e:
        out["e2e_file"] = file_bench(pkg, codec, d_in, starts, args.lines, vcfsynth.header(args.samples), args.samples)
        if "lib" in out["e2e_file"] and out.get("e2e", {}).get("value"):
            out["e2e_file"]["lib_vs_block_e2e"] = out["e2e_file"]["lib"]["compress_gbs"] / out["e2e"]["value"]
        out["index_query"] = index_query_bench(pkg, codec, dev)

    # ---- CPU baseline beside it: rank 0, N = 1 only, bounded sample ----
    if world == 1 and not args.no_cpu:
        n_proc = max(1, os.cpu_count() or 1)
        n_lines = min(args.lines, n_proc * args.cpu_lines_per_proc)
        sample = bytes(d_in[:int(starts[n_lines])].cpu().numpy())
        enc, kind, dec, used, total = cpu_reference_run(sample, vcfsynth.header(args.samples), n_proc, 2, 1, decode=not args.no_decode)
        one = sample[:int(starts[min(n_lines, 1500)])]
        enc1, _, _, _, _ = cpu_reference_run(one, vcfsynth.header(args.samples), 1, 1, 0)
        out["cpu_baseline"] = {"value": statistics.median(enc), "unit": UNIT, "cores": used, "kind": kind,
                               "sample": f"first {n_lines} lines ({total / 1e6:.0f} MB) of the workload, line-sharded over {used} "
                                         f"processes of the reference CLI, /dev/shm files; one process alone: {enc1[0]:.4f} GB/s",
                               "one_core_value": enc1[0],
                               "decode_value": statistics.median(dec) if dec else None}
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


if __name__ == "__main__":
    main()
