#!/usr/bin/env python
"""Headline benchmark: exact top-10 cosine search over a device-resident corpus.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)

Workload (BASELINE.json metric / configs[2]): synthetic unit-norm 10M x 768 corpus in a bf16
store (bf16 rows for the scan + the un-rounded fp32 rows for the exact re-ranking, DESIGN.md
3.4), cosine, k = 10, one query batch per step.  The corpus is FIXED at 10M rows and row-sharded
over the N ranks (strong scaling); the per-step exchange of B x k candidate keys is fused into
the scan kernel (peer memory over NVLink) or an NCCL all-gather + merge kernel.

One JSON line on rank 0 (keys per the driver contract):
  value     QPS with the query batch already in HBM (device-timed, CUDA events, max over
            ranks); back-to-back launches overlap by programmatic dependent launch (value_note)
  e2e       the same metric through the public host-buffer API with 3 queries in flight (submit /
            collect: a server with concurrent requests): every step's queries go pinned host ->
            device and its B x k result comes back to the host inside the timed region; the
            one-blocking-call-per-step numbers sit beside it (sequential_value, p50_ms)
  roofline  scan kernel: algorithmic bytes (rows x row_bytes) / its CUDA-event time vs
            MEASURED_PEAKS.json hbm_gbs
  verified  the engine's answers for this very corpus against a chunked fp32 brute force over
            ALL shards (per-shard lists all-gathered and merged on the host by (distance, row));
            recall_bf16_vs_fp32 = recall@k of the bf16 store against exact fp32 search on the
            un-rounded inputs, >= 1000 queries
  regimes   the other BASELINE configs on the same box: B = 1024 on the headline corpus, config 2
            (1M x 384 fp32, B = 1 / 32 / 1024, with `f32_tensor`: the bf16 shadow in force and how many
            queries had to be re-run exactly; plus 1M x 768 fp32, labelled as beyond BASELINE), config 4
            (10M x 384, `where` at 1 / 10 / 50 % with 5 %
            tombstones), config 5's per-GPU shard (25M x 384, B = 1024, top-100, l2); each with its
            own roofline
  cpu_baseline  the oracle's BLAS exact search on a bounded sample (N = 1 only)
"""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm must use every host core: torch.distributed.run exports OMP_NUM_THREADS=1 to its
    # workers, which numpy's BLAS would obey (set before numpy is imported)
    _n = str(os.cpu_count() or 1)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_v] = _n

import argparse  # noqa: E402
import json  # noqa: E402
import statistics  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "QPS, exact top-10 cosine, 10M x 768"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--space", default="cosine", choices=["cosine", "l2", "ip"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--regime", default="auto", choices=["auto", "stream", "tensor"])
    ap.add_argument("--rerank", type=int, default=1, help="bf16 stores: keep the fp32 re-ranking plane (default 1)")
    ap.add_argument("--cpu-sample-rows", type=int, default=500_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the global verification / recall pass")
    ap.add_argument("--verify-queries", type=int, default=1024)
    ap.add_argument("--hnsw-baseline", action="store_true", default=True,
                    help="also build the CPU HNSW restatement (Chroma defaults) on a small sample and report "
                         "its recall / QPS under cpu_baseline.hnsw_restatement (~1.1 ms per inserted vector)")
    ap.add_argument("--no-hnsw-baseline", dest="hnsw_baseline", action="store_false")
    ap.add_argument("--hnsw-sample-rows", type=int, default=10_000)
    ap.add_argument("--selectivity", type=float, default=0.0,
                    help="config 4: apply a `where` bitmap passing this fraction of rows (0 = no filter)")
    ap.add_argument("--tombstones", type=float, default=0.0, help="config 4: delete this fraction of rows first")
    ap.add_argument("--extra-batches", default="1024",
                    help="comma list of further batch sizes measured device-resident on the headline corpus")
    ap.add_argument("--inflight", type=int, default=3, help="host-buffer queries kept in flight in the e2e pass (1-4)")
    ap.add_argument("--configs", default="2,4,5",
                    help="comma list of the other BASELINE configs to measure under 'regimes' ('' = none)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload_name(args):
    return (f"synthetic unit-norm {args.rows}x{args.dim} {args.dtype} corpus, exact top-{args.k} {args.space}, "
            f"query batch {args.batch}")


def config_dict(args):
    """Identical in both arms (the driver compares them)."""
    return {"workload": workload_name(args), "rows": args.rows, "dim": args.dim, "corpus_dtype": args.dtype,
            "batch": args.batch, "k": args.k, "space": args.space,
            "queries": "unit-norm Gaussian, 10 % planted next to a corpus row (sigma 0.05)",
            "l2_flush": "inputs larger than L2 (corpus bytes >> 126 MB)"}


def gaussian_queries(n_batches, B, dim, seed=4321):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((n_batches, B, dim), dtype=np.float32)
    q /= np.linalg.norm(q, axis=2, keepdims=True)
    return q


def plant(q, rows, sigma=0.05, seed=99):
    """Every 10th query becomes a corpus row + N(0, sigma^2 / dim) noise, re-normalised."""
    flat = q.reshape(-1, q.shape[-1])
    rng = np.random.default_rng(seed)
    idx = np.arange(0, flat.shape[0], 10)[: rows.shape[0]]
    noisy = rows[: idx.shape[0]] + sigma / np.sqrt(flat.shape[1]) * rng.standard_normal((idx.shape[0], flat.shape[1]), dtype=np.float32)
    flat[idx] = noisy / np.linalg.norm(noisy, axis=1, keepdims=True)
    return q


# ------------------------------------------------------------------------------------------
# CPU arms
# ------------------------------------------------------------------------------------------
def _fill_normal_parallel(x, seed, threads):
    """x[:] = row-normalised N(0, 1), generated by `threads` independent generators (numpy releases the GIL)."""
    import concurrent.futures
    n = x.shape[0]
    step = (n + threads - 1) // threads

    def work(t):
        lo, hi = t * step, min(n, (t + 1) * step)
        if lo >= hi:
            return
        rng = np.random.default_rng(seed + t)
        for s in range(lo, hi, 65536):
            e = min(hi, s + 65536)
            blk = rng.standard_normal((e - s, x.shape[1]), dtype=np.float32)
            blk /= np.linalg.norm(blk, axis=1, keepdims=True)
            x[s:e] = blk
    with concurrent.futures.ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(threads)))


def cpu_exact_qps(args, seconds_budget=25.0, sample_rows=None):
    """The oracle's BLAS exact search (numpy fp32 Q @ X.T + argpartition) on a bounded sample of the same
    workload, all host cores.  Returns (qps scaled to the full corpus, seconds per step, description, cores)."""
    from oracle.exact_search import fast_topk_f32, prepare_corpus
    n = min(sample_rows or args.cpu_sample_rows, args.rows)
    cores = os.cpu_count() or 1
    x = np.empty((n, args.dim), dtype=np.float32)
    _fill_normal_parallel(x, 1234, min(cores, 16))
    x = prepare_corpus(args.space, x, "f32")
    q = prepare_corpus(args.space, gaussian_queries(1, args.batch, args.dim)[0], "f32")
    fast_topk_f32(args.space, q, x[: min(n, 65536)], args.k)          # warm BLAS threads
    t0 = time.perf_counter()
    reps = 0
    while True:
        fast_topk_f32(args.space, q, x, args.k)
        reps += 1
        if time.perf_counter() - t0 > seconds_budget / 2 or reps >= 20:
            break
    dt = (time.perf_counter() - t0) / reps
    qps_full = args.batch / dt * n / args.rows
    return qps_full, dt, f"{n} of {args.rows} rows x {args.dim} fp32, batch {args.batch}, {reps} reps; " \
                         f"time scaled linearly in rows", cores


def hnsw_baseline(args):
    """Reference-STYLE comparator: an HNSW graph at the parameters of the
    reference's shipped index (M 16, ef_construction 100) and Chroma's default search_ef 10,
    restated in C (oracle/hnsw_restatement.c -- NOT Chroma itself, which cannot be installed
    here), on a bounded sample.  Reports recall@k against exact search on the same sample, and
    recall@1 on queries planted next to a corpus row.  On isotropic random data in hundreds of
    dimensions graph search at ef = 10 is close to useless; real embeddings behave better."""
    from oracle.exact_search import fast_topk_f32, prepare_corpus
    from oracle.hnsw import HnswIndex
    n = min(args.hnsw_sample_rows, args.rows)
    rng = np.random.default_rng(1234)
    x = prepare_corpus("cosine", rng.standard_normal((n, args.dim), dtype=np.float32))
    nq = 512
    q = prepare_corpus("cosine", rng.standard_normal((nq, args.dim), dtype=np.float32))
    planted = rng.choice(n, nq // 2, replace=False)
    q[: nq // 2] = prepare_corpus("cosine", x[planted] + 0.05 * rng.standard_normal((nq // 2, args.dim), dtype=np.float32))
    t0 = time.perf_counter()
    idx = HnswIndex(x)
    build_s = time.perf_counter() - t0
    idx.query(q[:8], args.k, 10)
    t0 = time.perf_counter()
    ids, _ = idx.query(q, args.k, 10)
    dt = time.perf_counter() - t0
    want, _ = fast_topk_f32("l2", q, x, args.k)
    recall = float(np.mean([len(set(ids[i]) & set(want[i])) / args.k for i in range(nq)]))
    recall_planted = float(np.mean(ids[: nq // 2, 0] == planted))
    idx.close()
    return {"kind": "restatement of hnswlib at Chroma defaults (M=16, ef_construction=100, search_ef=10, l2); not Chroma",
            "sample": f"{n} x {args.dim} fp32 unit-norm rows, {nq} queries (half planted at sigma 0.05)",
            "build_seconds": round(build_s, 1), "qps": nq / dt, "threads": os.cpu_count(),
            "recall_at_k": recall, "recall_at_1_planted": recall_planted}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Its arithmetic is
    chromadb/hnswlib (not installable here: no wheel, no network), so this arm times the oracle port
    -- exact brute force, which is also what Chroma itself runs at the reference's shipped scale -- on
    ALL host cores, on the FULL fp32 corpus when the host has the memory for it (10M x 768 x 4 B =
    30.7 GB) and on the largest sample that fits otherwise (time scaled linearly in rows, said so in
    `sample`).  A step is one query batch against the corpus; W warm-up steps, then K timed steps --
    cut short (and `steps` says how many ran) once ~150 s of timed work have passed."""
    if int(os.environ.get("RANK", "0")) != 0:
        return                      # under torchrun only rank 0 works; the others leave the cores to it
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:           # noqa: BLE001 - the environment variables above already ask for all cores
        pass
    from oracle.exact_search import fast_topk_f32, prepare_corpus
    cores = os.cpu_count() or 1
    need = args.rows * args.dim * 4
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:           # noqa: BLE001
        avail = 8 << 30
    n = args.rows if need * 1.25 + (4 << 30) < avail else max(100_000, int((avail - (4 << 30)) / 1.25 / (args.dim * 4)))
    n = min(n, args.rows)
    t0 = time.perf_counter()
    x = np.empty((n, args.dim), dtype=np.float32)
    _fill_normal_parallel(x, 1234, min(cores, 32))
    if args.space == "cosine":
        pass                        # rows are unit-norm already: what hnswlib's cosine space stores
    gen_s = time.perf_counter() - t0
    W, K = max(0, args.warmup), max(1, args.steps)
    q_all = gaussian_queries(W + K, args.batch, args.dim)
    q_all = plant(q_all, x[: (W + K) * args.batch // 10 + 1].copy())
    # warm-up: at least one full pass (pages in the corpus, spins up the BLAS threads)
    t_w0 = time.perf_counter()
    for i in range(max(1, W)):
        fast_topk_f32(args.space, prepare_corpus(args.space, q_all[i % (W + K)], "f32"), x, args.k)
        if time.perf_counter() - t_w0 > 30:
            break
    per, t_total0 = [], time.perf_counter()
    for i in range(K):
        q = prepare_corpus(args.space, q_all[W + i], "f32")
        ts = time.perf_counter()
        fast_topk_f32(args.space, q, x, args.k)
        per.append(time.perf_counter() - ts)
        if time.perf_counter() - t_total0 > 150:
            break
    scale = n / args.rows                         # < 1 only if the full corpus did not fit
    ms_per_step = 1e3 * statistics.mean(per) / scale
    v = args.batch / (ms_per_step / 1e3)
    sample = (f"{n} of {args.rows} rows x {args.dim} fp32 resident in host RAM ({'FULL corpus' if n == args.rows else 'sample, time scaled linearly in rows'}), "
              f"batch {args.batch}, {len(per)} timed steps after {max(1, W)} warm-up passes; numpy/OpenBLAS sgemm + argpartition on {cores} threads; "
              f"corpus generated in {gen_s:.0f} s (not timed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": len(per), "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args),
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "p50_ms": 1e3 * statistics.median(per) / scale,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class Env:
    """torch / distributed plumbing of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def corpus_chunks(env, n_local, dim, space, seed, chunk=500_000):
    """The rank's synthetic shard, chunk by chunk on the device: N(0, 1) rows, unit-normalised.  The SAME
    generator sequence is replayed by the verification pass, so the brute force sees the rows the store got."""
    torch = env.torch
    gen = torch.Generator(device=env.dev)
    gen.manual_seed(seed)
    for s in range(0, n_local, chunk):
        m = min(chunk, n_local - s)
        xs = torch.randn((m, dim), generator=gen, device=env.dev, dtype=torch.float32)
        xs = torch.nn.functional.normalize(xs, dim=1)          # unit-norm rows in every space
        yield s, xs


def build_store(env, rag, rows_local, dim, dtype, space, seed, rerank, pk):
    """Fill a store from device-generated chunks; times the upsert kernel (K1) with CUDA events."""
    torch = env.torch
    store = rag.DeviceStore(dim, dtype, space, device=env.local_rank, capacity_hint=rows_local, rerank=bool(rerank))
    t0 = time.perf_counter()
    k_ms, k_rows = 0.0, 0
    for s, xs in corpus_chunks(env, rows_local, dim, space, seed):
        torch.cuda.synchronize(env.dev)
        store.upsert_device(xs.data_ptr(), xs.shape[0])      # K1 normalises (cosine) + converts on the way in; returns when done
        k_ms += store.last_upsert_ms()                       # CUDA events around the kernel on the store's admin stream
        k_rows += xs.shape[0]
        del xs
    build_s = time.perf_counter() - t0
    assert store.count() == rows_local
    row_bytes = dim * (2 if dtype == "bf16" else 4)
    per_row = dim * 4 + row_bytes + (dim * 4 if store.rerank else 0)
    gbs = k_rows * per_row / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
    info = {"seconds": round(build_s, 2), "rows_per_s": k_rows / (k_ms / 1e3) if k_ms else None,
            "upsert_kernel": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                              "algorithmic_bytes_per_row": per_row,
                              "timing": "CUDA events around upsert_kernel on the store's admin stream, summed over the 500k-row chunks"}}
    return store, info


def time_steps(env, fn, W, K):
    """W warm-up calls, then K calls back to back between two CUDA events; max over ranks.  Returns ms per step."""
    torch = env.torch
    for i in range(W):
        fn(i)
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    e0.record()
    for i in range(K):
        fn(W + i)
    e1.record()
    env.barrier()
    return env.max_over_ranks(e0.elapsed_time(e1)) / K


def kernel_time(store, q_host, k, mask_slot, regime, n=5):
    """Per-launch duration of the dominant kernel: CUDA events around the launch on the engine's own stream
    (synchronous API), mean of n."""
    kms, regime_seen = [], None
    for i in range(n):
        store.query(q_host[i % len(q_host)], k, mask_slot=mask_slot, regime=regime)
        info = store.last_query_info()
        kms.append(info["kernel_ms"])
        regime_seen = info["regime"]
    return statistics.mean(kms), regime_seen


def roofline_of(pk, regime_seen, B, rows, dim, dtype, kernel_ms, live_frac=1.0, masks=0, bitmap_rows=None, shadow=None):
    """The bound that applies: HBM (bytes the kernel is designed to touch: live & passing rows + bitmaps) or the
    tensor pipe (2 B N D).  For the tensor kernel the larger of the two fractions is reported.  fp32 stores in the
    tensor regime stream their bf16 shadow (`shadow`: "hi" = 2 bytes per element, "hilo" = 4)."""
    row_bytes = dim * (2 if dtype == "bf16" else 4)
    if dtype == "f32" and regime_seen == "tensor" and shadow == "hi":
        row_bytes = dim * 2
    alg_bytes = float(rows) * live_frac * row_bytes + (bitmap_rows if bitmap_rows is not None else rows) / 8.0 * (1 + masks)
    hbm = alg_bytes / (kernel_ms / 1e3) / 1e9
    kern = "gemm_topk_kernel" if regime_seen == "tensor" else "scan_stream_kernel"
    out = {"bound": "hbm", "achieved": hbm, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm / pk["hbm_gbs"],
           "traffic": None, "peak_source": pk["source"], "kernel": kern, "kernel_ms": kernel_ms,
           "algorithmic_bytes": alg_bytes, "frac_of_8TBs_nominal": hbm / 8000.0}
    if regime_seen == "tensor":
        flops = 2.0 * B * rows * dim
        tf = flops / (kernel_ms / 1e3) / 1e12
        if tf / pk["bf16_tflops"] > out["frac"]:
            out = {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                   "frac": tf / pk["bf16_tflops"], "traffic": None, "peak_source": pk["source"] + " burst",
                   "frac_of_sustained_peak": tf / pk["bf16_tflops_sustained"], "kernel": kern, "kernel_ms": kernel_ms,
                   "algorithmic_flops": flops}
            if dtype == "f32" and shadow != "hi":
                # fp32 rows are contracted as bf16 hi/lo pairs: 3 MMAs per algorithmic one
                out["executed_tflops"] = 3.0 * tf
                out["executed_frac"] = 3.0 * tf / pk["bf16_tflops"]
    if dtype == "f32" and regime_seen == "tensor":
        out["streams"] = f"bf16 shadow of the fp32 rows ({shadow}): {row_bytes} B per row"
        out["frac_on_fp32_row_bytes"] = float(rows) * live_frac * dim * 4 / (kernel_ms / 1e3) / 1e9 / pk["hbm_gbs"]
    return out


def attach_traffic(roof, rows, dim, dtype, B, selectivity=None):
    """DRAM traffic of the kernel from the committed ncu captures (profiles/traffic.json), valid only for
    the shard shape it was captured at."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return
    for ent in json.load(open(tpath)).get("kernels", []):
        if (ent["kernel"] == roof["kernel"] and ent["rows"] == rows and ent["dim"] == dim
                and ent["dtype"] == dtype and ent["batch"] == B and ent.get("where_selectivity") == selectivity):
            roof["traffic"] = ent["dram_bytes_per_launch"]
            roof["traffic_source"] = ent["source"]


def brute_force_topk(env, q_dev, rows_local, dim, space, seed, k, row_base, drop=None):
    """Exact fp32 top-k of this rank's shard by chunked torch matmul over the REGENERATED rows (un-rounded
    fp32; TF32 off).  Returns (dists [B, k] fp32, global rows [B, k] int64) on the device."""
    torch = env.torch
    torch.backends.cuda.matmul.allow_tf32 = False
    B = q_dev.shape[0]
    qn = torch.nn.functional.normalize(q_dev, dim=1) if space == "cosine" else q_dev
    best_d = torch.full((B, k), float("inf"), device=env.dev)
    best_r = torch.full((B, k), -1, dtype=torch.int64, device=env.dev)
    for s0, big in corpus_chunks(env, rows_local, dim, space, seed):       # same chunking as the build: same random stream
        for off in range(0, big.shape[0], 250_000):
            xs = big[off:off + 250_000]
            s = s0 + off
            if space == "l2":
                d = (qn * qn).sum(1, keepdim=True) + (xs * xs).sum(1)[None, :] - 2.0 * (qn @ xs.T)
            else:
                d = 1.0 - qn @ xs.T
            if drop is not None:
                d[:, drop[s:s + xs.shape[0]]] = float("inf")
            kk = min(k, d.shape[1])
            dd, ii = torch.topk(d, kk, dim=1, largest=False)
            cat_d = torch.cat([best_d, dd], 1)
            cat_r = torch.cat([best_r, ii + (s + row_base)], 1)
            o = torch.argsort(cat_d, dim=1, stable=True)[:, :k]
            best_d, best_r = torch.gather(cat_d, 1, o), torch.gather(cat_r, 1, o)
            del d
        del big
    return best_d, best_r


def verify_global(env, args, store, searcher, n_local, stride, seed, q_fused, q_batch):
    """Outside every timed region.  The engine's answers on THIS corpus (all shards) against exact fp32 brute
    force on the un-rounded rows and queries:
      * q_fused: a few single-query steps through the same call the e2e pass times (fused exchange at N > 1);
      * q_batch: >= 1000 queries in one batch (tensor regime; NCCL all-gather + merge kernel at N > 1).
    Each rank brute-forces its own shard, the per-shard lists are all-gathered and merged on the host by
    (distance, global row).  Also cross-checks the torch brute force against the numpy oracle on a 100k-row
    subsample of rank 0's shard.  Returns the `verified` record; exits non-zero on a mismatch."""
    torch, dist = env.torch, env.dist
    k = args.k
    out = {"ok": True}

    def global_truth(q_np):
        q_dev = torch.from_numpy(np.ascontiguousarray(q_np)).to(env.dev)
        d, r = brute_force_topk(env, q_dev, n_local, args.dim, args.space, seed, k, env.rank * stride)
        if env.world > 1:
            gd = [torch.empty_like(d) for _ in range(env.world)]
            gr = [torch.empty_like(r) for _ in range(env.world)]
            dist.all_gather(gd, d)
            dist.all_gather(gr, r)
            d, r = torch.cat(gd, 1), torch.cat(gr, 1)
        d, r = d.cpu().numpy(), r.cpu().numpy()
        o = np.lexsort((r, d), axis=1)[:, :k]           # host merge by (distance, global row)
        return np.take_along_axis(d, o, 1), np.take_along_axis(r, o, 1)

    def compare(name, got_r, got_d, want_d, want_r):
        B = got_r.shape[0]
        recall = float(np.mean([len(set(got_r[b].tolist()) & set(want_r[b].tolist())) / k for b in range(B)]))
        same = got_r == want_r
        dist_ok = bool(np.allclose(got_d[same], want_d[same], rtol=1e-5, atol=2e-6)) and \
            bool(np.allclose(np.sort(got_d, 1), np.sort(want_d, 1), rtol=1e-4, atol=1e-5))
        out[name] = {"queries": B, "recall_at_k": recall, "rows_identical": float(np.mean(np.all(same, axis=1))),
                     "max_abs_dist_err_on_common_rows": float(np.max(np.abs(got_d[same] - want_d[same]))) if same.any() else None}
        floor = 0.999 if (args.dtype == "f32" or store.rerank) else 0.97
        if recall < floor or not dist_ok:
            out["ok"] = False
        return recall

    # single-query steps, the e2e call
    got = [searcher.search(q_fused[i], k, regime=args.regime) for i in range(q_fused.shape[0])]
    want_d, want_r = global_truth(q_fused.reshape(-1, args.dim))
    compare("single_query_steps", np.concatenate([g[0] for g in got]), np.concatenate([g[1] for g in got]), want_d, want_r)
    out["single_query_steps"]["path"] = searcher.last_path or "one shard"
    # one big batch
    got_r, got_d, _ = searcher.search(q_batch, k, regime=args.regime)
    want_d, want_r = global_truth(q_batch)
    rec = compare("batch", got_r, got_d, want_d, want_r)
    out["batch"]["path"] = searcher.last_path or "one shard"
    out["queries"] = int(q_fused.shape[0] + q_batch.shape[0])
    # the brute force itself against the numpy oracle (fp64 accumulate) on a subsample of rank 0's shard
    if env.rank == 0:
        from oracle.exact_search import exact_search
        m = min(100_000, n_local)
        sub = next(corpus_chunks(env, n_local, args.dim, args.space, seed))[1][:m]       # first rows of the first build chunk
        m = sub.shape[0]
        qd = torch.from_numpy(np.ascontiguousarray(q_batch[:16])).to(env.dev)
        qn = torch.nn.functional.normalize(qd, dim=1)
        dd, ii = torch.topk(1.0 - qn @ sub.T if args.space != "l2" else torch.cdist(qn, sub) ** 2, k, dim=1, largest=False)
        orows, od = exact_search(args.space, q_batch[:16], sub.cpu().numpy(), k, None, "f32")
        agree = float(np.mean([len(set(ii[b].tolist()) & set(orows[b].tolist())) / k for b in range(16)]))
        out["brute_force_vs_numpy_oracle"] = {"rows": m, "queries": 16, "recall_at_k": agree,
                                              "max_abs_dist_diff": float(np.max(np.abs(np.sort(dd.cpu().numpy(), 1) - np.sort(np.stack(od), 1))))}
        if agree < 0.999:
            out["ok"] = False
    ok = torch.tensor([1 if out["ok"] else 0], device=env.dev)
    if env.world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["ok"] = bool(int(ok.item()))
    return out, rec


def sharding_desc(world, exchange_path):
    if world == 1:
        return "one shard (no exchange)"
    if exchange_path == "fused":
        return (f"row-wise x{world}; scan + all-gather of Bxk keys over NVLink peer memory + merge fused in ONE launch "
                f"per GPU (no NCCL call on the data path)")
    return f"row-wise x{world}, NCCL all-gather of Bxk keys + merge kernel"


def measure_batch(env, args, store, searcher, B, k, n_local, pk, mask_slot=-1, K=30, W=3, live_frac=1.0, masks=0, seed=99,
                  selectivity=None):
    """Device-resident QPS + roofline of one batch size on an existing store."""
    torch = env.torch
    q_dev = torch.from_numpy(gaussian_queries(W + K, B, store.dim, seed=seed)).to(env.dev)
    ms = time_steps(env, lambda i: searcher.search_device(q_dev[i], k, mask_slot=mask_slot, regime=args.regime), W, K)
    q_host = q_dev[:5].cpu().numpy()
    kernel_ms, regime_seen = kernel_time(store, q_host, k, mask_slot, args.regime)
    if regime_seen == "tensor":
        kernel_ms = ms        # the contraction is > 99.8 % of a step (profiles/): the timed region itself, at its clocks
    out = {"batch": B, "k": k, "value": B / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms, "regime": regime_seen,
           "steps": K, "warmup": W}
    shadow = None
    if store.dtype == "f32" and regime_seen == "tensor":
        ti = store.f32_tensor_info()
        shadow = ti["shadow"]
        out["f32_tensor"] = {"shadow": shadow, "queries": ti["queries"], "exact_reruns": ti["reruns"],
                             "note": "bf16 shadow contracted on the tensor cores, survivors re-ranked from the fp32 rows, "
                                     "each query certified by the guard or re-run on the exact stream kernel"}
    out["roofline"] = roofline_of(pk, regime_seen, B, n_local, store.dim, store.dtype, kernel_ms, live_frac, masks, shadow=shadow)
    attach_traffic(out["roofline"], n_local, store.dim, store.dtype, B, selectivity)
    return out


def other_configs(env, args, rag, ShardedSearcher, pk, which):
    """BASELINE configs 2, 4 and 5 (per-GPU shard) on this box, each on its own store (the headline store has
    been freed).  At N > 1 only config 5 runs: every rank holds a 25M-row shard -- at N = 8 that IS config 5."""
    torch = env.torch
    out = []

    def with_store(rows, dim, dtype, space, fn, rerank=args.rerank):
        store, binfo = build_store(env, rag, rows, dim, dtype, space, 4242 + env.rank, rerank, pk)
        searcher = ShardedSearcher(store, env.rank, env.world, row_base=env.rank * rows)
        try:
            fn(store, searcher, binfo)
        finally:
            searcher.close()
            store.close()
            torch.cuda.empty_cache()

    if 2 in which and env.world == 1:
        def cfg2(store, searcher, binfo):
            for B in (1, 32, 1024):
                r = measure_batch(env, args, store, searcher, B, 10, 1_000_000, pk, K=100 if B < 1024 else 30, W=5)
                r["config"] = "BASELINE configs[1]: 1M x 384 fp32, cosine, top-10"
                if B == 1:
                    r["build"] = binfo
                out.append(r)
        with_store(1_000_000, 384, "f32", "cosine", cfg2)

        def cfg2_wide(store, searcher, binfo):      # not a BASELINE config: fp32 rows of the headline width (dim > 384)
            for B in (32, 1024):
                r = measure_batch(env, args, store, searcher, B, 10, 1_000_000, pk, K=50 if B < 1024 else 20, W=5)
                r["config"] = "beyond BASELINE configs[1]: 1M x 768 fp32, cosine, top-10 (fp32 rows wider than 384)"
                out.append(r)
        with_store(1_000_000, 768, "f32", "cosine", cfg2_wide)
    if 4 in which and env.world == 1:
        def cfg4(store, searcher, binfo):
            n = 10_000_000
            rng = np.random.default_rng(77)
            dead = np.nonzero(rng.random(n) < 0.05)[0]
            store.delete(dead)
            live = 1.0 - len(dead) / n
            bucket = np.random.default_rng(99).integers(0, 100, n)          # metadata column bucket = hash(row) % 100
            for slot, pct in enumerate((1, 10, 50)):
                passing = bucket < pct                                      # where={"bucket": {"$lt": pct}} as a bitmap
                store.set_mask(slot, passing)
                r = measure_batch(env, args, store, searcher, 1, 10, n, pk, mask_slot=slot, K=100, W=5,
                                  live_frac=live * float(passing.mean()), masks=1, selectivity=pct / 100.0)
                r["config"] = (f"BASELINE configs[3]: 10M x 384 {args.dtype}, cosine, top-10, where selectivity {pct} %, "
                               f"5 % tombstones")
                r["where_selectivity"] = pct / 100.0
                r["tombstone_fraction"] = len(dead) / n
                out.append(r)
        with_store(10_000_000, 384, args.dtype, "cosine", cfg4)
    if 5 in which:
        def cfg5(store, searcher, binfo):
            r = measure_batch(env, args, store, searcher, 1024, 100, 25_000_000, pk, K=10, W=3)
            r["config"] = (f"BASELINE configs[4]: {25 * env.world}M x 384 bf16 over {env.world} GPU(s) (25M rows per GPU), "
                           f"l2, batch 1024, top-100" + ("" if env.world == 8 else "; the full config is 8 GPUs"))
            r["value_is"] = "aggregate over all ranks: every rank holds its own 25M-row shard (weak scaling in this entry)"
            r["rerank_plane"] = bool(store.rerank)
            r["build"] = binfo
            out.append(r)
        with_store(25_000_000, 384, "bf16", "l2", cfg5)
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    env = Env(args)
    torch, dist = env.torch, env.dist
    import local_rag_system_b200 as rag
    from local_rag_system_b200.sharded import ShardedSearcher, shard_plan
    world, rank, dev = env.world, env.rank, env.dev
    pk = peaks()

    # ---- build this rank's shard on the device (synthetic, seeded per shard) ----
    stride, counts = shard_plan(args.rows, world)
    n_local = counts[rank]
    seed = 1234 + rank
    store, build_info = build_store(env, rag, n_local, args.dim, args.dtype, args.space, seed, args.rerank, pk)

    mask_slot = -1
    live_frac = 1.0
    if args.tombstones > 0:
        rng = np.random.default_rng(77 + rank)
        dead = np.nonzero(rng.random(n_local) < args.tombstones)[0]
        store.delete(dead)
        live_frac *= 1.0 - len(dead) / max(n_local, 1)
    if args.selectivity > 0:
        rng = np.random.default_rng(99 + rank)
        passing = rng.random(n_local) < args.selectivity     # bucket = hash(row) % 100 < s, as a bitmap
        store.set_mask(0, passing)
        mask_slot = 0
        live_frac *= float(passing.mean())
    searcher = ShardedSearcher(store, rank, world, row_base=rank * stride)
    B, k, K, W = args.batch, args.k, args.steps, max(args.warmup, 3)

    # queries: unit-norm Gaussian, every 10th planted next to a row of rank 0's shard; identical on all ranks
    q_np = gaussian_queries(W + K, B, args.dim)
    n_plant = (W + K) * B // 10 + 1
    vq_np = gaussian_queries(1, args.verify_queries, args.dim, seed=777)[0]
    if rank == 0:
        prow = np.random.default_rng(5).choice(max(n_local, 1), min(n_plant + args.verify_queries // 10 + 1, n_local), replace=False)
        prows = store.fetch(prow, exact=True)
        q_np = plant(q_np, prows[:n_plant])
        vq_np = plant(vq_np[None], prows[n_plant:])[0]
    if world > 1:
        for arr in (q_np, vq_np):
            t = torch.from_numpy(arr).to(dev)
            dist.broadcast(t, 0)
            arr[...] = t.cpu().numpy()
    q_host = torch.from_numpy(q_np).pin_memory()
    q_dev = q_host.to(dev)

    verified, recall = None, None
    if not args.no_verify and args.tombstones == 0 and args.selectivity == 0:
        verified, recall = verify_global(env, args, store, searcher, n_local, stride, seed, q_np[:8], vq_np)
        if not verified["ok"]:
            if rank == 0:
                print(json.dumps({"verified": verified}), file=sys.stderr)
            raise SystemExit("verification against the fp32 brute force FAILED")

    # ---- device-resident timing: `value` ----
    # clocks / throttle reasons are sampled from here to the end of the end-to-end pass: warm-up, the timed
    # region, the latency pass and the e2e pass all keep the GPU under the same load (the timed region alone
    # can be shorter than one nvidia-smi sampling period)
    sampler = ClockSampler(env.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(W):
        searcher.search_device(q_dev[i], k, mask_slot=mask_slot, regime=args.regime)
    env.barrier()
    launches0 = store.kernel_launches()
    # throughput: K steps back to back, one event on each side (nothing between the launches, so
    # the scan kernel's programmatic dependent launch can overlap one query's tail with the next)
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ev_a.record()
    for i in range(K):
        searcher.search_device(q_dev[W + i], k, mask_slot=mask_slot, regime=args.regime)
    ev_b.record()
    env.barrier()
    exchange_path = searcher.last_path
    total_ms = env.max_over_ranks(ev_a.elapsed_time(ev_b))
    launches_timed = store.kernel_launches() - launches0
    # per-query latency distribution: the same K steps with an event between consecutive queries
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record()
    for i in range(K):
        searcher.search_device(q_dev[W + i], k, mask_slot=mask_slot, regime=args.regime)
        ev[i + 1].record()
    env.barrier()
    per_step = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(K))
    launches = launches_timed + (K if (world > 1 and exchange_path != "fused") else 0)   # + the cross-shard merge kernel per step
    value = B * K / (total_ms / 1e3)

    # ---- end to end through the host-buffer API: `e2e` ----
    # Every step: its queries go pinned host -> device, the search runs, the B x k result comes back to the host,
    # all inside the timed region.  Two passes:
    #   sequential  one blocking call per step (ShardedSearcher.search): the latency a single caller sees;
    #   in flight   submit() / collect() with 2 queries in the air (a server with concurrent requests, SURVEY.md
    #               8d "pipelined variant"): request i+1 is staged and launched while request i is scanned.
    for i in range(W):
        searcher.search(q_host[i].numpy(), k, mask_slot=mask_slot, regime=args.regime)
    env.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    lat = []
    for i in range(K):
        ts = time.perf_counter()
        searcher.search(q_host[W + i].numpy(), k, mask_slot=mask_slot, regime=args.regime)
        lat.append((time.perf_counter() - ts) * 1e3)
    e1.record()
    env.barrier()
    e2e_seq = B * K / (env.max_over_ranks(e0.elapsed_time(e1)) / 1e3)
    lat.sort()
    depth = max(1, min(4, args.inflight))
    for i in range(W):
        searcher.collect(searcher.submit(q_host[i].numpy(), k, mask_slot=mask_slot, regime=args.regime))
    env.barrier()
    t0 = time.perf_counter()
    pending = []
    for i in range(K):
        pending.append(searcher.submit(q_host[W + i].numpy(), k, mask_slot=mask_slot, regime=args.regime))
        if len(pending) >= depth:
            searcher.collect(pending.pop(0))
    last = None
    while pending:
        last = searcher.collect(pending.pop(0))
    wall_ms = (time.perf_counter() - t0) * 1e3          # every result is on the host: the host clock is the honest one
    env.barrier()
    e2e_value = B * K / (env.max_over_ranks(wall_ms) / 1e3)
    clocks = sampler.stop() if sampler else None
    if searcher.exchange is not None and searcher.exchange.timed_out():
        raise SystemExit(f"[rank {rank}] the fused exchange timed out waiting for a peer: results are invalid")

    # ---- roofline of the dominant kernel, CUDA events on its own stream ----
    kernel_ms, regime_seen = kernel_time(store, [q_host[W + i].numpy() for i in range(min(K, 50))], k, mask_slot, args.regime,
                                         n=min(K, 50))
    timing = "CUDA events around the kernel launch on the engine's stream (sync API), mean of %d" % min(K, 50)
    if regime_seen == "stream" and world == 1:
        # a step IS one launch of this kernel (query preparation, merge and re-ranking are fused into it): use the
        # per-step events of the latency pass (launches serialised by the events, no overlap)
        kernel_ms = statistics.mean(per_step)
        timing = ("per-launch CUDA events on the launching stream (one launch per step, launches separated by "
                  "the events), mean of %d" % K)
    if regime_seen == "tensor":
        kernel_ms = total_ms / K
        timing = "the timed region (the contraction is > 99.8 % of a step)"
    shadow = store.f32_tensor_info()["shadow"] if (args.dtype == "f32" and regime_seen == "tensor") else None
    roof = roofline_of(pk, regime_seen, B, n_local, args.dim, args.dtype, kernel_ms, live_frac, 1 if mask_slot >= 0 else 0,
                       shadow=shadow)
    roof["timing"] = timing
    attach_traffic(roof, n_local, args.dim, args.dtype, B, args.selectivity or None)

    regimes = []
    for xb in [int(v) for v in args.extra_batches.split(",") if v.strip()]:
        if xb == B:
            continue
        r = measure_batch(env, args, store, searcher, xb, k, n_local, pk, mask_slot=mask_slot, live_frac=live_frac,
                          masks=1 if mask_slot >= 0 else 0)
        r["config"] = f"headline corpus, batch {xb}"
        regimes.append(r)
    rerank_on = bool(store.rerank)
    searcher.close()
    store.close()
    torch.cuda.empty_cache()
    which = {int(v) for v in args.configs.split(",") if v.strip()}
    if which:
        regimes += other_configs(env, args, rag, ShardedSearcher, pk, which)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": config_dict(args),
            "value_note": ("pipelined throughput: K launches back to back, consecutive queries overlap by programmatic "
                           "dependent launch (so ms_per_step can be below roofline.kernel_ms); p50_ms is the per-query latency"),
            "details": {"rows_per_gpu": n_local, "sharding": sharding_desc(world, exchange_path), "regime": regime_seen,
                        "rerank_plane": rerank_on, "where_selectivity": args.selectivity or None,
                        "tombstone_fraction": args.tombstones or None, "build": build_info},
            "p50_ms": per_step[len(per_step) // 2], "p95_ms": per_step[int(len(per_step) * 0.95)],
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": B * args.dim * 4,
                    "d2h_bytes_per_step": B * k * 12 + B * 4,
                    "mode": (f"{depth} queries in flight: ShardedSearcher.submit / collect -> rag_store_query_submit / _wait "
                             "(a server with concurrent requests); host wall clock over the K steps, max over ranks"),
                    "sequential_value": e2e_seq, "p50_ms": lat[len(lat) // 2], "p95_ms": lat[int(len(lat) * 0.95)],
                    "sequential_call": ("one blocking call per step: ShardedSearcher.search -> rag_store_query (N = 1) / "
                                        "rag_store_query_fused (N > 1); p50_ms / p95_ms are its per-call latencies")},
            "gpu_launches": int(launches),
            "roofline": roof,
            "clocks": clocks,
            "verified": verified,
            "recall_bf16_vs_fp32": recall if args.dtype == "bf16" else None,
            "regimes": regimes,
        }
        if world == 1 and not args.no_cpu_baseline:
            qps, dt, sample, cores = cpu_exact_qps(args)
            line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                    "sample": sample}
            if args.hnsw_baseline:
                try:
                    line["cpu_baseline"]["hnsw_restatement"] = hnsw_baseline(args)
                except Exception as e:       # noqa: BLE001 - a comparator, never a reason to lose the bench line
                    line["cpu_baseline"]["hnsw_restatement"] = {"unavailable": f"{type(e).__name__}: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
