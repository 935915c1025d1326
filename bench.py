#!/usr/bin/env python
"""bench.py — headline benchmark of the zstdmt hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's own pthread + liblz4 path on the host cores)

Workload (BASELINE.json configs[1]): lz4-mt level 1, 8 GiB synthetic Silesia-mix, 1 MiB chunks, per GPU.
A step = one pass of the per-chunk compress hot path over the whole 8 GiB batch (8192 chunks):
  value : GB/s of (raw bytes in + framed bytes out), device-timed with CUDA events, inputs resident in HBM
  e2e   : same metric through the reference-shaped callback API (LZ4MT_compressCCtx, host buffers; the
          pinned staging copies and H2D/D2H are inside the timed region)
  roofline : the dominant kernel (lz4_blocks_pipe_kernel, the LZ4 block compressor) against the measured HBM copy peak
  cpu_baseline : the unmodified reference (oracle/_ref: lib/lz4-mt_*.c + liblz4 1.9.4) on this box's cores
Multi-GPU: chunks are dealt round-robin (chunk i -> rank i mod N), no collective on the data path;
weak scaling (8 GiB per GPU); time = max over ranks.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size-gib", type=float, default=float(os.environ.get("ZMT_BENCH_GIB", "8")))
    ap.add_argument("--chunk-mib", type=int, default=1)
    ap.add_argument("--ref-sample-gib", type=float, default=float(os.environ.get("ZMT_BENCH_REF_GIB", "2")))
    ap.add_argument("--no-extra", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.p, self.lines = index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _pump(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def traffic_per_launch(algorithmic_bytes):
    """DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full
    capture (profiles/r1_traffic.json, made at a smaller size), scaled by algorithmic bytes to this launch."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["lz4_compress_blocks_kernel"]
        return float(t["dram_bytes"]) / float(t["algorithmic_bytes"]) * algorithmic_bytes
    except Exception:
        return None


def algo_bytes_compress(in_bytes, framed_bytes):
    return in_bytes + framed_bytes


# =============================================================================== reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import _oracle as o
    import zstdmt_b200 as z
    chunk = args.chunk_mib << 20
    cores = os.cpu_count() or 1
    threads = min(cores, 128)                      # LZ4MT_THREAD_MAX
    n = int(min(args.ref_sample_gib, args.size_gib) * (1 << 30)) // chunk * chunk
    src = z.gen_stream(z.GEN_MIX, n, chunk)
    cap = n + n // 64 + (1 << 20)
    out = np.empty(cap, np.uint8)
    st = (ctypes.c_size_t * 5)()
    fn = o.ref().ref_lz4_compress_mem
    def step():
        t = time.perf_counter()
        rc = fn(threads, 1, chunk, src.ctypes.data, n, out.ctypes.data, cap, st)
        dt = time.perf_counter() - t
        assert rc == 0
        return dt, int(st[0])
    for _ in range(args.warmup):
        step()
    times = []
    for _ in range(args.steps):
        dt, outb = step(); times.append(dt)
    tot = sum(times)
    gbs = algo_bytes_compress(n, outb) * args.steps / tot / 1e9
    line = {
        "impl": "reference", "metric": "lz4-mt level-1 compress throughput, bytes in + framed bytes out", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "lz4-mt level 1, 8 GiB synthetic Silesia-mix generator, 1 MiB chunks (BASELINE configs[1])",
                   "chunk_bytes": chunk, "sample_bytes_per_step": n, "threads": threads},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "reference",
                         "sample": "%d MiB of the workload per step, LZ4MT_compressCCtx(T=%d, level 1, 1 MiB chunks), memory-to-memory callbacks" % (n >> 20, threads)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ratio": n / outb,
    }
    print(json.dumps(line))


# =============================================================================== B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import zstdmt_b200 as z

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("--gpus %d needs torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    L = z.lib()
    chunk = args.chunk_mib << 20
    n = int(args.size_gib * (1 << 30)) // chunk * chunk
    nchunks = n // chunk
    peak, peak_src = measured_peaks()

    # ---- inputs: chunk i of this rank is global chunk i*world + rank (round-robin deal)
    t0 = time.time()
    src = z.gen_stream(z.GEN_MIX, n, chunk, first=rank, stride=world)
    gen_s = time.time() - t0
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    step_mb = 256 << 20
    for o in range(0, n, step_mb):
        d_in[o:o + step_mb].copy_(torch.from_numpy(src[o:o + step_mb]))
    comp = z.Lz4DeviceCompressor(n, chunk)
    stream = torch.cuda.current_stream()

    # ---- device-timed steps
    for _ in range(max(args.warmup, 3)):
        comp.run(d_in, stream)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    L.zmt_prof_begin.restype = None
    L.zmt_prof_end.restype = ctypes.c_int
    L.zmt_prof_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        comp.run(d_in, stream)
    ev1.record(stream)
    barrier()
    ms = (ctypes.c_double * 16)(); cnt = (ctypes.c_int * 16)()
    L.zmt_prof_end(ms, cnt, 16)
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    dev_ms = allreduce(dev_ms, dist.ReduceOp.MAX if world > 1 else None)
    framed = int(comp.frame_off[-1].item())
    total_alg = allreduce(float(algo_bytes_compress(n, framed)), dist.ReduceOp.SUM if world > 1 else None)
    value = total_alg * args.steps / (dev_ms * 1e-3) / 1e9
    kernel_ms = ms[0] / max(cnt[0], 1)
    kernels = {"lz4_compress_blocks": ms[0] / max(cnt[0], 1), "xxh32": ms[1] / max(cnt[1], 1), "frame_sizes": ms[2] / max(cnt[2], 1),
               "scan": ms[3] / max(cnt[3], 1), "frame_pack": ms[4] / max(cnt[4], 1)}
    launches = int(sum(cnt[i] for i in range(5)))
    achieved = algo_bytes_compress(n, framed) / (kernel_ms * 1e-3) / 1e9

    # ---- parity gate at full size (not timed): GPU decode of the GPU stream == the input, on the device
    out, foff = comp.out, comp.frame_off
    foff_h = foff.cpu().numpy().astype(np.uint64)
    sizes = (np.diff(foff_h) - 12).astype(np.uint32)
    dec = z.Lz4DeviceDecompressor(foff_h[:-1], sizes, [chunk] * nchunks)
    dout, status = dec.run(out, stream)
    torch.cuda.synchronize()
    assert int(status.abs().sum().item()) == 0, "GPU decode of the GPU stream reported errors"
    assert torch.equal(dout[:n], d_in), "round trip mismatch at full size"
    extra = {"ratio": n / framed, "gen_seconds": gen_s, "kernel_ms": kernels}

    if not args.no_extra:
        # device-timed decode of our own stream (independent blocks)
        for _ in range(3):
            dec.run(out, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            dec.run(out, stream)
        e1.record(stream); torch.cuda.synchronize()
        dms = e0.elapsed_time(e1) / args.steps
        extra["lz4_decompress_device_gbs"] = (n + framed) / (dms * 1e-3) / 1e9
        extra["lz4_decompress_ms_per_step"] = dms
        # BASELINE config 3 class: decode of a stream framed by the REFERENCE (liblz4: linked blocks -> one warp per frame)
        try:
            import _oracle as o
            if o.have_ref() and rank == 0:
                rn = min(n, 1 << 30)
                rc, rframed, rst = o.ref_compress(o.CODEC_LZ4, src[:rn], threads=min(os.cpu_count() or 1, 128), level=1, chunk=chunk)
                assert rc == 0
                roffs, rsizes = z.scan_frames(rframed)
                rdec = z.Lz4DeviceDecompressor(roffs, rsizes, [chunk] * (rn // chunk))
                d_rf = torch.from_numpy(rframed).cuda()
                ro, rs = rdec.run(d_rf, stream); torch.cuda.synchronize()
                assert int(rs.abs().sum().item()) == 0 and torch.equal(ro[:rn], d_in[:rn]), "decode of reference-framed stream mismatch"
                for _ in range(2):
                    rdec.run(d_rf, stream)
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(args.steps):
                    rdec.run(d_rf, stream)
                e1.record(stream); torch.cuda.synchronize()
                rms = e0.elapsed_time(e1) / args.steps
                extra["lz4_decompress_reference_frames_device_gbs"] = (rn + rframed.size) / (rms * 1e-3) / 1e9
                extra["lz4_decompress_reference_frames_note"] = "%d MiB framed by the reference (liblz4 level 1, linked blocks): one warp per frame" % (rn >> 20)
                del rdec, d_rf, ro
        except AssertionError:
            raise
        except Exception as e:
            extra["lz4_decompress_reference_frames_note"] = "skipped: %r" % (e,)
    del dec, dout

    # ---- end to end through the reference-shaped callback API (host buffers)
    cap = z.mt_bound(n, chunk)
    h_out = np.empty(cap, np.uint8)
    st = (ctypes.c_size_t * 5)()
    threads = 4
    def e2e_step():
        rc = L.zmt_lz4_compress_mem(threads, 1, chunk, src.ctypes.data, n, h_out.ctypes.data, cap, st)
        assert rc == 0, rc
        return int(st[0])
    e2e_step()                                       # warm-up (allocates the pinned rings)
    barrier()
    sampler2 = ClockSampler(local); sampler2.start()
    t = time.perf_counter()
    for _ in range(args.steps):
        outb = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t
    sampler2.stop()
    e2e_s = allreduce(e2e_s, dist.ReduceOp.MAX if world > 1 else None)
    e2e_alg = allreduce(float(algo_bytes_compress(n, outb)), dist.ReduceOp.SUM if world > 1 else None)
    e2e_val = e2e_alg * args.steps / e2e_s / 1e9
    assert outb == framed
    # spot-check the e2e bytes against the device-path bytes
    assert np.array_equal(h_out[: 1 << 20], out[: 1 << 20].cpu().numpy())
    if not args.no_extra:
        # decompress end to end through LZ4MT_decompressDCtx (host buffers), on the stream just produced
        back = np.empty(n + 16, np.uint8)
        def d_step():
            rc = L.zmt_lz4_decompress_mem(threads, 0, h_out.ctypes.data, outb, back.ctypes.data, n + 16, st)
            assert rc == 0 and int(st[0]) == n, (rc, int(st[0]))
        d_step()
        barrier()
        t = time.perf_counter()
        for _ in range(args.steps):
            d_step()
        barrier()
        d_s = allreduce(time.perf_counter() - t, dist.ReduceOp.MAX if world > 1 else None)
        extra["lz4_decompress_e2e_gbs"] = e2e_alg * args.steps / d_s / 1e9
        assert np.array_equal(back[: 1 << 22], src[: 1 << 22]) and np.array_equal(back[n - (1 << 20): n], src[n - (1 << 20):])
        del back

    if not args.no_extra:
        # ---- BASELINE config 4 class: zstd-mt level 3 on synthetic text, 1 MiB chunks (2 GiB sample), device-timed + e2e
        del comp
        torch.cuda.empty_cache()
        zn = min(n, 2 << 30)
        ztxt = z.gen_stream(z.GEN_TEXT, zn, chunk, first=rank, stride=world)
        zd_in = d_in[:zn]; zd_in.copy_(torch.from_numpy(ztxt))
        zc = z.ZstdDeviceCompressor(zn, chunk)
        for _ in range(3):
            zc.run(zd_in, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            zc.run(zd_in, stream)
        e1.record(stream); torch.cuda.synchronize()
        zms = e0.elapsed_time(e1) / args.steps
        zframed = int(zc.frame_off[-1].item())
        zf_host = zc.out[:zframed].cpu().numpy()
        zdec = z.ZstdDeviceDecompressor(zf_host)
        zout, zst = zdec.run(zc.out, stream); torch.cuda.synchronize()
        assert int(zst.abs().sum().item()) == 0 and torch.equal(zout[:zn], zd_in), "zstd round trip mismatch"
        for _ in range(2):
            zdec.run(zc.out, stream)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            zdec.run(zc.out, stream)
        e1.record(stream); torch.cuda.synchronize()
        zdms = e0.elapsed_time(e1) / args.steps
        zh = np.empty(z.mt_bound(zn, chunk), np.uint8)
        L.zmt_zstd_compress_mem(threads, 3, chunk, ztxt.ctypes.data, zn, zh.ctypes.data, zh.size, st)
        t = time.perf_counter()
        rc = L.zmt_zstd_compress_mem(threads, 3, chunk, ztxt.ctypes.data, zn, zh.ctypes.data, zh.size, st)
        ze = time.perf_counter() - t
        assert rc == 0 and int(st[0]) == zframed
        zb = np.empty(zn + 16, np.uint8)
        L.zmt_zstd_decompress_mem(threads, 0, zh.ctypes.data, zframed, zb.ctypes.data, zn + 16, st)
        t = time.perf_counter()
        rc = L.zmt_zstd_decompress_mem(threads, 0, zh.ctypes.data, zframed, zb.ctypes.data, zn + 16, st)
        zde = time.perf_counter() - t
        assert rc == 0 and np.array_equal(zb[: 1 << 22], ztxt[: 1 << 22])
        extra["zstd"] = {"workload": "zstd-mt level 3 (predefined FSE tables), %d MiB synthetic text, 1 MiB chunks (BASELINE configs[3] class)" % (zn >> 20),
                         "ratio": zn / zframed, "compress_device_gbs": (zn + zframed) / (zms * 1e-3) / 1e9, "compress_ms": zms,
                         "decompress_device_gbs": (zn + zframed) / (zdms * 1e-3) / 1e9, "decompress_ms": zdms,
                         "compress_e2e_gbs": (zn + zframed) / ze / 1e9, "decompress_e2e_gbs": (zn + zframed) / zde / 1e9}
        if world == 1:
            try:
                import _oracle as o
                if o.have_ref():
                    T = min(os.cpu_count() or 1, 128)
                    capr = zn + zn // 64 + (1 << 20); outr = np.empty(capr, np.uint8); s5 = (ctypes.c_size_t * 5)()
                    best = None
                    for _ in range(2):
                        tt = time.perf_counter(); rc = o.ref().ref_zstd_compress_mem(T, 3, chunk, ztxt.ctypes.data, zn, outr.ctypes.data, capr, s5); dt = time.perf_counter() - tt
                        assert rc == 0; best = dt if best is None else min(best, dt)
                    rfr = int(s5[0])
                    extra["zstd"]["reference_cpu"] = {"threads": T, "compress_gbs": (zn + rfr) / best / 1e9, "ratio": zn / rfr}
                    bestd = None
                    for _ in range(2):
                        tt = time.perf_counter(); rc = o.ref().ref_zstd_decompress_mem(T, 0, outr.ctypes.data, rfr, zb.ctypes.data, zn + 16, s5); dt = time.perf_counter() - tt
                        assert rc == 0; bestd = dt if bestd is None else min(bestd, dt)
                    extra["zstd"]["reference_cpu"]["decompress_gbs"] = (zn + rfr) / bestd / 1e9
            except Exception as e:
                extra["zstd"]["reference_cpu"] = {"error": repr(e)}
        del zc, zdec, zout, zh, zb

    line = {
        "metric": "lz4-mt level-1 compress throughput, bytes in + framed bytes out", "value": value, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "lz4-mt level 1, 8 GiB synthetic Silesia-mix generator, 1 MiB chunks, per GPU (BASELINE configs[1])",
                   "chunk_bytes": chunk, "bytes_per_gpu": n, "chunks_per_gpu": nchunks, "parallelism": "round-robin chunks over %d GPU(s), no collective" % world,
                   "l2_policy": "inputs (%.1f GiB) larger than L2" % (n / 2 ** 30)},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": n + 4 * nchunks, "d2h_bytes_per_step": outb + 8 * (nchunks + nchunks // 64 + 1),
                "api": "LZ4MT_compressCCtx via in-memory fn_read/fn_write (csrc/memio_glue.c), threads=%d" % threads, "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_per_launch(algo_bytes_compress(n, framed)),
                     "kernel": "lz4_blocks_pipe_kernel (LZ4 block compressor, two-team pipeline)", "kernel_ms": kernel_ms, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes_compress(n, framed)},
        "extra": extra,
    }

    # ---- CPU baseline beside it (rank 0, N=1 only): the unmodified reference on this box's cores
    if world == 1:
        try:
            import _oracle as o
            if o.have_ref():
                cores = os.cpu_count() or 1
                T = min(cores, 128)
                ns = min(n, int(args.ref_sample_gib * (1 << 30)))
                capr = ns + ns // 64 + (1 << 20)
                outr = np.empty(capr, np.uint8); str_ = (ctypes.c_size_t * 5)()
                best = None
                for _ in range(3):
                    tt = time.perf_counter()
                    rc = o.ref().ref_lz4_compress_mem(T, 1, chunk, src.ctypes.data, ns, outr.ctypes.data, capr, str_)
                    dt = time.perf_counter() - tt
                    assert rc == 0
                    best = dt if best is None else min(best, dt)
                v = algo_bytes_compress(ns, int(str_[0])) / best / 1e9
                ns1 = min(ns, 256 << 20)
                tt = time.perf_counter()
                rc = o.ref().ref_lz4_compress_mem(1, 1, chunk, src.ctypes.data, ns1, outr.ctypes.data, capr, str_)
                dt1 = time.perf_counter() - tt
                line["cpu_baseline"] = {"value": v, "unit": "GB/s", "cores": T, "kind": "reference",
                                        "sample": "first %d MiB of the same workload, LZ4MT_compressCCtx(T=%d) best of 3; T=1 on %d MiB: %.3f GB/s"
                                                  % (ns >> 20, T, ns1 >> 20, algo_bytes_compress(ns1, int(str_[0])) / dt1 / 1e9)}
                # reference ratio on the sample (re-run value kept from the T=N run)
                rc = o.ref().ref_lz4_compress_mem(T, 1, chunk, src.ctypes.data, ns, outr.ctypes.data, capr, str_)
                line["cpu_baseline"]["ratio"] = ns / int(str_[0])
                if not args.no_extra:
                    # the reference decoding its own stream (T=nproc), same sample
                    fr = outr[: int(str_[0])].copy(); backr = np.empty(ns + 16, np.uint8); bestd = None
                    for _ in range(3):
                        tt = time.perf_counter()
                        rc = o.ref().ref_lz4_decompress_mem(T, 0, fr.ctypes.data, fr.size, backr.ctypes.data, ns + 16, str_)
                        dt = time.perf_counter() - tt
                        assert rc == 0
                        bestd = dt if bestd is None else min(bestd, dt)
                    line["cpu_baseline"]["lz4_decompress_gbs"] = (ns + fr.size) / bestd / 1e9
                line["cpu_baseline"]["cpu_model"] = cpu_model()
            else:
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
        except Exception as e:  # the baseline must never take the GPU result down with it
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
