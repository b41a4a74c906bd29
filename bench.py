#!/usr/bin/env python
"""Benchmark of the SSD box-geometry hot path (BASELINE.json metric: images/sec for
target-assign + loss + NMS, SSD300 / 8732 priors) -- one JSON line on stdout.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the reference's CPU path (NumPy port) on the host cores

A step is one pass of the chain over one synthetic batch per GPU (weak scaling: the per-GPU batch is
fixed).  `value` is measured with the inputs resident in HBM; `e2e` goes through the public host API
(ssdgeom.pipeline.HotPath.step_host) with pinned host buffers and the copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ssd-object-detection_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "images/sec for target-assign+loss+NMS (SSD300, 8732 anchors)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--table", default="ssd300", choices=["ssd300", "ssd512"])
    ap.add_argument("--max-gt", type=int, default=100)
    ap.add_argument("--gt-mode", default="max", choices=["max", "coco"])
    ap.add_argument("--cpu-images", type=int, default=0, help="images in the CPU-baseline sample (0 = host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pipeline", default="split", choices=["split", "two-stream"],
                    help="split: NMS and loss on their own streams (loss outranks NMS); two-stream: one stream per branch")
    ap.add_argument("--no-fused", action="store_true",
                    help="loss streams the logits itself (ce_kernel) instead of reusing the filter pass's row statistics")
    ap.add_argument("--mining", default="shard", choices=["shard", "global"],
                    help="hard-negative threshold over this rank's batch (reference split-batch semantics) or "
                         "over all ranks' batches (exact-global, 5 small all-reduces)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle/ssd_oracle.py, a NumPy port that keeps the reference's own
# arg-max sweeps) on the host cores.  Used by --impl reference and by the cpu_baseline leg only.
# ---------------------------------------------------------------------------------------------------------
def _cpu_image(task):
    from oracle import ssd_oracle as O
    (table, cls, boxes, pred_cls, pred_box) = task
    priors = _cpu_image.priors.get(table)
    if priors is None:
        from ssdgeom import synth
        t = synth.TABLES[table]
        priors = O.build_prior_box(t["sizes"], t["s_k_refer"], t["aspect_ratio"], t["input_size"])
        _cpu_image.priors[table] = priors
    lab, loc, mask = O.assign_encode(cls, boxes, priors, 0.5, sweeps=True)     # utils/bbox.py:44-101 as written
    kept, count, _, _ = O.detect(pred_cls, pred_box, priors)
    return lab, loc, mask, int(count.sum())


_cpu_image.priors = {}


class CpuArm:
    def __init__(self, args, images):
        import multiprocessing as mp
        from ssdgeom import synth
        self.cores = os.cpu_count() or 1
        self.images = images if images > 0 else self.cores
        self.table = args.table
        a = synth.num_priors(synth.TABLES[args.table])
        boxes, cls, off = synth.make_gt(1234, self.images, args.max_gt, args.gt_mode)
        pred_cls, pred_box = synth.make_predictions(1234, self.images, a)
        self.tasks = [(args.table, cls[off[i]:off[i + 1]], boxes[off[i]:off[i + 1]], pred_cls[i], pred_box[i])
                      for i in range(self.images)]
        self.pred = (pred_box, pred_cls)
        self.pool = mp.get_context("fork").Pool(min(self.cores, self.images))

    def step(self):
        from oracle import ssd_oracle as O
        res = self.pool.map(_cpu_image, self.tasks, chunksize=1)
        y_true = tuple(np.stack([r[k] for r in res]) for k in range(3))
        total, _ = O.ssd_loss(y_true, self.pred)                                  # models/ssd_model.py:341-396
        return total

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(args, args.cpu_images)
    for _ in range(max(args.warmup, 0)):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step()
    dt = time.perf_counter() - t0
    arm.close()
    value = arm.images * args.steps / dt
    cb = cpu_desc(arm, args, value)
    line = base_line(args, value, dt / args.steps * 1e3, n_gpus=args.gpus)
    line.update({"impl": "reference", "dtype": "f64", "cpu_baseline": cb, "gpu_launches": 0,
                 "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    from ssdgeom import synth
    line["config"]["priors"] = synth.num_priors(synth.TABLES[args.table])
    line["config"]["l2"] = "inputs larger than L2 (logits %.0f MB per GPU)" % (
        args.batch * line["config"]["priors"] * 81 * 4 / 1e6)
    print(json.dumps(line), flush=True)


def cpu_desc(arm, args, value):
    return {"value": value, "unit": "images/s", "cores": min(arm.cores, arm.images), "kind": "port",
            "sample": "%d %s images per step (GT mode %s, <=%d GT): match_bbox+apply_anchor_box with the reference's "
                      "arg-max sweeps and decode+per-class NMS per image in a %d-process pool, then _ssd_loss on the "
                      "sample; NumPy port (TensorFlow is not installed)" %
                      (arm.images, args.table, args.gt_mode, args.max_gt, min(arm.cores, arm.images))}


def workload_name(args):
    return "%s chained assign+encode -> multibox loss (3:1 mining) | decode+per-class NMS, batch %d/GPU, %d classes, " \
           "GT %s<=%d" % (args.table.upper(), args.batch, 81, args.gt_mode, args.max_gt)


def base_line(args, value, ms, n_gpus):
    return {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "global_batch": args.batch * n_gpus, "priors": None,
                       "l2": "inputs larger than L2 (logits %.0f MB per GPU)" % 0.0, "parallelism": "dp%d" % n_gpus,
                       "mining": "per-shard" if getattr(args, "mining", "shard") == "shard" else "exact-global"}}


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1, load0=None, load1=None):
        """Median SM clock and throttle reasons over the timed window [t0, t1]; when that window is too
        short for three samples, over the whole period the GPU was under this benchmark's load."""
        out = self._summary(t0 - 0.02, t1 + 0.02)
        out["window"] = "timed region"
        if out["samples"] < 3 and load0 is not None:
            out = self._summary(load0, load1)
            out["window"] = "whole loaded period (timed region shorter than 3 samples)"
        return out

    def _summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, row in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in row.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(names, f[5:9]):
                if v == "Active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ssdgeom import _native as N, device as D, synth
    from ssdgeom.pipeline import HotPath

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    N.check(N.lib().ssdg_set_device(local_rank), "set_device")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    table = synth.TABLES[args.table]
    b = args.batch
    boxes, cls, off = synth.make_gt(100 + rank, b, args.max_gt, args.gt_mode)
    tviews = {}

    def allreduce_on(buf, stream):   # exchange words of the cross-shard mining: NCCL, ordered on the loss stream
        key = (buf.ptr, stream.handle)
        if key not in tviews:
            tviews[key] = (torch.as_tensor(buf, device="cuda"), torch.cuda.ExternalStream(stream.handle))
        t, ts = tviews[key]
        if world > 1:
            with torch.cuda.stream(ts):
                dist.all_reduce(t)

    hp = HotPath(table, batch=b, max_gt=int(np.diff(off).max()), total_gt=boxes.shape[0], mining=args.mining,
                 global_priors=b * world * synth.num_priors(table) if args.mining == "global" else None,
                 allreduce=allreduce_on if args.mining == "global" else None)
    hp.split = args.pipeline == "split"
    hp.fused = not args.no_fused and hp.split
    a, c = hp.A, hp.classes
    # pinned host copies of one batch (also the source of the resident copy)
    h = {"gt_boxes": D.PinnedArray(boxes.shape, np.float32), "gt_cls": D.PinnedArray(cls.shape, np.float32),
         "gt_off": D.PinnedArray(off.shape, np.int32), "pred_cls": D.PinnedArray((b, a, c), np.float32),
         "pred_box": D.PinnedArray((b, a, 4), np.float32)}
    h["gt_boxes"].array[...] = boxes; h["gt_cls"].array[...] = cls; h["gt_off"].array[...] = off
    chunk = 16
    for i in range(0, b, chunk):               # generate in chunks to bound host memory
        n = min(chunk, b - i)
        pc, pb = synth.make_predictions(1000 * rank + i, n, a, c)
        h["pred_cls"].array[i:i + n] = pc; h["pred_box"].array[i:i + n] = pb
    o = {"result": D.PinnedArray((N.LOSS_RESULT_LEN,), np.float64), "kept": D.PinnedArray((b, c - 1, hp.top_k), np.int32),
         "count": D.PinnedArray((b, c - 1), np.int32)}
    host_in = [h[k].array for k in ("gt_boxes", "gt_cls", "gt_off", "pred_cls", "pred_box")]
    host_out = [o[k].array for k in ("result", "kept", "count")]
    hp.upload(*host_in)
    hp.s_main.sync()

    # multi-GPU exchange step: all-reduce of the separable loss sums (num_pos, num_neg, the three sums)
    s_main_t = torch.cuda.ExternalStream(hp.s_main.handle)
    res_t = torch.as_tensor(hp.loss["result"], device="cuda")

    t_streams = {}

    def exchange_on(stream):     # enqueued by HotPath on its loss stream: hidden under the NMS of the other branch
        if stream.handle not in t_streams:
            t_streams[stream.handle] = torch.cuda.ExternalStream(stream.handle)
        with torch.cuda.stream(t_streams[stream.handle]):
            dist.all_reduce(res_t[4:11])

    if world > 1 and args.mining == "shard":
        hp.loss_exchange = exchange_on

    def exchange():
        pass

    def full_step():
        hp.step()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.2)
    t_load0 = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        full_step()
    hp.s_main.sync()

    # ---- timed region: device-resident chain ------------------------------------------------------------
    ev0, ev1 = D.Event(), D.Event()
    barrier()
    t0 = time.perf_counter()
    ev0.record(hp.s_main)
    for _ in range(args.steps):
        full_step()
    hp.finish_exchange()          # the last step's loss exchange belongs to the timed region
    ev1.record(hp.s_main)
    hp.s_main.sync()
    barrier()
    t1 = time.perf_counter()
    ms_total = ev0.elapsed_ms(ev1)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = b * world * args.steps / (ms_total * 1e-3)
    res = hp.loss["result"].to_host()

    # ---- per-stage serial timing + dominant-kernel durations (CUDA events on the launch stream) ---------
    N.lib().ssdg_profile_enable(1)
    import ctypes as C
    stages = {}
    kernel_ms = {}

    def time_stage(name, fn, prof_ids):
        for _ in range(3):
            fn(hp.s_main)
        hp.s_main.sync()
        tot, ksum = 0.0, {k: 0.0 for k in prof_ids}
        a0, a1 = D.Event(), D.Event()
        for _ in range(args.steps):
            a0.record(hp.s_main)
            fn(hp.s_main)
            a1.record(hp.s_main)
            tot += a0.elapsed_ms(a1)
            for k in prof_ids:
                ms = C.c_float(0)
                N.lib().ssdg_profile_last_ms(k, C.byref(ms))
                ksum[k] += ms.value
        stages[name] = tot / args.steps
        for k in prof_ids:
            kernel_ms[k] = ksum[k] / args.steps

    time_stage("assign_ms", hp.assign, [N.PROF_MATCH])
    time_stage("loss_ms", hp.loss_stage, [N.PROF_CE])
    time_stage("detect_ms", hp.detect_stage, [N.PROF_FILTER, N.PROF_NMS])
    ce_alone_ms, filter_alone_ms = kernel_ms.get(N.PROF_CE, 0.0), kernel_ms.get(N.PROF_FILTER, 0.0)
    if hp.fused:   # the variants the chained step actually launches
        time_stage("detect_with_row_stats_ms", lambda s: hp.detect_stage(s, stats=True), [N.PROF_FILTER, N.PROF_NMS])
        time_stage("loss_from_row_stats_ms", lambda s: hp.loss_stage(s, stats=True), [N.PROF_CE])
    N.lib().ssdg_profile_enable(0)
    clocks = sampler.summary(t0, t1, t_load0, time.perf_counter())

    # ---- end to end through the host API -------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        pipelined = args.mining == "shard"   # double-buffered inputs: the H2D of step k+1 runs under step k
        def e2e_step():
            if pipelined:
                hp.submit(host_in, host_out)
            else:
                hp.step_host(*host_in, *host_out)
            exchange()
        for _ in range(2):
            e2e_step()
        hp.s_main.sync()
        barrier()
        e0, e1 = D.Event(), D.Event()
        e0.record(hp.s_main)
        if pipelined:
            D.stream_wait_event(hp.s_copy, e0)   # the first timed copy starts inside the timed region
        for _ in range(args.steps):
            e2e_step()
        e1.record(hp.s_main)
        hp.s_main.sync()
        barrier()
        e_ms = max_over_ranks(e0.elapsed_ms(e1))
        e2e = {"value": b * world * args.steps / (e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": hp.h2d_bytes,
               "d2h_bytes_per_step": hp.d2h_bytes, "ms_per_step": e_ms / args.steps,
               "mode": "HotPath.submit: double-buffered inputs, H2D of step k+1 under compute + D2H of step k" if pipelined
               else "HotPath.step_host: serial H2D, compute, D2H"}
    sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (ce_kernel: one pass over the logits) -------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
    def traffic_of(name):   # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full), same shape only
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", name)))
            if cap.get("batch") == b and cap.get("priors") == a and cap.get("classes") == c:
                return cap["dram_bytes_per_launch"]
        except Exception:
            pass
        return None

    if hp.fused:
        # the one pass over the logits: in logits + pred_box, out decoded boxes (16 B) + row statistics (12 B) per prior
        k_name, k_ms = "filter_kernel (softmax filter + decode + loss row statistics)", kernel_ms.get(N.PROF_FILTER, 0.0)
        k_bytes = b * a * (c * 4 + 16 + 16 + 12)
        traffic = traffic_of("filter_kernel_traffic.json")
    else:
        k_name, k_ms = "ce_kernel", ce_alone_ms
        k_bytes = b * a * (c * 4 + 16 + 16 + 4 + 1)            # logits + pred_box + gt_box + gt_cls + mask
        traffic = traffic_of("ce_kernel_traffic.json")
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": k_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": k_ms,
            "algorithmic_bytes_per_launch": k_bytes}
    ce_ms = ce_alone_ms
    filt_bytes = b * a * (c * 4 + 16)
    f_ms = filter_alone_ms
    extra = {"stages_ms": stages, "fused_logits_pass": bool(hp.fused),
             "ce_kernel_gbs_standalone": b * a * (c * 4 + 16 + 16 + 4 + 1) / (ce_ms * 1e-3) / 1e9 if ce_ms > 0 else None,
             "kernels_ms": {"match_kernel": kernel_ms.get(N.PROF_MATCH), "ce_kernel": ce_ms, "filter_kernel": f_ms,
                            "nms_kernel": kernel_ms.get(N.PROF_NMS),
                            "filter_kernel_with_row_stats": kernel_ms.get(N.PROF_FILTER) if hp.fused else None,
                            "lossprep_kernel": kernel_ms.get(N.PROF_CE) if hp.fused else None},
             "filter_kernel_gbs": filt_bytes / (f_ms * 1e-3) / 1e9 if f_ms > 0 else None,
             "match_pairs_per_s": (float(np.diff(off).sum()) * a) / (kernel_ms.get(N.PROF_MATCH, 0) * 1e-3)
             if kernel_ms.get(N.PROF_MATCH) else None,
             "chain_bytes_per_image": a * (c * 4 * (1 if hp.fused else 2) + 16 * 3 + 4 + 1 + 4 + 16 + 1),
             "loss": ({"total": res[0], "num_pos": res[4], "num_neg": res[5], "status": res[7], "scope": "this rank"}
                      if args.mining == "shard" else
                      {"total": (res[8] + res[10]) / res[11] + res[9] / res[5], "num_pos": res[11], "num_neg": res[5],
                       "status": res[7], "scope": "all ranks (exact-global mining)"})}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(args, args.cpu_images)
        arm.step()                                   # warm the pool
        t0c = time.perf_counter()
        n_cpu_steps = 2
        for _ in range(n_cpu_steps):
            arm.step()
        dtc = time.perf_counter() - t0c
        arm.close()
        cpu = cpu_desc(arm, args, arm.images * n_cpu_steps / dtc)

    line = base_line(args, value, ms_step, world)
    line["config"]["priors"] = a
    line["config"]["l2"] = "inputs larger than L2 (logits %.0f MB per GPU)" % (hp.pred_cls.nbytes / 1e6)
    line.update({"clocks": clocks, "e2e": e2e, "gpu_launches": hp.kernel_launches_per_step * args.steps,
                 "roofline": roof, "cpu_baseline": cpu, "detail": extra, "impl": "ours"})
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
