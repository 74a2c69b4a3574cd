#!/usr/bin/env python
"""bench.py -- segmenter images/s at 1024x2048 (BASELINE.json metric) on N B200s of one node.

A "step" segments one batch of B synthetic cfg2 images per GPU (soft maps, C=9, K=10 spiral offsets,
mergenet_b200.synth.cfg_cityscapes) through the whole path: edge pass -> record init + sort -> merge
scheduler -> labels.

  value      whole-job images/s, inputs already resident in HBM when the timed region starts
  e2e        the same through the host-buffer C ABI (mn_segment_batch_host): pinned host inputs,
             H2D + D2H inside the timed region
  roofline   the edge-construction kernel (HBM-bound): algorithmic bytes 4*N*[(C+K)+(C+2K)] per image
             / its CUDA-event duration on the launching stream, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) on the box's host cores, bounded sample
  --impl reference   times that CPU reference as its own arm (rank 0 only)

Multi-GPU: one process per GPU (torchrun), images partitioned by index, no data-path collective;
NCCL only gathers the per-image instance counts (the "result gather").  scaling = weak.
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

H, W, C, K = 1024, 2048, 9, 10
OPTS = (0.0, 1.0, 0.03)  # egs/cityscape/local/segment.py:134-136
CROP = (256, 512)        # CPU-baseline sample: a crop of the same workload


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_images(n, seed0, h=H, w=W):
    from mergenet_b200 import synth
    cps, sps = [], []
    for i in range(n):
        cp, sp, offs, _ = synth.cfg_cityscapes(h, w, seed=seed0 + i, n_shapes=max(4, int(400 * h * w / (1024 * 2048))),
                                               rmax=max(8, int(120 * h / 1024)), soft=True, noise_seed=7 + seed0 + i)
        cps.append(cp); sps.append(sp)
    return np.ascontiguousarray(np.stack(cps)), np.ascontiguousarray(np.stack(sps)), offs


def parity_with_reference_fixture(mask, classes, h, w):
    """SURVEY 8d "parity check attached to every timing": image 0 of rank 0 (seed 1000, noise seed 1007) against
    tests/golden/full/bench_image0_seed1000_noise1007.npz = canonical mask + classes computed by the UNMODIFIED
    reference in the build container (tests/golden/make_golden.py).  Relabel = first appearance in raster
    order, done here with numpy (nothing under oracle/ is used).  None when no fixture applies."""
    path = os.path.join(ROOT, "tests", "golden", "full", "bench_image0_seed1000_noise1007.npz")
    if (h, w, C, K) != (1024, 2048, 9, 10) or not os.path.exists(path):
        return None
    g = np.load(path)
    flat = np.asarray(mask).ravel()
    labels, first = np.unique(flat, return_index=True)
    perm = np.zeros(int(labels.max()) + 1, dtype=np.int64)
    nxt = 1
    for idx in np.argsort(first, kind="stable"):
        if labels[idx] != 0:
            perm[labels[idx]] = nxt
            nxt += 1
    cm = perm[flat].reshape(h, w)
    cc = np.zeros(nxt - 1, dtype=np.int64)
    for k, c in enumerate(classes, start=1):
        if k < perm.size and perm[k] > 0:
            cc[perm[k] - 1] = c
    return {"image": "rank 0, image 0", "fixture": "tests/golden/full/bench_image0_seed1000_noise1007.npz (unmodified reference)",
            "instances": int(nxt - 1), "reference_instances": int(len(g["cls"])),
            "mask_equal": bool(np.array_equal(cm, g["mask"])), "classes_equal": bool(list(cc) == list(g["cls"]))}


# ---- CPU reference arm / baseline -----------------------------------------------------------------
def _ref_worker(seed):
    import oracle
    from mergenet_b200 import synth
    cp, sp, offs, _ = synth.cfg_cityscapes(CROP[0], CROP[1], seed=seed, n_shapes=25, rmax=30, soft=True, noise_seed=7 + seed)
    t = time.time()
    if oracle.have_reference():
        mask, oc = oracle.ref_run_segmentation(cp, sp, C, offs, *OPTS)
        kind = "reference"
    else:
        mask, oc, _ = oracle.oracle_run_segmentation(cp, sp, C, offs, *OPTS)
        kind = "port"
    return time.time() - t, kind, len(oc)


def cpu_reference_sample(seed0=9000):
    """One wave of P = min(cores, 64) processes, each segmenting its own 256x512 crop-sized cfg2 image
    with the reference's C++ (oracle/_ref).  Returns full-resolution image equivalents per second."""
    import multiprocessing as mp
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cores = max(1, min(avail, 64))
    t = time.time()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_ref_worker, [seed0 + i for i in range(cores)])
    wall = time.time() - t
    frac = (CROP[0] * CROP[1]) / float(H * W)
    per_img = [r[0] for r in res]
    value = cores * frac / max(per_img)  # a wave finishes with its slowest member
    return {"value": value, "unit": "images/s", "cores": cores, "kind": res[0][1],
            "sample": "one wave of %d procs x one %dx%d cfg2 image each (%.1f s/image median, wall %.1f s); "
                      "scaled by pixel count to 1024x2048 equivalents (the reference is superlinear in size, "
                      "so this flatters it)" % (cores, CROP[0], CROP[1], float(np.median(per_img)), wall)}


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import oracle
    oracle.build()
    vals = []
    for i in range(args.warmup):
        cpu_reference_sample(9000 + 100 * i)
    t0 = time.time()
    last = None
    for i in range(args.steps):
        last = cpu_reference_sample(9500 + 100 * i)
        vals.append(last["value"])
    dt = time.time() - t0
    v = float(np.mean(vals))
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": "segmenter images/sec @1024x2048", "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 1024x2048 C=9 K=10 soft maps (CPU sample: %dx%d images, pixel-scaled)" % CROP,
                   "opts": list(OPTS)},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---- own arm ------------------------------------------------------------------------------------
def run_own_arm(args):
    import torch
    import torch.distributed as dist
    from mergenet_b200 import BatchSegmenter, SegmenterOptions, _lib
    if _lib.needs_build():
        _lib.build()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    distributed = world > 1
    torch.cuda.set_device(local)
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    h, w = args.height, args.width
    B = args.batch
    if B <= 0:
        # one persistent CTA per image: the more images in flight the better, up to one per SM; a
        # 1024x2048 image needs 1.36 GB of workspace + its maps + its outputs (device and staging copies)
        free_b, _ = torch.cuda.mem_get_info(local)
        per_img = (BatchSegmenter.workspace_bytes_per_image(h, w, C, K) + 4 * h * w * (C + K) + 8 * h * w)
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        B = int(max(1, min(sms, (free_b - (5 << 30)) // per_img)))
        B = (B // 8) * 8 if B >= 16 else B
        if distributed:  # every rank runs the same batch
            t = torch.tensor([B], dtype=torch.int64, device=torch.device("cuda", local))
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            B = int(t.item())
    # synthetic inputs: `distinct` different images per rank, tiled to the batch (generation is slow)
    distinct = min(B, args.distinct)
    cp, sp, offs = make_images(distinct, 1000 + rank * B, h, w)
    reps = (B + distinct - 1) // distinct
    cp = np.ascontiguousarray(np.concatenate([cp] * reps)[:B])
    sp = np.ascontiguousarray(np.concatenate([sp] * reps)[:B])
    opts = SegmenterOptions(*OPTS)
    seg = BatchSegmenter(B, h, w, C, offs, device=local)
    d_cp = torch.from_numpy(cp).to(dev)
    d_sp = torch.from_numpy(sp).to(dev)
    h_cp = torch.from_numpy(cp).pin_memory()
    h_sp = torch.from_numpy(sp).pin_memory()
    out_dev = (torch.empty((B, h, w), dtype=torch.int32, device=dev), torch.empty((B, h * w), dtype=torch.int32, device=dev),
               torch.empty((B,), dtype=torch.int32, device=dev))
    out_host = (np.empty((B, h, w), np.int32), np.empty((B, h * w), np.int32), np.empty((B,), np.int32))
    gathered = [torch.empty((B,), dtype=torch.int32, device=dev) for _ in range(world)] if distributed else None

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        seg.segment_device(d_cp, d_sp, opts, clip=False, out=out_dev)
        if distributed:  # the result gather (per-image instance counts) is the only collective
            dist.all_gather(gathered, out_dev[2])

    def step_host():
        seg.segment_host(h_cp.numpy(), h_sp.numpy(), opts, clip=False, out=out_host)

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        ev0.record()
        edge_ms, merge_ms, launches = [], [], 0
        for _ in range(steps):
            fn()
            tm = seg.timings()
            edge_ms.append(tm["edge_ms"]); merge_ms.append(tm["merge_ms"])
            launches += tm["edge_launches"] + tm["other_launches"]
        ev1.record()
        barrier()
        dev_ms = ev0.elapsed_time(ev1)
        wall_ms = 1e3 * (time.time() - t0)
        # the library synchronises its own (non-default) stream inside every call, so the device work of
        # a step lies inside [t0, t1]; take the larger of the event span and the host span.
        ms = max(dev_ms, wall_ms)
        if distributed:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, edge_ms, merge_ms, launches

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, edge_ms, merge_ms, launches = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    stats = [seg.stats(b) for b in range(B)]
    tm_dev = seg.timings()
    logprob0 = seg.total_logprob(0)[3]
    parity = None
    if rank == 0:
        n0 = int(out_dev[2][0].item())
        parity = parity_with_reference_fixture(out_dev[0][0].cpu().numpy(), out_dev[1][0, :n0].cpu().numpy().tolist(), h, w)
    value = world * B * args.steps / (ms / 1e3)

    # e2e through the host-buffer ABI (one warm-up, then the same number of steps).  The device-resident
    # copies of the inputs are released first: the host path stages its own.
    del d_cp, d_sp
    torch.cuda.empty_cache()
    step_host()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ms_h, _, _, _ = timed(step_host, e2e_steps)
    e2e_value = world * B * e2e_steps / (ms_h / 1e3)
    tm_h = seg.timings()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        n = h * w
        edge_bytes = 4.0 * n * ((C + K) + (C + 2 * K)) * B
        edge_s = float(np.mean(edge_ms)) / 1e3
        achieved = edge_bytes / edge_s / 1e9
        cpu = None
        if not args.no_cpu_baseline:
            import oracle
            oracle.build()
            cpu = cpu_reference_sample()
        rounds = [s["rounds"] for s in stats]
        events = [s["events"] for s in stats]
        merges = [s["merges"] for s in stats]
        merge_s = float(np.mean(merge_ms)) / 1e3
        line = {
            "metric": "segmenter images/sec @1024x2048" if (h, w) == (H, W) else "segmenter images/sec @%dx%d" % (h, w),
            "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: %dx%d, C=9, K=10 spiral offsets generate_offsets(40,10), soft maps "
                                   "sigmoid(3(2t-1)+N(0,1)); batch %d images per GPU (%d distinct), opts (sdb,omf,mlb)=%s; "
                                   "working set per step %.1f GB >> 126 MB L2 (no L2 flush needed)"
                                   % (h, w, B, distinct, list(OPTS), (cp.nbytes + sp.nbytes + seg.workspace_bytes_per_image(h, w, C, K) * B) / 1e9),
                       "batch_per_gpu": B, "partition": "images by index, contiguous per rank"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(cp.nbytes + sp.nbytes),
                    "d2h_bytes_per_step": int(out_host[0].nbytes + out_host[1].nbytes + out_host[2].nbytes),
                    "steps": e2e_steps, "h2d_ms": tm_h["h2d_ms"], "d2h_ms": tm_h["d2h_ms"]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "mn_edge_warp_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this
                         # kernel (profiles/r01_edge_warp_ncu_full.json: 408.0 MB per 1024x2048 image;
                         # algorithmic 402.7 MB + the 8.4 MB class-index plane the pass also writes)
                         "traffic": (408.0e6 * B) if (h, w, C, K) == (1024, 2048, 9, 10) else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": edge_bytes, "avg_launch_ms": edge_s * 1e3},
            "cpu_baseline": cpu,
            "parity": parity,
            "scheduler": {"merge_kernel_ms": merge_s * 1e3, "rounds_per_image": float(np.mean(rounds)),
                          "events_per_image": float(np.mean(events)), "merges_per_image": float(np.mean(merges)),
                          "us_per_round": 1e6 * merge_s / max(1.0, float(np.max(rounds))),
                          "merges_per_s_per_image": float(np.mean(merges)) / merge_s,
                          "merges_per_s_batch": float(np.sum(merges)) / merge_s},
            "phases_ms": {k: tm_dev[k] for k in ("edge_ms", "record_init_sort_ms", "merge_ms", "aggregate_ms", "label_ms")},
            # second HBM-bound pass: total log-prob terms from the maintained sums (segment.cc:272-287);
            # algorithmic bytes per image: obj 16 N + parent 4 N + record key/differentness halves 16 E
            "aggregation": {"kernel": "mn_logprob_kernel", "bound": "hbm", "unit": "GB/s",
                            "algorithmic_bytes_per_launch": float(B * (20 * n + 16 * n * K)),
                            "avg_launch_ms": tm_dev["aggregate_ms"],
                            "achieved": B * (20 * n + 16 * n * K) / max(1e-9, tm_dev["aggregate_ms"] * 1e-3) / 1e9,
                            "frac": B * (20 * n + 16 * n * K) / max(1e-9, tm_dev["aggregate_ms"] * 1e-3) / 1e9 / peak,
                            "traffic": (679.6e6 * B) if (h, w, C, K) == (1024, 2048, 9, 10) else None,
                            "total_logprob_image0": logprob0},
        }
        print(json.dumps(line))
    seg.close()
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batch", type=int, default=env_int("MN_BENCH_BATCH", 0),
                    help="images per GPU per step (0 = as many as fit the free HBM, at most one per SM)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic images per rank (tiled to the batch)")
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
